"""GPU-box diagnostic for test_bf16_mask_criteria_after_prefit[l]: where does the bf16 EVAL path leave the fp32 path?
Phase "fit": 40 fp32 steps on Hiera-L like the test, weights to /tmp.  Phase "eval" (one subprocess per switch setting,
the csrc switches are read once per process): bf16 forward vs the fp32 CUDA forward of the same weights."""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SD = "/tmp/prefit_l.pt"
CFG = "sam2_hiera_l.yaml"


def fit():
    from oracle import port
    from sam2_unet_b200 import SAM2UNet, TrainStep
    from sam2_unet_b200.params import fill_deterministic_
    dev = torch.device("cuda", 0)
    x, mask = port.synthetic_batch(8, 352, seed=11, correlated=True)
    m32 = SAM2UNet(model_cfg=CFG, dtype="fp32")
    fill_deterministic_(m32, 0)
    m32 = m32.to(dev)
    step = TrainStep(m32, lr=1e-3, weight_decay=5e-4, use_graph=False)
    for i in range(40):
        loss = step(x.to(dev), mask.to(dev))
    print("fit loss", loss.sum().item(), flush=True)
    sd = {k: v.detach().cpu().clone() for k, v in m32.state_dict().items()}
    # reference outputs + tape of the fp32 CUDA path, eval and train mode
    ref = {}
    eng = m32._engine(dev)
    for train in (False, True):
        with torch.no_grad():
            o = eng.forward(x.to(dev), train, save=True)
        tape = eng.tape
        ref[train] = {"outs": [t.float().cpu() for t in o], "blocks": [b["y"].float().cpu() for b in tape["blocks"]],
                      "dec": {k + n: v[n].float().cpu() for k, v in tape["dec"].items() for n in ("dst", "mid", "out") if n in v}}
        eng.tape = None
        m32.load_state_dict(sd, strict=True)
        eng = m32._engine(dev)
    torch.save({"sd": sd, "ref": ref, "x": x}, SD)


def evaluate(tag):
    from sam2_unet_b200 import SAM2UNet
    dev = torch.device("cuda", 0)
    blob = torch.load(SD)
    sd, ref, x = blob["sd"], blob["ref"], blob["x"]
    mb = SAM2UNet(model_cfg=CFG, dtype="bf16").to(dev)
    mb.load_state_dict(sd, strict=True)
    for train in (False, True):
        eng = mb._engine(dev)
        with torch.no_grad():
            o = eng.forward(x.to(dev), train, save=True)
        tape = eng.tape
        eng.tape = None
        line = []
        for g, r, name in zip(o, ref[train]["outs"], ("out", "out1", "out2")):
            sg, sr = torch.sigmoid(g.float().cpu()), torch.sigmoid(r)
            pg, pr = sg > 0.5, sr > 0.5
            iou = ((pg & pr).sum().item() + 1e-9) / ((pg | pr).sum().item() + 1e-9)
            line.append(f"{name} {float((sg - sr).abs().max()):.4f}/{iou:.4f}")
        print(f"[{tag}] train={train}: " + "  ".join(line), flush=True)
        if tag == "default":
            errs = []
            for i, (a, b) in enumerate(zip(ref[train]["blocks"], tape["blocks"])):
                b = b["y"].float().cpu()
                errs.append(f"{i}:{float((a - b).norm() / a.norm()):.4f}")
            print("   block rel-L2:", " ".join(errs[::4] + errs[-1:]), flush=True)
            mine = {k + n: v[n].float().cpu() for k, v in tape["dec"].items() for n in ("dst", "mid", "out") if n in v}
            for k in sorted(ref[train]["dec"]):
                a, b = ref[train]["dec"][k], mine[k]
                if a.shape == b.shape:
                    print(f"   dec {k}: rel-L2 {float((a - b).norm() / a.norm()):.4f}  max {float((a - b).abs().max() / a.abs().max()):.4f}", flush=True)
        mb.load_state_dict(sd, strict=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "eval":
        evaluate(sys.argv[2])
        sys.exit(0)
    fit()
    variants = [("default", {}), ("conv_igemm=0", {"S2U_CONV_IGEMM": "0"}), ("conv_tall=0", {"S2U_CONV_TALL": "0"}),
                ("merge_1x1=0", {"S2U_MERGE_1X1": "0"}), ("fuse_adapter=0", {"S2U_FUSE_ADAPTER": "0"}),
                ("attn_backend=1", {"S2U_ATTN_BACKEND": "1"})]
    for tag, env in variants:
        e = dict(os.environ)
        e.update(env)
        subprocess.run([sys.executable, os.path.abspath(__file__), "eval", tag], env=e, check=False)

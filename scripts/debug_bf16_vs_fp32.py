"""GPU-box diagnostic: per-block divergence between the bf16 and the fp32 engine on the same weights."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import port  # noqa: E402
from sam2_unet_b200 import SAM2UNet  # noqa: E402
from sam2_unet_b200.params import fill_deterministic_  # noqa: E402

dev = torch.device("cuda:0")


def run(cfg, S, overlap=True):
    x, _ = port.synthetic_batch(1, S, seed=5)
    outs = {}
    for dt in ("fp32", "bf16"):
        m = SAM2UNet(model_cfg=cfg, dtype=dt)
        fill_deterministic_(m, 0)
        m = m.to(dev).train()           # train mode + save=True keeps the per-block tape
        eng = m._engine(dev)
        eng.overlap = overlap
        with torch.no_grad():
            o = eng.forward(x.to(dev), False, save=True)
        tape = eng.tape
        outs[dt] = ([b["y"].float().clone() for b in tape["blocks"]], [v["dst"].float().clone() for k, v in sorted(tape["dec"].items()) if k.startswith("rfb")], [t.clone() for t in o])
        eng.tape = None
        del m, eng
    print(f"== {cfg} S={S} overlap={overlap}")
    for i, (a, b) in enumerate(zip(outs["fp32"][0], outs["bf16"][0])):
        err = ((a - b).abs().max() / a.abs().max()).item()
        flag = "  <<<<" if err > 0.05 else ""
        print(f"   block {i:2d} y maxnorm diff {err:.4f}{flag}")
    for i, (a, b) in enumerate(zip(outs["fp32"][1], outs["bf16"][1])):
        n = min(a.shape[1], 64)
        print(f"   rfb{i + 1} out diff {((a[:, :n] - b[:, :n]).abs().max() / a[:, :n].abs().max()).item():.4f}")
    for i, (a, b) in enumerate(zip(outs["fp32"][2], outs["bf16"][2])):
        print(f"   logits {i} diff {((a - b).abs().max() / a.abs().max()).item():.4f}  max|logit| {a.abs().max().item():.2f}")


run("sam2_hiera_b+.yaml", 352)
run("sam2_hiera_b+.yaml", 352, overlap=False)
run("sam2_hiera_l.yaml", 1024)

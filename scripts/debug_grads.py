"""Diagnostic (GPU box): per-group gradient error of the CUDA path vs the oracle port on the tiny trunk."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import port  # noqa: E402
from sam2_unet_b200 import SAM2UNet, structure_loss  # noqa: E402
from sam2_unet_b200.params import fill_deterministic_  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
cuda = torch.device("cuda:0")


def group(name):
    if name.startswith("encoder.blocks."):
        return "adapter." + name.split(".")[2]
    return name.split(".")[0] if not name.startswith("rfb") else ".".join(name.split(".")[:2])


def run(cfg, key, B, S, seed, dtype="fp32"):
    m = SAM2UNet(model_cfg=cfg, dtype=dtype)
    fill_deterministic_(m, 0)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.to(cuda).train()
    x, mask = port.synthetic_batch(B, S, seed=seed)
    bn = port.BNState()
    loss_ref, outs_ref, grads_ref = port.loss_and_grads(sd, port.TRUNKS[key], x, mask, True, bn)
    outs = m(x.to(cuda))
    loss = sum(structure_loss(o, mask.to(cuda)) for o in outs)
    loss.backward()
    print(f"== {cfg} B={B} S={S} {dtype}: loss {loss.item():.7f} ref {loss_ref.item():.7f}")
    for o, r in zip(outs, outs_ref):
        print("   logits maxnorm err", ((o.detach().cpu() - r).abs().max() / r.abs().max()).item())
    params = dict(m.named_parameters())
    groups = {}
    num = den = 0.0
    for k, g in grads_ref.items():
        if g is None:
            continue
        d = (params[k].grad.detach().cpu() - g).double()
        e = groups.setdefault(group(k), [0.0, 0.0])
        e[0] += float((d * d).sum())
        e[1] += float((g.double() ** 2).sum())
        num += float((d * d).sum())
        den += float((g.double() ** 2).sum())
    print("   global rel-L2", (num / den) ** 0.5)
    for gname, (n, d) in groups.items():
        print(f"   {gname:24s} rel-L2 {(n / max(d, 1e-30)) ** 0.5:.3e}   |g| {d ** 0.5:.3e}   share of err {n / num:.3f}")


if __name__ == "__main__":
    run("tiny_test.yaml", "test", 2, 160, 2)
    run("tiny_test.yaml", "test", 4, 224, 3)
    run("tiny_test.yaml", "test", 2, 160, 2, "bf16")

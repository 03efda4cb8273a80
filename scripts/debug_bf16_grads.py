"""GPU-box diagnostic: bf16 engine against the fp32 engine on the same Hiera-L weights and batch - per-block forward
divergence (rel-L2 of the residual stream) and per-group gradient divergence, for both attention kernel families."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sam2_unet_b200 import SAM2UNet, _lib, structure_loss  # noqa: E402
from sam2_unet_b200.params import fill_deterministic_  # noqa: E402
from sam2_unet_b200.synthetic import synthetic_batch  # noqa: E402

dev = torch.device("cuda:0")
B = int(os.environ.get("B", "4"))
x, mask = synthetic_batch(B, 352, seed=3)
x, mask = x.to(dev), mask.to(dev)


def run(dt, backend):
    _lib.call("s2u_set_attn_backend", backend)
    m = SAM2UNet(model_cfg="sam2_hiera_l.yaml", dtype=dt)
    fill_deterministic_(m, 0)
    m = m.to(dev).train()
    outs = m(x)
    eng = m._eng
    ys = [b["y"].float().clone() for b in eng.tape["blocks"]]
    loss = sum(structure_loss(o, mask) for o in outs)
    loss.backward()
    grads = {k: p.grad.detach().float().clone() for k, p in m.named_parameters() if p.grad is not None}
    _lib.call("s2u_set_attn_backend", 0)
    return ys, [o.detach().clone() for o in outs], grads, loss.item()


ref = run("fp32", 1)
for backend, name in ((1, "mma.sync"), (0, "tcgen05")):
    got = run("bf16", backend)
    print(f"== bf16 [{name}] loss {got[3]:.6f} vs fp32 {ref[3]:.6f}")
    line = []
    for i, (a, b) in enumerate(zip(ref[0], got[0])):
        line.append(f"{i}:{((a - b).norm() / a.norm()).item():.4f}")
    print("   block y rel-L2:", " ".join(line))
    for i, (a, b) in enumerate(zip(ref[1], got[1])):
        print(f"   logits {i} rel-L2 {((a - b).norm() / a.norm()).item():.4f}")
    groups = {}
    for k, g in ref[2].items():
        if k.startswith("encoder.blocks."):
            i = int(k.split(".")[2])
            key = f"adapters {i // 6 * 6:02d}-{i // 6 * 6 + 5:02d}"
        else:
            key = k.split(".")[0]
        d = (got[2][k] - g).double()
        acc = groups.setdefault(key, [0.0, 0.0])
        acc[0] += float((d * d).sum())
        acc[1] += float((g.double() ** 2).sum())
    tot = [sum(v[0] for v in groups.values()), sum(v[1] for v in groups.values())]
    print("   gradient rel-L2 global", (tot[0] / tot[1]) ** 0.5)
    for k in sorted(groups):
        print(f"      {k:18s} rel-L2 {(groups[k][0] / groups[k][1]) ** 0.5:.4f}   |g| {groups[k][1] ** 0.5:.3e}")

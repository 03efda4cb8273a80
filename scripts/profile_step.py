"""GPU-box diagnostic: CUDA-event breakdown of one eager Hiera-L train step, GEMMs grouped by shape."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sam2_unet_b200.synthetic import synthetic_batch  # noqa: E402
from sam2_unet_b200 import SAM2UNet, TrainStep, _lib  # noqa: E402
from sam2_unet_b200.params import fill_deterministic_  # noqa: E402

B = int(os.environ.get("PB", "12"))
cfg = os.environ.get("PCFG", "sam2_hiera_l.yaml")
dev = torch.device("cuda:0")
m = SAM2UNet(model_cfg=cfg, dtype="bf16")
fill_deterministic_(m, 0)
m = m.to(dev).train()
x, mask = synthetic_batch(B, 352, seed=0)
x, mask = x.to(dev), mask.to(dev)
step = TrainStep(m, use_graph=False)
for _ in range(2):
    step(x, mask)
torch.cuda.synchronize()
_lib.profile_begin()
step(x, mask)
rec = _lib.profile_end()
by, shapes = {}, {}
for name, t, a in rec:
    c = by.setdefault(name, [0, 0.0])
    c[0] += 1
    c[1] += t
    if name == "s2u_gemm":
        key = (a[6], a[7], a[8], a[16], "pre" if a[10] else "")
        s = shapes.setdefault(key, [0, 0.0])
        s[0] += 1
        s[1] += t
    if name == "s2u_gemm_wgrad":
        key = ("wgrad", a[6], a[7], a[8])
        s = shapes.setdefault(key, [0, 0.0])
        s[0] += 1
        s[1] += t
total = sum(v[1] for v in by.values())
print(f"total {total:.2f} ms")
for k, v in sorted(by.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:26s} {v[0]:5d} {v[1]:9.3f} ms")
print("\nGEMM shapes (M, N, K, flags, pre_out): calls, total ms, us/call, TFLOP/s, GB/s(min traffic)")
for k, v in sorted(shapes.items(), key=lambda kv: -kv[1][1]):
    if k[0] == "wgrad":
        _, M, P, Q = k
        fl = 2.0 * M * P * Q
        by_ = 2.0 * M * (P + Q)
    else:
        M, N, K = k[:3]
        fl = 2.0 * M * N * K
        by_ = 2.0 * (M * K + N * K + M * N * (2 if k[4] else 1) + (M * N if k[3] & 6 else 0))
    us = v[1] / v[0] * 1e3
    print(f"{str(k):44s} {v[0]:4d} {v[1]:8.3f} {us:9.1f} {fl / (us * 1e-6) / 1e12:8.1f} {by_ / (us * 1e-6) / 1e9:8.0f}")
json.dump({"total_ms": total, "by_op": {k: {"calls": v[0], "ms": v[1]} for k, v in by.items()},
           "gemm_shapes": {str(k): {"calls": v[0], "ms": v[1]} for k, v in shapes.items()}},
          open(os.environ.get("POUT", "gpurun_out/profile_step.json"), "w"), indent=1)

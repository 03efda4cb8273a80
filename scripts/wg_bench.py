"""Weight-gradient kernels at their dominant shapes (CUDA events, 50 back-to-back launches): S2U_WG_CTAS_PER_SM sweep."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sam2_unet_b200.engine import Ops

dev = torch.device("cuda:0")
ops = Ops(torch.bfloat16, dev, 0)
bf = torch.bfloat16


def t(fn, n=50):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for B, H, cin, kh, kw, dil in ((12, 88, 64, 3, 3, 1), (12, 88, 256, 3, 3, 1), (12, 44, 64, 3, 3, 3), (12, 88, 64, 1, 7, 1), (12, 22, 64, 3, 3, 1)):
    M = B * H * H
    x, dy = torch.randn(M, cin, device=dev).to(bf), torch.randn(M, 64, device=dev).to(bf)
    G = torch.zeros(64, cin, kh, kw, device=dev)
    print(f"conv_wgrad {B}x{H}x{H} cin {cin} {kh}x{kw} d{dil}: {t(lambda: ops.conv_wgrad(dy, 64, x, cin, G, B, H, H, cin, 64, kh, kw, dil)):.1f} us")
for R, C in ((5808, 576), (23232, 288), (92928, 144), (1452, 1152)):
    dh2, u = torch.randn(R, C, device=dev).to(bf), torch.randn(R, 32, device=dev).to(bf)
    dh1, x = torch.randn(R, 32, device=dev).to(bf), torch.randn(R, C, device=dev).to(bf)
    G2, G1 = torch.zeros(C, 32, device=dev), torch.zeros(32, C, device=dev)
    print(f"wgrad_pair R {R} C {C}: {t(lambda: ops.wgrad_pair(dh2, u, G2, 32, dh1, x, G1, C)):.1f} us")
for B, H, cin, n, kh, kw, dil in ((12, 88, 64, 64, 3, 3, 1), (12, 88, 64, 64, 3, 3, 7), (12, 88, 64, 64, 1, 7, 1), (12, 88, 256, 64, 3, 3, 1),
                                  (12, 88, 64, 256, 3, 3, 1), (12, 44, 64, 64, 3, 3, 1), (12, 22, 64, 64, 3, 3, 1)):
    M = B * H * H
    x = torch.randn(M, cin, device=dev).to(bf)
    w = (torch.randn(n, kh * kw * cin, device=dev) * 0.05).to(bf)
    y = torch.empty(M, n, device=dev, dtype=bf)
    us = t(lambda: ops.conv_igemm(x, cin, B, H, H, cin, w, n, kh, kw, dil, y, n))
    print(f"conv_igemm {B}x{H}x{H} {cin}->{n} {kh}x{kw} d{dil}: {us:.1f} us  {2.0 * M * n * kh * kw * cin / us / 1e6:.0f} TFLOP/s")

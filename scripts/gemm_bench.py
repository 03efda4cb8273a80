"""GPU-box micro-benchmark of the GEMM kernel on the model's dominant shapes (CUDA events, back-to-back launches)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sam2_unet_b200.engine import Ops  # noqa: E402

dev = torch.device("cuda:0")
ops = Ops(torch.bfloat16, dev)
shapes = [(5808, 2304, 576), (5808, 576, 2304), (5808, 1728, 576), (5808, 576, 1728), (5808, 576, 576),
          (92928, 576, 144), (92928, 144, 576), (92928, 144, 144), (23232, 1152, 288), (23232, 288, 1152),
          (23232, 288, 288), (1452, 4608, 1152), (1452, 1152, 4608), (92928, 64, 576), (5808, 32, 576),
          (5808, 576, 32), (8192, 8192, 8192)]
variants = [("pair128", 1024 + 128), ("pair144", 1024 + 144), ("pair192", 1024 + 192), ("pair256", 1024 + 256), ("auto", 2)]
flagsets = [("plain", 0, False), ("gelu+pre", 1, True), ("gelu+dg", 1 | 128, True), ("dgelu", 2, False),
            ("mulaux", 256, False), ("resid", 4, False), ("stream", 4 | 16 | 32 | 64, True)]
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def timeit(fn, iters=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for (M, N, K) in shapes:
    A = torch.randn(M, K, device=dev).bfloat16()
    W = (torch.randn(N, K, device=dev) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device=dev)
    aux = torch.randn(M, N, device=dev).bfloat16()
    C = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    C32 = torch.empty(M, N, device=dev)
    R32 = torch.randn(M, N, device=dev)
    fl = 2.0 * M * N * K
    t_ref = timeit(lambda: torch.matmul(A, W.t(), out=C))
    line = f"{str((M, N, K)):22s} cublas {t_ref:7.1f}us {fl / t_ref / 1e6:6.0f}TF |"
    for fname, fl_, use_pre in flagsets:
        if fname not in ("plain", "stream") and (M, N, K) not in ((5808, 2304, 576),):
            continue
        if fname == "stream" and N > 1152:
            continue
        for vname, be in variants:
            if N <= 64 and vname != "auto":
                continue
            if fname == "stream" and vname.startswith("old"):
                continue
            try:
                t = timeit(lambda: ops.gemm(A, W, C32 if fl_ & 16 else C, bias=bias, pre_out=pre if use_pre else None,
                                            aux=aux if fl_ & (2 | 256) else None,
                                            resid=(R32 if fl_ & 32 else aux) if fl_ & 4 else None, flags=fl_,
                                            backend=be))
                line += f" {fname}/{vname} {t:6.1f}us {fl / t / 1e6:5.0f}TF |"
            except Exception as e:  # noqa: BLE001
                line += f" {fname}/{vname} ERR |"
    print(line, flush=True)

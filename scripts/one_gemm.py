"""GPU-box helper: a few launches of the dominant GEMM (stage-3 fc1, GELU + saved GELU' epilogue) for ncu captures."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sam2_unet_b200.engine import Ops  # noqa: E402

dev = torch.device("cuda:0")
ops = Ops(torch.bfloat16, dev)
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 129
M, N, K = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (5808, 2304, 576)
A = torch.randn(M, K, device=dev).bfloat16()
W = (torch.randn(N, K, device=dev) * K ** -0.5).bfloat16()
bias = torch.randn(N, device=dev)
C = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
aux = torch.randn(M, N, device=dev).bfloat16()
for _ in range(5):
    ops.gemm(A, W, C, bias=bias, pre_out=pre if flags & 1 else None, aux=aux if flags & (2 | 256) else None, flags=flags)
torch.cuda.synchronize()
print("done")

"""One forward + backward of the stage-3 windowed attention (Hiera-L 352^2, batch 12) for ncu captures."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sam2_unet_b200 import _lib
from sam2_unet_b200.engine import Ops

B, H, W, nh, hd, window = 12, 22, 22, 8, 72, int(os.environ.get("WIN", "16"))
C = nh * hd
cuda = torch.device("cuda:0")
ops = Ops(torch.bfloat16, cuda, 0)
g = torch.Generator().manual_seed(0)
qkv = torch.randn(B, H, W, 3 * C, generator=g).to(cuda).bfloat16()
dout = torch.randn(B, H, W, C, generator=g).to(cuda).bfloat16()
bias = torch.randn(3 * C, generator=g).to(cuda)
out, lse = torch.empty(B, H, W, C, device=cuda, dtype=torch.bfloat16), torch.empty(B, H, W, nh, device=cuda)
dqkv = torch.empty_like(qkv)
_lib.call("s2u_set_attn_backend", int(os.environ.get("BACKEND", "2")))
for _ in range(3):
    ops.attn_fwd(qkv, bias, out, lse, B, H, W, nh, hd, window, False)
    ops.attn_bwd(qkv, bias, out, lse, dout, dqkv, B, H, W, nh, hd, window, False)
torch.cuda.synchronize()
print("done", float(out.float().abs().mean()), float(dqkv.float().abs().mean()))

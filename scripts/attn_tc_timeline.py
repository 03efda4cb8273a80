"""Per-CTA phase timeline of the tcgen05 attention forward (library built with -DS2U_ATC_TIMING)."""
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
bias = torch.randn(3 * C, generator=g).to(cuda)
out, lse = torch.empty(B, H, W, C, device=cuda, dtype=torch.bfloat16), torch.empty(B, H, W, nh, device=cuda)
buf = torch.zeros(8 * 8192, dtype=torch.int64, device=cuda)
os.environ["S2U_ATC_TIMING_BUF"] = str(buf.data_ptr())
_lib.call("s2u_set_attn_backend", 2)
BWD = os.environ.get("BWD", "0") == "1"
dout = torch.randn(B, H, W, C, generator=g).to(cuda).bfloat16()
dqkv = torch.empty_like(qkv)
if BWD:
    del os.environ["S2U_ATC_TIMING_BUF"]
    ops.attn_fwd(qkv, bias, out, lse, B, H, W, nh, hd, window, False)
    torch.cuda.synchronize()
    os.environ["S2U_ATC_TIMING_BUF"] = str(buf.data_ptr())
for _ in range(3):
    torch.cuda.synchronize()                      # isolated launches: no programmatic overlap with a predecessor
    if BWD:
        ops.attn_bwd(qkv, bias, out, lse, dout, dqkv, B, H, W, nh, hd, window, False)
    else:
        ops.attn_fwd(qkv, bias, out, lse, B, H, W, nh, hd, window, False)
torch.cuda.synchronize()
t = buf.view(-1, 8).cpu()
t = t[t[:, 0] > 0].double()
t0 = t[:, 0].min()
print("CTAs", t.shape[0], "kernel span us", (t[:, 7].max() - t0).item() / 1e3)
names = ["start->alloc", "pdl_wait", "->S issued(TMA)", "S mma done", "softmax", "PV done", "epilogue"]
if BWD:
    names = ["start -> zero rows done (thread 0: barriers + TMA issue)", "pdl_wait + D / lse", "sync (waits for TMEM alloc)", "TMA wait + S,dP mma", "elementwise", "dV dK dQ mma", "remaining steps + stores"]
if BWD and os.environ.get("PRO", "0") == "1":
    rel = (t - t[:, :1]) / 1e3
    for nm, col in (("zero rows done (thread 128)", 1), ("pdl_wait done", 2), ("D / lse done", 3), ("after sync", 4), ("TMA issued (thread 0)", 5), ("TMEM alloc done (thread 32)", 6), ("end", 7)):
        print(f"{nm:30s} mean {rel[:, col].mean().item():6.2f} us after start, max {rel[:, col].max().item():6.2f}")
    sys.exit(0)
if BWD:
    heavy = t[t[:, 6] > 0]                      # windows with 2 x 2 tile pairs
    light = t[t[:, 6] == 0]
    names = ["prologue", "step 0", "step 1", "store dK dV (block 0)", "step 2", "step 3", "store dK dV (block 1) + dQ"]
    d = heavy[:, 1:] - heavy[:, :-1]
    print("heavy CTAs", heavy.shape[0], "light", light.shape[0], "light life mean", ((light[:, 7] - light[:, 0]).mean() / 1e3).item())
else:
    d = t[:, 1:] - t[:, :-1]
for i, n in enumerate(names):
    print(f"{n:18s} mean {d[:, i].mean().item() / 1e3:6.2f} us  max {d[:, i].max().item() / 1e3:6.2f}")
life = (t[:, 7] - t[:, 0]) / 1e3
print("CTA life mean us", life.mean().item(), "quantiles", [round(float(torch.quantile(life, q)), 2) for q in (0.1, 0.5, 0.75, 0.9, 1.0)])
start = (t[:, 0] - t0) / 1e3
print("CTA start offsets us: quantiles", [round(float(torch.quantile(start, q)), 2) for q in (0.1, 0.5, 0.62, 0.9, 1.0)])

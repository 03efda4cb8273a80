"""One launch of each non-GEMM kernel family at its dominant Hiera-L 352x352 batch-12 shape, for ncu captures:

    python scripts/ncu_targets.py                      # plain run first (must exit 0)
    ncu --set full --clock-control none --import-source on -k regex:"<names>" -o gpurun_out/r2_kernels python scripts/ncu_targets.py

Kernels (in launch order after the warm-up pass): attention forward / fused backward (tcgen05, 16x16 windows of a 22x22
map), global attention forward / dQ / dK-dV (streaming), LayerNorm forward / backward, BatchNorm statistics / apply /
backward, structure_loss forward / backward."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sam2_unet_b200 import _lib
from sam2_unet_b200.engine import Ops

dev = torch.device("cuda:0")
ops = Ops(torch.bfloat16, dev, 0)
bf, f32 = torch.bfloat16, torch.float32
g = torch.Generator().manual_seed(0)


def r(*s, dt=bf):
    return torch.randn(*s, generator=g).to(dev).to(dt)


B, H, nh, hd = 12, 22, 8, 72
C = nh * hd
qkv, dout, bias = r(B, H, H, 3 * C), r(B, H, H, C), r(3 * C, dt=f32)
out, lse = torch.empty(B, H, H, C, device=dev, dtype=bf), torch.empty(B, H, H, nh, device=dev)
dqkv = torch.empty_like(qkv)
R, Cl = 5808, 576
ln = dict(x=r(R, Cl, dt=f32), g=r(Cl, dt=f32), b=r(Cl, dt=f32), y=torch.empty(R, Cl, device=dev, dtype=bf),
          mean=torch.empty(R, device=dev), rstd=torch.empty(R, device=dev), dy=r(R, Cl), dres=r(R, Cl),
          dx=torch.empty(R, Cl, device=dev, dtype=bf))
M, Cb = 92928, 64
bn = dict(x=r(M, Cb), sums=torch.zeros(_lib.load().s2u_bn_ws_doubles(Cb), device=dev, dtype=torch.float64), gamma=r(Cb, dt=f32),
          beta=r(Cb, dt=f32), rm=torch.zeros(Cb, device=dev), rv=torch.ones(Cb, device=dev), nb=torch.zeros(1, device=dev, dtype=torch.int64),
          scale=torch.empty(Cb, device=dev), shift=torch.empty(Cb, device=dev), mean=torch.empty(Cb, device=dev),
          rstd=torch.empty(Cb, device=dev), out=torch.empty(M, Cb, device=dev, dtype=bf), dy=r(M, Cb),
          dgamma=torch.zeros(Cb, device=dev), dbeta=torch.zeros(Cb, device=dev), c1=torch.empty(Cb, device=dev),
          c2=torch.empty(Cb, device=dev), dx=torch.empty(M, Cb, device=dev, dtype=bf))
Bl, Sl = 12, 352
ls = dict(p=[r(Bl, 1, Sl, Sl, dt=f32) for _ in range(3)], m=(torch.rand(Bl, 1, Sl, Sl, generator=g) > 0.7).float().to(dev),
          w=torch.empty(Bl, Sl, Sl, device=dev), sums=torch.zeros(3 * Bl * 2 + 3, device=dev, dtype=torch.float64),
          loss=torch.empty(3, device=dev), g=[torch.empty(Bl, 1, Sl, Sl, device=dev) for _ in range(3)])


ad = dict(xs=r(R, Cl, dt=f32), W1=r(32, Cl) * Cl ** -0.5, W2=r(Cl, 32) * 32 ** -0.5, b1=r(32, dt=f32), b2=r(Cl, dt=f32),
          xa=torch.empty(R, Cl, device=dev), n1=torch.empty(R, Cl, device=dev, dtype=bf), u=torch.empty(R, 32, device=dev, dtype=bf),
          g1=torch.empty(R, 32, device=dev, dtype=bf), g2=torch.empty(R, Cl, device=dev, dtype=bf),
          dh2=torch.empty(R, Cl, device=dev, dtype=bf), dh1=torch.empty(R, 32, device=dev, dtype=bf),
          dx=torch.empty(R, Cl, device=dev, dtype=bf), db1=torch.zeros(32, device=dev), db2=torch.zeros(Cl, device=dev))
ad["W2t"], ad["W1t"] = ad["W2"].t().contiguous(), ad["W1"].t().contiguous()


# implicit-GEMM convolution (3x3, 64 -> 64, dilation 3, 12 x 88 x 88) + its weight gradient, one-pass BatchNorm backward
Bc, Hc = 12, 88
cv = dict(x=r(Bc * Hc * Hc, 64), w=r(64, 64, 3, 3, dt=f32) * 0.04, wf=torch.empty(64, 576, device=dev, dtype=bf),
          wd=torch.empty(64, 576, device=dev, dtype=bf), y=torch.empty(Bc * Hc * Hc, 64, device=dev, dtype=bf),
          dy=r(Bc * Hc * Hc, 64), dx=torch.empty(Bc * Hc * Hc, 64, device=dev, dtype=bf), G=torch.zeros(64, 64, 3, 3, device=dev))
ops.conv_weight_pack(cv["w"], cv["wf"], cv["wd"], 64, 64, 3, 3)
# stage-1 windows (8 x 8 tokens, 2 heads) and the pooled transition block (4 heads): whole-window kernels
q1, q2 = r(12, 88, 88, 3 * 144), r(12, 88, 88, 3 * 288)
o1, o2 = torch.empty(12, 88, 88, 144, device=dev, dtype=bf), torch.empty(12, 44, 44, 288, device=dev, dtype=bf)
l1, l2 = torch.empty(12, 88, 88, 2, device=dev), torch.empty(12, 44, 44, 4, device=dev)
do1, do2 = r(12, 88, 88, 144), r(12, 44, 44, 288)
dq1, dq2 = torch.empty_like(q1), torch.empty_like(q2)
b1a, b2a = r(3 * 144, dt=f32), r(3 * 288, dt=f32)


def once():
    ops.conv_igemm(cv["x"], 64, Bc, Hc, Hc, 64, cv["wf"], 64, 3, 3, 3, cv["y"], 64, sums=bn["sums"])
    bn["sums"].zero_()
    ops.conv_igemm(cv["dy"], 64, Bc, Hc, Hc, 64, cv["wd"], 64, 3, 3, 3, cv["dx"], 64)
    ops.conv_wgrad(cv["dy"], 64, cv["x"], 64, cv["G"], Bc, Hc, Hc, 64, 64, 3, 3, 3)
    ops.attn_fwd(q1, b1a, o1, l1, 12, 88, 88, 2, 72, 8, False)
    ops.attn_bwd(q1, b1a, o1, l1, do1, dq1, 12, 88, 88, 2, 72, 8, False)
    ops.attn_fwd(q2, b2a, o2, l2, 12, 88, 88, 4, 72, 8, True)
    ops.attn_bwd(q2, b2a, o2, l2, do2, dq2, 12, 88, 88, 4, 72, 8, True)
    ops.adapter_ln_fwd(ad["xs"], ad["W1"], ad["b1"], ad["W2"], ad["b2"], ln["g"], ln["b"], ad["xa"], ad["n1"], ln["mean"],
                       ln["rstd"], ad["u"], ad["g1"], ad["g2"], R, Cl)
    ops.adapter_ln_bwd(ln["dy"], ad["xa"], ln["mean"], ln["rstd"], ln["g"], ln["dres"], ad["g2"], ad["g1"], ad["W2t"],
                       ad["W1t"], ad["dh2"], ad["dh1"], ad["dx"], ad["db1"], ad["db2"], R, Cl)
    _lib.call("s2u_set_attn_backend", 2)
    ops.attn_fwd(qkv, bias, out, lse, B, H, H, nh, hd, 16, False)
    ops.attn_bwd(qkv, bias, out, lse, dout, dqkv, B, H, H, nh, hd, 16, False)
    ops.attn_fwd(qkv, bias, out, lse, B, H, H, nh, hd, 0, False)
    ops.attn_bwd(qkv, bias, out, lse, dout, dqkv, B, H, H, nh, hd, 0, False)
    _lib.call("s2u_set_attn_backend", 0)
    ops.ln_fwd(ln["x"], ln["g"], ln["b"], ln["y"], ln["mean"], ln["rstd"], R, Cl)
    ops.ln_bwd(ln["dy"], ln["x"], ln["g"], ln["mean"], ln["rstd"], ln["dres"], ln["dx"], R, Cl)
    ops.bn_stats_finalize(bn["x"], Cb, bn["sums"], bn["gamma"], bn["beta"], bn["rm"], bn["rv"], bn["nb"], bn["scale"],
                          bn["shift"], bn["mean"], bn["rstd"], M, Cb)
    ops.bn_apply(bn["x"], Cb, bn["scale"], bn["shift"], None, 0, bn["out"], Cb, M, Cb, True)
    ops.bn_bwd(bn["dy"], Cb, bn["out"], Cb, bn["x"], Cb, bn["mean"], bn["rstd"], bn["gamma"], bn["sums"], bn["dgamma"],
               bn["dbeta"], bn["c1"], bn["c2"], bn["dx"], Cb, M, Cb)
    _lib.call("s2u_structure_loss_fwd", ls["p"][0].data_ptr(), ls["p"][1].data_ptr(), ls["p"][2].data_ptr(), ls["m"].data_ptr(),
              ls["w"].data_ptr(), ls["sums"].data_ptr(), ls["loss"].data_ptr(), Bl, Sl, Sl, 3, ops.stream)
    _lib.call("s2u_structure_loss_bwd", ls["p"][0].data_ptr(), ls["p"][1].data_ptr(), ls["p"][2].data_ptr(), ls["m"].data_ptr(),
              ls["w"].data_ptr(), ls["sums"].data_ptr(), 0, ls["g"][0].data_ptr(), ls["g"][1].data_ptr(), ls["g"][2].data_ptr(),
              Bl, Sl, Sl, 3, ops.stream)


once()
torch.cuda.synchronize()
once()
torch.cuda.synchronize()
print("ncu targets done", float(out.float().abs().mean()), float(ls["loss"].sum()))

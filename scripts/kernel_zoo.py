"""GPU-box micro-benchmark of the non-GEMM kernels at the dominant shapes of the Hiera-L 352x352 batch-12 step.
CUDA events around a graph replay of back-to-back launches; operands rotate through enough buffer sets to exceed the 126 MB L2, so
every launch streams from HBM.  Prints us/launch and algorithmic GB/s (bytes each kernel must move at least once)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sam2_unet_b200 import _lib  # noqa: E402
from sam2_unet_b200.engine import Ops  # noqa: E402

dev = torch.device("cuda:0")
ops = Ops(torch.bfloat16, dev)
bf, f32 = torch.bfloat16, torch.float32
only = sys.argv[1] if len(sys.argv) > 1 else ""


def bench(name, nbytes, make, run, iters=24):
    if only and only not in name:
        return
    nsets = max(2, int(200e6 // max(nbytes, 1)) + 1)
    nsets = min(nsets, 12)
    sets = [make() for _ in range(nsets)]
    for i in range(nsets):
        run(sets[i])
    torch.cuda.synchronize()
    # capture the launch sequence once and time graph replays: no CPU launch overhead in the measurement
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for i in range(iters):
                run(sets[i % nsets])
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    print(f"{name:46s} {us:8.1f} us   {nbytes / us / 1e3:7.0f} GB/s   ({nbytes / 1e6:.1f} MB)", flush=True)


def r(*s, dt=bf):
    return torch.randn(*s, device=dev).to(dt)


for (R, C) in [(5808, 576), (23232, 288), (92928, 144), (1452, 1152)]:
    def mk():
        return dict(x=r(R, C, dt=f32), g=r(C, dt=f32), b=r(C, dt=f32), y=torch.empty(R, C, device=dev, dtype=bf),
                    mean=torch.empty(R, device=dev), rstd=torch.empty(R, device=dev), dy=r(R, C), dres=r(R, C),
                    dx=torch.empty(R, C, device=dev, dtype=bf), pre=r(R, C),
                    dx2=torch.empty(R, C, device=dev, dtype=bf), cs=torch.zeros(C, device=dev))
    bench(f"ln_fwd {R}x{C}", R * C * 6, mk,
          lambda s: ops.ln_fwd(s["x"], s["g"], s["b"], s["y"], s["mean"], s["rstd"], R, C))

    def run_b(s):
        ops.ln_fwd(s["x"], s["g"], s["b"], s["y"], s["mean"], s["rstd"], R, C)
    sets_ready = None
    def mk2():
        s = mk()
        ops.ln_fwd(s["x"], s["g"], s["b"], s["y"], s["mean"], s["rstd"], R, C)
        return s
    bench(f"ln_bwd plain {R}x{C}", R * C * 10, mk2,
          lambda s: ops.ln_bwd(s["dy"], s["x"], s["g"], s["mean"], s["rstd"], s["dres"], s["dx"], R, C))
    bench(f"ln_bwd fused(saved gelu' + colsum) {R}x{C}", R * C * 14, mk2,
          lambda s: ops.ln_bwd(s["dy"], s["x"], s["g"], s["mean"], s["rstd"], s["dres"], s["dx"], R, C, pre=s["pre"],
                               dx2=s["dx2"], colsum=s["cs"], pre_is_grad=True))

for (M, C) in [(92928, 64), (23232, 64), (92928, 32)]:
    def mkb():
        return dict(x=r(M, C), sums=torch.zeros(_lib.load().s2u_bn_ws_doubles(C), device=dev, dtype=torch.float64), gamma=r(C, dt=f32),
                    beta=r(C, dt=f32), rm=torch.zeros(C, device=dev), rv=torch.ones(C, device=dev),
                    scale=torch.empty(C, device=dev), shift=torch.empty(C, device=dev), mean=torch.empty(C, device=dev),
                    rstd=torch.empty(C, device=dev), out=torch.empty(M, C, device=dev, dtype=bf), dy=r(M, C),
                    dgamma=torch.zeros(C, device=dev), dbeta=torch.zeros(C, device=dev), c1=torch.empty(C, device=dev),
                    c2=torch.empty(C, device=dev), dx=torch.empty(M, C, device=dev, dtype=bf))

    def stats(s):
        ops.bn_stats_finalize(s["x"], C, s["sums"], s["gamma"], s["beta"], s["rm"], s["rv"], None, s["scale"],
                              s["shift"], s["mean"], s["rstd"], M, C)
    bench(f"bn_stats_finalize {M}x{C}", M * C * 2, mkb, stats)
    bench(f"bn_apply relu {M}x{C}", M * C * 4, mkb,
          lambda s: ops.bn_apply(s["x"], C, s["scale"], s["shift"], None, 0, s["out"], C, M, C, True))

    def mkb2():
        s = mkb()
        stats(s)
        ops.bn_apply(s["x"], C, s["scale"], s["shift"], None, 0, s["out"], C, M, C, True)
        return s

    def bwd(s):
        ops.bn_bwd(s["dy"], C, s["out"], C, s["x"], C, s["mean"], s["rstd"], s["gamma"], s["sums"], s["dgamma"],
                   s["dbeta"], s["c1"], s["c2"], s["dx"], C, M, C)
    bench(f"bn_bwd (reduce+apply, relu mask) {M}x{C}", M * C * (3 * 2 + 3 * 2), mkb2, bwd)

B, H, nh, hd, win = 12, 22, 8, 72, 14
Cc = nh * hd
R = B * H * H


def mka():
    return dict(qkv=r(R, 3 * Cc), bias=r(3 * Cc, dt=f32), o=torch.empty(R, Cc, device=dev, dtype=bf),
                lse=torch.empty(R, nh, device=dev), do=r(R, Cc), dqkv=torch.empty(R, 3 * Cc, device=dev, dtype=bf))


flops_f = 4.0 * B * 4 * nh * 196 * 196 * hd


def mka2():
    s = mka()
    ops.attn_fwd(s["qkv"], s["bias"], s["o"], s["lse"], B, H, H, nh, hd, 16, False)
    return s


for _be, _nm in ((1, "mma.sync"), (2, "tcgen05")):
    _lib.call("s2u_set_attn_backend", _be)
    bench(f"attn_fwd win16 12x22x22 h8 d72 [{_nm}]", R * Cc * 8, mka,
          lambda s: ops.attn_fwd(s["qkv"], s["bias"], s["o"], s["lse"], B, H, H, nh, hd, 16, False))
    bench(f"attn_bwd win16 12x22x22 h8 d72 [{_nm}]", R * Cc * 16, mka2,
          lambda s: ops.attn_bwd(s["qkv"], s["bias"], s["o"], s["lse"], s["do"], s["dqkv"], B, H, H, nh, hd, 16, False))
    bench(f"attn_fwd global 12x22x22 h8 d72 [{_nm}]", R * Cc * 8, mka,
          lambda s: ops.attn_fwd(s["qkv"], s["bias"], s["o"], s["lse"], B, H, H, nh, hd, 0, False))
    bench(f"attn_bwd global 12x22x22 h8 d72 [{_nm}]", R * Cc * 16, mka2,
          lambda s: ops.attn_bwd(s["qkv"], s["bias"], s["o"], s["lse"], s["do"], s["dqkv"], B, H, H, nh, hd, 0, False))
_lib.call("s2u_set_attn_backend", 0)
bench("attn_fwd win14 12x22x22 h8 d72", R * Cc * 8, mka,
      lambda s: ops.attn_fwd(s["qkv"], s["bias"], s["o"], s["lse"], B, H, H, nh, hd, win, False))


bench("attn_bwd win14 12x22x22 h8 d72", R * Cc * 16, mka2,
      lambda s: ops.attn_bwd(s["qkv"], s["bias"], s["o"], s["lse"], s["do"], s["dqkv"], B, H, H, nh, hd, win, False))
bench("attn_fwd global 12x22x22 h8 d72", R * Cc * 8, mka,
      lambda s: ops.attn_fwd(s["qkv"], s["bias"], s["o"], s["lse"], B, H, H, nh, hd, 0, False))
bench("attn_bwd global 12x22x22 h8 d72", R * Cc * 16, mka2,
      lambda s: ops.attn_bwd(s["qkv"], s["bias"], s["o"], s["lse"], s["do"], s["dqkv"], B, H, H, nh, hd, 0, False))
print(f"(attention fwd algorithmic FLOPs, windowed: {flops_f / 1e9:.2f} GFLOP; bwd 2.5x)")
# the same kernels on windows that fill their 64-row tiles exactly (16x16 = 256 tokens, 32x32 map): tile-waste reference
H2 = 32
R2 = B * H2 * H2


def mkb_():
    return dict(qkv=r(R2, 3 * Cc), bias=r(3 * Cc, dt=f32), o=torch.empty(R2, Cc, device=dev, dtype=bf),
                lse=torch.empty(R2, nh, device=dev), do=r(R2, Cc), dqkv=torch.empty(R2, 3 * Cc, device=dev, dtype=bf))


def mkb2_():
    s = mkb_()
    ops.attn_fwd(s["qkv"], s["bias"], s["o"], s["lse"], B, H2, H2, nh, hd, 16, False)
    return s


bench("attn_fwd win16 12x32x32 h8 d72 (full tiles)", R2 * Cc * 8, mkb_,
      lambda s: ops.attn_fwd(s["qkv"], s["bias"], s["o"], s["lse"], B, H2, H2, nh, hd, 16, False))
bench("attn_bwd win16 12x32x32 h8 d72 (full tiles)", R2 * Cc * 16, mkb2_,
      lambda s: ops.attn_bwd(s["qkv"], s["bias"], s["o"], s["lse"], s["do"], s["dqkv"], B, H2, H2, nh, hd, 16, False))
print(f"(win16 fwd FLOPs: {4.0 * B * 4 * nh * 256 * 256 * hd / 1e9:.2f} GFLOP)")


# structure_loss forward + backward, three heads sharing one mask (train.py:21-29,76-79): algorithmic bytes =
# forward: 3 logit maps + mask read, weit written (20 B/pixel); backward: 3 logits + mask + weit read, 3 gradients
# written (32 B/pixel)
for (Bl, Sl) in [(12, 352), (4, 1024)]:
    npx = Bl * Sl * Sl

    def mkl():
        return dict(p=[r(Bl, 1, Sl, Sl, dt=f32) for _ in range(3)], m=(torch.rand(Bl, 1, Sl, Sl, device=dev) > 0.7).float(),
                    w=torch.empty(Bl, Sl, Sl, device=dev), sums=torch.zeros(3 * Bl * 2 + 3, device=dev, dtype=torch.float64),
                    loss=torch.empty(3, device=dev), g=[torch.empty(Bl, 1, Sl, Sl, device=dev) for _ in range(3)])

    def lfwd(s):
        _lib.call("s2u_structure_loss_fwd", s["p"][0].data_ptr(), s["p"][1].data_ptr(), s["p"][2].data_ptr(), s["m"].data_ptr(),
                  s["w"].data_ptr(), s["sums"].data_ptr(), s["loss"].data_ptr(), Bl, Sl, Sl, 3, ops.stream)

    def lbwd(s):
        _lib.call("s2u_structure_loss_bwd", s["p"][0].data_ptr(), s["p"][1].data_ptr(), s["p"][2].data_ptr(), s["m"].data_ptr(),
                  s["w"].data_ptr(), s["sums"].data_ptr(), 0, s["g"][0].data_ptr(), s["g"][1].data_ptr(), s["g"][2].data_ptr(),
                  Bl, Sl, Sl, 3, ops.stream)

    def mkl2():
        s = mkl()
        lfwd(s)
        return s
    bench(f"structure_loss fwd 3 heads {Bl}x{Sl}x{Sl} (+memset, finalize)", npx * 20, mkl, lfwd)
    bench(f"structure_loss bwd 3 heads {Bl}x{Sl}x{Sl}", npx * 32, mkl2, lbwd)

for (M, P) in [(5808, 576), (5808, 32)]:
    def mkc():
        return dict(a=r(M, P), out=torch.zeros(P, device=dev))
    bench(f"colsum {M}x{P}", M * P * 2, mkc, lambda s: ops.colsum(s["a"], s["out"]))

for (Bc, Hc, Cin, K) in [(12, 88, 64, 3), (12, 44, 64, 3), (12, 88, 64, 7)]:
    M = Bc * Hc * Hc
    taps = K * (K if K == 3 else 1)
    def mki():
        return dict(x=r(M, Cin), col=torch.empty(M, taps * Cin, device=dev, dtype=bf))
    kh, kw = (3, 3) if K == 3 else (1, 7)
    bench(f"im2col {Bc}x{Hc}x{Hc}x{Cin} k{kh}x{kw}", M * Cin * 2 * (1 + taps), mki,
          lambda s: ops.im2col(s["x"], Cin, s["col"], Bc, Hc, Hc, Cin, kh, kw, 1, 1, kh // 2, kw // 2))

# inference tail (test.py:66-76) on the device vs the reference's own CPU operators on this box
if not only or "tail" in only:
    import time
    import torch.nn.functional as F
    from oracle import port
    from sam2_unet_b200 import infer_tail
    for (S, pad, hw) in [(352, (0, 0, 0, 117), (480, 720)), (352, (0, 0, 0, 0), (1080, 1920)), (1024, (0, 0, 256, 0), (1536, 1152))]:
        lg = F.interpolate(torch.randn(1, 1, S // 8, S // 8) * 3, size=(S, S), mode="bicubic").contiguous()
        lgd = lg.to(dev)
        for _ in range(3):
            o = infer_tail(lgd, pad, hw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            o = infer_tail(lgd, pad, hw)
        e1.record()
        torch.cuda.synchronize()
        t_dev = e0.elapsed_time(e1) / 20 * 1e3
        t0 = time.perf_counter()
        for _ in range(5):
            o.cpu()
        t_d2h = (time.perf_counter() - t0) / 5 * 1e6
        t0 = time.perf_counter()
        for _ in range(5):
            port.infer_tail(lgd.cpu(), pad, hw)
        t_cpu = (time.perf_counter() - t0) / 5 * 1e6
        print(f"infer_tail {S}->{hw[0]}x{hw[1]}: device {t_dev:7.1f} us (3 launches, eager) + D2H of uint8 {t_d2h:7.1f} us"
              f"   | reference path (D2H of fp32 map + torch/numpy on the host): {t_cpu:9.1f} us", flush=True)

# evaluation metrics (eval.py:55-171) on the device vs the CPU restatement of the reference on this box
if not only or "eval" in only:
    import time
    import numpy as np
    from oracle import eval_port
    from sam2_unet_b200 import evaluate_segmentation_performance
    for (h, w, nb) in [(480, 720, 40), (1080, 1920, 80)]:
        rng = np.random.default_rng(0)
        yy, xx = np.mgrid[0:h, 0:w]
        gt = np.zeros((h, w), np.uint8)
        for _ in range(nb):
            cy, cx, rr = rng.integers(0, h), rng.integers(0, w), rng.integers(4, 40)
            gt[(yy - cy) ** 2 + (xx - cx) ** 2 <= rr * rr] = 255
        pred = np.roll(gt, (3, -2), (0, 1))
        pd, gd = torch.from_numpy(pred).to(dev), torch.from_numpy(gt).to(dev)
        for _ in range(2):
            evaluate_segmentation_performance(pd, gd)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            evaluate_segmentation_performance(pd, gd)
        t_dev = (time.perf_counter() - t0) / 5 * 1e3
        t0 = time.perf_counter()
        ref = eval_port.evaluate_segmentation_performance(pred, gt)
        t_cpu = (time.perf_counter() - t0) * 1e3
        print(f"eval metrics {h}x{w}, {ref['count_gt']} components: device path {t_dev:7.2f} ms wall (kernels + host "
              f"matching) | reference algorithm on the host: {t_cpu:9.1f} ms", flush=True)

# test-time input path (dataset.py:336-407) on the device vs the reference's torchvision transforms on this box
if not only or "prep" in only:
    import time
    import numpy as np
    from oracle import port
    from sam2_unet_b200 import preprocess_image
    for (h, w, S) in [(1080, 1920, 352), (2000, 1500, 1024)]:
        img = np.random.default_rng(0).integers(0, 256, (h, w, 3)).astype(np.uint8)
        imd = torch.from_numpy(img).to(dev)
        for _ in range(3):
            preprocess_image(imd, S)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            preprocess_image(imd, S)
        e1.record()
        torch.cuda.synchronize()
        t_dev = e0.elapsed_time(e1) / 20 * 1e3
        t0 = time.perf_counter()
        for _ in range(3):
            port.preprocess_image(img, S)
        t_cpu = (time.perf_counter() - t0) / 3 * 1e3
        print(f"preprocess {h}x{w} -> {S}: device {t_dev:7.1f} us (2 launches, eager; + H2D of the uint8 image) | "
              f"reference transforms on the host: {t_cpu:7.2f} ms", flush=True)

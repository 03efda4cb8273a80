"""tcgen05 attention (attention_tc.cu) against the mma.sync kernels and a PyTorch fp32 expression of the reference
(hieradet.py:56-81), plus timings of both kernel families.  GPU only:  python scripts/attn_tc_check.py [--bwd] [--time]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sam2_unet_b200 import _lib
from sam2_unet_b200.engine import Ops
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from test_kernels_gpu import _attn_reference, _rand

CASES = [  # B, H, W, nh, hd, window
    (1, 16, 16, 1, 72, 16), (1, 22, 22, 2, 72, 16), (2, 32, 32, 2, 72, 16), (1, 22, 22, 1, 72, 0), (2, 22, 22, 2, 72, 0),
    (1, 22, 22, 2, 96, 14), (1, 22, 22, 1, 96, 0), (1, 64, 64, 1, 72, 0), (12, 22, 22, 8, 72, 16),
]


def run(case, backend, bwd, cuda):
    B, H, W, nh, hd, window = case
    C = nh * hd
    ops = Ops(torch.bfloat16, cuda, 0)
    qkv = _rand((B, H, W, 3 * C), "bf16", cuda, 1)
    bias = _rand((3 * C,), "fp32", cuda, 2)
    out = torch.full((B, H, W, C), float("nan"), device=cuda, dtype=torch.bfloat16)
    lse = torch.full((B, H, W, nh), float("nan"), device=cuda)
    _lib.call("s2u_set_attn_backend", backend)
    ops.attn_fwd(qkv, bias, out, lse, B, H, W, nh, hd, window, False)
    res = {"out": out, "lse": lse}
    if bwd:
        dout = _rand((B, H, W, C), "bf16", cuda, 3)
        dqkv = torch.full((B, H, W, 3 * C), float("nan"), device=cuda, dtype=torch.bfloat16)
        ops.attn_bwd(qkv, bias, out, lse, dout, dqkv, B, H, W, nh, hd, window, False)
        res["dqkv"] = dqkv
        res["dout"] = dout
    torch.cuda.synchronize()
    res["qkv"], res["bias"] = qkv, bias
    return res


def err(a, b):
    a, b = a.float(), b.float()
    if not torch.isfinite(a).all():
        return float("nan")
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-6)).item()


def timeit(fn, iters=24, reps=5):
    """us per call: `iters` calls captured in one CUDA graph (no Python / launch latency in the number), replayed."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (iters * reps) * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bwd", action="store_true")
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--cases", type=str, default="")
    args = ap.parse_args()
    cuda = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    cases = CASES if not args.cases else [CASES[int(i)] for i in args.cases.split(",")]
    bad = 0
    for case in cases:
        B, H, W, nh, hd, window = case
        new = run(case, 2, args.bwd, cuda)
        old = run(case, 1, args.bwd, cuda)
        qr = new["qkv"].float().requires_grad_(True)
        ref = _attn_reference(qr, new["bias"].bfloat16().float(), B, H, W, nh, hd, window, False)
        line = f"{case}: fwd err new {err(new['out'], ref):.2e} old {err(old['out'], ref):.2e}  lse new-old {err(new['lse'], old['lse']):.2e}"
        ok = err(new["out"], ref) < 2e-2 and err(new["lse"], old["lse"]) < 1e-3
        if args.bwd:
            (gq,) = torch.autograd.grad(ref, qr, new["dout"].float())
            C = nh * hd
            for name, sl in (("dq", slice(0, C)), ("dk", slice(C, 2 * C)), ("dv", slice(2 * C, 3 * C))):
                en, eo = err(new["dqkv"][..., sl], gq[..., sl]), err(old["dqkv"][..., sl], gq[..., sl])
                line += f" | {name} new {en:.2e} old {eo:.2e}"
                ok = ok and en < 3e-2
        print(("ok   " if ok else "FAIL ") + line, flush=True)
        bad += 0 if ok else 1
    if args.time:
        for case in [(12, 22, 22, 8, 72, 16), (12, 22, 22, 8, 72, 0), (12, 32, 32, 8, 72, 16), (4, 64, 64, 8, 72, 16),
                     (4, 64, 64, 8, 72, 0)]:
            B, H, W, nh, hd, window = case
            C = nh * hd
            sets = []
            for i in range(6):                        # rotate through > L2 worth of operands
                qkv = _rand((B, H, W, 3 * C), "bf16", cuda, 10 + i)
                sets.append((qkv, torch.empty(B, H, W, C, device=cuda, dtype=torch.bfloat16),
                             torch.empty(B, H, W, nh, device=cuda), _rand((B, H, W, C), "bf16", cuda, 20 + i),
                             torch.empty(B, H, W, 3 * C, device=cuda, dtype=torch.bfloat16)))
            bias = _rand((3 * C,), "fp32", cuda, 2)
            dws = torch.empty(B, H, W, nh, device=cuda)
            for q, o, l, d, dq in sets:               # valid out / lse for the backward
                _lib.call("s2u_set_attn_backend", 1)
                Ops(torch.bfloat16, cuda, 0).attn_fwd(q, bias, o, l, B, H, W, nh, hd, window, False)
            for backend, name in ((1, "mma.sync"), (2, "tcgen05 ")):
                _lib.call("s2u_set_attn_backend", backend)
                k = [0]

                def f():
                    q, o, l, d, dq = sets[k[0] % len(sets)]
                    k[0] += 1
                    Ops(torch.bfloat16, cuda, 0).attn_fwd(q, bias, o, l, B, H, W, nh, hd, window, False)

                def g():
                    q, o, l, d, dq = sets[k[0] % len(sets)]
                    k[0] += 1
                    _lib.call("s2u_win_attn_bwd", q.data_ptr(), bias.data_ptr(), o.data_ptr(), l.data_ptr(), d.data_ptr(),
                              dq.data_ptr(), dws.data_ptr(), B, H, W, nh, hd, window, 0, 1,
                              torch.cuda.current_stream().cuda_stream)
                t_f = timeit(f)
                msg = f"{case} {name}: fwd {t_f:7.1f} us"
                if args.bwd:
                    msg += f"  bwd {timeit(g):7.1f} us"
                print(msg, flush=True)
    _lib.call("s2u_set_attn_backend", 0)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()

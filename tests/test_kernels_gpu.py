"""Op-level parity of every CUDA kernel (through the C ABI) against a plain PyTorch fp32 expression of the same
op on identical inputs.  fp32 mode: <= 1e-3 of the output scale (observed ~1e-6); bf16 mode: inputs are rounded
to bf16 first and the tolerance scales with the bf16 output rounding (2^-8)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DT = {"fp32": torch.float32, "bf16": torch.bfloat16}


def _ops(dtype, cuda, backend=0):
    from sam2_unet_b200.engine import Ops
    return Ops(DT[dtype], cuda, backend)


def _close(got, ref, tol, what=""):
    got, ref = got.float(), ref.float()
    scale = ref.abs().max().clamp_min(1e-6)
    err = (got - ref).abs().max() / scale
    assert torch.isfinite(got).all(), f"{what}: non-finite output"
    assert err.item() <= tol, f"{what}: max-normalised error {err.item():.3e} > {tol}"


def _tol(dtype, fp32=1e-4, bf16=2e-2):
    return fp32 if dtype == "fp32" else bf16


def _rand(shape, dtype, cuda, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(cuda).to(DT[dtype])


# --------------------------------------------------------------------------------------------------- GEMM

GEMM_SHAPES = [(300, 96, 32), (1000, 432, 144), (484, 2304, 576), (777, 32, 144), (128, 64, 64), (130, 72, 40),
               (5808, 576, 2304), (64, 256, 1152)]


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("shape", GEMM_SHAPES)
@pytest.mark.parametrize("flags", [0, 1, 4, 5, 2, 256, 1 | 128, 4 | 512])
def test_gemm_simt(cuda, dtype, shape, flags):
    M, N, K = shape
    ops = _ops(dtype, cuda, backend=1)
    A, W = _rand((M, K), dtype, cuda, 1), _rand((N, K), dtype, cuda, 2, K ** -0.5)
    bias = _rand((N,), "fp32", cuda, 3)
    aux, resid = _rand((M, N), dtype, cuda, 4), _rand((M, N), dtype, cuda, 5)
    C, pre = ops.empty(M, N), ops.empty(M, N)
    ops.gemm(A, W, C, bias=bias, pre_out=pre, aux=aux if flags & (2 | 256) else None,
             resid=resid if flags & 4 else None, flags=flags)
    ref_pre = A.float() @ W.float().t() + bias
    ref = ref_pre
    if flags & 1:
        ref = F.gelu(ref)
    if flags & 256:
        ref = ref * aux.float()
    if flags & 128:                                       # pre_out holds gelu'(pre-activation)
        p = ref_pre.clone().requires_grad_(True)
        (ref_pre,) = torch.autograd.grad(F.gelu(p).sum(), p)
    if flags & 2:
        a = aux.float().requires_grad_(True)
        (dg,) = torch.autograd.grad(F.gelu(a).sum(), a)
        ref = ref * dg
    if flags & 4:
        ref = ref + resid.float()
    if flags & 512:
        ref = F.relu(ref)
    _close(pre, ref_pre, _tol(dtype), "pre_out")
    _close(C, ref, _tol(dtype), "C")


@pytest.mark.parametrize("shape", GEMM_SHAPES + [(92928, 144, 144), (1452, 4608, 1152), (256, 1152, 4608)])
@pytest.mark.parametrize("kernel", ["auto", "ws32", "ws64", "ws128", "ws256", "pair32", "pair64", "pair128", "pair144",
                                    "pair192", "pair256"])
def test_gemm_umma(cuda, shape, kernel):
    """tcgen05 + TMA GEMM vs fp32 matmul of the same bf16 operands: the one-CTA persistent kernel and the CTA-pair
    (cta_group::2) kernel at every tile width, with every epilogue the engine uses."""
    M, N, K = shape
    if K % 8 or N % 8:
        pytest.skip("TMA path needs 16-byte row pitches (the engine routes such shapes to the SIMT kernel)")
    ops = _ops("bf16", cuda)
    A, W = _rand((M, K), "bf16", cuda, 1), _rand((N, K), "bf16", cuda, 2, K ** -0.5)
    bias = _rand((N,), "fp32", cuda, 3)
    resid = _rand((M, N), "bf16", cuda, 5)
    aux = _rand((M, N), "bf16", cuda, 7)
    C, pre = ops.empty(M, N), ops.empty(M, N)
    backend = 2 if kernel == "auto" else (16 + int(kernel[2:]) if kernel.startswith("ws") else 1024 + int(kernel[4:]))
    ops.gemm(A, W, C, bias=bias, pre_out=pre, resid=resid, flags=1 | 4, backend=backend)
    prod = A.float() @ W.float().t()
    ref_pre = prod + bias
    ref = F.gelu(ref_pre) + resid.float()
    _close(pre, ref_pre, 1e-2, "pre_out")
    _close(C, ref, 1e-2, "C")
    # plain (no epilogue)
    C2 = ops.empty(M, N)
    ops.gemm(A, W, C2, backend=backend)
    _close(C2, prod, 1e-2, "plain")
    # GELU' of a saved pre-activation multiplies the result (backward of fc2 / adapter layer 2)
    a = aux.float().requires_grad_(True)
    (dg,) = torch.autograd.grad(F.gelu(a).sum(), a)
    C5 = ops.empty(M, N)
    ops.gemm(A, W, C5, aux=aux, flags=2, backend=backend)
    _close(C5, prod * dg, 1e-2, "dgelu")
    # the same with GELU' evaluated by the forward epilogue: pre_out <- gelu'(pre), backward multiplies by it
    C6, d6 = ops.empty(M, N), ops.empty(M, N)
    ops.gemm(A, W, C6, bias=bias, pre_out=d6, flags=1 | 128, backend=backend)
    p = ref_pre.clone().requires_grad_(True)
    (dp,) = torch.autograd.grad(F.gelu(p).sum(), p)
    _close(C6, F.gelu(ref_pre), 1e-2, "gelu (save gelu')")
    _close(d6, dp, 1e-2, "saved gelu'")
    C7 = ops.empty(M, N)
    ops.gemm(A, W, C7, aux=aux, flags=256, backend=backend)
    _close(C7, prod * aux.float(), 1e-2, "mulaux")
    # fp32 residual stream: fp32 residual in, fp32 C out, bf16 copy of the final value
    r32 = _rand((M, N), "fp32", cuda, 6)
    C3, cp = torch.empty(M, N, device=cuda), ops.empty(M, N)
    ops.gemm(A, W, C3, bias=bias, resid=r32, pre_out=cp, flags=4 | 16 | 32 | 64, backend=backend)
    ref3 = prod + bias + r32
    _close(C3, ref3, 1e-4, "fp32 stream out")
    _close(cp, ref3, 1e-2, "bf16 copy of the stream")
    C4, h4 = torch.empty(M, N, device=cuda), ops.empty(M, N)
    ops.gemm(A, W, C4, bias=bias, resid=resid, pre_out=h4, flags=1 | 4 | 16, backend=backend)     # bf16 resid, fp32 out
    _close(C4, F.gelu(ref_pre) + resid.float(), 1e-3, "gelu + bf16 resid, fp32 out")
    _close(h4, ref_pre, 1e-2, "pre-activation copy")
    # eval-mode conv + folded BatchNorm: bias, residual, ReLU, written into a wider (concat) buffer
    wide = ops.empty(M, N + 64)
    ops.gemm(A, W, wide.view(-1)[8:], bias=bias, resid=resid, flags=4 | 512, ldc=N + 64, backend=backend)
    _close(wide[:, 8:8 + N], F.relu(ref_pre + resid.float()), 1e-2, "bias + resid + relu into a strided buffer")
    # adapter layer 2 forward: gelu + fp32 residual in / out with the pre-activation saved
    C8, h8 = torch.empty(M, N, device=cuda), ops.empty(M, N)
    ops.gemm(A, W, C8, bias=bias, resid=r32, pre_out=h8, flags=1 | 4 | 16 | 32, backend=backend)
    _close(C8, F.gelu(ref_pre) + r32, 1e-3, "gelu + fp32 resid, fp32 out")
    _close(h8, ref_pre, 1e-2, "pre-activation copy (stream)")


@pytest.mark.parametrize("shape", [(3000, 32, 144), (5808, 576, 32), (777, 64, 64), (92928, 64, 576), (100, 64, 2304),
                                   (1452, 1152, 32), (4096, 64, 136)])
def test_wgrad_tcgen05_shapes(cuda, shape):
    """bf16 weight-gradient GEMM (MN-major tcgen05 path) on the adapter / conv shapes, ragged row counts included."""
    M, P, Q = shape
    ops = _ops("bf16", cuda)
    A, B = _rand((M, P), "bf16", cuda, 1), _rand((M, Q), "bf16", cuda, 2)
    G = torch.zeros(P, Q, device=cuda)
    ops.wgrad(A, B, G)
    _close(G, A.float().t() @ B.float(), 2e-3, "wgrad")


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_wgrad_colsum(cuda, dtype):
    ops = _ops(dtype, cuda)
    M, P, Q = 3000, 64, 192
    A, B = _rand((M, P), dtype, cuda, 1), _rand((M, Q), dtype, cuda, 2)
    G = torch.zeros(P, Q, device=cuda)
    ops.wgrad(A, B, G)
    ops.wgrad(A, B, G)                                    # accumulates
    _close(G, 2 * A.float().t() @ B.float(), 1e-4, "wgrad")
    # conv-weight mapping: q = tap*Cin + ci -> [P][Cin][taps]
    taps, cin = 3, 64
    G2 = torch.zeros(P, cin, taps, device=cuda)
    ops.wgrad(A, B, G2, ldg=Q, q_inner=cin, q_taps=taps)
    ref = (A.float().t() @ B.float()).view(P, taps, cin).permute(0, 2, 1)
    _close(G2, ref, 1e-4, "wgrad conv layout")
    s = torch.zeros(Q, device=cuda)
    ops.colsum(B, s)
    _close(s, B.float().sum(0), 1e-4, "colsum")


# ------------------------------------------------------------------------------------------ norm / pointwise

@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("C", [32, 64, 96, 144, 288, 576, 1152])
def test_layernorm(cuda, dtype, C):
    ops = _ops(dtype, cuda)
    R = 1001
    x = _rand((R, C), dtype, cuda, 1) * 2 + 0.5
    g, b = _rand((C,), "fp32", cuda, 2) * 0.1 + 1, _rand((C,), "fp32", cuda, 3) * 0.1
    y, mean, rstd = ops.empty(R, C), ops.empty(R, dtype=torch.float32), ops.empty(R, dtype=torch.float32)
    ops.ln_fwd(x, g, b, y, mean, rstd, R, C)
    xr = x.float().requires_grad_(True)
    ref = F.layer_norm(xr, (C,), g, b, 1e-6)
    _close(y, ref, _tol(dtype), "ln fwd")
    dy, dres = _rand((R, C), dtype, cuda, 4), _rand((R, C), dtype, cuda, 5)
    dx = ops.empty(R, C)
    ops.ln_bwd(dy, x, g, mean, rstd, dres, dx, R, C)
    (gx,) = torch.autograd.grad(ref, xr, dy.float())
    _close(dx, gx + dres.float(), _tol(dtype), "ln bwd")
    # fused adapter tail: dx2 = dx * gelu'(pre), colsum += column sums of dx2
    pre = _rand((R, C), dtype, cuda, 6)
    dx_b, dx2, cs = ops.empty(R, C), ops.empty(R, C), torch.zeros(C, device=cuda)
    ops.ln_bwd(dy, x, g, mean, rstd, dres, dx_b, R, C, pre=pre, dx2=dx2, colsum=cs)
    assert torch.equal(dx_b, dx)
    pr = pre.float().requires_grad_(True)
    (dg,) = torch.autograd.grad(F.gelu(pr).sum(), pr)
    _close(dx2, dx.float() * dg, _tol(dtype, 1e-5), "ln bwd fused gelu'")
    _close(cs, dx2.float().sum(0), 1e-4, "ln bwd fused colsum")
    # the same with gelu'(pre) already evaluated (GEMM_SAVE_DGELU epilogue); the workspace must have come back clean,
    # and the column sums accumulate
    dgs = dg.to(DT[dtype])
    dx_c, dx2c = ops.empty(R, C), ops.empty(R, C)
    ops.ln_bwd(dy, x, g, mean, rstd, dres, dx_c, R, C, pre=dgs, dx2=dx2c, colsum=cs, pre_is_grad=True)
    assert torch.equal(dx_c, dx)
    _close(dx2c, dx.float() * dgs.float(), _tol(dtype, 1e-5), "ln bwd fused saved gelu'")
    _close(cs, dx2.float().sum(0) + dx2c.float().sum(0), 1e-4, "ln bwd fused colsum accumulates")
    if dtype == "bf16":                                   # fp32 residual stream in, bf16 out
        x32 = x.float()
        y2, m2, r2 = ops.empty(R, C), ops.empty(R, dtype=torch.float32), ops.empty(R, dtype=torch.float32)
        ops.ln_fwd(x32, g, b, y2, m2, r2, R, C)
        _close(y2, ref, _tol(dtype), "ln fwd (fp32 x)")
        dx3 = ops.empty(R, C)
        ops.ln_bwd(dy, x32, g, m2, r2, dres, dx3, R, C)
        _close(dx3, gx + dres.float(), _tol(dtype), "ln bwd (fp32 x)")


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_maxpool_dgelu(cuda, dtype):
    ops = _ops(dtype, cuda)
    B, H, W, C = 2, 10, 12, 64
    x = _rand((B, H, W, C), dtype, cuda, 1)
    out = ops.empty(B, H // 2, W // 2, C)
    ops.maxpool_fwd(x, out, B, H, W, C)
    xr = x.float().requires_grad_(True)
    ref = F.max_pool2d(xr.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
    _close(out, ref, 1e-6, "maxpool fwd")
    dout = _rand(tuple(out.shape), dtype, cuda, 2)
    dx = ops.empty(B, H, W, C)
    ops.maxpool_bwd(x, dout, dx, B, H, W, C)
    (gx,) = torch.autograd.grad(ref, xr, dout.float())
    _close(dx, gx, 1e-6, "maxpool bwd")
    pre, dy = _rand((64, 40), dtype, cuda, 3), _rand((64, 40), dtype, cuda, 4)
    o = ops.empty(64, 40)
    ops.dgelu_mul(dy, pre, o)
    pr = pre.float().requires_grad_(True)
    (gp,) = torch.autograd.grad(F.gelu(pr), pr, dy.float())
    _close(o, gp, _tol(dtype, 1e-5), "dgelu")


# ---------------------------------------------------------------------------------------------- attention

def _attn_reference(qkv, bias, B, H, W, nh, hd, window, pool):
    """hieradet.py:56-81 + utils.py:16-55 on an already-projected qkv map; padded tokens carry the bias."""
    C = nh * hd
    ws_h, ws_w = (window, window) if window > 0 else (H, W)
    ph, pw = (ws_h - H % ws_h) % ws_h, (ws_w - W % ws_w) % ws_w
    full = bias.view(1, 1, 1, -1).expand(B, H + ph, W + pw, 3 * C).clone()
    full[:, :H, :W] = qkv
    Hp, Wp = H + ph, W + pw
    t = full.view(B, Hp // ws_h, ws_h, Wp // ws_w, ws_w, 3 * C).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws_h, ws_w, 3 * C)
    q, k, v = t[..., :C], t[..., C:2 * C], t[..., 2 * C:]
    qh, qw = ws_h, ws_w
    if pool:
        q = F.max_pool2d(q.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
        qh, qw = ws_h // 2, ws_w // 2
    nw = t.shape[0]
    q = q.reshape(nw, qh * qw, nh, hd).transpose(1, 2)
    k = k.reshape(nw, ws_h * ws_w, nh, hd).transpose(1, 2)
    v = v.reshape(nw, ws_h * ws_w, nh, hd).transpose(1, 2)
    o = ((q @ k.transpose(-1, -2)) / math.sqrt(hd)).softmax(-1) @ v
    o = o.transpose(1, 2).reshape(nw, qh, qw, C)
    o = o.view(B, Hp // ws_h, Wp // ws_w, qh, qw, C).permute(0, 1, 3, 2, 4, 5).reshape(B, (Hp // ws_h) * qh,
                                                                                     (Wp // ws_w) * qw, C)
    Ho, Wo = (H // 2, W // 2) if pool else (H, W)
    return o[:, :Ho, :Wo]


ATTN_CASES = [  # B, H, W, nh, hd, window, pool
    (2, 16, 16, 2, 72, 8, False), (1, 22, 22, 2, 72, 16, False), (2, 11, 11, 4, 72, 8, False),
    (1, 22, 22, 1, 96, 0, False), (2, 16, 16, 2, 72, 8, True), (1, 22, 22, 2, 72, 16, True),
    (1, 12, 12, 2, 32, 4, True), (2, 10, 10, 1, 32, 6, True), (1, 5, 5, 2, 56, 3, False), (1, 22, 22, 2, 96, 14, False),
    # tiny windows (packed 4-per-CTA kernels): Hiera-L stage 2, its q-pool transition, ragged window counts
    (2, 44, 44, 4, 72, 4, False), (1, 44, 44, 8, 72, 4, True), (3, 12, 12, 2, 96, 4, False), (1, 6, 10, 1, 32, 4, False),
    # whole-window kernels (17..64 tokens per window): ragged maps, pooled ragged, odd window, other head dims
    (1, 20, 12, 4, 72, 8, True), (1, 12, 20, 2, 72, 8, False), (1, 24, 24, 1, 80, 8, False), (1, 14, 14, 2, 64, 7, False),
    (1, 88, 88, 2, 72, 8, False), (1, 88, 88, 4, 72, 8, True),
]


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("case", ATTN_CASES)
def test_window_attention(cuda, dtype, case):
    B, H, W, nh, hd, window, pool = case
    ops = _ops(dtype, cuda)
    C = nh * hd
    qkv = _rand((B, H, W, 3 * C), dtype, cuda, 1)
    bias = _rand((3 * C,), "fp32", cuda, 2)
    Ho, Wo = (H // 2, W // 2) if pool else (H, W)
    out, lse = ops.empty(B, Ho, Wo, C), ops.empty(B, Ho, Wo, nh, dtype=torch.float32)
    ops.attn_fwd(qkv, bias, out, lse, B, H, W, nh, hd, window, pool)
    qr = qkv.float().requires_grad_(True)
    b_eff = bias.to(DT[dtype]).float()
    ref = _attn_reference(qr, b_eff, B, H, W, nh, hd, window, pool)
    _close(out, ref, _tol(dtype, 1e-4, 2e-2), "attention fwd")
    dout = _rand((B, Ho, Wo, C), dtype, cuda, 3)
    dqkv = torch.full((B, H, W, 3 * C), float("nan"), device=cuda, dtype=DT[dtype])
    ops.attn_bwd(qkv, bias, out, lse, dout, dqkv, B, H, W, nh, hd, window, pool)
    (gq,) = torch.autograd.grad(ref, qr, dout.float())
    _close(dqkv, gq, _tol(dtype, 2e-4, 3e-2), "attention bwd")


# ----------------------------------------------------------------------------------------- stem and convs

@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_patch_embed(cuda, dtype):
    ops = _ops(dtype, cuda)
    B, S, E = 2, 96, 144
    x = _rand((B, 3, S, S), "fp32", cuda, 1)
    w, b = _rand((E, 3, 7, 7), "fp32", cuda, 2, 0.1), _rand((E,), "fp32", cuda, 3)
    pos = _rand((S // 4, S // 4, E), "fp32", cuda, 4)
    out = ops.empty(B, S // 4, S // 4, E)
    ops.patch_embed(x, w, b, pos, out, B, S, E)
    ref = F.conv2d(x, w, b, stride=4, padding=3).permute(0, 2, 3, 1) + pos
    _close(out, ref, _tol(dtype, 1e-5, 1e-2), "patch embed")
    o32, oc = torch.empty(B, S // 4, S // 4, E, device=cuda), ops.empty(B, S // 4, S // 4, E)
    ops.patch_embed(x, w, b, pos, o32, B, S, E, out_copy=oc)
    _close(o32, ref, 1e-5, "patch embed fp32 stream")
    assert torch.equal(oc, out)


def test_patch_embed_tensor_core(cuda):
    """bf16-mode stem: patch matrix with the image split into bf16 (hi, lo) halves x [W | W] on the GEMM kernel, bias +
    position table as the fp32 residual; against F.conv2d with the bf16-rounded weights in fp32."""
    ops = _ops("bf16", cuda)
    B, S, E = 2, 96, 144
    Hp = S // 4
    x = _rand((B, 3, S, S), "fp32", cuda, 1) * 2
    w, b = _rand((E, 3, 7, 7), "fp32", cuda, 2, 0.1), _rand((E,), "fp32", cuda, 3)
    pos = _rand((Hp, Hp, E), "fp32", cuda, 4)
    col = ops.empty(B * Hp * Hp, 320)
    ops.patch_im2col(x, col, B, S)
    patches = F.unfold(x, 7, padding=3, stride=4).transpose(1, 2).reshape(B * Hp * Hp, 147)
    hi = patches.bfloat16()
    assert torch.equal(col[:, :147], hi)
    assert torch.equal(col[:, 160:307], (patches - hi.float()).bfloat16())
    assert float(col[:, 147:160].abs().max()) == 0.0 and float(col[:, 307:].abs().max()) == 0.0
    w2 = torch.zeros(E, 320, device=cuda, dtype=torch.bfloat16)
    w2[:, :147] = w.reshape(E, 147).bfloat16()
    w2[:, 160:307] = w2[:, :147]
    rows = (pos.reshape(1, Hp * Hp, E) + b).expand(B, -1, -1).reshape(B * Hp * Hp, E).contiguous()
    ts, tc = torch.empty(B * Hp * Hp, E, device=cuda), ops.empty(B * Hp * Hp, E)
    ops.gemm(col, w2, ts, resid=rows, pre_out=tc, flags=4 | 16 | 32 | 64)
    ref = F.conv2d(x, w.bfloat16().float(), b, stride=4, padding=3).permute(0, 2, 3, 1) + pos
    _close(ts, ref.reshape(-1, E), 2e-4, "tensor-core stem (fp32 stream)")
    _close(tc, ref.reshape(-1, E), 1e-2, "tensor-core stem (bf16 copy)")


CONVS = [(64, 1, 3, 1), (64, 3, 1, 1), (64, 1, 7, 1), (64, 7, 1, 1), (64, 3, 3, 3), (64, 3, 3, 7), (256, 3, 3, 1),
         (128, 3, 3, 1)]


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("conv", CONVS)
def test_conv_im2col_gemm(cuda, dtype, conv):
    """conv forward, input gradient and weight gradient through im2col + GEMM vs F.conv2d autograd."""
    cin, kh, kw, dil = conv
    ops = _ops(dtype, cuda)
    B, H = 2, 11
    M, taps = B * H * H, kh * kw
    x = _rand((B, H, H, cin), dtype, cuda, 1)
    w = _rand((64, cin, kh, kw), "fp32", cuda, 2, (cin * taps) ** -0.5)
    wf, wd = ops.empty(64, taps * cin), ops.empty(cin, taps * 64)
    ops.conv_weight_pack(w, wf, wd, 64, cin, kh, kw)
    ph, pw = dil * (kh - 1) // 2, dil * (kw - 1) // 2
    col = ops.empty(M, taps * cin)
    ops.im2col(x, cin, col, B, H, H, cin, kh, kw, dil, dil, ph, pw)
    y = ops.empty(M, 64)
    ops.gemm(col, wf, y)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.to(DT[dtype]).float().requires_grad_(True)
    ref = F.conv2d(xr, wr, None, padding=(ph, pw), dilation=dil)
    _close(y.view(B, H, H, 64), ref.permute(0, 2, 3, 1), _tol(dtype), "conv fwd")
    dy = _rand((M, 64), dtype, cuda, 3)
    gx, gw = torch.autograd.grad(ref, (xr, wr), dy.float().view(B, H, H, 64).permute(0, 3, 1, 2))
    cold = ops.empty(M, taps * 64)
    ops.im2col(dy, 64, cold, B, H, H, 64, kh, kw, dil, dil, ph, pw)
    dx = ops.empty(M, cin)
    ops.gemm(cold, wd, dx)
    _close(dx.view(B, H, H, cin), gx.permute(0, 2, 3, 1), _tol(dtype), "conv dgrad")
    G = torch.zeros(64, cin, kh, kw, device=cuda)
    ops.wgrad(dy, col, G, ldg=taps * cin, q_inner=cin, q_taps=taps)
    _close(G, gw, _tol(dtype, 1e-4, 1e-2), "conv wgrad")


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("relu", [False, True])
@pytest.mark.parametrize("M", [2 * 11 * 11, 12 * 44 * 44 + 3, 12 * 88 * 88])
def test_batchnorm(cuda, dtype, relu, M):
    from sam2_unet_b200 import _lib
    ops = _ops(dtype, cuda)
    C = 64
    x = _rand((M, C), dtype, cuda, 1) * 1.5 + 0.3
    g, b = _rand((C,), "fp32", cuda, 2) * 0.1 + 1, _rand((C,), "fp32", cuda, 3) * 0.1
    rm, rv = torch.zeros(C, device=cuda), torch.ones(C, device=cuda)
    nbt = torch.zeros((), dtype=torch.long, device=cuda)
    sums = torch.zeros(_lib.load().s2u_bn_ws_doubles(C), dtype=torch.float64, device=cuda)
    scale, shift, mean, rstd = (torch.empty(C, device=cuda) for _ in range(4))
    if relu:                                              # both forms of the training forward
        ops.bn_stats(x, C, sums, M, C)
        ops.bn_finalize(sums, g, b, rm, rv, nbt, scale, shift, mean, rstd, M, C, True)
    else:
        ops.bn_stats_finalize(x, C, sums, g, b, rm, rv, nbt, scale, shift, mean, rstd, M, C)
    y = ops.empty(M, C)
    ops.bn_apply(x, C, scale, shift, None, 0, y, C, M, C, relu)
    xr = x.float().requires_grad_(True)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    rm2, rv2 = torch.zeros(C, device=cuda), torch.ones(C, device=cuda)
    ref = F.batch_norm(xr.t().reshape(1, C, M), rm2, rv2, gr, br, True, 0.1, 1e-5).reshape(C, M).t()
    if relu:
        ref = F.relu(ref)
    _close(y, ref, _tol(dtype, 1e-4, 2e-2), "bn fwd")
    _close(rm, rm2, 1e-4, "running mean")
    _close(rv, rv2, 1e-4, "running var")
    assert int(nbt) == 1 and float(sums.abs().max()) == 0.0
    dy = _rand((M, C), dtype, cuda, 4)
    dg, db = torch.zeros(C, device=cuda), torch.zeros(C, device=cuda)
    c1, c2 = torch.empty(C, device=cuda), torch.empty(C, device=cuda)
    dx = ops.empty(M, C)
    ops.bn_bwd(dy, C, y if relu else None, C, x, C, mean, rstd, g, sums, dg, db, c1, c2, dx, C, M, C)
    # the kernel masks with ITS OWN saved output; use the same mask in the reference to stay flip-free
    yr = F.batch_norm(xr.t().reshape(1, C, M), None, None, gr, br, True, 0.1, 1e-5).reshape(C, M).t()
    up = dy.float() * ((y.float() > 0).float() if relu else 1.0)
    gx, gg, gb = torch.autograd.grad(yr, (xr, gr, br), up)
    _close(dx, gx, _tol(dtype, 2e-4, 3e-2), "bn dx")
    _close(dg, gg, _tol(dtype, 2e-4, 2e-2), "bn dgamma")
    _close(db, gb, _tol(dtype, 2e-4, 2e-2), "bn dbeta")
    assert float(sums.abs().max()) == 0.0, "workspace (accumulators, barrier words) must be left zeroed"
    # eval mode uses the running statistics
    ops.bn_finalize(sums, g, b, rm, rv, nbt, scale, shift, None, None, M, C, False)
    ops.bn_apply(x, C, scale, shift, None, 0, y, C, M, C, False)
    ref = F.batch_norm(x.float().t().reshape(1, C, M), rm, rv, g, b, False, 0.1, 1e-5).reshape(C, M).t()
    _close(y, ref, _tol(dtype, 1e-4, 2e-2), "bn eval")


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_resample_and_heads(cuda, dtype):
    from sam2_unet_b200.resample_tables import ResampleTables
    ops = _ops(dtype, cuda)
    B, h, C = 2, 11, 64
    x = _rand((B, h, h, C), dtype, cuda, 1)
    tab = ResampleTables.get(h, 2 * h, True, None, cuda)
    out = ops.empty(B, 2 * h, 2 * h, 128)
    out.zero_()
    ops.resample_fwd(x, C, out.data_ptr() + 64 * out.element_size(), 128, B, h, 2 * h, C, tab)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=True)
    _close(out[..., 64:], ref.permute(0, 2, 3, 1), _tol(dtype, 1e-5, 1e-2), "upsample fwd")
    assert float(out[..., :64].abs().max()) == 0.0
    dout = _rand((B, 2 * h, 2 * h, 128), dtype, cuda, 2)
    dx = ops.empty(B, h, h, C)
    ops.resample_bwd(dout.data_ptr() + 64 * dout.element_size(), 128, dx, C, B, h, 2 * h, C, tab)
    (gx,) = torch.autograd.grad(ref, xr, dout[..., 64:].float().permute(0, 3, 1, 2))
    _close(dx, gx.permute(0, 2, 3, 1), _tol(dtype, 1e-5, 1e-2), "upsample bwd")
    # heads: 1x1 conv + x4/x8/x16 (align_corners=False)
    for scale in (4, 8, 16):
        M = B * h * h
        feat = _rand((M, 64), dtype, cuda, 3)
        w, b = _rand((1, 64, 1, 1), "fp32", cuda, 4, 0.2), _rand((1,), "fp32", cuda, 5)
        low = torch.empty(B, h, h, device=cuda)
        ops.head_fwd(feat, 64, w, b, low, M)
        S = h * scale
        t2 = ResampleTables.get(h, S, False, float(scale), cuda)
        full = torch.empty(B, 1, S, S, device=cuda)
        ops.resample1_fwd(low, full, B, h, S, t2)
        fr = feat.float().view(B, h, h, 64).permute(0, 3, 1, 2).requires_grad_(True)
        wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
        ref = F.interpolate(F.conv2d(fr, wr, br), scale_factor=scale, mode="bilinear")
        _close(full, ref, 1e-5, f"head x{scale} fwd")
        g = _rand((B, 1, S, S), "fp32", cuda, 6)
        dlow = torch.empty(B, h, h, device=cuda)
        ops.resample1_bwd(g, dlow, B, h, S, t2)
        dfeat = ops.empty(M, 64)
        dw, db = torch.zeros(64, device=cuda), torch.zeros(1, device=cuda)
        ops.head_bwd(feat, 64, w, dlow, dfeat, 64, False, dw, db, M)
        gf, gw, gb = torch.autograd.grad(ref, (fr, wr, br), g)
        _close(dfeat.view(B, h, h, 64), gf.permute(0, 2, 3, 1), _tol(dtype, 1e-4, 1e-2), "head dfeat")
        _close(dw, gw.view(-1), _tol(dtype, 1e-4, 1e-2), "head dw")
        _close(db, gb, 1e-4, "head db")


# ------------------------------------------------------------------------------------------ loss / optimizer

@pytest.mark.parametrize("shape", [(2, 96), (3, 352), (1, 64), (2, 50), (1, 130), (1, 132)])
def test_structure_loss(cuda, shape):
    from oracle import port
    from sam2_unet_b200.loss import structure_loss, structure_loss3
    B, S = shape
    _, mask = port.synthetic_batch(B, S, seed=3)
    preds = [(torch.randn(B, 1, S, S, generator=torch.Generator().manual_seed(i)) * 2).to(cuda).requires_grad_(True)
             for i in range(3)]
    mask = mask.to(cuda)
    ref = [port.structure_loss(p, mask) for p in preds]
    gref = torch.autograd.grad(sum(r * (i + 1) for i, r in enumerate(ref)), preds)
    l3 = structure_loss3(*preds, mask)
    got = torch.autograd.grad((l3 * torch.tensor([1.0, 2.0, 3.0], device=cuda)).sum(), preds)
    for i in range(3):
        assert abs(l3[i].item() - ref[i].item()) <= 1e-5 * max(1.0, abs(ref[i].item()))
        _close(got[i], gref[i], 1e-4, f"loss grad {i}")
    one = structure_loss(preds[0], mask)
    assert abs(one.item() - ref[0].item()) <= 1e-5 * max(1.0, abs(ref[0].item()))
    (g1,) = torch.autograd.grad(one, preds[0])
    _close(g1, gref[0], 1e-4, "single-head grad")


def test_adamw_kernel(cuda):
    from sam2_unet_b200 import _lib
    n = 10007
    p = torch.randn(n, device=cuda)
    ref_p = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([ref_p], lr=1e-3, weight_decay=5e-4)
    m, v = torch.zeros(n, device=cuda), torch.zeros(n, device=cuda)
    hyper = torch.tensor([1e-3, 1.0, 1.0, 1.0], device=cuda)
    st = torch.cuda.current_stream().cuda_stream
    for step in range(5):
        g = torch.randn(n, device=cuda, generator=torch.Generator(device="cuda").manual_seed(step))
        ref_p.grad = g.clone()
        opt.step()
        _lib.call("s2u_adamw", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, hyper.data_ptr(), 0.9, 0.999,
                  1e-8, 5e-4, st)
    _close(p, ref_p.detach(), 1e-5, "adamw")


# ------------------------------------------------------------------------------------------ inference tail

@pytest.mark.parametrize("case", [(352, (0, 0, 0, 0), (352, 352)), (352, (0, 0, 0, 117), (480, 720)),
                                  (352, (30, 0, 31, 0), (1001, 640)), (96, (5, 7, 0, 3), (50, 41)),
                                  (1024, (0, 0, 256, 0), (1536, 1152))])
def test_infer_tail(cuda, case):
    """Device-side crop + bilinear resize + sigmoid + min-max + uint8 (test.py:66-76) against the reference's own
    operators on the CPU (oracle.port.infer_tail).  Byte work: exact up to the last-ulp difference between the CPU's
    and the GPU's expf, which can move a value across a quantisation step - at most one level, on a handful of pixels."""
    from oracle import port
    from sam2_unet_b200 import infer_tail
    S, padding, hw = case
    g = torch.Generator(device="cpu").manual_seed(S + hw[0])
    logits = torch.randn(1, 1, S // 8, S // 8, generator=g) * 3
    logits = F.interpolate(logits, size=(S, S), mode="bicubic", align_corners=False).contiguous()   # smooth map
    ref = torch.from_numpy(port.infer_tail(logits, padding, hw))
    for _ in range(2):                                    # twice: the min / max workspace must come back re-armed
        got = infer_tail(logits.to(cuda), padding, hw).cpu()
    assert got.dtype == torch.uint8 and tuple(got.shape) == tuple(hw)
    diff = (got.int() - ref.int()).abs()
    assert int(diff.max()) <= 1, int(diff.max())
    assert float((diff > 0).float().mean()) <= 2e-3, float((diff > 0).float().mean())
    assert int(got.min()) == 0 and int(got.max()) >= 254
    with pytest.raises(Exception):
        infer_tail(logits, padding, hw)                   # CPU tensor: no fallback


# ------------------------------------------------------------------------------------------ eval metrics

def _blobs(h, w, n, seed, jitter=0):
    import numpy as np
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    m = np.zeros((h, w), np.uint8)
    for _ in range(n):
        cy, cx, r = rng.integers(0, h), rng.integers(0, w), rng.integers(2, max(3, min(h, w) // 6))
        cy, cx = cy + rng.integers(-jitter, jitter + 1), cx + rng.integers(-jitter, jitter + 1)
        m[(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = rng.integers(26, 256)
    return m


@pytest.mark.parametrize("case", [(64, 96, 6, 0), (353, 211, 25, 1), (480, 720, 60, 2), (32, 32, 0, 3), (128, 128, 300, 4)])
def test_eval_metrics_vs_oracle(cuda, case):
    """Device thresholding / counting / 8-connected labelling / overlap table + host matching against the CPU
    restatement of eval.py:55-171: integer work, so every entry of the result dictionary must be identical."""
    import numpy as np
    from oracle import eval_port
    from sam2_unet_b200 import evaluate_dataset, evaluate_segmentation_performance
    h, w, n, seed = case
    gt = _blobs(h, w, n, seed)
    pred = _blobs(h, w, n, seed, jitter=3)                     # same blobs, displaced: partial overlaps, merges, splits
    rng = np.random.default_rng(seed + 100)
    pred[rng.random((h, w)) < 0.002] = 200                     # salt: single-pixel components, diagonal contacts
    pred[rng.random((h, w)) < 0.01] = 20                       # below the 25.5 threshold: must not count
    results = []
    for p, g in ((pred, gt), (gt, gt), (np.zeros_like(gt), gt), (pred, np.zeros_like(gt))):
        ref = eval_port.evaluate_segmentation_performance(p, g)
        got = evaluate_segmentation_performance(torch.from_numpy(p).to(cuda), torch.from_numpy(g).to(cuda))
        assert set(got) == set(ref)
        for k in ref:
            assert got[k] == ref[k], (k, got[k], ref[k])
        results.append(got)
    agg, agg_ref = evaluate_dataset(results), eval_port.evaluate_dataset(results)      # eval.py:170-224
    assert list(agg) == list(agg_ref)
    for k in agg_ref:
        assert float(agg[k]) == float(agg_ref[k]), k
    with pytest.raises(Exception):
        evaluate_segmentation_performance(torch.from_numpy(pred), torch.from_numpy(gt))      # CPU tensors: no fallback


# ------------------------------------------------------------------------------------------ test-time input path

@pytest.mark.parametrize("case", [(300, 420, 352), (1080, 1920, 352), (352, 352, 352), (97, 41, 96), (200, 333, 1024),
                                  (2000, 1500, 1024)])
def test_preprocess_image(cuda, case):
    """uint8 HWC image -> /255 -> antialiased bilinear resize (longest side) -> centred zero pad -> normalise, against
    the reference's own torchvision transforms on the CPU (oracle.port.preprocess_image, dataset.py:336-407)."""
    import numpy as np
    from oracle import port
    from sam2_unet_b200 import preprocess_image
    H, W, S = case
    rng = np.random.default_rng(H + W)
    base = rng.integers(0, 256, (H // 8 + 2, W // 8 + 2, 3)).astype(np.float32)
    img = np.kron(base, np.ones((8, 8, 1), np.float32))[:H, :W]                    # blocky image with sharp edges
    img = np.clip(img + rng.normal(0, 12, img.shape), 0, 255).astype(np.uint8)
    ref, pad_ref = port.preprocess_image(img, S)
    got, pad = preprocess_image(torch.from_numpy(img).to(cuda), S)
    assert list(pad) == list(pad_ref) and tuple(got.shape) == (1, 3, S, S)
    err = (got[0].cpu() - ref).abs().max().item()
    assert err <= 2e-5, err                                                        # fp32 path, values in [-2.2, 2.7]
    with pytest.raises(Exception):
        preprocess_image(torch.from_numpy(img), S)                                 # CPU tensor: no fallback


# ------------------------------------------------------------------------------------------ training-time input path

def test_train_augment_against_reference_vectors(cuda):
    """TrainAugment (csrc/augment.cu) on the reference's own outputs: tests/golden/augment_train.npz was written by the
    unmodified FullDataset(mode="train") transform (dataset.py:288-313) with `random.seed(seed)`; the device pipeline
    draws from `random` in the same order, so the same seed must give the same sample.  Labels bit-exact; image within
    3e-5 (fp32, values in [-2.2, 2.7]; pow / division orders of the colour ops differ in the last bit)."""
    import os
    import random
    import numpy as np
    from oracle.make_golden_augment import inputs
    from sam2_unet_b200 import TrainAugment
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "augment_train.npz"))
    S = int(g["size"])
    aug = TrainAugment(S, cuda)
    for seed in g["seeds"].tolist():
        img, lab = inputs(seed)
        random.seed(seed)
        out = aug(img, lab)
        assert torch.equal(out["label"].cpu(), torch.from_numpy(g[f"label_{seed}"])), seed
        err = (out["image"].cpu() - torch.from_numpy(g[f"image_{seed}"])).abs().max().item()
        assert err <= 3e-5, (seed, err)
    with pytest.raises(Exception):
        TrainAugment(S, "cpu")                                                     # no CPU fallback


@pytest.mark.parametrize("size,hw", [(352, (300, 420)), (352, (1080, 1920)), (96, (97, 41)), (1024, (700, 500))])
def test_train_augment_against_oracle(cuda, size, hw):
    """30 seeded samples per shape against oracle/augment_port.py (pinned to the live reference by tests/test_cpu.py), at
    the network sizes; plus every colour op / blur size / rotation forced once through explicit parameters."""
    import random
    import numpy as np
    from oracle import augment_port as ap
    from sam2_unet_b200 import TrainAugment
    H, W = hw
    rng = np.random.default_rng(H * 7 + W)
    base = rng.integers(0, 256, (H // 8 + 2, W // 8 + 2, 3)).astype(np.float32)
    img = np.clip(np.kron(base, np.ones((8, 8, 1), np.float32))[:H, :W] + rng.normal(0, 12, (H, W, 3)), 0, 255).astype(np.uint8)
    lab = (np.kron(rng.random((H // 8 + 2, W // 8 + 2)), np.ones((8, 8)))[:H, :W] > 0.5).astype(np.uint8) * 255
    aug = TrainAugment(size, cuda)
    img_d, lab_d = torch.from_numpy(img).to(cuda), torch.from_numpy(lab).to(cuda)
    n = 30 if size <= 352 else 6
    plist = []
    for seed in range(n):
        random.seed(seed)
        plist.append(ap.draw(size, H, W))
    forced = [("brightness", 0.7), ("contrast", 1.3), ("saturation", 0.6), ("hue", -0.37), ("hue", 0.5), ("gamma", 1.4)]
    for i, op in enumerate(forced):
        plist.append({"geom": ("crop", 3, 5, H - 7, W - 9) if i % 2 else ("pad", 11, 4, H + 30, W + 17), "rot": i % 4,
                      "gray": False, "color": [op], "blur": (0, 3, 5)[i % 3]})
    for p in plist:
        ref = ap.apply(p, torch.from_numpy(img), torch.from_numpy(lab), size)
        q = dict(p, geom=(0 if p["geom"][0] == "pad" else 1,) + tuple(p["geom"][1:]))
        out = aug(img_d, lab_d, params=q)
        assert torch.equal(out["label"].cpu(), ref["label"]), p
        err = (out["image"].cpu() - ref["image"]).abs().max().item()
        assert err <= 3e-5, (p, err)


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("shape", [(5808, 576, 32), (1452, 1152, 32), (3000, 144, 32), (777, 64, 64)])
def test_wgrad_pair(cuda, dtype, shape):
    """The two adapter weight gradients of a block in one launch: dW2 [C,32] += dh2^T u, dW1 [32,C] += dh1^T x."""
    M, C, r = shape
    ops = _ops(dtype, cuda)
    dh2, u = _rand((M, C), dtype, cuda, 1), _rand((M, r), dtype, cuda, 2)
    dh1, x = _rand((M, r), dtype, cuda, 3), _rand((M, C), dtype, cuda, 4)
    G2, G1 = torch.zeros(C, r, device=cuda), torch.ones(r, C, device=cuda)       # accumulates into what is there
    ops.wgrad_pair(dh2, u, G2, r, dh1, x, G1, C)
    tol = 1e-4 if dtype == "fp32" else 2e-3
    _close(G2, dh2.float().t() @ u.float(), tol, "pair: dW2")
    _close(G1, 1 + dh1.float().t() @ x.float(), tol, "pair: dW1")


# ------------------------------------------------------------------------- fused adapter + LayerNorm (bf16 mode)

def _dgelu(x):
    return 0.5 * (1 + torch.erf(x * 0.7071067811865476)) + x * torch.exp(-0.5 * x * x) * 0.3989422804014327


@pytest.mark.parametrize("shape", [(5808, 576), (1000, 144), (23232 + 5, 288), (1452, 1152), (333, 32), (100, 64),
                                   (77, 128), (50, 256), (500, 96), (301, 448), (17, 768)])
def test_adapter_ln_fused(cuda, shape):
    """s2u_adapter_ln_fwd / _bwd (SAM2UNet.py:57-63 + hieradet.py:134) against the same math in fp32 torch."""
    from sam2_unet_b200 import _lib
    R, C = shape
    ops = _ops("bf16", cuda)
    assert ops.adapter_supported(C)
    xs = _rand((R, C), "fp32", cuda, 1, 2.0) + 0.3
    W1, W2 = _rand((32, C), "bf16", cuda, 2, C ** -0.5), _rand((C, 32), "bf16", cuda, 3, 32 ** -0.5)
    b1, b2 = _rand((32,), "fp32", cuda, 4, 0.5), _rand((C,), "fp32", cuda, 5, 0.5)
    gamma, beta = _rand((C,), "fp32", cuda, 6) + 1.0, _rand((C,), "fp32", cuda, 7)
    f32 = torch.float32
    xa, n1 = ops.empty(R, C, dtype=f32), ops.empty(R, C)
    mean, rstd = ops.empty(R, dtype=f32), ops.empty(R, dtype=f32)
    u, g1, g2 = ops.empty(R, 32), ops.empty(R, 32), ops.empty(R, C)
    ops.adapter_ln_fwd(xs, W1, b1, W2, b2, gamma, beta, xa, n1, mean, rstd, u, g1, g2, R, C)
    h = xs.bfloat16().float() @ W1.float().t() + b1
    u_ref = F.gelu(h)
    pre2 = u_ref.bfloat16().float() @ W2.float().t() + b2
    xa_ref = xs + F.gelu(pre2)
    _close(u, u_ref, 2e-2, "u")
    _close(g1, _dgelu(h), 2e-2, "gelu'(h1)")
    _close(g2, _dgelu(pre2), 2e-2, "gelu'(h2)")
    _close(xa, xa_ref, 5e-3, "xa")
    mu = xa_ref.mean(1)
    var = xa_ref.var(1, unbiased=False)
    _close(mean, mu, 5e-3, "mean")
    _close(rstd, torch.rsqrt(var + 1e-6), 5e-3, "rstd")
    _close(n1, F.layer_norm(xa_ref, (C,), gamma, beta, 1e-6), 2e-2, "n1")
    # inference variant: nothing saved, same outputs
    xa2, n2 = ops.empty(R, C, dtype=f32), ops.empty(R, C)
    ops.adapter_ln_fwd(xs, W1, b1, W2, b2, gamma, beta, xa2, n2, mean, rstd, None, None, None, R, C)
    assert torch.equal(xa2, xa) and torch.equal(n2, n1)

    # backward, from the tensors the forward saved
    dn1, dres = _rand((R, C), "bf16", cuda, 8), _rand((R, C), "bf16", cuda, 9)
    W2t, W1t = W2.t().contiguous(), W1.t().contiguous()
    for with_res in (True, False):
        dh2, dh1, dx = ops.empty(R, C), ops.empty(R, 32), ops.empty(R, C)
        db1 = torch.full((32,), 0.5, dtype=f32, device=cuda)
        db2 = torch.full((C,), -0.25, dtype=f32, device=cuda)
        ops.adapter_ln_bwd(dn1, xa, mean, rstd, gamma, dres if with_res else None, g2, g1, W2t, W1t, dh2, dh1, dx, db1,
                           db2, R, C)
        xh = (xa - mean[:, None]) * rstd[:, None]
        gd = dn1.float() * gamma
        dxa = rstd[:, None] * (gd - gd.mean(1, keepdim=True) - xh * (gd * xh).mean(1, keepdim=True))
        if with_res:
            dxa = dxa + dres.float()
        dh2_ref = dxa * g2.float()
        dh1_ref = (dh2_ref.bfloat16().float() @ W2.float()) * g1.float()
        dx_ref = dxa + dh1_ref.bfloat16().float() @ W1.float()
        _close(dh2, dh2_ref, 2e-2, "dh2")
        _close(dh1, dh1_ref, 2e-2, "dh1")
        _close(dx, dx_ref, 2e-2, "dx")
        _close(db1 - 0.5, dh1_ref.sum(0), 2e-3, "db1")
        _close(db2 + 0.25, dh2_ref.sum(0), 2e-3, "db2")


# ---------------------------------------------------------------------- implicit-GEMM convolution (bf16, no im2col)

@pytest.mark.parametrize("geom", [(2, 11), (3, 22), (1, 44), (2, 88), (1, 19)])
@pytest.mark.parametrize("conv", CONVS + [(64, 1, 5, 1), (64, 5, 1, 1), (64, 3, 3, 5)])
def test_conv_igemm(cuda, conv, geom):
    """s2u_conv_igemm (forward with bias / residual / ReLU and the BatchNorm statistics epilogue, input gradient with
    accumulation, strided input / output maps) and s2u_conv_wgrad vs F.conv2d autograd (SAM2UNet.py:83-86)."""
    cin, kh, kw, dil = conv
    B, H = geom
    ops = _ops("bf16", cuda)
    M, taps = B * H * H, kh * kw
    ldx = cin + 64                                                   # the input is a channel slice of a wider map
    xbuf = _rand((M, ldx), "bf16", cuda, 1)
    x = xbuf[:, 64:]
    w = _rand((64, cin, kh, kw), "fp32", cuda, 2, (cin * taps) ** -0.5)
    wf, wd = ops.empty(64, taps * cin), ops.empty(cin, taps * 64)
    ops.conv_weight_pack(w, wf, wd, 64, cin, kh, kw)
    ph, pw = dil * (kh - 1) // 2, dil * (kw - 1) // 2
    xr = x.float().reshape(B, H, H, cin).permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.bfloat16().float().requires_grad_(True)
    ref = F.conv2d(xr, wr, None, padding=(ph, pw), dilation=dil)
    refr = ref.permute(0, 2, 3, 1).reshape(M, 64)
    assert ops.conv_igemm_ok(cin, 64, ldx, 64)
    # forward + statistics
    from sam2_unet_b200 import _lib
    sums = torch.zeros(_lib.load().s2u_bn_ws_doubles(64), dtype=torch.float64, device=cuda)
    y = ops.empty(M, 64)
    ops.conv_igemm(xbuf.view(-1)[64:], ldx, B, H, H, cin, wf, 64, kh, kw, dil, y, 64, sums=sums)
    _close(y, refr, 2e-2, "conv fwd")
    rep = sums[:16 * 128].view(16, 128).sum(0)
    yf = y.double()
    assert torch.allclose(rep[:64], yf.sum(0), rtol=1e-5, atol=1e-3), "BatchNorm sums"
    assert torch.allclose(rep[64:], (yf * yf).sum(0), rtol=1e-5, atol=1e-3), "BatchNorm sums of squares"
    # the same with the finalisation fused (last CTA): scale / shift / saved statistics / running statistics must equal
    # the standalone statistics pass over the same output
    gmm, bta = _rand((64,), "fp32", cuda, 11) * 0.1 + 1, _rand((64,), "fp32", cuda, 12) * 0.1
    outs = {}
    for fused in (True, False):
        rm, rv = torch.zeros(64, device=cuda), torch.ones(64, device=cuda)
        nbt = torch.zeros((), dtype=torch.long, device=cuda)
        sc, sh_, mu, rs = (torch.empty(64, device=cuda) for _ in range(4))
        sums.zero_()
        if fused:
            y2 = ops.empty(M, 64)
            ops.conv_igemm_bn(xbuf.view(-1)[64:], ldx, B, H, H, cin, wf, kh, kw, dil, y2, 64, sums, gmm, bta, rm, rv, nbt,
                              sc, sh_, mu, rs)
            assert torch.equal(y2, y)
        else:
            ops.bn_stats_finalize(y, 64, sums, gmm, bta, rm, rv, nbt, sc, sh_, mu, rs, M, 64)
        assert float(sums.abs().max()) == 0.0 and int(nbt) == 1
        outs[fused] = (sc, sh_, mu, rs, rm, rv)
    for a, b_, nm in zip(outs[True], outs[False], ("scale", "shift", "mean", "rstd", "running_mean", "running_var")):
        assert torch.allclose(a, b_, rtol=1e-4, atol=1e-5), nm
    # bias + residual + ReLU into a strided output
    bias, res = _rand((64,), "fp32", cuda, 5), _rand((M, 64), "bf16", cuda, 6)
    obuf = torch.zeros(M, 256, dtype=torch.bfloat16, device=cuda)
    ops.conv_igemm(xbuf.view(-1)[64:], ldx, B, H, H, cin, wf, 64, kh, kw, dil, obuf.view(-1)[128:], 256, bias=bias,
                   resid=res, ld_res=64, relu=True)
    _close(obuf[:, 128:192], torch.relu(refr + bias + res.float()), 2e-2, "conv fwd epilogue")
    assert float(obuf[:, :128].abs().max()) == 0.0 and float(obuf[:, 192:].abs().max()) == 0.0
    # input gradient (accumulating) and weight gradient
    dy = _rand((M, 64), "bf16", cuda, 3)
    gx, gw = torch.autograd.grad(ref, (xr, wr), dy.float().view(B, H, H, 64).permute(0, 3, 1, 2))
    gxr = gx.permute(0, 2, 3, 1).reshape(M, cin)
    assert ops.conv_igemm_ok(64, cin, 64, cin)
    dx = ops.empty(M, cin)
    ops.conv_igemm(dy, 64, B, H, H, 64, wd, cin, kh, kw, dil, dx, cin)
    _close(dx, gxr, 2e-2, "conv dgrad")
    base = _rand((M, cin), "bf16", cuda, 7)
    dx2 = base.clone()
    ops.conv_igemm(dy, 64, B, H, H, 64, wd, cin, kh, kw, dil, dx2, cin, resid=dx2, ld_res=cin)
    _close(dx2, gxr + base.float(), 2e-2, "conv dgrad accumulate")
    G = torch.full((64, cin, kh, kw), 0.25, device=cuda)
    ops.conv_wgrad(dy, 64, xbuf.view(-1)[64:], ldx, G, B, H, H, cin, 64, kh, kw, dil)
    _close(G - 0.25, gw, 1e-2, "conv wgrad")

"""Forward / backward schedule of SAM2-UNet on the C-ABI kernels.

The engine walks the reference's data flow (/root/reference/SAM2UNet.py:164-173 and
sam2/modeling/backbones/hieradet.py:278-292,132-167) as an explicit list of kernel launches, keeps the
tensors the backward needs on a tape, and runs the hand-derived backward of every stage.  PyTorch is used
for device memory and streams only: every arithmetic operation on the path is a kernel of
libsam2unet_b200.so (no ATen math, no cuBLAS/cuDNN, no CPU fallback).

Data layout: NHWC tokens [B*H*W, C] in the compute dtype T (fp32 or bf16) everywhere inside the model;
NCHW only at the public edges (the input image, the three [B,1,S,S] fp32 logit maps).
Frozen trunk weights are held once in T in both orientations (W for the forward, W^T for the input
gradient: the trunk needs no weight gradient, SAM2UNet.py:146-147).  Trainable parameters live in one flat
fp32 master buffer (the nn.Parameters are views of it, state-dict layout unchanged) with a matching flat
gradient buffer, so AdamW and the data-parallel all-reduce are single flat operations.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from . import _lib
from .config import TrunkConfig
from .resample_tables import ResampleTables

GELU, DGELU, RESID, OUT_F32, RESID_F32, PRE_FINAL, SAVE_DGELU, MULAUX, RELU = 1, 2, 4, 16, 32, 64, 128, 256, 512


def _ptr(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


class Ops:
    """Thin typed wrappers over the C ABI; `dt` selects fp32 (0) or bf16 (1) activations."""

    def __init__(self, dtype: torch.dtype, device: torch.device, gemm_backend: int = 0):
        if dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("dtype must be torch.float32 or torch.bfloat16")
        self.T = dtype
        self.dt = 0 if dtype == torch.float32 else 1
        self.device = device
        self.gemm_backend = gemm_backend
        self._ln_ws = {}
        _lib.load()

    @property
    def stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def empty(self, *shape, dtype=None) -> torch.Tensor:
        return torch.empty(shape, dtype=dtype or self.T, device=self.device)

    # C[M,N] = epi(A[M,K] W[N,K]^T)
    def gemm(self, A, W, C, bias=None, pre_out=None, aux=None, resid=None, flags=0, M=None, N=None, K=None,
             lda=None, ldw=None, ldc=None, ld_pre=None, ld_aux=None, ld_res=None, backend=None):
        M = A.shape[0] if M is None else M
        K = A.shape[1] if K is None else K
        N = W.shape[0] if N is None else N
        _lib.call("s2u_gemm", A.data_ptr(), lda or A.stride(0), W.data_ptr(), ldw or W.stride(0), C.data_ptr(),
                  ldc or C.stride(0), M, N, K, _ptr(bias), _ptr(pre_out),
                  (ld_pre or pre_out.stride(0)) if pre_out is not None else 0, _ptr(aux),
                  (ld_aux or aux.stride(0)) if aux is not None else 0, _ptr(resid),
                  (ld_res or resid.stride(0)) if resid is not None else 0, flags, self.dt,
                  self.gemm_backend if backend is None else backend, self.stream)
        return C

    def wgrad(self, A, B, G, M=None, P=None, Q=None, lda=None, ldb=None, ldg=None, q_inner=0, q_taps=0):
        M = A.shape[0] if M is None else M
        P = A.shape[1] if P is None else P
        Q = B.shape[1] if Q is None else Q
        _lib.call("s2u_gemm_wgrad", A.data_ptr(), lda or A.stride(0), B.data_ptr(), ldb or B.stride(0), G.data_ptr(),
                  ldg if ldg is not None else Q, M, P, Q, q_inner, q_taps, self.dt, self.stream)

    def wgrad_pair(self, A0, B0, G0, ldg0, A1, B1, G1, ldg1):
        """G0 += A0^T B0 and G1 += A1^T B1 (same row count) in one launch."""
        _lib.call("s2u_gemm_wgrad_pair", A0.data_ptr(), A0.stride(0), B0.data_ptr(), B0.stride(0), G0.data_ptr(), ldg0,
                  A0.shape[1], B0.shape[1], A1.data_ptr(), A1.stride(0), B1.data_ptr(), B1.stride(0), G1.data_ptr(),
                  ldg1, A1.shape[1], B1.shape[1], A0.shape[0], self.dt, self.stream)

    def colsum(self, A, out, M=None, P=None, lda=None):
        _lib.call("s2u_colsum", A.data_ptr(), lda or A.stride(0), out.data_ptr(), A.shape[0] if M is None else M,
                  A.shape[1] if P is None else P, self.dt, self.stream)

    def ln_fwd(self, x, gamma, beta, y, mean, rstd, R, C):
        _lib.call("s2u_layernorm_fwd", x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), _ptr(mean),
                  _ptr(rstd), R, C, 1e-6, 1 if x.dtype == torch.float32 else 0, self.dt, self.stream)

    def ln_bwd(self, dy, x, gamma, mean, rstd, dres, dx, R, C, pre=None, dx2=None, colsum=None, pre_is_grad=False):
        ws = None
        if colsum is not None:
            # replicated column-sum accumulators, zeroed once and left zeroed by the kernel; one per (C, stream)
            key = (C, self.stream)
            ws = self._ln_ws.get(key)
            if ws is None:
                n = _lib.load().s2u_layernorm_ws_floats(C)
                ws = self._ln_ws[key] = torch.zeros(n, dtype=torch.float32, device=self.device)
        _lib.call("s2u_layernorm_bwd", dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), mean.data_ptr(),
                  rstd.data_ptr(), _ptr(dres), dx.data_ptr(), _ptr(pre), _ptr(dx2), _ptr(colsum), _ptr(ws),
                  1 if pre_is_grad else 0, R, C, 1 if x.dtype == torch.float32 else 0, self.dt, self.stream)

    def adapter_supported(self, C: int) -> bool:
        return self.T == torch.bfloat16 and bool(_lib.load().s2u_adapter_supported(C))

    def adapter_ln_fwd(self, xs, w1, b1, w2, b2, gamma, beta, xa, n1, mean, rstd, u, g1, g2, R, C):
        _lib.call("s2u_adapter_ln_fwd", xs.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                  gamma.data_ptr(), beta.data_ptr(), 1e-6, xa.data_ptr(), n1.data_ptr(), mean.data_ptr(),
                  rstd.data_ptr(), _ptr(u), _ptr(g1), _ptr(g2), R, C, self.stream)

    def adapter_ln_bwd(self, dn1, xa, mean, rstd, gamma, dres, g2, g1, w2t, w1t, dh2, dh1, dx, db1, db2, R, C):
        _lib.call("s2u_adapter_ln_bwd", dn1.data_ptr(), xa.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                  gamma.data_ptr(), _ptr(dres), g2.data_ptr(), g1.data_ptr(), w2t.data_ptr(), w1t.data_ptr(),
                  dh2.data_ptr(), dh1.data_ptr(), dx.data_ptr(), db1.data_ptr(), db2.data_ptr(), R, C, self.stream)

    def dgelu_mul(self, dy, pre, out):
        _lib.call("s2u_dgelu_mul", dy.data_ptr(), pre.data_ptr(), out.data_ptr(), dy.numel(), self.dt, self.stream)

    def add(self, a, b, out):
        _lib.call("s2u_add", a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), self.dt, self.stream)

    def maxpool_fwd(self, x, out, B, H, W, C):
        _lib.call("s2u_maxpool2_fwd", x.data_ptr(), out.data_ptr(), B, H, W, C, self.dt, self.stream)

    def maxpool_bwd(self, x, dout, dx, B, H, W, C):
        _lib.call("s2u_maxpool2_bwd", x.data_ptr(), dout.data_ptr(), dx.data_ptr(), B, H, W, C, self.dt, self.stream)

    def cast(self, src, dst, R, C, transpose=False):
        _lib.call("s2u_cast", src.data_ptr(), dst.data_ptr(), R, C, 1 if transpose else 0, self.dt, self.stream)

    def attn_fwd(self, qkv, bias, out, lse, B, H, W, nh, hd, window, pool):
        _lib.call("s2u_win_attn_fwd", qkv.data_ptr(), bias.data_ptr(), out.data_ptr(), lse.data_ptr(), B, H, W, nh,
                  hd, window, 1 if pool else 0, self.dt, self.stream)

    def attn_bwd(self, qkv, bias, out, lse, dout, dqkv, B, H, W, nh, hd, window, pool):
        dws = torch.empty_like(lse)
        _lib.call("s2u_win_attn_bwd", qkv.data_ptr(), bias.data_ptr(), out.data_ptr(), lse.data_ptr(),
                  dout.data_ptr(), dqkv.data_ptr(), dws.data_ptr(), B, H, W, nh, hd, window, 1 if pool else 0,
                  self.dt, self.stream)

    def patch_embed(self, x, w, bias, pos, out, B, S, E, out_copy=None):
        _lib.call("s2u_patch_embed", x.data_ptr(), w.data_ptr(), bias.data_ptr(), pos.data_ptr(), out.data_ptr(),
                  1 if out.dtype == torch.float32 else 0, _ptr(out_copy), B, S, E, self.dt, self.stream)

    def patch_im2col(self, x, out, B, S):
        _lib.call("s2u_patch_im2col", x.data_ptr(), out.data_ptr(), B, S, self.stream)

    def im2col(self, x, ldx, out, B, H, W, Cin, KH, KW, dh, dw, ph, pw):
        _lib.call("s2u_im2col", x.data_ptr(), ldx, out.data_ptr(), B, H, W, Cin, KH, KW, dh, dw, ph, pw, self.dt,
                  self.stream)

    def conv_igemm_ok(self, Cin, N, ldx, ld_out) -> bool:
        return self.T == torch.bfloat16 and bool(_lib.load().s2u_conv_igemm_supported(Cin, N, ldx, ld_out))

    def conv_igemm(self, x, ldx, B, H, W, Cin, wm, N, KH, KW, dil, out, ld_out, bias=None, resid=None, ld_res=0,
                   relu=False, sums=None):
        """out = epi(conv(x, wm)) without an im2col matrix; x / out / resid: tensors or raw device pointers."""
        xp = x.data_ptr() if isinstance(x, torch.Tensor) else x
        op = out.data_ptr() if isinstance(out, torch.Tensor) else out
        rp = 0 if resid is None else (resid.data_ptr() if isinstance(resid, torch.Tensor) else resid)
        _lib.call("s2u_conv_igemm", xp, ldx, B, H, W, Cin, wm.data_ptr(), N, KH, KW, dil, op, ld_out, _ptr(bias), rp,
                  ld_res, 1 if relu else 0, _ptr(sums), self.stream)

    def conv_igemm_bn(self, x, ldx, B, H, W, Cin, wm, KH, KW, dil, out, ld_out, sums, gamma, beta, rm, rv, nbt, scale,
                      shift, mean, rstd):
        """training forward: conv + BatchNorm batch statistics + finalisation in one launch"""
        xp = x.data_ptr() if isinstance(x, torch.Tensor) else x
        _lib.call("s2u_conv_igemm_bn", xp, ldx, B, H, W, Cin, wm.data_ptr(), KH, KW, dil, out.data_ptr(), ld_out,
                  sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(), rm.data_ptr(), rv.data_ptr(), _ptr(nbt),
                  scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), rstd.data_ptr(), 1e-5, 0.1, self.stream)

    def conv_wgrad(self, dy, ld_dy, x, ldx, G, B, H, W, Cin, Cout, KH, KW, dil):
        xp = x.data_ptr() if isinstance(x, torch.Tensor) else x
        _lib.call("s2u_conv_wgrad", dy.data_ptr(), ld_dy, xp, ldx, G.data_ptr(), B, H, W, Cin, Cout, KH, KW, dil,
                  self.stream)

    def conv_weight_pack(self, w, wf, wd, Cout, Cin, KH, KW):
        _lib.call("s2u_conv_weight_pack", w.data_ptr(), _ptr(wf), _ptr(wd), Cout, Cin, KH, KW, self.dt, self.stream)

    def bn_stats(self, x, ldx, sums, M, C):
        _lib.call("s2u_bn_stats", x.data_ptr(), ldx, sums.data_ptr(), M, C, self.dt, self.stream)

    def bn_stats_finalize(self, x, ldx, sums, gamma, beta, rm, rv, nbt, scale, shift, mean, rstd, M, C):
        _lib.call("s2u_bn_stats_finalize", x.data_ptr(), ldx, sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                  rm.data_ptr(), rv.data_ptr(), _ptr(nbt), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(),
                  rstd.data_ptr(), M, C, 1e-5, 0.1, self.dt, self.stream)

    def bn_finalize(self, sums, gamma, beta, rm, rv, nbt, scale, shift, mean, rstd, M, C, training):
        _lib.call("s2u_bn_finalize", sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(), rm.data_ptr(), rv.data_ptr(),
                  _ptr(nbt), scale.data_ptr(), shift.data_ptr(), _ptr(mean), _ptr(rstd), M, C, 1e-5, 0.1,
                  1 if training else 0, self.stream)

    def bn_apply(self, x, ldx, scale, shift, resid, ld_res, out, ld_out, M, C, relu):
        _lib.call("s2u_bn_apply", x.data_ptr() if isinstance(x, torch.Tensor) else x, ldx,
                  scale.data_ptr() if isinstance(scale, torch.Tensor) else scale,
                  shift.data_ptr() if isinstance(shift, torch.Tensor) else shift, _ptr(resid), ld_res,
                  out.data_ptr() if isinstance(out, torch.Tensor) else out, ld_out, M, C, 1 if relu else 0, self.dt,
                  self.stream)

    def relu_bwd(self, dy, ld_dy, y, ld_y, g, ld_g, M, C):
        _lib.call("s2u_relu_bwd", dy.data_ptr(), ld_dy, y.data_ptr(), ld_y, g.data_ptr(), ld_g, M, C, self.dt,
                  self.stream)

    def bn_bwd(self, dy, ld_dy, y, ld_y, x, ldx, mean, rstd, gamma, sums, dgamma, dbeta, c1, c2, dx, ld_dx, M, C):
        _lib.call("s2u_bn_bwd", dy.data_ptr(), ld_dy, _ptr(y), ld_y, x.data_ptr(), ldx, mean.data_ptr(),
                  rstd.data_ptr(), gamma.data_ptr(), sums.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(),
                  c1.data_ptr(), c2.data_ptr(), dx.data_ptr(), ld_dx, M, C, self.dt, self.stream)

    def resample_fwd(self, x, ldx, out_ptr, ld_out, B, Hi, Ho, C, tab: ResampleTables):
        f = tab.fwd_ptrs()
        _lib.call("s2u_resample_fwd", x.data_ptr(), ldx, out_ptr, ld_out, B, Hi, Hi, Ho, Ho, C, *f, *f, self.dt,
                  self.stream)

    def resample_bwd(self, dout_ptr, ld_do, dx, ld_dx, B, Hi, Ho, C, tab: ResampleTables):
        b = tab.bwd_ptrs()
        _lib.call("s2u_resample_bwd", dout_ptr, ld_do, dx.data_ptr(), ld_dx, B, Hi, Hi, Ho, Ho, C, *b, *b, self.dt,
                  self.stream)

    def resample1_fwd(self, x, out, B, Hi, Ho, tab):
        f = tab.fwd_ptrs()
        _lib.call("s2u_resample1_fwd", x.data_ptr(), out.data_ptr(), B, Hi, Hi, Ho, Ho, *f, *f, self.stream)

    def resample1_bwd(self, dout, dx, B, Hi, Ho, tab):
        b = tab.bwd_ptrs()
        _lib.call("s2u_resample1_bwd", dout.data_ptr(), dx.data_ptr(), B, Hi, Hi, Ho, Ho, *b, *b, self.stream)

    def head_fwd(self, feat, ldf, w, bias, out, M):
        _lib.call("s2u_head_fwd", feat.data_ptr(), ldf, w.data_ptr(), bias.data_ptr(), out.data_ptr(), M, 64, self.dt,
                  self.stream)

    def head_bwd(self, feat, ldf, w, dlogit, dfeat, ld_df, accumulate, dw, db, M):
        _lib.call("s2u_head_bwd", feat.data_ptr(), ldf, w.data_ptr(), dlogit.data_ptr(), dfeat.data_ptr(), ld_df,
                  1 if accumulate else 0, dw.data_ptr(), db.data_ptr(), M, 64, self.dt, self.stream)


# ------------------------------------------------------------------------------------------------- engine


@dataclass
class _ConvSpec:
    """One conv+BN unit of the RFB / decoder: where its parameters live in the flat buffers."""
    name: str            # state-dict prefix of the conv (e.g. "rfb1.branch1.1.conv")
    bn: str              # state-dict prefix of its BatchNorm
    cin: int
    kh: int
    kw: int
    dil: int


class Engine:
    def __init__(self, cfg: TrunkConfig, model: torch.nn.Module, dtype: torch.dtype, device, gemm_backend: int = 0):
        self.cfg = cfg
        self.model = model
        self.device = torch.device(device)
        self.ops = Ops(dtype, self.device, gemm_backend)
        self.T = dtype
        self._frozen: Dict[str, torch.Tensor] = {}
        self._shadow: Dict[str, torch.Tensor] = {}
        self._pos_cache: Dict[int, torch.Tensor] = {}
        self._bn_ws: Dict[str, Dict[str, torch.Tensor]] = {}
        self._fold_cache: Dict[str, tuple] = {}
        self.tape: Optional[dict] = None
        self._units = {cs.name: cs for cs in model.conv_units}
        self._prepare_frozen()
        self._shadow_version = -1
        self._refresh = None
        self._streams = None
        self.fuse_adapter = os.environ.get("S2U_FUSE_ADAPTER", "1") != "0"   # adapter + norm1 as one kernel per direction
        self.merge_1x1 = os.environ.get("S2U_MERGE_1X1", "1") != "0"         # an RFB's five 1x1 convs as one GEMM per direction
        # weight gradients feed nothing but the optimizer: they run on their own stream, off the dependency chain of the
        # backward pass (the decoder's and RFB4's backward are chains of tiny kernels the trunk backward waits for)
        self.async_wgrad = os.environ.get("S2U_ASYNC_WGRAD", "1") != "0"
        # RFB branches 1-3 on their own streams: measured no gain (17.95 vs 17.97 ms: the GPU time-slices whole kernels,
        # the step is the sum of their durations either way), so it stays an opt-in switch
        self.branch_streams = os.environ.get("S2U_BRANCH_STREAMS", "0") != "0"
        self._bstreams = None
        self._wg_stream = None
        self._wg_keep = []
        self.igemm = os.environ.get("S2U_CONV_IGEMM", "1") != "0"            # implicit-GEMM convolutions (no im2col)
        self.overlap = self.device.type == "cuda"   # RFB forward/backward on side streams, overlapped with the trunk

    # ------------------------------------------------------------------------------------------ weights

    def _prepare_frozen(self):
        """One-time re-layout of the frozen trunk (setup, not the hot path): W and W^T in the compute dtype."""
        sd = dict(self.model.named_parameters())
        T = self.T
        fz = self._frozen
        for i, spec in enumerate(self.cfg.blocks):
            p = f"encoder.blocks.{i}.block."
            for nm in ("attn.qkv", "attn.proj", "mlp.layers.0", "mlp.layers.1") + (("proj",) if spec.dim != spec.dim_out else ()):
                w = sd[p + nm + ".weight"].detach()
                fz[p + nm + ".w"] = w.to(T).contiguous()
                fz[p + nm + ".wt"] = w.t().to(T).contiguous()
                fz[p + nm + ".b"] = sd[p + nm + ".bias"].detach().float().contiguous()
            for nm in ("norm1", "norm2"):
                fz[p + nm + ".g"] = sd[p + nm + ".weight"].detach().float().contiguous()
                fz[p + nm + ".b"] = sd[p + nm + ".bias"].detach().float().contiguous()
        fz["pe.w"] = sd["encoder.patch_embed.proj.weight"].detach().float().contiguous()
        fz["pe.b"] = sd["encoder.patch_embed.proj.bias"].detach().float().contiguous()
        if T != torch.float32:
            # tensor-core stem: [W | W] against the (hi, lo) bf16 split of the image patches (s2u_patch_im2col)
            w2 = torch.zeros(fz["pe.w"].shape[0], 320, dtype=T, device=fz["pe.w"].device)
            w2[:, :147] = fz["pe.w"].reshape(-1, 147).to(T)
            w2[:, 160:307] = w2[:, :147]
            fz["pe.w2"] = w2
        self._pos_cache.clear()

    def _pos_table(self, hp: int) -> torch.Tensor:
        """hieradet.py:268-276: input-independent, so it is built once per resolution (setup, torch ops)."""
        if hp not in self._pos_cache:
            enc = self.model.encoder
            pe, pw = enc.pos_embed.detach().float(), enc.pos_embed_window.detach().float()
            pos = F.interpolate(pe, size=(hp, hp), mode="bicubic")
            pos = pos + pw.tile([a // b for a, b in zip(pos.shape, pw.shape)])
            self._pos_cache[hp] = pos.permute(0, 2, 3, 1).contiguous().view(hp, hp, -1)
        return self._pos_cache[hp]

    def _pos_bias_rows(self, B: int, hp: int) -> torch.Tensor:
        """position table + stem bias expanded to one fp32 row per token of the batch (setup, cached)."""
        key = ("rows", B, hp)
        if key not in self._pos_cache:
            t = self._pos_table(hp).reshape(1, hp * hp, -1) + self._frozen["pe.b"].view(1, 1, -1)
            self._pos_cache[key] = t.expand(B, -1, -1).reshape(B * hp * hp, -1).contiguous()
        return self._pos_cache[key]

    def refresh_shadows(self):
        """Compute-dtype operands of the TRAINABLE weights, rebuilt from the fp32 masters after every update by ONE
        kernel launch driven by a device-resident table (built on first use)."""
        flat = self.model.flat
        if self._shadow_version == flat.version:
            return
        ops, sh = self.ops, self._shadow
        if self._refresh is None:
            import numpy as np
            P = flat.views
            ent = []
            for i, spec in enumerate(self.cfg.blocks):
                p = f"encoder.blocks.{i}.prompt_learn."
                C = spec.dim
                for nm, (r, c) in (("0", (32, C)), ("2", (C, 32))):
                    key = p + nm
                    sh[key + ".w"], sh[key + ".wt"] = ops.empty(r, c), ops.empty(c, r)
                    ent.append((P[key + ".weight"].data_ptr(), sh[key + ".w"].data_ptr(), sh[key + ".wt"].data_ptr(), 0,
                                r, c, 0, 0, r * c))
            for cs in self.model.conv_units:
                key, taps = cs.name, cs.kh * cs.kw
                sh[key + ".wf"], sh[key + ".wd"] = ops.empty(64, taps * cs.cin), ops.empty(cs.cin, taps * 64)
                ent.append((P[key + ".weight"].data_ptr(), sh[key + ".wf"].data_ptr(), sh[key + ".wd"].data_ptr(), 1, 64,
                            cs.cin, cs.kh, cs.kw, 64 * cs.cin * taps))
            # the five 1x1 convs that read an RFB's input (branch0-3 entry convs, conv_res): ONE [320, Cin] forward
            # operand and ONE [Cin, 320] input-gradient operand, so each direction is a single GEMM
            esz = torch.empty(0, dtype=self.T).element_size()
            for k in range(4):
                names = [f"rfb{k + 1}.branch{b}.0.conv" for b in range(4)] + [f"rfb{k + 1}.conv_res.conv"]
                cin = self._units[names[0]].cin
                wf5, wd5 = ops.empty(320, cin), ops.empty(cin, 320)
                sh[f"rfb{k + 1}.in5.wf"], sh[f"rfb{k + 1}.in5.wd"] = wf5, wd5
                for j, nm in enumerate(names):
                    ent.append((P[nm + ".weight"].data_ptr(), wf5.data_ptr() + 64 * j * cin * esz,
                                wd5.data_ptr() + 64 * j * esz, 1, 64, cin, 1, 1, 64 * cin, 320))
            rec = np.zeros(len(ent), dtype=np.dtype([("src", "<u8"), ("d0p", "<u8"), ("d1p", "<u8"), ("kind", "<i4"),
                                                     ("d0", "<i4"), ("d1", "<i4"), ("d2", "<i4"), ("d3", "<i4"),
                                                     ("pad", "<i4")]))
            blocks = []
            for j, e in enumerate(ent):
                src, d0p, d1p, kind, d0, d1, d2, d3, n = e[:9]
                rec[j] = (src, d0p, d1p, kind, d0, d1, d2, d3, e[9] if len(e) > 9 else 0)
                blocks += [(j, ch) for ch in range((n + 1023) // 1024)]
            entries = torch.from_numpy(rec.view(np.uint8).copy()).to(self.device)
            blk = torch.tensor(blocks, dtype=torch.int32).to(self.device)
            self._refresh = (entries, blk, len(blocks), flat.master.data_ptr())
        entries, blk, nblk, base = self._refresh
        if base != flat.master.data_ptr():
            raise RuntimeError("flat parameter buffer moved; the engine must be rebuilt")
        _lib.call("s2u_refresh_shadows", entries.data_ptr(), blk.data_ptr(), nblk, ops.dt, ops.stream)
        self._shadow_version = flat.version

    def _bn_workspace(self, key: str, C: int) -> Dict[str, torch.Tensor]:
        ws = self._bn_ws.get(key)
        if ws is None:
            dev = self.device
            ws = dict(sums=torch.zeros(_lib.load().s2u_bn_ws_doubles(C), dtype=torch.float64, device=dev),
                      scale=torch.empty(C, dtype=torch.float32, device=dev),
                      shift=torch.empty(C, dtype=torch.float32, device=dev),
                      c1=torch.empty(C, dtype=torch.float32, device=dev),
                      c2=torch.empty(C, dtype=torch.float32, device=dev))
            self._bn_ws[key] = ws
        return ws

    # ------------------------------------------------------------------------------------------ forward

    def forward(self, x: torch.Tensor, training: bool, save: bool):
        """x [B,3,S,S] fp32 NCHW -> (out, out1, out2) fp32 [B,1,S,S]; `save` keeps the tape for backward()."""
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != x.shape[3]:
            raise ValueError("expected a square [B,3,S,S] image batch")
        B, S = x.shape[0], x.shape[2]
        if S % 32:
            raise ValueError("input side must be a multiple of 32 (hieradet.py:272-274 tiles the window embedding)")
        ops, cfg, fz = self.ops, self.cfg, self._frozen
        self.refresh_shadows()
        x = x.contiguous().float()
        tape = dict(B=B, S=S, blocks=[], training=training) if save else None
        H = S // 4
        E = cfg.embed_dim
        # the residual stream is kept in fp32 in both modes (in bf16 mode its rounding error would otherwise
        # random-walk over 2 adds per block); `tc` is its compute-dtype copy, the operand of the next GEMMs
        mixed = self.T != torch.float32
        ts = ops.empty(B * H * H, E, dtype=torch.float32)
        tc = ops.empty(B * H * H, E) if mixed else ts
        if mixed and E % 8 == 0:
            # stem on the tensor cores: patch matrix (bf16 hi/lo split) x [W | W], bias + position table as the fp32
            # residual of the stream epilogue
            col = ops.empty(B * H * H, 320)
            ops.patch_im2col(x, col, B, S)
            ops.gemm(col, fz["pe.w2"], ts, resid=self._pos_bias_rows(B, H), pre_out=tc,
                     flags=RESID | OUT_F32 | RESID_F32 | PRE_FINAL)
        else:
            ops.patch_embed(x, fz["pe.w"], fz["pe.b"], self._pos_table(H), ts, B, S, E, out_copy=tc if mixed else None)
        if tape is not None:
            tape["convs"] = {}
            tape["dec"] = {}
        # Each RFB only depends on its own stage output, so it runs on a side stream and overlaps the rest of the
        # trunk (RFB1 works on the 88x88 map while stages 2-4 run); the streams are joined before the Up stages.
        main = torch.cuda.current_stream(self.device) if self.overlap else None
        side = self._side_streams() if self.overlap else None
        feats, rfb_out, joins = [], [], []
        W = H
        for i, spec in enumerate(cfg.blocks):
            ts, tc, H, W = self._block_fwd(i, spec, ts, tc, B, H, W, tape)
            if spec.stage_end:
                k = len(feats)
                feats.append(tc)                             # keeps the tensor alive until the join
                if self.overlap:
                    fork = torch.cuda.Event()
                    fork.record(main)
                    side[k].wait_event(fork)
                    with torch.cuda.stream(side[k]):
                        rfb_out.append(self._rfb_fwd(k, tc, H, B, training, tape))
                        done = torch.cuda.Event()
                        done.record(side[k])
                    joins.append(done)
                else:
                    rfb_out.append(self._rfb_fwd(k, tc, H, B, training, tape))
        for done in joins:
            main.wait_event(done)
        outs = self._up_fwd(rfb_out, B, S, training, tape)
        if save:
            self.tape = tape
        if training:
            self.model.flat.bump()       # our kernels updated the BatchNorm running statistics through raw pointers
        return outs

    def _branches(self, k, enabled, bodies):
        """Run the three independent branch chains of RFB k (callables `bodies`): the first on the current stream, the
        other two on the RFB's branch streams, forked after everything enqueued so far and joined before returning.
        On the small maps (11x11, 22x22) every conv is a latency-bound launch chain that leaves the GPU idle."""
        if not (enabled and self.overlap and self.branch_streams):
            for body in bodies:
                body()
            return
        if self._bstreams is None:
            self._bstreams = [[torch.cuda.Stream(device=self.device) for _ in range(2)] for _ in range(4)]
        cur = torch.cuda.current_stream(self.device)
        fork = torch.cuda.Event()
        fork.record(cur)
        joins = []
        for st, body in zip(self._bstreams[k], bodies[1:]):
            st.wait_event(fork)
            with torch.cuda.stream(st):
                body()
                ev = torch.cuda.Event()
                ev.record(st)
                joins.append(ev)
        bodies[0]()
        for ev in joins:
            cur.wait_event(ev)

    def _wgrad_async(self, fn, keep):
        """Run the weight-gradient launch `fn` on the weight-gradient stream, after everything enqueued so far on the
        current stream; `keep` holds its operands alive until _wgrad_join()."""
        if not (self.overlap and self.async_wgrad):
            fn()
            return
        if self._wg_stream is None:
            self._wg_stream = torch.cuda.Stream(device=self.device)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._wg_stream.wait_event(ev)
        with torch.cuda.stream(self._wg_stream):
            fn()
        self._wg_keep.append(keep)

    def _wgrad_join(self):
        """The current stream waits for every weight gradient launched so far."""
        if self._wg_stream is not None and self._wg_keep:
            ev = torch.cuda.Event()
            ev.record(self._wg_stream)
            torch.cuda.current_stream(self.device).wait_event(ev)
            self._wg_keep = []

    def _side_streams(self):
        if self._streams is None:
            self._streams = [torch.cuda.Stream(device=self.device) for _ in range(4)]
        return self._streams

    def _block_fwd(self, i, spec, xs, x, B, H, W, tape):
        """xs: residual stream (fp32), x: its compute-dtype copy (the same tensor in fp32 mode)."""
        ops, fz, sh = self.ops, self._frozen, self._shadow
        mixed = self.T != torch.float32
        SF = (OUT_F32 | RESID_F32) if mixed else 0           # stream in, stream out
        P = self.model.flat.views
        p = f"encoder.blocks.{i}.block."
        a = f"encoder.blocks.{i}.prompt_learn."
        C, C2, nh = spec.dim, spec.dim_out, spec.num_heads
        hd = C2 // nh
        R = B * H * W
        f32 = torch.float32
        # adapter (SAM2UNet.py:61-63): xa = x + gelu(gelu(x W1^T + b1) W2^T + b2)
        # h1 / hid hold gelu'(pre-activation), evaluated by the forward epilogue next to gelu itself (SAVE_DGELU)
        xa, n1 = ops.empty(R, C, dtype=f32), ops.empty(R, C)
        mean1, rstd1 = ops.empty(R, dtype=f32), ops.empty(R, dtype=f32)
        if self.fuse_adapter and ops.adapter_supported(C):
            # adapter + norm1 (hieradet.py:134) in one launch, the 32-wide hidden activation never leaves the SM
            h1, u, h2 = (ops.empty(R, 32), ops.empty(R, 32), ops.empty(R, C)) if tape is not None else (None,) * 3
            ops.adapter_ln_fwd(xs, sh[a + "0.w"], P[a + "0.bias"], sh[a + "2.w"], P[a + "2.bias"], fz[p + "norm1.g"],
                               fz[p + "norm1.b"], xa, n1, mean1, rstd1, u, h1, h2, R, C)
        else:
            h1, u = (ops.empty(R, 32) if tape is not None else None), ops.empty(R, 32)
            ops.gemm(x, sh[a + "0.w"], u, bias=P[a + "0.bias"], pre_out=h1,
                     flags=GELU | (SAVE_DGELU if tape is not None else 0))
            h2 = ops.empty(R, C) if tape is not None else None
            ops.gemm(u, sh[a + "2.w"], xa, bias=P[a + "2.bias"], pre_out=h2, resid=xs,
                     flags=GELU | RESID | SF | (SAVE_DGELU if tape is not None else 0))
            # norm1 (hieradet.py:134)
            ops.ln_fwd(xa, fz[p + "norm1.g"], fz[p + "norm1.b"], n1, mean1, rstd1, R, C)
        pr = None
        Ho, Wo = H, W
        if C != C2:                                   # hieradet.py:137-138: shortcut = pool(proj(norm1(x)))
            pr = ops.empty(R, C2)
            ops.gemm(n1, fz[p + "proj.w"], pr, bias=fz[p + "proj.b"])
            Ho, Wo = H // 2, W // 2
            sc = ops.empty(B * Ho * Wo, C2)
            ops.maxpool_fwd(pr, sc, B, H, W, C2)
            sc_flags = OUT_F32 if mixed else 0               # pooled shortcut is in the compute dtype
        else:
            sc, sc_flags = xa, SF
        Ro = B * Ho * Wo
        # qkv on real tokens only; padding, windows, q-pool live inside the attention kernel
        qkv = ops.empty(R, 3 * C2)
        ops.gemm(n1, fz[p + "attn.qkv.w"], qkv, bias=fz[p + "attn.qkv.b"])
        o = ops.empty(Ro, C2)
        lse = ops.empty(Ro, nh, dtype=f32)
        ops.attn_fwd(qkv, fz[p + "attn.qkv.b"], o, lse, B, H, W, nh, hd, spec.window, spec.q_pool)
        y = ops.empty(Ro, C2, dtype=f32)
        ops.gemm(o, fz[p + "attn.proj.w"], y, bias=fz[p + "attn.proj.b"], resid=sc, flags=RESID | sc_flags)
        n2 = ops.empty(Ro, C2)
        mean2, rstd2 = ops.empty(Ro, dtype=f32), ops.empty(Ro, dtype=f32)
        ops.ln_fwd(y, fz[p + "norm2.g"], fz[p + "norm2.b"], n2, mean2, rstd2, Ro, C2)
        hid = ops.empty(Ro, 4 * C2) if tape is not None else None
        act = ops.empty(Ro, 4 * C2)
        ops.gemm(n2, fz[p + "mlp.layers.0.w"], act, bias=fz[p + "mlp.layers.0.b"], pre_out=hid,
                 flags=GELU | (SAVE_DGELU if tape is not None else 0))
        z = ops.empty(Ro, C2, dtype=f32)
        zc = ops.empty(Ro, C2) if mixed else z
        ops.gemm(act, fz[p + "mlp.layers.1.w"], z, bias=fz[p + "mlp.layers.1.b"], resid=y, flags=RESID | SF | (PRE_FINAL if mixed else 0),
                 pre_out=zc if mixed else None)
        if tape is not None:
            tape["blocks"].append(dict(x=x, h1=h1, u=u, h2=h2, xa=xa, mean1=mean1, rstd1=rstd1, pr=pr, qkv=qkv, o=o,
                                       lse=lse, y=y, mean2=mean2, rstd2=rstd2, hid=hid, H=H, W=W, Ho=Ho, Wo=Wo))
        return z, zc, Ho, Wo

    def _bn_folded(self, cs: "_ConvSpec"):
        """Eval mode: conv weight x BN scale in the GEMM layout [64][(ky, kx, ci)] (compute dtype) and the BN shift as
        the GEMM bias; rebuilt when the parameters change (setup, torch ops)."""
        flat = self.model.flat
        hit = self._fold_cache.get(cs.name)
        if hit is not None and hit[0] == flat.version:
            return hit[1], hit[2]
        P, Bf = flat.views, flat.buffers
        with torch.no_grad():
            scale = P[cs.bn + ".weight"].float() / torch.sqrt(Bf[cs.bn + ".running_var"].float() + 1e-5)
            shift = (P[cs.bn + ".bias"].float() - Bf[cs.bn + ".running_mean"].float() * scale).contiguous()
            w = P[cs.name + ".weight"].float()                                  # [64, Cin, kh, kw]
            wf = (w.permute(0, 2, 3, 1).reshape(w.shape[0], -1) * scale[:, None]).to(self.T).contiguous()
        self._fold_cache[cs.name] = (flat.version, wf, shift)
        return wf, shift

    # conv (+ BN (+ residual) (+ ReLU)) on NHWC rows.  `src`: (tensor, ld, channel offset) of the input map.
    def _conv_bn(self, cs: _ConvSpec, src, B, H, out, ld_out, out_off, relu, training, tape, resid=None, ld_res=0,
                 raw_in=None):
        """raw_in = (tensor, pitch): the conv output already computed (column slice of a merged GEMM)."""
        ops, sh = self.ops, self._shadow
        P, Bf = self.model.flat.views, self.model.flat.buffers
        x, ldx, xoff = src
        M = B * H * H
        taps = cs.kh * cs.kw
        esz = x.element_size()
        xin = x.view(-1)[xoff:] if xoff else x
        ig = self.igemm and taps > 1 and ops.conv_igemm_ok(cs.cin, 64, ldx, ld_out)
        col, ldcol, col_ptr_off = None, 0, 0
        if taps == 1:
            col, ldcol = x, ldx
            col_ptr_off = xoff
        elif not ig:
            col = ops.empty(M, taps * cs.cin)
            ph, pw = cs.dil * (cs.kh - 1) // 2, cs.dil * (cs.kw - 1) // 2
            ops.im2col(xin, ldx, col, B, H, H, cs.cin, cs.kh, cs.kw, cs.dil, cs.dil, ph, pw)
            ldcol, col_ptr_off = taps * cs.cin, 0
        A = None if ig else (col.view(-1)[col_ptr_off:] if col_ptr_off else col)
        if not training and tape is None:
            # inference: BatchNorm (running statistics) folded into the weights and a bias, residual and ReLU in the
            # GEMM epilogue - one launch per conv instead of GEMM + finalize + apply
            wf, shift = self._bn_folded(cs)
            C = out.view(-1)[out_off:] if out_off else out
            if ig:
                ops.conv_igemm(xin, ldx, B, H, H, cs.cin, wf, 64, cs.kh, cs.kw, cs.dil, C, ld_out, bias=shift,
                               resid=resid, ld_res=ld_res, relu=relu)
            else:
                ops.gemm(A, wf, C, bias=shift, resid=resid, ld_res=ld_res, M=M, N=64, K=taps * cs.cin, lda=ldcol,
                         ldw=taps * cs.cin, ldc=ld_out,
                         flags=(RESID if resid is not None else 0) | (RELU if relu else 0))
            return
        raw, ldraw = (ops.empty(M, 64), 64) if raw_in is None else raw_in
        ws = self._bn_workspace(cs.bn, 64)
        mean = rstd = None
        if raw_in is not None:
            if training:
                mean, rstd = ops.empty(64, dtype=torch.float32), ops.empty(64, dtype=torch.float32)
                ops.bn_stats_finalize(raw, ldraw, ws["sums"], P[cs.bn + ".weight"], P[cs.bn + ".bias"],
                                      Bf[cs.bn + ".running_mean"], Bf[cs.bn + ".running_var"],
                                      Bf[cs.bn + ".num_batches_tracked"], ws["scale"], ws["shift"], mean, rstd, M, 64)
            else:
                ops.bn_finalize(ws["sums"], P[cs.bn + ".weight"], P[cs.bn + ".bias"], Bf[cs.bn + ".running_mean"],
                                Bf[cs.bn + ".running_var"], Bf[cs.bn + ".num_batches_tracked"], ws["scale"], ws["shift"],
                                None, None, M, 64, False)
        elif ig:
            # conv + BatchNorm batch statistics (+ their finalisation by the last CTA) in one launch
            if training:
                mean, rstd = ops.empty(64, dtype=torch.float32), ops.empty(64, dtype=torch.float32)
                ops.conv_igemm_bn(xin, ldx, B, H, H, cs.cin, sh[cs.name + ".wf"], cs.kh, cs.kw, cs.dil, raw, 64,
                                  ws["sums"], P[cs.bn + ".weight"], P[cs.bn + ".bias"], Bf[cs.bn + ".running_mean"],
                                  Bf[cs.bn + ".running_var"], Bf[cs.bn + ".num_batches_tracked"], ws["scale"],
                                  ws["shift"], mean, rstd)
            else:
                ops.conv_igemm(xin, ldx, B, H, H, cs.cin, sh[cs.name + ".wf"], 64, cs.kh, cs.kw, cs.dil, raw, 64)
                ops.bn_finalize(ws["sums"], P[cs.bn + ".weight"], P[cs.bn + ".bias"], Bf[cs.bn + ".running_mean"],
                                Bf[cs.bn + ".running_var"], Bf[cs.bn + ".num_batches_tracked"], ws["scale"], ws["shift"],
                                None, None, M, 64, False)
        else:
            ops.gemm(A, sh[cs.name + ".wf"], raw, M=M, N=64, K=taps * cs.cin, lda=ldcol, ldw=taps * cs.cin, ldc=64)
            if training:
                mean, rstd = ops.empty(64, dtype=torch.float32), ops.empty(64, dtype=torch.float32)
                ops.bn_stats_finalize(raw, 64, ws["sums"], P[cs.bn + ".weight"], P[cs.bn + ".bias"],
                                      Bf[cs.bn + ".running_mean"], Bf[cs.bn + ".running_var"],
                                      Bf[cs.bn + ".num_batches_tracked"], ws["scale"], ws["shift"], mean, rstd, M, 64)
            else:
                ops.bn_finalize(ws["sums"], P[cs.bn + ".weight"], P[cs.bn + ".bias"], Bf[cs.bn + ".running_mean"],
                                Bf[cs.bn + ".running_var"], Bf[cs.bn + ".num_batches_tracked"], ws["scale"], ws["shift"],
                                None, None, M, 64, False)
        out_ptr = out.data_ptr() + out_off * esz
        ops.bn_apply(raw, ldraw, ws["scale"], ws["shift"], resid, ld_res, out_ptr, ld_out, M, 64, relu)
        if tape is not None:
            tape["convs"][cs.name] = dict(col=col, ldcol=ldcol, coff=col_ptr_off, raw=raw, ldraw=ldraw, mean=mean,
                                          rstd=rstd, B=B, H=H, ig=ig, xin=xin, ldx=ldx)

    def _rfb_fwd(self, k, f, H, B, training, tape):
        """RFB_modified k on the stage-k feature map f [B*H*H, Cin] (SAM2UNet.py:117-125); returns (dst, ld, H) where
        dst is the left half of up-stage k's concat buffer [skip | upsampled] (plain [M,64] for the deepest level)."""
        ops = self.ops
        units = self._units
        r = f"rfb{k + 1}."
        M = B * H * H
        Cin = f.shape[1]
        if k < 3:
            dst, ld_dst = ops.empty(M, 128), 128
        else:
            dst, ld_dst = ops.empty(M, 64), 64
        cat = ops.empty(M, 256)
        src = (f, Cin, 0)
        # the five 1x1 convs on f (branch0-3 entry convs, conv_res) as ONE GEMM with N = 320: f is read once
        merged = (training or tape is not None) and self.merge_1x1
        raw5 = None
        if merged:
            raw5 = ops.empty(M, 320)
            ops.gemm(f, self._shadow[r + "in5.wf"], raw5, M=M, N=320, K=Cin, lda=Cin, ldw=Cin, ldc=320)
        rin = (lambda j: (raw5.view(-1)[64 * j:], 320)) if merged else (lambda j: None)
        self._conv_bn(units[r + "branch0.0.conv"], src, B, H, cat, 256, 0, False, training, tape, raw_in=rin(0))
        keep = []

        def branch_fwd(bi):
            def body():
                t0, t1, t2 = ops.empty(M, 64), ops.empty(M, 64), ops.empty(M, 64)
                keep.append((t0, t1, t2))
                self._conv_bn(units[r + f"branch{bi}.0.conv"], src, B, H, t0, 64, 0, False, training, tape, raw_in=rin(bi))
                self._conv_bn(units[r + f"branch{bi}.1.conv"], (t0, 64, 0), B, H, t1, 64, 0, False, training, tape)
                self._conv_bn(units[r + f"branch{bi}.2.conv"], (t1, 64, 0), B, H, t2, 64, 0, False, training, tape)
                self._conv_bn(units[r + f"branch{bi}.3.conv"], (t2, 64, 0), B, H, cat, 256, 64 * bi, False, training, tape)
            return body

        self._branches(k, True, [branch_fwd(bi) for bi in (1, 2, 3)])
        res = ops.empty(M, 64)
        self._conv_bn(units[r + "conv_res.conv"], src, B, H, res, 64, 0, False, training, tape, raw_in=rin(4))
        self._conv_bn(units[r + "conv_cat.conv"], (cat, 256, 0), B, H, dst, ld_dst, 0, True, training, tape,
                      resid=res, ld_res=64)
        if tape is not None:
            tape["dec"][r] = dict(f=f, cat=cat, dst=dst, ld_dst=ld_dst, H=H, Cin=Cin, merged=merged)
        return dst, ld_dst, H

    def _up_fwd(self, rfb_out, B, S, training, tape):
        """up1(x4, x3) -> side1 ; up2(., x2) -> side2 ; up3(., x1) -> head   (SAM2UNet.py:167-172)."""
        ops = self.ops
        units = self._units
        P = self.model.flat.views
        cur, cur_ld, cur_h = rfb_out[3]
        outs = {}
        for stage, (k, head_name, scale) in enumerate(((2, "side1", 16), (1, "side2", 8), (0, "head", 4))):
            u = f"up{stage + 1}."
            catb, _, H = rfb_out[k]
            if H != 2 * cur_h:
                raise _lib.KernelError("decoder expects exact 2x pyramid levels (input side multiple of 32)")
            M = B * H * H
            tab = ResampleTables.get(cur_h, H, True, None, self.device)
            ops.resample_fwd(cur, cur_ld, catb.data_ptr() + 64 * catb.element_size(), 128, B, cur_h, H, 64, tab)
            mid, nxt = ops.empty(M, 64), ops.empty(M, 64)
            self._conv_bn(units[u + "conv.double_conv.0"], (catb, 128, 0), B, H, mid, 64, 0, True, training, tape)
            self._conv_bn(units[u + "conv.double_conv.3"], (mid, 64, 0), B, H, nxt, 64, 0, True, training, tape)
            low = ops.empty(B, H, H, dtype=torch.float32)
            ops.head_fwd(nxt, 64, P[head_name + ".weight"], P[head_name + ".bias"], low, M)
            full = ops.empty(B, 1, S, S, dtype=torch.float32)
            ops.resample1_fwd(low, full, B, H, S, ResampleTables.get(H, S, False, float(scale), self.device))
            outs[head_name] = full
            if tape is not None:
                tape["dec"][u] = dict(inp=cur, inp_ld=cur_ld, inp_h=cur_h, catb=catb, mid=mid, out=nxt, H=H,
                                      head=head_name, scale=scale)
            cur, cur_ld, cur_h = nxt, 64, H
        return outs["head"], outs["side1"], outs["side2"]

    # ----------------------------------------------------------------------------------------- backward

    def backward(self, g_out, g_out1, g_out2, on_bucket=None):
        """Gradients of (out, out1, out2) [B,1,S,S] fp32 -> accumulates into model.flat.grad; frees the tape.
        `on_bucket(lo, hi)` is called as soon as the flat gradient range [lo, hi) is final (data-parallel overlap)."""
        tape = self.tape
        if tape is None:
            raise RuntimeError("backward() without a saved forward")
        if not tape["training"]:
            raise NotImplementedError("backward through eval-mode BatchNorm is not implemented")
        self.tape = None
        B, S = tape["B"], tape["S"]
        nblocks = len(self.cfg.blocks)
        buckets = self.model.flat.buckets if on_bucket is not None else []
        main = torch.cuda.current_stream(self.device) if self.overlap else None
        side = self._side_streams() if self.overlap else None
        d_rfb = self._up_bwd(tape, {"head": g_out, "side1": g_out1, "side2": g_out2}, B, S)
        # the four RFB backwards are independent: side streams, deepest first (its gradient is needed first);
        # RFB1's backward (the 88x88 map) overlaps the whole trunk backward
        d_feats, joins = [None] * 4, [None] * 4
        for k in (3, 2, 1, 0):
            if self.overlap:
                fork = torch.cuda.Event()
                fork.record(main)
                side[k].wait_event(fork)
                with torch.cuda.stream(side[k]):
                    d_feats[k] = self._rfb_bwd(k, tape, d_rfb[k], B)
                    joins[k] = torch.cuda.Event()
                    joins[k].record(side[k])
            else:
                d_feats[k] = self._rfb_bwd(k, tape, d_rfb[k], B)
        # trunk: walk the blocks backwards; stage-end blocks receive the RFB's gradient of that feature map
        dz = None
        stage = len(d_feats) - 1
        for i in range(nblocks - 1, -1, -1):
            spec = self.cfg.blocks[i]
            if spec.stage_end:
                if joins[stage] is not None:
                    main.wait_event(joins[stage])
                g = d_feats[stage]
                if stage == 0:                           # every RFB / decoder gradient is final now
                    for lo, hi, ready in buckets:
                        if ready == nblocks:
                            self._wgrad_join()
                            on_bucket(lo, hi)
                stage -= 1
                if dz is None:
                    dz = g
                else:
                    self.ops.add(dz, g, dz)
            dz = self._block_bwd(i, spec, tape["blocks"][i], dz, B)
            tape["blocks"][i] = None
            for lo, hi, ready in buckets:
                if ready == i:
                    self._wgrad_join()
                    on_bucket(lo, hi)
        self._wgrad_join()

    def _block_bwd(self, i, spec, tp, dz, B):
        ops, fz, sh = self.ops, self._frozen, self._shadow
        G = self.model.flat.grad_views
        p = f"encoder.blocks.{i}.block."
        a = f"encoder.blocks.{i}.prompt_learn."
        C, C2, nh = spec.dim, spec.dim_out, spec.num_heads
        hd = C2 // nh
        H, W, Ho, Wo = tp["H"], tp["W"], tp["Ho"], tp["Wo"]
        R, Ro = B * H * W, B * Ho * Wo
        # z = y + fc2(gelu(fc1(LN2(y))))
        dh = ops.empty(Ro, 4 * C2)
        ops.gemm(dz, fz[p + "mlp.layers.1.wt"], dh, aux=tp["hid"], flags=MULAUX)
        dn2 = ops.empty(Ro, C2)
        ops.gemm(dh, fz[p + "mlp.layers.0.wt"], dn2)
        del dh
        dy = ops.empty(Ro, C2)
        ops.ln_bwd(dn2, tp["y"], fz[p + "norm2.g"], tp["mean2"], tp["rstd2"], dz, dy, Ro, C2)
        # y = shortcut + proj(attn)
        do = dn2                                        # reuse
        ops.gemm(dy, fz[p + "attn.proj.wt"], do)
        dqkv = ops.empty(R, 3 * C2)
        ops.attn_bwd(tp["qkv"], fz[p + "attn.qkv.b"], tp["o"], tp["lse"], do, dqkv, B, H, W, nh, hd, spec.window,
                     spec.q_pool)
        dn1 = ops.empty(R, C)
        ops.gemm(dqkv, fz[p + "attn.qkv.wt"], dn1)
        del dqkv
        dres = dy
        if C != C2:
            dpr = ops.empty(R, C2)
            ops.maxpool_bwd(tp["pr"], dy, dpr, B, H, W, C2)
            ops.gemm(dpr, fz[p + "proj.wt"], dn1, resid=dn1, flags=RESID)
            dres = None
        # LN1 backward fused with the head of the adapter backward (xa = x + gelu(h2), h2 = u W2^T + b2,
        # u = gelu(h1), h1 = x W1^T + b1): dxa, dh2 = dxa * gelu'(h2) and db2 = colsum(dh2) in one pass
        if self.fuse_adapter and ops.adapter_supported(C):
            dh2, dh1, dx = ops.empty(R, C), ops.empty(R, 32), ops.empty(R, C)
            ops.adapter_ln_bwd(dn1, tp["xa"], tp["mean1"], tp["rstd1"], fz[p + "norm1.g"], dres, tp["h2"], tp["h1"],
                               sh[a + "2.wt"], sh[a + "0.wt"], dh2, dh1, dx, G[a + "0.bias"], G[a + "2.bias"], R, C)
            u_, x_ = tp["u"], tp["x"]
            self._wgrad_async(lambda: ops.wgrad_pair(dh2, u_, G[a + "2.weight"], 32, dh1, x_, G[a + "0.weight"], C),
                              (dh2, dh1, u_, x_))
            return dx
        dxa, dh2 = ops.empty(R, C), ops.empty(R, C)
        ops.ln_bwd(dn1, tp["xa"], fz[p + "norm1.g"], tp["mean1"], tp["rstd1"], dres, dxa, R, C, pre=tp["h2"], dx2=dh2,
                   colsum=G[a + "2.bias"], pre_is_grad=True)         # h2 holds gelu'(pre-activation)
        dh1 = ops.empty(R, 32)
        ops.gemm(dh2, sh[a + "2.wt"], dh1, aux=tp["h1"], flags=MULAUX)  # (dh2 W2) * gelu'(h1)
        # dW2 [C, 32] = dh2^T u and dW1 [32, C] = dh1^T x in one launch
        ops.wgrad_pair(dh2, tp["u"], G[a + "2.weight"], 32, dh1, tp["x"], G[a + "0.weight"], C)
        ops.colsum(dh1, G[a + "0.bias"])
        dx = ops.empty(R, C)
        ops.gemm(dh1, sh[a + "0.wt"], dx, resid=dxa, flags=RESID)
        return dx

    def _conv_bn_bwd(self, cs: _ConvSpec, tape, dy, ld_dy, dy_off, y, ld_y, y_off, dst, accumulate, draw_out=None):
        """Backward of one conv+BN unit.  dy / y: (tensor, pitch, channel offset) of the output gradient and, when
        the unit ends in a ReLU, of its saved output.  dst = (tensor, ld, offset) receiving d(input)."""
        ops, sh = self.ops, self._shadow
        P, G = self.model.flat.views, self.model.flat.grad_views
        tp = tape["convs"][cs.name]
        B, H = tp["B"], tp["H"]
        M = B * H * H
        taps = cs.kh * cs.kw
        ws = self._bn_workspace(cs.bn, 64)
        # d(raw): its own [M,64] tensor, or a column slice (pitch 320) of the merged input-gradient GEMM's operand
        draw, ld_draw = (ops.empty(M, 64), 64) if draw_out is None else draw_out
        dyv = dy.view(-1)[dy_off:] if dy_off else dy
        yv = None
        if y is not None:
            yv = y.view(-1)[y_off:] if y_off else y
        ops.bn_bwd(dyv, ld_dy, yv, ld_y, tp["raw"], tp["ldraw"], tp["mean"], tp["rstd"], P[cs.bn + ".weight"], ws["sums"],
                   G[cs.bn + ".weight"], G[cs.bn + ".bias"], ws["c1"], ws["c2"], draw, ld_draw, M, 64)
        if tp["ig"]:
            # weight gradient straight from the un-expanded input map (tap-shifted TMA boxes)
            xin_, ldx_ = tp["xin"], tp["ldx"]
            self._wgrad_async(lambda: ops.conv_wgrad(draw, 64, xin_, ldx_, G[cs.name + ".weight"], B, H, H, cs.cin, 64,
                                                     cs.kh, cs.kw, cs.dil), (draw, xin_))
        else:
            col = tp["col"]
            colv = col.view(-1)[tp["coff"]:] if tp["coff"] else col
            ldcol_ = tp["ldcol"]
            self._wgrad_async(lambda: ops.wgrad(draw, colv, G[cs.name + ".weight"], M=M, P=64, Q=taps * cs.cin,
                                                lda=ld_draw, ldb=ldcol_, ldg=taps * cs.cin, q_inner=cs.cin,
                                                q_taps=taps), (draw, colv, col))
        if dst is None:
            return
        dt, ld_d, d_off = dst
        dv = dt.view(-1)[d_off:] if d_off else dt
        if taps > 1 and self.igemm and ops.conv_igemm_ok(64, cs.cin, 64, ld_d):
            # input gradient = the same implicit convolution of d(raw) with the flipped / transposed weights
            ops.conv_igemm(draw, 64, B, H, H, 64, sh[cs.name + ".wd"], cs.cin, cs.kh, cs.kw, cs.dil, dv, ld_d,
                           resid=dv if accumulate else None, ld_res=ld_d)
            return
        if taps == 1:
            A, lda = draw, 64
        else:
            A = ops.empty(M, taps * 64)
            ph, pw = cs.dil * (cs.kh - 1) // 2, cs.dil * (cs.kw - 1) // 2
            ops.im2col(draw, 64, A, B, H, H, 64, cs.kh, cs.kw, cs.dil, cs.dil, ph, pw)
            lda = taps * 64
        ops.gemm(A, sh[cs.name + ".wd"], dv, M=M, N=cs.cin, K=taps * 64, lda=lda, ldw=taps * 64, ldc=ld_d,
                 resid=dv if accumulate else None, ld_res=ld_d, flags=RESID if accumulate else 0)

    def _up_bwd(self, tape, g_heads, B, S):
        """Backward of the heads and the three Up stages; returns, per pyramid level k, (tensor, pitch) holding the
        gradient of RFB k's output (the left half of the concat-buffer gradient; plain [M,64] for the deepest)."""
        ops = self.ops
        units = self._units
        P, G = self.model.flat.views, self.model.flat.grad_views
        dec = tape["dec"]
        d_rfb = [None, None, None, None]
        d_cur = None                                    # gradient of the current decoder feature map [M,64]
        for stage in (2, 1, 0):                          # up3, up2, up1
            u = f"up{stage + 1}."
            tp = dec[u]
            H, M = tp["H"], B * tp["H"] * tp["H"]
            k = 2 - stage
            tab = ResampleTables.get(H, S, False, float(tp["scale"]), self.device)
            dlow = ops.empty(B, H, H, dtype=torch.float32)
            ops.resample1_bwd(g_heads[tp["head"]].contiguous().float(), dlow, B, H, S, tab)
            dfeat = ops.empty(M, 64) if d_cur is None else d_cur
            ops.head_bwd(tp["out"], 64, P[tp["head"] + ".weight"], dlow, dfeat, 64, d_cur is not None,
                         G[tp["head"] + ".weight"], G[tp["head"] + ".bias"], M)
            dmid = ops.empty(M, 64)
            self._conv_bn_bwd(units[u + "conv.double_conv.3"], tape, dfeat, 64, 0, tp["out"], 64, 0, (dmid, 64, 0),
                              False)
            dcat = ops.empty(M, 128)
            self._conv_bn_bwd(units[u + "conv.double_conv.0"], tape, dmid, 64, 0, tp["mid"], 64, 0, (dcat, 128, 0),
                              False)
            d_rfb[k] = (dcat, 128)                      # left half [:, :64] is d(rfb_k output)
            h_in = tp["inp_h"]
            dprev = ops.empty(B * h_in * h_in, 64)
            ops.resample_bwd(dcat.data_ptr() + 64 * dcat.element_size(), 128, dprev, 64, B, h_in, H, 64,
                             ResampleTables.get(h_in, H, True, None, self.device))
            d_cur = dprev
        d_rfb[3] = (d_cur, 64)
        return d_rfb

    def _rfb_bwd(self, k, tape, d_out, B):
        """Backward of RFB k given the gradient of its output; returns the gradient of the stage-k feature map."""
        ops = self.ops
        units = self._units
        r = f"rfb{k + 1}."
        tp = tape["dec"][r]
        H, Cin = tp["H"], tp["Cin"]
        M = B * H * H
        dy, ld_dy = d_out
        # out = relu(bn(conv_cat(cat)) + bn(conv_res(f))): g = dy * (out > 0) feeds both BN backwards
        g = ops.empty(M, 64)
        ops.relu_bwd(dy, ld_dy, tp["dst"], tp["ld_dst"], g, 64, M, 64)
        df = ops.empty(M, Cin)
        dcatb = ops.empty(M, 256)
        merged = tp["merged"]
        draw5 = ops.empty(M, 320) if merged else None
        dsl = (lambda j: (draw5.view(-1)[64 * j:], 320)) if merged else (lambda j: None)
        dfd = (lambda acc: None) if merged else (lambda acc: (df, Cin, 0))
        self._conv_bn_bwd(units[r + "conv_cat.conv"], tape, g, 64, 0, None, 0, 0, (dcatb, 256, 0), False)
        self._conv_bn_bwd(units[r + "conv_res.conv"], tape, g, 64, 0, None, 0, 0, dfd(False), False, draw_out=dsl(4))
        self._conv_bn_bwd(units[r + "branch0.0.conv"], tape, dcatb, 256, 0, None, 0, 0, dfd(True), True, draw_out=dsl(0))
        keep = []

        def branch_bwd(bi):
            def body():
                d2, d1, d0 = ops.empty(M, 64), ops.empty(M, 64), ops.empty(M, 64)
                keep.append((d2, d1, d0))
                self._conv_bn_bwd(units[r + f"branch{bi}.3.conv"], tape, dcatb, 256, 64 * bi, None, 0, 0, (d2, 64, 0), False)
                self._conv_bn_bwd(units[r + f"branch{bi}.2.conv"], tape, d2, 64, 0, None, 0, 0, (d1, 64, 0), False)
                self._conv_bn_bwd(units[r + f"branch{bi}.1.conv"], tape, d1, 64, 0, None, 0, 0, (d0, 64, 0), False)
                self._conv_bn_bwd(units[r + f"branch{bi}.0.conv"], tape, d0, 64, 0, None, 0, 0, dfd(True), True,
                                  draw_out=dsl(bi))
            return body

        # (without the merged input-gradient GEMM the three chains accumulate into df one after the other)
        self._branches(k, merged, [branch_bwd(bi) for bi in (1, 2, 3)])
        if merged:                                      # df = [d(raw) of the five 1x1 convs] . [Cin, 320]^T in one GEMM
            ops.gemm(draw5, self._shadow[r + "in5.wd"], df, M=M, N=Cin, K=320, lda=320, ldw=320, ldc=Cin)
        tape["dec"][r]["keep"] = (g, dcatb, dy, draw5, keep)  # alive until the streams are joined
        return df

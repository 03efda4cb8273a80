"""`SAM2UNet` — drop-in for /root/reference/SAM2UNet.py:128-173 running on the sm_100a kernels.

Same constructor (`SAM2UNet(checkpoint_path="")`), same submodule / state-dict names, same
`forward(x) -> (out, out1, out2)` logits [B,1,S,S] fp32, same `requires_grad` pattern (trunk frozen,
adapters + RFB + decoder + heads trainable, SAM2UNet.py:146-162).  Keyword-only extras select the trunk
yaml (the fork hard-codes Hiera-S, SAM2UNet.py:131) and the compute dtype.

The module only holds parameters; `forward` hands the batch to `engine.Engine`, which launches the CUDA
kernels.  There is no CPU path and no ATen/cuBLAS/cuDNN fallback: calling it with a CPU tensor raises.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import _lib
from .config import TrunkConfig, trunk_config
from .engine import Engine, _ConvSpec
from .params import _RFB, _Trunk, _Up, trainable_filter

_DTYPES = {"fp32": torch.float32, "float32": torch.float32, "bf16": torch.bfloat16, "bfloat16": torch.bfloat16}


class FlatParams:
    """All trainable parameters in ONE fp32 buffer (+ one gradient buffer); the nn.Parameters alias it.

    Order = order in which backward completes the gradients (heads, decoder, RFBs, then adapters of the last
    block down to block 0), so data-parallel buckets are contiguous ranges that become ready front to back.
    The never-used `up4.*` (SAM2UNet.py:159, never called in forward) sits after `n_active`: it gets no
    gradient and no optimizer update, exactly like a `grad is None` parameter in the reference.
    """

    def __init__(self, model: "SAM2UNet", device: torch.device):
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        nblocks = len(model.cfg.blocks)

        def rank(name: str):
            if name.startswith(("head.", "side")):
                return (0, 0)
            if name.startswith("up4."):
                return (9, 0)
            if name.startswith("up"):
                return (1, -int(name[2]))
            if name.startswith("rfb"):
                return (2, int(name[3]))
            if name.startswith("encoder.blocks."):
                return (3, nblocks - int(name.split(".")[2]))
            return (8, 0)

        named.sort(key=lambda kv: rank(kv[0]))          # stable: keeps registration order inside a group
        sizes = [p.numel() for _, p in named]
        offs, total = [], 0
        for s in sizes:
            offs.append(total)
            total += (s + 7) // 8 * 8                    # 32-byte aligned slots
        self.n_total = total
        self.n_active = next((o for (n, _), o in zip(named, offs) if n.startswith("up4.")), total)
        self.master = torch.zeros(total, dtype=torch.float32, device=device)
        self.grad = torch.zeros(total, dtype=torch.float32, device=device)
        self.views: Dict[str, torch.Tensor] = {}
        self.grad_views: Dict[str, torch.Tensor] = {}
        self.params: Dict[str, nn.Parameter] = {}
        self.offsets: Dict[str, int] = {}
        with torch.no_grad():
            for (n, p), o, s in zip(named, offs, sizes):
                v = self.master[o:o + s].view(p.shape)
                v.copy_(p.detach().to(device=device, dtype=torch.float32))
                p.data = v
                self.views[n] = v
                self.grad_views[n] = self.grad[o:o + s].view(p.shape)
                self.params[n] = p
                self.offsets[n] = o
        self.buffers: Dict[str, torch.Tensor] = {n: b for n, b in model.named_buffers()}
        self._active_params = [p for n, p in self.params.items() if self.offsets[n] < self.n_active]
        self._float_buffers = [b for b in self.buffers.values() if b.dtype.is_floating_point]
        self._manual = 0
        # bucket boundaries for the gradient all-reduce: decoder side first, then ~8 blocks of adapters each
        self.buckets = self._buckets(named, offs)

    def _buckets(self, named, offs, blocks_per_bucket: int = 12):
        """[(lo, hi, ready_block)]: flat range + index of the trunk block whose backward completes it
        (ready_block == number of blocks means "ready as soon as the decoder/RFB backward is done")."""
        nblocks = len({n.split(".")[2] for n, _ in named if n.startswith("encoder.blocks.")})
        first_adapter = next((o for (n, _), o in zip(named, offs) if n.startswith("encoder.")), self.n_active)
        out = [(0, first_adapter, nblocks)]
        start = {}
        for (n, _), o in zip(named, offs):
            if n.startswith("encoder.blocks."):
                start.setdefault(int(n.split(".")[2]), o)
        order = sorted(start, reverse=True)              # backward visits the last block first
        for j in range(0, len(order), blocks_per_bucket):
            grp = order[j:j + blocks_per_bucket]
            lo = start[grp[0]]
            nxt = order[j + blocks_per_bucket] if j + blocks_per_bucket < len(order) else None
            hi = start[nxt] if nxt is not None else self.n_active
            out.append((lo, hi, grp[-1]))
        return [b for b in out if b[1] > b[0]]

    @property
    def version(self):
        """Changes whenever a trainable parameter or a BatchNorm buffer may have changed.  The parameters alias the
        flat master buffer through `p.data`, so an in-place update by a stock optimizer (torch.optim.AdamW, EMA,
        `p.copy_`) bumps `p._version` but NOT `master._version`: the counter is therefore derived from the parameters
        and buffers themselves; writes by our own kernels (raw pointers) are announced with bump()."""
        v = self._manual + self.master._version
        for p in self._active_params:
            v += p._version
        for b in self._float_buffers:
            v += b._version
        return v

    def bump(self):
        self._manual += 1


class _SAM2UNetFn(torch.autograd.Function):
    """One autograd node for the whole network: forward and backward are the engine's kernel schedules."""

    @staticmethod
    def forward(ctx, x, anchor, model):
        eng = model._engine(x.device)
        outs = eng.forward(x, model.training, save=True)
        ctx.model = model
        ctx.engine = eng
        return outs

    @staticmethod
    def backward(ctx, g0, g1, g2):
        model, eng = ctx.model, ctx.engine
        flat = model.flat
        zeros = None
        gs = []
        for g, ref in zip((g0, g1, g2), eng.tape_out_shapes):
            if g is None:
                if zeros is None:
                    zeros = torch.zeros(ref, dtype=torch.float32, device=flat.master.device)
                g = zeros
            gs.append(g.contiguous().float())
        active = [(n, p) for n, p in flat.params.items() if flat.offsets[n] < flat.n_active]
        foreign = [(n, p) for n, p in active if p.grad is not None and p.grad.data_ptr() != flat.grad_views[n].data_ptr()]
        fresh = all(p.grad is None for _, p in active)
        stash = None
        if fresh or foreign:
            if foreign and not fresh:                   # keep what was accumulated in our views
                stash = flat.grad.clone()
            flat.grad.zero_()
        eng.backward(*gs)
        for n, p in active:
            if p.grad is None:
                p.grad = flat.grad_views[n]
        if foreign:
            for n, p in foreign:
                p.grad.add_(flat.grad_views[n])
            if stash is not None:
                flat.grad.add_(stash)
        return None, None, None


class SAM2UNet(nn.Module):
    def __init__(self, checkpoint_path: str = "", *, model_cfg: str = "sam2_hiera_s.yaml", dtype="bf16",
                 gemm_backend: int = 0) -> None:
        super().__init__()
        self.cfg: TrunkConfig = trunk_config(model_cfg)
        self.compute_dtype = _DTYPES[dtype] if isinstance(dtype, str) else dtype
        self.gemm_backend = gemm_backend
        self.encoder = _Trunk(self.cfg)
        for p in self.encoder.parameters():            # SAM2UNet.py:146-147
            p.requires_grad = False
        for blk in self.encoder.blocks:                 # adapters are created after the freeze (SAM2UNet.py:148-151)
            for p in blk.prompt_learn.parameters():
                p.requires_grad = True
        dims = self.cfg.stage_dims
        self.rfb1, self.rfb2, self.rfb3, self.rfb4 = (_RFB(d, 64) for d in dims)
        self.up1, self.up2, self.up3, self.up4 = (_Up(128, 64) for _ in range(4))
        self.side1 = nn.Conv2d(64, 1, kernel_size=1)
        self.side2 = nn.Conv2d(64, 1, kernel_size=1)
        self.head = nn.Conv2d(64, 1, kernel_size=1)
        self.conv_units: List[_ConvSpec] = self._conv_units(dims)
        self._eng: Optional[Engine] = None
        self.flat: Optional[FlatParams] = None
        self._anchor = None
        self.register_load_state_dict_post_hook(lambda m, keys: m._invalidate())
        if len(checkpoint_path) > 0:
            self.load_hiera_checkpoint(checkpoint_path)

    @staticmethod
    def _conv_units(dims) -> List[_ConvSpec]:
        units = []
        for k, cin in enumerate(dims):
            r = f"rfb{k + 1}."
            units.append(_ConvSpec(r + "branch0.0.conv", r + "branch0.0.bn", cin, 1, 1, 1))
            for bi, ks in ((1, 3), (2, 5), (3, 7)):
                b = r + f"branch{bi}."
                units.append(_ConvSpec(b + "0.conv", b + "0.bn", cin, 1, 1, 1))
                units.append(_ConvSpec(b + "1.conv", b + "1.bn", 64, 1, ks, 1))
                units.append(_ConvSpec(b + "2.conv", b + "2.bn", 64, ks, 1, 1))
                units.append(_ConvSpec(b + "3.conv", b + "3.bn", 64, 3, 3, ks))
            units.append(_ConvSpec(r + "conv_cat.conv", r + "conv_cat.bn", 256, 3, 3, 1))
            units.append(_ConvSpec(r + "conv_res.conv", r + "conv_res.bn", cin, 1, 1, 1))
        for k in (1, 2, 3):
            u = f"up{k}.conv.double_conv."
            units.append(_ConvSpec(u + "0", u + "1", 128, 3, 3, 1))
            units.append(_ConvSpec(u + "3", u + "4", 64, 3, 3, 1))
        return units

    # ------------------------------------------------------------------------------------- checkpoints

    def load_hiera_checkpoint(self, path: str) -> None:
        """Load an upstream `sam2_hiera_*.pt` (dict with key "model", trunk keys `image_encoder.trunk.*`,
        /root/reference/sam2/build_sam.py:79-89 + SAM2UNet.py:132-144) into the frozen trunk, strictly."""
        sd = torch.load(path, map_location="cpu", weights_only=True)
        sd = sd["model"] if "model" in sd else sd
        pfx = "image_encoder.trunk."
        mine = {}
        for k, v in sd.items():
            if not k.startswith(pfx):
                continue
            k2 = k[len(pfx):]
            if k2.startswith("blocks."):
                parts = k2.split(".")
                k2 = ".".join(parts[:2] + ["block"] + parts[2:])
            mine["encoder." + k2] = v
        want = {n for n, _ in self.encoder.named_parameters() if ".prompt_learn." not in n}
        want = {"encoder." + n for n in want}
        missing, unexpected = sorted(want - set(mine)), sorted(set(mine) - want)
        if missing or unexpected:
            raise RuntimeError(f"trunk checkpoint mismatch: missing {missing[:5]}, unexpected {unexpected[:5]}")
        self.load_state_dict(mine, strict=False)

    # ----------------------------------------------------------------------------------------- engine

    def _invalidate(self):
        self._eng = None
        self.flat = None

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        self._invalidate()
        return out

    def _engine(self, device) -> Engine:
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.KernelError("SAM2UNet runs on CUDA (sm_100a) only: there is no CPU fallback; move the model "
                                   "and the batch to a B200")
        if self._eng is None or self._eng.device != device:
            p0 = next(self.parameters())
            if p0.device != device:
                raise RuntimeError(f"model is on {p0.device}, input on {device}")
            self.flat = FlatParams(self, device)
            self._anchor = torch.zeros((), device=device, requires_grad=True)
            self._eng = Engine(self.cfg, self, self.compute_dtype, device, self.gemm_backend)
        return self._eng

    def forward(self, x: torch.Tensor):
        eng = self._engine(x.device)
        if torch.is_grad_enabled() and self.training:
            B, S = x.shape[0], x.shape[-1]
            eng.tape_out_shapes = [(B, 1, S, S)] * 3
            return _SAM2UNetFn.apply(x, self._anchor, self)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and x.requires_grad:
            raise NotImplementedError("gradients in eval mode (frozen BatchNorm statistics) are not implemented")
        return eng.forward(x, self.training, save=False)

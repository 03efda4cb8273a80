"""Fused AdamW over the flat parameter buffer and the fused train step.

Reference: torch.optim.AdamW(lr=1e-3, weight_decay=5e-4, betas=(0.9, 0.999), eps=1e-8) over
`model.parameters()` where frozen / unused parameters have `grad is None` and are skipped
(/root/reference/train.py:48-52,74,83), CosineAnnealingLR stepped once per epoch (train.py:54,87).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib
from .ddp import GradSync
from .model import SAM2UNet


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.Optimizer-compatible (param_groups, LR schedulers, zero_grad) AdamW whose `step()` is ONE
    kernel over the model's flat fp32 master buffer (28 B/parameter of HBM traffic)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, *, model: SAM2UNet,
                 grad_scale: float = 1.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.model = model
        self.grad_scale = grad_scale
        self._t = 0
        self._m = self._v = self._hyper = self._staged = None

    def _ensure_state(self):
        flat = self.model.flat
        if flat is None:
            raise RuntimeError("FusedAdamW.step() before the first forward: the flat parameter buffer does not exist")
        if self._m is None or self._m.numel() != flat.n_active or self._m.device != flat.master.device:
            dev = flat.master.device
            self._m = torch.zeros(flat.n_active, dtype=torch.float32, device=dev)
            self._v = torch.zeros(flat.n_active, dtype=torch.float32, device=dev)
            self._hyper = torch.tensor([0.0, 1.0, 1.0, 1.0], dtype=torch.float32, device=dev)
            self._staged = None
        return flat

    def set_hyper(self):
        """Stage lr / gradient scale on the device when they changed (stream-ordered fills, no host buffer to race
        with; the bias-correction products beta^t advance on the device inside s2u_adamw)."""
        self._ensure_state()
        lr = float(self.param_groups[0]["lr"])
        if self._staged != (lr, self.grad_scale):
            self._hyper[0:1].fill_(lr)
            self._hyper[1:2].fill_(self.grad_scale)
            self._staged = (lr, self.grad_scale)

    def launch(self):
        """Enqueue the update kernel (capturable); hyper-parameters must have been staged by set_hyper()."""
        flat = self._ensure_state()
        g = self.param_groups[0]
        b1, b2 = g["betas"]
        st = torch.cuda.current_stream(flat.master.device).cuda_stream
        _lib.call("s2u_adamw", flat.master.data_ptr(), flat.grad.data_ptr(), self._m.data_ptr(), self._v.data_ptr(),
                  flat.n_active, self._hyper.data_ptr(), b1, b2, g["eps"], g["weight_decay"], st)
        flat.bump()

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        flat = self._ensure_state()
        active = [p for n, p in flat.params.items() if flat.offsets[n] < flat.n_active]
        if all(p.grad is None for p in active):
            return loss
        for n, p in flat.params.items():                # gradients delivered by something else than our backward
            if flat.offsets[n] < flat.n_active and p.grad is not None and \
                    p.grad.data_ptr() != flat.grad_views[n].data_ptr():
                flat.grad_views[n].copy_(p.grad)
        self.set_hyper()
        self.launch()
        self._t += 1
        return loss


def cosine_lr(epoch: int, t_max: int, base_lr: float = 1e-3, eta_min: float = 1e-7) -> float:
    """Closed form of CosineAnnealingLR(T_max, eta_min) after `epoch` scheduler steps (train.py:54,87)."""
    return eta_min + (base_lr - eta_min) * (1 + math.cos(math.pi * epoch / t_max)) / 2


class _Prefetch:
    """Host -> device copy of the NEXT batch on a copy stream, overlapped with the step that is running now.

    `put(tensors)` starts the copies into staging buffers; `take(tensors, dsts)` (called by the step that consumes the
    same host tensors) makes the compute stream wait for them and moves staging -> the step's static inputs with
    device-to-device copies.  Returns False when nothing matching was prefetched (the caller then copies directly)."""

    def __init__(self):
        self.stream = None
        self.stage = None
        self.src = None
        self.ready = None
        self.free = None

    def put(self, dev, tensors):
        if self.stream is None:
            self.stream = torch.cuda.Stream(dev)
        if self.stage is None or any(a.shape != b.shape for a, b in zip(self.stage, tensors)):
            self.stage = [torch.empty_like(t, device=dev) for t in tensors]
            self.free = None
        with torch.cuda.stream(self.stream):
            if self.free is not None:
                self.stream.wait_event(self.free)          # the previous step has finished reading the staging buffers
            for d, t in zip(self.stage, tensors):
                d.copy_(t, non_blocking=True)
            self.ready = torch.cuda.Event()
            self.ready.record(self.stream)
        self.src = tensors

    def take(self, tensors, dsts) -> bool:
        if self.src is None or len(self.src) != len(tensors) or any(a is not b for a, b in zip(self.src, tensors)):
            return False
        cur = torch.cuda.current_stream(dsts[0].device)
        cur.wait_event(self.ready)
        for d, st in zip(dsts, self.stage):
            d.copy_(st, non_blocking=True)
        self.free = torch.cuda.Event()
        self.free.record(cur)
        self.src = None
        return True


class TrainStep:
    """The whole step of train.py:66-86 — forward, three structure_loss terms, backward, (all-reduce,) AdamW —
    as one launch sequence without autograd, optionally captured in a CUDA graph.

        step = TrainStep(model, lr=1e-3, weight_decay=5e-4)
        loss = step(x, mask)          # device tensor [3]; `.sum().item()` only when you want to look at it
    """

    def __init__(self, model: SAM2UNet, lr=1e-3, weight_decay=5e-4, betas=(0.9, 0.999), eps=1e-8,
                 use_graph: bool = True, process_group=None, sync_grads: bool = True):
        self.model = model
        self.pg = process_group
        # sync_grads=False: a rank-local step without the all-reduce (profiling one rank must not issue collectives)
        self.world = (dist.get_world_size(process_group) if (process_group is not None or dist.is_initialized()) else 1) \
            if sync_grads else 1
        self.optim = FusedAdamW([p for p in model.parameters() if p.requires_grad], lr=lr, betas=betas, eps=eps,
                                weight_decay=weight_decay, model=model, grad_scale=1.0 / self.world)
        self.use_graph = use_graph
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._static = None
        self._warm = 0
        self._cache = {}
        self._pre = _Prefetch()

    def prefetch(self, x: torch.Tensor, mask: torch.Tensor) -> None:
        """Start copying the NEXT (pinned) host batch to the device while the current step runs; the following
        `step(x, mask)` with the same two tensors picks the copy up instead of copying again."""
        dev = next(self.model.parameters()).device
        if x.device.type == "cpu" and x.dtype == torch.float32 and mask.dtype == torch.float32 and \
                x.is_contiguous() and mask.is_contiguous():
            self._pre.put(dev, [x, mask])

    # the raw launch sequence (capturable: no host sync, no allocation outside torch's caching allocator)
    def _run(self, x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        model = self.model
        eng = model._engine(x.device)
        flat = model.flat
        dev = x.device
        B, S = x.shape[0], x.shape[-1]
        out, out1, out2 = eng.forward(x, True, save=True)
        st = torch.cuda.current_stream(dev).cuda_stream
        weit = torch.empty(B, S, S, dtype=torch.float32, device=dev)
        sums = torch.empty(3 * B * 2 + 3, dtype=torch.float64, device=dev)
        loss = torch.empty(3, dtype=torch.float32, device=dev)
        _lib.call("s2u_structure_loss_fwd", out.data_ptr(), out1.data_ptr(), out2.data_ptr(), mask.data_ptr(),
                  weit.data_ptr(), sums.data_ptr(), loss.data_ptr(), B, S, S, 3, st)
        g = [torch.empty_like(out) for _ in range(3)]
        _lib.call("s2u_structure_loss_bwd", out.data_ptr(), out1.data_ptr(), out2.data_ptr(), mask.data_ptr(),
                  weit.data_ptr(), sums.data_ptr(), 0, g[0].data_ptr(), g[1].data_ptr(), g[2].data_ptr(), B, S, S, 3,
                  st)
        flat.grad.zero_()
        sync = GradSync(flat.grad, self.pg)
        eng.backward(g[0], g[1], g[2], on_bucket=sync.on_bucket if self.world > 1 else None)
        sync.finish()
        self.optim.launch()
        return loss

    def __call__(self, x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        model = self.model
        if not model.training:
            model.train()
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise _lib.KernelError("TrainStep needs the model on a CUDA device (no CPU fallback)")
        x0, m0 = x, mask
        x = x.contiguous().float()
        mask = mask.contiguous().float()
        model._engine(dev)
        self.optim.set_hyper()
        if not self.use_graph:
            # host batches (pinned or pageable) are copied to the device here: this is the H2D leg of the step
            loss = self._run(x.to(dev, non_blocking=True), mask.to(dev, non_blocking=True))
        else:
            # the captured graph holds raw pointers into the flat parameter / gradient / shadow buffers: a rebuild of
            # those (load_state_dict, .to(), .float()) must re-capture, and the optimizer moments start over
            key = (tuple(x.shape), dev, id(model.flat), model.flat.master.data_ptr())
            if self._static is None or self._static[0] != key:
                # one captured graph per batch shape (a smaller last batch of an epoch alternates with the full one)
                if self._static is not None:
                    self._cache[self._static[0]] = (self._static, self._graph, self._warm)
                hit = self._cache.pop(key, None)
                if hit is not None:
                    self._static, self._graph, self._warm = hit
                else:
                    self._static = (key, torch.empty_like(x, device=dev), torch.empty_like(mask, device=dev), None)
                    self._graph = None
                    self._warm = 0
                stale = [k for k in self._cache if k[2:] != key[2:]]      # graphs of a rebuilt parameter buffer
                for k in stale:
                    del self._cache[k]
            _, sx, sm, sl = self._static
            if not self._pre.take([x0, m0], [sx, sm]):  # prefetched by the previous iteration, else copy now
                sx.copy_(x, non_blocking=True)
                sm.copy_(mask, non_blocking=True)
            if self._graph is None and self._warm < 2:
                loss = self._run(sx, sm)                # eager warm-up: allocator pools, smem attributes, tensor maps
                self._warm += 1
            elif self._graph is None:
                torch.cuda.synchronize(dev)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    sl = self._run(sx, sm)
                self._graph = graph
                self._static = (key, sx, sm, sl)
                graph.replay()
                loss = sl
            else:
                self._graph.replay()
                self.model.flat.bump()
                loss = sl
        self.optim._t += 1
        for p in self.model.flat.params.values():       # this path bypasses autograd: keep .grad unset
            p.grad = None
        return loss


class Predictor:
    """Inference forward (test.py:61, train.py:101) replayed from a CUDA graph: `out, out1, out2 = Predictor(model)(x)`.

    The eager `model(x)` call issues ~700 launches from Python per Hiera-L forward; at small batch that launch stream,
    not the GPU, is the bottleneck.  The graph is re-captured when the input shape or the weights change."""

    def __init__(self, model: SAM2UNet, use_graph: bool = True):
        self.model = model
        self.use_graph = use_graph
        self._key = None
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._sx = self._outs = None
        self._warm = 0
        self._pre = _Prefetch()

    def prefetch(self, x: torch.Tensor) -> None:
        """Start copying the NEXT (pinned) host batch while the current forward runs (see TrainStep.prefetch)."""
        dev = next(self.model.parameters()).device
        if x.device.type == "cpu" and x.dtype == torch.float32 and x.is_contiguous():
            self._pre.put(dev, [x])

    @torch.no_grad()
    def __call__(self, x: torch.Tensor):
        model = self.model
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise _lib.KernelError("Predictor needs the model on a CUDA device (no CPU fallback)")
        if model.training:
            model.eval()
        eng = model._engine(dev)
        x0 = x
        x = x.contiguous().float()
        if not self.use_graph:
            return eng.forward(x.to(dev, non_blocking=True), False, save=False)
        key = (tuple(x.shape), dev, id(model.flat), model.flat.master.data_ptr(), model.flat.version)
        if key != self._key:
            self._key, self._graph, self._warm = key, None, 0
            self._sx = torch.empty_like(x, device=dev)
        if not self._pre.take([x0], [self._sx]):
            self._sx.copy_(x, non_blocking=True)
        if self._graph is None and self._warm < 2:
            self._warm += 1
            return eng.forward(self._sx, False, save=False)
        if self._graph is None:
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._outs = eng.forward(self._sx, False, save=False)
            self._graph = graph
        self._graph.replay()
        return self._outs

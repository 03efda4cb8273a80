"""`structure_loss(pred, mask)` — drop-in for /root/reference/train.py:21-29 on the fused CUDA kernels.

`structure_loss` keeps the reference signature (one prediction map, differentiable w.r.t. `pred`);
`structure_loss3` evaluates the three training terms of train.py:76-79 in one pass so that the 31x31
box filter of the shared mask is computed once instead of three times.
"""
from __future__ import annotations

import torch

from . import _lib


def _check(pred: torch.Tensor, mask: torch.Tensor):
    if pred.device.type != "cuda":
        raise _lib.KernelError("structure_loss runs on CUDA only (no CPU fallback)")
    if pred.dim() != 4 or pred.shape[1] != 1 or pred.shape != mask.shape:
        raise ValueError("structure_loss expects pred and mask of shape [B,1,H,W]")


class _StructureLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mask, *preds):
        n = len(preds)
        B, _, H, W = mask.shape
        dev = mask.device
        preds = [p.contiguous().float() for p in preds]
        mask = mask.contiguous().float()
        weit = torch.empty(B, H, W, dtype=torch.float32, device=dev)
        sums = torch.empty(n * B * 2 + n, dtype=torch.float64, device=dev)
        loss = torch.empty(n, dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        ptrs = [p.data_ptr() for p in preds] + [0] * (3 - n)
        _lib.call("s2u_structure_loss_fwd", *ptrs, mask.data_ptr(), weit.data_ptr(), sums.data_ptr(), loss.data_ptr(),
                  B, H, W, n, st)
        ctx.save_for_backward(mask, weit, sums, *preds)
        ctx.n = n
        return loss

    @staticmethod
    def backward(ctx, gloss):
        mask, weit, sums, *preds = ctx.saved_tensors
        n = ctx.n
        B, _, H, W = mask.shape
        grads = [torch.empty_like(p) for p in preds]
        gscale = gloss.contiguous().float()
        st = torch.cuda.current_stream(mask.device).cuda_stream
        pp = [p.data_ptr() for p in preds] + [0] * (3 - n)
        gp = [g.data_ptr() for g in grads] + [0] * (3 - n)
        _lib.call("s2u_structure_loss_bwd", *pp, mask.data_ptr(), weit.data_ptr(), sums.data_ptr(), gscale.data_ptr(),
                  *gp, B, H, W, n, st)
        return (None, *grads)


def structure_loss(pred: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """train.py:21-29: scalar loss of one logits map against a {0,1} mask."""
    _check(pred, mask)
    return _StructureLossFn.apply(mask, pred)[0]


def structure_loss3(pred0: torch.Tensor, pred1: torch.Tensor, pred2: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """Per-head losses [3] of train.py:76-78 (their sum is the training loss, train.py:79)."""
    for p in (pred0, pred1, pred2):
        _check(p, mask)
    return _StructureLossFn.apply(mask, pred0, pred1, pred2)

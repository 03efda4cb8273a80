"""Data parallelism: the batch is sharded over ranks (one process per GPU) and the ONLY collective on the path is
a bucketed all-reduce (sum) of the flat gradient buffer, launched bucket by bucket while backward is still running.

The reference has no distributed code at all; semantics follow PyTorch DDP:
gradients are averaged over ranks (the 1/world factor is folded into the AdamW kernel's gradient scale), BatchNorm
statistics stay per rank, unused parameters (`up4.*`) are simply not part of the reduced range.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


class GradSync:
    """Launches `all_reduce(flat_grad[lo:hi])` as soon as the engine reports a bucket final (Engine.backward's
    `on_bucket`), on NCCL's own stream, and joins them before the optimizer runs."""

    def __init__(self, flat_grad: torch.Tensor, process_group=None):
        self.grad = flat_grad
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.works: List = []
        self.ranges: List = []

    def on_bucket(self, lo: int, hi: int) -> None:
        self.ranges.append((lo, hi))
        if self.world > 1 and hi > lo:
            self.works.append(dist.all_reduce(self.grad[lo:hi], op=dist.ReduceOp.SUM, group=self.pg, async_op=True))

    def finish(self) -> None:
        for w in self.works:
            w.wait()                       # stream-level join for NCCL, blocking for gloo
        self.works.clear()
        self.ranges.clear()


def shard_batch(x: torch.Tensor, rank: Optional[int] = None, world: Optional[int] = None) -> torch.Tensor:
    """Rank's contiguous slice of a global batch (weak scaling keeps the per-rank size fixed instead)."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    per = x.shape[0] // world
    return x[rank * per:(rank + 1) * per]

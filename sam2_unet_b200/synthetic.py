"""Seeded synthetic workload of the benchmarks and smoke runs: N(0,1) images and binary masks made of three random
discs (non-trivial 31x31 boundary weights for structure_loss, /root/reference/train.py:21-29).  `correlated=True`
shifts the image by the mask so that a few optimisation steps can fit it (the pre-fit of the bf16 mask criteria)."""
from __future__ import annotations

from typing import Tuple

import torch


def synthetic_batch(B: int, S: int, seed: int = 0, correlated: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (image [B,3,S,S] fp32, mask [B,1,S,S] fp32 in {0,1}) on the CPU, a pure function of (B, S, seed)."""
    g = torch.Generator().manual_seed(1234 + seed)
    x = torch.randn(B, 3, S, S, generator=g)
    yy, xx = torch.meshgrid(torch.arange(S, dtype=torch.float32), torch.arange(S, dtype=torch.float32), indexing="ij")
    mask = torch.zeros(B, 1, S, S)
    for b in range(B):
        for _ in range(3):
            cy, cx = (torch.rand(2, generator=g) * S).tolist()
            r = (0.08 + 0.2 * torch.rand(1, generator=g).item()) * S
            mask[b, 0] = torch.maximum(mask[b, 0], ((yy - cy) ** 2 + (xx - cx) ** 2 <= r * r).float())
    if correlated:
        x = 0.5 * x + 2 * mask - 0.5
    return x, mask

"""Per-axis tables for the separable bilinear resampling kernels (csrc/resample.cu).

Index/weight conventions follow ATen's `upsample_bilinear2d`, which the reference reaches through
nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True) (/root/reference/SAM2UNet.py:35) and
F.interpolate(scale_factor=s, mode="bilinear") with align_corners=False (SAM2UNet.py:168-172).
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch


def axis_forward(n_in: int, n_out: int, align_corners: bool, scale: float | None = None):
    """(i0, i1, w0, w1) per output index, computed in float32 like ATen's area_pixel_compute_source_index."""
    o = np.arange(n_out, dtype=np.float32)
    if align_corners:
        r = np.float32((n_in - 1) / (n_out - 1)) if n_out > 1 else np.float32(0)
        src = o * r
    else:
        # with scale_factor given (recompute_scale_factor unset) ATen uses 1/scale_factor
        r = np.float32(1.0 / scale) if scale is not None else np.float32(n_in / n_out)
        src = np.maximum((o + np.float32(0.5)) * r - np.float32(0.5), np.float32(0))
    i0 = np.minimum(src.astype(np.int64), n_in - 1)
    i1 = np.minimum(i0 + 1, n_in - 1)
    w1 = (src - i0.astype(np.float32)).astype(np.float32)
    w1 = np.clip(w1, 0, 1)
    w0 = (np.float32(1) - w1).astype(np.float32)
    return i0.astype(np.int32), i1.astype(np.int32), w0, w1


def axis_backward(n_in: int, n_out: int, fwd) -> Tuple[np.ndarray, np.ndarray, int]:
    """Transpose of the forward operator as a padded adjacency list: idx/w of shape [n_in, taps]."""
    i0, i1, w0, w1 = fwd
    lists = [[] for _ in range(n_in)]
    for o in range(n_out):
        if i0[o] == i1[o]:
            lists[i0[o]].append((o, float(w0[o]) + float(w1[o])))
        else:
            lists[i0[o]].append((o, float(w0[o])))
            lists[i1[o]].append((o, float(w1[o])))
    taps = max(1, max(len(l) for l in lists))
    idx = np.zeros((n_in, taps), dtype=np.int32)
    w = np.zeros((n_in, taps), dtype=np.float32)
    for i, l in enumerate(lists):
        for k, (o, wt) in enumerate(l):
            idx[i, k] = o
            w[i, k] = wt
    return idx, w, taps


class ResampleTables:
    """Device-resident tables for one (n_in -> n_out) square resampling."""

    _cache: Dict[tuple, "ResampleTables"] = {}

    def __init__(self, n_in: int, n_out: int, align_corners: bool, scale, device):
        f = axis_forward(n_in, n_out, align_corners, scale)
        bi, bw, taps = axis_backward(n_in, n_out, f)
        self.n_in, self.n_out, self.taps = n_in, n_out, taps
        self.i0, self.i1 = (torch.from_numpy(a).to(device) for a in f[:2])
        self.w0, self.w1 = (torch.from_numpy(a).to(device) for a in f[2:])
        self.bidx = torch.from_numpy(bi).to(device)
        self.bw = torch.from_numpy(bw).to(device)

    @classmethod
    def get(cls, n_in, n_out, align_corners, scale, device) -> "ResampleTables":
        key = (n_in, n_out, align_corners, scale, str(device))
        if key not in cls._cache:
            cls._cache[key] = cls(n_in, n_out, align_corners, scale, device)
        return cls._cache[key]

    def fwd_ptrs(self):
        return (self.i0.data_ptr(), self.i1.data_ptr(), self.w0.data_ptr(), self.w1.data_ptr())

    def bwd_ptrs(self):
        return (self.bidx.data_ptr(), self.bw.data_ptr(), self.taps)

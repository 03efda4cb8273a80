"""Training-time input pipeline on the device: the reference's `FullDataset(mode="train")` transform
(/root/reference/dataset.py:288-313: ToTensor -> ResizeLongestSideAndPad -> RandomRotate -> ToGray ->
ColorAugmentations -> GaussianBlur -> Normalize) with the pixel work in CUDA kernels (csrc/augment.cu).

The random DECISIONS are drawn here, on the host, from Python's `random` in exactly the order the reference's transform
classes consume it - so `random.seed(s)` reproduces the reference's sample for the same image - and only the uint8 image
and label cross PCIe (1 + 1/3 bytes per source pixel instead of 16 bytes per output pixel).  No CPU fallback."""
from __future__ import annotations

import math
import random
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import _lib

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)
_OPS = {"gray": 0, "brightness": 1, "contrast": 2, "saturation": 3, "hue": 4, "gamma": 5, "normalize": 6}


def draw_train_params(size: int, H: int, W: int) -> Dict:
    """The decisions of one training sample, consuming `random` like dataset.py does (line numbers of the reference)."""
    p: Dict = {}
    if random.random() < 0.5:                                    # :56 pad branch of ResizeLongestSideAndPad
        sh, sw = random.uniform(1.0, 1.5), random.uniform(1.0, 1.5)          # :58-59
        pad_h, pad_w = int(round(H * sh)) - H, int(round(W * sw)) - W        # :62-67
        pad_top = random.randint(0, pad_h)                                   # :70
        pad_left = random.randint(0, pad_w)                                  # :72
        # :76 passes [left, right, top, bottom] where torchvision's F.pad expects [left, top, right, bottom]; what the
        # reference really computes is left = pad_left, top = pad_right, right = pad_top, bottom = pad_bottom
        left, top, right, bottom = pad_left, pad_w - pad_left, pad_top, pad_h - pad_top
        p["geom"] = (0, top, left, H + top + bottom, W + left + right)
    else:                                                        # :81 crop branch
        sh, sw = random.uniform(0.5, 1.0), random.uniform(0.5, 1.0)          # :83-84
        nh, nw = max(1, int(round(H * sh))), max(1, int(round(W * sw)))      # :87-92
        y1 = random.randint(0, H - nh)                                       # :95
        x1 = random.randint(0, W - nw)                                       # :96
        p["geom"] = (1, y1, x1, nh, nw)
    p["rot"] = 0
    if random.random() < 0.75:                                   # :161 RandomRotate
        p["rot"] = random.choice([90, 180, 270]) // 90
    p["gray"] = random.random() < 0.5                            # :191 ToGray
    ops: List[Tuple[str, float]] = []
    if random.random() < 0.8:                                    # :217 ColorAugmentations
        choice = random.randint(0, 3)
        if choice == 0:
            b, c = random.uniform(0.5, 1.5), random.uniform(0.5, 1.5)
            ops = [("brightness", b), ("contrast", c)]
        elif choice == 1:
            b, c = random.uniform(0.5, 1.5), random.uniform(0.5, 1.5)
            s, h = random.uniform(0.5, 1.5), random.uniform(-0.5, 0.5)
            ops = [("brightness", b), ("contrast", c), ("saturation", s), ("hue", h)]
        elif choice == 2:
            s, h = random.uniform(0.5, 1.5), random.uniform(-0.5, 0.5)
            ops = [("saturation", s), ("hue", h)]
        else:
            ops = [("gamma", random.uniform(0.5, 1.5))]
    p["color"] = ops
    p["blur"] = 0
    if random.random() < 0.2:                                    # :278 GaussianBlur
        p["blur"] = random.choice([3, 5])
    return p


def _gaussian_weights(k: int) -> np.ndarray:
    """torchvision _get_gaussian_kernel1d with sigma = 0.15 k + 0.35, in fp32 like the reference"""
    sigma = np.float32(k * 0.15 + 0.35)
    x = np.linspace(-(k - 1) * 0.5, (k - 1) * 0.5, k, dtype=np.float32)
    pdf = np.exp(np.float32(-0.5) * (x / sigma) ** 2).astype(np.float32)
    return (pdf / pdf.sum(dtype=np.float32)).astype(np.float32)


class TrainAugment:
    """`aug(image_u8, label_u8) -> {"image": [3,S,S] fp32, "label": [1,S,S] fp32}` on the device.

    image_u8: uint8 [H, W, 3] (RGB, what `PIL.Image.convert("RGB")` holds), label_u8: uint8 [H, W] ("L"); numpy arrays
    or torch tensors on the host or on the device.  Draws from Python's `random` exactly like the reference."""

    def __init__(self, size: int, device="cuda"):
        self.size = int(size)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.KernelError("TrainAugment runs on CUDA only (no CPU fallback)")
        _lib.load()
        self._ws = torch.zeros(1, dtype=torch.float64, device=self.device)
        self._mean = (_lib.F * 3)(*MEAN)
        self._std = (_lib.F * 3)(*STD)

    def _dev_u8(self, a) -> torch.Tensor:
        t = torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a
        if t.dtype != torch.uint8:
            raise ValueError("expected uint8 pixels")
        return t.to(self.device, non_blocking=True).contiguous()

    def __call__(self, image_u8, label_u8, params: Dict = None) -> Dict[str, torch.Tensor]:
        img, lab = self._dev_u8(image_u8), self._dev_u8(label_u8)
        if img.dim() != 3 or img.shape[2] != 3 or lab.shape != img.shape[:2]:
            raise ValueError("expected image [H, W, 3] and label [H, W]")
        H, W = lab.shape
        S = self.size
        p = params if params is not None else draw_train_params(S, H, W)
        st = torch.cuda.current_stream(self.device).cuda_stream
        mode, oy, ox, ph, pw = p["geom"]
        scale = S / max(ph, pw)                                              # dataset.py:106-113
        nh, nw = int(round(ph * scale)), int(round(pw * scale))
        pt, pl = (S - nh) // 2, (S - nw) // 2                                # :126-131
        tmp = torch.empty(3 * ph * nw, dtype=torch.float32, device=self.device)
        out = torch.empty(3, S, S, dtype=torch.float32, device=self.device)
        olab = torch.empty(1, S, S, dtype=torch.float32, device=self.device)
        _lib.call("s2u_aug_resize_pad", img.data_ptr(), lab.data_ptr(), H, W, mode, oy, ox, ph, pw, S, nh, nw, pl, pt,
                  tmp.data_ptr(), out.data_ptr(), olab.data_ptr(), st)
        if p["rot"]:
            o2, l2 = torch.empty_like(out), torch.empty_like(olab)
            _lib.call("s2u_aug_rot90", out.data_ptr(), o2.data_ptr(), 3, S, p["rot"], st)
            _lib.call("s2u_aug_rot90", olab.data_ptr(), l2.data_ptr(), 1, S, p["rot"], st)
            out, olab = o2, l2

        def color(op: str, f: float = 0.0):
            _lib.call("s2u_aug_color", out.data_ptr(), S, _OPS[op], float(f), float(1.0 - float(f)), self._ws.data_ptr(),
                      self._mean, self._std, st)

        if p["gray"]:
            color("gray")
        for name, f in p["color"]:
            color(name, f)
        if p["blur"]:
            k = p["blur"]
            w = _gaussian_weights(k)
            o2 = torch.empty_like(out)
            _lib.call("s2u_aug_blur", out.data_ptr(), o2.data_ptr(), S, k, (_lib.F * k)(*[float(v) for v in w]), st)
            out = o2
        color("normalize")
        return {"image": out, "label": olab}

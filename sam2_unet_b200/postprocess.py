"""Post-forward inference tail on the device (reference: test.py:66-76, train.py:103-112).

    res_padded, _, _ = model(image)                       # [1,1,S,S] logits of the letterboxed image
    png = infer_tail(res_padded, padding, gt.shape)       # uint8 [H,W] on the device: crop, bilinear resize to the
                                                          # ground-truth size, sigmoid, min-max normalise, x255

The reference moves the fp32 map to the host and finishes in numpy; here the whole tail is two kernel launches and one
byte per pixel crosses PCIe.  No CPU fallback: CPU tensors raise.
"""
from typing import Dict, Sequence, Tuple

import torch

from . import _lib

_ws: Dict[Tuple[int, int], torch.Tensor] = {}


def _workspace(dev: torch.device, stream: int) -> torch.Tensor:
    key = (dev.index or 0, stream)
    ws = _ws.get(key)
    if ws is None:
        ws = _ws[key] = torch.empty(2, dtype=torch.int32, device=dev)
        _lib.call("s2u_infer_tail_init", ws.data_ptr(), stream)
    return ws


def infer_tail(res_padded: torch.Tensor, padding: Sequence[int], out_hw: Sequence[int]) -> torch.Tensor:
    """res_padded: [1,1,S,S] / [S,S] fp32 logits on a CUDA device; padding = (left, top, right, bottom) as produced by
    the reference's TestDataset; out_hw = ground-truth (height, width).  Returns a uint8 [H,W] device tensor."""
    if res_padded.device.type != "cuda":
        raise _lib.KernelError("infer_tail needs a CUDA tensor (no CPU fallback)")
    x = res_padded.detach()
    if x.dim() == 4:
        if x.shape[0] != 1 or x.shape[1] != 1:
            raise ValueError("infer_tail post-processes one single-channel map at a time, like test.py")
        x = x[0, 0]
    if x.dim() != 2 or x.shape[0] != x.shape[1]:
        raise ValueError("expected a square [S,S] logit map")
    x = x.contiguous().float()
    left, top, right, bottom = (int(v) for v in padding)
    H, W = int(out_hw[0]), int(out_hw[1])
    out = torch.empty(H, W, dtype=torch.uint8, device=x.device)
    st = torch.cuda.current_stream(x.device).cuda_stream
    _lib.call("s2u_infer_tail", x.data_ptr(), x.shape[0], left, top, right, bottom, H, W,
              _workspace(x.device, st).data_ptr(), out.data_ptr(), st)
    return out


_MEAN = (0.485, 0.456, 0.406)        # dataset.py:389 (NormalizeImage)
_STD = (0.229, 0.224, 0.225)


def preprocess_image(img: torch.Tensor, size: int):
    """Test-time input path of the reference (dataset.py:336-407) on the device: `img` uint8 [H,W,3] RGB CUDA tensor ->
    (x fp32 [1,3,size,size], padding [left, top, right, bottom]) exactly like TestDataset.load_data's image / padding:
    /255, antialiased bilinear resize of the longest side to `size`, centred zero padding, ImageNet normalisation."""
    import ctypes
    if img.device.type != "cuda":
        raise _lib.KernelError("preprocess_image needs a CUDA tensor (no CPU fallback)")
    if img.dtype != torch.uint8 or img.dim() != 3 or img.shape[2] != 3:
        raise ValueError("expected a uint8 [H,W,3] image")
    img = img.contiguous()
    H, W = int(img.shape[0]), int(img.shape[1])
    scale = size / max(H, W)                                 # dataset.py:364-371
    new_h, new_w = int(round(H * scale)), int(round(W * scale))
    pad_h, pad_w = size - new_h, size - new_w
    pad_top, pad_left = pad_h // 2, pad_w // 2
    padding = [pad_left, pad_top, pad_w - pad_left, pad_h - pad_top]
    tmp = torch.empty(3 * H * new_w, dtype=torch.float32, device=img.device)
    out = torch.empty(1, 3, size, size, dtype=torch.float32, device=img.device)
    mean = (ctypes.c_float * 3)(*_MEAN)
    std = (ctypes.c_float * 3)(*_STD)
    st = torch.cuda.current_stream(img.device).cuda_stream
    _lib.call("s2u_preprocess", img.data_ptr(), H, W, size, new_h, new_w, pad_left, pad_top,
              ctypes.cast(mean, ctypes.c_void_p).value, ctypes.cast(std, ctypes.c_void_p).value, tmp.data_ptr(),
              out.data_ptr(), st)
    return out, padding

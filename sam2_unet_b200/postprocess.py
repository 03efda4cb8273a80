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

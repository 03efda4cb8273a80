"""ctypes binding of the C-ABI library (include/sam2unet_b200.h).

There is no fallback: if the shared library is missing or a kernel launch fails, this raises.  The
library is built in-tree by `python -m sam2_unet_b200.build` (or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_double, c_float, c_int, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsam2unet_b200.so")

P, I, L, F = c_void_p, c_int, c_longlong, c_float

# name -> argument ctypes, mirrors include/sam2unet_b200.h (tests check both against the .so)
SIGNATURES = {
    "s2u_gemm": [P, I, P, I, P, I, I, I, I, P, P, I, P, I, P, I, I, I, I, P],
    "s2u_gemm_simt_fallbacks": [I],
    "s2u_gemm_wgrad": [P, I, P, I, P, I, L, I, I, I, I, I, P],
    "s2u_gemm_wgrad_pair": [P, I, P, I, P, I, I, I, P, I, P, I, P, I, I, I, L, I, P],
    "s2u_colsum": [P, I, P, L, I, I, P],
    "s2u_layernorm_fwd": [P, P, P, P, P, P, L, I, F, I, I, P],
    "s2u_layernorm_ws_floats": [I],
    "s2u_layernorm_bwd": [P, P, P, P, P, P, P, P, P, P, P, I, L, I, I, I, P],
    "s2u_adapter_supported": [I],
    "s2u_adapter_ln_fwd": [P, P, P, P, P, P, P, F, P, P, P, P, P, P, P, L, I, P],
    "s2u_adapter_ln_bwd": [P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, L, I, P],
    "s2u_dgelu_mul": [P, P, P, L, I, P],
    "s2u_add": [P, P, P, L, I, P],
    "s2u_maxpool2_fwd": [P, P, I, I, I, I, I, P],
    "s2u_maxpool2_bwd": [P, P, P, I, I, I, I, I, P],
    "s2u_cast": [P, P, I, I, I, I, P],
    "s2u_refresh_shadows": [P, P, I, I, P],
    "s2u_win_attn_fwd": [P, P, P, P, I, I, I, I, I, I, I, I, P],
    "s2u_set_attn_backend": [I],
    "s2u_win_attn_bwd": [P, P, P, P, P, P, P, I, I, I, I, I, I, I, I, P],
    "s2u_patch_embed": [P, P, P, P, P, I, P, I, I, I, I, P],
    "s2u_patch_im2col": [P, P, I, I, P],
    "s2u_im2col": [P, I, P, I, I, I, I, I, I, I, I, I, I, I, P],
    "s2u_conv_igemm_supported": [I, I, I, I],
    "s2u_conv_igemm": [P, I, I, I, I, I, P, I, I, I, I, P, I, P, P, I, I, P, P],
    "s2u_conv_igemm_bn": [P, I, I, I, I, I, P, I, I, I, P, I, P, P, P, P, P, P, P, P, P, P, F, F, P],
    "s2u_conv_wgrad": [P, I, P, I, P, I, I, I, I, I, I, I, I, P],
    "s2u_conv_weight_pack": [P, P, P, I, I, I, I, I, P],
    "s2u_bn_ws_doubles": [I],
    "s2u_bn_stats": [P, I, P, L, I, I, P],
    "s2u_bn_stats_finalize": [P, I, P, P, P, P, P, P, P, P, P, P, L, I, F, F, I, P],
    "s2u_bn_finalize": [P, P, P, P, P, P, P, P, P, P, L, I, F, F, I, P],
    "s2u_bn_apply": [P, I, P, P, P, I, P, I, L, I, I, I, P],
    "s2u_relu_bwd": [P, I, P, I, P, I, L, I, I, P],
    "s2u_bn_bwd": [P, I, P, I, P, I, P, P, P, P, P, P, P, P, P, I, L, I, I, P],
    "s2u_resample_fwd": [P, I, P, I, I, I, I, I, I, I, P, P, P, P, P, P, P, P, I, P],
    "s2u_resample_bwd": [P, I, P, I, I, I, I, I, I, I, P, P, I, P, P, I, I, P],
    "s2u_resample1_fwd": [P, P, I, I, I, I, I, P, P, P, P, P, P, P, P, P],
    "s2u_resample1_bwd": [P, P, I, I, I, I, I, P, P, I, P, P, I, P],
    "s2u_head_fwd": [P, I, P, P, P, L, I, I, P],
    "s2u_head_bwd": [P, I, P, P, P, I, I, P, P, L, I, I, P],
    "s2u_structure_loss_fwd": [P, P, P, P, P, P, P, I, I, I, I, P],
    "s2u_structure_loss_bwd": [P, P, P, P, P, P, P, P, P, P, I, I, I, I, P],
    "s2u_infer_tail_init": [P, P],
    "s2u_infer_tail": [P, I, I, I, I, I, I, I, P, P, P],
    "s2u_seg_counts": [P, P, L, F, P, P],
    "s2u_cc_label": [P, F, I, I, P, P, P, I, P],
    "s2u_cc_stats": [P, P, L, P, P, P, P, I, P, P],
    "s2u_preprocess": [P, I, I, I, I, I, I, I, P, P, P, P, P],
    "s2u_aug_resize_pad": [P, P, I, I, I, I, I, I, I, I, I, I, I, I, P, P, P, P],
    "s2u_aug_rot90": [P, P, I, I, I, P],
    "s2u_aug_color": [P, I, I, F, F, P, P, P, P],
    "s2u_aug_blur": [P, P, I, I, P, P],
    "s2u_adamw": [P, P, P, P, L, P, F, F, F, F, P],
}

_lib = None
_launches = 0


class KernelError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load the library once; raise (never fall back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise KernelError(
                f"{LIB_PATH} not found: build it with `python -m sam2_unet_b200.build` — there is no CPU or "
                "library fallback for the SAM2-UNet hot path")
        lib = ctypes.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = c_int
        _lib = lib
    return _lib


def _describe(rc: int) -> str:
    if rc == -1:
        return "invalid argument"
    if rc == -2:
        return "unsupported shape / alignment for this kernel"
    if rc <= -100:
        return f"cuTensorMapEncodeTiled failed with CUresult {-rc - 100}"
    return f"CUDA error {rc}"


_profile = None      # when profiling: list of (name, start_event, end_event, args)


def call(name: str, *args) -> None:
    """Invoke a C-ABI entry point and raise on a non-zero status."""
    global _launches
    if _profile is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(load(), name)(*args)
        e1.record()
        _profile.append((name, e0, e1, args))
    else:
        rc = getattr(load(), name)(*args)
    _launches += 1
    if rc != 0:
        raise KernelError(f"{name} failed: {_describe(rc)}")


def profile_begin() -> None:
    """Bracket every following C-ABI call with CUDA events on the current stream (eager runs only)."""
    global _profile
    _profile = []


def profile_end():
    """-> list of (name, milliseconds, args) for the calls made since profile_begin()."""
    global _profile
    import torch
    torch.cuda.synchronize()
    rec, _profile = _profile, None
    return [(n, e0.elapsed_time(e1), a) for n, e0, e1, a in rec]


def launch_count() -> int:
    """Number of C-ABI calls made so far (each launches at least one kernel of this package)."""
    return _launches

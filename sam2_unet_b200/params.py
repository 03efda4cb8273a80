"""Parameter containers with the reference's state-dict layout.

These nn.Modules only HOLD parameters and buffers under the exact key names of the reference
(/root/reference/SAM2UNet.py:129-162 and sam2/modeling/backbones/hieradet.py:200-259), so that
`.pth` files interchange with `load_state_dict(strict=True)` (train.py:46, test.py:45).  Their
`forward` is never used: the compute runs in `engine.Engine` on the CUDA kernels.
"""
from __future__ import annotations

import zlib

import torch
import torch.nn as nn

from .config import TrunkConfig


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - containers are not callable compute
        raise RuntimeError("parameter container: compute runs in sam2_unet_b200.engine, not here")


class _PatchEmbed(_Holder):                       # backbones/utils.py:58-88
    def __init__(self, embed_dim):
        super().__init__()
        self.proj = nn.Conv2d(3, embed_dim, kernel_size=7, stride=4, padding=3)


class _Attn(_Holder):                             # hieradet.py:35-54
    def __init__(self, dim, dim_out):
        super().__init__()
        self.qkv = nn.Linear(dim, dim_out * 3)
        self.proj = nn.Linear(dim_out, dim_out)


class _MLP(_Holder):                              # sam2_utils.py:109-125
    def __init__(self, dim, hidden):
        super().__init__()
        self.layers = nn.ModuleList([nn.Linear(dim, hidden), nn.Linear(hidden, dim)])


class _Block(_Holder):                            # hieradet.py:84-130
    def __init__(self, dim, dim_out):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attn(dim, dim_out)
        self.norm2 = nn.LayerNorm(dim_out, eps=1e-6)
        self.mlp = _MLP(dim_out, int(dim_out * 4.0))
        if dim != dim_out:
            self.proj = nn.Linear(dim, dim_out)


class _AdapterBlock(_Holder):                     # SAM2UNet.py:52-65
    def __init__(self, dim, dim_out):
        super().__init__()
        self.block = _Block(dim, dim_out)
        self.prompt_learn = nn.Sequential(nn.Linear(dim, 32), nn.GELU(), nn.Linear(32, dim), nn.GELU())


class _Trunk(_Holder):                            # hieradet.py:170-259
    def __init__(self, cfg: TrunkConfig):
        super().__init__()
        self.patch_embed = _PatchEmbed(cfg.embed_dim)
        self.pos_embed = nn.Parameter(torch.zeros(1, cfg.embed_dim, *cfg.window_pos_embed_bkg_spatial_size))
        self.pos_embed_window = nn.Parameter(torch.zeros(1, cfg.embed_dim, cfg.window_spec[0], cfg.window_spec[0]))
        self.blocks = nn.Sequential(*[_AdapterBlock(b.dim, b.dim_out) for b in cfg.blocks])


class _ConvBN(_Holder):                           # SAM2UNet.py:68-86 (BasicConv2d: conv -> bn, no activation)
    def __init__(self, cin, cout, kernel_size, padding=0, dilation=1):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, kernel_size=kernel_size, padding=padding, dilation=dilation, bias=False)
        self.bn = nn.BatchNorm2d(cout)


class _RFB(_Holder):                              # SAM2UNet.py:89-125
    def __init__(self, cin, cout=64):
        super().__init__()
        self.branch0 = nn.Sequential(_ConvBN(cin, cout, 1))
        for name, k, d in (("branch1", 3, 3), ("branch2", 5, 5), ("branch3", 7, 7)):
            setattr(self, name, nn.Sequential(
                _ConvBN(cin, cout, 1),
                _ConvBN(cout, cout, (1, k), padding=(0, k // 2)),
                _ConvBN(cout, cout, (k, 1), padding=(k // 2, 0)),
                _ConvBN(cout, cout, 3, padding=d, dilation=d)))
        self.conv_cat = _ConvBN(4 * cout, cout, 3, padding=1)
        self.conv_res = _ConvBN(cin, cout, 1)


class _DoubleConv(_Holder):                       # SAM2UNet.py:9-26
    def __init__(self, cin, cout, mid):
        super().__init__()
        self.double_conv = nn.Sequential(
            nn.Conv2d(cin, mid, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(mid), nn.ReLU(inplace=True),
            nn.Conv2d(mid, cout, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class _Up(_Holder):                               # SAM2UNet.py:29-49
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = _DoubleConv(cin, cout, cin // 2)


def trainable_filter(name: str) -> bool:
    """True for parameters the reference trains: everything except the trunk's ORIGINAL parameters
    (SAM2UNet.py:146-151 freezes the trunk, then adds the adapters)."""
    return (not name.startswith("encoder.")) or (".prompt_learn." in name)


@torch.no_grad()
def fill_deterministic_(module: nn.Module, seed: int = 0) -> None:
    """Overwrite every parameter and buffer with values that depend only on (seed, key name, shape).

    Used by the parity tests and fixtures: the same call on the reference model and on this package's
    model gives bit-identical weights on any machine (torch CPU generator).  Norm affine, positional
    embeddings and BN running statistics are randomised too (they are trivial at the reference's init,
    SURVEY.md section 8c), so that every term of the forward is exercised.
    """
    sd = module.state_dict()
    for name in sorted(sd.keys()):
        t = sd[name]
        g = torch.Generator(device="cpu")
        g.manual_seed((zlib.crc32(name.encode()) + 7919 * seed) % (2 ** 31 - 1))
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            t.zero_()
            continue
        shape = tuple(t.shape)
        r = torch.randn(shape, generator=g, dtype=torch.float32)
        is_norm = (".norm1." in name or ".norm2." in name or ".bn." in name
                   or (".double_conv." in name and t.dim() == 1))
        if leaf == "running_mean":
            v = 0.1 * r
        elif leaf == "running_var":
            v = 1.0 + 0.2 * r.abs()
        elif is_norm and leaf == "weight":
            v = 1.0 + 0.1 * r
        elif is_norm and leaf == "bias":
            v = 0.1 * r
        elif name.endswith("pos_embed") or name.endswith("pos_embed_window"):
            v = 0.1 * r
        elif t.dim() >= 2:
            fan_in = t[0].numel()
            v = r * (1.0 / fan_in) ** 0.5
        else:                                   # linear / conv biases
            v = 0.05 * r
        t.copy_(v.to(t.dtype))

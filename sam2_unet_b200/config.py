"""Trunk configuration: the `trunk:` stanza of the reference's Hydra yamls, without Hydra.

The reference builds its Hiera trunk through Hydra from `sam2_configs/sam2_hiera_{t,s,b+,l}.yaml`
(only lines 9-16 of each file — the `trunk:` stanza — reach SAM2-UNet, see
/root/reference/SAM2UNet.py:131-144) on top of the class defaults in
/root/reference/sam2/modeling/backbones/hieradet.py:175-199.  This module holds those values
directly and derives the per-block table exactly as hieradet.py:200-259 does.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

# hieradet.py:175-199 class defaults; each variant overrides a few fields (yaml trunk stanza).
_DEFAULTS = dict(
    embed_dim=96,
    num_heads=1,
    q_pool=3,
    q_stride=(2, 2),
    stages=(2, 3, 16, 3),
    dim_mul=2.0,
    head_mul=2.0,
    window_pos_embed_bkg_spatial_size=(14, 14),
    window_spec=(8, 4, 14, 7),
    global_att_blocks=(12, 16, 20),
)

_VARIANTS = {
    # sam2_configs/sam2_hiera_t.yaml:11-15
    "sam2_hiera_t.yaml": dict(embed_dim=96, num_heads=1, stages=(1, 2, 7, 2), global_att_blocks=(5, 7, 9),
                              window_pos_embed_bkg_spatial_size=(7, 7)),
    # sam2_configs/sam2_hiera_s.yaml:11-15 (the fork's hard-coded default, SAM2UNet.py:131)
    "sam2_hiera_s.yaml": dict(embed_dim=96, num_heads=1, stages=(1, 2, 11, 2), global_att_blocks=(7, 10, 13),
                              window_pos_embed_bkg_spatial_size=(7, 7)),
    # sam2_configs/sam2_hiera_b+.yaml:11-12
    "sam2_hiera_b+.yaml": dict(embed_dim=112, num_heads=2),
    # sam2_configs/sam2_hiera_l.yaml:11-16
    "sam2_hiera_l.yaml": dict(embed_dim=144, num_heads=2, stages=(2, 6, 36, 4), global_att_blocks=(23, 33, 43),
                              window_pos_embed_bkg_spatial_size=(7, 7), window_spec=(8, 4, 16, 8)),
    # not in the reference: a small trunk that exercises every block kind (plain window, transition with
    # q-pool, padded window, global) in seconds on a CPU.  Used by the parity tests only.
    "tiny_test.yaml": dict(embed_dim=32, num_heads=1, stages=(1, 2, 3, 2), global_att_blocks=(4,),
                           window_pos_embed_bkg_spatial_size=(7, 7), window_spec=(8, 4, 6, 3)),
}

_ALIASES = {"t": "sam2_hiera_t.yaml", "tiny": "sam2_hiera_t.yaml", "s": "sam2_hiera_s.yaml",
            "small": "sam2_hiera_s.yaml", "b+": "sam2_hiera_b+.yaml", "base_plus": "sam2_hiera_b+.yaml",
            "l": "sam2_hiera_l.yaml", "large": "sam2_hiera_l.yaml", "test": "tiny_test.yaml"}


@dataclass(frozen=True)
class BlockSpec:
    """One MultiScaleBlock (hieradet.py:84-167) as the engine sees it."""
    index: int
    stage: int            # 0-based stage of the block's OUTPUT
    dim: int
    dim_out: int
    num_heads: int
    window: int           # 0 = global attention
    q_pool: bool          # 2x2 max-pool of q (and of the projected shortcut)
    stage_end: bool       # emits a feature map after this block


@dataclass(frozen=True)
class TrunkConfig:
    name: str
    embed_dim: int
    num_heads: int
    stages: Tuple[int, ...]
    global_att_blocks: Tuple[int, ...]
    window_spec: Tuple[int, ...]
    window_pos_embed_bkg_spatial_size: Tuple[int, int]
    q_pool: int = 3
    q_stride: Tuple[int, int] = (2, 2)
    dim_mul: float = 2.0
    head_mul: float = 2.0
    blocks: Tuple[BlockSpec, ...] = field(default=(), compare=False)

    @property
    def stage_dims(self) -> List[int]:
        return [int(self.embed_dim * self.dim_mul ** i) for i in range(len(self.stages))]

    @property
    def head_dim(self) -> int:
        return self.embed_dim // self.num_heads


def _block_table(cfg: dict) -> Tuple[BlockSpec, ...]:
    # hieradet.py:203-259
    stages = tuple(cfg["stages"])
    depth = sum(stages)
    stage_ends = [sum(stages[:i]) - 1 for i in range(1, len(stages) + 1)]
    q_pool_blocks = [x + 1 for x in stage_ends[:-1]][: cfg["q_pool"]]
    embed_dim, num_heads = cfg["embed_dim"], cfg["num_heads"]
    cur_stage = 1
    out = []
    for i in range(depth):
        dim_out = embed_dim
        window = cfg["window_spec"][cur_stage - 1]          # lags one block (hieradet.py:237-240)
        if cfg["global_att_blocks"] is not None and i in cfg["global_att_blocks"]:
            window = 0
        if i - 1 in stage_ends:
            dim_out = int(embed_dim * cfg["dim_mul"])
            num_heads = int(num_heads * cfg["head_mul"])
            cur_stage += 1
        out.append(BlockSpec(index=i, stage=cur_stage - 1, dim=embed_dim, dim_out=dim_out, num_heads=num_heads,
                             window=window, q_pool=i in q_pool_blocks, stage_end=i in stage_ends))
        embed_dim = dim_out
    return tuple(out)


def canonical_name(name: str) -> str:
    name = _ALIASES.get(name.lower(), name)
    if name not in _VARIANTS:
        raise ValueError(f"unknown trunk config {name!r}; known: {sorted(_VARIANTS)}")
    return name


def trunk_config(name: str = "sam2_hiera_s.yaml", **overrides) -> TrunkConfig:
    """Resolve a yaml name (or alias t/s/b+/l) to the trunk hyper-parameters and block table."""
    name = canonical_name(name)
    cfg = dict(_DEFAULTS)
    cfg.update(_VARIANTS[name])
    cfg.update(overrides)
    if len(cfg["stages"]) != len(cfg["window_spec"]):
        raise ValueError("stages and window_spec must have the same length (hieradet.py:203)")
    blocks = _block_table(cfg)
    return TrunkConfig(name=name, embed_dim=cfg["embed_dim"], num_heads=cfg["num_heads"], stages=tuple(cfg["stages"]),
                       global_att_blocks=tuple(cfg["global_att_blocks"] or ()), window_spec=tuple(cfg["window_spec"]),
                       window_pos_embed_bkg_spatial_size=tuple(cfg["window_pos_embed_bkg_spatial_size"]),
                       q_pool=cfg["q_pool"], q_stride=tuple(cfg["q_stride"]), dim_mul=cfg["dim_mul"],
                       head_mul=cfg["head_mul"], blocks=blocks)


def load_trunk_yaml(path: str) -> TrunkConfig:
    """Read the `model.image_encoder.trunk` stanza of a reference-style yaml file with PyYAML."""
    import yaml

    with open(path) as f:
        doc = yaml.safe_load(f)
    stanza = dict(doc["model"]["image_encoder"]["trunk"])
    stanza.pop("_target_", None)
    cfg = dict(_DEFAULTS)
    for k, v in stanza.items():
        cfg[k] = tuple(v) if isinstance(v, list) else v
    blocks = _block_table(cfg)
    return TrunkConfig(name=path, embed_dim=cfg["embed_dim"], num_heads=cfg["num_heads"], stages=tuple(cfg["stages"]),
                       global_att_blocks=tuple(cfg["global_att_blocks"] or ()), window_spec=tuple(cfg["window_spec"]),
                       window_pos_embed_bkg_spatial_size=tuple(cfg["window_pos_embed_bkg_spatial_size"]),
                       q_pool=cfg["q_pool"], q_stride=tuple(cfg["q_stride"]), dim_mul=cfg["dim_mul"],
                       head_mul=cfg["head_mul"], blocks=blocks)

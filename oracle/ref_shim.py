"""ORACLE (test infrastructure): import the UNMODIFIED reference from /root/reference in this container.

The reference needs hydra, omegaconf and torchsummary, none of which is installed (and there is no
network).  This module pre-seeds `sys.modules` with minimal stand-ins — `hydra.compose` reads the yaml
with PyYAML, `hydra.utils.instantiate` is a recursive `_target_` importer — and then imports the
reference's own `SAM2UNet.py` / `sam2/...` files from where they lie.  Nothing of the reference is
copied.  It only works where /root/reference exists (the build container), so only
`oracle/make_golden.py` and the `not gpu` pin tests (skipped when the path is absent) use it.
"""
from __future__ import annotations

import ast
import importlib
import os
import sys
import types

import torch
import torch.nn as nn
import yaml

REF_ROOT = os.environ.get("SAM2UNET_REFERENCE", "/root/reference")

# trunk-stanza override applied by the stand-in `compose` (used for the small "test" trunk)
_TRUNK_OVERRIDE: dict = {}


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "SAM2UNet.py"))


def _coerce(node):
    """PyYAML reads `1e-6` as a string; Hydra/OmegaConf would give a float."""
    if isinstance(node, dict):
        return {k: _coerce(v) for k, v in node.items()}
    if isinstance(node, list):
        return [_coerce(v) for v in node]
    if isinstance(node, str):
        try:
            return float(node)
        except ValueError:
            return node
    return node


class _Cfg(dict):
    __getattr__ = dict.__getitem__


def _wrap(node):
    if isinstance(node, dict):
        return _Cfg({k: _wrap(v) for k, v in node.items()})
    if isinstance(node, list):
        return [_wrap(v) for v in node]
    return node


def _compose(config_name, overrides=()):
    with open(os.path.join(REF_ROOT, "sam2_configs", config_name)) as f:
        cfg = _coerce(yaml.safe_load(f))
    for ov in overrides:                         # "++a.b.c=value"
        path, value = ov.lstrip("+").split("=", 1)
        node = cfg
        keys = path.split(".")
        for k in keys[:-1]:
            node = node.setdefault(k, {})
        node[keys[-1]] = yaml.safe_load(value)
    if _TRUNK_OVERRIDE:
        cfg["model"]["image_encoder"]["trunk"].update(_TRUNK_OVERRIDE["trunk"])
        cfg["model"]["image_encoder"]["neck"]["backbone_channel_list"] = _TRUNK_OVERRIDE["channels"]
    return _wrap(cfg)


def _instantiate(node, _recursive_=True, **_):
    if isinstance(node, dict):
        kwargs = {k: _instantiate(v) for k, v in node.items() if k != "_target_"}
        if "_target_" in node:
            mod, attr = node["_target_"].rsplit(".", 1)
            return getattr(importlib.import_module(mod), attr)(**kwargs)
        return kwargs
    if isinstance(node, list):
        return [_instantiate(v) for v in node]
    return node


def _install_stubs():
    if "hydra" in sys.modules and getattr(sys.modules["hydra"], "_s2u_stub", False):
        return
    hydra = types.ModuleType("hydra")
    hydra._s2u_stub = True
    hydra.compose = _compose
    hydra.initialize_config_module = lambda *a, **k: None
    hutils = types.ModuleType("hydra.utils")
    hutils.instantiate = _instantiate
    hydra.utils = hutils
    omega = types.ModuleType("omegaconf")
    omega.OmegaConf = type("OmegaConf", (), {"resolve": staticmethod(lambda cfg: None)})
    tsum = types.ModuleType("torchsummary")
    tsum.summary = lambda *a, **k: None
    sys.modules.update({"hydra": hydra, "hydra.utils": hutils, "omegaconf": omega, "torchsummary": tsum})
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


_YAML = {"t": "sam2_hiera_t.yaml", "s": "sam2_hiera_s.yaml", "b+": "sam2_hiera_b+.yaml", "l": "sam2_hiera_l.yaml"}


def build_reference(variant: str = "s", trunk_override: dict | None = None) -> nn.Module:
    """Instantiate the reference `SAM2UNet` (SAM2UNet.py:128-162) on CPU for a trunk variant.

    The fork hard-codes the yaml name and the four RFB widths for Hiera-S (SAM2UNet.py:131,152-155);
    for the other variants the yaml name is substituted and rfb1-4 are rebuilt with the reference's own
    `RFB_modified` at widths E*{1,2,4,8}.  `forward` is the reference's, untouched.
    """
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    _install_stubs()
    ref = importlib.import_module("SAM2UNet")
    real_build = importlib.import_module("sam2.build_sam").build_sam2
    yaml_name = _YAML.get(variant, "sam2_hiera_s.yaml")
    _TRUNK_OVERRIDE.clear()
    if trunk_override is not None:
        dims = [trunk_override["embed_dim"] * 2 ** i for i in range(len(trunk_override["stages"]))]
        _TRUNK_OVERRIDE.update(trunk=dict(trunk_override), channels=dims[::-1])

    def patched(cfg_name, ckpt=None, **kw):
        return real_build(yaml_name, ckpt, device="cpu", **kw)

    ref.build_sam2 = patched
    try:
        model = ref.SAM2UNet()
    finally:
        ref.build_sam2 = real_build
        _TRUNK_OVERRIDE.clear()
    dims = [b.dim_out for b in (blk.block for blk in model.encoder.blocks)]
    ends = model.encoder.stage_ends
    widths = [dims[e] for e in ends]
    if widths != [96, 192, 384, 768]:
        for i, wdt in enumerate(widths):
            setattr(model, f"rfb{i + 1}", ref.RFB_modified(wdt, 64))
    return model


def reference_structure_loss():
    """Return the reference's `structure_loss` (train.py:21-29) compiled from its own source text.

    `train.py` cannot be imported here (it pulls dataset.py / eval.py -> skimage), so the function's AST
    node is extracted from the file and executed with torch / F in scope — the body is the reference's.
    """
    with open(os.path.join(REF_ROOT, "train.py")) as f:
        tree = ast.parse(f.read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "structure_loss")
    ns = {"torch": torch, "F": torch.nn.functional}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), os.path.join(REF_ROOT, "train.py"), "exec"), ns)
    return ns["structure_loss"]

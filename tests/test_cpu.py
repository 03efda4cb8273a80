"""CPU-side tests (`-m "not gpu"`): the oracle against the committed golden vectors (which the reference itself
produced, oracle/make_golden.py), the host logic (trunk tables, state-dict layout, flat parameter buffer,
gradient buckets, resampling tables), and that the C-ABI library loads and exports every declared symbol.
No kernel is launched here."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


# ------------------------------------------------------------------------------------ oracle vs golden vectors

def _sd(cfg):
    from sam2_unet_b200 import SAM2UNet
    from sam2_unet_b200.params import fill_deterministic_
    m = SAM2UNet(model_cfg=cfg, dtype="fp32")
    fill_deterministic_(m, 0)
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def _proj(n, k=4):
    g = torch.Generator().manual_seed(n % 1000003)
    return torch.randn(k, n, generator=g, dtype=torch.float64) / n ** 0.5


def test_pin_report_is_green():
    rep = json.load(open(os.path.join(GOLD, "pin_report.json")))
    for case in ("tiny_160_b2", "hiera_l_352"):
        r = rep[case]
        assert max(r["eval_logits_maxnorm"]) < 1e-4 and max(r["train_logits_maxnorm"]) < 1e-4
        assert r["loss_rel"] < 1e-5 and r["grad_global_rel_l2"] < 5e-3


def test_oracle_tiny_against_reference_vectors():
    from oracle import port
    sd = _sd("tiny_test.yaml")
    x, mask = port.synthetic_batch(2, 160, seed=2)
    gold_t, gold_e = np.load(os.path.join(GOLD, "tiny_train.npz")), np.load(os.path.join(GOLD, "tiny_eval.npz"))
    with torch.no_grad():
        ev = port.forward(sd, port.TRUNKS["test"], x, False)
    for o, nm in zip(ev, ("out", "out1", "out2")):
        assert np.abs(o.numpy() - gold_e[nm]).max() <= 1e-4 * np.abs(gold_e[nm]).max()
    loss, outs, grads = port.loss_and_grads(sd, port.TRUNKS["test"], x, mask, True, port.BNState())
    assert abs(loss.item() - float(gold_t["loss"])) <= 1e-5 * abs(float(gold_t["loss"]))
    for o, nm in zip(outs, ("out", "out1", "out2")):
        assert np.abs(o.numpy() - gold_t[nm]).max() <= 1e-4 * np.abs(gold_t[nm]).max()
    num = den = 0.0
    for k, g in grads.items():
        if g is None:
            assert "gnorm/" + k not in gold_t.files and k.startswith("up4.")
            continue
        v = g.double().reshape(-1)
        ref_proj = gold_t["gproj/" + k]
        num += float(((_proj(v.numel()) @ v).numpy() - ref_proj).__pow__(2).sum())
        den += float((ref_proj ** 2).sum())
        assert abs(v.norm().item() - float(gold_t["gnorm/" + k])) <= 2e-2 * float(gold_t["gnorm/" + k]) + 1e-9, k
    assert (num / den) ** 0.5 <= 1e-3


def test_oracle_config1_hiera_l_forward():
    """BASELINE.json config 1 (Hiera-L 352^2, batch 1, fp32 forward) — the oracle reproduces the reference's output."""
    from oracle import port
    sd = _sd("sam2_hiera_l.yaml")
    x, _ = port.synthetic_batch(1, 352, seed=0)
    gold = np.load(os.path.join(GOLD, "hiera_l_352_fwd.npz"))
    with torch.no_grad():
        outs = port.forward(sd, port.TRUNKS["l"], x, False)
    for o, nm in zip(outs, ("out", "out1", "out2")):
        assert np.abs(o[0, 0, ::3, ::3].numpy() - gold[nm]).max() <= 1e-4 * np.abs(gold[nm]).max()


def test_structure_loss_oracle_and_closed_form():
    from oracle import port
    gold = np.load(os.path.join(GOLD, "structure_loss.npz"))
    for i, (B, S) in enumerate(((2, 96), (3, 352), (1, 64))):
        _, mask = port.synthetic_batch(B, S, seed=3)
        pred = (torch.randn(B, 1, S, S, generator=torch.Generator().manual_seed(i)) * 2).requires_grad_(True)
        loss = port.structure_loss(pred, mask)
        (g,) = torch.autograd.grad(loss, pred)
        assert abs(loss.item() - float(gold[f"loss_{i}"])) <= 1e-6
        assert np.abs(g[:, 0, ::7, ::5].numpy() - gold[f"grad_{i}"]).max() <= 1e-4 * np.abs(gold[f"grad_{i}"]).max()
        cf_loss, cf_grad = port.structure_loss_closed_form(pred.detach().double(), mask.double())
        assert abs(cf_loss.item() - loss.item()) <= 1e-6
        assert (cf_grad.float() - g).abs().max() <= 1e-4 * g.abs().max()


def test_oracle_against_live_reference_when_present():
    from oracle import port, ref_shim
    if not ref_shim.available():
        pytest.skip("/root/reference is not present on this machine")
    from sam2_unet_b200.params import fill_deterministic_
    ref = ref_shim.build_reference("t")
    fill_deterministic_(ref, 1)
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    x, _ = port.synthetic_batch(1, 96, seed=9)
    ref.eval()
    with torch.no_grad():
        a = ref(x)
        b = port.forward(sd, port.TRUNKS["t"], x, False)
    for u, v in zip(a, b):
        assert (u - v).abs().max() <= 1e-4 * u.abs().max()


def test_adamw_and_cosine_schedule_match_torch():
    from oracle import port
    from sam2_unet_b200 import cosine_lr
    p = torch.randn(1000)
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([{"params": [ref], "initial_lr": 1e-3}], lr=1e-3, weight_decay=5e-4)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, 20, eta_min=1e-7)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for t in range(1, 6):
        g = torch.randn(1000, generator=torch.Generator().manual_seed(t))
        ref.grad = g.clone()
        lr = opt.param_groups[0]["lr"]
        assert abs(lr - cosine_lr(t - 1, 20)) < 1e-12 and abs(lr - port.cosine_lr(t - 1, 20)) < 1e-12
        opt.step()
        port.adamw_step(p, g, m, v, t, lr=lr)
        sched.step()
    assert (p - ref.detach()).abs().max() < 1e-6


# ----------------------------------------------------------------------------------------------- host logic

def test_trunk_tables_match_the_reference_variants():
    from oracle import port
    from sam2_unet_b200 import trunk_config
    want = {"t": (12, [1, 3, 10], [5, 7, 9]), "s": (16, [1, 3, 14], [7, 10, 13]), "b+": (24, [2, 5, 21], [12, 16, 20]),
            "l": (48, [2, 8, 44], [23, 33, 43])}
    for name, (depth, pools, globs) in want.items():
        cfg = trunk_config(name)
        assert len(cfg.blocks) == depth
        assert [b.index for b in cfg.blocks if b.q_pool] == pools
        assert [b.index for b in cfg.blocks if b.window == 0] == globs
        table = port.block_table(port.TRUNKS[name])
        for b, t in zip(cfg.blocks, table):
            assert (b.dim, b.dim_out, b.num_heads, b.window, b.q_pool, b.stage_end) == \
                   (t["dim"], t["dim_out"], t["heads"], t["window"], t["pool"], t["end"])
    lcfg = trunk_config("l")
    assert lcfg.stage_dims == [144, 288, 576, 1152] and lcfg.blocks[44].window == 16 and lcfg.blocks[45].window == 8
    assert all(b.dim_out // b.num_heads == 72 for b in lcfg.blocks)
    with pytest.raises(ValueError):
        trunk_config("nope.yaml")


def test_state_dict_layout_and_requires_grad_pattern():
    from sam2_unet_b200 import SAM2UNet
    m = SAM2UNet(model_cfg="sam2_hiera_l.yaml")
    sd = m.state_dict()
    assert len(sd) == 1192                                           # SURVEY.md section 8b, measured on the reference
    assert sum(p.numel() for p in m.parameters()) == 216529891
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 4380595
    for key in ("encoder.pos_embed", "encoder.pos_embed_window", "encoder.patch_embed.proj.weight",
                "encoder.blocks.2.block.proj.weight", "encoder.blocks.0.block.attn.qkv.bias",
                "encoder.blocks.47.block.mlp.layers.1.weight", "encoder.blocks.5.prompt_learn.2.bias",
                "rfb3.branch2.3.bn.running_var", "rfb1.conv_cat.conv.weight", "up4.conv.double_conv.4.num_batches_tracked",
                "side2.bias", "head.weight"):
        assert key in sd, key
    assert sd["rfb2.branch3.3.conv.weight"].shape == (64, 64, 3, 3) and sd["up1.conv.double_conv.0.weight"].shape == (64, 128, 3, 3)
    for n, p in m.named_parameters():
        frozen = n.startswith("encoder.") and ".prompt_learn." not in n
        assert p.requires_grad == (not frozen), n
    s = SAM2UNet()                                                   # the fork's default trunk is Hiera-S
    assert s.cfg.name == "sam2_hiera_s.yaml" and s.rfb1.conv_res.conv.weight.shape[1] == 96
    with pytest.raises(Exception):
        s(torch.randn(1, 3, 64, 64))                                  # CPU tensor: the product path has no fallback


def test_state_dict_matches_live_reference_keys():
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("/root/reference is not present on this machine")
    from sam2_unet_b200 import SAM2UNet
    ref = ref_shim.build_reference("s")
    mine = SAM2UNet()
    rs, ms = ref.state_dict(), mine.state_dict()
    assert list(rs.keys()) == list(ms.keys()) or set(rs) == set(ms)
    for k in rs:
        assert rs[k].shape == ms[k].shape, k
    mine.load_state_dict(rs, strict=True)
    ref.load_state_dict(mine.state_dict(), strict=True)
    assert {n for n, p in ref.named_parameters() if p.requires_grad} == {n for n, p in mine.named_parameters() if p.requires_grad}


def test_flat_parameters_alias_and_buckets():
    from sam2_unet_b200 import SAM2UNet
    from sam2_unet_b200.model import FlatParams
    m = SAM2UNet(model_cfg="tiny_test.yaml")
    before = {k: v.clone() for k, v in m.state_dict().items()}
    flat = FlatParams(m, torch.device("cpu"))
    for k, v in m.state_dict().items():
        assert torch.equal(v, before[k]), k
    p = dict(m.named_parameters())["side1.weight"]
    assert p.data_ptr() == flat.views["side1.weight"].data_ptr()
    with torch.no_grad():
        p.add_(1.0)
    off = flat.offsets["side1.weight"]
    assert torch.equal(flat.master[off:off + 64].view_as(p), p)
    assert all(flat.offsets[n] >= flat.n_active for n in flat.offsets if n.startswith("up4."))
    assert all(flat.offsets[n] < flat.n_active for n in flat.offsets if not n.startswith("up4."))
    lo_prev = 0
    for lo, hi, ready in flat.buckets:                               # contiguous cover of the active range
        assert lo == lo_prev and hi > lo
        lo_prev = hi
    assert lo_prev == flat.n_active and flat.buckets[0][2] == len(m.cfg.blocks) and flat.buckets[-1][2] == 0
    for n, o in flat.offsets.items():                                # an adapter's bucket is ready only after its block
        if n.startswith("encoder.blocks."):
            blk = int(n.split(".")[2])
            ready = next(r for lo, hi, r in flat.buckets if lo <= o < hi)
            assert ready <= blk


def test_resample_tables_match_aten():
    from sam2_unet_b200.resample_tables import axis_backward, axis_forward
    for n_in, n_out, ac, sc in ((11, 22, True, None), (44, 88, True, None), (22, 352, False, 16.0), (88, 352, False, 4.0)):
        f = axis_forward(n_in, n_out, ac, sc)
        x = torch.randn(1, 1, n_in, n_in, requires_grad=True)
        ref = F.interpolate(x, scale_factor=2 if ac else sc, mode="bilinear", align_corners=ac if ac else None)
        i0, i1, w0, w1 = (torch.from_numpy(a) for a in f)
        xv = x.detach()[0, 0]
        rows = w0[:, None] * xv[i0.long()] + w1[:, None] * xv[i1.long()]
        out = w0[None, :] * rows[:, i0.long()] + w1[None, :] * rows[:, i1.long()]
        assert (out - ref[0, 0]).abs().max() < 1e-5
        bi, bw, taps = axis_backward(n_in, n_out, f)
        g = torch.randn_like(ref)
        (gx,) = torch.autograd.grad(ref, x, g)
        bi_t, bw_t = torch.from_numpy(bi).long(), torch.from_numpy(bw)
        tmp = (bw_t[:, :, None] * g[0, 0][bi_t]).sum(1)
        dx = (bw_t[None, :, :] * tmp[:, bi_t]).sum(2)
        assert (dx - gx[0, 0]).abs().max() < 1e-4


# ------------------------------------------------------------------------------------------------ C ABI

def _header_decls():
    src = open(os.path.join(ROOT, "include", "sam2unet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.findall(r"int\s+(s2u_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S)


def test_library_exports_every_declared_symbol():
    from sam2_unet_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from sam2_unet_b200 import build
        build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    decls = _header_decls()
    assert len(decls) >= 29
    for name, _ in decls:
        assert hasattr(lib, name), name
    assert {n for n, _ in decls} == set(_lib.SIGNATURES)


def test_ctypes_signatures_match_the_header():
    from sam2_unet_b200 import _lib
    kinds = {"P": ctypes.c_void_p, "I": ctypes.c_int, "L": ctypes.c_longlong, "F": ctypes.c_float}

    def kind(arg):
        return "P" if "*" in arg else "L" if "long long" in arg else "F" if "float" in arg else "I"

    for name, args in _header_decls():
        want = [kinds[kind(a.strip())] for a in args.split(",")]
        assert want == _lib.SIGNATURES[name], name


def test_missing_library_fails_loudly(monkeypatch):
    from sam2_unet_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libsam2unet_b200.so")
    with pytest.raises(_lib.KernelError):
        _lib.call("s2u_add", 0, 0, 0, 8, 0, 0)


# --------------------------------------------------------------------------- eval metrics oracle (eval.py:55-171)

def test_eval_oracle_known_answers():
    """Known-answer pins of the CPU restatement of eval.py's per-image metrics (skimage is absent from the image, so
    these hand-computed cases are what anchors the labelling: 8-connectivity, raster label order, greedy matching)."""
    import numpy as np
    from oracle import eval_port
    z = np.zeros((40, 60), np.uint8)
    a = z.copy(); a[5:15, 5:15] = 255; a[20:30, 30:45] = 255           # two components
    r = eval_port.evaluate_segmentation_performance(a, a)
    assert r["semantic_iou"] == 1.0 and r["dice_coefficient"] == 1.0 and r["count_gt"] == 2 and r["count_pred"] == 2
    assert r["instance_precision_50"] == r["instance_recall_75"] == r["instance_f1_75"] == 1.0
    b = z.copy(); b[5:15, 7:17] = 255; b[20:30, 30:45] = 255           # first square shifted by 2: IoU 80 / 120
    r = eval_port.evaluate_segmentation_performance(b, a)
    assert r["semantic_iou"] == (80 + 150) / (120 + 150) and r["dice_coefficient"] == 2 * 230 / 500
    assert r["instance_precision_50"] == 1.0 and r["instance_precision_75"] == 0.5 and r["instance_recall_75"] == 0.5
    d = z.copy(); d[3, 3] = d[4, 4] = d[5, 5] = 255; d[10, 10] = 30    # diagonal chain = ONE component; 30 > 25.5 counts
    r = eval_port.evaluate_segmentation_performance(d, z)
    assert r["count_pred"] == 2 and r["count_gt"] == 0 and r["semantic_iou"] == 0.0 and r["instance_f1_50"] == 0.0
    r = eval_port.evaluate_segmentation_performance(z, a)
    assert r["count_pred"] == 0 and r["instance_precision_50"] == 0.0 and r["dice_coefficient"] == 0.0


def _blob_masks(seed, H=72, W=96):
    """A prediction / ground-truth pair with several touching, diagonal and competing components (uint8 0..255)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    gt = np.zeros((H, W), np.uint8)
    pred = np.zeros((H, W), np.float32)
    for _ in range(int(rng.integers(1, 7))):
        cy, cx, r = rng.uniform(0, H), rng.uniform(0, W), rng.uniform(3, 14)
        gt[(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = 255
        dy, dx, s = rng.normal(0, 3), rng.normal(0, 3), rng.uniform(0.6, 1.3)
        pred[(yy - cy - dy) ** 2 + (xx - cx - dx) ** 2 <= (r * s) ** 2] = rng.uniform(20, 255)
    speck = rng.random((H, W)) < 0.004                                  # isolated / diagonal specks
    pred[speck] = 200
    return pred.astype(np.uint8), gt


def _label_flood_fill(binary):
    """An independent statement of what skimage.measure.label(binary) documents for a 2-D input with the default
    connectivity (= ndim: the 8-neighbourhood), background 0, labels 1..n in raster order of each component's first
    pixel: plain stack-based flood fill, no scipy / cv2."""
    H, W = binary.shape
    out = np.zeros((H, W), np.int64)
    n = 0
    for y in range(H):
        for x in range(W):
            if binary[y, x] and not out[y, x]:
                n += 1
                out[y, x] = n
                stack = [(y, x)]
                while stack:
                    cy, cx = stack.pop()
                    for ny in (cy - 1, cy, cy + 1):
                        for nx in (cx - 1, cx, cx + 1):
                            if 0 <= ny < H and 0 <= nx < W and binary[ny, nx] and not out[ny, nx]:
                                out[ny, nx] = n
                                stack.append((ny, nx))
    return out


def test_eval_oracle_labelling_three_ways():
    """The one call of eval.py that cannot be run here (skimage.measure.label, eval.py:105-106; scikit-image is not in
    this image): the oracle's scipy labelling, the flood-fill statement of skimage's documented semantics and OpenCV's
    8-connected labelling give the same components, and the first two the same label ORDER (which the greedy matching
    of eval.py:120-150 depends on)."""
    from scipy import ndimage
    cv2 = pytest.importorskip("cv2")
    for seed in range(12):
        pred, gt = _blob_masks(seed)
        for m in (pred > 25.5, gt > 25.5):
            b = m.astype(np.uint8)
            a, na = ndimage.label(b, structure=np.ones((3, 3), np.uint8))
            f = _label_flood_fill(b)
            assert na == f.max() and np.array_equal(a, f)
            nc, c = cv2.connectedComponents(b, connectivity=8)
            assert nc - 1 == na
            pairs = np.unique(np.stack([a[b > 0], c[b > 0]], 1), axis=0)          # a bijection between the label sets
            assert len(pairs) == na and len(set(pairs[:, 0])) == na and len(set(pairs[:, 1])) == na


@pytest.mark.skipif(not os.path.isfile("/root/reference/eval.py"), reason="reference not present")
def test_eval_oracle_against_live_reference_eval_py():
    """The UNMODIFIED /root/reference/eval.py executed here, with `skimage.measure` supplied by a stand-in module whose
    `label` is the flood fill above and whose `regionprops` lists the labels in ascending order (skimage's documented
    order): every line of evaluate_segmentation_performance / evaluate_dataset except the labelling call itself is the
    reference's own code.  oracle/eval_port.py must reproduce its dictionaries exactly."""
    import importlib.util
    import sys
    import types
    pytest.importorskip("cv2")
    from oracle import eval_port
    sk, skm = types.ModuleType("skimage"), types.ModuleType("skimage.measure")
    skm.label = _label_flood_fill
    skm.regionprops = lambda lab: [types.SimpleNamespace(label=int(v)) for v in range(1, int(lab.max()) + 1)]
    sk.measure = skm
    saved = {k: sys.modules.get(k) for k in ("skimage", "skimage.measure")}
    sys.modules.update({"skimage": sk, "skimage.measure": skm})
    try:
        spec = importlib.util.spec_from_file_location("ref_eval", "/root/reference/eval.py")
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    theirs, mine = [], []
    for seed in range(24):
        pred, gt = _blob_masks(100 + seed)
        if seed == 5:
            pred[:] = 0
        if seed == 6:
            gt[:] = 0
        a, b = ref.evaluate_segmentation_performance(pred, gt), eval_port.evaluate_segmentation_performance(pred, gt)
        assert set(a) == set(b)
        for k in a:
            assert float(a[k]) == float(b[k]), (seed, k, a[k], b[k])
        theirs.append(a)
        mine.append(b)
    A, B = ref.evaluate_dataset(theirs), eval_port.evaluate_dataset(mine)
    assert list(A) == list(B)
    for k in A:
        assert float(A[k]) == float(B[k]), k
    with pytest.raises(ValueError):
        eval_port.evaluate_segmentation_performance(np.zeros((4, 4), np.uint8), np.zeros((4, 5), np.uint8))


def _fake_upstream_checkpoint(model, path, seed=7):
    """A file shaped like Meta's sam2_hiera_*.pt: {"model": {"image_encoder.trunk.<hiera key>": tensor, ...other
    SAM2Base keys...}} (/root/reference/sam2/build_sam.py:79-89; SAM2-UNet keeps the trunk only, SAM2UNet.py:136-144)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for n, p in model.encoder.named_parameters():
        if ".prompt_learn." in n:
            continue
        sd["image_encoder.trunk." + n.replace(".block.", ".", 1)] = torch.randn(p.shape, generator=g)
    sd["image_encoder.neck.convs.0.conv.weight"] = torch.randn(4, 4, 1, 1, generator=g)     # dropped by SAM2-UNet
    sd["sam_mask_decoder.iou_token.weight"] = torch.randn(1, 8, generator=g)
    sd["maskmem_tpos_enc"] = torch.randn(7, 1, 1, 8, generator=g)
    torch.save({"model": sd}, path)
    return sd


def test_load_hiera_checkpoint_synthesised(tmp_path):
    """`SAM2UNet(checkpoint_path)` / `load_hiera_checkpoint` on an upstream-shaped file: every trunk tensor arrives under
    its `encoder.blocks.N.block.*` name, adapters and decoder keep their init, non-trunk keys are ignored, the
    requires_grad pattern is untouched, and a file that lacks a trunk tensor is rejected (strict, like build_sam2)."""
    from sam2_unet_b200 import SAM2UNet
    torch.manual_seed(0)
    base = SAM2UNet(model_cfg="tiny_test.yaml")
    path = str(tmp_path / "sam2_hiera_fake.pt")
    sd = _fake_upstream_checkpoint(base, path)
    torch.manual_seed(0)
    m = SAM2UNet(path, model_cfg="tiny_test.yaml")
    mine = m.state_dict()
    n_trunk = 0
    for k, v in sd.items():
        if not k.startswith("image_encoder.trunk."):
            continue
        k2 = k[len("image_encoder.trunk."):]
        if k2.startswith("blocks."):
            parts = k2.split(".")
            k2 = ".".join(parts[:2] + ["block"] + parts[2:])
        assert torch.equal(mine["encoder." + k2], v), k
        n_trunk += 1
    assert n_trunk == sum(1 for n, _ in base.encoder.named_parameters() if ".prompt_learn." not in n)
    ref = base.state_dict()
    for k, v in mine.items():                             # everything that is not the trunk: same seeded init as `base`
        if ".prompt_learn." in k or not k.startswith("encoder."):
            assert torch.equal(v, ref[k]), k
    assert all(p.requires_grad == (".prompt_learn." in n) for n, p in m.encoder.named_parameters())
    assert all(p.requires_grad for n, p in m.named_parameters() if not n.startswith("encoder."))
    broken = {k: v for k, v in sd.items() if not k.endswith("blocks.1.attn.qkv.weight")}
    torch.save({"model": broken}, path)
    with pytest.raises(RuntimeError):
        SAM2UNet(path, model_cfg="tiny_test.yaml")


@pytest.mark.skipif(not os.path.isfile("/root/reference/SAM2UNet.py"), reason="reference not present")
def test_load_hiera_checkpoint_against_live_reference(tmp_path):
    """The same upstream-shaped file through the reference's own loader (build_sam2 -> _load_checkpoint, strict into the
    full SAM2Base, then SAM2UNet.py:136-144 keeps the trunk) and through ours: identical trunk tensors."""
    from oracle import ref_shim
    from sam2_unet_b200 import SAM2UNet
    import importlib
    ref_shim._install_stubs()
    torch.manual_seed(0)
    ref_model = ref_shim.build_reference("s")             # Hiera-S: the fork's hard-coded trunk
    full_sd = None
    # a complete SAM2Base state dict is needed for the reference's strict load: take it from a fresh build_sam2
    build = importlib.import_module("sam2.build_sam")
    sam = build.build_sam2("sam2_hiera_s.yaml", None, device="cpu")
    g = torch.Generator().manual_seed(3)
    full_sd = {k: torch.randn(v.shape, generator=g).to(v.dtype) if v.dtype.is_floating_point else v.clone()
               for k, v in sam.state_dict().items()}
    path = str(tmp_path / "sam2_hiera_small_fake.pt")
    torch.save({"model": full_sd}, path)
    sam2 = build.build_sam2("sam2_hiera_s.yaml", path, device="cpu")       # the reference's loader, strict
    theirs = {"encoder." + k.replace("blocks.", "blocks.", 1): v for k, v in sam2.image_encoder.trunk.state_dict().items()}
    m = SAM2UNet(path, model_cfg="sam2_hiera_s.yaml")
    mine = m.state_dict()
    n = 0
    for k, v in theirs.items():
        k2 = k
        if k2.startswith("encoder.blocks."):
            parts = k2.split(".")
            k2 = ".".join(parts[:3] + ["block"] + parts[3:])
        assert torch.equal(mine[k2], v), k
        n += 1
    assert n == sum(1 for nme, _ in m.encoder.named_parameters() if ".prompt_learn." not in nme)
    del ref_model


# ------------------------------------------------------------------------------------------ training-time input pipeline

def _augment_inputs(seed):
    from oracle.make_golden_augment import inputs
    return inputs(seed)


def test_augment_oracle_against_reference_vectors():
    """oracle/augment_port.py (the CPU restatement of dataset.py:288-313) reproduces tests/golden/augment_train.npz, which
    the unmodified reference transform wrote (oracle/make_golden_augment.py): same `random.seed`, same uint8 inputs."""
    import random
    from oracle import augment_port as ap
    g = np.load(os.path.join(GOLD, "augment_train.npz"))
    S = int(g["size"])
    assert len(g["seeds"]) >= 6
    for seed in g["seeds"].tolist():
        img, lab = _augment_inputs(seed)
        random.seed(seed)
        out = ap.apply(ap.draw(S, *lab.shape), torch.from_numpy(img), torch.from_numpy(lab), S)
        assert torch.equal(out["label"], torch.from_numpy(g[f"label_{seed}"])), seed
        err = (out["image"] - torch.from_numpy(g[f"image_{seed}"])).abs().max().item()
        assert err <= 1e-5, (seed, err)


def test_augment_draw_matches_the_oracle_draw():
    """The product's host-side decision draw consumes `random` exactly like the oracle's (and so like the reference)."""
    import random
    from oracle import augment_port as ap
    from sam2_unet_b200.augment import draw_train_params
    for seed in range(200):
        random.seed(seed)
        a = ap.draw(352, 300 + seed, 517 - seed)
        sa = random.random()
        random.seed(seed)
        b = draw_train_params(352, 300 + seed, 517 - seed)
        sb = random.random()
        assert sa == sb
        assert a["geom"][1:] == b["geom"][1:] and (a["geom"][0] == "crop") == (b["geom"][0] == 1)
        assert (a["rot"], a["gray"], a["color"], a["blur"]) == (b["rot"], b["gray"], b["color"], b["blur"])


@pytest.mark.skipif(not os.path.isfile("/root/reference/dataset.py"), reason="reference not present")
def test_augment_oracle_against_live_reference():
    """60 more seeds straight through the reference's transform classes (needs torchvision, cv2 and PIL here)."""
    import importlib.util
    import random
    pytest.importorskip("cv2")
    from PIL import Image
    from torchvision import transforms
    from oracle import augment_port as ap
    spec = importlib.util.spec_from_file_location("ref_dataset", "/root/reference/dataset.py")
    rd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rd)
    S = 64
    tf = transforms.Compose([rd.ToTensor(), rd.ResizeLongestSideAndPad(S), rd.RandomRotate(), rd.ToGray(),
                             rd.ColorAugmentations(), rd.GaussianBlur(), rd.Normalize()])
    for seed in range(500, 560):
        img, lab = _augment_inputs(seed)
        random.seed(seed)
        ref = tf({"image": Image.fromarray(img), "label": Image.fromarray(lab)})
        random.seed(seed)
        out = ap.apply(ap.draw(S, *lab.shape), torch.from_numpy(img), torch.from_numpy(lab), S)
        assert torch.equal(out["label"], ref["label"]), seed
        assert (out["image"] - ref["image"]).abs().max().item() <= 1e-5, seed


@pytest.mark.skipif(not os.path.isfile("/root/reference/eval.py"), reason="reference not present")
def test_print_eval_report_matches_the_reference(tmp_path, capsys):
    """The report text (eval.py:23-52) of the package equals the reference function's, character for character."""
    import importlib.util
    import sys
    import types
    pytest.importorskip("cv2")
    from sam2_unet_b200.evalmetrics import print_eval_report
    sk, skm = types.ModuleType("skimage"), types.ModuleType("skimage.measure")
    skm.label = skm.regionprops = None
    sk.measure = skm
    saved = {k: sys.modules.get(k) for k in ("skimage", "skimage.measure")}
    sys.modules.update({"skimage": sk, "skimage.measure": skm})
    try:
        spec = importlib.util.spec_from_file_location("ref_eval2", "/root/reference/eval.py")
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    res = {"semantic_iou": 0.81234567, "dice_coefficient": 0.9, "count_gt": 3, "count_pred": 12, "instance_f1_50": 0.0}
    for title in ("Segmentation Evaluation", "[3/17] a_rather_long_file_name_0001.png", "x"):
        a, b = str(tmp_path / "a.txt"), str(tmp_path / "b.txt")
        ref.print_eval_report(res, title=title, log_path=a)
        theirs = capsys.readouterr().out
        print_eval_report(res, title=title, log_path=b)
        mine = capsys.readouterr().out
        assert mine == theirs
        assert open(a).read() == open(b).read()

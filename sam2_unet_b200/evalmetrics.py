"""The reference's per-image segmentation metrics (eval.py:55-171) with the pixel work on the device.

    result = evaluate_segmentation_performance(pred_mask, gt_mask)      # uint8 [H,W] CUDA tensors, e.g. infer_tail()

returns the same dictionary as the reference function: semantic IoU and Dice, component counts, and instance
precision / recall / F1 at IoU 0.5 and 0.75.  Thresholding, the four pixel counts, 8-connected component labelling,
component areas and pairwise intersections are kernels (csrc/eval_metrics.cu, exact integer results); the greedy
matching over the few dozen components is the reference's own loop on the host.  `evaluate_dataset` is the reference's
aggregation (eval.py:170-224), pure host arithmetic.  No CPU fallback for the pixel work: CPU tensors raise.
"""
from typing import Dict, List

import numpy as np
import torch

from . import _lib

IOU_THRESHOLDS = [0.5, 0.75]          # eval.py:9
SCORE_THRESHOLD = 0.1                 # eval.py:11
SEMANTIC_IOU = "semantic_iou"
DICE_COEFFICIENT = "dice_coefficient"
COUNT_GT = "count_gt"
COUNT_PRED = "count_pred"
INSTANCE_PRECISION = "instance_precision"
INSTANCE_RECALL = "instance_recall"
INSTANCE_F1 = "instance_f1"
MIOU = "mIoU"
MDICE = "mDice"

_ROOT_CAP = 1 << 16
_PAIR_CAP = 1 << 17


def _label(mask: torch.Tensor, threshold: float, st: int):
    H, W = mask.shape
    dev = mask.device
    labels = torch.empty(H * W, dtype=torch.int32, device=dev)
    roots = torch.empty(_ROOT_CAP, dtype=torch.int32, device=dev)
    nroots = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.call("s2u_cc_label", mask.data_ptr(), float(threshold), H, W, labels.data_ptr(), roots.data_ptr(),
              nroots.data_ptr(), _ROOT_CAP, st)
    return labels, roots, nroots


def evaluate_segmentation_performance(pred_mask: torch.Tensor, gt_mask: torch.Tensor,
                                      threshold: float = 255 * SCORE_THRESHOLD) -> Dict[str, float]:
    if pred_mask.shape != gt_mask.shape:
        raise ValueError(f"Shape mismatch: Pred {tuple(pred_mask.shape)} vs GT {tuple(gt_mask.shape)}")
    if pred_mask.device.type != "cuda" or gt_mask.device.type != "cuda":
        raise _lib.KernelError("evaluate_segmentation_performance needs CUDA tensors (no CPU fallback)")
    if pred_mask.dtype != torch.uint8 or gt_mask.dtype != torch.uint8 or pred_mask.dim() != 2:
        raise ValueError("expected two uint8 [H,W] masks")
    pred_mask, gt_mask = pred_mask.contiguous(), gt_mask.contiguous()
    dev = pred_mask.device
    n = pred_mask.numel()
    st = torch.cuda.current_stream(dev).cuda_stream
    counts = torch.zeros(4, dtype=torch.int64, device=dev)
    _lib.call("s2u_seg_counts", pred_mask.data_ptr(), gt_mask.data_ptr(), n, float(threshold), counts.data_ptr(), st)
    pl, proots, pn = _label(pred_mask, threshold, st)
    gl, groots, gn = _label(gt_mask, threshold, st)
    parea = torch.zeros(n, dtype=torch.int32, device=dev)
    garea = torch.zeros(n, dtype=torch.int32, device=dev)
    keys = torch.full((_PAIR_CAP,), -1, dtype=torch.int64, device=dev)
    vals = torch.zeros(_PAIR_CAP, dtype=torch.int32, device=dev)
    overflow = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.call("s2u_cc_stats", pl.data_ptr(), gl.data_ptr(), n, parea.data_ptr(), garea.data_ptr(), keys.data_ptr(),
              vals.data_ptr(), _PAIR_CAP, overflow.data_ptr(), st)
    # ---- everything below is host logic on a few hundred integers
    inter, union, psum, gsum = (int(v) for v in counts.cpu())
    npred, ngt = int(pn.cpu()), int(gn.cpu())
    if npred > _ROOT_CAP or ngt > _ROOT_CAP or int(overflow.cpu()):
        raise _lib.KernelError("more connected components / overlapping pairs than the metric workspace holds")
    pr = sorted(proots[:npred].cpu().tolist())            # root = first pixel in raster order: skimage's label order
    gr = sorted(groots[:ngt].cpu().tolist())
    pa = dict(zip(pr, parea[torch.tensor(pr, dtype=torch.long, device=dev)].cpu().tolist())) if pr else {}
    ga = dict(zip(gr, garea[torch.tensor(gr, dtype=torch.long, device=dev)].cpu().tolist())) if gr else {}
    k = keys.cpu()
    used = (k != -1).nonzero().flatten()
    v = vals.cpu()
    gidx = {r: i for i, r in enumerate(gr)}
    overlaps: Dict[int, List] = {}
    for s in used.tolist():
        key = int(k[s]) & 0xFFFFFFFFFFFFFFFF
        overlaps.setdefault(key >> 32, []).append((gidx[key & 0xFFFFFFFF], int(v[s])))
    s_iou = inter / union if union > 0 else 0.0           # eval.py:88-101
    dice = (2 * inter) / (psum + gsum) if (psum + gsum) > 0 else 0.0
    result = {SEMANTIC_IOU: s_iou, DICE_COEFFICIENT: dice, COUNT_GT: ngt, COUNT_PRED: npred}
    for thresh in IOU_THRESHOLDS:                         # eval.py:120-165
        tp = 0
        matched = set()
        for p in pr:
            best_iou, best_idx = 0, -1
            for idx, i_pg in sorted(overlaps.get(p, [])):   # components without overlap have IoU 0: never the best
                if idx in matched:
                    continue
                u = pa[p] + ga[gr[idx]] - i_pg
                iou = i_pg / u if u > 0 else 0
                if iou > best_iou:
                    best_iou, best_idx = iou, idx
            if best_iou >= thresh:
                tp += 1
                matched.add(best_idx)
        precision = tp / npred if npred > 0 else 0.0
        recall = tp / ngt if ngt > 0 else 0.0
        f1 = 2 * (precision * recall) / (precision + recall) if (precision + recall) > 0 else 0.0
        suffix = int(thresh * 100)
        result[f"{INSTANCE_PRECISION}_{suffix}"] = precision
        result[f"{INSTANCE_RECALL}_{suffix}"] = recall
        result[f"{INSTANCE_F1}_{suffix}"] = f1
    return result


def evaluate_dataset(all_image_results: List[Dict[str, float]]) -> Dict[str, float]:
    """eval.py:170-224: mean IoU / Dice over images, instance precision / recall / F1 from the summed counts."""
    if not all_image_results:
        return {}
    mean_iou = float(np.mean([r[SEMANTIC_IOU] for r in all_image_results]))       # np.mean like the reference: the
    mean_dice = float(np.mean([r[DICE_COEFFICIENT] for r in all_image_results]))  # same (pairwise) summation order
    total_gt = sum(r[COUNT_GT] for r in all_image_results)
    total_pred = sum(r[COUNT_PRED] for r in all_image_results)
    final = {MIOU: mean_iou, MDICE: mean_dice, "images_count": total_gt}
    for thresh in IOU_THRESHOLDS:
        suffix = int(thresh * 100)
        total_tp = sum(r[f"{INSTANCE_PRECISION}_{suffix}"] * r[COUNT_PRED] for r in all_image_results)
        precision = total_tp / total_pred if total_pred > 0 else 0.0
        recall = total_tp / total_gt if total_gt > 0 else 0.0
        f1 = (2 * precision * recall) / (precision + recall) if (precision + recall) > 0 else 0.0
        final[f"Precision_{suffix}"] = precision
        final[f"Recall_{suffix}"] = recall
        final[f"F1_Score_{suffix}"] = f1
    return final


def print_eval_report(results: Dict, title: str = "Evaluation Results", log_path: str = None) -> str:
    """eval.py:23-52: the reference's report layout (centred title between rules, one `name : value` line per entry with
    underscores shown as spaces, floats to four decimals), printed and - when `log_path` is given - appended to that
    file.  Returns the text."""
    width = max(len(title) + 2, 25)
    lines = ["\n" + "=" * width, f"{title:^{width}}", "-" * width]
    for metric, value in results.items():
        name = str(metric).replace("_", " ")
        if isinstance(value, float):
            lines.append(f"{name:<{width - 8}}: {value:>6.4f}")
        else:
            lines.append(f"{name:<{width - 8}}: {value:>6}")
    lines.append("=" * width + "\n")
    text = "\n".join(lines)
    print(text)
    if log_path:
        with open(log_path, "a") as f:
            f.write(text)
    return text

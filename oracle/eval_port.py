"""CPU restatement of the reference's evaluation metrics (TEST INFRASTRUCTURE: only tests/ may import this).

Follows /root/reference/eval.py:55-171 (`evaluate_segmentation_performance`) and :170-224 (`evaluate_dataset`) line by
line.  The reference labels connected components with skimage.measure.label / regionprops, which is NOT installed in
this image (scikit-image, un-pinned in the reference's requirements): the labelling is restated with
scipy.ndimage.label and a full 3x3 structuring element, i.e. skimage's documented default for 2-D input (connectivity =
ndim = 8-neighbourhood, labels numbered in raster order of each component's first pixel).  Pinning: tests/test_cpu.py
runs the UNMODIFIED reference eval.py (with a stand-in `skimage.measure` module, because skimage cannot be imported
here) against this file and requires identical dictionaries, and checks the labelling three ways (scipy, a flood fill
stating skimage's documented semantics, OpenCV).  PARITY UNPINNED against skimage's own binary for that one call.
"""
import numpy as np
from scipy import ndimage

IOU_THRESHOLDS = [0.5, 0.75]
SCORE_THRESHOLD = 0.1


def evaluate_segmentation_performance(pred_mask: np.ndarray, gt_mask: np.ndarray,
                                      threshold: float = 255 * SCORE_THRESHOLD) -> dict:
    if pred_mask.shape != gt_mask.shape:
        raise ValueError(f"Shape mismatch: Pred {pred_mask.shape} vs GT {gt_mask.shape}")
    pred_bin = (pred_mask > threshold).astype(np.uint8)                    # eval.py:81-82
    gt_bin = (gt_mask > threshold).astype(np.uint8)
    intersection = np.logical_and(pred_bin, gt_bin).sum()                  # eval.py:85-86
    union = np.logical_or(pred_bin, gt_bin).sum()
    s_iou = intersection / union if union > 0 else 0.0
    dice = (2 * intersection) / (pred_bin.sum() + gt_bin.sum()) if (pred_bin.sum() + gt_bin.sum()) > 0 else 0.0
    full = np.ones((3, 3), dtype=np.uint8)
    pred_label, n_pred = ndimage.label(pred_bin, structure=full)           # eval.py:105-106 (skimage label)
    gt_label, n_gt = ndimage.label(gt_bin, structure=full)
    result = {"semantic_iou": float(s_iou), "dice_coefficient": float(dice), "count_gt": n_gt, "count_pred": n_pred}
    for thresh in IOU_THRESHOLDS:                                          # eval.py:120-165
        tp = 0
        matched = set()
        for p in range(1, n_pred + 1):
            best_iou, best_idx = 0, -1
            p_mask = pred_label == p
            for idx in range(n_gt):
                if idx in matched:
                    continue
                g_mask = gt_label == idx + 1
                inter = np.logical_and(p_mask, g_mask).sum()
                uni = np.logical_or(p_mask, g_mask).sum()
                iou = inter / uni if uni > 0 else 0
                if iou > best_iou:
                    best_iou, best_idx = iou, idx
            if best_iou >= thresh:
                tp += 1
                matched.add(best_idx)
        precision = tp / n_pred if n_pred > 0 else 0.0
        recall = tp / n_gt if n_gt > 0 else 0.0
        f1 = 2 * (precision * recall) / (precision + recall) if (precision + recall) > 0 else 0.0
        suffix = int(thresh * 100)
        result[f"instance_precision_{suffix}"] = precision
        result[f"instance_recall_{suffix}"] = recall
        result[f"instance_f1_{suffix}"] = f1
    return result


def evaluate_dataset(all_image_results: list) -> dict:
    """eval.py:170-224: mean IoU / Dice over images; instance precision / recall / F1 from the true-positive counts
    re-derived as precision * count_pred and summed over the data set (note: the reference stores the GROUND-TRUTH
    component total under the key "images_count")."""
    if not all_image_results:
        return {}
    total_gt = sum(r["count_gt"] for r in all_image_results)
    total_pred = sum(r["count_pred"] for r in all_image_results)
    out = {"mIoU": np.mean([r["semantic_iou"] for r in all_image_results]),
           "mDice": np.mean([r["dice_coefficient"] for r in all_image_results]), "images_count": total_gt}
    for thresh in IOU_THRESHOLDS:
        suffix = int(thresh * 100)
        tp = sum(r[f"instance_precision_{suffix}"] * r["count_pred"] for r in all_image_results)
        precision = tp / total_pred if total_pred > 0 else 0.0
        recall = tp / total_gt if total_gt > 0 else 0.0
        f1 = (2 * precision * recall) / (precision + recall) if (precision + recall) > 0 else 0.0
        out[f"Precision_{suffix}"] = precision
        out[f"Recall_{suffix}"] = recall
        out[f"F1_Score_{suffix}"] = f1
    return out

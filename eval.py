#!/usr/bin/env python
"""eval.py - drop-in for the reference's evaluation entry point (/root/reference/eval.py:226-266): every ground-truth
mask of --gt_path is scored against the PNG of the same stem in --pred_path (semantic IoU / Dice, connected-component
instance precision / recall / F1 at IoU 0.5 and 0.75), a report per image and one for the data set are printed and
appended to <pred_path>/log.txt in the reference's layout.  The pixel work (thresholding, counting, 8-connected
labelling, overlap table) runs on the device: `sam2_unet_b200.evaluate_segmentation_performance`; images are decoded
on the host with PIL in grayscale mode (the reference uses cv2.IMREAD_GRAYSCALE; 8-bit single-channel PNGs, which is
what test.py writes, decode identically)."""
from __future__ import annotations

import argparse
import os

import numpy as np
import torch

from sam2_unet_b200 import evaluate_dataset, evaluate_segmentation_performance, print_eval_report

# names the reference module exports (eval.py:8-20)
IOU_THRESHOLDS = [0.5, 0.75]
SCORE_THRESHOLD = 0.1


def _gray(path: str) -> torch.Tensor:
    from PIL import Image
    return torch.from_numpy(np.array(Image.open(path).convert("L")))


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--pred_path", type=str, required=True, help="Path to the prediction results")
    parser.add_argument("--gt_path", type=str, required=True, help="Path to the ground truth masks")
    args = parser.parse_args()
    device = torch.device("cuda")
    gt_list = sorted(os.listdir(args.gt_path))
    log_path = os.path.join(args.pred_path, "log.txt")
    results = []
    for i, mask_name in enumerate(gt_list):
        gt = _gray(os.path.join(args.gt_path, mask_name)).to(device)
        pred = _gray(os.path.join(args.pred_path, mask_name[:-4] + ".png")).to(device)
        result = evaluate_segmentation_performance(pred, gt)
        print_eval_report(result, title=f"[{i + 1}/{len(gt_list)}] {mask_name}", log_path=log_path)
        results.append(result)
    print_eval_report(evaluate_dataset(results), title="Segmentation Evaluation", log_path=log_path)

#!/usr/bin/env python
"""test.py — drop-in for the reference's inference entry point (/root/reference/test.py:13-86): loads a SAM2-UNet
checkpoint, runs every test image through the model (batch 1, no_grad), removes the padding, resizes to the ground
truth size, applies sigmoid + min-max normalisation and writes 8-bit PNGs; prints the mean forward time (measured
with CUDA events here; the reference's time.time() without a sync only measures the launch)."""
from __future__ import annotations

import argparse
import os

import numpy as np
import torch

from sam2_unet_b200 import SAM2UNet, infer_tail, preprocess_image

if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--checkpoint", type=str, required=True, help="path to the checkpoint of sam2-unet")
    parser.add_argument("--save_path", type=str, required=True, help="path to save the predicted masks")
    parser.add_argument("--test_image_path", type=str, default="data_test/images/")
    parser.add_argument("--test_gt_path", type=str, default="data_test/masks/")
    parser.add_argument("--size", default=960, type=int)
    parser.add_argument("--model_cfg", default="sam2_hiera_s.yaml")
    parser.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    args = parser.parse_args()

    from PIL import Image
    device = torch.device("cuda")
    model = SAM2UNet(model_cfg=args.model_cfg, dtype=args.dtype).to(device)
    model.load_state_dict(torch.load(args.checkpoint, map_location=device), strict=True)
    model.eval()
    os.makedirs(args.save_path, exist_ok=True)
    names = sorted(f for f in os.listdir(args.test_image_path) if f.endswith((".jpg", ".png")))
    times = []
    for name in names:
        img_u8 = np.ascontiguousarray(np.asarray(Image.open(os.path.join(args.test_image_path, name)).convert("RGB")))
        gt_path = os.path.join(args.test_gt_path, name[:-4] + ".png")
        gt_shape = np.asarray(Image.open(gt_path)).shape[:2] if os.path.exists(gt_path) else img_u8.shape[:2]
        # TestDataset's transforms (dataset.py:336-407) on the device: /255, antialiased resize of the longest side,
        # centred zero padding, normalisation; `padding` = (left, top, right, bottom) like load_data's
        x, padding = preprocess_image(torch.from_numpy(img_u8).to(device), args.size)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.no_grad():
            e0.record()
            res, _, _ = model(x)
            e1.record()
        # remove padding, resize to the ground-truth size, sigmoid, min-max, uint8 (test.py:66-76): on the device,
        # one byte per pixel comes back
        res = infer_tail(res, padding, tuple(gt_shape))
        Image.fromarray(res.cpu().numpy()).save(os.path.join(args.save_path, name[:-4] + ".png"))
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) / 1e3)
        print("Saving " + name)
        print("process_time:", times[-1])
    print("mean_test_time:", float(np.mean(times)) if times else float("nan"))

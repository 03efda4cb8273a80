"""Throughput of the training-time input pipeline: `sam2_unet_b200.TrainAugment` on the
device (uint8 image + mask cross PCIe, all transforms as CUDA kernels) next to the reference's own
FullDataset(mode="train").transform on ONE host core (a DataLoader worker, train.py:35 runs 8 of them), same seeds.

    python scripts/augment_bench.py [--size 352] [--n 64] > gpurun_out/augment_bench.jsonl
"""
import argparse
import importlib.util
import json
import os
import random
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=352)
    ap.add_argument("--n", type=int, default=64)
    args = ap.parse_args()
    from sam2_unet_b200 import TrainAugment
    dev = torch.device("cuda", 0)
    aug = TrainAugment(args.size, dev)
    ref_tf = None
    for root in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        path = os.path.join(root, "dataset.py")
        if os.path.isfile(path):
            try:
                from torchvision import transforms
                spec = importlib.util.spec_from_file_location("ref_dataset", path)
                rd = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(rd)
                ref_tf = transforms.Compose([rd.ToTensor(), rd.ResizeLongestSideAndPad(args.size), rd.RandomRotate(),
                                             rd.ToGray(), rd.ColorAugmentations(), rd.GaussianBlur(), rd.Normalize()])
            except Exception as e:                                   # noqa: BLE001
                print(json.dumps({"reference_transform_unavailable": repr(e)}))
            break
    for H, W in ((480, 640), (1080, 1920)):
        rng = np.random.default_rng(H)
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        lab = (rng.random((H, W)) > 0.5).astype(np.uint8) * 255
        img_h, lab_h = torch.from_numpy(img).pin_memory(), torch.from_numpy(lab).pin_memory()
        for _ in range(5):
            aug(img_h, lab_h)
        torch.cuda.synchronize(dev)
        random.seed(0)
        t0 = time.perf_counter()
        for _ in range(args.n):
            aug(img_h, lab_h)
        torch.cuda.synchronize(dev)
        gpu_s = (time.perf_counter() - t0) / args.n
        line = {"source": [H, W], "size": args.size, "device_img_per_s": 1.0 / gpu_s, "device_ms_per_image": gpu_s * 1e3,
                "h2d_bytes_per_image": H * W * 4}
        if ref_tf is not None:
            from PIL import Image
            torch.set_num_threads(1)
            pim, plab = Image.fromarray(img), Image.fromarray(lab)
            random.seed(0)
            n = max(4, args.n // 8)
            t0 = time.perf_counter()
            for _ in range(n):
                ref_tf({"image": pim, "label": plab})
            cpu_s = (time.perf_counter() - t0) / n
            line.update({"reference_cpu_img_per_s_one_worker": 1.0 / cpu_s, "reference_cpu_ms_per_image": cpu_s * 1e3,
                         "reference_cpu_img_per_s_8_workers_if_linear": 8.0 / cpu_s})
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()

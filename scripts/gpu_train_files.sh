#!/bin/bash
# train.py end to end on FILES (not --synthetic): images of several sizes on disk -> host decode -> device augmentation
# (sam2_unet_b200.TrainAugment) -> TrainStep; evaluation through preprocess_image / infer_tail / the device metrics.
set -u
D=gpurun_out/tf
mkdir -p $D/train/images $D/train/masks $D/test/images $D/test/masks
python - <<'PY'
import numpy as np
from PIL import Image
rng = np.random.default_rng(0)
def sample(h, w):
    yy, xx = np.mgrid[0:h, 0:w]
    m = np.zeros((h, w), bool)
    for _ in range(3):
        cy, cx, r = rng.uniform(0, h), rng.uniform(0, w), rng.uniform(0.08, 0.25) * min(h, w)
        m |= (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r
    img = rng.normal(90, 30, (h, w, 3)) + m[..., None] * 90
    return np.clip(img, 0, 255).astype(np.uint8), m.astype(np.uint8) * 255
sizes = [(300, 420), (512, 384), (352, 352), (240, 640), (600, 450), (333, 517)]
for i in range(48):
    img, m = sample(*sizes[i % len(sizes)])
    Image.fromarray(img).save(f"gpurun_out/tf/train/images/im{i:03d}.jpg")
    Image.fromarray(m).save(f"gpurun_out/tf/train/masks/im{i:03d}.png")
for i in range(6):
    img, m = sample(*sizes[i])
    Image.fromarray(img).save(f"gpurun_out/tf/test/images/t{i}.jpg")
    Image.fromarray(m).save(f"gpurun_out/tf/test/masks/t{i}.png")
PY
timeout 600 python train.py --save_path $D/ck --random_trunk --size 352 --model_cfg sam2_hiera_t.yaml --epoch 3 --batch_size 12 \
  --train_image_path $D/train/images/ --train_mask_path $D/train/masks/ --test_image_path $D/test/images/ \
  --test_gt_path $D/test/masks/ 2>&1 | grep -v "^epoch-" | tail -12 | cut -c1-220
echo "train.py rc=$?"
rm -rf $D

#!/bin/bash
# test.py and eval.py end to end on synthetic files: a checkpoint from a short train.py run, a few random images
set -u
mkdir -p gpurun_out/tp/img gpurun_out/tp/gt gpurun_out/tp/out
python - <<'PY'
import numpy as np
from PIL import Image
rng = np.random.default_rng(0)
for i, (h, w) in enumerate([(300, 420), (512, 384), (352, 352)]):
    Image.fromarray(rng.integers(0, 255, (h, w, 3), dtype=np.uint8)).save(f"gpurun_out/tp/img/im{i}.jpg")
    Image.fromarray((rng.random((h, w)) > 0.5).astype(np.uint8) * 255).save(f"gpurun_out/tp/gt/im{i}.png")
PY
timeout 300 python train.py --save_path gpurun_out/tp/ck --synthetic 24 --size 352 --model_cfg sam2_hiera_t.yaml --epoch 1 --batch_size 12 2>&1 | tail -2
timeout 300 python test.py --checkpoint "$(ls gpurun_out/tp/ck/*.pth | head -1)" --test_image_path gpurun_out/tp/img/ --test_gt_path gpurun_out/tp/gt/ --save_path gpurun_out/tp/out --size 352 --model_cfg sam2_hiera_t.yaml 2>&1 | tail -4
python - <<'PY'
import numpy as np
from PIL import Image
for i in range(3):
    a = np.asarray(Image.open(f"gpurun_out/tp/out/im{i}.png"))
    print(a.shape, a.dtype, a.min(), a.max())
PY
timeout 200 python eval.py --pred_path gpurun_out/tp/out --gt_path gpurun_out/tp/gt/ 2>&1 | tail -16
rm -rf gpurun_out/tp

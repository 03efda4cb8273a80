#!/bin/bash
# end-of-round sanity: other configs through bench.py, the train.py command line
set -u
mkdir -p gpurun_out
for cfg in sam2_hiera_t.yaml sam2_hiera_s.yaml sam2_hiera_b+.yaml; do
  timeout 200 python bench.py --cfg $cfg --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-260
done
timeout 300 python bench.py --size 1024 --batch 4 --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-260
timeout 300 python bench.py --size 1024 --batch 1 --mode infer --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
timeout 300 python train.py --save_path gpurun_out/ck --synthetic 24 --size 352 --model_cfg sam2_hiera_t.yaml --epoch 2 --batch_size 12 2>&1 | tail -4 | cut -c1-200
rm -rf gpurun_out/ck

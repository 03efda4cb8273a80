#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -k "attention" -q --no-header -p no:cacheprovider > gpurun_out/t_attn.log 2>&1; echo "attn rc=$?"; grep -E "^(FAILED|E   .*Error)|passed|failed" gpurun_out/t_attn.log | head -40
timeout 900 python -m pytest tests/test_model_gpu.py -q --no-header -p no:cacheprovider > gpurun_out/t_model.log 2>&1; echo "model rc=$?"; grep -E "^(FAILED|E   .*Error)|passed|failed" gpurun_out/t_model.log | head -40
timeout 900 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/profile_train2.json > gpurun_out/bench2.log 2>&1; echo "bench rc=$?"; tail -3 gpurun_out/bench2.log

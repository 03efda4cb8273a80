#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -k "attention" -q --no-header -p no:cacheprovider > gpurun_out/t_attn.log 2>&1; echo "attn rc=$?"; grep -E "^(FAILED|E   .*Error)|passed|failed" gpurun_out/t_attn.log | head -30
timeout 900 python -m pytest tests/test_model_gpu.py -q --no-header -p no:cacheprovider > gpurun_out/t_model.log 2>&1; echo "model rc=$?"; grep -E "^(FAILED|E   .*Error)|passed|failed" gpurun_out/t_model.log | head -20
timeout 600 python scripts/profile_step.py > gpurun_out/profile_step4.log 2>&1; echo "prof rc=$?"; head -22 gpurun_out/profile_step4.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench5.log 2>&1; echo "bench rc=$?"; tail -2 gpurun_out/bench5.log | cut -c1-300

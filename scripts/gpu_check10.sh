#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -k "layernorm" -q --no-header -p no:cacheprovider 2>&1 | tail -2
timeout 900 python -m pytest tests/test_model_gpu.py -k "prefit" -q -s --no-header -p no:cacheprovider > gpurun_out/t_prefit.log 2>&1; echo "prefit rc=$?"; grep -E "^\{|^(FAILED|E   )|passed|failed" gpurun_out/t_prefit.log | head -20 | cut -c1-900
timeout 600 python scripts/profile_step.py > gpurun_out/profile_step6.log 2>&1; echo "prof rc=$?"; head -12 gpurun_out/profile_step6.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench8.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench8.log | cut -c1-250

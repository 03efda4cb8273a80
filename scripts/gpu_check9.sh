#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -k "not umma and not simt" -q --no-header -p no:cacheprovider > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?"; grep -E "^(FAILED|E   .*Error)|passed|failed" gpurun_out/t_kernels.log | head -20
timeout 900 python -m pytest tests/test_model_gpu.py -q --no-header -p no:cacheprovider > gpurun_out/t_model.log 2>&1; echo "model rc=$?"; grep -E "^(FAILED|E   .*Error)|passed|failed" gpurun_out/t_model.log | head -20
timeout 600 python scripts/profile_step.py > gpurun_out/profile_step5.log 2>&1; echo "prof rc=$?"; head -24 gpurun_out/profile_step5.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench7.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench7.log | cut -c1-250

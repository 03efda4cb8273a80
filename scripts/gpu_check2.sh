#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -k "not umma" -q --no-header -p no:cacheprovider > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?"; tail -5 gpurun_out/t_kernels.log
timeout 600 python scripts/debug_grads.py > gpurun_out/debug_grads.log 2>&1; echo "debug rc=$?"; cat gpurun_out/debug_grads.log | tail -80
timeout 900 python -m pytest tests/test_model_gpu.py -q --no-header -p no:cacheprovider > gpurun_out/t_model.log 2>&1; echo "model rc=$?"; tail -30 gpurun_out/t_model.log
timeout 900 python bench.py --steps 5 --warmup 3 --profile-out gpurun_out/profile_train.json > gpurun_out/bench1.log 2>&1; echo "bench rc=$?"; tail -5 gpurun_out/bench1.log

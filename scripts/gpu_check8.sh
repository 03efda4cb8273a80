#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -k "umma" -q -x --no-header -p no:cacheprovider 2>&1 | tail -2
timeout 900 python -m pytest tests/test_model_gpu.py -q --no-header -p no:cacheprovider > gpurun_out/t_model.log 2>&1; echo "model rc=$?"; grep -E "^(FAILED|E   .*Error)|passed|failed" gpurun_out/t_model.log | head -20
timeout 300 python scripts/gemm_bench.py 2>&1 | head -2 | cut -c1-1200
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench6.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench6.log | cut -c1-250
timeout 600 python bench.py --mode infer --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_infer.log 2>&1; echo "infer rc=$?"; tail -1 gpurun_out/bench_infer.log | cut -c1-900
timeout 600 python bench.py --mode infer --batch 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_infer1.log 2>&1; echo "infer1 rc=$?"; tail -1 gpurun_out/bench_infer1.log | cut -c1-250
timeout 600 python bench.py --size 1024 --batch 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1024.log 2>&1; echo "1024 rc=$?"; tail -1 gpurun_out/bench_1024.log | cut -c1-250

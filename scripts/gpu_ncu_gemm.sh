#!/bin/bash
# ncu --set full capture of the dominant GEMM launches inside one graph-replayed bench step
set -u
mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --ncu-range > gpurun_out/ncu_plain.log 2>&1 || { echo "plain failed"; tail -5 gpurun_out/ncu_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_umma_pair_kernel -s 100 -c 10 -f -o gpurun_out/r1_gemm_pair python bench.py --steps 2 --warmup 3 --no-cpu-baseline --ncu-range > gpurun_out/ncu_gemm.log 2>&1
echo "gemm rc=$?"; ls -la gpurun_out/*.ncu-rep | tail -3

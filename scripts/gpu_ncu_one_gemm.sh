#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 100 python scripts/one_gemm.py 129 > gpurun_out/one_gemm.log 2>&1 || { echo "plain failed"; tail -3 gpurun_out/one_gemm.log; exit 1; }
timeout 400 ncu --set full --section SourceCounters --clock-control none --import-source on -k regex:gemm_umma_pair_kernel -s 3 -c 1 -f -o gpurun_out/r1_one_gemm python scripts/one_gemm.py 129 > gpurun_out/ncu_one_gemm.log 2>&1
echo "rc=$?"; ls -la gpurun_out/r1_one_gemm.ncu-rep

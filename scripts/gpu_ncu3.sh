#!/bin/bash
# ncu evidence for profiles/: launch list of one graph-replayed bench step + full capture of the dominant GEMM
set -u
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --ncu-range > gpurun_out/ncu_plain.log 2>&1 || { echo "plain failed"; tail -5 gpurun_out/ncu_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --ncu-range > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"; wc -l gpurun_out/r1_launches.csv
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_umma_ws_kernel -s 60 -c 8 -f -o gpurun_out/r1_gemm_ws python bench.py --steps 2 --warmup 3 --no-cpu-baseline --ncu-range > gpurun_out/ncu_gemm.log 2>&1
echo "gemm rc=$?"; ls -la gpurun_out/*.ncu-rep | tail -3

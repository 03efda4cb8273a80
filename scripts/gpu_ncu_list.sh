#!/bin/bash
# ncu launch list of one graph-replayed bench step (per-kernel gpu__time_duration, serialised)
set -u
mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --ncu-range > gpurun_out/ncu_plain.log 2>&1 || { echo "plain failed"; tail -5 gpurun_out/ncu_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --ncu-range > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"; wc -l gpurun_out/r1_launches.csv

#!/bin/bash
set -u
mkdir -p gpurun_out
python scripts/ncu_step.py > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1.csv python scripts/ncu_step.py > gpurun_out/ncu_run.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_run.log; wc -l gpurun_out/launches_r1.csv

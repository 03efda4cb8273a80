#!/bin/bash
# ncu --set full (with source counters) of the windowed-attention forward kernel from the kernel zoo
set -u
mkdir -p gpurun_out
timeout 200 python scripts/kernel_zoo.py "attn_fwd win14" > gpurun_out/zoo_attn.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/zoo_attn.log; exit 1; }
timeout 600 ncu --set full --section SourceCounters --clock-control none --import-source on -k "regex:fwd_kernel" -s 3 -c 1 -f -o gpurun_out/r1_attn_fwd python scripts/kernel_zoo.py "attn_fwd win14" > gpurun_out/ncu_attn.log 2>&1
echo "rc=$?"; ls -la gpurun_out/r1_attn_fwd.ncu-rep

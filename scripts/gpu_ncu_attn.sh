#!/bin/bash
# ncu --set full of the windowed-attention kernels (stage-3 shape) from the kernel zoo
set -u
mkdir -p gpurun_out
timeout 200 python scripts/kernel_zoo.py "attn_bwd win14" > gpurun_out/zoo_attn.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/zoo_attn.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:bwd_dq_kernel|bwd_dkv_kernel|^fwd_kernel|amma::fwd_kernel" -s 10 -c 4 -f -o gpurun_out/r1_attn python scripts/kernel_zoo.py "attn_bwd win14" > gpurun_out/ncu_attn.log 2>&1
echo "rc=$?"; ls -la gpurun_out/r1_attn.ncu-rep

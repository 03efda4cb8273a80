#!/bin/bash
set -u
mkdir -p gpurun_out
python scripts/ncu_step.py > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.log; exit 1; }
for spec in "fwd_kernel:14:attn_fwd" "bwd_dq_kernel:8:attn_dq" "bwd_dkv_kernel:8:attn_dkv" "gemm_umma_ws_kernel<256:40:gemm256"; do
  IFS=: read -r pat skip name <<< "$spec"
  ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$pat" -s "$skip" -c 1 -f -o "gpurun_out/prof_$name" python scripts/ncu_step.py > "gpurun_out/ncu_$name.log" 2>&1
  echo "$name rc=$?"
done
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# First-contact GPU run: tcgen05 GEMM in its own process (a trap poisons the context), then every kernel test,
# then the end-to-end parity tests.  Logs go to gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== umma ==" ; timeout 600 python -m pytest tests/test_kernels_gpu.py -k "umma" -q -x --no-header -p no:cacheprovider > gpurun_out/t_umma.log 2>&1; echo "umma rc=$?"; tail -15 gpurun_out/t_umma.log
echo "== kernels ==" ; timeout 900 python -m pytest tests/test_kernels_gpu.py -k "not umma" -q --no-header -p no:cacheprovider > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?"; tail -40 gpurun_out/t_kernels.log
echo "== model ==" ; timeout 900 python -m pytest tests/test_model_gpu.py -q --no-header -p no:cacheprovider > gpurun_out/t_model.log 2>&1; echo "model rc=$?"; tail -40 gpurun_out/t_model.log

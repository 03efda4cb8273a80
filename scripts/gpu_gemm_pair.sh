#!/bin/bash
# guarded first run of the CTA-pair GEMM: one small parity case under a short timeout, then the GEMM tests, then timing
set -x
timeout 120 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "test_gemm_umma and pair256 and 484" 2>&1 | tail -15
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "test_gemm" 2>&1 | tail -15
timeout 300 python scripts/gemm_bench.py 2>&1 | tail -20

#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -k "single_block" -q --no-header -p no:cacheprovider > gpurun_out/t_blocks.log 2>&1; echo "blocks rc=$?"; grep -E "^(FAILED|E  )|passed|failed" gpurun_out/t_blocks.log | head -60

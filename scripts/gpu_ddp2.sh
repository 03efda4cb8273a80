#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.log 2>&1; echo "bench2 rc=$?"; tail -4 gpurun_out/bench_2gpu.log | cut -c1-700

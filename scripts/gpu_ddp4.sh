#!/bin/bash
set -u
mkdir -p gpurun_out
N=${1:-4}
nvidia-smi -L | wc -l
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu.log 2>&1; echo "bench$N rc=$?"; tail -2 gpurun_out/bench_${N}gpu.log | cut -c1-330

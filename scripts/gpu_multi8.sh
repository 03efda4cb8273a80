#!/bin/bash
# 8-GPU box: BASELINE config 3 (global batch 96 over 2 / 4 / 8 GPUs = strong scaling), config 5 (Hiera-L 1024x1024,
# batch 4 per GPU, 8 GPUs) and the 2-GPU NCCL correctness test.  One JSON line per run in gpurun_out/r2_multi8.jsonl.
set -u
mkdir -p gpurun_out
out=gpurun_out/r2_multi8.jsonl
: > $out
run() {  # N, extra args...
  local n=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
    bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline --no-infer "$@" 2>gpurun_out/multi8_err_$n.log | tail -1 >> $out
}
if [ "${1:-8}" = "1" ]; then   # the single-GPU points of the same tables (run on a 1-GPU box)
  timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-infer --global-batch 96 2>/dev/null | tail -1 >> $out
  timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-infer --size 1024 --batch 4 2>/dev/null | tail -1 >> $out
  cp $out gpurun_out/r2_multi1.jsonl
  cat $out | cut -c1-200
  exit 0
fi
for n in 2 4 8; do run $n --global-batch 96; done
run 8 --size 1024 --batch 4
timeout 600 python -m pytest tests/test_ddp_gpu.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/r2_ddp_test.txt
cat gpurun_out/r2_ddp_test.txt
python - <<'PY'
import json
print("| workload | GPUs | scaling | global batch | img/s | ms / step | img/s end to end |")
print("|---|---:|---|---:|---:|---:|---:|")
for l in open("gpurun_out/r2_multi8.jsonl"):
    if not l.startswith("{"):
        continue
    r = json.loads(l)
    c = r["config"]
    print(f"| {c['workload']} | {r['n_gpus']} | {r['scaling']} | {c['global_batch']} | {r['value']:.0f} | {r['ms_per_step']:.2f} | {r['e2e']['value']:.0f} |")
PY

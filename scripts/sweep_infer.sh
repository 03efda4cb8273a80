#!/bin/bash
# BASELINE config 4: inference batch sweep 1-256 at 352x352 across the Hiera-T/S/B+/L trunks (one B200, bf16,
# CUDA-graphed Predictor).  Writes one JSON line per point to gpurun_out/r2_sweep_infer.jsonl and a table.
set -u
mkdir -p gpurun_out
out=gpurun_out/r2_sweep_infer.jsonl
: > $out
for cfg in sam2_hiera_t.yaml sam2_hiera_s.yaml sam2_hiera_b+.yaml sam2_hiera_l.yaml; do
  for b in 1 4 16 64 256; do
    timeout 300 python bench.py --mode infer --cfg $cfg --batch $b --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 >> $out
  done
done
python - <<'PY'
import json
rows = [json.loads(l) for l in open("gpurun_out/r2_sweep_infer.jsonl") if l.startswith("{")]
print("| trunk | batch | img/s (device) | ms / batch | img/s end to end (H2D of the images + D2H of the three maps) |")
print("|---|---:|---:|---:|---:|")
for r in rows:
    c = r["config"]
    print(f"| {c['workload'].split()[1]} | {c['global_batch']} | {r['value']:.0f} | {r['ms_per_step']:.2f} | {r['e2e']['value']:.0f} |")
PY

#!/bin/bash
# Weak-scaling evidence (images are independent: batch shards, no collective on the path):
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 900 -- 'bash profiles/scale.sh'
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
: > gpurun_out/r2_scale.jsonl
for n in 1 2 4 8; do
  [ "$n" -le "$NG" ] || continue
  if [ "$n" -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline >> gpurun_out/r2_scale.jsonl 2>> gpurun_out/r2_scale.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline >> gpurun_out/r2_scale.jsonl 2>> gpurun_out/r2_scale.err
  fi
done
# BASELINE config 4 at the GPU count of the box: the RD training step with the bucketed NCCL gradient all-reduce
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29600 \
  bench.py --gpus $NG --workload train --steps 5 --warmup 3 > gpurun_out/r2_train_${NG}gpu.json 2>> gpurun_out/r2_scale.err
python - <<'PY'
import json
rows = [json.loads(l) for l in open("gpurun_out/r2_scale.jsonl") if l.startswith("{")]
base = rows[0]["value"] if rows else 0
for r in rows:
    print(r["n_gpus"], round(r["value"], 1), round(r["e2e"]["value"], 1), round(r["e2e_compact"]["value"], 1),
          round(r["value"] / (base * r["n_gpus"]), 3), r["clocks"])
PY

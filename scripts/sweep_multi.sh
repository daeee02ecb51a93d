#!/bin/bash
# BASELINE configs[4] on N GPUs (default 8): k in {21,25,31} x coverage in {10,30,100}, C2 genome per GPU (weak scaling), insert
# + filter only.  One JSON line per case into gpurun_out/sweep_n<N>.jsonl.
#   gpurun --gpus 8 --timeout 1500 -- 'bash scripts/sweep_multi.sh 8'
N=${1:-8}
mkdir -p gpurun_out
OUT=gpurun_out/sweep_n${N}.jsonl
: > $OUT
for COV in 10 30 100; do
  for K in 21 25 31; do
    if [ "$N" -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --k $K --coverage $COV --no-graph --no-cpu-baseline >> $OUT 2>> gpurun_out/sweep_n${N}.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
        bench.py --gpus $N --steps 10 --warmup 3 --k $K --coverage $COV --no-graph --no-cpu-baseline >> $OUT 2>> gpurun_out/sweep_n${N}.err
    fi
  done
done
N=$N python - <<'PY'
import json, os
for line in open("gpurun_out/sweep_n%s.jsonl" % os.environ["N"]):
    try:
        d = json.loads(line)
    except Exception:
        continue
    print(d["config"]["k"], d["config"]["workload"][-30:], "%.1f G k-mers/s" % (d["value"] / 1e9), "%.2f ms" % d["ms_per_step"])
PY

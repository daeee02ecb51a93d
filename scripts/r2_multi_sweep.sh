#!/bin/bash
# Where the sharded insert and the sharded graph build lose time at N GPUs: traces + tuning variants (tuning run: no parity leg).
#   gpurun --gpus N --timeout 1500 -- 'bash scripts/r2_multi_sweep.sh'
mkdir -p gpurun_out
NGPU=$(python -c "import torch; print(torch.cuda.device_count())")
LOG=gpurun_out/r2_msweep_${NGPU}.log
run() { # name, tune, extra args...
  name=$1; tune=$2; shift 2
  GENOME_B200_TUNE="$tune" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NGPU --master-addr 127.0.0.1 --master-port 29612 \
    bench.py --gpus $NGPU --steps 10 --warmup 3 --no-cpu-baseline --no-named "$@" > gpurun_out/r2_ms${NGPU}_$name.json 2> gpurun_out/r2_ms${NGPU}_$name.err
  echo "-- $name ($tune) rc=$?"
  grep -E "^\[(pmap|sgraph|pgraph)\]" gpurun_out/r2_ms${NGPU}_$name.err | tail -${TAILN:-0}
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2_ms${NGPU}_$name.json').read().strip().splitlines()[-1])
    g = d.get('graph') or {}
    print('   %.3f ms/step device, e2e %.3f ms, insert %.3f ms, graph build %s ms (kernels %s), simplify %s ms' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['insert_ms'], g.get('build_ms'), g.get('build_kernels_ms'), g.get('components_retain_simplify_ms')))
except Exception as e:
    print('   failed', e)
PY
}
{
  TAILN=60 run trace "trace=1"
  TAILN=40 run sgraph_trace "trace=1,pgraph_sharded=1"
  run default ""
  run sgraph "pgraph_sharded=1"
  run batches4 "batches=4" --no-graph
  run batches1 "batches=1" --no-graph
  run route2 "route=2" --no-graph
  run slices16 "slice_bits=4" --no-graph
  run slices4 "slice_bits=2" --no-graph
} > $LOG 2>&1
cat $LOG

#!/bin/bash
# Where the sharded insert and the sharded graph build lose time at N GPUs: traces + tuning variants (tuning run: no parity leg).
#   gpurun --gpus N --timeout 1500 -- 'bash scripts/r2_multi_sweep.sh [tests]'
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
  if [ "$1" = "tests" ]; then
    echo "== sharded parity tests at $NGPU GPUs"
    timeout 1200 python -m pytest tests/test_parity_multigpu.py -q -m gpu -x ${TESTSEL:+-k "$TESTSEL"} 2>&1 | tail -8
  fi
  TAILN=6 run dma_trace "trace=1,a2a=2" --no-graph
  TAILN=12 run sgraph_trace "trace=1,pgraph_sharded=1"
  run default ""
  run dma "a2a=2" --no-graph
  run dma_b3 "a2a=2,batches=3" --no-graph
  run dma_b4 "a2a=2,batches=4" --no-graph
  run dma_b4_s16 "a2a=2,batches=4,slice_bits=4" --no-graph
} > $LOG 2>&1
cat $LOG

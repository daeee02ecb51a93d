#!/bin/bash
# Round 2 (session c), N-GPU call: sharded parity tests at this box's size, then bench.py under the routing / graph-build variants.
#   gpurun --gpus N --timeout 1500 -- 'bash scripts/r2l_multi.sh [notests]'
mkdir -p gpurun_out
NGPU=$(python -c "import torch; print(torch.cuda.device_count())")
LOG=gpurun_out/r2l_multi_${NGPU}.log
run() { # name, tune, extra args...
  name=$1; tune=$2; shift 2
  GENOME_B200_TUNE="$tune" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NGPU --master-addr 127.0.0.1 --master-port 29612 \
    bench.py --gpus $NGPU --steps 10 --warmup 3 "$@" > gpurun_out/r2l_n${NGPU}_$name.json 2> gpurun_out/r2l_n${NGPU}_$name.err
  echo "-- $name ($tune $*) rc=$?"
  grep -E "^\[(pmap|sgraph|pgraph)\]" gpurun_out/r2l_n${NGPU}_$name.err | tail -${TAILN:-0}
  tail -2 gpurun_out/r2l_n${NGPU}_$name.err | cut -c1-300
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2l_n${NGPU}_$name.json').read().strip().splitlines()[-1])
    g = d.get('graph') or {}
    print('   %.3f ms/step device, e2e %.3f ms, insert %.3f ms, parity %s, graph build %s ms (kernels %s), simplify %s ms' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['insert_ms'], d.get('parity_checked'), g.get('build_ms'), g.get('build_kernels_ms'), g.get('components_retain_simplify_ms')))
except Exception as e:
    print('   failed', e)
PY
}
{
  if [ "$1" != "notests" ]; then
    echo "== sharded parity tests at $NGPU GPUs"
    timeout 1200 python -m pytest tests/test_parity_multigpu.py -q -m gpu -x -k "matches_oracle or sharded_graph_build or dma-push or superkmer" --durations=5 2>&1 | tail -12
  fi
  run default "" --no-named
  run sgraph "pgraph_sharded=1" --no-named --no-cpu-baseline
  TAILN=8 run dma_sgraph_trace "a2a=2,pgraph_sharded=1,trace=1" --no-named --no-cpu-baseline
  run superkmer_sgraph "wire_superkmer=1,pgraph_sharded=1" --no-named --no-cpu-baseline
} > $LOG 2>&1
cat $LOG

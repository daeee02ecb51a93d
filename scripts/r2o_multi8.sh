#!/bin/bash
# 8-GPU confirmation: sharded parity at 8 ranks, the driver-shaped bench line (parity, named config, sharded graph build), the replicated
# graph build beside it
mkdir -p gpurun_out
NGPU=$(python -c "import torch; print(torch.cuda.device_count())")
LOG=gpurun_out/r2o_multi_${NGPU}.log
run() { # name, tune, extra args...
  name=$1; tune=$2; shift 2
  GENOME_B200_TUNE="$tune" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NGPU --master-addr 127.0.0.1 --master-port 29612 \
    bench.py --gpus $NGPU "$@" > gpurun_out/r2o_n${NGPU}_$name.json 2> gpurun_out/r2o_n${NGPU}_$name.err
  echo "-- $name ($tune $*) rc=$?"
  grep -E "^\[(pmap|sgraph|pgraph)\]" gpurun_out/r2o_n${NGPU}_$name.err | tail -${TAILN:-0}
  grep -v "OMP_NUM_THREADS\|^\*\*\*\*" gpurun_out/r2o_n${NGPU}_$name.err | tail -3 | cut -c1-400
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2o_n${NGPU}_$name.json').read().strip().splitlines()[-1])
    g = d.get('graph') or {}
    print('   %.3f ms/step device, e2e %.3f ms, insert %.3f ms, parity %s, graph build %s ms (kernels %s), simplify %s ms' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['insert_ms'], d.get('parity_checked'), g.get('build_ms'), g.get('build_kernels_ms'), g.get('components_retain_simplify_ms')))
    n = (d.get('named_configs') or {}).get('C3')
    if n: print('   C3 named:', n['ms_per_step'], 'ms/step', n['value'], n.get('identities'), (n.get('graph') or {}).get('build_ms'))
except Exception as e:
    print('   failed', e)
PY
}
{
  echo "== sharded parity tests at $NGPU GPUs"
  timeout 900 python -m pytest tests/test_parity_multigpu.py -q -m gpu -x -k "matches_oracle and $NGPU or sharded_graph_build and $NGPU" --durations=4 2>&1 | tail -12
  run default "" --steps 20 --warmup 3
  TAILN=14 run replicated_graph_trace "pgraph_sharded=0,trace=1" --steps 5 --warmup 3 --no-named --no-cpu-baseline
  TAILN=14 run sharded_graph_trace "trace=1" --steps 5 --warmup 3 --no-named --no-cpu-baseline
} > $LOG 2>&1
cat $LOG

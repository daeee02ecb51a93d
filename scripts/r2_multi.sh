#!/bin/bash
# Multi-GPU validation + bench of the sharded paths:  gpurun --gpus N --timeout 1500 -- 'bash scripts/r2_multi.sh'
# Writes gpurun_out/r2_multi_N.log and gpurun_out/r2_bench_nN_*.json
mkdir -p gpurun_out
export GENOME_B200_UNVALIDATED=1
NGPU=$(python -c "import torch; print(torch.cuda.device_count())")
LOG=gpurun_out/r2_multi_${NGPU}.log
{
  echo "== $NGPU GPUs: sharded parity tests (oracle equality per shard, lookups, graph build)"
  timeout 900 python -m pytest tests/test_parity_multigpu.py -q -m gpu --durations=6 2>&1 | tail -30
  run() { # name, env...
    name=$1; shift
    env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NGPU --master-addr 127.0.0.1 --master-port 29611 \
      bench.py --gpus $NGPU --steps 20 --warmup 3 > gpurun_out/r2_bench_n${NGPU}_$name.json 2> gpurun_out/r2_bench_n${NGPU}_$name.err
    echo "-- bench $name rc=$?"; tail -3 gpurun_out/r2_bench_n${NGPU}_$name.err
  }
  run default X=1
  run sgraph GENOME_B200_PGRAPH=sharded
  run superkmer GENOME_B200_WIRE=superkmer
  run superkmer_sgraph GENOME_B200_WIRE=superkmer GENOME_B200_PGRAPH=sharded
  python - <<PY
import json
for f in ('default', 'sgraph', 'superkmer', 'superkmer_sgraph'):
    try:
        d = json.loads(open('gpurun_out/r2_bench_n${NGPU}_%s.json' % f).read().strip().splitlines()[-1])
        print(f, '%.3f ms/step device' % d['ms_per_step'], '%.3f ms e2e' % d['e2e']['ms_per_step'], 'insert %.3f ms' % d['roofline']['insert_ms'], 'graph', d['graph'] and (d['graph']['build_ms'], d['graph']['components_retain_simplify_ms']))
    except Exception as e:
        print(f, 'failed', e)
PY
} > $LOG 2>&1
tail -40 $LOG

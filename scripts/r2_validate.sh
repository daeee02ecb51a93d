#!/bin/bash
# First GPU call of round 2: everything that was written after round 1's GPU budget ran out, in one gpurun call.
#   gpurun --timeout 2400 -- 'bash scripts/r2_validate.sh'            (1 GPU: virtual shards, graph map handle, CheckGraph)
#   gpurun --gpus 2 --timeout 1500 -- 'bash scripts/r2_validate.sh'   (adds the NCCL + IPC fabric at 2 ranks)
# Writes gpurun_out/r2_validate.log, gpurun_out/r2_sgraph_timing.json.  No -x: every opt-in test reports on its own.
mkdir -p gpurun_out
export GENOME_B200_UNVALIDATED=1
{
  echo "== opt-in device tests (one process per test function: a CUDA fault in one must not take the others down)"
  for sel in "tests/test_sgraph_gpu.py -k virtual_shards_match_oracle" "tests/test_sgraph_gpu.py -k noncanonical_and_cycle" \
             "tests/test_sgraph_gpu.py -k equal_single_gpu_build_at_size" "tests/test_sgraph_gpu.py -k random_dense" \
             "tests/test_graphmap_gpu.py" "tests/test_scripts_gpu.py" \
             "tests/test_countless_gpu.py -k countless_insert_matches_oracle" "tests/test_countless_gpu.py -k overflow_path" \
             "tests/test_countless_gpu.py -k overflow_list" "tests/test_countless_gpu.py -k ragged_stream" \
             "tests/test_countless_gpu.py -k at_size" "tests/test_countless_gpu.py -k superkmer_records" \
             "tests/test_countless_gpu.py -k cas128"; do
    echo "-- $sel"
    timeout 900 python -m pytest $sel -q -m gpu 2>&1 | tail -6
  done
  echo "== ungated parity suite incl. insert_path=direct|partitioned and the full-size C1/C2 oracle equality"
  timeout 1200 python -m pytest tests/test_parity_gpu.py -q -m gpu --durations=8 2>&1 | tail -25
  echo "== sharded graph build through the NCCL fabric, one rank (works on a one-GPU box)"
  timeout 600 python -m pytest tests/test_parity_multigpu.py -q -m gpu -k "(sharded_graph_build or superkmer_wire) and 1" 2>&1 | tail -10
  NGPU=$(python -c "import torch; print(torch.cuda.device_count())")
  if [ "$NGPU" -ge 2 ]; then
    echo "== sharded graph build over $NGPU ranks"
    timeout 900 python -m pytest tests/test_parity_multigpu.py -q -m gpu -k "sharded_graph_build or pmap_matches_oracle or superkmer_wire" 2>&1 | tail -24
  fi
  echo "== bench: default vs single-pass bucket pass (GENOME_B200_COUNTLESS=1)"
  timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
  GENOME_B200_COUNTLESS=1 timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2_bench_countless.json 2> gpurun_out/r2_bench_countless.err
  GENOME_B200_CAS128=1 timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2_bench_cas128.json 2> gpurun_out/r2_bench_cas128.err
  GENOME_B200_CAS128=1 GENOME_B200_COUNTLESS=1 timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2_bench_countless_cas128.json 2> gpurun_out/r2_bench_countless_cas128.err
  python -c "
import json
for f in ('default', 'countless', 'cas128', 'countless_cas128'):
    try:
        d = json.loads(open('gpurun_out/r2_bench_%s.json' % f).read().strip().splitlines()[-1])
        print(f, '%.3f ms/step device' % d['ms_per_step'], '%.3f ms/step e2e' % d['e2e']['ms_per_step'], d['roofline'].get('phases_ms'), d['roofline'].get('insert_ms'))
    except Exception as e:
        print(f, 'failed', e)
"
  echo "== timing: single-GPU build vs virtual shards (C2)"
  timeout 600 python scripts/sgraph_timing.py > gpurun_out/r2_sgraph_timing.json 2> gpurun_out/r2_sgraph_timing.err
  tail -5 gpurun_out/r2_sgraph_timing.json
  tail -5 gpurun_out/r2_sgraph_timing.err
  echo "== timing: the same on a graph that does not fit L2 (C4 x 0.05: ~50 M kept k-mers)"
  timeout 900 python scripts/sgraph_timing.py C4 0.05 > gpurun_out/r2_sgraph_timing_c4.json 2> gpurun_out/r2_sgraph_timing_c4.err
  tail -5 gpurun_out/r2_sgraph_timing_c4.err
} > gpurun_out/r2_validate.log 2>&1
tail -60 gpurun_out/r2_validate.log

#!/bin/bash
mkdir -p gpurun_out
NGPU=$(python -c "import torch; print(torch.cuda.device_count())")
for t in "batches=1" "batches=2" "batches=3" "batches=4" "batches=1,slice_bits=5"; do
  GENOME_B200_TUNE="$t" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NGPU --master-addr 127.0.0.1 --master-port 29612 \
    bench.py --gpus $NGPU --steps 10 --warmup 3 --no-named --no-cpu-baseline --no-graph > gpurun_out/r2n_tmp.json 2> gpurun_out/r2n_tmp.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/r2n_tmp.json').read().strip().splitlines()[-1])
print('$t', 'N=$NGPU  %.3f ms/step device, e2e %.3f ms, insert %.3f ms' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['insert_ms']))
PY
done

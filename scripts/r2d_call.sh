#!/bin/bash
# Round 2 (session c): structure-of-arrays table -- GPU suite, phase times, a short bench, the launch list
mkdir -p gpurun_out
python scripts/insert_phases.py C2 5 2>&1 | tail -6
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2d_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r2d_bench_n1.err
python - <<'P'
import json
d = json.loads(open('gpurun_out/r2d_bench_n1.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'parity', d['parity_checked'])
print(json.dumps(d['roofline']['phases_ms']), d['roofline']['insert_ms'], d['roofline']['insert_vs_64B_sector_roofline'])
print(json.dumps(d['graph']))
print(json.dumps(d['roofline_graph']))
P

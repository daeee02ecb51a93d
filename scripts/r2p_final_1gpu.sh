#!/bin/bash
# Round 2, final 1-GPU evidence: GPU suite, bench (ours + reference arm), launch list, ncu --set full of the two insert kernels.
# (compute-sanitizer is closed on this pool.)   gpurun --timeout 1500 -- 'bash scripts/r2p_final_1gpu.sh'
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2p_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2p_bench_n1.json 2> gpurun_out/r2p_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r2p_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2p_bench_ref.json 2> gpurun_out/r2p_bench_ref.err; echo "ref rc=$?"
python - <<'P'
import json
d = json.loads(open('gpurun_out/r2p_bench_n1.json').read().strip().splitlines()[-1])
r = json.loads(open('gpurun_out/r2p_bench_ref.json').read().strip().splitlines()[-1])
print('ours ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], 'parity', d['parity_checked'], 'launches', d['gpu_launches'])
print('phases', json.dumps(d['roofline']['phases_ms']), 'insert', d['roofline']['insert_ms'], 'vs sector roofline', d['roofline']['insert_vs_64B_sector_roofline'], 'frac', d['roofline']['frac'])
print('reference', r['value'], r['ms_per_step'], 'ratio e2e', d['e2e']['value'] / r['value'])
P
python scripts/insert_phases.py C2 2 > gpurun_out/r2p_phases_plain.log 2>&1 && \
$NCU --set full --import-source on -k regex:"bucket_slabs|insert_slabs" -s 2 -c 2 -o gpurun_out/prof_insert_r2p -f \
    python scripts/insert_phases.py C2 2 > gpurun_out/r2p_ncu_insert.log 2>&1
tail -2 gpurun_out/r2p_ncu_insert.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2p_bench_plain.json 2> gpurun_out/r2p_bench_plain.err && \
$NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file gpurun_out/launches_r2p.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2p_ncu_launches.log 2>&1
ls -la gpurun_out/*r2p*

#!/bin/bash
# Round 2, 1-GPU call: insert experiments, the rewritten bench.py (parity + rooflines), the whole GPU suite, fresh ncu evidence.
mkdir -p gpurun_out
NCU="ncu --clock-control none"
python scripts/r2_insert_sweep.py > gpurun_out/r2b_insert_sweep.jsonl 2> gpurun_out/r2b_insert_sweep.err; echo "sweep rc=$?"; cat gpurun_out/r2b_insert_sweep.jsonl; tail -5 gpurun_out/r2b_insert_sweep.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r2b_bench_n1.err; head -c 3000 gpurun_out/r2b_bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/r2b_bench_ref.json
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b_pytest.log
python scripts/insert_phases.py C2 2 > gpurun_out/r2b_phases_plain.log 2>&1 && \
$NCU --set full --import-source on -k regex:"part_scatter|insert_keys" -s 2 -c 2 -o gpurun_out/prof_insert_r2b -f \
    python scripts/insert_phases.py C2 2 > gpurun_out/r2b_ncu_insert.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_plain.json 2> gpurun_out/r2b_bench_plain.err && \
$NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file gpurun_out/launches_r2b.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_ncu_launches.log 2>&1
ls -la gpurun_out/*r2b*

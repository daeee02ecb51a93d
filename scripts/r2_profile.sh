#!/bin/bash
# ncu evidence for round 2 (ONE GPU; each capture only after the same command has exited 0 without ncu):
#   1. launch list of a short bench run                      -> gpurun_out/launches_r2.csv
#   2. --set full of the insert kernels (default + single pass) -> gpurun_out/prof_insert_r2.ncu-rep
#   3. --set full of the sharded build's heavy kernels through virtual shards -> gpurun_out/prof_sgraph_r2.ncu-rep
#   gpurun --timeout 1700 -- 'bash scripts/r2_profile.sh'
# Summaries for profiles/ are made here afterwards:  python scripts/summarise_profiles.py launches|full ...
mkdir -p gpurun_out
NCU="ncu --clock-control none"
set -x
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_plain.json 2> gpurun_out/bench_r2_plain.err || exit 1
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_r2.log 2>&1

python scripts/insert_phases.py C2 2 > gpurun_out/insert_phases_r2.log 2>&1 || exit 1
$NCU --set full --import-source on -k regex:"part_count|part_scatter|insert_keys" -s 3 -c 3 -o gpurun_out/prof_insert_r2 -f \
    python scripts/insert_phases.py C2 2 > gpurun_out/ncu_insert_r2.log 2>&1
GENOME_B200_COUNTLESS=1 python scripts/insert_phases.py C2 2 > gpurun_out/insert_phases_r2_countless.log 2>&1 && \
GENOME_B200_COUNTLESS=1 $NCU --set full --import-source on -k regex:"part_scatter|insert_keys|make_slab" -s 3 -c 3 -o gpurun_out/prof_insert_countless_r2 -f \
    python scripts/insert_phases.py C2 2 > gpurun_out/ncu_insert_countless_r2.log 2>&1

python scripts/sgraph_timing.py C2 > gpurun_out/sgraph_timing_r2.json 2> gpurun_out/sgraph_timing_r2.err || exit 1
# the 8-rank virtual build: one launch of each heavy functor (masks, classify, jump, segment jump, close, bases)
SG_P=8 $NCU --set full --import-source on -k regex:"items_kernel" -c 160 -o gpurun_out/prof_sgraph_r2 -f \
    python scripts/sgraph_timing.py C2 > gpurun_out/ncu_sgraph_r2.log 2>&1
# paired-end path support (DESIGN 3.6 was timed but never profiled): filter + walk kernels
python scripts/walk_timing.py > gpurun_out/walk_timing_r2.json 2> gpurun_out/walk_timing_r2.err && \
$NCU --set full --import-source on -k regex:"walk_|posmap_" -c 12 -o gpurun_out/prof_walk_r2 -f \
    python scripts/walk_timing.py > gpurun_out/ncu_walk_r2.log 2>&1
ls -la gpurun_out/*.ncu-rep

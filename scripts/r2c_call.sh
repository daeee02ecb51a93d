#!/bin/bash
# Round 2 (session c), 1-GPU call: L2 request-pattern variants, then ncu --set full of the two insert kernels as they ship.
mkdir -p gpurun_out
python scripts/r2c_l2_modes.py > gpurun_out/r2c_l2_modes.jsonl 2> gpurun_out/r2c_l2_modes.err; echo "modes rc=$?"; cat gpurun_out/r2c_l2_modes.jsonl; tail -3 gpurun_out/r2c_l2_modes.err
python scripts/insert_phases.py C2 2 > gpurun_out/r2c_phases_plain.log 2>&1 && \
ncu --clock-control none --set full --import-source on -k regex:"part_scatter|insert_slabs" -s 2 -c 2 -o gpurun_out/prof_insert_r2c -f \
    python scripts/insert_phases.py C2 2 > gpurun_out/r2c_ncu_insert.log 2>&1
tail -4 gpurun_out/r2c_ncu_insert.log
ls -la gpurun_out/*r2c*

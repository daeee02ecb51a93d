#!/bin/bash
# Round 2, GPU call after r2r: the converged per-probe membership kernel and the per-probe ops of the sharded build (PartsOp + ProbeOp +
# CombineOp) -- device tests under both forms, timing on C2 (single GPU and 8 virtual ranks), ncu --set full of the new kernels.
#   gpurun --timeout 240 -- 'bash scripts/r2s_call.sh'
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 110 python -m pytest -x -q -m gpu -p no:cacheprovider --durations=6 tests/test_sgraph_gpu.py \
    "tests/test_parity_gpu.py::test_hash_tie_kmers_in_real_reads" "tests/test_parity_gpu.py::test_build_graph_matches_oracle" \
    "tests/test_parity_gpu.py::test_noncanonical_keys_both_orientations" "tests/test_parity_gpu.py::test_graph_operators_match_oracle" \
    > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/r2s_pytest.log
timeout 60 python scripts/masks_timing.py C2 1.0 6 > gpurun_out/r2s_masks_timing.json 2> gpurun_out/r2s_masks_timing.err; echo "timing rc=$?"
cat gpurun_out/r2s_masks_timing.json; tail -3 gpurun_out/r2s_masks_timing.err
timeout 100 $NCU --set full --import-source on --kernel-name-base mangled -k regex:"masks_flat_kernel|ProbeOp|PartsOp|CombineOp" -c 5 \
    -o gpurun_out/prof_masks_r2s -f python scripts/masks_timing.py C2 1.0 1 > gpurun_out/r2s_ncu_masks.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/r2s_ncu_masks.log
ls -la gpurun_out/ | grep r2s

#!/bin/bash
# Round 2, last GPU call (a few box minutes left): the device tests added since the final 1-GPU evidence (hash-tie k-mers, per-probe
# membership kernel, Kryo graph file), both membership kernels timed on C2, ncu --set full of the membership kernels (single-GPU and
# the sharded build's MasksOp over 8 virtual ranks).   gpurun --timeout 270 -- 'bash scripts/r2r_last_call.sh'
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 100 python -m pytest -x -q -m gpu -p no:cacheprovider \
    "tests/test_parity_gpu.py::test_hash_tie_kmers_in_real_reads" "tests/test_parity_gpu.py::test_build_graph_matches_oracle" \
    "tests/test_parity_gpu.py::test_noncanonical_keys_both_orientations" "tests/test_parity_gpu.py::test_canonical_rule_on_random_kmers" \
    "tests/test_scripts_gpu.py::test_graph_file_round_trip" "tests/test_scripts_gpu.py::test_graph_edit_mutators" \
    > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2r_pytest.log
timeout 60 python scripts/masks_timing.py C2 1.0 6 > gpurun_out/r2r_masks_timing.json 2> gpurun_out/r2r_masks_timing.err; echo "timing rc=$?"
cat gpurun_out/r2r_masks_timing.json; tail -3 gpurun_out/r2r_masks_timing.err
timeout 130 $NCU --set full --import-source on --kernel-name-base mangled -k regex:"masks_kernel|masks_flat_kernel|MasksOp|jump_kernel" -c 12 \
    -o gpurun_out/prof_masks_r2r -f python scripts/masks_timing.py C2 1.0 1 > gpurun_out/r2r_ncu_masks.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2r_ncu_masks.log
ls -la gpurun_out/ | tail -8

#!/bin/bash
# Round 2, GPU call after r2s: sublist list ranking (walk_sublists_kernel + pointer jumping over splitters) -- every graph parity test under
# all three build forms, timing on C2, ncu --set full of the ranking kernels.   gpurun --timeout 200 -- 'bash scripts/r2t_call.sh'
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 110 python -m pytest -x -q -m gpu -p no:cacheprovider --durations=4 tests/test_sgraph_gpu.py \
    "tests/test_parity_gpu.py::test_hash_tie_kmers_in_real_reads" "tests/test_parity_gpu.py::test_build_graph_matches_oracle" \
    "tests/test_parity_gpu.py::test_noncanonical_keys_both_orientations" "tests/test_parity_gpu.py::test_graph_operators_match_oracle" \
    "tests/test_parity_gpu.py::test_perfect_cycle_is_dropped" "tests/test_parity_gpu.py::test_error_free_linear_genome_is_two_edges" \
    > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -10 gpurun_out/r2t_pytest.log
timeout 60 python scripts/masks_timing.py C2 1.0 6 > gpurun_out/r2t_timing.json 2> gpurun_out/r2t_timing.err; echo "timing rc=$?"
cat gpurun_out/r2t_timing.json; tail -3 gpurun_out/r2t_timing.err
MT_ONLY=sublists timeout 60 $NCU --set full --import-source on -k regex:"walk_sublists_kernel|jump_kernel" -c 5 \
    -o gpurun_out/prof_rank_r2t -f python scripts/masks_timing.py C2 1.0 1 > gpurun_out/r2t_ncu.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/r2t_ncu.log
ls -la gpurun_out/ | grep r2t

#!/bin/bash
# compute-sanitizer memcheck over the small parity cases (one GPU): out-of-bounds and misaligned accesses that a passing
# test does not show.  Slow (10-50x): only the small cases, never the full-size ones.
#   gpurun --timeout 1700 -- 'bash scripts/r2_memcheck.sh'
mkdir -p gpurun_out
export GENOME_B200_UNVALIDATED=1
SAN=/usr/local/cuda/bin/compute-sanitizer
{
  for sel in "tests/test_parity_gpu.py -k 'insert_counts or build_graph or operators or noncanonical or perfect_cycle or empty_and_short'" \
             "tests/test_graphmap_gpu.py" "tests/test_walk_gpu.py -k 'not script'" "tests/test_sgraph_gpu.py -k 'match_oracle or noncanonical'"; do
    echo "== memcheck: $sel"
    eval timeout 1500 $SAN --tool memcheck --error-exitcode 86 --print-limit 20 python -m pytest $sel -q -m gpu -x 2>&1 | grep -v "^$" | tail -25
  done
} > gpurun_out/r2_memcheck.log 2>&1
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r2_memcheck.log
tail -40 gpurun_out/r2_memcheck.log

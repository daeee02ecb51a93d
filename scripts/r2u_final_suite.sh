#!/bin/bash
# Round 2, last GPU seconds: the whole device suite on the final code (graph files first: that is where the code changed since the
# r2p run; the insert tests last).   gpurun --timeout 146 -- 'bash scripts/r2u_final_suite.sh'
mkdir -p gpurun_out
timeout 136 python -m pytest -q -m gpu -p no:cacheprovider --durations=8 tests/test_sgraph_gpu.py tests/test_scripts_gpu.py tests/test_graphmap_gpu.py \
    tests/test_walk_gpu.py tests/test_parity_gpu.py tests/test_countless_gpu.py tests/test_parity_multigpu.py > gpurun_out/r2u_pytest.log 2>&1
echo "pytest rc=$?"; tail -16 gpurun_out/r2u_pytest.log

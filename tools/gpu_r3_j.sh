#!/bin/bash
# session 3J: rulebook kernels without per-tap integer divisions: parity (rulebook / Morton / encoder tests), microbench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lidar.py tests/test_gpu_morton.py tests/test_gpu_sparse_tc.py -x -q -m gpu --timeout 600 > gpurun_out/j_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/j_tests.log
timeout 300 python tools/rulebook_bench.py > gpurun_out/j_rulebook.txt 2>&1; echo "rc=$?"; tail -11 gpurun_out/j_rulebook.txt

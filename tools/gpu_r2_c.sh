#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python tools/gpu_nan_hunt.py distill_lidar > gpurun_out/c_nan.log 2>&1
echo "nan hunt rc=$?"; grep -v "^\s*File\|^\s*return\|^\s*self\.\|\^\^\^" gpurun_out/c_nan.log | tail -60
timeout 900 python -m pytest tests/test_gpu_sparse_tc.py tests/test_gpu_lidar.py tests/test_gpu_glue.py tests/test_head_batching.py tests/test_gpu_graph.py tests/test_gpu_assigner.py -m gpu -q > gpurun_out/c_tests.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/c_tests.log

#!/bin/bash
# round-2 GPU session B: debug the step (loss terms, NaN hunt), op-level profile, full tests
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python tools/gpu_debug_step.py distill_lidar 14 --profile > gpurun_out/b_debug_distill_lidar.log 2>&1
echo "debug rc=$?"; grep -E "^step" gpurun_out/b_debug_distill_lidar.log | head -20
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/b_tests.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/b_tests.log

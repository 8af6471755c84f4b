#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python tools/gpu_debug_step.py distill_lidar 12 --profile > gpurun_out/d_debug_distill_lidar.log 2>&1
echo "debug rc=$?"; grep -E "^step" gpurun_out/d_debug_distill_lidar.log | cut -c1-230 | head -14
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/d_tests.log 2>&1
echo "tests rc=$?"; tail -8 gpurun_out/d_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/d_bench.err

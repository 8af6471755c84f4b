#!/bin/bash
# round-2 GPU session A: tests, default bench, launch list of one distillation step
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/a_smi.txt 2>&1
nproc >> gpurun_out/a_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/a_tests.log
tail -5 gpurun_out/a_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/a_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 12000 --csv \
  --log-file gpurun_out/a_launches_distill_lidar.csv python bench.py --steps 1 --warmup 3 --profile --no-secondary --no-cpu-baseline \
  > gpurun_out/a_prof.log 2>&1
echo "ncu rc=$?"

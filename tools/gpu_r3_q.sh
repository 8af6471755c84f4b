#!/bin/bash
# session 3Q: tower conv bias absorbed by the train-mode BatchNorm: head / step / graph tests, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_head_batching.py tests/test_gpu_step.py tests/test_gpu_graph.py tests/test_gpu_glue.py -x -q --timeout 600 > gpurun_out/q_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/q_tests.log
B="python bench.py --steps 8 --warmup 3 --no-secondary --no-cpu-baseline"
timeout 600 $B > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/q_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/q_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["step_ms_min_med_max"], "launches", d["gpu_launches"])
print(d["loss_terms_after_timed_steps"])
PY

#!/bin/bash
# session 3S: tower layer-2 dgrad as an fprop: head / step / graph tests, bench, top dense kernels of the step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_head_batching.py tests/test_gpu_step.py tests/test_gpu_graph.py -x -q --timeout 600 > gpurun_out/s_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/s_tests.log
B="python bench.py --steps 8 --warmup 3 --no-secondary --no-cpu-baseline"
timeout 600 $B > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/s_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/s_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["step_ms_min_med_max"])
print(d["loss_terms_after_timed_steps"])
PY
timeout 600 python tools/gpu_step_timeline.py > gpurun_out/s_timeline.txt 2>&1; grep -n "top kernels" -A22 gpurun_out/s_timeline.txt | cut -c1-150

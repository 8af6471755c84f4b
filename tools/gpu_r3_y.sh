#!/bin/bash
# session 3Y: geometry prefetch with the explicit memory protocol: step test, then the default-flag bench (K=10, W=3)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_step.py -x -q -m gpu --timeout 600 > gpurun_out/y3_tests.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/y3_tests.log
timeout 600 python bench.py --no-secondary --no-cpu-baseline > gpurun_out/y3_bench.json 2> gpurun_out/y3_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/y3_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/y3_bench.json").read().strip().splitlines()[-1])
print(d["steps"], d["warmup"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
print("   step_ms", d.get("step_ms"))
PY

#!/bin/bash
# session S: batched head loss + graphed ground-truth preparation
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_glue.py tests/test_head_batching.py -x -q -m gpu --timeout 600 > gpurun_out/s_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/s_tests.log
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/s_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/s_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"], "launches", d["gpu_launches"])
print({k: (v.get("value"), v.get("ms_per_step")) for k, v in d.get("secondary", {}).items() if isinstance(v, dict)})
print(d["loss_terms_after_timed_steps"])
PY
timeout 600 python tools/gpu_step_timeline.py distill_lidar > gpurun_out/timeline_stdout.txt 2>&1
head -16 gpurun_out/timeline_distill_lidar.txt

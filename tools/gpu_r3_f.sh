#!/bin/bash
# session 3F: wgrad descriptors hoisted (one per operand and stage); NG table back to 4
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sparse_tc.py tests/test_gpu_lidar.py -x -q -m gpu --timeout 600 > gpurun_out/f_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/f_tests.log
B="python bench.py --steps 8 --warmup 3 --no-secondary --no-cpu-baseline"
timeout 600 $B > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/f_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/f_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["step_ms_min_med_max"])
for e in d["roofline_all"][:24]:
    print("  %-72s %6.3f ms/step  %7.4f ms/call  frac %.4f" % (e["kernel"], e["ms_per_step"], e["avg_call_ms"], e.get("frac", 0)))
PY

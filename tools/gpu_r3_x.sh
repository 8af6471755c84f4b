#!/bin/bash
# session 3X: next step's geometry built on a side stream under the backward: test + A/B bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_step.py -x -q -m gpu --timeout 600 > gpurun_out/x3_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/x3_tests.log
B="python bench.py --steps 12 --warmup 4 --no-secondary --no-cpu-baseline"
UD_PREFETCH_GEOMETRY=0 timeout 600 $B > gpurun_out/x3_bench_0.json 2> gpurun_out/x3_bench_0.err; echo "bench 0 rc=$?"; tail -2 gpurun_out/x3_bench_0.err
timeout 600 $B > gpurun_out/x3_bench_1.json 2> gpurun_out/x3_bench_1.err; echo "bench 1 rc=$?"; tail -2 gpurun_out/x3_bench_1.err
python - <<'PY'
import json
for n in ("0", "1"):
    d = json.loads(open("gpurun_out/x3_bench_%s.json" % n).read().strip().splitlines()[-1])
    print("prefetch", n, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["step_ms_min_med_max"], d["loss_terms_after_timed_steps"]["loss_feature"])
    print("   step_ms", d.get("step_ms"))
PY

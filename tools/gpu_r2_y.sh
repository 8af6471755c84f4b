#!/bin/bash
# session Y: teacher camera branch on a side stream
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_glue.py -x -q -m gpu --timeout 600 > gpurun_out/y_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/y_tests.log
B="python bench.py --steps 8 --warmup 3 --no-secondary --no-cpu-baseline"
UD_OVERLAP_CAMERA=0 timeout 600 $B > gpurun_out/y_bench_inline.json 2> gpurun_out/y_bench_inline.err; echo "inline rc=$?"
timeout 600 $B > gpurun_out/y_bench_side.json 2> gpurun_out/y_bench_side.err; echo "side rc=$?"; tail -2 gpurun_out/y_bench_side.err
python - <<'PY'
import json
for f in ("y_bench_inline", "y_bench_side"):
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["step_ms_min_med_max"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"].get("traffic"))
    print("   ", d["loss_terms_after_timed_steps"])
PY

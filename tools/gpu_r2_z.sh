#!/bin/bash
# session Z: gradient stealing (no per-parameter accumulate kernels)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_step.py -x -q -m gpu --timeout 600 > gpurun_out/z_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/z_tests.log
B="python bench.py --steps 8 --warmup 3 --no-secondary --no-cpu-baseline"
UD_GRAD_STEAL=0 timeout 600 $B > gpurun_out/z_bench_views.json 2> gpurun_out/z_bench_views.err; echo "views rc=$?"
timeout 600 $B > gpurun_out/z_bench_steal.json 2> gpurun_out/z_bench_steal.err; echo "steal rc=$?"; tail -2 gpurun_out/z_bench_steal.err
python - <<'PY'
import json
for f in ("z_bench_views", "z_bench_steal"):
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["step_ms_min_med_max"])
    print("   ", d["loss_terms_after_timed_steps"])
PY

#!/bin/bash
# session 3R: geometry-first rulebooks: encoder parity, bench with / without, timeline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lidar.py tests/test_gpu_step.py tests/test_reference_over_shim.py -x -q --timeout 600 > gpurun_out/r_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r_tests.log
B="python bench.py --steps 8 --warmup 3 --no-secondary --no-cpu-baseline"
UD_PREBUILD_RULEBOOKS=0 timeout 600 $B > gpurun_out/r_bench_lazy.json 2> gpurun_out/r_bench_lazy.err; echo "bench lazy rc=$?"
timeout 600 $B > gpurun_out/r_bench_pre.json 2> gpurun_out/r_bench_pre.err; echo "bench prebuild rc=$?"; tail -2 gpurun_out/r_bench_pre.err
python - <<'PY'
import json
for n in ("lazy", "pre"):
    d = json.loads(open("gpurun_out/r_bench_%s.json" % n).read().strip().splitlines()[-1])
    print(n, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["step_ms_min_med_max"])
PY
timeout 600 python tools/gpu_step_timeline.py > gpurun_out/r_timeline.txt 2>&1; head -16 gpurun_out/r_timeline.txt | tail -14

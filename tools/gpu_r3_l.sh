#!/bin/bash
# session 3L: NHWC-native box distill + rulebook kernels: distill / step tests, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_distill.py tests/test_gpu_step.py tests/test_gpu_graph.py -x -q -m gpu --timeout 600 > gpurun_out/l_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/l_tests.log
B="python bench.py --steps 8 --warmup 3 --no-secondary --no-cpu-baseline"
timeout 600 $B > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/l_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/l_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["step_ms_min_med_max"], "launches", d["gpu_launches"])
print(d["loss_terms_after_timed_steps"])
for e in d["roofline_all"]:
    if "distill" in e["kernel"] or "rulebook" in e["kernel"]:
        print("  %-60s %6.3f ms/step  %7.4f ms/call  frac %.4f" % (e["kernel"], e["ms_per_step"], e["avg_call_ms"], e.get("frac", 0)))
PY

#!/bin/bash
# last validation of the round: all GPU tests, smoke, the default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/last_tests.log 2>&1
echo "gpu tests rc=$?"; tail -2 gpurun_out/last_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/last_smoke.log 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/last_smoke.log
timeout 900 python bench.py > gpurun_out/last_bench_default.json 2> gpurun_out/last_bench_default.err
echo "bench default rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/last_bench_default.json").read().strip().splitlines()[-1])
print(d.get("value"), d.get("ms_per_step"), "e2e", d.get("e2e", {}).get("value"), "cpu", d.get("cpu_baseline", {}).get("value"),
      "parity", d.get("parity", {}).get("ok"), d.get("parity", {}).get("bev_rel_err"), "clocks", d.get("clocks"), "launches", d.get("gpu_launches"))
print(d["roofline"])
for k, v in d.get("secondary", {}).items():
    if isinstance(v, dict): print("   ", k, v.get("value"), v.get("ms_per_step"))
PY

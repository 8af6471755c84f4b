#!/bin/bash
# session V: eval-BN / residual / ReLU folded into the conv epilogue for frozen encoders
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_lidar.py tests/test_gpu_sparse_tc.py tests/test_gpu_step.py tests/test_gpu_glue.py -x -q -m gpu --timeout 600 > gpurun_out/v_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/v_tests.log
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/v_bench.err
timeout 600 python bench.py --workload lidar --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/v_bench_lidar.json 2> gpurun_out/v_bench_lidar.err
python - <<'PY'
import json
for f in ("gpurun_out/v_bench.json", "gpurun_out/v_bench_lidar.json"):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", d.get("parity"))
    print({k: (v.get("value"), v.get("ms_per_step")) for k, v in d.get("secondary", {}).items() if isinstance(v, dict)})
    for e in d["roofline_all"][:8]:
        print("  %-70s %6.3f ms/step  %7.4f ms/call  frac %.4f" % (e["kernel"], e["ms_per_step"], e["avg_call_ms"], e["frac"]))
PY

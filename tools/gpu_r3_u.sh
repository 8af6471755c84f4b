#!/bin/bash
# session 3U: weights packed from the parameter layout: parity, A/B bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sparse_tc.py tests/test_gpu_lidar.py tests/test_gpu_step.py -x -q -m gpu --timeout 600 > gpurun_out/u_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/u_tests.log
B="python bench.py --steps 8 --warmup 3 --no-secondary --no-cpu-baseline"
UD_SPCONV_PACK_FROM_PARAM=0 timeout 600 $B > gpurun_out/u_bench_0.json 2> gpurun_out/u_bench_0.err; echo "bench 0 rc=$?"
timeout 600 $B > gpurun_out/u_bench_1.json 2> gpurun_out/u_bench_1.err; echo "bench 1 rc=$?"; tail -2 gpurun_out/u_bench_1.err
python - <<'PY'
import json
for n in ("0", "1"):
    d = json.loads(open("gpurun_out/u_bench_%s.json" % n).read().strip().splitlines()[-1])
    print("pack_from_param", n, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["step_ms_min_med_max"], "launches", d["gpu_launches"])
PY

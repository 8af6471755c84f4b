#!/bin/bash
# session 3W: zero bias gradient under a train-mode BatchNorm (no column-sum pass): tests, A/B bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lidar.py tests/test_gpu_step.py -x -q -m gpu --timeout 600 > gpurun_out/w_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/w_tests.log
B="python bench.py --steps 8 --warmup 3 --no-secondary --no-cpu-baseline"
UD_ZERO_BIAS_GRAD_UNDER_BN=0 timeout 600 $B > gpurun_out/w_bench_0.json 2> gpurun_out/w_bench_0.err; echo "bench 0 rc=$?"
timeout 600 $B > gpurun_out/w_bench_1.json 2> gpurun_out/w_bench_1.err; echo "bench 1 rc=$?"; tail -2 gpurun_out/w_bench_1.err
python - <<'PY'
import json
for n in ("0", "1"):
    d = json.loads(open("gpurun_out/w_bench_%s.json" % n).read().strip().splitlines()[-1])
    st = [e for e in d["roofline_all"] if e["entry"] == "ud_bn_stats"]
    print("zero_bias_grad", n, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["step_ms_min_med_max"], "bn_stats calls", st[0]["calls_per_step"] if st else 0)
PY

#!/bin/bash
# session 3H: float4 BN backward (parity + bench); pre-split rows under the two-CTA conv kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sparse_tc.py tests/test_gpu_lidar.py tests/test_gpu_step.py -x -q -m gpu --timeout 600 > gpurun_out/h_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/h_tests.log
B="python bench.py --steps 8 --warmup 3 --no-secondary --no-cpu-baseline"
timeout 600 $B > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/h_bench.err
UD_SPCONV_PRESPLIT=1 timeout 600 $B > gpurun_out/h_bench_presplit.json 2> gpurun_out/h_bench_presplit.err; echo "bench presplit rc=$?"; tail -2 gpurun_out/h_bench_presplit.err
python - <<'PY'
import json
rows = {}
for v in ("h_bench", "h_bench_presplit"):
    d = json.loads(open("gpurun_out/%s.json" % v).read().strip().splitlines()[-1])
    print(v, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["step_ms_min_med_max"])
    for e in d["roofline_all"]:
        rows.setdefault(e["kernel"], {})[v] = (e["avg_call_ms"], e["ms_per_step"], e.get("frac", 0))
print("%-72s %9s %9s   (ms/call; ms/step; frac)" % ("kernel", "default", "presplit"))
for k, v in list(rows.items())[:30]:
    if len(v) == 2:
        print("%-72s %9.4f %9.4f   %6.3f %6.3f  %.3f" % (k, v["h_bench"][0], v["h_bench_presplit"][0], v["h_bench"][1], v["h_bench_presplit"][1], v["h_bench"][2]))
PY

#!/bin/bash
# session 3E: weight-gradient kernel with two CTAs per SM (NG 2): parity, then bench with NG forced to 4 vs the table
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sparse_tc.py tests/test_gpu_lidar.py -x -q -m gpu --timeout 600 > gpurun_out/e_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/e_tests.log
B="python bench.py --steps 8 --warmup 3 --no-secondary --no-cpu-baseline"
for n in 4 0; do
  UD_WG_NG=$n timeout 600 $B > gpurun_out/e_bench_$n.json 2> gpurun_out/e_bench_$n.err; echo "bench ng=$n rc=$?"; tail -2 gpurun_out/e_bench_$n.err
done
python - <<'PY'
import json
rows = {}
for n in (4, 0):
    d = json.loads(open("gpurun_out/e_bench_%d.json" % n).read().strip().splitlines()[-1])
    print("ng", n, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["step_ms_min_med_max"], d["loss_terms_after_timed_steps"])
    for e in d["roofline_all"]:
        rows.setdefault(e["kernel"], {})[n] = (e["avg_call_ms"], e["ms_per_step"])
print("%-72s %9s %9s   (ms/call; ms/step)" % ("kernel", "ng4", "table"))
for k, v in rows.items():
    if "wgrad" in k and len(v) == 2:
        print("%-72s %9.4f %9.4f   %6.3f %6.3f" % (k, v[4][0], v[0][0], v[4][1], v[0][1]))
PY

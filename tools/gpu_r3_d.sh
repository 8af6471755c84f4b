#!/bin/bash
# session 3D: two CTAs per SM (NSETS 2) for the conv kernel: parity, then bench with NSETS forced to 4 / forced to 2 / table
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sparse_tc.py tests/test_gpu_lidar.py -x -q -m gpu --timeout 600 > gpurun_out/d_tests.log 2>&1
echo "tests(table) rc=$?"; tail -3 gpurun_out/d_tests.log
UD_DG_NSETS=2 timeout 900 python -m pytest tests/test_gpu_sparse_tc.py -x -q -m gpu --timeout 600 > gpurun_out/d_tests2.log 2>&1
echo "tests(nsets 2 everywhere) rc=$?"; tail -3 gpurun_out/d_tests2.log
B="python bench.py --steps 8 --warmup 3 --no-secondary --no-cpu-baseline"
for n in 4 2 0; do
  UD_DG_NSETS=$n timeout 600 $B > gpurun_out/d_bench_$n.json 2> gpurun_out/d_bench_$n.err; echo "bench nsets=$n rc=$?"; tail -2 gpurun_out/d_bench_$n.err
done
python - <<'PY'
import json
rows = {}
for n in (4, 2, 0):
    d = json.loads(open("gpurun_out/d_bench_%d.json" % n).read().strip().splitlines()[-1])
    print("nsets", n, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["step_ms_min_med_max"])
    for e in d["roofline_all"]:
        rows.setdefault(e["kernel"], {})[n] = (e["avg_call_ms"], e["ms_per_step"])
print("%-72s %9s %9s %9s   (ms/call; ms/step)" % ("kernel", "nsets4", "nsets2", "table"))
for k, v in rows.items():
    if "spconv" in k and len(v) == 3:
        print("%-72s %9.4f %9.4f %9.4f   %6.3f %6.3f %6.3f" % (k, v[4][0], v[2][0], v[0][0], v[4][1], v[2][1], v[0][1]))
PY

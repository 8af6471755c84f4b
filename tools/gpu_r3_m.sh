#!/bin/bash
# session 3M: weight gradient with five producer groups (NG 5) vs four
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
UD_WG_NG=5 timeout 900 python -m pytest tests/test_gpu_sparse_tc.py -x -q -m gpu --timeout 600 -k "wgrad or backward or grad" > gpurun_out/m_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/m_tests.log
B="python bench.py --steps 6 --warmup 3 --no-secondary --no-cpu-baseline"
for n in 5 0; do
  UD_WG_NG=$n timeout 600 $B > gpurun_out/m_bench_$n.json 2> gpurun_out/m_bench_$n.err; echo "bench ng=$n rc=$?"; tail -2 gpurun_out/m_bench_$n.err
done
python - <<'PY'
import json
rows = {}
for n in (5, 0):
    d = json.loads(open("gpurun_out/m_bench_%d.json" % n).read().strip().splitlines()[-1])
    print("ng", n, d["value"], d["ms_per_step"])
    for e in d["roofline_all"]:
        rows.setdefault(e["kernel"], {})[n] = (e["avg_call_ms"], e["ms_per_step"])
for k, v in rows.items():
    if "wgrad" in k and len(v) == 2:
        print("%-60s ng5 %9.4f  ng4 %9.4f" % (k, v[5][0], v[0][0]))
PY

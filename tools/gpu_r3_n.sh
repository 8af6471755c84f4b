#!/bin/bash
# session 3N: weight gradient with index-segment prefetch vs without
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sparse_tc.py -x -q -m gpu --timeout 600 -k "wgrad or backward or grad" > gpurun_out/n_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/n_tests.log
B="python bench.py --steps 6 --warmup 3 --no-secondary --no-cpu-baseline"
timeout 600 $B > gpurun_out/n_bench_pf.json 2> gpurun_out/n_bench_pf.err; echo "bench pf rc=$?"
UD_B200_LIB=$PWD/cvpr2023-unidistill_b200/variants/lib_wgnopf.so timeout 600 $B > gpurun_out/n_bench_nopf.json 2> gpurun_out/n_bench_nopf.err; echo "bench nopf rc=$?"
python - <<'PY'
import json
rows = {}
for n in ("pf", "nopf"):
    d = json.loads(open("gpurun_out/n_bench_%s.json" % n).read().strip().splitlines()[-1])
    print(n, d["value"], d["ms_per_step"])
    for e in d["roofline_all"]:
        rows.setdefault(e["kernel"], {})[n] = (e["avg_call_ms"], e["ms_per_step"])
for k, v in rows.items():
    if "wgrad" in k and len(v) == 2:
        print("%-60s prefetch %9.4f  none %9.4f" % (k, v["pf"][0], v["nopf"][0]))
PY

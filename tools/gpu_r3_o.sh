#!/bin/bash
# session 3O: weight gradient: + dout prefetch
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 6 --warmup 3 --no-secondary --no-cpu-baseline"
timeout 600 $B > gpurun_out/o_bench_pf.json 2> gpurun_out/o_bench_pf.err; echo "bench pf rc=$?"
UD_B200_LIB=$PWD/cvpr2023-unidistill_b200/variants/lib_wgpfd.so timeout 600 $B > gpurun_out/o_bench_pfd.json 2> gpurun_out/o_bench_pfd.err; echo "bench pfd rc=$?"
python - <<'PY'
import json
rows = {}
for n in ("pf", "pfd"):
    d = json.loads(open("gpurun_out/o_bench_%s.json" % n).read().strip().splitlines()[-1])
    print(n, d["value"], d["ms_per_step"])
    for e in d["roofline_all"]:
        rows.setdefault(e["kernel"], {})[n] = (e["avg_call_ms"], e["ms_per_step"])
for k, v in rows.items():
    if "wgrad" in k and len(v) == 2:
        print("%-60s idx prefetch %9.4f  + dout %9.4f" % (k, v["pf"][0], v["pfd"][0]))
PY

#!/bin/bash
# session 3G: which side binds the weight-gradient kernel: no gathers (wg1), no MMAs (wg2), no gathers + no smem stores (wg3)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-secondary --no-cpu-baseline"
for v in wg1 wg2 wg3; do
  UD_B200_LIB=$PWD/cvpr2023-unidistill_b200/variants/lib_$v.so timeout 600 $B > gpurun_out/g_bench_$v.json 2> gpurun_out/g_bench_$v.err; echo "bench $v rc=$?"; tail -2 gpurun_out/g_bench_$v.err
done
python - <<'PY'
import json
rows = {}
for v in ("wg1", "wg2", "wg3"):
    d = json.loads(open("gpurun_out/g_bench_%s.json" % v).read().strip().splitlines()[-1])
    for e in d["roofline_all"]:
        rows.setdefault(e["kernel"], {})[v] = e["avg_call_ms"]
print("%-60s %9s %9s %9s" % ("kernel (ms/call)", "no gather", "no MMA", "no g+st"))
for k, v in rows.items():
    if "wgrad" in k and len(v) == 3:
        print("%-60s %9.4f %9.4f %9.4f" % (k, v["wg1"], v["wg2"], v["wg3"]))
PY

#!/bin/bash
# session 3C: one mbarrier arrive per producer warp (dg, window, wgrad kernels): parity, role cycles, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sparse_tc.py tests/test_gpu_lidar.py -x -q -m gpu --timeout 600 > gpurun_out/c_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/c_tests.log
for v in probe exp3; do
  UD_B200_LIB=$PWD/cvpr2023-unidistill_b200/variants/lib_$v.so timeout 300 python tools/gpu_dg_probe.py > gpurun_out/dg_${v}_warparrive.txt 2>&1
  echo "== $v"; tail -22 gpurun_out/dg_${v}_warparrive.txt | cut -c1-210 | grep -v "^  480000\|  910319  27   16\|120140"
done
timeout 900 python bench.py --steps 8 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/c_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
for e in d["roofline_all"][:16]:
    print("  %-50s %6.3f ms/step  %7.4f ms/call  frac %.4f" % (e["kernel"], e["ms_per_step"], e["avg_call_ms"], e["frac"]))
PY

#!/bin/bash
# session T: L2 prefetch warp in the conv kernel: parity, role cycles, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sparse_tc.py tests/test_gpu_lidar.py -x -q -m gpu --timeout 600 > gpurun_out/t_tests.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/t_tests.log
UD_B200_LIB=$PWD/cvpr2023-unidistill_b200/variants/lib_probe.so timeout 300 python tools/gpu_dg_probe.py > gpurun_out/dg_probe_prefetch.txt 2>&1
tail -22 gpurun_out/dg_probe_prefetch.txt | cut -c1-200
timeout 900 python bench.py --steps 8 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/t_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/t_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
for e in d["roofline_all"][:14]:
    print("  %-50s %6.3f ms/step  %7.4f ms/call  frac %.4f" % (e["kernel"], e["ms_per_step"], e["avg_call_ms"], e["frac"]))
PY

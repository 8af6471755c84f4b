#!/bin/bash
# session U: dual accumulation (2 MMAs per 16 channels for narrow outputs): parity, role cycles, bench per variant
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sparse_tc.py tests/test_gpu_lidar.py -x -q -m gpu --timeout 600 > gpurun_out/u_tests.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/u_tests.log
UD_B200_LIB=$PWD/cvpr2023-unidistill_b200/variants/lib_probe.so timeout 300 python tools/gpu_dg_probe.py > gpurun_out/dg_probe_dual.txt 2>&1
tail -22 gpurun_out/dg_probe_dual.txt | cut -c1-230
B="python bench.py --steps 8 --warmup 3 --no-secondary --no-cpu-baseline"
timeout 900 $B > gpurun_out/u_bench_dual64.json 2> gpurun_out/u_bench_dual64.err; echo "dual64 rc=$?"
for v in 32 0; do
  UD_B200_LIB=$PWD/cvpr2023-unidistill_b200/variants/lib_dual$v.so timeout 900 $B > gpurun_out/u_bench_dual$v.json 2> gpurun_out/u_bench_dual$v.err; echo "dual$v rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/u_bench_dual*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    conv = {e["kernel"].split(" ")[0].replace("k_spconv_umma_", ""): round(e["avg_call_ms"], 4) for e in d["roofline_all"] if "umma_dg" in e["kernel"]}
    print(f.split("/")[-1], d["value"], d["ms_per_step"], conv, "parity", d.get("parity", {}).get("bev_rel_err"))
PY

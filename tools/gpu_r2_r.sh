#!/bin/bash
# session R: graphed student trunk (tests + bench, with and without) and gather-warp variants of the conv kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_step.py -x -q -m gpu --timeout 600 > gpurun_out/r_tests_step.log 2>&1
echo "step tests rc=$?"; tail -3 gpurun_out/r_tests_step.log
B="python bench.py --steps 6 --warmup 3 --no-secondary --no-cpu-baseline"
UD_STUDENT_GRAPH=0 timeout 600 $B > gpurun_out/r_bench_nograph.json 2> gpurun_out/r_bench_nograph.err
echo "nograph rc=$?"
timeout 600 $B > gpurun_out/r_bench_graph.json 2> gpurun_out/r_bench_graph.err
echo "graph rc=$?"; tail -3 gpurun_out/r_bench_graph.err
for v in 0 1 3 7 9; do
  UD_B200_LIB=$PWD/cvpr2023-unidistill_b200/variants/lib_v$v.so timeout 600 $B > gpurun_out/r_bench_v$v.json 2> gpurun_out/r_bench_v$v.err
  echo "v$v rc=$?"
done
UD_B200_LIB=$PWD/cvpr2023-unidistill_b200/variants/lib_v7.so timeout 900 python -m pytest tests/test_gpu_sparse_tc.py -x -q -m gpu --timeout 600 > gpurun_out/r_tests_v7.log 2>&1
echo "v7 parity rc=$?"; tail -2 gpurun_out/r_tests_v7.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    conv = {e["kernel"].split(" ")[0]: e["ms_per_step"] for e in d["roofline_all"] if "spconv" in e["kernel"]}
    print(f.split("/")[-1], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], {k: conv[k] for k in list(conv)[:6]})
PY
timeout 600 python tools/gpu_step_timeline.py distill_lidar > gpurun_out/timeline_stdout.txt 2>&1
head -16 gpurun_out/timeline_distill_lidar.txt

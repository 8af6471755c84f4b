#!/bin/bash
# session 3T: stem convolution with a zero fourth input channel (frozen teacher), new tests (prebuilt rulebooks, NSETS 2 == 4)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_graph.py tests/test_fold_bn.py "tests/test_gpu_lidar.py::test_prebuilt_rulebooks_equal_lazy_ones" "tests/test_gpu_sparse_tc.py::test_one_and_two_ctas_per_sm_give_identical_outputs" -x -q --timeout 600 > gpurun_out/t_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/t_tests.log
B="python bench.py --steps 8 --warmup 3 --no-secondary --no-cpu-baseline"
timeout 600 $B > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/t_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/t_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["step_ms_min_med_max"])
PY

#!/bin/bash
# session 3V: projection-shortcut bias folded into the block's last convolution: test + bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused_dense.py tests/test_gpu_step.py tests/test_gpu_graph.py -x -q -m gpu --timeout 600 > gpurun_out/v_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/v_tests.log
B="python bench.py --steps 8 --warmup 3 --no-secondary --no-cpu-baseline"
timeout 600 $B > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/v_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/v_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["step_ms_min_med_max"])
PY

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_morton.py tests/test_gpu_lidar.py tests/test_gpu_sparse_tc.py -m gpu -q -s > gpurun_out/e_tests.log 2>&1
echo "tests rc=$?"; grep -E "passed|failed|worst parameter|Error|assert" gpurun_out/e_tests.log | tail -12
timeout 600 python tools/gpu_debug_step.py distill_lidar 84 > gpurun_out/e_debug_distill_lidar.log 2>&1
echo "debug rc=$?"; grep -E "^step|^   |largest" gpurun_out/e_debug_distill_lidar.log | cut -c1-200 | tail -32
timeout 900 python bench.py --steps 10 --warmup 3 --no-secondary > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/e_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/e_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity'))
for e in d['roofline_all'][:14]: print('%8.3f ms x%-5s %-55s %s %s'%(e['ms_per_step'],e['calls_per_step'],e['kernel'],e.get('achieved'),e.get('frac')))
PY

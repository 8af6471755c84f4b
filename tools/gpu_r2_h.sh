#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sparse_tc.py tests/test_gpu_lidar.py tests/test_gpu_morton.py -m gpu -q -x > gpurun_out/j_tests.log 2>&1
echo "tests rc=$?"; grep -E "passed|failed|^E  " gpurun_out/j_tests.log | cut -c1-250 | tail -8
timeout 900 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/j_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/j_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['loss_terms_after_timed_steps']['loss'])
for e in d['roofline_all'][:16]: print('%8.3f ms x%-5s %-55s %s %s'%(e['ms_per_step'],e['calls_per_step'],e['kernel'],e.get('achieved'),e.get('frac')))
PY

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sparse_tc.py -m gpu -q -x --timeout 180 -k "window" > gpurun_out/n_tests_win.log 2>&1
echo "window tests rc=$?"; grep -E "passed|failed|^E  |Timeout" gpurun_out/n_tests_win.log | cut -c1-250 | tail -8
timeout 900 python -m pytest tests/test_gpu_sparse_tc.py tests/test_gpu_lidar.py -m gpu -q --timeout 300 > gpurun_out/n_tests.log 2>&1
echo "tests rc=$?"; grep -E "passed|failed|^E  |Timeout" gpurun_out/n_tests.log | cut -c1-250 | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/n_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/n_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['loss_terms_after_timed_steps']['loss'])
for e in d['roofline_all'][:12]: print('%8.3f ms x%-5s %-55s %s %s'%(e['ms_per_step'],e['calls_per_step'],e['kernel'],e.get('achieved'),e.get('frac')))
PY

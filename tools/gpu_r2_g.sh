#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_step.py tests/test_gpu_graph.py tests/test_head_batching.py tests/test_gpu_lidar.py tests/test_input_side.py tests/test_gpu_glue.py -m gpu -q -s > gpurun_out/g_tests.log 2>&1
echo "tests rc=$?"; grep -E "passed|failed|tensor-core path|^E  " gpurun_out/g_tests.log | cut -c1-330 | tail -12
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/g_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity',{}).get('bev_rel_err'), d['loss_terms_after_timed_steps']['loss'])
for k,v in d.get('secondary',{}).items(): print(k, v['value'], v['ms_per_step'], v['e2e']['value'])
for e in d['roofline_all'][:12]: print('%8.3f ms x%-5s %-55s %s %s'%(e['ms_per_step'],e['calls_per_step'],e['kernel'],e.get('achieved'),e.get('frac')))
PY

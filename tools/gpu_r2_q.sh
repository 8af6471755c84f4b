#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/q_tests.log 2>&1
echo "tests rc=$?"; grep -E "passed|failed|^E  |Timeout" gpurun_out/q_tests.log | cut -c1-250 | tail -8
timeout 1800 python bench.py --all --steps 5 --warmup 3 > gpurun_out/q_bench_all.json 2> gpurun_out/q_bench_all.err
echo "bench --all rc=$?"; tail -c 500 gpurun_out/q_bench_all.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/q_bench_all.json').read().strip().splitlines()[-1])
print('primary', d['value'], d['ms_per_step'], 'e2e', d['e2e'], d.get('parity',{}).get('bev_rel_err'), d.get('cpu_baseline'))
for k,v in d.get('secondary',{}).items():
    if isinstance(v, dict): print(k, v.get('value'), v.get('ms_per_step'), v.get('e2e'))
    else: print(k, [(s['points_per_sample'], s['bev'], s['value']) for s in v])
PY

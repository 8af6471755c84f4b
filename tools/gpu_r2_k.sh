#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/k_tests.log 2>&1
echo "tests rc=$?"; grep -E "passed|failed|^E  " gpurun_out/k_tests.log | cut -c1-250 | tail -8
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/k_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/k_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['loss_terms_after_timed_steps']['loss'])
for k,v in d.get('secondary',{}).items(): print(k, v['value'], v['ms_per_step'], v['e2e']['value'])
for e in d['roofline_all'][:22]: print('%8.3f ms x%-5s %-55s %s %s'%(e['ms_per_step'],e['calls_per_step'],e['kernel'],e.get('achieved'),e.get('frac')))
PY

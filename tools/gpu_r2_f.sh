#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/f_tests.log 2>&1
echo "tests rc=$?"; grep -E "passed|failed|worst parameter|^E  " gpurun_out/f_tests.log | tail -12
timeout 300 python tools/gpu_head_microbench.py > gpurun_out/f_head_mb.txt 2>&1; cat gpurun_out/f_head_mb.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/f_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/f_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity',{}).get('bev_rel_err'), d['loss_terms_after_timed_steps']['loss'])
for k,v in d.get('secondary',{}).items(): print(k, v['value'], v['ms_per_step'], v['e2e']['value'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 12000 --csv \
  --log-file gpurun_out/f_launches_distill_lidar.csv python bench.py --steps 1 --warmup 3 --profile --no-secondary --no-cpu-baseline \
  > gpurun_out/f_prof.log 2>&1
echo "ncu rc=$?"

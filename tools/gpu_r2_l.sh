#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python bench.py --all --steps 5 --warmup 3 > gpurun_out/l_bench_all.json 2> gpurun_out/l_bench_all.err
echo "bench --all rc=$?"; tail -c 500 gpurun_out/l_bench_all.err
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_spconv_umma_dg -c 27 \
  --csv --page raw --log-file gpurun_out/l_ncu_full_dg.csv python bench.py --steps 1 --warmup 3 --profile --no-secondary --no-cpu-baseline \
  > gpurun_out/l_ncu.log 2>&1
echo "ncu full rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 12000 --csv \
  --log-file gpurun_out/l_launches_distill_lidar.csv python bench.py --steps 1 --warmup 3 --profile --no-secondary --no-cpu-baseline \
  > gpurun_out/l_prof.log 2>&1
echo "ncu launches rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/l_bench_all.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
for k,v in d.get('secondary',{}).items():
    if isinstance(v, dict): print(k, v.get('value'), v.get('ms_per_step'))
    else: print(k, [(s['points_per_sample'], s['bev'], s['value']) for s in v])
PY

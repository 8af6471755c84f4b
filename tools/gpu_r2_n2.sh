#!/bin/bash
# N-GPU scaling probe of the headline step: torchrun, one rank per GPU
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > gpurun_out/n${N}_smi.txt 2>&1
nproc >> gpurun_out/n${N}_smi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 --no-secondary > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err
echo "bench N=$N rc=$?"; tail -c 600 gpurun_out/n${N}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/n${N}_bench.json').read().strip().splitlines()[-1])
print('N=%d value %.2f  ms/step %.2f  e2e %.2f  ranks %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['rank_ms_per_step_min_med_max']))
print(d['allreduce']); print(d['clocks']); print(d['loss_terms_after_timed_steps'])
PY

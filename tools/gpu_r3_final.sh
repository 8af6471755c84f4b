#!/bin/bash
# final validation + measurement set of round 2, second session (r03 files)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/f3_tests.log 2>&1
echo "gpu tests rc=$?"; tail -2 gpurun_out/f3_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f3_smoke.log 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/f3_smoke.log
timeout 900 python bench.py > gpurun_out/f3_bench_default.json 2> gpurun_out/f3_bench_default.err
echo "bench default rc=$?"
timeout 900 python bench.py --impl reference > gpurun_out/f3_bench_reference.json 2> gpurun_out/f3_bench_reference.err
echo "bench reference rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_spconv_umma_dg -c 27 \
  --csv --page raw --log-file gpurun_out/f3_ncu_full_dg.csv python bench.py --steps 1 --warmup 3 --profile --no-secondary --no-cpu-baseline \
  > gpurun_out/f3_ncu.log 2>&1
echo "ncu full rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 12000 --csv \
  --log-file gpurun_out/f3_launches_distill_lidar.csv python bench.py --steps 1 --warmup 3 --profile --no-secondary --no-cpu-baseline \
  > gpurun_out/f3_prof.log 2>&1
echo "ncu launches rc=$?"
python - <<'PY'
import json
for f in ("f3_bench_default", "f3_bench_reference"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), "e2e", d.get("e2e", {}).get("value"), "cpu", d.get("cpu_baseline", {}).get("value"),
              "parity", d.get("parity"), "clocks", d.get("clocks"))
        for k, v in d.get("secondary", {}).items():
            if isinstance(v, dict): print("   ", k, v.get("value"), v.get("ms_per_step"))
    except Exception as e:
        print(f, "unreadable", e)
PY

#!/bin/bash
# session 3K: per-kernel times of the rulebook sequence (ncu launch list, time only)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/k_rulebook_launches.csv python tools/rulebook_bench.py > gpurun_out/k_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/k_rulebook_launches.csv")))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        start = i
        break
for r in rows[start + 1:]:
    if len(r) >= 15:
        print("%-70s grid %-12s %8.1f us" % (r[4][:70], r[8], float(r[14].replace(",", "")) / 1e3))
PY

#!/bin/bash
# session 3I: rulebook microbench + ncu of the rulebook kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/rulebook_bench.py > gpurun_out/i_rulebook.txt 2>&1; echo "rc=$?"; cat gpurun_out/i_rulebook.txt | tail -14
REPS=1 timeout 600 ncu --set full --clock-control none -k regex:'k_conv_tables|k_conv_mark|k_subm_table|k_hash_build|k_popc|k_conv_out_coords|k_transpose_table' -c 24 -o gpurun_out/i_rulebook_ncu -f python tools/rulebook_bench.py > gpurun_out/i_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/i_ncu.log
ls -la gpurun_out/i_rulebook_ncu.ncu-rep

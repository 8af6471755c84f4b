#!/bin/bash
# session 3A: which side of the conv pipeline binds -- role cycles with the gathers removed (exp1), the MMAs removed (exp2),
# gathers + TMEM stores removed (exp3)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in probe exp1 exp2 exp3; do
  UD_B200_LIB=$PWD/cvpr2023-unidistill_b200/variants/lib_$v.so timeout 300 python tools/gpu_dg_probe.py > gpurun_out/dg_$v.txt 2>&1
  echo "== $v"; tail -22 gpurun_out/dg_$v.txt | cut -c1-210
done

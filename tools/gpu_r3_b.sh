#!/bin/bash
# session 3B: serialised MMA completion timing (exp4)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in ${VARIANTS:-exp4}; do
  UD_B200_LIB=$PWD/cvpr2023-unidistill_b200/variants/lib_$v.so timeout 300 python tools/gpu_dg_probe.py > gpurun_out/dg_$v.txt 2>&1
  echo "== $v"; tail -22 gpurun_out/dg_$v.txt | cut -c1-210
done

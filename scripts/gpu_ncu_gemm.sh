#!/bin/bash
# usage: bash scripts/gpu_ncu_gemm.sh [variants...]   (default: gelu dgelu)
mkdir -p gpurun_out
vs="${@:-gelu dgelu}"
for v in $vs; do
  python scripts/gpu_gemm_one.py $v > gpurun_out/time_gemm_$v.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 4 -c 1 -f -o gpurun_out/gemm_$v python scripts/gpu_gemm_one.py $v > gpurun_out/ncu_gemm_$v.log 2>&1
done

#!/bin/bash
mkdir -p gpurun_out
for v in gelu dgelu; do
  ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 4 -c 1 -f -o gpurun_out/gemm_$v python scripts/gpu_gemm_one.py $v > gpurun_out/ncu_gemm_$v.log 2>&1
done

#!/bin/bash
# per-shape GEMM times of the headline step with and without the tail split
mkdir -p gpurun_out
for t in 0 1; do
  PIXPARSE_B200_GEMM_TAIL_SPLIT=$t timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras --no-gpu-reference \
     --profile-all gpurun_out/r02tail_prof$t.txt > gpurun_out/r02tail_prof$t.json 2> gpurun_out/r02tail_prof$t.err
  echo "== tail=$t"; grep -E "N=768 K=(3072|2304)|epi=2 M=32288 N=768 K=768|b200_gemm_bf16$" gpurun_out/r02tail_prof$t.txt | head -12
done

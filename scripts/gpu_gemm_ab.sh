#!/bin/bash
# time the train-step GEMM variants standalone with each library given (paths relative to pixparse_b200/csrc)
mkdir -p gpurun_out
for rep in 1 2; do
  for l in "$@"; do
    echo "== $l"
    for v in gelu dgelu resid store wgrad; do PIXPARSE_B200_LIB=pixparse_b200/csrc/$l python scripts/gpu_gemm_one.py $v; done
  done
done > gpurun_out/gemm_ab.log 2>&1

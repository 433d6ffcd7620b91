#!/bin/bash
# time the train-step GEMM variants standalone: CTA pairs (default) vs SINGLE_CTA=1, or several libraries
# usage: bash scripts/gpu_gemm_ab.sh [lib.so ...]   (paths relative to pixparse_b200/csrc; default: the in-tree build)
mkdir -p gpurun_out
libs="${@:-libpixparse_b200.so}"
for rep in 1; do
  for l in $libs; do
    for mode in ${MODES:-auto allpair single}; do
      echo "== $l $mode"
      for v in ${VARIANTS:-gelu dgelu resid store wgrad wgrad_qkv wgrad_proj}; do
        if [ $mode = single ]; then export SINGLE_CTA=1; elif [ $mode = allpair ]; then export SINGLE_CTA=-1; else unset SINGLE_CTA; fi
        PIXPARSE_B200_LIB=pixparse_b200/csrc/$l timeout 120 python scripts/gpu_gemm_one.py $v
      done
    done
  done
done > gpurun_out/gemm_ab.log 2>&1

#!/bin/bash
# A/B of two builds of the library on the same box: bash scripts/gpu_ab.sh libA.so libB.so  (paths relative to pixparse_b200/csrc)
mkdir -p gpurun_out
for rep in 1 2; do
  for l in "$@"; do
    PIXPARSE_B200_LIB=pixparse_b200/csrc/$l timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_${l%.so}_$rep.json 2> gpurun_out/ab_${l%.so}_$rep.err
  done
done

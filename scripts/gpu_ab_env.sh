#!/bin/bash
# A/B of an environment switch on the same box: bash scripts/gpu_ab_env.sh VAR=value   (runs bench.py without / with it, twice)
mkdir -p gpurun_out
for rep in 1 2; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/abenv_off_$rep.json 2> gpurun_out/abenv_off_$rep.err
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/abenv_on_$rep.json 2> gpurun_out/abenv_on_$rep.err
done

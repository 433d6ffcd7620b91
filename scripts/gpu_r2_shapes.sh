#!/bin/bash
# headline bench with the per-(M, N, K) in-step GEMM / attention table
mkdir -p gpurun_out
TAG=${TAG:-r02c}
PIXPARSE_B200_PROFILE_SHAPES=1 timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline --no-extras --no-gpu-reference \
   --profile-all gpurun_out/${TAG}_step_profile.txt > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit=$?" >> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_step_profile.txt

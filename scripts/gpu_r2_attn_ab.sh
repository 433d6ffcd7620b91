#!/bin/bash
# attention A/B on one box: parity tests with the new library, then the kernel timings for each library given
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py tests/test_decode_gpu.py -q -k "attention or attn or decode or mask or drop" 2>&1 | tail -8
for l in "$@"; do
  echo "=== $l"
  PIXPARSE_B200_LIB=pixparse_b200/csrc/$l ITERS=20 python scripts/gpu_attn_profile.py 2>&1 | tee gpurun_out/r02_attn_${l%.so}.log
  PIXPARSE_B200_LIB=pixparse_b200/csrc/$l ITERS=20 NODROP=1 python scripts/gpu_attn_profile.py 2>&1 | grep dec | sed 's/^/nodrop /' | tee -a gpurun_out/r02_attn_${l%.so}.log
done

#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_decode_fuse.log
: > $L
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_kernels_gpu.py -m gpu -q --maxfail=12 -k "decode" 2>&1 | tail -15 >> $L
for f in 3 0 1 2 3 0; do
  echo "PIXPARSE_B200_DECODE_FUSE=$f" >> $L
  PIXPARSE_B200_DECODE_FUSE=$f python scripts/gpu_decode_profile.py 512 2>&1 | grep "tokens x" >> $L
done
cat $L

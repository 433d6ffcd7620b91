#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_decode_pdl.log
: > $L
for pdl in 1 0 1 0; do
  echo "PIXPARSE_B200_DECODE_PDL=$pdl" >> $L
  PIXPARSE_B200_DECODE_PDL=$pdl python scripts/gpu_decode_profile.py 512 2>&1 | grep "tokens x" >> $L
done
cat $L
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_kernels_gpu.py -m gpu -q --maxfail=12 -k "decode" 2>&1 | tail -5

#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -q -k "wgrad_gemm_fused or gemm" 2>&1 | tail -3
for l in libab_biasg_lds128.so libpixparse_b200.so; do
  PIXPARSE_B200_LIB=pixparse_b200/csrc/$l python scripts/gpu_wgrad_biasgrad.py 2>&1 | tee gpurun_out/r02_biasg_${l%.so}.log
done

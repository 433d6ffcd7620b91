#!/bin/bash
mkdir -p gpurun_out
TAG=${TAG:-r02_decode}
python scripts/gpu_decode_profile.py 64 > gpurun_out/${TAG}_time.log 2>&1
python scripts/gpu_decode_profile.py 512 >> gpurun_out/${TAG}_time.log 2>&1
cat gpurun_out/${TAG}_time.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'decode_|layernorm' --csv \
    --log-file gpurun_out/${TAG}_launches.csv python scripts/gpu_decode_profile.py 32 > gpurun_out/${TAG}_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/${TAG}_launches.csv 36 | tee gpurun_out/${TAG}_launches_summary.txt
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_kernels_gpu.py -m gpu -q --maxfail=12 -k "decode" -s 2>&1 | tail -30

#!/bin/bash
mkdir -p gpurun_out
TAG=${TAG:-r02_decode}
python scripts/gpu_decode_profile.py 64 > gpurun_out/${TAG}_time.log 2>&1
python scripts/gpu_decode_profile.py 512 >> gpurun_out/${TAG}_time.log 2>&1
cat gpurun_out/${TAG}_time.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'decode_|layernorm' --csv \
    --log-file gpurun_out/${TAG}_launches.csv python scripts/gpu_decode_profile.py 32 > gpurun_out/${TAG}_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/${TAG}_launches.csv 36 | tee gpurun_out/${TAG}_launches_summary.txt

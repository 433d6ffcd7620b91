#!/bin/bash
# final evidence of the round: the driver's sequence (GPU suite, smoke, default bench line of both arms) and the ncu launch list
# of the same bench command (cold-cache, serialised: shares, not absolutes)
TAG=r02f SKIP_REF= bash scripts/gpu_r2_verify.sh
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02f_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-gpu-reference > gpurun_out/r02f_ncu_bench.log 2>&1
echo "ncu exit=$?"
python scripts/summarize_launches.py gpurun_out/r02f_launches.csv 3 > gpurun_out/r02f_launches_summary.txt 2>&1
head -12 gpurun_out/r02f_launches_summary.txt

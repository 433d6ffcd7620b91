#!/bin/bash
# per-launch device times of one bench run (cold-cache, serialised: compare SHARES, not absolutes)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "exit=$?" >> gpurun_out/ncu_bench.log

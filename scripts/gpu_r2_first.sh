#!/bin/bash
# round 2, first GPU call: test-suite sanity, library attention baselines, full bench line (extras + stock-torch reference)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r02_pytest_first.log
OUT=gpurun_out/r02_sdpa_baseline.json timeout 600 python scripts/gpu_sdpa_baseline.py > gpurun_out/r02_sdpa_baseline.log 2>&1
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_first.json 2> gpurun_out/r02_bench_first.err
echo "bench exit=$?" >> gpurun_out/r02_bench_first.err
tail -5 gpurun_out/r02_pytest_first.log
tail -3 gpurun_out/r02_bench_first.err

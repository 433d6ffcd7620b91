#!/bin/bash
# round-end evidence: ncu --set full on the dominant GEMM variants + per-launch list of one bench run
mkdir -p gpurun_out
for v in wgrad store; do
  ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 4 -c 1 -f -o gpurun_out/gemm_$v python scripts/gpu_gemm_one.py $v > gpurun_out/ncu_gemm_$v.log 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_final.log 2>&1
echo "exit=$?" >> gpurun_out/ncu_bench_final.log

#!/bin/bash
# ncu --set full of one launch of each GEMM variant at its encoder train-step shape (round-2 kernel) -> raw metric CSVs
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__m_xbar2l1tex_read_bytes.sum,launch__registers_per_thread"
for v in gelu dgelu store wgrad resid resid3072; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 4 -c 1 -f -o gpurun_out/r02_gemm_$v python scripts/gpu_gemm_one.py $v > gpurun_out/r02_ncu_gemm_$v.log 2>&1
  ncu -i gpurun_out/r02_gemm_$v.ncu-rep --page raw --csv --metrics $M > gpurun_out/r02_ncu_gemm_$v.csv 2>> gpurun_out/r02_ncu_gemm_$v.log
  tail -1 gpurun_out/r02_ncu_gemm_$v.csv | cut -c1-400
done

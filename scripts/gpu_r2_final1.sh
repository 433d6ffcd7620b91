#!/bin/bash
# final single-GPU evidence: full GPU suite, smoke, default bench line, ncu of the HBM-bound kernels (new CE kernel included)
mkdir -p gpurun_out
bash scripts/gpu_r2_full.sh
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,launch__registers_per_thread"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'adamw_kernel|sumsq_partial|ce_fwd_bwd|layernorm_fwd_kernel|layernorm_bwd_kernel|preprocess|reduce_shards|decode_linear|decode_attention' \
    -f -o gpurun_out/r02_hbm python scripts/gpu_ncu_hbm.py > gpurun_out/r02_ncu_hbm.log 2>&1
ncu -i gpurun_out/r02_hbm.ncu-rep --page raw --csv --metrics $M > gpurun_out/r02_ncu_hbm_raw.csv 2>> gpurun_out/r02_ncu_hbm.log
tail -2 gpurun_out/r02_ncu_hbm.log

#!/bin/bash
# ncu --set full: (1) every HBM-bound kernel once at bench shapes, (2) the attention kernels at the encoder shape,
# (3) per-launch durations of one whole step (launch list). Summaries are extracted here into gpurun_out/*.txt
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,launch__registers_per_thread,launch__grid_size,launch__block_size"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'adamw_kernel|sumsq_partial|ce_fwd_bwd|layernorm_fwd_kernel|layernorm_bwd_kernel|preprocess|reduce_shards|decode_linear|decode_attention' \
    -f -o gpurun_out/r02_hbm python scripts/gpu_ncu_hbm.py > gpurun_out/r02_ncu_hbm.log 2>&1
ncu -i gpurun_out/r02_hbm.ncu-rep --page raw --csv --metrics $M > gpurun_out/r02_ncu_hbm_raw.csv 2>> gpurun_out/r02_ncu_hbm.log
ONLY=encoder ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attention_fwd_kernel|attention_bwd_kernel' -s 4 -c 2 -f -o gpurun_out/r02_attn python scripts/gpu_attn_profile.py > gpurun_out/r02_ncu_attn.log 2>&1
ncu -i gpurun_out/r02_attn.ncu-rep --page raw --csv --metrics $M,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_barrier.ratio,smsp__average_warp_latency_issue_stalled_wait.ratio,smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio,smsp__average_warp_latency_issue_stalled_mio_throttle.ratio,smsp__average_warp_latency_issue_stalled_sleeping.ratio,smsp__average_warp_latency_issue_stalled_not_selected.ratio > gpurun_out/r02_ncu_attn_raw.csv 2>> gpurun_out/r02_ncu_attn.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --no-gpu-reference > gpurun_out/r02_ncu_bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches_summary.txt 2>&1
tail -3 gpurun_out/r02_ncu_hbm.log; tail -3 gpurun_out/r02_ncu_attn.log; head -30 gpurun_out/r02_launches_summary.txt
ls -la gpurun_out/r02_hbm.ncu-rep gpurun_out/r02_attn.ncu-rep

#!/bin/bash
# ncu --set full of the attention kernels at the encoder shape (ONLY=encoder) -> gpurun_out/attn_{fwd,bwd}.ncu-rep
mkdir -p gpurun_out
ONLY=encoder python scripts/gpu_attn_profile.py > gpurun_out/attn_time.log 2>&1
ONLY=encoder ITERS=1 ncu --set full --clock-control none --import-source on -k regex:attention_bwd_kernel -s 1 -c 1 -f -o gpurun_out/attn_bwd python scripts/gpu_attn_profile.py > gpurun_out/ncu_attn.log 2>&1
ONLY=encoder ITERS=1 ncu --set full --clock-control none --import-source on -k regex:attention_fwd_kernel -s 1 -c 1 -f -o gpurun_out/attn_fwd python scripts/gpu_attn_profile.py >> gpurun_out/ncu_attn.log 2>&1
echo "exit=$?" >> gpurun_out/ncu_attn.log

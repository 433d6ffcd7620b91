#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/gpu_p2p_probe.py > gpurun_out/p2p_probe_$N.log 2>&1
echo "exit=$?" >> gpurun_out/p2p_probe_$N.log
grep -v "^W\|Warning\|warn" gpurun_out/p2p_probe_$N.log | tail -30

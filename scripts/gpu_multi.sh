#!/bin/bash
# usage: gpu_multi.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/ddp_check.py > gpurun_out/ddp_check_$N.log 2>&1
echo "exit=$?" >> gpurun_out/ddp_check_$N.log
MODEL=cruller_base B=2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/ddp_check.py >> gpurun_out/ddp_check_$N.log 2>&1
echo "exit=$?" >> gpurun_out/ddp_check_$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err
echo "exit=$?" >> gpurun_out/bench_$N.err

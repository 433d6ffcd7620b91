#!/bin/bash
# usage: gpu_multi_bench.sh N [ENV=VAL ...]   (bench only, under gpurun --gpus N); result in gpurun_out/bench_N_<tag>.json
N=${1:-2}; shift
tag=$(echo "$*" | tr ' =' '__'); tag=${tag:-default}
mkdir -p gpurun_out
env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}_$tag.json 2> gpurun_out/bench_${N}_$tag.err
echo "exit=$?" >> gpurun_out/bench_${N}_$tag.err

#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
for mode in p2p nccl; do
PIXPARSE_B200_REDUCER=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-gpu-reference > gpurun_out/r02_exposed_${N}_$mode.json 2> gpurun_out/r02_exposed_${N}_$mode.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_exposed_${N}_$mode.json").read().strip().splitlines()[-1])
print("N=$N $mode: ms/step", round(d["ms_per_step"], 3), "exposed wait after backward (rank 0)", d.get("exchange_exposed_ms"), d.get("per_rank_ms"))
PY
done

#!/bin/bash
# usage: gpu_r2_p2p_ab.sh N "modes" reps
N=${1:-2}; MODES=${2:-"none p2p nccl"}; REPS=${3:-2}
mkdir -p gpurun_out
L=gpurun_out/r02_p2p_ab_$N.log
: > $L
for rep in $(seq 1 $REPS); do
for mode in $MODES; do
  PIXPARSE_B200_REDUCER=$mode python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-gpu-reference > gpurun_out/r02_p2p_bench_${N}_${mode}_$rep.json 2> gpurun_out/r02_p2p_bench_${N}_${mode}_$rep.err
  python - <<PY >> $L
import json
try:
    d = json.loads(open("gpurun_out/r02_p2p_bench_${N}_${mode}_$rep.json").read().strip().splitlines()[-1])
    print("bench N=$N reducer=$mode rep=$rep: pages/s", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "sm_mhz", d["clocks"]["sm_mhz"], d.get("per_rank_ms"))
except Exception as e:
    print("bench N=$N reducer=$mode: no line:", e)
PY
done
done
cat $L

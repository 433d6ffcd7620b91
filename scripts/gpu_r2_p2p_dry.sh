#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
L=gpurun_out/r02_p2p_dry_$N.log
: > $L
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
for rep in 1 2; do
for cfg in "p2p 1" "p2p 0" "none 0"; do
  set -- $cfg
  PIXPARSE_B200_REDUCER=$1 PIXPARSE_B200_P2P_DRY=$2 run 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-gpu-reference > gpurun_out/r02_pd_${N}_$1_$2_$rep.json 2> gpurun_out/r02_pd_${N}_$1_$2_$rep.err
  python - <<PY >> $L
import json
try:
    d = json.loads(open("gpurun_out/r02_pd_${N}_$1_$2_$rep.json").read().strip().splitlines()[-1])
    print("bench N=$N reducer=$1 dry=$2 rep=$rep: ms/step", round(d["ms_per_step"], 3), "sm_mhz", d["clocks"]["sm_mhz"], d.get("per_rank_ms"))
except Exception as e:
    print("bench N=$N reducer=$1 dry=$2: no line:", e)
PY
done
done
cat $L

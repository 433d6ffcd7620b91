#!/bin/bash
# usage: gpu_r2_p2p_barrier.sh N : ddp_check with the memop barrier, then bench A/B: p2p+memop barrier / p2p+nccl barrier / none
N=${1:-2}
mkdir -p gpurun_out
L=gpurun_out/r02_p2p_barrier_$N.log
: > $L
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
PIXPARSE_B200_REDUCER=p2p run 29511 scripts/ddp_check.py 2>&1 | grep "ddp_check\|identical\|Error\|error" >> $L
MODEL=cruller_base B=2 PIXPARSE_B200_REDUCER=p2p run 29512 scripts/ddp_check.py 2>&1 | grep "ddp_check\|identical\|Error\|error" >> $L
for rep in 1 2; do
for cfg in "p2p memop" "p2p nccl" "none memop"; do
  set -- $cfg
  PIXPARSE_B200_REDUCER=$1 PIXPARSE_B200_P2P_BARRIER=$2 run 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-gpu-reference > gpurun_out/r02_pb_${N}_$1_$2_$rep.json 2> gpurun_out/r02_pb_${N}_$1_$2_$rep.err
  python - <<PY >> $L
import json
try:
    d = json.loads(open("gpurun_out/r02_pb_${N}_$1_$2_$rep.json").read().strip().splitlines()[-1])
    print("bench N=$N reducer=$1 barrier=$2 rep=$rep: pages/s", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "sm_mhz", d["clocks"]["sm_mhz"], d.get("per_rank_ms"))
except Exception as e:
    print("bench N=$N reducer=$1 barrier=$2: no line:", e)
PY
done
done
cat $L
grep -i "error\|Traceback" gpurun_out/r02_pb_${N}_p2p_memop_1.err | head -5

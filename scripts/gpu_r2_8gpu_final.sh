#!/bin/bash
# under gpurun --gpus N: final multi-GPU evidence with the copy-engine exchange + stream-memop barriers
N=${1:-8}
mkdir -p gpurun_out
L=gpurun_out/r02_${N}gpu_final.log
: > $L
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
PIXPARSE_B200_REDUCER=p2p run 29511 scripts/ddp_check.py 2>&1 | grep "ddp_check\|identical\|Error\|error" >> $L
line() {
python - <<PY >> $L
import json
try:
    d = json.loads(open("$1").read().strip().splitlines()[-1])
    print("$2: pages/s", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "mfu_burst", round(d["mfu"]["vs_measured_burst"], 4), "sm_mhz", d["clocks"]["sm_mhz"], d.get("per_rank_ms"))
except Exception as e:
    print("$2: no line:", e)
PY
}
for cfg in "p2p memop" "p2p nccl" "none memop" "nccl memop"; do
  set -- $cfg
  PIXPARSE_B200_REDUCER=$1 PIXPARSE_B200_P2P_BARRIER=$2 run 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-gpu-reference > gpurun_out/r02f_base_${N}gpu_$1_$2.json 2> gpurun_out/r02f_base_${N}gpu_$1_$2.err
  line gpurun_out/r02f_base_${N}gpu_$1_$2.json "base N=$N reducer=$1 barrier=$2"
done
PIXPARSE_B200_REDUCER=p2p run 29514 bench.py --gpus $N --config large --steps 6 --warmup 3 --no-cpu-baseline --no-extras --no-gpu-reference > gpurun_out/r02f_large_${N}gpu.json 2> gpurun_out/r02f_large_${N}gpu.err
line gpurun_out/r02f_large_${N}gpu.json "large N=$N reducer=p2p barrier=memop"
cat $L
grep -i "error\|Traceback\|falls back" gpurun_out/r02f_base_${N}gpu_p2p_memop.err | head -5

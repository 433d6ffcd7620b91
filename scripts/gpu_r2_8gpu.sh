#!/bin/bash
# under gpurun --gpus 8: correctness of the copy-engine exchange at 8 ranks, headline bench (p2p vs nccl), cruller_large (configs[2])
N=${1:-8}
mkdir -p gpurun_out
L=gpurun_out/r02_${N}gpu.log
: > $L
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
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
for mode in p2p nccl; do
  PIXPARSE_B200_REDUCER=$mode run 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-gpu-reference > gpurun_out/r02_bench_base_${N}gpu_$mode.json 2> gpurun_out/r02_bench_base_${N}gpu_$mode.err
  line gpurun_out/r02_bench_base_${N}gpu_$mode.json "base N=$N reducer=$mode"
done
for mode in ${LARGE_MODES:-p2p}; do
  PIXPARSE_B200_REDUCER=$mode run 29514 bench.py --gpus $N --config large --steps 6 --warmup 3 --no-cpu-baseline --no-extras --no-gpu-reference > gpurun_out/r02_bench_large_${N}gpu_$mode.json 2> gpurun_out/r02_bench_large_${N}gpu_$mode.err
  line gpurun_out/r02_bench_large_${N}gpu_$mode.json "large N=$N reducer=$mode"
done
cat $L
grep -v "Warning\|warn\|return func" gpurun_out/r02_bench_large_${N}gpu_p2p.err | tail -5

#!/bin/bash
# usage: gpu_r2_p2p.sh N  (under gpurun --gpus N): correctness of the copy-engine gradient exchange, then bench A/B vs NCCL
N=${1:-2}
mkdir -p gpurun_out
L=gpurun_out/r02_p2p_$N.log
: > $L
for mode in p2p nccl; do
  PIXPARSE_B200_REDUCER=$mode python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/ddp_check.py 2>&1 | grep "ddp_check\|identical\|Error\|error" >> $L
  MODEL=cruller_base B=2 PIXPARSE_B200_REDUCER=$mode python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/ddp_check.py 2>&1 | grep "ddp_check\|identical\|Error\|error" >> $L
done
for rep in 1 2; do
for mode in p2p nccl; do
  PIXPARSE_B200_REDUCER=$mode python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-gpu-reference > gpurun_out/r02_p2p_bench_${N}_${mode}_$rep.json 2> gpurun_out/r02_p2p_bench_${N}_${mode}_$rep.err
  python - <<PY >> $L
import json
try:
    d = json.loads(open("gpurun_out/r02_p2p_bench_${N}_${mode}_$rep.json").read().strip().splitlines()[-1])
    print("bench N=$N reducer=$mode rep=$rep: pages/s", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "sm_mhz", d["clocks"]["sm_mhz"])
except Exception as e:
    print("bench N=$N reducer=$mode: no line:", e)
PY
done
done
if [ "$WITH_N1" = "1" ]; then
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-gpu-reference > gpurun_out/r02_p2p_bench_1.json 2> gpurun_out/r02_p2p_bench_1.err
python - <<PY >> $L
import json
d = json.loads(open("gpurun_out/r02_p2p_bench_1.json").read().strip().splitlines()[-1])
print("bench N=1: pages/s", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3))
PY
fi
cat $L
tail -5 gpurun_out/r02_p2p_bench_${N}_p2p_1.err

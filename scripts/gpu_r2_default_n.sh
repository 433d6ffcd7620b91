#!/bin/bash
# the driver's own scaling command at N GPUs (default flags: extras included), then the reference arm the same way
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_default_$N.json 2> gpurun_out/r02_default_$N.err
echo "exit=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_default_$N.json").read().strip().splitlines()[-1])
print("pages/s", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), d.get("per_rank_ms"))
for k, v in d.get("extra_configs", {}).items():
    print(k, {a: b for a, b in v.items() if a in ("value", "ms_per_step", "error", "n_gpus")})
PY
grep -i "error\|Traceback\|fall" gpurun_out/r02_default_$N.err | head -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | tail -c 300

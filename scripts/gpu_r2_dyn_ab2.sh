#!/bin/bash
# dynamic tiles + side-stream weight gradients (encoder and decoder, events right behind the producing kernel): parity, then A/B
mkdir -p gpurun_out
TAG=${TAG:-r02dyn2}
timeout 900 python -m pytest tests -m gpu -q -x -k "side_stream or gemm or golden or tiny_model" 2>&1 | tail -6 > gpurun_out/${TAG}_pytest.log
echo "pytest exit=${PIPESTATUS[0]}" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
if grep -q "exit=0" gpurun_out/${TAG}_pytest.log; then
for rep in 1 2 3; do
  for cfg in "0 0" "1 1"; do
    set -- $cfg
    PIXPARSE_B200_DYN_SCHED=$1 PIXPARSE_B200_SIDE_WGRAD=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-gpu-reference \
      > gpurun_out/${TAG}_dyn$1_side$2_$rep.json 2> gpurun_out/${TAG}_dyn$1_side$2_$rep.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_dyn$1_side$2_$rep.json").read().strip().splitlines()[-1])
    print("dyn=$1 side=$2 rep=$rep pages/s", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "sm_mhz", d["clocks"]["sm_mhz"], "loss", d.get("loss"))
except Exception as e:
    print("dyn=$1 side=$2 rep=$rep: no bench line:", e)
PY
  done
done
fi

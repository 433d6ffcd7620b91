#!/bin/bash
# same-box A/B of two source trees: the repo root ("new") against a checkout of an earlier commit under _old/ ("old")
mkdir -p gpurun_out
ROOT=$(pwd)
for rep in 1 2 3; do
  for which in old new; do
    if [ $which = old ]; then cd $ROOT/_old; else cd $ROOT; fi
    timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-gpu-reference > $ROOT/gpurun_out/r02ab_${which}_$rep.json 2> $ROOT/gpurun_out/r02ab_${which}_$rep.err
    cd $ROOT
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02ab_${which}_$rep.json").read().strip().splitlines()[-1])
    print("$which rep=$rep pages/s", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "gemm_tf", round(d["roofline"]["achieved"], 1), "sm_mhz", d["clocks"]["sm_mhz"])
except Exception as e:
    print("$which rep=$rep: no bench line:", e)
PY
  done
done

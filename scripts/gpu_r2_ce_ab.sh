#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_ce_ab.log
: > $L
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_golden.py -m gpu -q --maxfail=12 -k "cross_entropy or parity or golden" 2>&1 | tail -5 >> $L
for m in 1 0 1 0; do
  PIXPARSE_B200_CE_PIPE=$m python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-gpu-reference > gpurun_out/r02_ce_$m.json 2> gpurun_out/r02_ce_$m.err
  python - <<PY >> $L
import json
d = json.loads(open("gpurun_out/r02_ce_$m.json").read().strip().splitlines()[-1])
k = d["roofline_hbm"]["kernels"]["ce_fwd_bwd_kernel"]
print("CE_PIPE=$m: step", round(d["ms_per_step"], 3), "ms; ce alone", round(k["ms"], 4), "ms", round(k["achieved"]), "GB/s frac", round(k["frac"], 3), "loss", d["loss"])
PY
done
cat $L

#!/bin/bash
# GEMM tail split: kernel + model parity with it on (the default), then the headline bench with and without it (same box, alternating)
mkdir -p gpurun_out
TAG=${TAG:-r02tail}
timeout 900 python -m pytest tests -m gpu -q -x -k "${KEXPR:-gemm or golden or model or dropout}" 2>&1 | tail -40 > gpurun_out/${TAG}_pytest.log
echo "pytest exit=${PIPESTATUS[0]}" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
if grep -q "exit=0" gpurun_out/${TAG}_pytest.log; then
for rep in 1 2 3; do
  for t in 0 1; do
    PIXPARSE_B200_GEMM_TAIL_SPLIT=$t timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-gpu-reference \
      ${PROFILE:+--profile-all gpurun_out/${TAG}_tail${t}_${rep}_profile.txt} > gpurun_out/${TAG}_tail${t}_$rep.json 2> gpurun_out/${TAG}_tail${t}_$rep.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_tail${t}_$rep.json").read().strip().splitlines()[-1])
    print("tail=$t rep=$rep pages/s", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "sm_mhz", d["clocks"]["sm_mhz"], "loss", d.get("loss"))
except Exception as e:
    print("tail=$t rep=$rep: no bench line:", e)
PY
  done
done
else
  head -60 gpurun_out/${TAG}_pytest.log
fi

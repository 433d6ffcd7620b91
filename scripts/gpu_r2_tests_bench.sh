#!/bin/bash
# GPU test-suite (all failures, not just the first) + headline bench with the per-entry-point profile
mkdir -p gpurun_out
TAG=${TAG:-r02}
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -k "${KEXPR:-test}" 2>&1 | tail -60 > gpurun_out/${TAG}_pytest.log
echo "pytest exit=$?" >> gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline ${BENCH_ARGS:---no-extras --no-gpu-reference} \
   --profile-all gpurun_out/${TAG}_step_profile.txt > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit=$?" >> gpurun_out/${TAG}_bench.err
tail -25 gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("pages/s", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "mfu_burst", round(d["mfu"]["vs_measured_burst"], 4))
except Exception as e:
    print("no bench line:", e)
PY
head -20 gpurun_out/${TAG}_step_profile.txt

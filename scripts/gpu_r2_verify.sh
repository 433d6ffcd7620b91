#!/bin/bash
# what the driver runs at round end, on one box: GPU suite, smoke(), the default bench line (both arms), with wall times
mkdir -p gpurun_out
TAG=${TAG:-r02v}
t0=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/${TAG}_pytest.log
echo "pytest exit=${PIPESTATUS[0]} wall=$(( $(date +%s) - t0 ))s" >> gpurun_out/${TAG}_pytest.log
t0=$(date +%s)
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke exit=$? wall=$(( $(date +%s) - t0 ))s" >> gpurun_out/${TAG}_smoke.log
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit=$? wall=$(( $(date +%s) - t0 ))s" >> gpurun_out/${TAG}_bench.err
if [ -z "$SKIP_REF" ]; then
t0=$(date +%s)
timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
echo "ref exit=$? wall=$(( $(date +%s) - t0 ))s" >> gpurun_out/${TAG}_bench_ref.err
fi
tail -3 gpurun_out/${TAG}_pytest.log; tail -2 gpurun_out/${TAG}_smoke.log; tail -2 gpurun_out/${TAG}_bench.err; tail -1 gpurun_out/${TAG}_bench_ref.err 2>/dev/null
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("pages/s", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "mfu_burst", round(d["mfu"]["vs_measured_burst"], 4), "clocks", d.get("clocks"))
print("roofline", d["roofline"]["frac"], "gpu_ref", (d.get("gpu_reference") or {}).get("value"))
PY

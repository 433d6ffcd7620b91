#!/bin/bash
# full GPU test-suite + the default bench line (all configs, stock-torch arm, cpu baseline) + smoke
mkdir -p gpurun_out
TAG=${TAG:-r02_full}
timeout 2400 python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -40 > gpurun_out/${TAG}_pytest.log
echo "pytest exit=$?" >> gpurun_out/${TAG}_pytest.log
tail -8 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 1500 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit=$?" >> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("pages/s", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "mfu_burst", round(d["mfu"]["vs_measured_burst"], 4), "clocks", d["clocks"])
for k, v in d.get("extra_configs", {}).items():
    print(k, round(v.get("value", 0), 1), v.get("unit"), v.get("mfu"))
print("gpu_reference", {k: v for k, v in d.get("gpu_reference", {}).items() if k != "what"})
print("cpu_baseline", d.get("cpu_baseline"))
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
tail -c 600 gpurun_out/${TAG}_bench_reference.json

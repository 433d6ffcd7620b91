#!/bin/bash
# programmatic dependent launch on/off on the same box: full GPU test-suite with it on, then the headline bench twice each way
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 2>&1 | tail -15 > gpurun_out/r02_pdl_pytest.log
tail -4 gpurun_out/r02_pdl_pytest.log
for rep in 1 2; do
  for v in 0 1; do
    PIXPARSE_B200_PDL=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-gpu-reference \
       > gpurun_out/r02_pdl${v}_$rep.json 2> gpurun_out/r02_pdl${v}_$rep.err
    python -c "
import json; d=json.loads(open('gpurun_out/r02_pdl${v}_$rep.json').read().strip().splitlines()[-1])
print('PDL=$v rep $rep: pages/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'gemm', round(d['roofline']['gemm_ms_per_step'],2), 'attn', round(d['roofline']['attention_ms_per_step'],2), d['clocks']['sm_mhz'])"
  done
done

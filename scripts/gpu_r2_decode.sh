#!/bin/bash
# decode kernels: parity tests, then the configs[4] bench line (cruller_large_6layers greedy decode, 16 pages)
mkdir -p gpurun_out
TAG=${TAG:-r02_decode}
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_kernels_gpu.py -m gpu -q --maxfail=12 -k "${KEXPR:-decode}" -s 2>&1 | tail -80 > gpurun_out/${TAG}_pytest.log
echo "pytest exit=$?" >> gpurun_out/${TAG}_pytest.log
tail -40 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --config eval_ocr --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit=$?" >> gpurun_out/${TAG}_bench.err
tail -5 gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json

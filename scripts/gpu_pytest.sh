#!/bin/bash
# run the GPU test-suite on the box; keep the log in gpurun_out/
mkdir -p gpurun_out
timeout ${1:-900} python -m pytest tests -m gpu -x -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
echo "exit=$?" >> gpurun_out/pytest_gpu.log

#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_sanitizer.log
: > $L
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool: tests/test_kernels_gpu.py -k 'decode or cross_entropy'" >> $L
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "decode or cross_entropy" 2>&1 | grep -v "^$" | tail -12 >> $L
  echo "exit=$?" >> $L
done
cat $L

#!/bin/bash
# compute-sanitizer over the GEMM tests incl. the opt-in modes (tail split, dynamic tiles through the two-stream model test)
mkdir -p gpurun_out
L=gpurun_out/r02_sanitizer_modes.log
: > $L
echo "== compute-sanitizer --tool memcheck: tests/test_kernels_gpu.py -k gemm" >> $L
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm" 2>&1 | grep -v "^$" | tail -8 >> $L
echo "exit=${PIPESTATUS[0]}" >> $L
echo "== compute-sanitizer --tool memcheck: tests/test_model_gpu.py -k 'side_stream and cruller_test' (dynamic tiles + side-stream weight gradients)" >> $L
timeout 150 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_model_gpu.py -m gpu -q -x -k "side_stream and cruller_test" 2>&1 | grep -v "^$" | tail -8 >> $L
echo "exit=${PIPESTATUS[0]}" >> $L
cat $L

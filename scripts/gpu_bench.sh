#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
timeout ${TMO:-900} python bench.py --steps ${STEPS:-10} --warmup 3 "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "exit=$?" >> gpurun_out/bench.err

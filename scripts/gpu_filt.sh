#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env timeout -s KILL 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?"; }
run tfilt python -m pytest tests/test_gpu_filter.py -q -m gpu --timeout 300 -x
tail -n 2 gpurun_out/tfilt.log
run filtb python scripts/filter_bench.py; cat gpurun_out/filtb.log

#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env timeout -s KILL 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?"; }
run tl4 python -m pytest tests/test_gpu_lstm.py tests/test_gpu_gemm.py tests/test_gpu_step.py -q -m gpu --timeout 300 -x
tail -n 3 gpurun_out/tl4.log
run cfg4 python scripts/bench_cfg4.py; tail -n 1 gpurun_out/cfg4.log
CSN_NO_PDL=1 timeout -s KILL 300 python scripts/bench_cfg4.py 2>&1 | tail -n 1
PB=512 timeout -s KILL 300 python scripts/bench_cfg4.py 2>&1 | tail -n 1

#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env timeout -s KILL 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?"; }
run t14 python -m pytest tests/test_retrieval.py -q -m gpu --timeout 300 -x
tail -n 12 gpurun_out/t14.log
run cfg5 python scripts/bench_cfg5.py; cat gpurun_out/cfg5.log | tail -8

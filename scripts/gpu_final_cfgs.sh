#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env timeout -s KILL 240 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?"; tail -n 5 gpurun_out/$name.log | grep -v "^$" | tail -n 4; }
run cfg3 python scripts/bench_cfg3.py
run cfg4 python scripts/bench_cfg4.py
run cfg5 python scripts/bench_cfg5.py
run filtb python scripts/filter_bench.py

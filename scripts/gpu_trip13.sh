#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env timeout -s KILL 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?"; }
run t13 python -m pytest tests -q -m gpu --timeout 300
tail -n 3 gpurun_out/t13.log
run smoke13 python __graft_entry__.py smoke; tail -n 3 gpurun_out/smoke13.log
run bench13 python bench.py
tail -n 1 gpurun_out/bench13.log | cut -c1-3000
bash scripts/gpu_ncu.sh r01e

#!/bin/bash
# Full validation: every gpu test, smoke, filter/loss micro-benches, default bench.
mkdir -p gpurun_out
run() { name=$1; shift; env timeout -s KILL 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?"; }
run t11 python -m pytest tests -q -m gpu --timeout 300
tail -n 3 gpurun_out/t11.log
run smoke11 python __graft_entry__.py smoke; tail -n 3 gpurun_out/smoke11.log
run filt11 python scripts/filter_bench.py; cat gpurun_out/filt11.log
run loss11 python scripts/loss_bench.py; cat gpurun_out/loss11.log
run bench11 python bench.py
tail -n 2 gpurun_out/bench11.log

#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env timeout -s KILL 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?"; }
run tloss python -m pytest tests/test_gpu_loss.py tests/test_gpu_step.py -q -m gpu --timeout 300 -x
tail -n 5 gpurun_out/tloss.log
run lossb python scripts/loss_bench.py; cat gpurun_out/lossb.log

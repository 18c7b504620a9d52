#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env timeout -s KILL 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?"; }
run tfilt python -m pytest tests/test_gpu_filter.py tests/test_gpu_step.py tests/test_gpu_lstm.py -q -m gpu --timeout 300 -x
tail -n 4 gpurun_out/tfilt.log
run filtb python scripts/filter_bench.py; cat gpurun_out/filtb.log
CSN_FILTER_NO_FAST=1 python scripts/filter_bench.py 2>&1 | head -3
run benchf python bench.py --steps 100 --warmup 10 --no_cpu_baseline
tail -n 1 gpurun_out/benchf.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print(round(d['value']), 'trials/s', round(d['ms_per_step'],4), 'ms; e2e', round(d['e2e']['value']), d['stages_ms'], 'roofline', round(d['roofline']['frac'],4), 'filter', round(d['roofline_filter']['frac'],3), 'loss', round(d['roofline_loss']['frac'],3), 'launches', d['gpu_launches'])
"

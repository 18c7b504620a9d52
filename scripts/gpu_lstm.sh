#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env timeout -s KILL 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?"; }
run tlstm python -m pytest tests/test_gpu_lstm.py tests/test_gpu_step.py -q -m gpu --timeout 300 -x
tail -n 15 gpurun_out/tlstm.log
run benchl python bench.py --steps 100 --warmup 5 --no_cpu_baseline
tail -n 1 gpurun_out/benchl.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print(round(d['value']), 'trials/s', round(d['ms_per_step'],4), 'ms; e2e', round(d['e2e']['value']), d['stages_ms'], d['clocks'], 'roofline', round(d['roofline']['frac'],4), 'filter', round(d['roofline_filter']['frac'],3), 'loss', round(d['roofline_loss']['frac'],3))
"

#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env timeout -s KILL 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?"; }
run t12 python -m pytest tests -q -m gpu --timeout 300 -x
tail -n 6 gpurun_out/t12.log
run loss12 python scripts/loss_bench.py; cat gpurun_out/loss12.log
run bench12 python bench.py --steps 100 --warmup 10 --no_cpu_baseline
tail -n 1 gpurun_out/bench12.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print(round(d['value']), 'trials/s', round(d['ms_per_step'],4), 'ms; e2e', round(d['e2e']['value']), d['stages_ms'], d['clocks'], 'roofline', round(d['roofline']['frac'],4), 'filter', round(d['roofline_filter']['frac'],3), 'loss', round(d['roofline_loss']['frac'],3), 'launches', d['gpu_launches'])
"

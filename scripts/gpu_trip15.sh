#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env timeout -s KILL 240 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?"; }
run t15 python -m pytest tests -q -m gpu --timeout 90
tail -n 3 gpurun_out/t15.log
run smoke15 python __graft_entry__.py smoke; tail -n 3 gpurun_out/smoke15.log
run bench15 python bench.py --no_cpu_baseline
tail -n 1 gpurun_out/bench15.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print(round(d['value']), 'trials/s', round(d['ms_per_step'],4), 'ms; e2e', round(d['e2e']['value']), d['stages_ms'], 'roofline', round(d['roofline']['frac'],4), 'filter', round(d['roofline_filter']['frac'],3), 'loss', round(d['roofline_loss']['frac'],3), 'launches', d['gpu_launches'], d['dp_exchange'])
"

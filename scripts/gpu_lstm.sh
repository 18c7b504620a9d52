#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env timeout -s KILL 150 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?"; }
run tlstm python -m pytest tests/test_gpu_lstm.py tests/test_gpu_step.py -q -m gpu --timeout 60 -x
tail -n 4 gpurun_out/tlstm.log
run prof python scripts/prof_lstm_steps.py; grep -A6 "^backward" gpurun_out/prof.log | grep -v "^periods\|^wait" | head -8
run benchl python bench.py --steps 200 --warmup 10 --no_cpu_baseline
tail -n 1 gpurun_out/benchl.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print(round(d['value']), 'trials/s', round(d['ms_per_step'],4), 'ms; e2e', round(d['e2e']['value']), d['stages_ms'])
"

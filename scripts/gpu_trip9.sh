#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env timeout -s KILL 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?"; }
for nv in 2 4 8; do CSN_LSTM_NV=$nv run prof9_nv$nv python scripts/prof_lstm_steps.py; echo "NV=$nv"; grep median gpurun_out/prof9_nv$nv.log; done
for pb in 128 64; do PB=$pb CSN_LSTM_NV=2 run prof9_b$pb python scripts/prof_lstm_steps.py; echo "B=$pb NV=2"; grep median gpurun_out/prof9_b$pb.log; done
CSN_LSTM_NV=4 run bench9_nv4 python bench.py --steps 20 --warmup 5 --no_cpu_baseline
grep -o '"value": [0-9.]*, "unit": "trials/s", "n_gpus"' gpurun_out/bench9_nv4.log

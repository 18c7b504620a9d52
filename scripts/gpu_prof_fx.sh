#!/bin/bash
mkdir -p gpurun_out
env timeout -s KILL 300 python scripts/prof_lstm_steps.py > gpurun_out/prof_fx.log 2>&1; echo "fx exit $?"
echo "== FX"; grep -A12 "^forward kernel" gpurun_out/prof_fx.log | head -14; grep -A4 "periods" gpurun_out/prof_fx.log | head -3
bash scripts/gpu_lstm.sh

#!/bin/bash
# cluster recurrence with the GEMMs overlapped on the free SMs: parity, layer timings (overlap on / off), cfg4 step
timeout 300 python -m pytest tests/test_gpu_lstm.py tests/test_gpu_determinism.py -x -q -k "256 or 512" 2>&1 | tail -3
for B in 128 512; do
  PH=512 PB=$B PI=128 timeout 120 python scripts/lstm_layer_bench.py 2>&1 | tail -1
  PH=512 PB=$B PI=512 timeout 120 python scripts/lstm_layer_bench.py 2>&1 | tail -1
  CSN_LSTM_NO_OVERLAP=1 PH=512 PB=$B PI=128 timeout 120 python scripts/lstm_layer_bench.py 2>&1 | tail -1
  CSN_LSTM_NO_OVERLAP=1 PH=512 PB=$B PI=512 timeout 120 python scripts/lstm_layer_bench.py 2>&1 | tail -1
done
timeout 200 python scripts/bench_cfg4.py 2>&1 | tail -1
CSN_LSTM_NO_OVERLAP=1 timeout 200 python scripts/bench_cfg4.py 2>&1 | tail -1

#!/bin/bash
# cluster recurrence (H = 256 / 512): parity cases, then layer and step timings with and without it
timeout 300 python -m pytest tests/test_gpu_lstm.py -x -q -k "256 or 512" 2>&1 | tail -8
PH=512 PB=128 PI=128 timeout 120 python scripts/lstm_layer_bench.py 2>&1 | tail -2
PH=512 PB=128 PI=512 timeout 120 python scripts/lstm_layer_bench.py 2>&1 | tail -2
CSN_LSTM_NO_CLUSTER=1 PH=512 PB=128 PI=128 timeout 120 python scripts/lstm_layer_bench.py 2>&1 | tail -2
timeout 200 python scripts/bench_cfg4.py 2>&1 | tail -2

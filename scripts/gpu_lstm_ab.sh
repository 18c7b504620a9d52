#!/bin/bash
CSN_LSTM_SERVERS=1 python -m pytest tests/test_gpu_lstm.py -q -x --timeout 120 2>&1 | tail -2
CSN_LSTM_SERVERS=1 python scripts/lstm_layer_bench.py
python scripts/lstm_layer_bench.py
CSN_LSTM_SERVERS=1 python scripts/prof_lstm_steps.py 2>&1 | grep -B4 -A10 "^forward B"

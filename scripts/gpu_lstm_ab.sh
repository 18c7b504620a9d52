#!/bin/bash
python -m pytest tests/test_gpu_lstm.py tests/test_gpu_determinism.py -q -x --timeout 120 2>&1 | tail -3
for sh in 0.20 0.27 0.33 0.40 1.0; do echo share $sh; CSN_LSTM_CONSUMER_SHARE=$sh CSN_LSTM_NO_SERVERS=1 python scripts/lstm_layer_bench.py; done
CSN_LSTM_NO_SERVERS=1 CSN_LSTM_NO_CONSUMERS=1 python scripts/lstm_layer_bench.py

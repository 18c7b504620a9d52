#!/bin/bash
for b in 0 4 8 16 32; do echo XBURST $b; CSN_LSTM_XBURST=$b python scripts/lstm_layer_bench.py; done
CSN_LSTM_XBURST=8 python -m pytest tests/test_gpu_lstm.py -q -x --timeout 120 2>&1 | tail -2
CSN_LSTM_XBURST=8 python scripts/prof_lstm_steps.py 2>&1 | grep -A10 "^forward B"

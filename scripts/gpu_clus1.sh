#!/bin/bash
# cluster recurrence (H = 256 / 512): parity cases, then layer and step timings: one / two trial groups per cluster, no cluster
timeout 300 python -m pytest tests/test_gpu_lstm.py -x -q -k "256 or 512" 2>&1 | tail -4
CSN_CLUSTER_GROUPS=1 timeout 300 python -m pytest tests/test_gpu_lstm.py -x -q -k "256 or 512" 2>&1 | tail -2
CSN_CLUSTER_GROUPS=2 timeout 300 python -m pytest tests/test_gpu_lstm.py -x -q -k "256 or 512" 2>&1 | tail -2
for B in 128 112 64 256 512; do
  PH=512 PB=$B PI=128 timeout 120 python scripts/lstm_layer_bench.py 2>&1 | tail -1
  CSN_CLUSTER_GROUPS=1 PH=512 PB=$B PI=128 timeout 120 python scripts/lstm_layer_bench.py 2>&1 | tail -1
  CSN_CLUSTER_GROUPS=2 PH=512 PB=$B PI=128 timeout 120 python scripts/lstm_layer_bench.py 2>&1 | tail -1
done
timeout 200 python scripts/bench_cfg4.py 2>&1 | tail -1

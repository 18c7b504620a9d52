#!/bin/bash
# exchange through L2 + multicast: microbenchmark data check, parity, layer / step timings with and without it
for i in 25 27; do timeout 20 scripts/_bin/cluster_xchg_bench $i; done
timeout 300 python -m pytest tests/test_gpu_lstm.py tests/test_gpu_determinism.py -x -q -k "256 or 512" 2>&1 | tail -3
for B in 128 64 512; do
  PH=512 PB=$B PI=128 timeout 120 python scripts/lstm_layer_bench.py 2>&1 | tail -1
  CSN_CLUSTER_NO_MULTICAST=1 PH=512 PB=$B PI=128 timeout 120 python scripts/lstm_layer_bench.py 2>&1 | tail -1
done
timeout 200 python scripts/bench_cfg4.py 2>&1 | tail -1
CSN_CLUSTER_NO_MULTICAST=1 timeout 200 python scripts/bench_cfg4.py 2>&1 | tail -1

#!/bin/bash
# cfg3 with side roles under the per-chain CTA budgets: parity, then step time for several budget splits (teacher, global, local, backward)
timeout 600 python -m pytest tests/test_gpu_multicrop_step.py tests/test_gpu_step.py tests/test_gpu_lstm.py -x -q 2>&1 | tail -2
for b in 64,64,64,64 44,44,60,74 40,40,68,74 38,38,72,74 48,48,52,74; do
  echo "budgets $b"; CTA_BUDGET=$b timeout 200 python scripts/cfg3_step.py 2>&1 | tail -1
done
echo "side roles off under budgets"; CSN_LSTM_BUDGET_SIDE=0 timeout 200 python scripts/cfg3_step.py 2>&1 | tail -1

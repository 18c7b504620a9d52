#!/bin/bash
for b in 48,48,52,74 50,50,48,74 46,46,56,74 52,52,44,74 48,48,52,64 48,48,52,90 56,56,36,74 48,48,52,148; do
  echo "budgets $b"; CTA_BUDGET=$b timeout 200 python scripts/cfg3_step.py 2>&1 | tail -1 | cut -c1-60
done

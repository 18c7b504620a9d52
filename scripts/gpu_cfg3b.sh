#!/bin/bash
mkdir -p gpurun_out
for cfg in "0 0" "1 48" "1 64" "1 40"; do set -- $cfg
  CONCURRENT=$1 CTA_BUDGET=$2 timeout -s KILL 300 python scripts/bench_cfg3.py > gpurun_out/cfg3_$1_$2.log 2>&1; echo "conc=$1 budget=$2 exit $?"; tail -n 2 gpurun_out/cfg3_$1_$2.log | head -1
done

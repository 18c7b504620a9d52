#!/bin/bash
mkdir -p gpurun_out
python scripts/loss_bench.py > gpurun_out/loss_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dino_loss_staged -s 4 -c 2 -o gpurun_out/prof_loss python scripts/loss_bench.py > gpurun_out/ncu_loss.log 2>&1
echo "ncu exit $?"; cat gpurun_out/loss_plain.log | tail -n 4

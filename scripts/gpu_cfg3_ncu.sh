#!/bin/bash
# launch list of the cfg3 step (MultiCropDistillStep): 3 warm-up + 2 timed steps
mkdir -p gpurun_out
NSTEPS=2 python scripts/cfg3_step.py > gpurun_out/cfg3_plain.log 2>&1 &&
NSTEPS=2 timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_cfg3.csv python scripts/cfg3_step.py > gpurun_out/ncu_cfg3.log 2>&1
echo "ncu exit $?"

#!/bin/bash
# cfg4 (H = 512 cluster recurrence) under ncu.  ncu serialises kernels, so the GEMMs that normally run BESIDE the recurrence
# (and feed / follow it through flags) are switched to the sequential order: CSN_LSTM_NO_OVERLAP=1.
mkdir -p gpurun_out
TAG=${1:-r02}
export CSN_LSTM_NO_OVERLAP=1
NSTEPS=2 python scripts/bench_cfg4.py > gpurun_out/plain_cfg4_${TAG}.log 2>&1 &&
NSTEPS=1 timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg4_${TAG}.csv python scripts/bench_cfg4.py > gpurun_out/ncu_launches_cfg4_${TAG}.log 2>&1
echo "cfg4 launch list exit $?"
NSTEPS=1 timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"lstm_(fwd|bwd)_cluster_kernel" -s 4 -c 4 -o gpurun_out/prof_cfg4_${TAG} python scripts/bench_cfg4.py > gpurun_out/ncu_full_cfg4_${TAG}.log 2>&1
echo "cfg4 full capture exit $?"
ls -la gpurun_out | grep cfg4_${TAG}

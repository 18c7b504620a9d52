#!/bin/bash
# Per-kernel launch list + full capture of the hot kernels (one GPU; plain run first, as the recipe requires).
mkdir -p gpurun_out
TAG=${1:-r02}
CMD="python bench.py --steps 2 --warmup 3 --no_cpu_baseline --configs="
$CMD > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"lstm_(fwd|bwd)_tc_kernel|sosfilt_(stream|warp)|bptt_finalize|dino_loss_staged_kernel|head_dino_kernel|dp_adam_peer" -s 30 -c 12 -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out | grep ${TAG}

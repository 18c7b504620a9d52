#!/bin/bash
# Per-kernel launch list + full capture of the recurrence kernels (one GPU; plain run first, as the recipe requires).
mkdir -p gpurun_out
python -m pytest tests/test_gpu_filter.py -q -m gpu -k "errors_and_empty" > gpurun_out/t_filter_fix.log 2>&1; echo "filter_fix exit $?"
CMD="python bench.py --steps 2 --warmup 3 --no_cpu_baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"lstm_(fwd|bwd)_tc_kernel|sosfilt_stream|gemm_tc_kernel" -s 16 -c 8 -o gpurun_out/prof_r01 $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out | head -30

#!/bin/bash
# First GPU trip: non-tensor-core kernels first, then tcgen05 bring-up in increasing order of risk.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; timeout -s KILL 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary.txt; }
run t_filter python -m pytest tests/test_gpu_filter.py -q -m gpu --timeout 300
run t_loss python -m pytest tests/test_gpu_loss.py -q -m gpu --timeout 300
run t_gemm_f32 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k "gemm_f32" --timeout 300
run t_lstm_f32 python -m pytest tests/test_gpu_lstm.py -q -m gpu -k "f32_layer or zero_weights" --timeout 300
run t_step_f32 python -m pytest tests/test_gpu_step.py -q -m gpu -k "not bfloat16" --timeout 300
run t_umma python -m pytest tests/test_gpu_gemm.py -q -m gpu -k "umma_tile" --timeout 120
run t_gemm_tc python -m pytest tests/test_gpu_gemm.py -q -m gpu -k "bf16" --timeout 120
run t_lstm_tc python -m pytest tests/test_gpu_lstm.py -q -m gpu -k "bf16" --timeout 200
run t_step_bf16 python -m pytest tests/test_gpu_step.py -q -m gpu -k "bfloat16" --timeout 200
run smoke python __graft_entry__.py smoke
run bench_fp32 python bench.py --steps 3 --warmup 3 --precision fp32 --no_cpu_baseline
run bench_bf16 python bench.py --steps 10 --warmup 3
cat gpurun_out/summary.txt
tail -5 gpurun_out/bench_bf16.log

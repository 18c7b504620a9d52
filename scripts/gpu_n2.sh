#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 20 --warmup 5 --no_cpu_baseline > gpurun_out/bench_n2.log 2>&1
echo "n2 exit $?"
grep '^{' gpurun_out/bench_n2.log | cut -c1-300; tail -n 3 gpurun_out/bench_n2.log | cut -c1-300

#!/bin/bash
# 2-GPU check: the world-2 GPU tests, then the bench line (cfg2 + the cfg3/cfg4/cfg5 sub-records) exactly as the driver launches it
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_dp_peer.py tests/test_gpu_multicrop_step.py -x -q -m gpu > gpurun_out/t_n2.log 2>&1
echo "tests exit $?"; tail -n 3 gpurun_out/t_n2.log
S=$(date +%s)
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/bench_n2.log 2>&1
echo "n2 exit $? wall $(( $(date +%s) - S )) s"
grep '^{' gpurun_out/bench_n2.log | cut -c1-400; tail -n 3 gpurun_out/bench_n2.log | cut -c1-300

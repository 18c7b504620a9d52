#!/bin/bash
# 8-GPU box: the bench line exactly as the driver launches it at N = 8 and N = 4 (cfg2 + cfg3/cfg4/cfg5 sub-records)
mkdir -p gpurun_out
for N in 8 4; do
  S=$(date +%s)
  timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2971$N bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/bench_n$N.log 2>&1
  echo "n$N exit $? wall $(( $(date +%s) - S )) s"
  grep '^{' gpurun_out/bench_n$N.log | cut -c1-300; tail -n 2 gpurun_out/bench_n$N.log | cut -c1-300
done

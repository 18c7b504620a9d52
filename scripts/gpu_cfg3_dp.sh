#!/bin/bash
# cfg3 on two ranks: CTAs given to the exchange kernels that run beside BPTT (CSN_DP_XCHG_BLOCKS; 0 = default, 2 per SM)
run() { CSN_DP_XCHG_BLOCKS=$1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29733 scripts/cfg3_step.py 2>&1 | grep "cfg3 B" | cut -c1-70; }
for b in 0 16 32 64 148 444 592 1184; do echo "blocks $b"; run $b; done

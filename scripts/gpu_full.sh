#!/bin/bash
# the driver's round-end sequence on one GPU: full -m gpu suite, smoke(), bench (ours + reference arm)
mkdir -p gpurun_out
TAG=${1:-x}
timeout -s KILL 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/t_${TAG}.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/t_${TAG}.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_${TAG}.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/smoke_${TAG}.log
S=$(date +%s)
timeout -s KILL 900 python bench.py > gpurun_out/b_${TAG}.log 2>&1; echo "bench exit $? wall $(( $(date +%s) - S )) s"
grep '^{' gpurun_out/b_${TAG}.log | cut -c1-200

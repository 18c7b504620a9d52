#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout_s> <gpus> '<command>'  -- retries while the pod answers busy / transient
T=$1; G=$2; shift 2
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then OUT=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1); else OUT=$(/usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$@" 2>&1); fi
  RC=$?
  echo "$OUT" | tail -6
  if echo "$OUT" | grep -q "status=transient\|retry in a few minutes\|status=busy\|no box"; then sleep 45; continue; fi
  exit $RC
done
exit 3

#!/bin/bash
# usage: gpurun_retry.sh <gpus> <timeout_s> '<command>'   -- retries while the pool answers "busy" (nothing is charged for those)
G=$1; TO=$2; shift 2
for i in 1 2 3 4 5 6 7 8; do
  out=$(gpurun --gpus $G --timeout $TO -- "$@" 2>&1)
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient\|status=busy\|rc=3"; then sleep 120; continue; fi
  break
done

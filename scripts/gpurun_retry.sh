#!/bin/bash
# retry a gpurun call while the pod answers busy/transient (nothing is charged for those)
# usage: [GPUS=N] scripts/gpurun_retry.sh <timeout_s> <command...>
T=$1; shift
G=""
if [ -n "$GPUS" ] && [ "$GPUS" != "1" ]; then G="--gpus $GPUS"; fi
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun $G --timeout "$T" -- "$@" 2>&1)
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient\|status=busy\|rc=None"; then sleep 90; continue; fi
  break
done

#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> [--gpus N] -- <command>: retries while the pod answers "transient / busy" (nothing charged)
T=$1; shift
for attempt in 1 2 3 4 5 6 7 8; do
  out=$(/usr/local/graft/bin/gpurun --timeout $T "$@" 2>&1)
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient\|status=busy\|exit code 3"; then sleep 90; continue; fi
  break
done

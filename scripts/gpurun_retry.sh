#!/bin/bash
# usage: scripts/gpurun_retry.sh <gpurun args...>   -- retries while the pod answers "transient" (no box / slot free; nothing charged)
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3

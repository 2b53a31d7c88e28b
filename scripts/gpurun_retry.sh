#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout_s> <command...>   (retries while the pod answers busy; build container only)
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > /tmp/gpurun_last.txt 2>&1
  if grep -q "status=transient" /tmp/gpurun_last.txt; then sleep 90; continue; fi
  break
done
tail -40 /tmp/gpurun_last.txt

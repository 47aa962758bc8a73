#!/bin/bash
# usage: tools/gpurun_retry.sh [gpurun options] -- 'command'   -- retries while the pod answers "transient" (nothing charged)
for attempt in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|rc=3\b"; then sleep 90; continue; fi
  echo "$out"
  exit 0
done
echo "$out"
exit 3

#!/bin/bash
# gpurun with retries while the pod answers "transient" (busy): tools/gpurun_retry.sh [gpurun flags] -- '<command>'
for attempt in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then
    sleep 90
    continue
  fi
  echo "$out" | tail -60
  exit 0
done
echo "gpurun: still busy after 30 attempts"
exit 3

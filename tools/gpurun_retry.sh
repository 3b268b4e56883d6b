#!/bin/bash
# usage: tools/gpurun_retry.sh LOG [gpurun args...]   - retries while the pod answers "transient" (nothing charged)
LOG=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  if grep -q "status=transient\|status=busy" "$LOG"; then sleep 90; continue; fi
  break
done

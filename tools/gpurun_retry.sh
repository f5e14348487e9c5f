#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <gpurun args...>   — retries while the pod answers busy/transient (nothing charged)
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1; rc=$?
  if grep -q "status=transient\|status=busy\|nothing was charged" "$log" || [ $rc -eq 3 ]; then sleep 60; continue; fi
  break
done
echo "rc=$rc after $i tries"; tail -60 "$log"

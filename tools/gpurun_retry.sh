#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3 / status=transient): tools/gpurun_retry.sh <log> <gpurun args...>
log=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient" "$log" || [ $rc -eq 3 ]; then sleep 150; continue; fi
  exit $rc
done
exit 3

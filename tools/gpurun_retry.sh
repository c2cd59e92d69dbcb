#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> '<command>' [extra gpurun args]  -- retries while the pod is busy
T=$1; shift; CMD=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$T" "$@" -- "$CMD"; rc=$?
  if [ $rc -ne 3 ] && ! grep -q '"status": "transient"' /root/repo/gpurun_out/.last_call.json 2>/dev/null; then exit $rc; fi
  echo "[retry] attempt $i busy; sleeping 120 s"; sleep 120
done
exit 3

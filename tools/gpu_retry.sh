#!/usr/bin/env bash
# Local helper: keep asking gpurun for a box until it is not "busy" (exit code 3).
#   tools/gpu_retry.sh <log> [gpurun args...] -- '<command>'
LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3

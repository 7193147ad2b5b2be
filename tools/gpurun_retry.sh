#!/bin/bash
# Retry wrapper around gpurun for a busy pod (exit code 3 = no slot right now, nothing charged): tools/gpurun_retry.sh [gpurun args...]
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3

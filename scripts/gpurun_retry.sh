#!/usr/bin/env bash
# gpurun with retries while the pod has no free slot (exit code 3):  scripts/gpurun_retry.sh [gpurun args...] -- 'command'
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3

#!/bin/bash
# gpurun with retries while the pod answers "busy" (nothing is charged for those).  usage: gpurun_retry.sh LOG TIMEOUT CMD
LOG=$1; TMO=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TMO -- "$@" > $LOG 2>&1
  if ! grep -q "status=transient" $LOG; then exit 0; fi
  sleep 120
done
exit 3

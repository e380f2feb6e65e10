#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3): [GPUS=N] tools/gpurun_retry.sh <timeout> '<command>'
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout "$1" -- "$2"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3

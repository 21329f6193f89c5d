#!/bin/bash
# usage: gpurun_retry.sh <timeout> <command...>   -- retries while the pod answers "busy" (rc 3)
T=$1; shift
for attempt in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout $T "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3

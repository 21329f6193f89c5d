#!/bin/bash
# usage: gpurun_retry_n.sh <gpus> <timeout> -- <command>
G=$1; T=$2; shift; shift
for attempt in $(seq 1 10); do
  /usr/local/graft/bin/gpurun --gpus $G --timeout $T "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3

#!/bin/bash
# usage: tools/gpurun_retry8.sh <gpus> <timeout> <command...>
G=$1; T=$2; shift; shift
for i in $(seq 1 15); do
  /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] busy, attempt $i"; sleep 120
done
exit 3

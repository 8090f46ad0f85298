#!/bin/bash
# usage: tools/gpurun_retry_n.sh <gpus> <timeout> "<command>" <stdout-file>  -- retries while the pod is busy
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --gpus $1 --timeout $2 -- "$3" > $4 2>&1
  if ! grep -q "status=transient\|rc=3\b" $4; then exit 0; fi
  sleep 120
done

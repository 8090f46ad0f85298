#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <script> <stdout-file>   -- retries while the pod answers "transient"/busy
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $1 -- "bash $2" > $3 2>&1
  if ! grep -q "status=transient\|rc=3\b" $3; then exit 0; fi
  sleep 90
done

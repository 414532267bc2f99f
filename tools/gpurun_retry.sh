#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <script> <log>   -- retries while the pod answers busy (exit 3) / transient
T=$1; S=$2; L=$3
for i in 1 2 3 4 5 6 7 8; do
  gpurun --timeout $T -- "bash $S" > $L 2>&1
  if grep -q "status=ok\|status=fail\|status=timeout" $L; then break; fi
  sleep 90
done
tail -60 $L

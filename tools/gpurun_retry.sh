#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <command...>   (retries while the pod answers "transient"/busy)
to=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $to -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|exit code 3\|no box\|busy"; then sleep 90; continue; fi
  echo "$out" | tail -25; exit 0
done
echo "gave up"; exit 1

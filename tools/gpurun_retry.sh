#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <timeout_s> [--gpus N] -- <command>   (retries while the pod answers "transient"/busy)
log=$1; shift; to=$1; shift
extra=()
while [ "$1" != "--" ]; do extra+=("$1"); shift; done
shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" "${extra[@]}" -- "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient\|rc=3\|no box\|busy" "$log" && ! grep -q "status=ok\|status=fail" "$log"; then
    sleep 90
    continue
  fi
  break
done
echo "gpurun_retry finished rc=$rc attempt=$attempt" >> "$log"

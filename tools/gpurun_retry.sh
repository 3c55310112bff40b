#!/bin/bash
# gpurun with retries while the pod has no free GPU slot (exit code 3 / status=transient).
# usage: tools/gpurun_retry.sh [--gpus N] <timeout-seconds> '<command>'
GP=""
if [ "$1" = "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun $GP --timeout "$T" -- "$@" 2>&1)
  rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$out"
  exit $rc
done
echo "gave up: no GPU slot"
exit 3

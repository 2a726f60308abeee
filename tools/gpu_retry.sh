#!/bin/bash
# Diagnostic helper (not part of the product): gpurun with retries while the pod has no free GPU slot (exit code 3).
# usage: tools/gpu_retry.sh [--gpus N] TIMEOUT 'command'
GP=""
if [ "$1" == "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun $GP --timeout $T -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 100
done
exit 3

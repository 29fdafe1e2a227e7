#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3: nothing charged).  usage: tools/gpu_retry.sh [--gpus N] TIMEOUT 'command'
GP=""
if [ "$1" = "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun $GP --timeout $T -- "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 60
done
exit 3

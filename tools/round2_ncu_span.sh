#!/bin/bash
# ncu --set full capture of the histogram kernel under the given env (after the same command ran clean without ncu)
# usage: bash tools/round2_ncu_span.sh tag "ENV=.." [route]
TAG=$1; ENVS=$2; ROUTE=${3:-spans}
env $ENVS SNPG_NO_GRAPH=1 python tools/prof_step.py --steps 2 --route $ROUTE > /dev/null 2>&1 && \
env $ENVS SNPG_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k 'regex:k_span_hist|k_fused_hist' --launch-skip 1 -c 1 -f -o gpurun_out/r2_${TAG} python tools/prof_step.py --steps 2 --route $ROUTE > gpurun_out/ncu_${TAG}.log 2>&1

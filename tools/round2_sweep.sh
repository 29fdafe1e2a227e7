#!/bin/bash
# A/B sweep on one B200: each line of SWEEP is a set of env assignments; prints the step time and per-kernel times.
# usage: bash tools/round2_sweep.sh tag "ENV1=a ENV2=b" "ENV1=c" ...
TAG=$1; shift
O=gpurun_out/r2_sweep_${TAG}.txt
: > $O
for envs in "$@"; do
  env $envs python bench.py --no-cpu --no-extras --steps 40 --warmup 5 2> gpurun_out/sweep.err | python -c "
import json,sys
lines=sys.stdin.read().strip().splitlines()
if not lines: print('$envs FAILED'); sys.exit()
d=json.loads(lines[-1]); r=d['roofline']
print('%-40s step_ms %.4f hist GB/s %.0f frac %.3f' % ('$envs', d['ms_per_step'], r['achieved'], r['frac']), {k: round(v['ms_per_step']*1000, 1) for k, v in d['kernels'].items()})" >> $O
  tail -2 gpurun_out/sweep.err >> $O
done
cat $O

#!/bin/bash
# A/B lines: bash tools/round2_ab.sh tag "ENV.. | bench args" ...   (each argument: env assignments, then '|', then bench.py arguments)
TAG=$1; shift
O=gpurun_out/r2_ab_${TAG}.txt
: > $O
for spec in "$@"; do
  envs="${spec%%|*}"; args="${spec#*|}"
  env $envs python bench.py --no-cpu --no-extras --steps 40 --warmup 5 $args 2> gpurun_out/sweep.err | python -c "
import json,sys
lines=sys.stdin.read().strip().splitlines()
if not lines: print('''$spec FAILED'''); sys.exit()
d=json.loads(lines[-1]); r=d['roofline']
print('%-50s step_ms %.4f %s GB/s %.0f frac %.3f' % ('''$spec''', d['ms_per_step'], r['kernel'], r['achieved'], r['frac']), {k: round(v['ms_per_step']*1000, 1) for k, v in d['kernels'].items()})" >> $O
  tail -2 gpurun_out/sweep.err >> $O
done
cat $O

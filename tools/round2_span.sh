#!/bin/bash
# one gpurun call: parity of every histogram route, then an A/B sweep of the TMA-fed span kernel's shapes
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bit_exact or ragged or more_than or within_job or fuzz or annotation" > gpurun_out/r2_span_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_span_pytest.log
O=gpurun_out/r2_span_sweep.txt
: > $O
run() {
  env $1 python bench.py --no-cpu --no-extras --steps 40 --warmup 5 $2 2> gpurun_out/sweep.err | python -c "
import json,sys
lines=sys.stdin.read().strip().splitlines()
if not lines: print('$1 $2 FAILED'); sys.exit()
d=json.loads(lines[-1]); r=d['roofline']
print('%-46s step_ms %.4f %s GB/s %.0f frac %.3f' % ('$1 $2', d['ms_per_step'], r['kernel'], r['achieved'], r['frac']), {k: round(v['ms_per_step']*1000, 1) for k, v in d['kernels'].items()})" >> $O
  tail -2 gpurun_out/sweep.err >> $O
}
run "SNPG_X=0" "--route fused"
run "SNPG_SPAN_SHAPE=0" ""
run "SNPG_SPAN_SHAPE=1" ""
run "SNPG_SPAN_SHAPE=2" ""
run "SNPG_SPAN_SHAPE=3" ""
run "SNPG_SPAN_SHAPE=1 SNPG_SPAN_ROWS=64" ""
run "SNPG_SPAN_SHAPE=1 SNPG_SPAN_ROWS=16" ""
cat $O

#!/bin/bash
# Regenerates the round's evidence under gpurun_out/ (run with gpurun on one B200; copy what is kept into profiles/).
# Bench numbers first, without a profiler; the ncu passes only after the same command has exited 0.
R=${1:-r1}
O=gpurun_out
set -x
python bench.py > $O/${R}_bench_fused.json 2> $O/bench_fused.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > $O/${R}_bench_reference.json 2> $O/bench_reference.err
python bench.py --route tiles --no-cpu --steps 50 > $O/${R}_bench_tiles.json 2> $O/bench_tiles.err
python bench.py --route codes --no-cpu --steps 50 > $O/${R}_bench_codes.json 2> $O/bench_codes.err
for p in 0.3 0.03 0.003; do python bench.py --p-poly $p --no-cpu --steps 50 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('p_poly $p step_ms %.4f GB/s %.0f %.3f' % (d['ms_per_step'], r['achieved'], r['frac']), {k: round(v['ms_per_step']*1000, 1) for k, v in d['kernels'].items()})"; done > $O/${R}_density_sweep.txt
python tools/bench_config4.py > $O/${R}_config4.json 2> $O/config4.err
python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${R}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_launches.log 2>&1
SNPG_NO_GRAPH=1 python tools/prof_step.py --steps 2 > /dev/null 2>&1 && \
SNPG_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k 'regex:k_fused_hist|k_bootstrap|k_vote_refs|k_fused_fixup|k_stats|k_moments' --launch-skip 5 -c 5 -f -o $O/${R}_full python tools/prof_step.py --steps 2 > $O/ncu_full.log 2>&1
ls -la $O

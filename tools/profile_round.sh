#!/bin/bash
# Regenerates the round's evidence under gpurun_out/ (run with gpurun on one B200; copy what is kept into profiles/).
# Bench numbers first, without a profiler; the ncu passes only after the same command has exited 0.
R=${1:-r2}
O=gpurun_out
set -x
python bench.py > $O/${R}_bench_fused.json 2> $O/bench_fused.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > $O/${R}_bench_reference.json 2> $O/bench_reference.err
python bench.py --route tiles --no-cpu --steps 50 > $O/${R}_bench_tiles.json 2> $O/bench_tiles.err
python bench.py --route codes --no-cpu --steps 50 > $O/${R}_bench_codes.json 2> $O/bench_codes.err
python bench.py --route spans --no-cpu --steps 50 > $O/${R}_bench_spans.json 2> $O/bench_spans.err
for k in 0 1; do python bench.py --boot-kernel $k --no-cpu --steps 50 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('bootstrap kernel $k (0 gather, 1 weights; default 2 classes) step_ms %.4f' % d['ms_per_step'], {k: round(v['ms_per_step']*1000, 1) for k, v in d['kernels'].items()})"; done > $O/${R}_boot_kernels.txt
python bench.py --max-changes 1 --max-alleles 1 --no-cpu --steps 50 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('one single-nucleotide allele per polymorphic codon: step_ms %.4f GB/s %.0f %.3f' % (d['ms_per_step'], r['achieved'], r['frac']))" > $O/${R}_single_allele.txt
for p in 0.3 0.03 0.003; do python bench.py --p-poly $p --no-cpu --steps 50 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('p_poly $p step_ms %.4f GB/s %.0f %.3f' % (d['ms_per_step'], r['achieved'], r['frac']), {k: round(v['ms_per_step']*1000, 1) for k, v in d['kernels'].items()})"; done > $O/${R}_density_sweep.txt
python tools/bench_config4.py > $O/${R}_config4.json 2> $O/config4.err
python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${R}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_launches.log 2>&1
SNPG_NO_GRAPH=1 python tools/prof_step.py --steps 2 > /dev/null 2>&1 && \
SNPG_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k 'regex:k_fused_hist|k_bootstrap|k_boot_|k_vote_refs|k_fused_fixup|k_stats|k_moments' --launch-skip 7 -c 7 -f -o $O/${R}_full python tools/prof_step.py --steps 2 > $O/ncu_full.log 2>&1
for k in k_fused_hist k_bootstrap_mma k_tile_hist k_boot_split; do echo "== $k"; sh tools/sass_of.sh $k | awk '{print $2}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -12; done > $O/${R}_sass_mix.txt
ls -la $O

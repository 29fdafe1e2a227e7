#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_c3_ab.txt; : > $O
for r in 127 255 511; do echo "SNPG_BOOT_MINROW=$r" >> $O; SNPG_BOOT_MINROW=$r python tools/prof_config3.py >> $O 2>&1
SNPG_BOOT_MINROW=$r python bench.py --no-cpu --no-extras --steps 40 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), {k:round(v['ms_per_step']*1e3,1) for k,v in d['kernels'].items()})" >> $O 2>&1
done
cat $O

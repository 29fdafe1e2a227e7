#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2_boot_sweep2.txt
for cfg in "SNPG_BOOT_NB=1" "SNPG_BOOT_NB=2" "SNPG_BOOT_NB=1 SNPG_BOOT_SPLIT=1" "SNPG_BOOT_NB=2 SNPG_BOOT_SPLIT=1"; do
  env $cfg python bench.py --no-cpu --steps 50 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$cfg', round(d['ms_per_step'],4), {k:round(v['ms_per_step']*1e3,1) for k,v in d['kernels'].items()})" >> gpurun_out/r2_boot_sweep2.txt 2>&1
done

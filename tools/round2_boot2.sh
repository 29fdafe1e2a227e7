#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bootstrap" > gpurun_out/r2_boot_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_boot_pytest.log
rm -f gpurun_out/r2_boot_sweep.txt
for sh in 4 5 6 7 8; do
  SNPG_BOOT_CLASS_SHIFT=$sh python bench.py --no-cpu --steps 50 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('shift $sh', round(d['ms_per_step'],4), {k:round(v['ms_per_step']*1e3,1) for k,v in d['kernels'].items()})" >> gpurun_out/r2_boot_sweep.txt 2>&1
done
bash tools/round2_boot_ncu.sh

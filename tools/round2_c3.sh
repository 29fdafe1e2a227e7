#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_jobs.py tests/test_gpu_parity.py -m gpu -x -q -k "window or job or config3" > gpurun_out/r2_c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_c3_pytest.log
python bench.py --no-cpu --steps 50 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), {k:round(v['ms_per_step']*1e3,1) for k,v in d['kernels'].items()}); print(d.get('config3'))" > gpurun_out/r2_c3.txt 2>&1

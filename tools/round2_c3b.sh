#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_jobs.py -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
python bench.py > gpurun_out/r2_bench_fused.json 2> gpurun_out/bench_fused.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_fused.json').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), {k:round(v['ms_per_step']*1e3,1) for k,v in d['kernels'].items()}, 'e2e', d['e2e']['ms_per_step'], 'c3', d['config3']['ms_per_step'], 'c4', d['config4']['ms_per_step'], d['check'])" > gpurun_out/r2_c3_ab.txt 2>&1
cat gpurun_out/r2_c3_ab.txt; tail -n 3 gpurun_out/r2_pytest.log

#!/bin/bash
# class-aggregated bootstrap: its tests, the other bootstrap/job tests, then the bench line with both bootstrap kernels
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bootstrap" > gpurun_out/r2_boot_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_boot_pytest.log
python -m pytest tests/test_gpu_jobs.py tests/test_rwindows.py tests/test_pooled.py -m gpu -x -q > gpurun_out/r2_boot_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_boot_pytest2.log
python bench.py --no-cpu --steps 50 --warmup 5 > gpurun_out/r2_boot_bench.json 2> gpurun_out/r2_boot_bench.err; echo "rc=$?" >> gpurun_out/r2_boot_bench.err

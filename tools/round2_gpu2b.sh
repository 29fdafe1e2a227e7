#!/bin/bash
mkdir -p gpurun_out
python tools/cli_wall.py --gpus 2 > gpurun_out/r2_cli_wall_n2.txt 2>&1; echo "cli rc=$?" >> gpurun_out/r2_cli_wall_n2.txt
python -m pytest tests/test_gpu_jobs.py tests/test_perl_host.py -m gpu -x -q > gpurun_out/r2_n2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_n2_pytest.log

#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
bash tools/profile_round.sh r2 > gpurun_out/r2_profile_round.log 2>&1
python tools/cli_wall.py > gpurun_out/r2_cli_wall.txt 2>&1; echo "cli rc=$?" >> gpurun_out/r2_cli_wall.txt

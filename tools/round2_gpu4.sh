#!/bin/bash
# one gpurun call: the whole GPU suite, then the wall time of the Perl command line with the library's trace
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
python tools/cli_wall.py > gpurun_out/r2_cli_wall.txt 2>&1; echo "cli rc=$?" >> gpurun_out/r2_cli_wall.txt

#!/bin/bash
# whole GPU suite, smoke, the bench line (both arms), launch list and full capture of the step's kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke.log 2>&1
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "rc=$?" >> gpurun_out/r2_bench.err

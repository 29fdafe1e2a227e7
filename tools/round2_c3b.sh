#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
bash tools/round2_ab.sh roll "SNPG_SPAN_VARIANT=0|" "SNPG_SPAN_VARIANT=5|" "SNPG_SPAN_VARIANT=0|" > /dev/null 2>&1
cp gpurun_out/r2_ab_roll.txt gpurun_out/r2_c3_ab.txt
python tools/chunk_sweep.py >> gpurun_out/r2_c3_ab.txt 2>&1; python tools/chunk_sweep.py --hiv --n 2500 5000 >> gpurun_out/r2_c3_ab.txt 2>&1
python tools/prof_config3.py >> gpurun_out/r2_c3_ab.txt 2>&1
cat gpurun_out/r2_c3_ab.txt; tail -n 3 gpurun_out/r2_pytest.log

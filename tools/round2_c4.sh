#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "between or config4 or taxa or group" > gpurun_out/r2_c4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_c4_pytest.log
python tools/bench_config4.py --calls > gpurun_out/r2_c4.json 2> gpurun_out/c4.err
python tools/bench_config4.py >> gpurun_out/r2_c4.json 2>> gpurun_out/c4.err
SNPG_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --launch-skip 200 -c 120 --csv --log-file gpurun_out/r2_c4_launches.csv python tools/bench_config4.py --steps 3 > gpurun_out/r2_c4_ncu.log 2>&1
tail -n 3 gpurun_out/r2_c4_pytest.log; cat gpurun_out/r2_c4.json; tail -n 3 gpurun_out/c4.err

#!/bin/bash
mkdir -p gpurun_out
SNPG_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_c3_launches.csv python tools/bench_config3.py > gpurun_out/r2_c3_ncu.log 2>&1

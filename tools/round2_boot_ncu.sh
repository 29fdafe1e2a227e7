#!/bin/bash
mkdir -p gpurun_out
SNPG_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_boot_launches.csv python tools/prof_step.py --steps 3 > gpurun_out/r2_boot_ncu.log 2>&1

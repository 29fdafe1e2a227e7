#!/bin/bash
mkdir -p gpurun_out
SNPG_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:"k_boot_split|k_boot_classify" -c 2 -o gpurun_out/r2_split python tools/prof_step.py --steps 1 > gpurun_out/ncu_split.log 2>&1

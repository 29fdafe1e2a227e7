#!/bin/bash
mkdir -p gpurun_out
SNPG_NO_GRAPH=1 python tools/prof_config3.py --steps 2 > /dev/null 2>&1 && \
SNPG_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k 'regex:k_bootstrap_mma|k_boot_classify' --launch-skip 6 -c 3 -f -o gpurun_out/r2_c3boot python tools/prof_config3.py --steps 2 > gpurun_out/ncu_c3boot.log 2>&1

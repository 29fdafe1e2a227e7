#!/bin/bash
mkdir -p gpurun_out
for v in 7 8; do
  SNPG_SPAN_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/roll_pytest_v$v.log 2>&1; echo "rc=$?" >> gpurun_out/roll_pytest_v$v.log
done
bash tools/round2_ab.sh roll "SNPG_SPAN_VARIANT=0|" "SNPG_SPAN_VARIANT=7|" "SNPG_SPAN_VARIANT=8|" "SNPG_SPAN_VARIANT=5|" \
  "SNPG_SPAN_VARIANT=0 SNPG_SPAN_CARVE=66|" "SNPG_SPAN_VARIANT=0 SNPG_SPAN_CARVE=100|" "SNPG_SPAN_VARIANT=7 SNPG_SPAN_CARVE=34|" "SNPG_SPAN_VARIANT=8 SNPG_SPAN_CARVE=26|" "SNPG_SPAN_VARIANT=5 SNPG_SPAN_CARVE=50|"

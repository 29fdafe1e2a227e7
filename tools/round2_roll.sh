#!/bin/bash
# double-buffered span kernel: parity under the variants, then A/B bench lines
mkdir -p gpurun_out
for v in 5; do
  SNPG_SPAN_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_jobs.py -m gpu -x -q > gpurun_out/roll_pytest_v$v.log 2>&1; echo "rc=$?" >> gpurun_out/roll_pytest_v$v.log
done
bash tools/round2_ab.sh roll "SNPG_SPAN_VARIANT=0|" "SNPG_SPAN_VARIANT=4|" "SNPG_SPAN_VARIANT=5|" "SNPG_SPAN_VARIANT=6|" \
  "SNPG_SPAN_VARIANT=5 SNPG_SPAN_ROWS=64|" "SNPG_SPAN_VARIANT=5 SNPG_SPAN_ROWS=48|" "SNPG_SPAN_VARIANT=5 SNPG_SPAN_TAIL=4|"
bash tools/round2_ncu_span.sh v9 "SNPG_SPAN_VARIANT=5" fused

python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
bash tools/round2_sweep.sh hist1 "SNPG_SPAN_VARIANT=0" "SNPG_SPAN_VARIANT=2" "SNPG_SPAN_VARIANT=3" "SNPG_SPAN_VARIANT=2 SNPG_BOOT_NB=1" "SNPG_SPAN_VARIANT=2 SNPG_BOOT_SPLIT=1 SNPG_BOOT_NB=1" > /dev/null 2>&1
SNPG_NO_GRAPH=1 python tools/prof_step.py --steps 2 > /dev/null 2>&1 && \
SNPG_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k 'regex:k_fused_hist|k_bootstrap' --launch-skip 2 -c 2 -f -o gpurun_out/r2_full python tools/prof_step.py --steps 2 > gpurun_out/ncu_full.log 2>&1

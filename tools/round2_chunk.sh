#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_chunk_sweep.txt; : > $O
for r in auto 4 8 12 16 24 32; do
  if [ $r = auto ]; then python tools/chunk_sweep.py >> $O 2>&1; python tools/chunk_sweep.py --hiv --n 2500 5000 >> $O 2>&1
  else SNPG_SPAN_ROWS=$r python tools/chunk_sweep.py >> $O 2>&1; SNPG_SPAN_ROWS=$r python tools/chunk_sweep.py --hiv --n 2500 5000 >> $O 2>&1; fi
done
python tools/prof_config3.py >> $O 2>&1
cat $O

timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/t9.log
timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/b9_def.json 2> gpurun_out/b9_def.err
SNPG_SPAN_TAIL=2 timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/b9_t2.json 2> gpurun_out/b9_t2.err
SNPG_SPAN_TAIL=8 timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/b9_t8.json 2> gpurun_out/b9_t8.err
SNPG_SPAN_ROWS=24 timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/b9_r24.json 2> gpurun_out/b9_r24.err
SNPG_SPAN_ROWS=40 timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/b9_r40.json 2> gpurun_out/b9_r40.err

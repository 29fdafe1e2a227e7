SNPG_BOOT_FORM=0 timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/b11_f0u2.json 2> gpurun_out/b11_f0.err

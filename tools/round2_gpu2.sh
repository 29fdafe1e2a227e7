#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_jobs.py -m gpu -x -q -k "team or processes or unconnected" > gpurun_out/r2_n2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_n2_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 5 \
  > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench rc=$?" >> gpurun_out/r2_bench_n2.err
tail -n 2 gpurun_out/r2_n2_pytest.log; tail -n 2 gpurun_out/r2_bench_n2.err

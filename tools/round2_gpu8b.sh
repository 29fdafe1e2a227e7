#!/bin/bash
# one gpurun --gpus 8 call: the bench line at N = 8 (with config 5), N = 4 and N = 2
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 50 --warmup 5 \
  > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench rc=$?" >> gpurun_out/r2_bench_n8.err
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 50 --warmup 5 --no-extras \
  > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; echo "bench rc=$?" >> gpurun_out/r2_bench_n4.err
CUDA_VISIBLE_DEVICES=0,1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 5 --no-extras \
  > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench rc=$?" >> gpurun_out/r2_bench_n2.err
tail -n 2 gpurun_out/r2_bench_n8.err gpurun_out/r2_bench_n4.err gpurun_out/r2_bench_n2.err

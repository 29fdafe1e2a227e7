#!/bin/bash
# one gpurun --gpus 8 call: multi-GPU tests, the bench line at N = 8 (with config 5 checked against an unsharded run) and
# at N = 4 (four of the eight GPUs), the Perl command line on 8 GPUs
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_n8_gpus.txt
python -m pytest tests/test_gpu_jobs.py -m gpu -x -q -k "team or processes or unconnected" > gpurun_out/r2_n8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_n8_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 50 --warmup 5 \
  > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench rc=$?" >> gpurun_out/r2_bench_n8.err
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 50 --warmup 5 \
  > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; echo "bench rc=$?" >> gpurun_out/r2_bench_n4.err
python tools/cli_wall.py --gpus 8 > gpurun_out/r2_cli_wall_n8.txt 2>&1; echo "cli rc=$?" >> gpurun_out/r2_cli_wall_n8.txt

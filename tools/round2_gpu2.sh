#!/bin/bash
# one gpurun --gpus N call: the multi-GPU tests, then the bench line at N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_n${N}_gpus.txt
python -m pytest tests/test_gpu_jobs.py -m gpu -x -q -k "team or processes or unconnected" > gpurun_out/r2_n${N}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_n${N}_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 \
  > gpurun_out/r2_bench_n${N}.json 2> gpurun_out/r2_bench_n${N}.err; echo "bench rc=$?" >> gpurun_out/r2_bench_n${N}.err

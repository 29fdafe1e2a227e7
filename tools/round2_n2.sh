#!/bin/bash
# one gpurun --gpus 2 call: the bench line at N = 2 (ONE alignment split over two GPUs, checked bit for bit against one GPU) and the
# CPU arm launched the way the driver launches it (rank 0 alone works)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 5 --no-extras \
  > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench rc=$?" >> gpurun_out/r2_bench_n2.err
tail -3 gpurun_out/r2_bench_n2.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2_bench_n2.json").read().strip().splitlines()[-1])
print(d["n_gpus"], round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 3), d.get("check"), d.get("exchange"))
P

#!/usr/bin/env python
"""BASELINE config 5 on its own: ONE within-group alignment of 1,000,000 sequences x 29,903 nt, 10,000 bootstraps, its codon
axis and its bootstrap replicates split over the ranks, through the library's whole-job call (the ranks exchange over
peer memory).  The same routine as bench.py's `config5` key (bench.config5), which also compares every output with an
unsharded run of the gathered alignment on rank 0, bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/bench_config5.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from snpgenie_b200 import synth  # noqa: E402
from snpgenie_b200.engine import Engine  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
stream = torch.cuda.Stream(device=dev)
with torch.cuda.stream(stream):
    res = bench.config5(torch, dist, Engine, synth, stream, dev, rank, world, local)
if rank == 0:
    print(json.dumps(res))
dist.barrier()
dist.destroy_process_group()
ok = res.get("equals_unsharded_run_bitwise") if rank == 0 else None
sys.exit(0 if (ok is None or all(ok.values())) else 1)

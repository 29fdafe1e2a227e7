#!/usr/bin/env python
"""torchrun check of the codon-range-sharded path (one alignment split over ranks, NCCL gathers).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/run_sharded_check.py [--n 4000] [--boot 2000]
Rank 0 also runs the whole alignment unsharded and compares: per-codon stats and product sums must be
identical, bootstrap SEs statistically equal.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from snpgenie_b200 import multigpu, synth  # noqa: E402
from snpgenie_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4000)
ap.add_argument("--boot", type=int, default=4000)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
products = synth.sars2_products()
aln = synth.make_alignment(a.n, 29903, products, 99, gap_frac=0.01, n_frac=0.01)    # same on every rank (same seed)
with Engine(device=local, seed=5, rank=rank, world=world) as eng:
    eng.set_products(products, 29903)
    eng.load_group(0, aln)
    sums, se, stats = multigpu.within_group_sharded(eng, 0, a.boot)
ok = True
if rank == 0:
    with Engine(device=local, seed=5) as full:
        full.set_products(products, 29903)
        full.load_group(0, aln)
        fsums, fse = full.within_all(0, B=a.boot)
        fstats = np.concatenate([full.within(0, i)["stats4"] for i in range(len(products))])
    ok &= np.array_equal(stats.cpu().numpy(), fstats)
    ok &= np.allclose(sums, fsums, rtol=1e-13)
    ok &= bool(np.all(np.abs(se[:, :3] / fse[:, :3] - 1) < 0.1))
    print("sharded check:", "OK" if ok else "MISMATCH", "world", world, "SE ratio", np.round(se[:, 2] / fse[:, 2], 3))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)

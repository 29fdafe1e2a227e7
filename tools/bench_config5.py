#!/usr/bin/env python
"""Side measurement (not the bench line): BASELINE config 5 -- ONE within-group alignment of 1,000,000 sequences x
29,903 nt, 10,000 bootstraps, its codon axis split over the ranks, each rank bootstrapping its share
of the replicates, two small NCCL all-gathers.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
        tools/bench_config5.py [--seqs 1000000] [--boot 10000] [--steps 10] [--warmup 2]

The alignment is generated on the device: a rank fills only the columns of its own codon range (the only ones its
kernels read) of a full-pitch buffer -- per-column ancestral base, each codon column polymorphic with probability
0.3 with 1-3 minor alleles of frequency ~ Beta(0.5, 5) differing at one position.  With --check (small n only) rank 0
also runs the whole alignment unsharded on the gathered columns and compares the product sums.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from snpgenie_b200 import multigpu, synth  # noqa: E402
from snpgenie_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seqs", type=int, default=1000000)
ap.add_argument("--boot", type=int, default=10000)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--check", action="store_true")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
products = synth.sars2_products()
L, n = 29903, a.seqs
pitch = (L + 127) // 128 * 128
codons = sum(p.n_codons for p in products)

# global codon axis -> the three alignment columns of every codon (all CDS are '+' and single-segment)
cols = np.concatenate([np.arange(p.starts[0] - 1, p.stops[0]).reshape(-1, 3) for p in products])     # [codons, 3]
first, count = multigpu.shard_plan(codons, rank, world)
mine = cols[first:first + count]
gen = torch.Generator(device=dev)
gen.manual_seed(20261019 + 5)                      # same ancestral row on every rank
anc = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)[torch.randint(0, 4, (pitch,), device=dev, generator=gen)]
d = anc.repeat(n, 1)                               # [n, pitch]; this rank's kernels only ever read its own codon range
rng = np.random.default_rng(1000 + rank)
nt = np.frombuffer(b"ACGT", dtype=np.uint8)
for c in np.flatnonzero(rng.random(count) < 0.3):
    u = torch.rand(n, device=dev)
    acc = 0.0
    for f in rng.beta(0.5, 5.0, size=int(rng.integers(1, 4))):
        col = int(mine[c, rng.integers(0, 3)])
        alt = int(rng.choice(nt[nt != int(anc[col])]))
        d[:, col].masked_fill_((u >= acc) & (u < acc + f), alt)
        acc += f
torch.cuda.synchronize()
dist.barrier()

with Engine(device=local, seed=5, rank=rank, world=world) as eng:
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    eng.set_products(products, L)

    def step():
        eng.load_group_ptr(0, d.data_ptr(), n, L, pitch, device=True)
        return multigpu.within_group_sharded(eng, 0, a.boot)

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.steps):
        sums, se, stats = step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ok = None
    if a.check:
        # every rank's columns onto rank 0's buffer, then the same analysis unsharded
        for r in range(world):
            f, cnt = multigpu.shard_plan(codons, r, world)
            c_lo, c_hi = int(cols[f:f + cnt].min()), int(cols[f:f + cnt].max()) + 1
            part = d[:, c_lo:c_hi].contiguous() if r == rank else torch.empty((n, c_hi - c_lo), dtype=torch.uint8, device=dev)
            dist.broadcast(part, src=r)
            if rank == 0:
                d[:, c_lo:c_hi] = part
        if rank == 0:
            with Engine(device=local, seed=5) as full:
                full.set_products(products, L)
                full.load_group_ptr(0, d.data_ptr(), n, L, pitch, device=True)
                fsums, fse = full.within_all(0, B=a.boot)
            ok = bool(np.allclose(sums, fsums, rtol=1e-12)) and bool(np.all(np.abs(se[:, :3] / fse[:, :3] - 1) < 0.15))
    if rank == 0:
        pairs = codons * (n * (n - 1) // 2)
        print(json.dumps({"workload": f"config5: within-group, {n} seqs x {L} nt, 10 CDS, {a.boot} bootstraps, codon axis and "
                                      f"replicates split over {world} GPUs; alignment resident in HBM",
                          "n_gpus": world, "ms_per_step": float(ms.item()), "codon_pair_comparisons_per_s": pairs / (float(ms.item()) * 1e-3),
                          "bytes_read_per_gpu": int(n * 3 * count), "side_experiment": True, "matches_unsharded": ok,
                          "dN_minus_dS_SE_product0": float(se[0, 2])}))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok in (None, True) else 1)

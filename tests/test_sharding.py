"""N>1 host logic on CPU: shard arithmetic and gather assembly with world_size-2 gloo."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from snpgenie_b200 import multigpu


def test_shard_plan_partitions_exactly():
    for total in (0, 1, 7, 9677, 10000, 10 ** 6 + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [multigpu.shard_plan(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total, B, P, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # per-codon stats: every rank computes only its codon slice
        full_stats = torch.arange(total * 4, dtype=torch.float64).reshape(total, 4) * 0.25
        first, count = multigpu.shard_plan(total, rank, world)
        gathered = multigpu.all_gather_ragged(full_stats[first:first + count].clone(), total)
        ok_stats = torch.equal(gathered, full_stats)
        # replicate sums: every rank computes only its replicate range for all products
        full_rep = torch.arange(P * B * 4, dtype=torch.float64).reshape(P, B, 4)
        b_first, b_count = multigpu.shard_plan(B, rank, world)
        rep = multigpu.all_gather_replicates(full_rep[:, b_first:b_first + b_count].clone(), B)
        ok_rep = torch.equal(rep, full_rep)
        out.put((rank, ok_stats, ok_rep, tuple(gathered.shape), tuple(rep.shape)))
    finally:
        dist.destroy_process_group()


def test_gather_assembly_world2_gloo():
    world, total, B, P = 2, 9677, 1001, 10
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, B, P, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_stats, ok_rep, s1, s2 in res:
        assert ok_stats and ok_rep, rank
        assert s1 == (total, 4) and s2 == (P, B, 4)

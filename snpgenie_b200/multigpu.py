"""Multi-GPU plumbing done by the CALLER with torch.distributed -- the NCCL BASELINE of the exchange.

The product path is inside the library: Engine(rank=, world=).connect_ranks() and then Engine.job(), whose kernels
exchange through peer memory (include/snpgenie_cuda.h, "several GPUs").  What is kept here is the same analysis with
the two gathers done by NCCL all_gather calls between device-pointer ABI calls; tools/exchange_compare.py times one
against the other.  For ONE alignment split over ranks the sequence is

    rank r:  Engine(rank=r, world=w).set_products(...)      # owns a slice of the global codon axis
             load_group(...)                                 # stages only the columns of that slice
             within_stats_device -> local stats4 [count, 4]
    all ranks: all_gather of the slices -> stats4 [total, 4]  (C x 4 doubles, ~0.3 MB at config 2)
    rank r:  bootstrap_all_device over its replicate range  -> rep [P, b_count, 4]
    all ranks: all_gather of the replicate sums             -> rep [P, B, 4]
             product_sums_device / rep_moments_device       -> per-product sums and SEs

Everything here works on any torch.distributed backend: NCCL tensors on the GPU box, gloo tensors
in the CPU tests (tests/test_sharding.py), where the device calls are replaced by host arithmetic.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _abi


def shard_plan(total: int, rank: int, world: int) -> tuple[int, int]:
    """(first, count) of `total` units owned by `rank` -- the library's own split (snpg_shard_plan)."""
    first, count = C.c_int64(), C.c_int64()
    rc = _abi.lib.snpg_shard_plan(total, rank, world, C.byref(first), C.byref(count))
    if rc:
        raise _abi.SnpgError(rc, "snpg_shard_plan")
    return first.value, count.value


def all_gather_ragged(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """Concatenate per-rank slices of a [count_r, ...] tensor split by shard_plan(total, r, world)."""
    world = dist.get_world_size(group)
    counts = [shard_plan(total, r, world)[1] for r in range(world)]
    width = max(counts)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def all_gather_replicates(local_rep: torch.Tensor, B: int, group=None) -> torch.Tensor:
    """local_rep [P, b_count_r, 4] for replicate range shard_plan(B, r, world) -> [P, B, 4]."""
    full = all_gather_ragged(local_rep.transpose(0, 1).contiguous(), B, group)     # [B, P, 4]
    return full.transpose(0, 1).contiguous()


def within_group_sharded(eng, group_id: int, B: int, group=None):
    """Per-product sums [P,4] and SEs [P,6] for one alignment whose codon axis is split over ranks."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    # The library launches on the ctx's own stream and torch/NCCL on theirs: order them explicitly around every
    # hand-over (the library's calls synchronise their stream before they return; torch's side is drained here).
    sync = torch.cuda.synchronize
    first, count, total = eng.shard_range()
    assert (first, count) == shard_plan(total, rank, world)
    dev = torch.device("cuda", torch.cuda.current_device())
    local = torch.zeros((max(count, 1), 4), dtype=torch.float64, device=dev)
    sync()
    eng.within_stats_device(group_id, local.data_ptr())
    stats = all_gather_ragged(local[:count], total, group).contiguous()
    sync()
    sums = eng.product_sums_device(stats.data_ptr())
    se = None
    if B >= 2:
        b_first, b_count = shard_plan(B, rank, world)
        P = len(eng.products)
        rep = torch.zeros((P, max(b_count, 1), 4), dtype=torch.float64, device=dev)
        sync()
        if b_count:
            eng.bootstrap_all_device(stats.data_ptr(), b_first, b_count, rep.data_ptr())
        full = all_gather_replicates(rep[:, :b_count], B, group)
        sync()
        se = eng.rep_moments_device(full.data_ptr(), P, B)
    return sums, se, stats

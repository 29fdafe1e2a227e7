#!/usr/bin/env python
"""Small driver for ncu captures: a few device-resident steps of the config-2 hot path.

    python tools/prof_step.py [--steps 3] [--route spans|fused|tiles|codes] [--p-poly 0.3] [--n 10000]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from snpgenie_b200 import _abi, synth  # noqa: E402
from snpgenie_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--route", default="fused")
ap.add_argument("--n", type=int, default=10000)
ap.add_argument("--boot", type=int, default=10000)
ap.add_argument("--p-poly", type=float, default=0.3)
ap.add_argument("--max-changes", type=int, default=3)
ap.add_argument("--max-alleles", type=int, default=3)
a = ap.parse_args()
products = synth.sars2_products()
L = 29903
aln = synth.make_alignment(a.n, L, products, 20261021, p_poly=a.p_poly, max_changes=a.max_changes, max_alleles=a.max_alleles)
pitch = (L + 127) // 128 * 128
d = torch.zeros((a.n, pitch), dtype=torch.uint8, device="cuda:0")
d[:, :L] = torch.from_numpy(aln).cuda()
with Engine(device=0) as eng:
    eng.set_option(_abi.OPT_KEEP_CODES, 1 if a.route == "codes" else 0)
    eng.set_option(_abi.OPT_FUSED_KERNEL, {"tiles": _abi.FUSED_TMA_TILES, "spans": _abi.FUSED_TMA_SPANS}.get(a.route, _abi.FUSED_LDG_SPANS))
    eng.set_products(products, L)
    for _ in range(a.steps):       # the whole-job call bench.py times (run with SNPG_NO_GRAPH=1 for plain launches)
        sums, se = eng.within_job(0, d.data_ptr(), a.n, L, pitch, True, B=a.boot)
    print("ok", sums[0], se[0][:3], "tile launches", eng.tile_launches, "span-tma launches", eng.span_tma_launches)

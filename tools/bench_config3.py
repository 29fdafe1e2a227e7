#!/usr/bin/env python
"""Side measurement (not the bench line): BASELINE config 3 -- HIV-shaped within-group analysis, 5,000 sequences x
9,719 nt, 9 CDS (overlapping frames, a two-segment product, two '-' strand CDS), 5 % of the bases damaged (codon-aligned
gaps, N, IUPAC), 1,000 bootstraps, sliding windows of 100 codons, step 10, with per-window bootstraps.

    python tools/bench_config3.py [--steps 20] [--warmup 3]
One step = snpg_within_job on the device-resident alignment + the per-codon statistics and windows of every product.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from snpgenie_b200 import synth  # noqa: E402
from snpgenie_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--boot", type=int, default=1000)
a = ap.parse_args()
products = synth.hiv_products()
L, n, W, STEP = 9719, 5000, 100, 10
aln = synth.make_alignment(n, L, products, 20261019 + 3, gap_frac=0.025, n_frac=0.02, iupac_frac=0.005)
pitch = (L + 127) // 128 * 128
d = torch.zeros((n, pitch), dtype=torch.uint8, device="cuda:0")
d[:, :L] = torch.from_numpy(aln).cuda()
torch.cuda.synchronize()
codons = sum(p.n_codons for p in products)
with Engine(device=0) as eng:
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    eng.set_products(products, L)

    def step():
        sums, se = eng.within_job(0, d.data_ptr(), n, L, pitch, True, B=a.boot)
        nwin = 0
        for i, p in enumerate(products):
            st = eng.within(0, i)["stats4"]
            if p.n_codons >= W:
                r = eng.windows(st, W, STEP, B=a.boot)
                nwin += len(r[0]) if isinstance(r, tuple) else 0
        return sums, se, nwin

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(a.steps):
        sums, se, nwin = step()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / a.steps
    t0 = time.perf_counter()
    for _ in range(a.steps):
        eng.within_job(0, d.data_ptr(), n, L, pitch, True, B=a.boot)
    torch.cuda.synchronize()
    ms_job = (time.perf_counter() - t0) * 1e3 / a.steps
    pairs = codons * (n * (n - 1) // 2)
    print(json.dumps({"workload": "config3: within-group, 5000 seqs x 9719 nt, 9 CDS (overlapping, 2-segment, 2 on '-'), 5% damaged bases, "
                                  "1000 bootstraps, windows 100 codons step 10 with per-window bootstraps; alignment resident in HBM",
                      "ms_per_step": ms, "ms_within_job_only": ms_job, "windows_per_step": nwin, "window_bootstrap_reps_per_step": nwin * a.boot,
                      "codon_pair_comparisons_per_s": pairs / (ms * 1e-3), "side_experiment": True, "timing": "host wall clock over the calls"}))

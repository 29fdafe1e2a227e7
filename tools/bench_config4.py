#!/usr/bin/env python
"""Side measurement (not the bench line): BASELINE config 4 -- between-group, 4 groups x 5,000 x 29,903,
all 6 group pairs, 1,000 bootstraps -- through snpg_load_group_device x 4 + snpg_between_all.

    python tools/bench_config4.py [--steps 20] [--warmup 3]
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
ap.add_argument("--calls", action="store_true", help="load the groups with separate calls, then snpg_between_all (default: one snpg_between_job)")
a = ap.parse_args()
products = synth.sars2_products()
L, n, G = 29903, 5000, 4
codons = sum(p.n_codons for p in products)
pitch = (L + 127) // 128 * 128
rng = np.random.default_rng(4)
d_groups = []
for g in range(G):
    aln = synth.make_alignment(n, L, products, 404 + g)
    cols = rng.integers(300, 29500, size=200)                    # group-specific fixed differences
    aln[:, cols] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=200)]
    d = torch.zeros((n, pitch), dtype=torch.uint8, device="cuda:0")
    d[:, :L] = torch.from_numpy(aln).cuda()
    d_groups.append(d)
torch.cuda.synchronize()
with Engine(device=0) as eng:
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    eng.set_products(products, L)

    def step_calls():
        for g in range(G):
            eng.load_group_ptr(g, d_groups[g].data_ptr(), n, L, pitch, device=True)
        return eng.between_all(list(range(G)), B=a.boot)

    def step_job():
        return eng.between_job([d.data_ptr() for d in d_groups], [n] * G, L, [pitch] * G, device=True, B=a.boot)

    step = step_calls if a.calls else step_job

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.steps):
        sums, se = step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    pairs = 6 * codons * n * n
    print(json.dumps({"workload": "config4: between-group, 4 groups x 5000 x 29903, 6 group pairs, 10 CDS, 1000 bootstraps; alignments resident in HBM",
                      "calls": "4 x snpg_load_group_device + snpg_between_all" if a.calls else "one snpg_between_job", "ms_per_step": ms, "codon_pair_comparisons_per_s": pairs / (ms * 1e-3),
                      "bootstrap_reps_per_step": 6 * len(products) * a.boot, "side_experiment": True,
                      "dN_minus_dS_SE_pair0_product0": float(se[0, 0, 2])}))

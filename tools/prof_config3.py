#!/usr/bin/env python
"""Config 3 (HIV-shaped job with sliding-window bootstraps) through ONE whole-job call, a few times: for launch lists
and A/B timings.    python tools/prof_config3.py [--steps 20]    (SNPG_NO_GRAPH=1 for plain launches under ncu)"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from snpgenie_b200 import synth  # noqa: E402
from snpgenie_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20)
a = ap.parse_args()
products = synth.hiv_products()
L, n = 9719, 5000
aln = synth.make_alignment(n, L, products, 20261019 + 3, gap_frac=0.025, n_frac=0.02, iupac_frac=0.005)
pitch = (L + 127) // 128 * 128
dev = torch.device("cuda:0")
d = torch.zeros((n, pitch), dtype=torch.uint8, device=dev)
d[:, :L] = torch.from_numpy(aln).to(dev)
stream = torch.cuda.Stream(device=dev)
with Engine(device=0, seed=3) as e:
    e.set_stream(stream.cuda_stream)
    e.set_products(products, L)
    fn = lambda: e.job(0, d.data_ptr(), n, L, pitch, device=True, B=1000, window=100, step=10, B_win=1000)  # noqa: E731
    for _ in range(4):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.steps):
        out = fn()
    e1.record(stream)
    torch.cuda.synchronize()
    print("config3 ms_per_step %.4f windows %d replays %d" % (e0.elapsed_time(e1) / a.steps, int(out["win_ofs"][-1]), e.graph_replays))

#!/usr/bin/env python
"""Span-histogram kernel time on small launches (a rank's share of a sharded job, a short alignment): n rows of config 2's
alignment, loaded repeatedly with per-kernel timing.   SNPG_SPAN_ROWS=R python tools/chunk_sweep.py [--n 1250 2500 5000]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from snpgenie_b200 import _abi, synth  # noqa: E402
from snpgenie_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, nargs="+", default=[1250, 2500, 5000])
ap.add_argument("--hiv", action="store_true", help="config 3's alignment shape and damage instead of config 2's")
a = ap.parse_args()
products = synth.hiv_products() if a.hiv else synth.sars2_products()
L = 9719 if a.hiv else 29903
kw = dict(gap_frac=0.025, n_frac=0.02, iupac_frac=0.005) if a.hiv else {}
aln = synth.make_alignment(max(a.n), L, products, 20261021, **kw)
pitch = (L + 127) // 128 * 128
d = torch.zeros((max(a.n), pitch), dtype=torch.uint8, device="cuda:0")
d[:, :L] = torch.from_numpy(aln).cuda()
with Engine(device=0) as e:
    e.set_option(_abi.OPT_PROFILE, 1 << _abi.KERNEL_KINDS.index("fused"))
    e.set_products(products, L)
    out = []
    for n in a.n:
        for _ in range(5):
            e.load_group_ptr(0, d.data_ptr(), n, L, pitch, device=True)
        e.profile(reset=True)
        for _ in range(30):
            e.load_group_ptr(0, d.data_ptr(), n, L, pitch, device=True)
        ms, calls = e.profile()["fused"]
        out.append("n=%d %.1f us" % (n, ms / calls * 1e3))
    print("SNPG_SPAN_ROWS=%s %s:" % (os.environ.get("SNPG_SPAN_ROWS", "auto"), "hiv" if a.hiv else "sars2"), "  ".join(out))

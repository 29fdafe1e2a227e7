#!/usr/bin/env python
"""Small end-to-end exercise of every kernel, for compute-sanitizer (memcheck / racecheck)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from snpgenie_b200 import _abi, hostio, synth  # noqa: E402
from snpgenie_b200.engine import Engine  # noqa: E402

products = synth.hiv_products() + [hostio.Product("split", 1, [11, 40, 70], [20, 47, 78]), hostio.Product("end", 1, [9711], [9719])]
aln = synth.make_alignment(333, 9719, products, 5, gap_frac=0.03, n_frac=0.02, iupac_frac=0.01, lower_frac=0.02, u_frac=0.01)
for keep in (True, False):
    with Engine(device=0, keep_codes=keep) as e:
        e.set_option(_abi.OPT_SITE_COUNTS, 1)
        e.set_products(products, 9719)
        e.load_group(0, aln)
        e.load_group(1, aln[:77])
        sums, se = e.within_all(0, B=300)
        for i in range(len(products)):
            e.hist(0, i)
            r = e.within(0, i)
            e.between(0, 1, i)
            if keep:
                e.codon_indices(1, i)
        e.site_counts(0, 9719)
        st = e.within(0, 0)["stats4"]
        e.bootstrap_product(st, 200, keep=(np.arange(len(st)) % 2).astype(np.uint8), want_sums=True)
        e.windows(st, 20, 7, B=100)
        print("keep_codes", keep, "ok", float(sums.sum()))
# sharded contexts and a one-row alignment shorter than a span
for r in range(3):
    with Engine(device=0, rank=r, world=3) as e:
        e.set_products(products, 9719)
        e.load_group(0, aln[:50])
        e.within_all(0, B=50)
with Engine(device=0) as e:
    e.set_products([hostio.Product("one", 1, [2], [4])], 7)
    e.load_group(0, np.frombuffer(b"GATGCAT", dtype=np.uint8).reshape(1, 7).copy())
    e.within_all(0, B=10)
print("done")

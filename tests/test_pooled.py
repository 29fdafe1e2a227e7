"""Row f4: the pooled-mode codon contraction of snpgenie.pl (:6631-6760), snpg_pooled_diffs."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc  # noqa: E402

CODE = {a + b + c: 16 * i + 4 * j + k for i, a in enumerate("ACGT") for j, b in enumerate("ACGT") for k, c in enumerate("ACGT")}


def test_oracle_pooled_by_hand():
    """ATT (3/4) and ACT (1/4) at coverage 100: one nonsynonymous difference (Ile -> Thr), weighted by the number of ATT x ACT
    read pairs over all pairs of reads: 75 * 25 / 4950."""
    f = np.zeros((3, 64))
    f[0, CODE["ATT"]], f[0, CODE["ACT"]] = 0.75, 0.25
    f[1, CODE["GGA"]] = 1.0                                   # a single codon: no differences
    f[2, CODE["TTA"]], f[2, CODE["CTA"]], f[2, CODE["TAA"]] = 0.5, 0.25, 0.25     # Leu/Leu synonymous; pairs with the stop codon count 0
    nd, sd = orc.pooled_diffs(f, np.array([100.0, 40.0, 8.0]))
    assert np.isclose(nd[0], 75 * 25 / 4950, rtol=1e-15) and sd[0] == 0
    assert nd[1] == 0 and sd[1] == 0
    assert nd[2] == 0 and np.isclose(sd[2], (0.5 * 8) * (0.25 * 8) / 28, rtol=1e-15)


@pytest.mark.gpu
def test_pooled_diffs_match_oracle_bitwise():
    from snpgenie_b200 import _abi
    from snpgenie_b200.engine import Engine
    rng = np.random.default_rng(4)
    C = 5000
    f = np.zeros((C, 64))
    for c in range(C):
        k = rng.integers(1, 6)
        codons = rng.choice(64, size=k, replace=False)
        w = rng.random(k) + 0.05
        f[c, codons] = w / w.sum()
    cov = rng.uniform(2, 5000, C)
    want_n, want_s = orc.pooled_diffs(f, cov)
    with Engine(device=0) as eng:
        got_n, got_s = eng.pooled_diffs(f, cov)
        assert np.array_equal(got_n, want_n) and np.array_equal(got_s, want_s)      # same products, same additions, same order
        # the hand-derived case
        g = np.zeros((1, 64))
        g[0, CODE["ATT"]], g[0, CODE["ACT"]] = 0.75, 0.25
        n1, s1 = eng.pooled_diffs(g, np.array([100.0]))
        assert np.isclose(n1[0], 75 * 25 / 4950, rtol=1e-15) and s1[0] == 0
        # where the script dies: two codons at a position covered by one read
        with pytest.raises(_abi.SnpgError) as e:
            eng.pooled_diffs(g, np.array([1.0]))
        assert e.value.code == _abi.E_INPUT and "impossible number of sequenced alleles" in str(e.value)

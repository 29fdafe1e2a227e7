"""Row f2: the statistics of SNPGenie_sliding_windows.R (snpg_sliding_windows_r).

PARITY UNPINNED for the R script itself: there is no R in the authoring container.  The oracle's restatement
(oracle.sliding_windows_r, citing SNPGenie_sliding_windows.R file:line) is pinned here against hand-derived vectors and
scipy's normal distribution; the CUDA path is then compared with the oracle: deterministic columns to 1e-12, bootstrap
columns statistically (the RNG streams differ)."""
import math
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc  # noqa: E402

DET = ("sw_start", "sw_center", "sw_end", "sw_num_replicates", "sw_N_diffs", "sw_S_diffs", "sw_N_sites", "sw_S_sites", "sw_dN", "sw_dS",
       "sw_dNdS", "sw_dN_m_dS")


def _table(C, seed, p_poly=0.4):
    rng = np.random.default_rng(seed)
    t = np.zeros((C, 4))
    t[:, 0] = rng.uniform(1.8, 2.7, C)                  # N_sites
    t[:, 1] = 3 - t[:, 0]                               # S_sites
    poly = rng.random(C) < p_poly
    t[poly, 2] = rng.uniform(0, 0.2, poly.sum())        # N_diffs
    t[poly, 3] = rng.uniform(0, 0.15, poly.sum())       # S_diffs
    return t


# ----------------------------------------------------------------- the oracle against hand-derived vectors (CPU)

def test_oracle_windows_by_hand():
    """Four codons, windows of 2 step 1: sums, ratios and window numbering worked out by hand (R :445-451, :512-519)."""
    t = np.array([[2.0, 1.0, 0.2, 0.1], [2.5, 0.5, 0.0, 0.0], [2.0, 1.0, 0.4, 0.3], [2.25, 0.75, 0.1, 0.0]])
    o = orc.sliding_windows_r(t, 2, 1, B=50)
    assert o["sw_start"].tolist() == [1, 2, 3] and o["sw_end"].tolist() == [2, 3, 4] and o["sw_center"].tolist() == [1.5, 2.5, 3.5]
    assert np.allclose(o["sw_N_sites"], [4.5, 4.5, 4.25]) and np.allclose(o["sw_S_sites"], [1.5, 1.5, 1.75])
    assert np.allclose(o["sw_N_diffs"], [0.2, 0.4, 0.5]) and np.allclose(o["sw_S_diffs"], [0.1, 0.3, 0.3])
    assert np.allclose(o["sw_dN"], [0.2 / 4.5, 0.4 / 4.5, 0.5 / 4.25]) and np.allclose(o["sw_dS"], [0.1 / 1.5, 0.3 / 1.5, 0.3 / 1.75])
    assert np.allclose(o["sw_dNdS"], o["sw_dN"] / o["sw_dS"]) and np.allclose(o["sw_dN_m_dS"], o["sw_dN"] - o["sw_dS"])
    assert (o["sw_num_replicates"] == 50).all()
    # counts partition the replicates; the ASL P values follow R :317-319 and :492-502
    assert np.allclose(o["sw_boot_dN_gt_dS_count"] + o["sw_boot_dN_eq_dS_count"] + o["sw_boot_dN_lt_dS_count"], 50)
    assert np.allclose(o["sw_ASL_dN_gt_dS_P"], o["sw_boot_dN_lt_dS_count"] / 50)
    two = 2 * np.minimum(o["sw_ASL_dN_gt_dS_P"], o["sw_ASL_dN_lt_dS_P"])
    assert np.allclose(o["sw_ASL_dNdS_P"], np.where(two == 0, 1 / 50, two))


def test_oracle_jukes_cantor_and_filter_by_hand():
    t = np.array([[2.0, 1.0, 0.2, 0.1], [2.5, 0.5, 0.1, 0.05], [2.0, 1.0, 0.4, 0.3], [2.25, 0.75, 0.1, 0.0], [2.0, 1.0, 0.0, 0.2]])
    nd = np.array([10, 3, 10, 10, 10])
    o = orc.sliding_windows_r(t, 2, 1, B=20, n_defined=nd, min_defined=6, correction="JC")
    # codon 2 is dropped (R :441); rows are then indexed by codon NUMBERS 1.. (R :445-449): windows are rows (1,2), (2,3), (3,4) of
    # the kept table = codons (1,3), (3,4), (4,5); the loop would go on to i = 4 (max codon_num 5 - 2 + 1), whose rows do not exist
    assert o["sw_start"].tolist() == [1, 3, 4] and o["sw_end"].tolist() == [3, 4, 5]
    dn = 0.6 / 4.0
    assert math.isclose(o["sw_dN"][0], -0.75 * math.log(1 - 4 / 3 * dn), rel_tol=1e-15)
    assert math.isclose(o["sw_dS"][0], -0.75 * math.log(1 - 4 / 3 * (0.4 / 2.0)), rel_tol=1e-15)


def test_oracle_p_values_are_the_normal_tail():
    from scipy.stats import norm
    for z in (0.0, 0.3, 1.0, 1.959963984540054, 3.5, 8.0):
        assert math.isclose(orc._pnorm_two_sided(z), 2 * norm.cdf(-abs(z)), rel_tol=1e-12)
        assert math.isclose(orc._pnorm_two_sided(-z), 2 * norm.cdf(-abs(z)), rel_tol=1e-12)
    assert math.isclose(orc._pnorm_two_sided(1.959963984540054), 0.05, rel_tol=1e-12)       # the textbook value
    assert orc._pnorm_two_sided(float("inf")) == 0.0 and math.isnan(orc._pnorm_two_sided(float("nan")))


def test_oracle_undefined_windows_follow_r_arithmetic():
    """A window without synonymous sites: dS = 0/0 = NaN, every comparison with it is NA (R :314-319), the two-sided ASL stays 1."""
    t = np.array([[3.0, 0.0, 0.1, 0.0], [3.0, 0.0, 0.0, 0.0], [2.0, 1.0, 0.1, 0.1], [2.0, 1.0, 0.0, 0.0]])
    o = orc.sliding_windows_r(t, 2, 2, B=30)
    assert np.isnan(o["sw_dS"][0]) and np.isnan(o["sw_dN_m_dS"][0]) and np.isnan(o["sw_boot_dN_m_dS_SE"][0])
    assert np.isnan(o["sw_boot_dN_gt_dS_count"][0]) and o["sw_ASL_dNdS_P"][0] == 1.0
    assert np.isfinite(o["sw_dS"][1])


# ----------------------------------------------------------------- the CUDA path against the oracle (GPU)

@pytest.mark.gpu
@pytest.mark.parametrize("correction", ["NONE", "JC"])
def test_rwindows_match_oracle(correction):
    from snpgenie_b200.engine import Engine
    t = _table(240, 3)
    nd = np.full(240, 50, dtype=np.uint32)
    B = 4000
    want = orc.sliding_windows_r(t, 40, 5, B=B, n_defined=nd, min_defined=6, correction=correction, seed=11)
    with Engine(device=0, seed=9) as eng:
        got = eng.sliding_windows_r(t, 40, 5, B=B, n_defined=nd, min_defined=6, correction=correction)
        again = eng.sliding_windows_r(t, 40, 5, B=B, n_defined=nd, min_defined=6, correction=correction)
    assert len(got["sw_start"]) == len(want["sw_start"]) == (240 - 40) // 5 + 1
    for k in DET:
        assert np.allclose(got[k], want[k], rtol=1e-12, atol=0), k
    # bootstrap columns: two different RNG streams -> relative SE of an SD over B replicates ~ 1/sqrt(2B); 6 sigma
    tol = 6 / math.sqrt(2 * B) * 1.5
    for k in ("sw_boot_dN_SE", "sw_boot_dS_SE", "sw_boot_dN_m_dS_SE"):
        assert np.all(np.abs(got[k] / want[k] - 1) < tol), (k, np.max(np.abs(got[k] / want[k] - 1)))
    fin = np.isfinite(want["sw_boot_dN_over_dS_SE"]) & np.isfinite(got["sw_boot_dN_over_dS_SE"])
    assert fin.mean() > 0.9
    # the P values are functions of the estimate and the SE
    from scipy.stats import norm
    for est, se, p in (("sw_dN_m_dS", "sw_boot_dN_m_dS_SE", "sw_boot_dN_m_dS_P"), ("sw_dNdS", "sw_boot_dN_over_dS_SE", "sw_boot_dN_over_dS_P")):
        ok = np.isfinite(got[se]) & (got[se] > 0)
        assert np.allclose(got[p][ok], 2 * norm.cdf(-np.abs(got[est][ok] / got[se][ok])), rtol=1e-10, atol=1e-300), p
    # counts and achieved significance levels (R :314-319, :492-502)
    tot = got["sw_boot_dN_gt_dS_count"] + got["sw_boot_dN_eq_dS_count"] + got["sw_boot_dN_lt_dS_count"]
    assert np.array_equal(tot, np.full_like(tot, B))
    assert np.allclose(got["sw_ASL_dN_gt_dS_P"], got["sw_boot_dN_lt_dS_count"] / B) and np.allclose(got["sw_ASL_dN_lt_dS_P"], got["sw_boot_dN_gt_dS_count"] / B)
    two = 2 * np.minimum(got["sw_ASL_dN_gt_dS_P"], got["sw_ASL_dN_lt_dS_P"])
    assert np.allclose(got["sw_ASL_dNdS_P"], np.where(two == 0, 1 / B, two))
    assert np.all(np.abs(got["sw_ASL_dN_gt_dS_P"] - want["sw_ASL_dN_gt_dS_P"]) < 6 * 0.5 / math.sqrt(B) + 1e-9)
    # a second call draws a new stream: same deterministic columns, SEs close but not identical
    assert np.array_equal(again["sw_dN"], got["sw_dN"]) and not np.array_equal(again["sw_boot_dN_SE"], got["sw_boot_dN_SE"])


@pytest.mark.gpu
def test_rwindows_filter_numbering_and_undefined():
    from snpgenie_b200.engine import Engine
    t = np.array([[2.0, 1.0, 0.2, 0.1], [2.5, 0.5, 0.1, 0.05], [2.0, 1.0, 0.4, 0.3], [2.25, 0.75, 0.1, 0.0], [2.0, 1.0, 0.0, 0.2]])
    nd = np.array([10, 3, 10, 10, 10], dtype=np.uint32)
    with Engine(device=0) as eng:
        got = eng.sliding_windows_r(t, 2, 1, B=200, n_defined=nd, min_defined=6, correction="JC")
        want = orc.sliding_windows_r(t, 2, 1, B=200, n_defined=nd, min_defined=6, correction="JC")
        for k in DET:
            assert np.allclose(got[k], want[k], rtol=1e-12, equal_nan=True), k
        assert got["sw_start"].tolist() == [1, 3, 4]
        # a window without synonymous sites: NaN as in R, NA counts, two-sided ASL 1
        u = np.array([[3.0, 0.0, 0.1, 0.0], [3.0, 0.0, 0.0, 0.0], [2.0, 1.0, 0.1, 0.1], [2.0, 1.0, 0.0, 0.0]])
        g = eng.sliding_windows_r(u, 2, 2, B=64)
        assert np.isnan(g["sw_dS"][0]) and np.isnan(g["sw_boot_dN_m_dS_SE"][0]) and np.isnan(g["sw_boot_dN_gt_dS_count"][0])
        assert g["sw_ASL_dNdS_P"][0] == 1.0 and np.isfinite(g["sw_dS"][1])
        # fewer kept codons than the window: no windows (R :444, :537-542)
        assert len(eng.sliding_windows_r(t, 5, 1, B=10, n_defined=nd, min_defined=6)["sw_start"]) == 0
        # identical replicates: SE exactly 0, Z infinite, P 0
        c = np.tile(np.array([[2.0, 1.0, 0.25, 0.1]]), (6, 1))
        g = eng.sliding_windows_r(c, 3, 1, B=100)
        assert np.all(g["sw_boot_dN_SE"] == 0) and np.all(g["sw_boot_dN_m_dS_SE"] == 0) and np.all(g["sw_boot_dN_m_dS_P"] == 0)
        # ... and 0 / 0 where the estimate is 0 as well: NaN, as R's pnorm(NaN)
        c[:, 3] = 0.125
        g = eng.sliding_windows_r(c, 3, 1, B=100)
        assert np.all(g["sw_dN_m_dS"] == 0) and np.all(np.isnan(g["sw_boot_dN_m_dS_P"]))

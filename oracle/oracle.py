"""ctypes wrapper around oracle/liboracle.so (the CPU restatement).

TEST INFRASTRUCTURE ONLY -- see the header of ng_oracle.c.  Importable from
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never from snpgenie_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

UNDEF, OTHER = 64, 65


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "ng_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_product_positions.restype = C.c_int64
        _LIB.orc_windows.restype = C.c_int64
        _LIB.orc_sd.restype = C.c_double
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def tables():
    """(sites [64,2], diffs [2,64,64]) regenerated from first principles."""
    sites = np.zeros((64, 2), dtype=np.float64)
    diffs = np.zeros((2, 64, 64), dtype=np.float64)
    lib().orc_build_tables(_p(sites, C.c_double), _p(diffs, C.c_double))
    return sites, diffs


def product_positions(starts, stops, strand: int, aln_len: int) -> np.ndarray:
    st = np.ascontiguousarray(starts, dtype=np.int64)
    sp = np.ascontiguousarray(stops, dtype=np.int64)
    total = lib().orc_product_positions(C.c_int(len(st)), _p(st, C.c_int64), _p(sp, C.c_int64),
                                        C.c_int(strand), C.c_int64(aln_len), None)
    if total < 0:
        raise ValueError("bad product coordinates (%d)" % total)
    pos = np.zeros(total, dtype=np.int64)
    lib().orc_product_positions(C.c_int(len(st)), _p(st, C.c_int64), _p(sp, C.c_int64),
                                C.c_int(strand), C.c_int64(aln_len), _p(pos, C.c_int64))
    return pos


def extract(aln: np.ndarray, pos: np.ndarray, strand: int):
    """(raw [n,C,3] normalised chars, codes [n,C]) for one product."""
    aln = np.ascontiguousarray(aln, dtype=np.uint8)
    n, _ = aln.shape
    ncodon = len(pos) // 3
    raw = np.zeros((n, ncodon, 3), dtype=np.uint8)
    codes = np.zeros((n, ncodon), dtype=np.uint8)
    lib().orc_extract(_p(aln, C.c_uint8), C.c_int64(n), C.c_int64(aln.strides[0]), _p(pos, C.c_int64),
                      C.c_int64(ncodon), C.c_int(strand), _p(raw, C.c_uint8), _p(codes, C.c_uint8))
    return raw, codes


def hist(codes: np.ndarray) -> np.ndarray:
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    n, ncodon = codes.shape
    h = np.zeros((ncodon, 66), dtype=np.uint32)
    lib().orc_hist(_p(codes, C.c_uint8), C.c_int64(n), C.c_int64(ncodon), _p(h, C.c_uint32))
    return h


def other_distinct(raw: np.ndarray, codes: np.ndarray) -> np.ndarray:
    """Per codon: 1 if >= 2 distinct 'other' strings occur (numpy; small cases)."""
    n, ncodon, _ = raw.shape
    key = (raw[:, :, 0].astype(np.uint32) << 16) | (raw[:, :, 1].astype(np.uint32) << 8) | raw[:, :, 2]
    out = np.zeros(ncodon, dtype=np.uint8)
    for c in np.flatnonzero((codes == OTHER).any(axis=0)):
        out[c] = len(np.unique(key[codes[:, c] == OTHER, c])) >= 2
    return out


def _alloc(ncodon):
    return dict(
        poly=np.zeros(ncodon, dtype=np.uint8), comparisons=np.zeros(ncodon, dtype=np.float64),
        N_sites=np.zeros(ncodon), S_sites=np.zeros(ncodon), N_diffs=np.zeros(ncodon), S_diffs=np.zeros(ncodon),
    )


def within_pairloop(raw: np.ndarray, sites=None, diffs=None) -> dict:
    if sites is None:
        sites, diffs = tables()
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    n, ncodon, _ = raw.shape
    r = _alloc(ncodon)
    r["n_defined"] = np.zeros(ncodon, dtype=np.uint32)
    r["first_key"] = np.zeros(ncodon, dtype=np.uint32)
    lib().orc_within_pairloop(_p(raw, C.c_uint8), C.c_int64(n), C.c_int64(ncodon), _p(sites, C.c_double), _p(diffs, C.c_double),
                              _p(r["poly"], C.c_uint8), _p(r["n_defined"], C.c_uint32), _p(r["comparisons"], C.c_double),
                              _p(r["N_sites"], C.c_double), _p(r["S_sites"], C.c_double), _p(r["N_diffs"], C.c_double),
                              _p(r["S_diffs"], C.c_double), _p(r["first_key"], C.c_uint32))
    return r


def between_pairloop(raw_a: np.ndarray, raw_b: np.ndarray, sites=None, diffs=None) -> dict:
    if sites is None:
        sites, diffs = tables()
    raw_a = np.ascontiguousarray(raw_a, dtype=np.uint8)
    raw_b = np.ascontiguousarray(raw_b, dtype=np.uint8)
    na, ncodon, _ = raw_a.shape
    nb = raw_b.shape[0]
    r = _alloc(ncodon)
    r["ndef_a"] = np.zeros(ncodon, dtype=np.uint32)
    r["ndef_b"] = np.zeros(ncodon, dtype=np.uint32)
    r["first_key"] = np.zeros(ncodon, dtype=np.uint32)
    lib().orc_between_pairloop(_p(raw_a, C.c_uint8), C.c_int64(na), _p(raw_b, C.c_uint8), C.c_int64(nb), C.c_int64(ncodon),
                               _p(sites, C.c_double), _p(diffs, C.c_double),
                               _p(r["poly"], C.c_uint8), _p(r["ndef_a"], C.c_uint32), _p(r["ndef_b"], C.c_uint32),
                               _p(r["comparisons"], C.c_double), _p(r["N_sites"], C.c_double), _p(r["S_sites"], C.c_double),
                               _p(r["N_diffs"], C.c_double), _p(r["S_diffs"], C.c_double), _p(r["first_key"], C.c_uint32))
    return r


def within_hist(h: np.ndarray, odist=None, sites=None, diffs=None) -> dict:
    if sites is None:
        sites, diffs = tables()
    h = np.ascontiguousarray(h, dtype=np.uint32)
    ncodon = h.shape[0]
    r = _alloc(ncodon)
    r["n_defined"] = np.zeros(ncodon, dtype=np.uint32)
    od = None if odist is None else _p(np.ascontiguousarray(odist, dtype=np.uint8), C.c_uint8)
    lib().orc_within_hist(_p(h, C.c_uint32), od, C.c_int64(ncodon), _p(sites, C.c_double), _p(diffs, C.c_double),
                          _p(r["poly"], C.c_uint8), _p(r["n_defined"], C.c_uint32), _p(r["comparisons"], C.c_double),
                          _p(r["N_sites"], C.c_double), _p(r["S_sites"], C.c_double), _p(r["N_diffs"], C.c_double),
                          _p(r["S_diffs"], C.c_double))
    return r


def between_hist(ha: np.ndarray, hb: np.ndarray, odist=None, sites=None, diffs=None) -> dict:
    if sites is None:
        sites, diffs = tables()
    ha = np.ascontiguousarray(ha, dtype=np.uint32)
    hb = np.ascontiguousarray(hb, dtype=np.uint32)
    ncodon = ha.shape[0]
    r = _alloc(ncodon)
    r["ndef_a"] = np.zeros(ncodon, dtype=np.uint32)
    r["ndef_b"] = np.zeros(ncodon, dtype=np.uint32)
    od = None if odist is None else _p(np.ascontiguousarray(odist, dtype=np.uint8), C.c_uint8)
    lib().orc_between_hist(_p(ha, C.c_uint32), _p(hb, C.c_uint32), od, C.c_int64(ncodon), _p(sites, C.c_double), _p(diffs, C.c_double),
                           _p(r["poly"], C.c_uint8), _p(r["ndef_a"], C.c_uint32), _p(r["ndef_b"], C.c_uint32),
                           _p(r["comparisons"], C.c_double), _p(r["N_sites"], C.c_double), _p(r["S_sites"], C.c_double),
                           _p(r["N_diffs"], C.c_double), _p(r["S_diffs"], C.c_double))
    return r


def bootstrap_product(stats4: np.ndarray, B: int, seed: int = 1):
    """(sums [B,4], se [6] = SD of dN, dS, dN-dS, JC dN, JC dS, JC dN-dS)."""
    stats4 = np.ascontiguousarray(stats4, dtype=np.float64)
    sums = np.zeros((B, 4))
    se = np.zeros(6)
    lib().orc_bootstrap_product(_p(stats4, C.c_double), C.c_int64(stats4.shape[0]), C.c_int64(B), C.c_uint64(seed),
                                _p(sums, C.c_double), _p(se, C.c_double))
    return sums, se


def windows(stats4: np.ndarray, W: int, step: int, B: int = 0, seed: int = 1):
    """(win_sums [nwin,4], win_se [nwin,3] or None)."""
    stats4 = np.ascontiguousarray(stats4, dtype=np.float64)
    ncodon = stats4.shape[0]
    if W > ncodon:
        return np.zeros((0, 4)), None
    nwin = (ncodon - W) // step + 1
    sums = np.zeros((nwin, 4))
    se = np.zeros((nwin, 3)) if B > 1 else None
    lib().orc_windows(_p(stats4, C.c_double), C.c_int64(ncodon), C.c_int64(W), C.c_int64(step), C.c_int64(B), C.c_uint64(seed),
                      _p(sums, C.c_double), None if se is None else _p(se, C.c_double))
    return sums, se


def max_threads() -> int:
    return int(lib().orc_max_threads())


def use_all_cores() -> int:
    """Pin the OpenMP team to every core this process may run on (ignores OMP_NUM_THREADS)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().orc_set_threads(C.c_int(n))
    return max_threads()


def site_counts(aln: np.ndarray) -> np.ndarray:
    """[L,4] counts of A,C,G,T per alignment column after uc and U->T (snpgenie_within_group.pl:1222-1237)."""
    a = np.asarray(aln, dtype=np.uint8).copy()
    lower = (a >= ord("a")) & (a <= ord("z"))
    a[lower] -= 32
    a[a == ord("U")] = ord("T")
    return np.stack([(a == ord(ch)).sum(axis=0) for ch in "ACGT"], axis=1).astype(np.uint32)


# ---------------------------------------------------------------------------------------------------------------
# "Next" row f2: SNPGenie_sliding_windows.R restated in numpy.  PARITY UNPINNED: there is no R in the authoring
# container, so this restatement could not be run against the script itself; tests/test_rwindows.py pins its
# deterministic columns against hand-derived vectors and its P values against scipy's normal distribution.
# ---------------------------------------------------------------------------------------------------------------
RWIN_NAMES = ("sw_start", "sw_center", "sw_end", "sw_num_replicates", "sw_N_diffs", "sw_S_diffs", "sw_N_sites", "sw_S_sites", "sw_dN", "sw_dS",
              "sw_dNdS", "sw_dN_m_dS", "sw_boot_dN_SE", "sw_boot_dS_SE", "sw_boot_dN_over_dS_SE", "sw_boot_dN_over_dS_P", "sw_boot_dN_m_dS_SE",
              "sw_boot_dN_m_dS_P", "sw_boot_dN_gt_dS_count", "sw_boot_dN_eq_dS_count", "sw_boot_dN_lt_dS_count", "sw_ASL_dN_gt_dS_P",
              "sw_ASL_dN_lt_dS_P", "sw_ASL_dNdS_P")


def _pnorm_two_sided(z):
    """2 * pnorm(-abs(z)) (SNPGenie_sliding_windows.R:304, :311)."""
    from math import erfc, sqrt
    return float("nan") if z != z else erfc(abs(z) / sqrt(2.0))


def sliding_windows_r(stats4: np.ndarray, window: int, step: int, B: int = 1000, n_defined=None, min_defined: int = 6, codon_num=None,
                      numerator: str = "N", denominator: str = "S", correction: str = "NONE", seed: int = 1) -> dict:
    """SNPGenie_sliding_windows.R:241-321 (bootstrap function), :326-402 (its Jukes-Cantor twin), :441-548 (the window
    loop).  stats4 [C,4] = N_sites, S_sites, N_diffs, S_diffs.  One set of resamples serves the script's four boot() calls."""
    stats4 = np.asarray(stats4, dtype=np.float64)
    C_ = stats4.shape[0]
    cn = np.arange(1, C_ + 1) if codon_num is None else np.asarray(codon_num, dtype=np.int64)
    rows = np.arange(C_)
    if n_defined is not None:
        rows = rows[np.asarray(n_defined)[rows] >= min_defined]                 # :441 filter(num_defined_seqs >= MIN_DEFINED_CODONS)
    rows = rows[np.argsort(cn[rows], kind="stable")]                            # :442 arrange(codon_num)
    tab, num = stats4[rows], cn[rows]
    ns, ds = {"N": 0, "S": 1}[numerator], {"N": 0, "S": 1}[denominator]
    rng = np.random.default_rng(seed)
    out = {k: [] for k in RWIN_NAMES}
    if len(rows) < window:                                                      # :444
        return {k: np.array(v) for k, v in out.items()}

    def stat(x):                                                                # :252-276: the four statistics of summed columns
        with np.errstate(divide="ignore", invalid="ignore"):
            dn, dss = np.float64(x[..., 2 + ns]) / x[..., ns], np.float64(x[..., 2 + ds]) / x[..., ds]
            if correction == "JC":                                              # :329, :336
                dn, dss = -0.75 * np.log(1 - (4.0 / 3.0) * dn), -0.75 * np.log(1 - (4.0 / 3.0) * dss)
            return dn, dss, dn / dss, dn - dss

    for i in range(int(num.min()), int(num.max()) - window + 2, step):          # :445 seq(min, max - W + 1, by = STEP)
        if i < 1 or i - 1 + window > len(rows):
            break                                                               # (rows of NA: the script stops with an error)
        w = tab[i - 1:i - 1 + window]                                           # :449 rows i .. i + W - 1
        lo, hi = int(num[i - 1:i - 1 + window].min()), int(num[i - 1:i - 1 + window].max())
        sums = np.zeros(4)
        for r in w:                                                             # sum() in row order
            sums += r
        dn, dss, ratio, diff = stat(sums)
        idx = rng.integers(0, window, size=(B, window))                         # boot(): W draws with replacement
        t = stat(w[idx].sum(axis=1))
        with np.errstate(invalid="ignore"):
            se = [np.std(x, ddof=1) if np.all(np.isfinite(x)) else np.nan for x in t]      # :283-299 sd(boot$t)
            se = [0.0 if (np.all(x == x[0]) and np.isfinite(x[0])) else s for x, s in zip(t, se)]
        tm = t[3]
        if np.any(np.isnan(tm)):                                                # :314-319: sums of NA comparisons are NA
            gt = eq = lt = p_gt = p_lt = np.nan
            asl = 1.0                                                           # :492-498
        else:
            gt, eq, lt = float((tm > 0).sum()), float((tm == 0).sum()), float((tm < 0).sum())
            p_gt, p_lt = lt / (gt + eq + lt), gt / (gt + eq + lt)
            asl = 2 * p_gt if p_gt < p_lt else 2 * p_lt
        if asl == 0:
            asl = 1.0 / B                                                       # :500-502
        with np.errstate(divide="ignore", invalid="ignore"):
            z_ratio, z_diff = np.float64(ratio) / se[2], np.float64(diff) / se[3]
        vals = (lo, (lo + hi) / 2.0, hi, B, sums[2 + ns], sums[2 + ds], sums[ns], sums[ds], dn, dss, ratio, diff, se[0], se[1], se[2],
                _pnorm_two_sided(z_ratio), se[3], _pnorm_two_sided(z_diff), gt, eq, lt, p_gt, p_lt, asl)
        for k, v in zip(RWIN_NAMES, vals):
            out[k].append(float(v))
    return {k: np.array(v) for k, v in out.items()}


# ---------------------------------------------------------------------------------------------------------------
# "Next" row f4: the pooled-mode codon contraction of snpgenie.pl (:6631-6760), literally.  The tables are the pinned ones
# (tests/golden/ng_tables.tsv = a dump of the reference's own subs); the loop itself is checked against hand-derived values.
# ---------------------------------------------------------------------------------------------------------------
def pooled_diffs(freq: np.ndarray, coverage: np.ndarray, diffs=None):
    """freq [C,64] codon frequencies, coverage [C] -> (N_diffs [C], S_diffs [C]) as snpgenie.pl:6871 prints them."""
    if diffs is None:
        _, diffs = tables()
    freq = np.asarray(freq, dtype=np.float64)
    nd, sd = np.zeros(freq.shape[0]), np.zeros(freq.shape[0])
    for c in range(freq.shape[0]):
        cov = float(coverage[c])
        comps = ((cov * cov) - cov) / 2                                     # :6647-6649
        here = [i for i in range(64) if freq[c, i] > 0]                     # :6644 sort(keys): alphabetical = code order
        for a in range(len(here)):
            for b in range(a + 1, len(here)):
                i, j = here[a], here[b]
                num = (freq[c, i] * cov) * (freq[c, j] * cov)               # :6683-6685
                w = num / comps                                             # :6752 (the script dies when comps == 0)
                nd[c] += diffs[0, i, j] * w                                 # :6753, :6760
                sd[c] += diffs[1, i, j] * w
    return nd, sd

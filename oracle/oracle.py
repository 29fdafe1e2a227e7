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

"""Host-side logic of the ctypes mirror that needs no GPU: Engine.job / Engine.between_job build their argument
structs and output arrays once per argument tuple (the library call is replaced by a recorder here)."""
import ctypes as C

import numpy as np

import snpgenie_b200.engine as eng
from snpgenie_b200 import _abi, synth


class _Recorder:
    def __init__(self):
        self.calls = []

    def snpg_within_job_ex(self, ctx, group, ptr, device, n, length, stride, spec_ref):
        self.calls.append(("job", group, ptr, device, n, length, stride, C.addressof(spec_ref._obj)))
        return 0

    def snpg_between_job(self, ctx, n, ptrs, device, n_seqs, length, strides, tot, grp, B, sums, se):
        self.calls.append(("between", n, [ptrs[i] for i in range(n)], device, [n_seqs[i] for i in range(n)], length, B))
        return 0

    def snpg_window_plan(self, ctx, W, step, ofs):
        ofs[len(synth.sars2_products())] = 7
        return 0


class _HostOnlyEngine(eng.Engine):
    peer_connected = False
    team_size = 1

    def __init__(self):          # no snpg_create: only the Python side is exercised
        self.products = synth.sars2_products()
        self.world = 1
        self._ctx = None
        self._pinned = {}
        self._jobs = {}

    def shard_range(self):
        return 0, 9677, 9677

    def close(self):
        pass


def test_job_struct_is_built_once_per_argument_tuple(monkeypatch):
    rec = _Recorder()
    monkeypatch.setattr(eng._abi, "lib", rec)
    e = _HostOnlyEngine()
    a = e.job(0, 4096, 100, 29903, 29952, device=True, B=50, pinned=False)
    b = e.job(0, 4096, 100, 29903, 29952, device=True, B=50, pinned=False)
    assert a is b and len(e._jobs) == 1
    assert rec.calls[0] == rec.calls[1]                                    # same struct, same arguments
    assert set(a) == {"prod_sums", "prod_se", "prod_undefined", "poly", "n_defined", "comparisons", "stats4", "conserved"}
    assert a["stats4"].shape == (9677, 4) and a["comparisons"].dtype == np.uint64
    c = e.job(0, 4096, 100, 29903, 29952, device=True, B=0, pinned=False)    # another tuple: another struct
    assert c is not a and "prod_se" not in c and len(e._jobs) == 2 and rec.calls[2][-1] != rec.calls[0][-1]
    w = e.job(0, 4096, 100, 29903, 29952, device=True, B=50, window=100, step=10, B_win=20, pinned=False)
    assert w["win_sums"].shape == (7, 4) and w["win_se"].shape == (7, 3)
    e._jobs.clear()          # what set_products / peer_connect do: the axis may have changed
    d = e.job(0, 4096, 100, 29903, 29952, device=True, B=50, pinned=False)
    assert d is not a and np.shares_memory(d["stats4"], a["stats4"])       # (the arrays themselves are kept by shape)


def test_between_job_reuses_its_arrays(monkeypatch):
    rec = _Recorder()
    monkeypatch.setattr(eng._abi, "lib", rec)
    e = _HostOnlyEngine()
    s1, se1 = e.between_job([16, 32, 48], [5, 6, 7], 29903, [29952] * 3, device=True, B=10)
    s2, se2 = e.between_job([16, 32, 48], [5, 6, 7], 29903, [29952] * 3, device=True, B=10)
    assert s1 is s2 and se1 is se2 and s1.shape == (3, 10, 4) and se1.shape == (3, 10, 6)
    assert rec.calls[0] == rec.calls[1] == ("between", 3, [16, 32, 48], 1, [5, 6, 7], 29903, 10)
    s3, se3 = e.between_job([16, 32], [5, 6], 29903, [29952] * 2, device=True, B=0)
    assert s3.shape == (1, 10, 4) and se3 is None

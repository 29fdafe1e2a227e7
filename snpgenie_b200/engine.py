"""Host-side mirror of the reference's within-/between-group engine, over the C ABI.

`Engine` owns one snpg_ctx (one GPU).  Method names follow the reference's
vocabulary (products, groups, codons, windows, bootstraps); every compute call
goes through libsnpgenie_cuda.so -- there is no Python or CPU implementation
behind these methods.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from .hostio import Product, products_to_csr

STATS = ("N_sites", "S_sites", "N_diffs", "S_diffs")
CODONS = [a + b + c for a in "ACGT" for b in "ACGT" for c in "ACGT"]


def _ptr(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def ng_tables():
    """(sites [64,2], diffs [2,64,64]) as the kernels use them (host-built, no GPU needed)."""
    sites = np.zeros((64, 2))
    diffs = np.zeros((2, 64, 64))
    rc = _abi.lib.snpg_get_tables(_ptr(sites, C.c_double), _ptr(diffs, C.c_double))
    if rc:
        raise _abi.SnpgError(rc, "snpg_get_tables")
    return sites, diffs


class Engine:
    def __init__(self, device: int = 0, seed: int = 20261019, rank: int = 0, world: int = 1, keep_codes: bool = False,
                 gpus: int | list[int] | None = None):
        """keep_codes=True materialises the packed codon matrix (K1+K2 route: codon_indices(), exact
        between-group first-defined rule); the default builds histograms with the fused kernel.
        gpus=N (or a list of device ids): ONE engine over N GPUs of this process (snpg_create_multi) -- codon ranges and
        bootstrap replicates are split over them and exchanged through peer memory.  rank/world: this engine is one
        rank of a multi-process job (one process per GPU); connect the ranks with peer_export()/peer_connect()."""
        self._ctx = C.c_void_p()
        if gpus is not None:
            devs = list(range(gpus)) if isinstance(gpus, int) else list(gpus)
            rc = _abi.lib.snpg_create_multi(C.byref(self._ctx), len(devs), (C.c_int * len(devs))(*devs), seed)
        else:
            rc = _abi.lib.snpg_create(C.byref(self._ctx), device, seed)
        if rc:
            raise _abi.SnpgError(rc, (_abi.lib.snpg_last_error(None) or b"").decode())
        self.products: list[Product] = []
        self.rank, self.world = rank, world
        self._pinned = {}
        self._jobs = {}         # job(): the filled-in snpg_job struct and output arrays of every argument tuple seen
        if world > 1:
            self._ck(_abi.lib.snpg_set_shard(self._ctx, rank, world))
        if keep_codes:
            self.set_option(_abi.OPT_KEEP_CODES, 1)

    # -- plumbing ---------------------------------------------------------
    def _ck(self, rc):
        if rc:
            raise _abi.SnpgError(rc, (_abi.lib.snpg_last_error(self._ctx) or b"").decode())

    def close(self):
        if self._ctx:
            _abi.lib.snpg_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(_abi.lib.snpg_launch_count(self._ctx))

    @property
    def tile_launches(self) -> int:
        """Launches of the TMA-tiled histogram kernel (SNPG_FUSED_TMA_TILES)."""
        return int(_abi.lib.snpg_tile_launch_count(self._ctx))

    @property
    def span_tma_launches(self) -> int:
        """Launches of the TMA-fed span histogram kernel (SNPG_FUSED_TMA_SPANS)."""
        return int(_abi.lib.snpg_span_tma_launch_count(self._ctx))

    @property
    def graph_replays(self) -> int:
        """How often within_job ran as a replayed CUDA graph."""
        return int(_abi.lib.snpg_graph_replays(self._ctx))

    def set_option(self, option: int, value: int):
        self._ck(_abi.lib.snpg_set_option(self._ctx, option, value))

    def set_stream(self, cuda_stream_ptr: int | None):
        """Launch on a caller-owned stream.  None restores the engine's own stream; 0 -- the handle torch reports for the
        legacy default stream -- is passed on as cudaStreamLegacy (the ABI reads NULL as "own stream")."""
        if cuda_stream_ptr is not None and cuda_stream_ptr == 0:
            cuda_stream_ptr = 1
        self._ck(_abi.lib.snpg_set_stream(self._ctx, cuda_stream_ptr))

    # -- several GPUs -----------------------------------------------------
    @property
    def team_size(self) -> int:
        return int(_abi.lib.snpg_team_size(self._ctx))

    def peer_export(self, max_bootstraps: int) -> bytes:
        """This rank's exchange handle (call after set_products); gather all ranks' handles in rank order."""
        buf = C.create_string_buffer(_abi.PEER_HANDLE_BYTES)
        self._ck(_abi.lib.snpg_peer_export(self._ctx, max_bootstraps, buf))
        return buf.raw

    def peer_connect(self, handles: list[bytes]):
        blob = b"".join(handles)
        self._ck(_abi.lib.snpg_peer_connect(self._ctx, blob, len(handles)))
        self._jobs.clear()

    @property
    def peer_connected(self) -> bool:
        return bool(_abi.lib.snpg_peer_connected(self._ctx))

    def connect_ranks(self, max_bootstraps: int, group=None):
        """Export + all-gather of the handles over torch.distributed (any backend) + connect."""
        import torch.distributed as dist
        mine = self.peer_export(max_bootstraps)
        handles = [None] * dist.get_world_size(group)
        dist.all_gather_object(handles, mine, group=group)
        self.peer_connect(handles)

    def profile(self, reset: bool = True) -> dict:
        """{kernel kind: (milliseconds, launches)} accumulated since the last reset (kinds selected by OPT_PROFILE)."""
        ms = np.zeros(len(_abi.KERNEL_KINDS))
        calls = np.zeros(len(_abi.KERNEL_KINDS), dtype=np.int64)
        self._ck(_abi.lib.snpg_get_profile(self._ctx, _ptr(ms, C.c_double), _ptr(calls, C.c_int64), int(reset)))
        return {k: (float(ms[i]), int(calls[i])) for i, k in enumerate(_abi.KERNEL_KINDS)}

    # -- annotation ---------------------------------------------------------
    def set_products(self, products: list[Product], aln_len: int):
        seg_ofs, starts, stops, strand = products_to_csr(products)
        self._ck(_abi.lib.snpg_set_products(self._ctx, len(products), _ptr(seg_ofs, C.c_int64), _ptr(starts, C.c_int64),
                                            _ptr(stops, C.c_int64), _ptr(strand, C.c_int8), aln_len))
        self.products = list(products)
        self._jobs.clear()

    def n_codons(self, product: int) -> int:
        out = C.c_int64()
        self._ck(_abi.lib.snpg_n_codons(self._ctx, product, C.byref(out)))
        return out.value

    def shard_range(self):
        f, c, t = C.c_int64(), C.c_int64(), C.c_int64()
        self._ck(_abi.lib.snpg_shard_range(self._ctx, C.byref(f), C.byref(c), C.byref(t)))
        return f.value, c.value, t.value

    def local_codons(self, product: int) -> int:
        """Codons of `product` inside this ctx's shard (== n_codons when world == 1)."""
        first, count, _ = self.shard_range()
        ofs = np.cumsum([0] + [self.n_codons(p) for p in range(len(self.products))])
        lo = min(max(ofs[product] - first, 0), count)
        hi = min(max(ofs[product + 1] - first, 0), count)
        return int(hi - lo)

    # -- ingest -------------------------------------------------------------
    def load_group(self, group: int, aln: np.ndarray):
        """aln: uint8 [n_seqs, aln_len] host array (raw FASTA residues)."""
        assert aln.dtype == np.uint8 and aln.ndim == 2 and aln.strides[1] == 1
        self._ck(_abi.lib.snpg_load_group(self._ctx, group, aln.ctypes.data, aln.shape[0], aln.shape[1], aln.strides[0]))

    def load_group_ptr(self, group: int, ptr: int, n_seqs: int, aln_len: int, row_stride: int, device: bool):
        fn = _abi.lib.snpg_load_group_device if device else _abi.lib.snpg_load_group
        self._ck(fn(self._ctx, group, ptr, n_seqs, aln_len, row_stride))

    def read_fasta(self, slot: int, path: str):
        """Native FASTA reader (reference semantics) into a pinned host buffer: (ptr, n_seqs, aln_len)."""
        ptr, n, length = C.c_void_p(), C.c_uint64(), C.c_uint64()
        self._ck(_abi.lib.snpg_read_fasta(self._ctx, slot, path.encode(), C.byref(ptr), C.byref(n), C.byref(length)))
        return ptr.value, n.value, length.value

    def fasta_wait(self, slot: int):
        """With OPT_INGEST_ASYNC: returns when the whole buffer of the slot is parsed (library calls wait by themselves)."""
        self._ck(_abi.lib.snpg_fasta_wait(self._ctx, slot))

    def fasta_view(self, ptr: int, n_seqs: int, aln_len: int) -> np.ndarray:
        """numpy view of a read_fasta buffer (valid until the slot is read into again)."""
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n_seqs, aln_len))

    def codon_strings(self, slot: int, product: int, codon: int, n_seqs: int) -> np.ndarray:
        out = np.zeros((n_seqs, 3), dtype=np.uint8)
        self._ck(_abi.lib.snpg_codon_strings(self._ctx, slot, product, codon, _ptr(out, C.c_uint8)))
        return out

    def group_size(self, group: int) -> int:
        n = C.c_uint64()
        self._ck(_abi.lib.snpg_group_size(self._ctx, group, C.byref(n)))
        return n.value

    # -- test hooks -----------------------------------------------------------
    def codon_indices(self, group: int, product: int) -> np.ndarray:
        out = np.zeros((self.group_size(group), self.local_codons(product)), dtype=np.uint8)
        self._ck(_abi.lib.snpg_get_codon_indices(self._ctx, group, product, _ptr(out, C.c_uint8)))
        return out

    def hist(self, group: int, product: int):
        nc = self.local_codons(product)
        h = np.zeros((nc, _abi.HIST_BINS), dtype=np.uint32)
        od = np.zeros(nc, dtype=np.uint8)
        self._ck(_abi.lib.snpg_get_hist(self._ctx, group, product, _ptr(h, C.c_uint32), _ptr(od, C.c_uint8)))
        return h, od

    def site_counts(self, group: int, aln_len: int) -> np.ndarray:
        """[aln_len, 4] counts of A, C, G, T per alignment column (needs OPT_SITE_COUNTS before set_products)."""
        out = np.zeros((aln_len, 4), dtype=np.uint32)
        self._ck(_abi.lib.snpg_get_site_counts(self._ctx, group, _ptr(out, C.c_uint32)))
        return out

    # -- per-codon engine -------------------------------------------------------
    def within(self, group: int, product: int) -> dict:
        nc = self.local_codons(product)
        r = dict(poly=np.zeros(nc, np.uint8), n_defined=np.zeros(nc, np.uint32), comparisons=np.zeros(nc, np.uint64),
                 stats4=np.zeros((nc, 4)), conserved=np.zeros(nc, np.uint8))
        self._ck(_abi.lib.snpg_within(self._ctx, group, product, _ptr(r["poly"], C.c_uint8), _ptr(r["n_defined"], C.c_uint32),
                                      _ptr(r["comparisons"], C.c_uint64), _ptr(r["stats4"], C.c_double), _ptr(r["conserved"], C.c_uint8)))
        for q, k in enumerate(STATS):
            r[k] = r["stats4"][:, q]
        return r

    def between(self, group_a: int, group_b: int, product: int) -> dict:
        nc = self.local_codons(product)
        r = dict(poly=np.zeros(nc, np.uint8), ndef_a=np.zeros(nc, np.uint32), ndef_b=np.zeros(nc, np.uint32),
                 comparisons=np.zeros(nc, np.uint64), stats4=np.zeros((nc, 4)), conserved=np.zeros(nc, np.uint8))
        self._ck(_abi.lib.snpg_between(self._ctx, group_a, group_b, product, _ptr(r["poly"], C.c_uint8), _ptr(r["ndef_a"], C.c_uint32),
                                       _ptr(r["ndef_b"], C.c_uint32), _ptr(r["comparisons"], C.c_uint64), _ptr(r["stats4"], C.c_double),
                                       _ptr(r["conserved"], C.c_uint8)))
        for q, k in enumerate(STATS):
            r[k] = r["stats4"][:, q]
        return r

    # -- resampling -------------------------------------------------------------
    def bootstrap_product(self, stats4: np.ndarray, B: int, keep: np.ndarray | None = None, want_sums: bool = False):
        stats4 = np.ascontiguousarray(stats4, dtype=np.float64)
        keep8 = None if keep is None else np.ascontiguousarray(keep, dtype=np.uint8)
        sums = np.zeros((B, 4)) if want_sums else None
        se = np.zeros(6)
        self._ck(_abi.lib.snpg_bootstrap_product(self._ctx, _ptr(stats4, C.c_double), _ptr(keep8, C.c_uint8), stats4.shape[0], B,
                                                 _ptr(sums, C.c_double), _ptr(se, C.c_double)))
        return sums, se

    def windows(self, stats4: np.ndarray, W: int, step: int, B: int = 0, keep: np.ndarray | None = None):
        stats4 = np.ascontiguousarray(stats4, dtype=np.float64)
        keep8 = None if keep is None else np.ascontiguousarray(keep, dtype=np.uint8)
        nwin = C.c_int64()
        self._ck(_abi.lib.snpg_windows(self._ctx, _ptr(stats4, C.c_double), _ptr(keep8, C.c_uint8), stats4.shape[0], W, step, 0,
                                       None, None, C.byref(nwin)))
        sums = np.zeros((nwin.value, 4))
        se = np.zeros((nwin.value, 3)) if B >= 2 else None
        self._ck(_abi.lib.snpg_windows(self._ctx, _ptr(stats4, C.c_double), _ptr(keep8, C.c_uint8), stats4.shape[0], W, step, B,
                                       _ptr(sums, C.c_double), _ptr(se, C.c_double), C.byref(nwin)))
        return sums, se

    def sliding_windows_r(self, stats4: np.ndarray, window: int, step: int, B: int = 1000, n_defined: np.ndarray | None = None,
                          min_defined: int = 6, codon_num: np.ndarray | None = None, numerator: str = "N", denominator: str = "S",
                          correction: str = "NONE") -> dict:
        """SNPGenie_sliding_windows.R on one product's per-codon table (snpg_sliding_windows_r): the script's arguments
        (2)-(8), its output columns as a dict of arrays, one entry per window."""
        stats4 = np.ascontiguousarray(stats4, dtype=np.float64)
        nd = None if n_defined is None else np.ascontiguousarray(n_defined, dtype=np.uint32)
        cn = None if codon_num is None else np.ascontiguousarray(codon_num, dtype=np.int64)
        sel = {"N": 0, "S": 1}
        opt = _abi.RWindows(C.sizeof(_abi.RWindows), sel[numerator], sel[denominator], 1 if correction == "JC" else 0, window, step, B, min_defined)
        n = C.c_int64()
        args = (self._ctx, _ptr(stats4, C.c_double), _ptr(nd, C.c_uint32), _ptr(cn, C.c_int64), stats4.shape[0], C.byref(opt), C.byref(n))
        self._ck(_abi.lib.snpg_sliding_windows_r(*args, None, 0))
        out = np.zeros((n.value, _abi.RWIN_COLS))
        if n.value:
            self._ck(_abi.lib.snpg_sliding_windows_r(*args, _ptr(out, C.c_double), n.value))
        return {k: out[:, i] for i, k in enumerate(_abi.RWIN_NAMES)}

    def pooled_diffs(self, freq: np.ndarray, coverage: np.ndarray):
        """snpgenie.pl's pooled-mode contraction (snpg_pooled_diffs): freq [C,64] codon frequencies, coverage [C] ->
        (N_diffs [C], S_diffs [C])."""
        freq = np.ascontiguousarray(freq, dtype=np.float64)
        cov = np.ascontiguousarray(coverage, dtype=np.float64)
        assert freq.ndim == 2 and freq.shape[1] == 64 and cov.shape == (freq.shape[0],)
        nd, sd = np.zeros(freq.shape[0]), np.zeros(freq.shape[0])
        self._ck(_abi.lib.snpg_pooled_diffs(self._ctx, _ptr(freq, C.c_double), _ptr(cov, C.c_double), freq.shape[0], _ptr(nd, C.c_double),
                                            _ptr(sd, C.c_double)))
        return nd, sd

    # -- whole job ----------------------------------------------------------------
    def within_all(self, group: int, B: int = 0):
        P = len(self.products)
        sums = np.zeros((P, 4))
        se = np.zeros((P, 6)) if B >= 2 else None
        self._ck(_abi.lib.snpg_within_all(self._ctx, group, B, _ptr(sums, C.c_double), _ptr(se, C.c_double)))
        return sums, se

    def between_all(self, groups: list[int], B: int = 0, min_taxa_total: int = 0, min_taxa_group: int = 0):
        """Every pair of `groups` x every product in one ABI call: filtered product sums [pairs, P, 4] and, for
        B >= 2, bootstrap SEs [pairs, P, 6].  Pairs in the order (g0,g1), (g0,g2), ..., (g[n-2],g[n-1])."""
        P, n = len(self.products), len(groups)
        n_pairs = n * (n - 1) // 2
        sums = np.zeros((n_pairs, P, 4))
        se = np.zeros((n_pairs, P, 6)) if B >= 2 else None
        g = (C.c_int * n)(*groups)
        self._ck(_abi.lib.snpg_between_all(self._ctx, g, n, min_taxa_total, min_taxa_group, B, _ptr(sums, C.c_double), _ptr(se, C.c_double)))
        return sums, se

    def between_job(self, ptrs: list[int], n_seqs: list[int], aln_len: int, row_strides: list[int], device: bool, B: int = 0,
                    min_taxa_total: int = 0, min_taxa_group: int = 0):
        """Loads the alignments at `ptrs` as groups 0..n-1 and analyses every pair x every product in ONE ABI call (one
        synchronisation): what `load_group_ptr` x n + `between_all` return, bit for bit.  The returned arrays are owned by the
        engine and overwritten by the next call with the same arguments."""
        key = ("between", tuple(ptrs), tuple(n_seqs), aln_len, tuple(row_strides), device, B, min_taxa_total, min_taxa_group)
        hit = self._jobs.get(key)
        if hit is None:
            P, n = len(self.products), len(ptrs)
            n_pairs = n * (n - 1) // 2
            sums = np.zeros((n_pairs, P, 4))
            se = np.zeros((n_pairs, P, 6)) if B >= 2 else None
            hit = (sums, se, (n, (C.c_void_p * n)(*ptrs), int(device), (C.c_uint64 * n)(*n_seqs), aln_len, (C.c_uint64 * n)(*row_strides),
                              min_taxa_total, min_taxa_group, B, _ptr(sums, C.c_double), _ptr(se, C.c_double)))
            if len(self._jobs) < 64:
                self._jobs[key] = hit           # the arrays of an argument tuple are reused by the next call with it
        self._ck(_abi.lib.snpg_between_job(self._ctx, *hit[2]))
        return hit[0], hit[1]

    def within_job(self, group: int, ptr: int, n_seqs: int, aln_len: int, row_stride: int, device: bool, B: int = 0):
        """Load + within_all in one ABI call (one synchronisation): per-product sums [P,4], SEs [P,6]."""
        P = len(self.products)
        sums = np.zeros((P, 4))
        se = np.zeros((P, 6)) if B >= 2 else None
        self._ck(_abi.lib.snpg_within_job(self._ctx, group, ptr, int(device), n_seqs, aln_len, row_stride, B,
                                          _ptr(sums, C.c_double), _ptr(se, C.c_double)))
        return sums, se

    def staged_bytes_per_row(self) -> int:
        first, n = C.c_int64(), C.c_int64()
        self._ck(_abi.lib.snpg_staged_columns(self._ctx, C.byref(first), C.byref(n)))
        return n.value

    def probe_peaks(self) -> dict:
        out = np.zeros(4)
        self._ck(_abi.lib.snpg_probe_peaks(self._ctx, _ptr(out, C.c_double)))
        return {"philox_gdraws_per_s": float(out[0]) / 1e9, "dmma_tflops": float(out[1])}

    def total_codons(self) -> int:
        return self.shard_range()[2]

    def window_plan(self, W: int, step: int) -> np.ndarray:
        """win_ofs [P+1]: product p owns the windows [win_ofs[p], win_ofs[p+1]) of a job's window outputs."""
        ofs = np.zeros(len(self.products) + 1, dtype=np.int64)
        self._ck(_abi.lib.snpg_window_plan(self._ctx, W, step, _ptr(ofs, C.c_int64)))
        return ofs

    def _host_array(self, name, shape, dtype, pinned):
        """Output arrays of job(): reused between calls (a replayed CUDA graph writes to fixed addresses)."""
        key = (name, tuple(shape), np.dtype(dtype).str, pinned)
        a = self._pinned.get(key)
        if a is None:
            if pinned:
                import torch
                t = torch.empty(tuple(shape), dtype=getattr(torch, np.dtype(dtype).name)).pin_memory()
                a = (t, t.numpy())
            else:
                a = (None, np.zeros(shape, dtype=dtype))
            self._pinned[key] = a
        return a[1]

    def job(self, group: int, ptr: int | None, n_seqs: int = 0, aln_len: int = 0, row_stride: int = 0, device: bool = False, B: int = 0,
            window: int = 0, step: int = 1, B_win: int = 0, per_codon: bool = True, pinned: bool = True) -> dict:
        """The whole within-group job in ONE ABI call (snpg_within_job_ex): product sums and SEs, the per-codon
        tables over the global codon axis and, with window > 0, the sliding windows with their bootstrap SEs.
        ptr=None: the group is already loaded.  The returned arrays are owned by the engine and overwritten by the next
        job of the same shape.  (The struct and the arrays of an argument tuple are built once: a step of the resident
        benchmark is 0.25 ms, and assembling them anew costs 28 us of interpreter time between two GPU jobs.)"""
        key = (group, ptr, n_seqs, aln_len, row_stride, device, B, window, step, B_win, per_codon, pinned)
        hit = self._jobs.get(key)
        if hit is not None:
            self._ck(_abi.lib.snpg_within_job_ex(self._ctx, group, ptr, hit[2], n_seqs, aln_len, row_stride, hit[1]))
            return hit[0]
        P = len(self.products)
        total = self.total_codons() if (self.world == 1 or self.peer_connected or self.team_size > 1) else self.shard_range()[1]
        out = {"prod_sums": self._host_array("prod_sums", (P, 4), np.float64, False)}
        spec = _abi.Job()
        spec.struct_size = C.sizeof(_abi.Job)
        spec.n_bootstraps = B
        spec.prod_sums = out["prod_sums"].ctypes.data
        if B >= 2:
            out["prod_se"] = self._host_array("prod_se", (P, 6), np.float64, False)
            out["prod_undefined"] = self._host_array("prod_undefined", (P, 6), np.uint32, False)
            spec.prod_se, spec.prod_undefined = out["prod_se"].ctypes.data, out["prod_undefined"].ctypes.data
        if per_codon:
            for name, shape, dt in (("poly", (total,), np.uint8), ("n_defined", (total,), np.uint32), ("comparisons", (total,), np.uint64),
                                    ("stats4", (total, 4), np.float64), ("conserved", (total,), np.uint8)):
                out[name] = self._host_array(name, shape, dt, pinned)
                setattr(spec, name, out[name].ctypes.data)
        if window > 0:
            ofs = self.window_plan(window, step)
            out["win_ofs"] = ofs
            nw = int(ofs[-1])
            spec.window, spec.step, spec.n_window_bootstraps = window, step, B_win
            out["win_sums"] = self._host_array("win_sums", (nw, 4), np.float64, False)
            spec.win_sums = out["win_sums"].ctypes.data
            if B_win >= 2:
                out["win_se"] = self._host_array("win_se", (nw, 3), np.float64, False)
                spec.win_se = out["win_se"].ctypes.data
        self._ck(_abi.lib.snpg_within_job_ex(self._ctx, group, ptr, int(device), n_seqs, aln_len, row_stride, C.byref(spec)))
        if len(self._jobs) < 64:
            self._jobs[key] = (out, C.byref(spec), int(device), spec)       # (spec is kept alive for its byref)
        return out

    # -- device-pointer plumbing for multi-GPU ------------------------------------
    def within_stats_device(self, group: int, d_stats4_ptr: int):
        self._ck(_abi.lib.snpg_within_stats_device(self._ctx, group, d_stats4_ptr))

    def bootstrap_all_device(self, d_stats4_full_ptr: int, b_first: int, b_count: int, d_rep_ptr: int):
        self._ck(_abi.lib.snpg_bootstrap_all_device(self._ctx, d_stats4_full_ptr, b_first, b_count, d_rep_ptr))

    def product_sums_device(self, d_stats4_full_ptr: int) -> np.ndarray:
        sums = np.zeros((len(self.products), 4))
        self._ck(_abi.lib.snpg_product_sums_device(self._ctx, d_stats4_full_ptr, _ptr(sums, C.c_double)))
        return sums

    def rep_moments_device(self, d_rep_ptr: int, n_items: int, B: int) -> np.ndarray:
        se = np.zeros((n_items, 6))
        self._ck(_abi.lib.snpg_rep_moments_device(self._ctx, d_rep_ptr, n_items, B, _ptr(se, C.c_double)))
        return se

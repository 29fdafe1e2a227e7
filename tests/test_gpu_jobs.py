"""GPU parity of the whole-job entry (snpg_within_job_ex), its CUDA-graph replay and the multi-GPU exchange.

The job must return, in one call, exactly what the call-by-call API returns (which test_gpu_parity.py holds against
the oracle and the reference goldens): per-codon tables over the global codon axis, product sums, sliding-window sums
bitwise, SEs statistically (the Philox streams of different call kinds differ by construction).  The multi-GPU path must
return the single-GPU results bit for bit -- ranks on one device exercise the same protocol as ranks on eight.
"""
import multiprocessing as mp
import os

import numpy as np
import pytest

from oracle import oracle as orc
from snpgenie_b200 import synth

pytestmark = pytest.mark.gpu

STATS = ("N_sites", "S_sites", "N_diffs", "S_diffs")


def _hiv(n=400, seed=11):
    products = synth.hiv_products()
    aln = synth.make_alignment(n, 9719, products, seed, gap_frac=0.025, n_frac=0.02, iupac_frac=0.005)
    return products, aln


def _snapshot(out):
    return {k: np.array(v, copy=True) for k, v in out.items()}


def _call_by_call(eng, products):
    rows = [eng.within(0, i) for i in range(len(products))]
    return {k: np.concatenate([r[k] for r in rows]) for k in ("poly", "n_defined", "comparisons", "stats4", "conserved")}


@pytest.mark.parametrize("pinned", [True, False])
def test_job_returns_what_the_calls_return(pinned):
    from snpgenie_b200.engine import Engine
    products, aln = _hiv()
    W, step = 40, 7
    with Engine(device=0, seed=5) as eng:
        eng.set_products(products, aln.shape[1])
        out = _snapshot(eng.job(0, aln.ctypes.data, aln.shape[0], aln.shape[1], aln.strides[0], device=False, B=1500,
                                window=W, step=step, B_win=800, pinned=pinned))
        want = _call_by_call(eng, products)
        for k, v in want.items():
            assert np.array_equal(out[k], v), k                       # same kernels' arithmetic: bit-identical
        sums, se = eng.within_all(0, B=1500)
        assert np.array_equal(out["prod_sums"], sums)
        assert np.all(np.abs(out["prod_se"][:, :3] / se[:, :3] - 1) < 0.2)
        assert (out["prod_undefined"] == 0).all()
        # windows: product p owns win_ofs[p]..win_ofs[p+1]; sums bitwise the oracle's ascending additions
        ofs = np.cumsum([0] + [p.n_codons for p in products])
        for i, p in enumerate(products):
            st = want["stats4"][ofs[i]:ofs[i + 1]]
            w0, w1 = out["win_ofs"][i], out["win_ofs"][i + 1]
            if p.n_codons < W:
                assert w0 == w1
                continue
            osums, _ = orc.windows(st, W, step)
            assert np.array_equal(out["win_sums"][w0:w1], osums), p.name
            _, ose = orc.windows(st, W, step, B=800, seed=3)
            got = out["win_se"][w0:w1]
            big = ose[:, 2] > 1e-3
            if big.any():
                assert np.median(np.abs(got[big, 2] / ose[big, 2] - 1)) < 0.1, p.name
            assert ((got == 0) == (ose == 0)).all()                    # an all-conserved window has SE exactly 0


def test_job_graph_replay_is_identical():
    """Device-resident alignment: the third identical call on replays a CUDA graph; everything but the SEs (a fresh
    Philox stream per call) is bit-identical, and the SEs agree statistically."""
    import torch
    from snpgenie_b200.engine import Engine
    products, aln = _hiv(n=700)
    pitch = (aln.shape[1] + 127) // 128 * 128
    d = torch.zeros((aln.shape[0], pitch), dtype=torch.uint8, device="cuda:0")
    d[:, :aln.shape[1]] = torch.from_numpy(aln).to("cuda:0")
    torch.cuda.synchronize()
    with Engine(device=0, seed=5) as eng:
        eng.set_products(products, aln.shape[1])
        outs = []
        for _ in range(5):
            outs.append(_snapshot(eng.job(0, d.data_ptr(), aln.shape[0], aln.shape[1], pitch, device=True, B=1000, window=50, step=10, B_win=300)))
        assert eng.graph_replays >= 3
        for o in outs[1:]:
            for k in ("poly", "n_defined", "comparisons", "stats4", "conserved", "prod_sums", "win_sums"):
                assert np.array_equal(o[k], outs[0][k]), k
            assert not np.array_equal(o["prod_se"], outs[0]["prod_se"])          # new replicates every call
            assert np.all(np.abs(o["prod_se"][:, :3] / outs[0]["prod_se"][:, :3] - 1) < 0.25)
        # a reload of the same group with another alignment between replays must not be served by the old graph
        aln2 = synth.make_alignment(500, aln.shape[1], products, 99)
        eng.load_group(0, aln2)
        assert eng.group_size(0) == 500
        o = _snapshot(eng.job(0, d.data_ptr(), aln.shape[0], aln.shape[1], pitch, device=True, B=1000, window=50, step=10, B_win=300))
        assert eng.group_size(0) == aln.shape[0]
        assert np.array_equal(o["stats4"], outs[0]["stats4"])


def test_config2_full_size_job_every_codon():
    """BASELINE config 2 at full size through the benched path -- snpg_within_job on a device pointer, graph replays
    included -- with EVERY one of the 9,677 histograms and per-codon rows checked against the oracle."""
    import torch
    from snpgenie_b200.engine import Engine
    products = synth.sars2_products()
    n, L = 10000, 29903
    aln = synth.make_alignment(n, L, products, 20261021)
    pitch = (L + 127) // 128 * 128
    d = torch.zeros((n, pitch), dtype=torch.uint8, device="cuda:0")
    d[:, :L] = torch.from_numpy(aln).to("cuda:0")
    torch.cuda.synchronize()
    sites, diffs = orc.tables()
    with Engine(device=0, seed=1) as eng:
        eng.set_products(products, L)
        for _ in range(4):
            out = eng.job(0, d.data_ptr(), n, L, pitch, device=True, B=2000)
        assert eng.graph_replays >= 2
        out = _snapshot(out)
        ofs = np.cumsum([0] + [p.n_codons for p in products])
        assert ofs[-1] == 9677
        for i, p in enumerate(products):
            pos = orc.product_positions(p.starts, p.stops, p.strand, L)
            raw, codes = orc.extract(aln, pos, p.strand)
            want_h = orc.hist(codes)
            h, od = eng.hist(0, i)
            assert np.array_equal(h, want_h), p.name                    # all histograms, bit-exact
            want = orc.within_hist(want_h, orc.other_distinct(raw, codes), sites, diffs)
            sl = slice(ofs[i], ofs[i + 1])
            assert np.array_equal(out["poly"][sl], want["poly"])
            assert np.array_equal(out["n_defined"][sl], want["n_defined"])
            assert np.array_equal(out["comparisons"][sl].astype(np.float64), want["comparisons"])
            for q, k in enumerate(STATS):
                assert np.allclose(out["stats4"][sl, q], want[k], rtol=1e-12, atol=1e-15), (p.name, k)
                assert np.allclose(out["prod_sums"][i, q], want[k].sum(), rtol=1e-12), (p.name, k)
        assert np.all(out["prod_se"][:, :3] > 0)


def test_boot_slots_option():
    """SNPG_OPT_BOOT_SLOTS: FAITHFUL draws from C+1 slots (a replicate is on average C/(C+1) of the product, wg:885),
    CORRECTED from the C codons."""
    from snpgenie_b200 import _abi
    from snpgenie_b200.engine import Engine
    rng = np.random.default_rng(3)
    C, B = 9, 40000
    stats4 = rng.random((C, 4)) + 0.5
    with Engine(device=0, seed=8) as eng:
        faithful, _ = eng.bootstrap_product(stats4, B, want_sums=True)
        eng.set_option(_abi.OPT_BOOT_SLOTS, _abi.SLOTS_CORRECTED)
        corrected, _ = eng.bootstrap_product(stats4, B, want_sums=True)
    tot = stats4.sum(axis=0)
    assert np.allclose(faithful.mean(axis=0), tot * C / (C + 1), rtol=0.01)
    assert np.allclose(corrected.mean(axis=0), tot, rtol=0.01)


def test_jukes_cantor_undefined_is_reported():
    """A replicate with dN or dS >= 3/4 has no Jukes-Cantor distance (the reference dies: bgp:1047); the job counts them."""
    from snpgenie_b200.engine import Engine
    from snpgenie_b200.hostio import Product
    # four equally frequent alleles that differ at all three positions: 6 of 7 sequence pairs differ everywhere
    n, L = 8, 90
    aln = np.repeat(np.frombuffer(b"ACGT", dtype=np.uint8), 2)[:, None].repeat(L, axis=1).copy()
    with Engine(device=0, seed=2) as eng:
        eng.set_products([Product("p", 1, [1], [90])], L)
        out = eng.job(0, aln.ctypes.data, n, L, aln.strides[0], device=False, B=500)
        s = out["prod_sums"][0]
        dn, ds = s[2] / s[0], s[3] / s[1]
        assert dn >= 0.75, (dn, ds)                                 # the point of the construction
        assert out["prod_undefined"][0, :3].sum() == 0
        assert out["prod_undefined"][0, 3] == 500                   # every replicate has the same, undefined, JC dN
        assert (out["prod_undefined"][0, 4] == 500) == (ds >= 0.75)
        assert not np.isfinite(out["prod_se"][0, 3])


# ----------------------------------------------------------------- several GPUs (ranks)

def _reference_run(products, aln, B, window=0, step=1, B_win=0, seed=5):
    from snpgenie_b200.engine import Engine
    with Engine(device=0, seed=seed) as eng:
        eng.set_products(products, aln.shape[1])
        out = _snapshot(eng.job(0, aln.ctypes.data, aln.shape[0], aln.shape[1], aln.strides[0], device=False, B=B, window=window,
                                step=step, B_win=B_win))
        hists = [eng.hist(0, i) for i in range(len(products))]
    return out, hists


@pytest.mark.parametrize("ranks", [2, 3])
def test_team_of_ranks_matches_one_gpu_bit_for_bit(ranks):
    """snpg_create_multi: codon ranges and replicate ranges split over the ranks, exchanged through peer memory.  The
    ranks share GPU 0 when the box has fewer GPUs (same protocol, same kernels)."""
    import torch
    from snpgenie_b200.engine import Engine
    n_dev = torch.cuda.device_count()
    if ranks > 2 and n_dev < ranks:
        # kernels of different ranks wait for each other: on ONE device that needs every rank's streams to sit in their
        # own hardware queue (CUDA_DEVICE_MAX_CONNECTIONS, 8 by default) -- two ranks fit, three do not
        pytest.skip("needs one GPU per rank")
    products, aln = _hiv(n=300, seed=21)
    B = 999
    want, want_h = _reference_run(products, aln, B, window=30, step=5, B_win=200)
    devs = [r % n_dev for r in range(ranks)]
    with Engine(seed=5, gpus=devs) as eng:
        assert eng.team_size == ranks
        eng.set_products(products, aln.shape[1])
        for rep in range(3):
            out = _snapshot(eng.job(0, aln.ctypes.data, aln.shape[0], aln.shape[1], aln.strides[0], device=False, B=B, window=30, step=5, B_win=200))
            for k in ("poly", "n_defined", "comparisons", "stats4", "conserved", "prod_sums", "win_sums"):
                assert np.array_equal(out[k], want[k]), (rep, k)
            if rep == 0:
                # same seed, same job number: the same replicates, drawn by different ranks -> the same SEs, bit for bit
                assert np.array_equal(out["prod_se"], want["prod_se"])
        for i in range(len(products)):
            h, od = eng.hist(0, i)
            assert np.array_equal(h, want_h[i][0]) and np.array_equal(od, want_h[i][1])
            r = eng.within(0, i)
            sl = slice(sum(p.n_codons for p in products[:i]), sum(p.n_codons for p in products[:i + 1]))
            assert np.array_equal(r["stats4"], want["stats4"][sl])
        sums, se = eng.within_all(0, B=300)
        assert np.array_equal(sums, want["prod_sums"])


def _ipc_rank(rank, world, q_in, q_out, seed):
    """One process per rank (all on GPU 0 when the box has one GPU): handles travel through multiprocessing queues."""
    import numpy as np
    import torch
    from snpgenie_b200 import synth
    from snpgenie_b200.engine import Engine
    products = synth.hiv_products()
    aln = synth.make_alignment(300, 9719, products, 21, gap_frac=0.025, n_frac=0.02, iupac_frac=0.005)
    dev = rank % torch.cuda.device_count()
    with Engine(device=dev, seed=seed, rank=rank, world=world) as eng:
        eng.set_products(products, aln.shape[1])
        q_out.put((rank, eng.peer_export(2000)))
        handles = q_in.get(timeout=120)
        eng.peer_connect(handles)
        res = []
        for _ in range(3):
            out = eng.job(0, aln.ctypes.data, aln.shape[0], aln.shape[1], aln.strides[0], device=False, B=999, pinned=False)
            res.append({k: np.array(v, copy=True) for k, v in out.items()})
        q_out.put((rank, res))
        q_in.get(timeout=120)            # keep the exchange region mapped until every rank is done


def test_ranks_in_separate_processes_over_cuda_ipc():
    world = 2
    products, aln = _hiv(n=300, seed=21)
    want, _ = _reference_run(products, aln, 999)
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    q_in = [ctx.Queue() for _ in range(world)]
    procs = [ctx.Process(target=_ipc_rank, args=(r, world, q_in[r], q_out, 5)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        handles = dict(q_out.get(timeout=180) for _ in range(world))
        for q in q_in:
            q.put([handles[r] for r in range(world)])
        results = dict(q_out.get(timeout=180) for _ in range(world))
    finally:
        for q in q_in:
            q.put(None)
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.kill()
    for p in procs:
        assert p.exitcode == 0
    for rank in range(world):
        for rep, out in enumerate(results[rank]):
            for k in ("poly", "n_defined", "comparisons", "stats4", "conserved", "prod_sums"):
                assert np.array_equal(out[k], want[k]), (rank, rep, k)
            if rep == 0:
                assert np.array_equal(out["prod_se"], want["prod_se"]), rank


def test_unconnected_ranks_fail_loudly():
    from snpgenie_b200 import _abi
    from snpgenie_b200.engine import Engine
    products, aln = _hiv(n=64)
    with Engine(device=0, seed=1, rank=0, world=2) as eng:
        eng.set_products(products, aln.shape[1])
        with pytest.raises(_abi.SnpgError) as e:
            eng.peer_connect([b"\0" * 128, b"\0" * 128])
        assert e.value.code == _abi.E_STATE
        h = eng.peer_export(100)
        with pytest.raises(_abi.SnpgError) as e:
            eng.peer_connect([h, b"\0" * 128])
        assert e.value.code == _abi.E_ARG
        with pytest.raises(_abi.SnpgError) as e:          # windows need the whole axis
            eng.job(0, aln.ctypes.data, aln.shape[0], aln.shape[1], aln.strides[0], window=10, step=1)
        assert e.value.code == _abi.E_STATE


@pytest.mark.parametrize("device", [True, False])
def test_between_job_equals_loads_plus_between_all_bit_for_bit(device):
    """snpg_between_job (the loads queued behind each other, the group pairs on parallel streams, one synchronisation)
    returns what snpg_load_group x n + snpg_between_all return -- also when those run with per-kernel timing on, which
    keeps the pairs on one stream, one after the other."""
    import torch
    from snpgenie_b200 import _abi
    from snpgenie_b200.engine import Engine
    products = synth.hiv_products()
    L = 9719
    pitch = (L + 127) // 128 * 128
    alns = [synth.make_alignment(n, L, products, 900 + g, gap_frac=0.03, n_frac=0.02, iupac_frac=0.004) for g, n in enumerate((70, 45, 52, 64, 38))]     # 10 pairs: more than the six parallel streams
    bufs = []
    for a in alns:
        t = torch.zeros((a.shape[0], pitch), dtype=torch.uint8, device="cuda:0" if device else "cpu")
        t[:, :L] = torch.from_numpy(a)
        bufs.append(t)
    ns = [a.shape[0] for a in alns]
    want = {}
    for profile in (0, 1):
        with Engine(device=0, seed=5) as e:
            e.set_option(_abi.OPT_PROFILE, 0xFFFF if profile else 0)
            e.set_products(products, L)
            for g, t in enumerate(bufs):
                e.load_group_ptr(g, t.data_ptr(), ns[g], L, pitch, device=device)
            want[profile] = e.between_all([0, 1, 2, 3, 4], B=500, min_taxa_total=90, min_taxa_group=30)
    assert np.array_equal(want[0][0], want[1][0]) and np.array_equal(want[0][1], want[1][1], equal_nan=True)   # (Jukes-Cantor SEs of distant groups are NaN)
    with Engine(device=0, seed=5) as e:
        e.set_products(products, L)
        sums, se = (x.copy() for x in e.between_job([t.data_ptr() for t in bufs], ns, L, [pitch] * 5, device=device, B=500, min_taxa_total=90, min_taxa_group=30))
        assert np.array_equal(sums, want[0][0]) and np.array_equal(se, want[0][1], equal_nan=True)
        assert np.isfinite(sums).all() and (se[:, :, :3] >= 0).all() and sums[:, :, 0].max() > 0
        # a second job on the same ctx: same sums, another Philox stream
        sums2, se2 = e.between_job([t.data_ptr() for t in bufs], ns, L, [pitch] * 5, device=device, B=500, min_taxa_total=90, min_taxa_group=30)
        assert np.array_equal(sums2, sums) and not np.array_equal(se2[:, :, :3], se[:, :, :3])
        with pytest.raises(Exception):
            e.between_job([bufs[0].data_ptr()], ns[:1], L, [pitch], device=device, B=0)

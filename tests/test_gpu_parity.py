"""GPU parity: libsnpgenie_cuda.so (through the C ABI) against the oracle and the reference goldens.

Bars (BASELINE.json north_star): codon indices and histograms bit-exact; per-codon
and per-product N/S sites and differences within 1e-12 relative; window sums
bitwise (same additions in the same order); bootstrap SE statistically.
"""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from snpgenie_b200 import hostio, synth
from tests import refio
from tests.refio import STATS

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-12


def close(a, b, rtol=RTOL, atol=1e-15):
    return np.allclose(a, b, rtol=rtol, atol=atol)


@pytest.fixture(scope="module")
def eng():
    from snpgenie_b200.engine import Engine
    e = Engine(device=0, seed=1234, keep_codes=True)
    yield e
    e.close()


@pytest.fixture(scope="module", params=["classes", "weights"])
def eng_boot(request):
    """The statistical bootstrap tests run on the class-aggregated kernel (the default: class counts by exact binomials, the
    other draws one by one) and on the weight-matrix kernel (every draw one by one)."""
    from snpgenie_b200 import _abi
    from snpgenie_b200.engine import Engine
    e = Engine(device=0, seed=1234)
    e.set_option(_abi.OPT_BOOTSTRAP_KERNEL, _abi.BOOT_CLASSES if request.param == "classes" else _abi.BOOT_WEIGHTS)
    yield e
    e.close()


@pytest.fixture(scope="module", params=["codes", "fused", "tiles", "spans"])
def eng2(request):
    """All histogram routes: K1 pack + K2 hist over the kept code matrix, the fused vector-load span kernel (the default),
    the fused TMA-tiled kernel and the TMA-fed span kernel."""
    from snpgenie_b200 import _abi
    from snpgenie_b200.engine import Engine
    e = Engine(device=0, seed=1234)
    e.has_codes = request.param == "codes"
    e.route = request.param
    e.set_option(_abi.OPT_KEEP_CODES, 1 if e.has_codes else 0)
    e.set_option(_abi.OPT_FUSED_KERNEL, {"tiles": _abi.FUSED_TMA_TILES, "spans": _abi.FUSED_TMA_SPANS}.get(request.param, _abi.FUSED_LDG_SPANS))
    yield e
    e.close()


def _oracle_product(aln, p):
    pos = orc.product_positions(p.starts, p.stops, p.strand, aln.shape[1])
    raw, codes = orc.extract(aln, pos, p.strand)
    return raw, codes


# ----------------------------------------------------------------- K1 / K2 bit-exact

@pytest.mark.parametrize("case,minus", [("w_cfg1", False), ("w_gaps", False), ("w_heavy", False), ("w_minus", True)])
def test_codon_indices_and_histograms_bit_exact(eng2, case, minus):
    eng = eng2
    d, aln, _, _ = refio.within_case(case)
    products = hostio.read_gtf_products(os.path.join(d, "cds.gtf"), include_minus=minus)
    eng.set_products(products, aln.shape[1])
    before, before_s = eng.tile_launches, eng.span_tma_launches
    eng.load_group(0, aln)
    # the "tiles" and "spans" routes must really be the TMA kernels, the other two must not touch them
    assert (eng.tile_launches > before) == (eng.route == "tiles")
    assert (eng.span_tma_launches > before_s) == (eng.route == "spans")
    for i, p in enumerate(products):
        raw, codes = _oracle_product(aln, p)
        if eng.has_codes:
            assert (eng.codon_indices(0, i) == codes).all(), (case, p.name)
        h, od = eng.hist(0, i)
        assert (h == orc.hist(codes)).all(), (case, p.name)
        assert (od == orc.other_distinct(raw, codes)).all(), (case, p.name)


@pytest.mark.parametrize("n", [1, 2, 31, 127, 128, 129, 1000, 8191, 8192, 8193, 20000])
def test_histogram_ragged_row_counts(eng2, n):
    """Row counts around every padding / chunking boundary, uniform random codons (worst case for K2)."""
    eng = eng2
    rng = np.random.default_rng(n)
    L = 3 * 70 + 5
    aln = np.frombuffer(b"ACGTNacgtu-RY", dtype=np.uint8)[rng.integers(0, 13, size=(n, L))]
    aln[:, :60] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(n, 60))]
    products = [hostio.Product("p", 1, [3], [212])]
    eng.set_products(products, L)
    eng.load_group(0, aln)
    raw, codes = _oracle_product(aln, products[0])
    if eng.has_codes:
        assert (eng.codon_indices(0, 0) == codes).all()
    h, od = eng.hist(0, 0)
    assert (h == orc.hist(codes)).all()
    assert (h.sum(axis=1) == n).all()
    assert (od == orc.other_distinct(raw, codes)).all()


def test_device_resident_and_strided_input_match_host_input(eng2):
    import torch
    eng = eng2
    products = synth.hiv_products()
    aln = synth.make_alignment(700, 9719, products, 77, gap_frac=0.025, n_frac=0.02, iupac_frac=0.005)
    eng.set_products(products, 9719)
    eng.load_group(0, aln)
    ref = [eng.hist(0, i)[0] for i in range(len(products))]
    # device buffer with a padded row pitch
    pitch = 9728
    d = torch.zeros((700, pitch), dtype=torch.uint8, device="cuda:0")
    d[:, :9719] = torch.from_numpy(aln).cuda()
    eng.load_group_ptr(1, d.data_ptr(), 700, 9719, pitch, device=True)
    for i in range(len(products)):
        assert (eng.hist(1, i)[0] == ref[i]).all()
        assert (ref[i] == orc.hist(_oracle_product(aln, products[i])[1])).all()
        if eng.has_codes:
            assert (eng.codon_indices(1, i) == eng.codon_indices(0, i)).all()
    # a device buffer the fused kernel cannot take (odd pitch): same result through the fallback
    d2 = torch.zeros((700, 9719 + 5), dtype=torch.uint8, device="cuda:0")
    d2[:, :9719] = torch.from_numpy(aln).cuda()
    eng.load_group_ptr(3, d2.data_ptr(), 700, 9719, 9719 + 5, device=True)
    for i in range(len(products)):
        assert (eng.hist(3, i)[0] == ref[i]).all()
    # host buffer with a row stride (a view into a wider matrix)
    wide = np.zeros((700, 10000), dtype=np.uint8)
    wide[:, :9719] = aln
    eng.load_group(2, wide[:, :9719])
    for i in range(len(products)):
        assert (eng.hist(2, i)[0] == ref[i]).all()


# ----------------------------------------------------------------- K3 vs reference goldens

def _check_within_case(eng, case, products, aln, codon_rows, prod_rows):
    eng.set_products(products, aln.shape[1])
    eng.load_group(0, aln)
    sums, _ = eng.within_all(0)
    for i, p in enumerate(products):
        gold = refio.codon_arrays(codon_rows, p.name, p.n_codons)
        r = eng.within(0, i)
        assert (r["poly"] == gold["poly"]).all(), (case, p.name)
        for k in STATS:
            assert close(r[k], gold[k]), (case, p.name, k)
        raw, codes = _oracle_product(aln, p)
        res = orc.within_pairloop(raw)
        assert (r["n_defined"] == res["n_defined"]).all()
        nd = res["n_defined"].astype(np.uint64)
        assert (r["comparisons"] == np.where(gold["poly"] == 1, nd * (nd - 1) // 2, 0)).all()
        # conserved codon column (wg:726): the codon string, empty when nothing is defined
        for c in np.flatnonzero(gold["poly"] == 0):
            code = int(r["conserved"][c])
            want = gold["comparisons_str"][c]
            if code < 64:
                assert "ACGT"[code >> 4] + "ACGT"[(code >> 2) & 3] + "ACGT"[code & 3] == want
            elif code == 64:
                assert want == ""
            else:
                assert want != "" and any(ch not in "ACGT" for ch in want)
        pr = prod_rows[p.name]
        for q, k in enumerate(STATS):
            assert close(sums[i, q], float(pr[k])), (case, p.name, k)
        assert close(sums[i, 2] / sums[i, 0], float(pr["dN"])) and close(sums[i, 3] / sums[i, 1], float(pr["dS"]))


@pytest.mark.parametrize("case", ["w_cfg1", "w_gaps", "w_heavy"])
def test_within_matches_reference(eng2, case):
    eng = eng2
    d, aln, codon_rows, prod_rows = refio.within_case(case)
    products = hostio.read_gtf_products(os.path.join(d, "cds.gtf"))
    _check_within_case(eng, case, products, aln, codon_rows, prod_rows)


def test_minus_strand_matches_revcom_route(eng2):
    eng = eng2
    d, aln, codon_rows, prod_rows = refio.within_case("w_minus")
    products = [p for p in hostio.read_gtf_products(os.path.join(d, "cds.gtf"), include_minus=True) if p.strand < 0]
    _check_within_case(eng, "w_minus", products, aln, codon_rows, prod_rows)


def test_between_matches_reference(eng):
    d = os.path.join(refio.GOLD, "b_3groups")
    names = ("grpA.fasta", "grpB.fasta", "grpC.fasta")
    groups = {g: hostio.read_fasta(os.path.join(d, g))[1] for g in names}
    products = hostio.read_gtf_products(os.path.join(d, "cds.gtf"))
    _, rows = refio.read_tsv(os.path.join(d, "between_group_codon_results.txt"))
    maps = refio.between_label_maps(os.path.join(d, "stdout.log"))
    eng.set_products(products, groups[names[0]].shape[1])
    for gi, g in enumerate(names):
        eng.load_group(gi, groups[g])
    checked = 0
    for i, p in enumerate(products):
        prow = [r for r in rows if r["product"] == p.name]
        for lab1, lab2 in sorted({(r["group_1"], r["group_2"]) for r in prow}):
            g1, g2 = maps[p.name][lab1], maps[p.name][lab2]
            r = eng.between(names.index(g1), names.index(g2), i)
            ra, rb = _oracle_product(groups[g1], p), _oracle_product(groups[g2], p)
            res = orc.between_pairloop(ra[0], rb[0])     # sequence order = file order (deterministic tie-break)
            for row in (x for x in prow if (x["group_1"], x["group_2"]) == (lab1, lab2)):
                c = int(row["codon"].replace("codon_", "")) - 1
                assert int(row["num_defined_codons_g1"] or 0) == r["ndef_a"][c]
                assert int(row["num_defined_codons_g2"] or 0) == r["ndef_b"][c]
                assert (row["variability"] == "polymorphic") == bool(r["poly"][c])
                for k in STATS:
                    # always equal to the oracle run in file order ...
                    assert close(r[k][c], res[k][c]), (p.name, g1, g2, c, k)
                    # ... and to the reference unless its hash-ordered "first defined codon" is ambiguous
                    if close(res[k][c], float(row[k])):
                        assert close(r[k][c], float(row[k]))
                checked += 1
    assert checked == 3 * sum(p.n_codons for p in products)


# ----------------------------------------------------------------- mid-size vs oracle closed form

def test_config3_shape_against_oracle(eng2):
    """HIV-shaped: overlapping frames, 2-segment product, '-' strand CDS, 5 % gaps/ambiguous."""
    eng = eng2
    products = synth.hiv_products()
    aln = synth.make_alignment(1500, 9719, products, 20261022, gap_frac=0.025, n_frac=0.02, iupac_frac=0.005,
                               lower_frac=0.01, u_frac=0.01)
    eng.set_products(products, 9719)
    eng.load_group(0, aln)
    sums, _ = eng.within_all(0)
    for i, p in enumerate(products):
        raw, codes = _oracle_product(aln, p)
        if eng.has_codes:
            assert (eng.codon_indices(0, i) == codes).all()
        h, od = eng.hist(0, i)
        assert (h == orc.hist(codes)).all()
        assert (od == orc.other_distinct(raw, codes)).all()
        want = orc.within_hist(h, orc.other_distinct(raw, codes))
        got = eng.within(0, i)
        assert (got["poly"] == want["poly"]).all()
        for k in STATS:
            assert close(got[k], want[k]), (p.name, k)
            assert close(sums[i, STATS.index(k)], want[k].sum())
    # literal O(n^2) pair loop on one product, as the reference would do it
    raw, _ = _oracle_product(aln[:400], products[2])
    eng.load_group(1, aln[:400])
    lit = orc.within_pairloop(raw)
    got = eng.within(1, 2)
    assert (got["poly"] == lit["poly"]).all()
    for k in STATS:
        assert close(got[k], lit[k]), k


@pytest.mark.parametrize("route", ["codes", "first_defined"])
def test_between_mid_size_against_oracle(route):
    """route "codes": the kept code matrix serves the first-defined-codon rule; "first_defined": the fast histogram route
    with SNPG_OPT_FIRST_DEFINED (what the between-group command line uses) must give the same answers."""
    from snpgenie_b200 import _abi
    from snpgenie_b200.engine import Engine
    with Engine(device=0, seed=1234, keep_codes=(route == "codes")) as eng:
        if route == "first_defined":
            eng.set_option(_abi.OPT_FIRST_DEFINED, 1)
        _between_mid_size(eng)


def _between_mid_size(eng):
    products = synth.sars2_products()[2:5]
    a = synth.make_alignment(900, 29903, products, 5, gap_frac=0.01, n_frac=0.01)
    b = synth.make_alignment(700, 29903, products, 5, gap_frac=0.01, n_frac=0.01, p_poly=0.5)
    b[:, 21600:21700] = a[0, 21600:21700]
    b[:, 21563 + 3 * 50:21563 + 3 * 53] = ord("-")     # group b undefined where group a varies or not
    a[::2, 21563 + 3 * 51] = ord("A"); a[1::2, 21563 + 3 * 51] = ord("C")
    a[:3, 21563 + 3 * 51] = ord("N")                   # ... and the first defined codon of group a is not in its first rows
    eng.set_products(products, 29903)
    eng.load_group(0, a)
    eng.load_group(1, b)
    for i, p in enumerate(products):
        ra, ca = _oracle_product(a, p)
        rb, cb = _oracle_product(b, p)
        od = orc.other_distinct(np.concatenate([ra, rb]), np.concatenate([ca, cb]))
        want = orc.between_hist(orc.hist(ca), orc.hist(cb), od)
        got = eng.between(0, 1, i)
        assert (got["poly"] == want["poly"]).all()
        assert (got["ndef_a"] == want["ndef_a"]).all() and (got["ndef_b"] == want["ndef_b"]).all()
        # one group undefined at a codon where the other varies: the reference takes the first defined
        # codon in sequence order (bg:853-890), which the histogram alone cannot give
        amb = (want["poly"] == 0) & (np.minimum(want["ndef_a"], want["ndef_b"]) == 0) & (np.maximum(want["ndef_a"], want["ndef_b"]) > 0)
        for k in STATS:
            assert close(got[k][~amb], want[k][~amb]), (p.name, k)
        sites, _ = orc.tables()
        for c in np.flatnonzero(amb):
            col = ca[:, c] if want["ndef_a"][c] > 0 else cb[:, c]
            first = int(col[col != orc.UNDEF][0])
            exp = sites[first] if first < 64 else (0.0, 0.0)
            assert int(got["conserved"][c]) == first
            assert got["N_sites"][c] == exp[0] and got["S_sites"][c] == exp[1]
            assert got["N_diffs"][c] == 0 and got["S_diffs"][c] == 0


# ----------------------------------------------------------------- resampling

def _stats_from_golden(case="w_gaps", product="geneB"):
    d, aln, codon_rows, prod_rows = refio.within_case(case)
    p = [x for x in hostio.read_gtf_products(os.path.join(d, "cds.gtf")) if x.name == product][0]
    return refio.stats4_of(refio.codon_arrays(codon_rows, p.name, p.n_codons)), prod_rows[p.name]


def test_bootstrap_product_statistics(eng_boot):
    eng = eng_boot
    stats4, pr = _stats_from_golden()
    C, B = stats4.shape[0], 40000
    sums, se = eng.bootstrap_product(stats4, B, want_sums=True)
    # each replicate draws C times from C+1 slots (slot 0 empty): E[sum] = C/(C+1) * total
    expect = stats4.sum(axis=0) * C / (C + 1)
    sd = np.sqrt(C * (np.mean(stats4 ** 2, axis=0) * C / (C + 1) - (stats4.mean(axis=0) * C / (C + 1)) ** 2))
    assert np.all(np.abs(sums.mean(axis=0) - expect) < 5 * sd / np.sqrt(B) + 1e-12)
    # SE recomputed on the host from the replicate sums (the reference's formulas)
    dn = np.where(sums[:, 0] > 0, sums[:, 2] / sums[:, 0], 0)
    ds = np.where(sums[:, 1] > 0, sums[:, 3] / sums[:, 1], 0)
    assert close(se[0], np.std(dn, ddof=1), rtol=1e-9) and close(se[1], np.std(ds, ddof=1), rtol=1e-9)
    assert close(se[2], np.std(dn - ds, ddof=1), rtol=1e-9)
    jn, js = -0.75 * np.log(1 - 4 / 3 * dn), -0.75 * np.log(1 - 4 / 3 * ds)
    assert close(se[3], np.std(jn, ddof=1), rtol=1e-9) and close(se[5], np.std(jn - js, ddof=1), rtol=1e-9)
    # against the oracle's own stream and the reference's B=200 run
    _, ose = orc.bootstrap_product(stats4, B, seed=9)
    assert np.all(np.abs(se / ose - 1) < 0.03)
    for mine, nm in zip(se[:3], ("SE_dN", "SE_dS", "SE_dN_minus_dS")):
        assert abs(mine / float(pr[nm]) - 1) < 0.2
    # deterministic per (seed, call), different across calls
    _, se2 = eng.bootstrap_product(stats4, B)
    assert not np.array_equal(se, se2) and np.all(np.abs(se / se2 - 1) < 0.03)


@pytest.mark.parametrize("C", [1, 5, 6, 7, 8, 13, 64, 203, 1000])
def test_bootstrap_draws_are_uniform_over_slots(eng_boot, C):
    """Every one of the C+1 slots (slot 0 = the reference's empty codon, wg:885) is drawn with probability
    1/(C+1), jointly multinomial: the replicate sum of an indicator row counts the draws of that slot.  (For the
    class-aggregated kernel an indicator row is a class of one codon and the zero rows are class 0: the test then checks
    the binomial chain.)"""
    eng = eng_boot
    B = 40000
    p = 1.0 / (C + 1)
    mean, var = C * p, C * p * (1 - p)
    rng = np.random.default_rng(C)
    order = rng.permutation(C)
    codons = order if C <= 13 else order[:12]          # every slot for the small tables, a sample otherwise
    for i in range(0, len(codons), 4):
        t = codons[i:i + 4]
        stats4 = np.zeros((C, 4))
        for q, c in enumerate(t):
            stats4[c, q] = 1.0
        sums, _ = eng.bootstrap_product(stats4, B, want_sums=True)
        assert np.all(sums == np.round(sums)) and sums.min() >= 0 and sums.sum(axis=1).max() <= C
        for q in range(len(t)):
            assert abs(sums[:, q].mean() - mean) < 5 * np.sqrt(var / B), (C, t[q])
            assert abs(sums[:, q].var(ddof=1) / var - 1) < 0.06, (C, t[q])
        if len(t) >= 2:     # two slots of one replicate: Cov = -C p^2
            cov = np.cov(sums[:, 0], sums[:, 1])[0, 1]
            assert abs(cov + C * p * p) < 6 * var / np.sqrt(B), (C, cov)
    # all codons together: C draws minus those of the empty slot
    ones = np.zeros((C, 4))
    ones[:, 0] = 1.0
    sums, _ = eng.bootstrap_product(ones, B, want_sums=True)
    assert sums[:, 0].max() <= C and abs(sums[:, 0].mean() - C * C * p) < 5 * np.sqrt(var / B)
    assert abs(sums[:, 0].var(ddof=1) / var - 1) < 0.06


@pytest.mark.parametrize("C", [1, 2, 3, 5, 31, 62, 127, 128, 129, 500, 1274, 4406, 6300, 6500])
def test_bootstrap_kernels_agree(C):
    """The weight-matrix kernel (slot counters + FP64 tensor-core contraction, the default) and the gather kernel
    consume the same Philox stream draw for draw: replicate sums equal to rounding, SEs with them.  Covers one to
    eight replicates per warp, partial quads, the largest table that fits and one that does not (C = 6500: both
    take the global-memory gathers)."""
    from snpgenie_b200 import _abi
    from snpgenie_b200.engine import Engine
    rng = np.random.default_rng(C)
    sites, _ = orc.tables()
    stats4 = np.zeros((C, 4))
    conserved = rng.random(C) < 0.7
    stats4[:, :2] = np.where(conserved[:, None], np.asarray(sites)[rng.integers(0, 64, size=C)], rng.random((C, 2)) * 3)
    stats4[:, 2:] = np.where(conserved[:, None], 0.0, rng.random((C, 2)))
    B = 300
    out = []
    for kernel in (_abi.BOOT_GATHER, _abi.BOOT_WEIGHTS):
        with Engine(device=0, seed=77) as e:
            e.set_option(_abi.OPT_BOOTSTRAP_KERNEL, kernel)
            out.append(e.bootstrap_product(stats4, B, want_sums=True))
            keep = (np.arange(C) % 3 != 1).astype(np.uint8)
            out.append(e.bootstrap_product(stats4, B, keep=keep, want_sums=True))
    for (s0, se0), (s1, se1) in ((out[0], out[2]), (out[1], out[3])):
        assert np.allclose(s0, s1, rtol=1e-12, atol=1e-12), C
        assert np.allclose(se0, se1, rtol=1e-9, atol=1e-15, equal_nan=True), C     # JC of a random dN > 3/4 is NaN in both
    assert out[0][0][:, 0].max() > 0 or C < 3


@pytest.mark.parametrize("C,frac_general", [(40, 0.3), (420, 0.3), (1274, 0.05), (4406, 0.3), (6300, 0.3), (6300, 0.0), (300, 1.0)])
def test_bootstrap_classes_match_draw_by_draw_distribution(C, frac_general):
    """SNPG_BOOT_CLASSES against SNPG_BOOT_WEIGHTS on tables shaped like a product's (conserved codons in a few classes
    of identical (N_sites, S_sites), the rest with differences): different uses of the Philox stream, the same
    distribution -- means of the replicate sums within 5 standard errors, variances within 8 %, SEs within 5 %.  Class
    sums are whole multiples when there is one class, the result does not depend on anything but (seed, call), and more
    than kBootMaxClasses distinct conserved rows are handled (the smallest groups are drawn one by one)."""
    from snpgenie_b200 import _abi
    from snpgenie_b200.engine import Engine
    rng = np.random.default_rng(C + int(100 * frac_general))
    sites, _ = orc.tables()
    sites = np.asarray(sites)
    stats4 = np.zeros((C, 4))
    general = rng.random(C) < frac_general
    stats4[:, :2] = np.where(general[:, None], rng.random((C, 2)) * 3, sites[rng.integers(0, 64, size=C)])
    stats4[:, 2:] = np.where(general[:, None], rng.random((C, 2)), 0.0)
    if C == 420:                                # 40 distinct conserved rows: more than the classes a product may have
        odd = np.flatnonzero(~general)[:40]
        stats4[odd, 0] = 2.0 + np.arange(len(odd)) / 64.0
        stats4[odd, 1] = 1.0 - np.arange(len(odd)) / 64.0
    B = 20000
    res = {}
    for kernel in (_abi.BOOT_CLASSES, _abi.BOOT_WEIGHTS, _abi.BOOT_CLASSES):
        with Engine(device=0, seed=4242) as e:
            e.set_option(_abi.OPT_BOOTSTRAP_KERNEL, kernel)
            res.setdefault(kernel, []).append(e.bootstrap_product(stats4, B, want_sums=True))
    (s_c, se_c), (s_c2, se_c2) = res[_abi.BOOT_CLASSES]
    s_w, se_w = res[_abi.BOOT_WEIGHTS][0]
    assert np.array_equal(s_c, s_c2) and np.array_equal(se_c, se_c2)          # deterministic
    assert not np.array_equal(s_c, s_w)
    for q in range(4):
        sd = s_w[:, q].std(ddof=1)
        if sd == 0:
            assert s_c[:, q].std() == 0 and np.allclose(s_c[:, q], s_w[:, q])
            continue
        assert abs(s_c[:, q].mean() - s_w[:, q].mean()) < 5 * sd * np.sqrt(2 / B), (C, q)
        assert abs(s_c[:, q].var(ddof=1) / s_w[:, q].var(ddof=1) - 1) < 0.08, (C, q)
    ok = np.isfinite(se_w) & (se_w > 0)
    assert np.all(np.abs(se_c[ok] / se_w[ok] - 1) < 0.05), (se_c, se_w)
    # N_sites + S_sites of a replicate: the cross term between general draws and class counts is right when the total
    # behaves as for draw-by-draw sampling
    t_c, t_w = s_c[:, 0] + s_c[:, 1], s_w[:, 0] + s_w[:, 1]
    assert abs(t_c.var(ddof=1) / t_w.var(ddof=1) - 1) < 0.08
    # one class only: every replicate sum is a whole multiple of the class's row
    one = np.tile(np.array([[2.5, 0.5, 0.0, 0.0]]), (C, 1))
    with Engine(device=0, seed=1) as e:
        s1, _ = e.bootstrap_product(one, 4000, want_sums=True)
    k = s1[:, 0] / 2.5
    assert np.all(k == np.round(k)) and np.all(s1[:, 1] == k * 0.5) and k.max() <= C and np.all(s1[:, 2:] == 0)
    p0 = 1.0 / (C + 1)
    assert abs((C - k).mean() - C * p0) < 5 * np.sqrt(C * p0 * (1 - p0) / 4000)      # draws of the empty slot


def test_bootstrap_keep_mask(eng_boot):
    eng = eng_boot
    stats4, _ = _stats_from_golden()
    keep = (np.arange(stats4.shape[0]) % 3 != 0).astype(np.uint8)
    sums, se = eng.bootstrap_product(stats4, 5000, keep=keep, want_sums=True)
    masked = stats4 * keep[:, None]
    sums2, se2 = eng.bootstrap_product(masked, 5000, want_sums=True)
    assert np.all(np.abs(se / se2 - 1) < 0.1)
    assert sums.max() <= masked.max() * stats4.shape[0]


def test_windows_bitwise_and_se(eng):
    stats4, _ = _stats_from_golden("w_cfg1", "ORF2")
    for W, step in ((10, 3), (100, 10), (1, 1), (stats4.shape[0], 1)):
        sums, _ = eng.windows(stats4, W, step)
        osums, _ = orc.windows(stats4, W, step)
        assert sums.shape == osums.shape
        assert np.array_equal(sums, osums), (W, step)      # same additions, same order
    sums, se = eng.windows(stats4, 100, 10, B=3000)
    _, ose = orc.windows(stats4, 100, 10, B=3000, seed=4)
    for q in range(3):
        ok = ose[:, q] > 0
        assert np.median(np.abs(se[ok, q] / ose[ok, q] - 1)) < 0.05
        assert np.all(se[~ok, q] == 0)
    # the reference's own processor output (w_gaps, W=10 step=3, B=200)
    d, aln, codon_rows, _ = refio.within_case("w_gaps")
    _, wrows = refio.read_tsv(os.path.join(d, "sw_10codons_results.txt"))
    for p in hostio.read_gtf_products(os.path.join(d, "cds.gtf")):
        st = refio.stats4_of(refio.codon_arrays(codon_rows, p.name, p.n_codons))
        sums, se = eng.windows(st, 10, 3, B=4000)
        rows = [r for r in wrows if r["product"] == p.name]
        assert len(rows) == sums.shape[0]
        for k, r in enumerate(rows):
            for q, nm in enumerate(STATS):
                assert close(sums[k, q], float(r[nm]))
        ref_se = np.array([float(r["SE_dN_minus_dS"]) if r["SE_dN_minus_dS"] not in ("NA", "") else 0.0 for r in rows])
        ok = ref_se > 0
        assert 0.8 < np.median(se[ok, 2] / ref_se[ok]) < 1.25


def test_between_processor_filters_via_keep_mask(eng):
    d = os.path.join(refio.GOLD, "b_3groups")
    _, rows = refio.read_tsv(os.path.join(d, "between_group_codon_results.txt"))
    _, frows = refio.read_tsv(os.path.join(d, "between_group_product_results_filtered.txt"))
    _, wrows = refio.read_tsv(os.path.join(d, "between_group_sw_8codons_results.txt"))
    for fr in frows[:4]:
        sel = [r for r in rows if r["product"] == fr["product"] and sorted((r["group_1"], r["group_2"])) == [fr["group_1"], fr["group_2"]]]
        sel.sort(key=lambda r: int(r["codon"].replace("codon_", "")))
        stats4 = np.array([[float(r[k]) for k in STATS] for r in sel])
        n1 = np.array([int(r["num_defined_codons_g1"] or 0) for r in sel])
        n2 = np.array([int(r["num_defined_codons_g2"] or 0) for r in sel])
        keep = (((n1 + n2) >= 9) & (np.minimum(n1, n2) >= 4)).astype(np.uint8)
        _, se = eng.bootstrap_product(stats4, 20000, keep=keep)
        for mine, nm in zip(se, ("SE_dN", "SE_dS", "SE_dN_minus_dS", "SE_JC_dN", "SE_JC_dS", "SE_JC_dN_minus_dS")):
            assert abs(mine / float(fr[nm]) - 1) < 0.2, (fr["product"], nm)
        sums, _ = eng.windows(stats4, 8, 2, keep=keep)
        wsel = [r for r in wrows if r["product"] == fr["product"] and (r["group_1"], r["group_2"]) == (fr["group_1"], fr["group_2"])]
        assert len(wsel) == sums.shape[0]
        for k, r in enumerate(wsel):
            for q, nm in enumerate(STATS):
                assert close(sums[k, q], 0.0 if r[nm] == "NA" else float(r[nm]))


def test_within_all_bootstrap_matches_per_product_calls(eng):
    d, aln, codon_rows, prod_rows = refio.within_case("w_cfg1")
    products = hostio.read_gtf_products(os.path.join(d, "cds.gtf"))
    eng.set_products(products, aln.shape[1])
    eng.load_group(0, aln)
    sums, se = eng.within_all(0, B=20000)
    for i, p in enumerate(products):
        pr = prod_rows[p.name]
        for mine, nm in zip(se[i, :3], ("SE_dN", "SE_dS", "SE_dN_minus_dS")):
            assert abs(mine / float(pr[nm]) - 1) < 0.25, (p.name, nm, mine, pr[nm])   # reference used B=100


# ----------------------------------------------------------------- errors

def test_error_paths(eng):
    from snpgenie_b200 import _abi
    with pytest.raises(_abi.SnpgError) as e:
        eng.set_products([hostio.Product("bad", 1, [1], [10])], 100)
    assert e.value.code == _abi.E_INPUT and "complete codons" in str(e.value)
    with pytest.raises(_abi.SnpgError) as e:
        eng.set_products([hostio.Product("out", 1, [95], [103])], 100)
    assert e.value.code == _abi.E_INPUT
    eng.set_products([hostio.Product("ok", 1, [1], [9])], 100)
    with pytest.raises(_abi.SnpgError) as e:
        eng.load_group(0, np.zeros((4, 50), dtype=np.uint8))
    assert e.value.code == _abi.E_INPUT
    with pytest.raises(_abi.SnpgError) as e:
        eng.within(7, 0)
    assert e.value.code == _abi.E_STATE
    with pytest.raises(_abi.SnpgError) as e:
        eng.windows(np.zeros((5, 4)), 10, 1)
    assert e.value.code == _abi.E_INPUT


# ----------------------------------------------------------------- full size (BASELINE config 2)

def test_config2_full_size_properties(eng2):
    """10,000 x 29,903, 10 CDS: size-independent properties + exact checks on sampled columns."""
    eng = eng2
    products = synth.sars2_products()
    n, L = 10000, 29903
    aln = synth.make_alignment(n, L, products, 20261021)
    eng.set_products(products, L)
    eng.load_group(0, aln)
    sums, se = eng.within_all(0, B=2000)
    sites, diffs = orc.tables()
    rng = np.random.default_rng(0)
    tot_codons = 0
    for i, p in enumerate(products):
        h, od = eng.hist(0, i)
        tot_codons += h.shape[0]
        assert (h.sum(axis=1) == n).all()                       # every sequence lands in exactly one bin
        assert h[:, 64].sum() == 0 and h[:, 65].sum() == 0      # no gaps / ambiguity in this config
        got = eng.within(0, i)
        want = orc.within_hist(h, od, sites, diffs)              # closed form on the device's own histogram
        assert (got["poly"] == want["poly"]).all()
        for k in STATS:
            assert close(got[k], want[k]), (p.name, k)
        # sampled columns: histogram recomputed from the raw bytes on the host
        pos = orc.product_positions(p.starts, p.stops, p.strand, L)
        for c in rng.choice(h.shape[0], size=min(40, h.shape[0]), replace=False):
            raw, codes = orc.extract(aln, pos[3 * c:3 * c + 3], 1)
            assert (orc.hist(codes)[0] == h[c]).all()
        assert close(sums[i], got["stats4"].sum(axis=0), rtol=1e-13)
        assert np.all(se[i, :3] > 0)
    assert tot_codons == 9677


def test_fused_route_irregular_codons_and_many_layers():
    """Codons split across CDS segments and > 4 products over one span take the compact K1+K2 route."""
    from snpgenie_b200 import _abi
    from snpgenie_b200.engine import Engine
    rng = np.random.default_rng(3)
    L, n = 400, 300
    aln = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(n, L))]
    aln[rng.random((n, L)) < 0.03] = ord("-")
    aln[rng.random((n, L)) < 0.01] = ord("R")
    aln[1:, 100:200] = aln[0, 100:200]                     # a conserved stretch (fast path) ...
    aln[5, 120] = ord("a")                                  # ... with one lower-case byte (decode path, same code)
    products = [
        hostio.Product("split", 1, [11, 40, 70], [20, 47, 78]),     # 10+8+9 nt: two codons straddle segment joins
        hostio.Product("f0", 1, [101], [190]), hostio.Product("f1", 1, [102], [191]), hostio.Product("f2", 1, [103], [192]),
        hostio.Product("r0", -1, [101], [190]), hostio.Product("r1", -1, [102], [191]), hostio.Product("r2", -1, [103], [192]),
        hostio.Product("tiny", 1, [300, 310, 320], [301, 311, 324]),  # 2+2+5 nt
        hostio.Product("end", 1, [392], [400]),                      # touches the last alignment column
    ]
    with Engine(device=0) as e:
        e.set_option(_abi.OPT_KEEP_CODES, 0)
        e.set_products(products, L)
        e.load_group(0, aln)
        for i, p in enumerate(products):
            raw, codes = _oracle_product(aln, p)
            h, od = e.hist(0, i)
            assert (h == orc.hist(codes)).all(), p.name
            assert (od == orc.other_distinct(raw, codes)).all(), p.name
        with pytest.raises(_abi.SnpgError) as err:
            e.codon_indices(0, 0)
        assert err.value.code == _abi.E_STATE


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("keep", [1, 0])
def test_codon_range_shards_reassemble(world, keep):
    """Every rank's ctx owns a slice of the global codon axis; concatenated slices == the unsharded run.

    (Ranks are emulated as successive contexts on one GPU; the exchange itself is tested with gloo on CPU.)"""
    from snpgenie_b200 import _abi
    from snpgenie_b200.engine import Engine
    products = synth.hiv_products()
    aln = synth.make_alignment(600, 9719, products, 31, gap_frac=0.02, n_frac=0.02, iupac_frac=0.004)
    with Engine(device=0, keep_codes=True) as full:
        full.set_products(products, 9719)
        full.load_group(0, aln)
        ref_hist = [full.hist(0, i)[0] for i in range(len(products))]
        ref_stats = [full.within(0, i)["stats4"] for i in range(len(products))]
    got_hist = [[] for _ in products]
    got_stats = [[] for _ in products]
    covered = 0
    for r in range(world):
        with Engine(device=0, rank=r, world=world) as e:
            e.set_option(_abi.OPT_KEEP_CODES, keep)
            e.set_products(products, 9719)
            first, count, total = e.shard_range()
            assert total == sum(p.n_codons for p in products)
            covered += count
            e.load_group(0, aln)
            for i in range(len(products)):
                if e.local_codons(i):
                    got_hist[i].append(e.hist(0, i)[0])
                    got_stats[i].append(e.within(0, i)["stats4"])
    assert covered == sum(p.n_codons for p in products)
    for i in range(len(products)):
        assert np.array_equal(np.concatenate(got_hist[i]), ref_hist[i])
        assert np.array_equal(np.concatenate(got_stats[i]), ref_stats[i])


# ----------------------------------------------------------------- "next" rows f1: site and variant tables

@pytest.mark.parametrize("keep", [1, 0])
@pytest.mark.parametrize("case", ["w_cfg1", "w_gaps", "w_heavy"])
def test_site_counts_match_reference_site_table(case, keep):
    from snpgenie_b200 import _abi, report
    from snpgenie_b200.engine import Engine
    d, aln, _, _ = refio.within_case(case)
    products = hostio.read_gtf_products(os.path.join(d, "cds.gtf"))
    with Engine(device=0) as e:
        e.set_option(_abi.OPT_KEEP_CODES, keep)
        e.set_option(_abi.OPT_SITE_COUNTS, 1)
        e.set_products(products, aln.shape[1])
        e.load_group(0, aln)
        counts = e.site_counts(0, aln.shape[1])
        for i, p in enumerate(products):                       # the codon path is unaffected by the wider staging
            assert (e.hist(0, i)[0] == orc.hist(_oracle_product(aln, p)[1])).all()
    assert np.array_equal(counts, orc.site_counts(aln))
    tab = report.site_table(counts)
    _, rows = refio.read_tsv(os.path.join(d, "within_group_site_results.txt"))
    assert len(rows) == aln.shape[1]
    for k, r in enumerate(rows):
        assert [int(r[x]) for x in "ACGT"] == list(counts[k])
        assert r["maj_nt"] == tab["maj_nt"][k] and int(r["maj_nt_count"]) == tab["maj_nt_count"][k]
        assert int(r["coverage"]) == tab["coverage"][k]
        if r["pi"] == "NA":
            assert np.isnan(tab["pi"][k])
        else:
            assert close(tab["pi"][k], float(r["pi"]))


def test_site_counts_large_random_and_device_input():
    import torch
    from snpgenie_b200 import _abi
    from snpgenie_b200.engine import Engine
    rng = np.random.default_rng(8)
    n, L = 3001, 1003
    aln = np.frombuffer(b"ACGTacgtUuNn-RY", dtype=np.uint8)[rng.integers(0, 15, size=(n, L))]
    aln[:, 100:400] = aln[0, 100:400]                          # long conserved stretch
    products = [hostio.Product("p", 1, [4], [1002])]
    pitch = 1024      # 16-byte multiple that covers the last span plus the 4 look-ahead bytes
    dbuf = torch.zeros((n, pitch), dtype=torch.uint8, device="cuda:0")
    dbuf[:, :L] = torch.from_numpy(aln).cuda()
    with Engine(device=0) as e:
        e.set_option(_abi.OPT_KEEP_CODES, 0)
        e.set_option(_abi.OPT_SITE_COUNTS, 1)
        e.set_products(products, L)
        e.load_group(0, aln)
        e.load_group_ptr(1, dbuf.data_ptr(), n, L, pitch, device=True)
        want = orc.site_counts(aln)
        assert np.array_equal(e.site_counts(0, L), want)
        assert np.array_equal(e.site_counts(1, L), want)
        assert (e.hist(0, 0)[0] == e.hist(1, 0)[0]).all()
    with Engine(device=0) as e2:                                # not requested -> loud error
        e2.set_products(products, L)
        e2.load_group(0, aln)
        with pytest.raises(_abi.SnpgError) as err:
            e2.site_counts(0, L)
        assert err.value.code == _abi.E_STATE


def test_variant_table_from_histograms(eng):
    from snpgenie_b200 import report
    for case, exact_counts in (("w_cfg1", True), ("w_gaps", False)):
        d, aln, codon_rows, _ = refio.within_case(case)
        products = hostio.read_gtf_products(os.path.join(d, "cds.gtf"))
        _, vrows = refio.read_tsv(os.path.join(d, "within_group_variant_results.txt"))
        gold = {(r["product_name"], int(r["codon_num"])): r for r in vrows}
        eng.set_products(products, aln.shape[1])
        eng.load_group(0, aln)
        n_rows = 0
        for i, p in enumerate(products):
            h, _ = eng.hist(0, i)
            res = eng.within(0, i)
            raw, codes = _oracle_product(aln, p)
            for c in np.flatnonzero(res["poly"]):
                others = {}
                for s in range(raw.shape[0]):
                    if codes[s, c] == orc.OTHER:
                        k = bytes(raw[s, c]).decode()
                        others[k] = others.get(k, 0) + 1
                row = report.variant_row(h[c], others)
                g = gold[(p.name, c + 1)]
                assert g["codons"].split(",") == row["codons"]
                assert g["amino_acids"].split(",") == row["amino_acids"]
                assert g["variant_type"] == row["variant_type"] and int(g["num_alleles"]) == row["num_alleles"]
                if exact_counts:      # without gaps the reference's counts are the true counts
                    assert [int(x) for x in g["codon_counts"].split(",")] == row["codon_counts"]
                    assert g["consensus_codon"] == row["consensus_codon"]
                    assert int(g["max_codon_count"]) == row["max_codon_count"] and int(g["min_codon_count"]) == row["min_codon_count"]
                    assert int(g["num_defined_seqs"]) == row["num_defined_seqs"]
                    assert g["codon_freqs"] == ",".join("%.3f" % f for f in row["codon_freqs"])
                else:                 # quirk q1: the reference's counts are scaled by (n_true-1)/(n_truncated-1); ours are integers
                    assert row["num_defined_seqs"] == int(res["n_defined"][c])
                n_rows += 1
        assert n_rows == len(gold)


def test_config4_full_size_between_properties(eng):
    """BASELINE config 4: 4 groups x 5,000 x 29,903, all 6 group pairs, every product."""
    products = synth.sars2_products()
    L, n = 29903, 5000
    base = synth.make_alignment(n, L, products, 404)
    rng = np.random.default_rng(4)
    nt = np.frombuffer(b"ACGT", dtype=np.uint8)
    eng.set_products(products, L)
    for g in range(4):
        a = base if g == 0 else synth.make_alignment(n, L, products, 404, p_poly=0.3 + 0.05 * g)
        if g:
            cols = rng.integers(300, 29500, size=200)            # group-specific fixed differences
            a[:, cols] = nt[rng.integers(0, 4, size=200)]
        eng.load_group(g, a)
        del a
    sites, diffs = orc.tables()
    hists = {g: [eng.hist(g, i) for i in range(len(products))] for g in range(4)}
    for i, p in enumerate(products):
        w0 = eng.within(0, i)
        self_pair = eng.between(0, 0, i)
        pol = w0["poly"] == 1
        # a group against itself counts every ordered pair incl. (s,s): n^2 comparisons vs C(n,2)
        assert close(self_pair["N_diffs"][pol] * n / (n - 1), w0["N_diffs"][pol]) and close(self_pair["S_diffs"][pol] * n / (n - 1), w0["S_diffs"][pol])
        for a in range(4):
            for b in range(a + 1, 4):
                ab, ba = eng.between(a, b, i), eng.between(b, a, i)
                assert (ab["ndef_a"] == n).all() and (ab["ndef_b"] == n).all()
                assert (ab["poly"] == ba["poly"]).all()
                assert (ab["comparisons"][ab["poly"] == 1] == n * n).all()
                for k in STATS:
                    assert close(ab[k], ba[k], rtol=1e-13), (p.name, a, b, k)
                (ha, oa), (hb, ob) = hists[a][i], hists[b][i]
                want = orc.between_hist(ha, hb, np.maximum(oa, ob), sites, diffs)
                assert (ab["poly"] == want["poly"]).all()
                for k in STATS:
                    assert close(ab[k], want[k]), (p.name, a, b, k)


def test_between_all_matches_per_pair_calls(eng):
    """snpg_between_all (every group pair x every product, filters, sums, bootstrap in one call) against the
    per-(pair, product) calls plus host-side filtering -- the route the Perl processor takes."""
    products = synth.hiv_products()
    L = 9719
    rng = np.random.default_rng(77)
    groups = []
    for g, n in enumerate((40, 25, 31)):
        a = synth.make_alignment(n, L, products, 700 + g, gap_frac=0.06, n_frac=0.03, iupac_frac=0.004)
        cols = rng.integers(800, 9000, size=60)
        a[:, cols] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=60)]
        groups.append(a)
    eng.set_products(products, L)
    for g, a in enumerate(groups):
        eng.load_group(g, a)
    pairs = [(0, 1), (0, 2), (1, 2)]
    for tot, grp in ((0, 0), (60, 0), (0, 24), (58, 23)):
        sums, se = eng.between_all([0, 1, 2], B=4000, min_taxa_total=tot, min_taxa_group=grp)
        assert sums.shape == (3, len(products), 4) and se.shape == (3, len(products), 6)
        for k, (a, b) in enumerate(pairs):
            for i, p in enumerate(products):
                r = eng.between(a, b, i)
                n1, n2 = r["ndef_a"].astype(np.int64), r["ndef_b"].astype(np.int64)
                keep = np.ones(len(n1), dtype=bool)
                if tot:
                    keep &= (n1 + n2) >= tot
                if grp:
                    keep &= np.minimum(n1, n2) >= grp
                want = (r["stats4"] * keep[:, None]).sum(axis=0)
                assert close(sums[k, i], want, rtol=1e-12, atol=1e-12), (tot, grp, a, b, p.name)
        # SE of one pair/product against the stand-alone bootstrap with the same keep mask
        r = eng.between(0, 2, 1)
        n1, n2 = r["ndef_a"].astype(np.int64), r["ndef_b"].astype(np.int64)
        keep = ((tot == 0) | ((n1 + n2) >= tot)) & ((grp == 0) | (np.minimum(n1, n2) >= grp))
        _, se1 = eng.bootstrap_product(r["stats4"], 8000, keep=keep.astype(np.uint8))
        ok = se1[:3] > 0
        assert np.all(np.abs(se[1, 1, :3][ok] / se1[:3][ok] - 1) < 0.1), (tot, grp, se[1, 1], se1)
    # no bootstrap requested: sums only
    sums0, se0 = eng.between_all([0, 1, 2], B=0)
    sums1, _ = eng.between_all([0, 1, 2], B=4000)
    assert se0 is None and np.array_equal(sums0, sums1)
    # a different order of groups is a different order of pairs
    sums2, _ = eng.between_all([2, 0], B=0)
    assert close(sums2[0], sums0[1], rtol=1e-13)


@pytest.mark.parametrize("route", ["spans", "fused", "tiles", "codes"])
def test_more_than_2_20_rows(route):
    """Config-5-sized row counts (> 2^20 sequences) cross the per-launch row limit of the device-resident path
    and the row-block staging of the host path; counts must stay exact."""
    import torch
    from snpgenie_b200 import _abi
    from snpgenie_b200.engine import Engine
    n, L = (1 << 20) + 70001, 160
    rng = np.random.default_rng(5)
    anc = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=L)]
    aln = np.repeat(anc[None, :], n, axis=0)
    hit = rng.random((n, L)) < 0.01
    aln[hit] = np.frombuffer(b"ACGTN-", dtype=np.uint8)[rng.integers(0, 6, size=int(hit.sum()))]
    aln[-1, :] = ord("T")                                        # the very last row must be counted
    products = [hostio.Product("a", 1, [2], [151]), hostio.Product("b", -1, [5], [154])]
    want = []
    for p in products:
        raw, codes = _oracle_product(aln, p)
        want.append(orc.hist(codes))
    with Engine(device=0, keep_codes=(route == "codes")) as e:
        e.set_option(_abi.OPT_FUSED_KERNEL, {"tiles": _abi.FUSED_TMA_TILES, "spans": _abi.FUSED_TMA_SPANS}.get(route, _abi.FUSED_LDG_SPANS))
        e.set_products(products, L)
        e.load_group(0, aln)
        for i in range(len(products)):
            assert (e.hist(0, i)[0] == want[i]).all(), (route, "host", i)
        d = torch.from_numpy(aln).cuda()                         # pitch 160: a multiple of 16
        e.load_group_ptr(1, d.data_ptr(), n, L, L, device=True)
        for i in range(len(products)):
            assert (e.hist(1, i)[0] == want[i]).all(), (route, "device", i)
        assert e.group_size(1) == n


@pytest.mark.parametrize("annotation", ["sars2", "hiv"])
def test_within_job_matches_call_by_call(annotation):
    """snpg_within_job folds the per-codon statistics into the fix-up kernel when every codon is regular (sars2) and
    keeps the two kernels otherwise (hiv: a codon split over two CDS segments): either way the product sums are those
    of load + within_all bit for bit, from a host buffer and from a device buffer."""
    import torch
    from snpgenie_b200.engine import Engine
    products, L = (synth.sars2_products(), 29903) if annotation == "sars2" else (synth.hiv_products(), 9719)
    n = 333
    aln = synth.make_alignment(n, L, products, 17, gap_frac=0.02, n_frac=0.01, iupac_frac=0.005)
    pitch = (L + 127) // 128 * 128
    d = torch.zeros((n, pitch), dtype=torch.uint8, device="cuda:0")
    d[:, :L] = torch.from_numpy(aln).cuda()
    with Engine(device=0) as e:
        e.set_products(products, L)
        e.load_group(1, aln)
        plain, _ = e.within_all(1, B=0)
        per_codon = [e.within(1, i)["stats4"] for i in range(len(products))]
        for _ in range(3):                         # plain call, then captured and replayed
            sums, se = e.within_job(0, d.data_ptr(), n, L, pitch, True, B=200)
            assert np.array_equal(sums, plain) and np.all(np.isfinite(se[:, :3]))
        sums, _ = e.within_job(2, aln.ctypes.data, n, L, L, False, B=50)
        assert np.array_equal(sums, plain)
        for i in range(len(products)):             # the call-by-call statistics of the job's group agree too
            assert np.array_equal(e.within(0, i)["stats4"], per_codon[i])
            assert (e.hist(0, i)[0] == e.hist(1, i)[0]).all()


def test_within_job_graph_replay_matches_plain_calls():
    """snpg_within_job on a device-resident alignment is replayed as a CUDA graph from the second identical call:
    same kernels, so the product sums are bit-identical and the SEs (a fresh Philox stream per call) agree
    statistically; options, sizes and profiling changes re-capture."""
    import torch
    from snpgenie_b200 import _abi
    from snpgenie_b200.engine import Engine
    products = synth.sars2_products()
    L, n = 29903, 700
    pitch = (L + 127) // 128 * 128
    aln = synth.make_alignment(n, L, products, 99)
    d = torch.zeros((n, pitch), dtype=torch.uint8, device="cuda:0")
    d[:, :L] = torch.from_numpy(aln).cuda()
    with Engine(device=0) as e:
        e.set_products(products, L)
        outs = [e.within_job(0, d.data_ptr(), n, L, pitch, True, B=600) for _ in range(5)]
        assert e.graph_replays == 4
        e.load_group_ptr(1, d.data_ptr(), n, L, pitch, device=True)
        plain, _ = e.within_all(1, B=0)
        for sums, se in outs:
            assert np.array_equal(sums, plain)
            assert np.all(np.isfinite(se)) and np.all(np.abs(se[:, :3] / outs[0][1][:, :3] - 1) < 0.25)
        assert not np.array_equal(outs[3][1], outs[4][1])          # every call draws afresh
        for i in range(len(products)):                              # the graph really loaded the group
            raw, codes = _oracle_product(aln, products[i])
            assert (e.hist(0, i)[0] == orc.hist(codes)).all()
        # profiling on: still a graph, and the kernel clock keeps counting
        e.set_option(_abi.OPT_PROFILE, -1)
        r0 = e.graph_replays
        e.profile(reset=True)
        for _ in range(4):
            sums, _ = e.within_job(0, d.data_ptr(), n, L, pitch, True, B=600)
            assert np.array_equal(sums, plain)
        prof = e.profile()
        assert e.graph_replays == r0 + 3 and prof["fused"][1] == 4 and prof["fused"][0] > 0 and prof["bootstrap"][1] == 4
        # another replicate count, another histogram kernel, a host buffer: each a different job
        e.set_option(_abi.OPT_PROFILE, 0)
        for B in (50, 50, 50):
            sums, se = e.within_job(0, d.data_ptr(), n, L, pitch, True, B=B)
            assert np.array_equal(sums, plain)
        e.set_option(_abi.OPT_FUSED_KERNEL, _abi.FUSED_TMA_TILES)
        t0 = e.tile_launches
        for _ in range(3):
            sums, _ = e.within_job(0, d.data_ptr(), n, L, pitch, True, B=50)
            assert np.array_equal(sums, plain)
        assert e.tile_launches == t0 + 3
        host = np.ascontiguousarray(aln)
        sums, _ = e.within_job(0, host.ctypes.data, n, L, L, False, B=50)
        assert np.array_equal(sums, plain)


# ----------------------------------------------------------------- edge cases

@pytest.mark.parametrize("keep", [True, False])
def test_tiny_and_degenerate_inputs(keep):
    from snpgenie_b200.engine import Engine
    with Engine(device=0, keep_codes=keep) as e:
        # one codon, one sequence; alignment shorter than one 16-byte span
        e.set_products([hostio.Product("one", 1, [2], [4])], 7)
        e.load_group(0, np.frombuffer(b"GATGCAT", dtype=np.uint8).reshape(1, 7).copy())
        h, _ = e.hist(0, 0)
        assert h.sum() == 1 and h[0, 14] == 1                      # ATG = 0*16 + 3*4 + 2
        r = e.within(0, 0)
        assert r["poly"][0] == 0 and r["conserved"][0] == 14 and r["N_sites"][0] == 3 and r["comparisons"][0] == 0
        # every sequence undefined at a codon; a CR inside the sequence is an 'other' residue (chomp strips LF only)
        aln = np.frombuffer(b"NNNACG\rAA" + b"---ACG\rAA" + b"nnnACGTAA", dtype=np.uint8).reshape(3, 9).copy()
        e.set_products([hostio.Product("p", 1, [1], [9])], 9)
        e.load_group(0, aln)
        raw, codes = _oracle_product(aln, hostio.Product("p", 1, [1], [9]))
        h, od = e.hist(0, 0)
        assert (h == orc.hist(codes)).all() and h[0, 64] == 3 and h[2, 65] == 2 and h[2, 48] == 1
        r = e.within(0, 0)
        assert r["conserved"][0] == 64 and (r["stats4"][0] == 0).all() and r["n_defined"][0] == 0
        assert r["poly"][2] == 1 and r["N_diffs"][2] == 0 and r["comparisons"][2] == 3      # TAA vs '\rAA': defined, 0 sites
        lit = orc.within_pairloop(raw)
        for k in STATS:
            assert close(r[k], lit[k])
        # two identical groups: between-group sees conserved codons only where a single allele exists
        e.load_group(1, aln)
        b = e.between(0, 1, 0)
        assert b["poly"][0] == 0 and b["conserved"][0] == 64 and b["ndef_a"][0] == 0
        assert b["poly"][1] == 0 and b["conserved"][1] == 6                                   # ACG


def test_reload_group_with_different_sizes(eng2):
    eng = eng2
    products = synth.hiv_products()
    eng.set_products(products, 9719)
    for n, seed in ((400, 1), (33, 2), (900, 3)):
        aln = synth.make_alignment(n, 9719, products, seed, gap_frac=0.02, n_frac=0.01)
        eng.load_group(0, aln)                                     # same slot, new size: buffers are reused
        for i in (0, 4, 7):
            assert (eng.hist(0, i)[0] == orc.hist(_oracle_product(aln, products[i])[1])).all(), (n, i)


def test_windows_edge_shapes(eng):
    stats4 = np.arange(40, dtype=np.float64).reshape(10, 4) / 7
    for W, step, nwin in ((10, 1, 1), (1, 1, 10), (3, 4, 2), (4, 100, 1)):
        sums, se = eng.windows(stats4, W, step, B=50)
        osums, _ = orc.windows(stats4, W, step)
        assert sums.shape == (nwin, 4) and np.array_equal(sums, osums)
        assert se.shape == (nwin, 3) and np.all(np.isfinite(se))
        if W == 1:
            assert np.all(se == 0)                                 # one codon resampled with itself never varies


# ----------------------------------------------------------------- "next" row f3: native FASTA ingest

def test_native_fasta_reader_matches_reference_semantics(tmp_path, eng):
    from snpgenie_b200 import _abi
    # the golden FASTA files (one wrapped at 80 columns, others single-line)
    for case in ("w_cfg1", "w_gaps", "w_minus"):
        path = os.path.join(refio.GOLD, case, "aln.fasta")
        _, want = hostio.read_fasta(path)
        ptr, n, L = eng.read_fasta(0, path)
        assert (n, L) == want.shape
        assert np.array_equal(eng.fasta_view(ptr, n, L), want)
    # CRLF (the CR stays, as after chomp), '>' in mid-line makes a header, blank lines, no trailing newline
    odd = tmp_path / "odd.fa"
    odd.write_bytes(b">a\r\nACGT\r\nAC\r\n\nxx>b\nACGT\rAN\r\n>c desc\nAC\nGT\rA-\r")
    _, want = hostio.read_fasta(str(odd))
    ptr, n, L = eng.read_fasta(1, str(odd))
    got = eng.fasta_view(ptr, n, L)
    assert got.shape == want.shape == (3, 8) and np.array_equal(got, want)
    assert bytes(got[0]) == b"ACGT\rAC\r"
    # errors the reference dies on
    bad = tmp_path / "bad.fa"
    bad.write_bytes(b">a\nACGT\n>b\nACG\n")
    with pytest.raises(_abi.SnpgError) as e:
        eng.read_fasta(2, str(bad))
    assert e.value.code == _abi.E_INPUT and "same length" in str(e.value)
    with pytest.raises(_abi.SnpgError) as e:
        eng.read_fasta(2, str(tmp_path / "missing.fa"))
    assert e.value.code == _abi.E_INPUT
    # the pinned buffer feeds load_group directly; codon strings come back normalised
    d, aln, _, _ = refio.within_case("w_minus")
    products = hostio.read_gtf_products(os.path.join(d, "cds.gtf"), include_minus=True)
    ptr, n, L = eng.read_fasta(0, os.path.join(d, "aln.fasta"))
    eng.set_products(products, L)
    eng.load_group_ptr(0, ptr, n, L, L, device=False)
    for i, p in enumerate(products):
        raw, codes = _oracle_product(aln, p)
        assert (eng.hist(0, i)[0] == orc.hist(codes)).all()
        for c in (0, 7, p.n_codons - 1):
            assert np.array_equal(eng.codon_strings(0, i, c, n), raw[:, c])


@pytest.mark.parametrize("threads", ["1", "3", "16"])
def test_native_fasta_reader_threads_and_background_copy(tmp_path, threads):
    """The reader cuts the file into one byte range per thread (records and lines straddle the cuts) and, with
    OPT_INGEST_ASYNC, copies the rows in the background while the load already moves the first row blocks."""
    import subprocess
    import sys
    # a 40 MB FASTA wrapped at 61 columns with CRs, blank lines and one multi-megabyte header-free stretch per record
    code = r"""
import os, sys, numpy as np
sys.path.insert(0, %r)
from snpgenie_b200 import _abi, synth, hostio
from snpgenie_b200.engine import Engine
from oracle import oracle as orc
rng = np.random.default_rng(7)
n, L = 1300, 29903
products = synth.sars2_products()
aln = synth.make_alignment(n, L, products, 5, gap_frac=0.01, n_frac=0.01)
path = os.path.join(%r, "big.fa")
with open(path, "wb") as fh:
    for r in range(n):
        fh.write(b">seq%%d some description > with a bracket\n" %% r)
        row = aln[r].tobytes()
        w = 61 if r %% 3 else 997
        for k in range(0, L, w):
            fh.write(row[k:k + w] + b"\n")
        if r %% 50 == 0:
            fh.write(b"\n\n")
with Engine(device=0) as eng:
    for async_ in (0, 1):
        eng.set_option(_abi.OPT_INGEST_ASYNC, async_)
        ptr, nn, LL = eng.read_fasta(0, path)
        assert (nn, LL) == (n, L)
        eng.set_products(products, L)
        eng.load_group_ptr(0, ptr, nn, LL, LL, device=False)      # waits row block by row block when the copy is still running
        for i in (0, 3, 9):
            pos = orc.product_positions(products[i].starts, products[i].stops, products[i].strand, L)
            _, codes = orc.extract(aln, pos, products[i].strand)
            assert (eng.hist(0, i)[0] == orc.hist(codes)).all(), (async_, i)
        eng.fasta_wait(0)
        assert np.array_equal(eng.fasta_view(ptr, nn, LL), aln), async_
    # a second file into the same slot while nothing is pending; ragged input still fails in pass 1
    bad = os.path.join(%r, "bad.fa")
    open(bad, "wb").write(b">a\nACGT\nAC\n>b\nACGTA\n")
    try:
        eng.read_fasta(0, bad)
        raise SystemExit("ragged FASTA accepted")
    except _abi.SnpgError as e:
        assert "same length" in str(e)
print("ok")
""" % (ROOT, str(tmp_path), str(tmp_path))
    env = dict(os.environ, SNPG_INGEST_THREADS=threads)
    p = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    assert p.returncode == 0 and "ok" in p.stdout, p.stdout[-3000:]


def test_philox_known_answers(eng):
    """Random123 kat_vectors for philox4x32-10, computed by the device function the bootstraps call."""
    import ctypes as C
    from snpgenie_b200 import _abi
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        c, k, o = np.array(ctr, np.uint32), np.array(key, np.uint32), np.zeros(4, np.uint32)
        rc = _abi.lib.snpg_test_philox(eng._ctx, c.ctypes.data_as(C.POINTER(C.c_uint32)), k.ctypes.data_as(C.POINTER(C.c_uint32)),
                                       o.ctypes.data_as(C.POINTER(C.c_uint32)))
        assert rc == 0 and tuple(int(x) for x in o) == want


# ----------------------------------------------------------------- randomized annotations (index-map fuzz)

def _random_products(rng, L):
    prods = []
    for k in range(int(rng.integers(1, 7))):
        nseg = int(rng.integers(1, 5))
        cuts = np.sort(rng.choice(np.arange(1, L), size=2 * nseg, replace=False))
        starts, stops = list(cuts[0::2] + 0), list(cuts[1::2] + 0)
        total = sum(e - s + 1 for s, e in zip(starts, stops))
        trim = total % 3                      # shorten the last segment to complete the last codon
        while trim and stops[-1] - trim < starts[-1]:
            starts.pop(); stops.pop()
            if not starts:
                break
            total = sum(e - s + 1 for s, e in zip(starts, stops)); trim = total % 3
        if not starts:
            continue
        stops[-1] -= trim
        if sum(e - s + 1 for s, e in zip(starts, stops)) < 3:
            continue
        order = rng.permutation(len(starts))  # segments arrive in file order, the library sorts them
        prods.append(hostio.Product(f"p{k}", int(rng.choice([1, -1])), [int(starts[i]) for i in order], [int(stops[i]) for i in order]))
    return prods


@pytest.mark.parametrize("seed", range(24))
def test_random_annotations_both_routes(seed):
    from snpgenie_b200.engine import Engine
    rng = np.random.default_rng(1000 + seed)
    L = int(rng.integers(20, 400))
    n = int(rng.integers(1, 300))
    prods = _random_products(rng, L)
    if not prods:
        prods = [hostio.Product("p", 1, [1], [3])]
    alphabet = np.frombuffer(b"ACGTACGTACGTacgtUuNn-RY", dtype=np.uint8)
    aln = alphabet[rng.integers(0, len(alphabet), size=(n, L))]
    keep_cols = rng.random(L) < 0.6                            # conserved columns exercise the fast path
    aln[:, keep_cols] = aln[0, keep_cols]
    sorted_prods = [hostio.Product(p.name, p.strand, sorted(p.starts), [e for _, e in sorted(zip(p.starts, p.stops))]) for p in prods]
    from snpgenie_b200 import _abi
    for keep, fused in ((True, 0), (False, _abi.FUSED_TMA_TILES), (False, _abi.FUSED_LDG_SPANS), (False, _abi.FUSED_TMA_SPANS)):
        with Engine(device=0, keep_codes=keep) as e:
            e.set_option(_abi.OPT_FUSED_KERNEL, fused)
            e.set_products(prods, L)
            e.load_group(0, aln)
            for i, p in enumerate(sorted_prods):
                raw, codes = _oracle_product(aln, p)
                h, od = e.hist(0, i)
                assert (h == orc.hist(codes)).all(), (seed, keep, p)
                assert (od == orc.other_distinct(raw, codes)).all()
                if keep:
                    assert (e.codon_indices(0, i) == codes).all()
                got, want = e.within(0, i), orc.within_pairloop(raw)
                assert (got["poly"] == want["poly"]).all()
                for k in STATS:
                    assert close(got[k], want[k]), (seed, keep, p.name, k)

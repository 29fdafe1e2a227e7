"""Pin the CPU oracle against the unmodified reference (goldens from tools/make_goldens.py).

CPU-only.  Tolerance: tables bit-exact; per-codon and per-product values 1e-12
relative (the goldens carry 15 significant digits);
bootstrap standard errors statistically (the reference re-seeds rand() per
replicate, so its stream cannot be replayed).
"""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from snpgenie_b200 import hostio
from tests import refio
from tests.refio import STATS

RTOL = 1e-12
CODONS = [a + b + c for a in "ACGT" for b in "ACGT" for c in "ACGT"]


def close(a, b, rtol=RTOL, atol=1e-15):
    return np.allclose(a, b, rtol=rtol, atol=atol)


def test_tables_bit_exact(golden_dir):
    sites, diffs = orc.tables()
    S = np.zeros((64, 3, 2))
    D = np.zeros((2, 64, 64))
    for line in open(os.path.join(golden_dir, "ng_tables.tsv")):
        f = line.rstrip("\n").split("\t")
        if f[0] == "S":
            S[CODONS.index(f[1]), int(f[2]) - 1] = (float(f[3]), float(f[4]))
        else:
            D[:, CODONS.index(f[1]), CODONS.index(f[2])] = (float(f[3]), float(f[4]))
    ref_sites = S[:, 0] + S[:, 1] + S[:, 2]          # callers sum pos1+pos2+pos3 (wg:636-646)
    assert (ref_sites == sites).all()
    assert (D == diffs).all()
    # SURVEY appendix C known answers
    idx = CODONS.index
    assert diffs[0, idx("AAA"), idx("TGC")] == 3 and diffs[1, idx("AAA"), idx("TGC")] == 0
    assert diffs[0, idx("AAA"), idx("TTC")] == 2.75
    assert (diffs[:, idx("TAA")] == 0).all() and (diffs[:, :, idx("TGA")] == 0).all()
    assert (diffs[0] == diffs[0].T).all() and (diffs[1] == diffs[1].T).all()
    assert int((diffs[0] + diffs[1] == 0).sum()) - 64 == 372   # STOP-zeroed ordered entries


def _check_within(case, products=None):
    d, aln, codon_rows, prod_rows = refio.within_case(case)
    if products is None:
        products = hostio.read_gtf_products(os.path.join(d, "cds.gtf"))
    sites, diffs = orc.tables()
    for p in products:
        pos = orc.product_positions(p.starts, p.stops, p.strand, aln.shape[1])
        raw, codes = orc.extract(aln, pos, p.strand)
        ncodon = len(pos) // 3
        gold = refio.codon_arrays(codon_rows, p.name, ncodon)
        res = orc.within_pairloop(raw, sites, diffs)
        assert (res["poly"] == gold["poly"]).all(), case
        for k in STATS:
            assert close(res[k], gold[k]), (case, p.name, k)
        # the closed form from histograms agrees with the literal loop
        h = orc.hist(codes)
        res2 = orc.within_hist(h, orc.other_distinct(raw, codes), sites, diffs)
        assert (res2["poly"] == gold["poly"]).all()
        for k in STATS:
            assert close(res2[k], gold[k]), (case, p.name, k, "hist")
        assert (res2["n_defined"] == res["n_defined"]).all()
        # comparisons column: allele names only (pair order is Perl hash order)
        for c in range(ncodon):
            defined = [bytes(raw[i, c]).decode() for i in range(raw.shape[0]) if codes[i, c] != orc.UNDEF]
            if gold["poly"][c]:
                got = refio.pair_set(gold["comparisons_str"][c])
                assert {x for s in got for x in s} == set(defined)
            else:
                assert gold["comparisons_str"][c] == (defined[0] if defined else "")
        # product line (wg:981-1057)
        pr = prod_rows[p.name]
        tot = {k: res[k].sum() for k in STATS}
        for k in tot:
            assert close(tot[k], float(pr[k])), (case, p.name, k)
        assert close(tot["N_diffs"] / tot["N_sites"], float(pr["dN"]))
        assert close(tot["S_diffs"] / tot["S_sites"], float(pr["dS"]))


@pytest.mark.parametrize("case", ["w_cfg1", "w_gaps", "w_heavy"])
def test_within_matches_reference(case):
    _check_within(case)


def test_minus_strand_matches_revcom_route():
    """Native '-' strand handling == fasta2revcom.pl + gtf2revcom.pl + within driver."""
    d = os.path.join(refio.GOLD, "w_minus")
    products = hostio.read_gtf_products(os.path.join(d, "cds.gtf"), include_minus=True)
    assert sorted(p.strand for p in products) == [-1, -1, 1]
    # the reference saw geneA as '-' (ignored) and revB/revC as '+': only those two have rows
    _check_within("w_minus", [p for p in products if p.strand < 0])


def test_within_bootstrap_se_statistical():
    d, aln, codon_rows, prod_rows = refio.within_case("w_gaps")
    products = hostio.read_gtf_products(os.path.join(d, "cds.gtf"))
    for p in products:
        gold = refio.codon_arrays(codon_rows, p.name, p.n_codons)
        _, se = orc.bootstrap_product(refio.stats4_of(gold), 20000, seed=3)
        pr = prod_rows[p.name]
        # reference used B=200: relative MC error of an SD ~ 1/sqrt(2B) = 5%; allow 4 sigma
        for mine, ref in zip(se[:3], (pr["SE_dN"], pr["SE_dS"], pr["SE_dN_minus_dS"])):
            assert abs(mine - float(ref)) / float(ref) < 0.2, (p.name, mine, ref)


def test_within_windows_match_processor():
    d, aln, codon_rows, _ = refio.within_case("w_gaps")
    _, wrows = refio.read_tsv(os.path.join(d, "sw_10codons_results.txt"))
    products = hostio.read_gtf_products(os.path.join(d, "cds.gtf"))
    for p in products:
        gold = refio.codon_arrays(codon_rows, p.name, p.n_codons)
        sums, se = orc.windows(refio.stats4_of(gold), 10, 3, B=4000, seed=5)
        rows = [r for r in wrows if r["product"] == p.name]
        assert len(rows) == sums.shape[0]
        for k, r in enumerate(rows):
            assert int(r["first_codon"]) == 3 * k + 1 and int(r["last_codon"]) == 3 * k + 10
            for q, nm in enumerate(STATS):
                assert close(sums[k, q], float(r[nm])), (p.name, k, nm)
        ref_se = np.array([float(r["SE_dN_minus_dS"]) if r["SE_dN_minus_dS"] not in ("NA", "") else 0.0 for r in rows])
        ok = ref_se > 0
        ratio = se[ok, 2] / ref_se[ok]
        assert 0.8 < np.median(ratio) < 1.25 and (np.abs(ratio - 1) < 0.35).mean() > 0.9


def _between_inputs():
    d = os.path.join(refio.GOLD, "b_3groups")
    groups = {g: hostio.read_fasta(os.path.join(d, g))[1] for g in ("grpA.fasta", "grpB.fasta", "grpC.fasta")}
    products = hostio.read_gtf_products(os.path.join(d, "cds.gtf"))
    _, rows = refio.read_tsv(os.path.join(d, "between_group_codon_results.txt"))
    maps = refio.between_label_maps(os.path.join(d, "stdout.log"))
    return d, groups, products, rows, maps


def test_between_matches_reference():
    d, groups, products, rows, maps = _between_inputs()
    sites, diffs = orc.tables()
    L = next(iter(groups.values())).shape[1]
    checked = 0
    for p in products:
        pos = orc.product_positions(p.starts, p.stops, p.strand, L)
        raw = {g: orc.extract(a, pos, p.strand) for g, a in groups.items()}
        prow = [r for r in rows if r["product"] == p.name]
        pairs = sorted({(r["group_1"], r["group_2"]) for r in prow})
        assert len(pairs) == 3
        for lab1, lab2 in pairs:
            g1, g2 = maps[p.name][lab1], maps[p.name][lab2]      # undo quirk q2
            res = orc.between_pairloop(raw[g1][0], raw[g2][0], sites, diffs)
            ha, hb = orc.hist(raw[g1][1]), orc.hist(raw[g2][1])
            both_raw = np.concatenate([raw[g1][0], raw[g2][0]])
            both_codes = np.concatenate([raw[g1][1], raw[g2][1]])
            res2 = orc.between_hist(ha, hb, orc.other_distinct(both_raw, both_codes), sites, diffs)
            for r in (x for x in prow if (x["group_1"], x["group_2"]) == (lab1, lab2)):
                c = int(r["codon"].replace("codon_", "")) - 1
                assert int(r["num_defined_codons_g1"] or 0) == res["ndef_a"][c] == res2["ndef_a"][c]
                assert int(r["num_defined_codons_g2"] or 0) == res["ndef_b"][c] == res2["ndef_b"][c]
                assert (r["variability"] == "polymorphic") == bool(res["poly"][c]) == bool(res2["poly"][c])
                # one side empty and the other varying: the reference takes the first defined codon in
                # (hash-random) sequence order -- not a function of the histogram, excluded
                ambiguous = (not res["poly"][c]) and min(res["ndef_a"][c], res["ndef_b"][c]) == 0 \
                    and max((ha[c, :64] > 0).sum() + (ha[c, 65] > 0), (hb[c, :64] > 0).sum() + (hb[c, 65] > 0)) > 1
                for k in STATS:
                    if not ambiguous:
                        assert close(res[k][c], float(r[k])), (p.name, g1, g2, c, k)
                        assert close(res2[k][c], float(r[k])), (p.name, g1, g2, c, k, "hist")
                checked += 1
    assert checked == 3 * sum(p.n_codons for p in products)


def test_between_processor_sums_and_filters():
    d, groups, products, rows, maps = _between_inputs()
    _, frows = refio.read_tsv(os.path.join(d, "between_group_product_results_filtered.txt"))
    _, wrows = refio.read_tsv(os.path.join(d, "between_group_sw_8codons_results.txt"))
    assert len(frows) == 9
    for fr in frows:
        # the processor sorts the two labels of each row (:326) and keys on them
        sel = [r for r in rows if r["product"] == fr["product"]
               and sorted((r["group_1"], r["group_2"])) == [fr["group_1"], fr["group_2"]]]
        sel.sort(key=lambda r: int(r["codon"].replace("codon_", "")))
        stats4 = np.array([[float(r[k]) for k in STATS] for r in sel])
        n1 = np.array([int(r["num_defined_codons_g1"] or 0) for r in sel])
        n2 = np.array([int(r["num_defined_codons_g2"] or 0) for r in sel])
        keep = ((n1 + n2) >= 9) & (np.minimum(n1, n2) >= 4)   # --codon_min_taxa_total=9 --codon_min_taxa_group=4
        masked = stats4 * keep[:, None]
        tot = masked.sum(axis=0)
        for q, nm in enumerate(STATS):
            assert close(tot[q], float(fr[nm])), (fr["product"], nm)
        dN, dS = tot[2] / tot[0], tot[3] / tot[1]
        assert close(dN, float(fr["dN"])) and close(dS, float(fr["dS"]))
        assert close(-0.75 * np.log(1 - 4 / 3 * dN), float(fr["JC_dN"]))
        _, se = orc.bootstrap_product(masked, 20000, seed=11)
        for mine, nm in zip(se, ("SE_dN", "SE_dS", "SE_dN_minus_dS", "SE_JC_dN", "SE_JC_dS", "SE_JC_dN_minus_dS")):
            assert abs(mine - float(fr[nm])) / float(fr[nm]) < 0.2, (fr["product"], nm, mine, fr[nm])
        sums, _ = orc.windows(masked, 8, 2)
        wsel = [r for r in wrows if r["product"] == fr["product"] and (r["group_1"], r["group_2"]) == (fr["group_1"], fr["group_2"])]
        assert len(wsel) == sums.shape[0]
        for k, r in enumerate(wsel):
            for q, nm in enumerate(STATS):
                ref = 0.0 if r[nm] == "NA" else float(r[nm])
                assert close(sums[k, q], ref), (fr["product"], k, nm)

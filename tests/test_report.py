"""Host-side table derivations (report.py) against the reference's variant and site tables, on CPU.

The histograms here come from the oracle, so this file checks only the host logic; the GPU tests
feed the same functions with device histograms.
"""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from snpgenie_b200 import hostio, report
from tests import refio


@pytest.mark.parametrize("case", ["w_cfg1", "w_gaps", "w_heavy"])
def test_site_table_matches_reference(case):
    d, aln, _, _ = refio.within_case(case)
    tab = report.site_table(orc.site_counts(aln))
    _, rows = refio.read_tsv(os.path.join(d, "within_group_site_results.txt"))
    assert len(rows) == aln.shape[1]
    for k, r in enumerate(rows):
        assert (int(r["A"]), int(r["C"]), int(r["G"]), int(r["T"])) == (tab["A"][k], tab["C"][k], tab["G"][k], tab["T"][k])
        assert r["maj_nt"] == tab["maj_nt"][k] and int(r["maj_nt_count"]) == tab["maj_nt_count"][k]
        assert int(r["coverage"]) == tab["coverage"][k]
        if r["pi"] == "NA":
            assert np.isnan(tab["pi"][k])
        else:
            assert np.isclose(tab["pi"][k], float(r["pi"]), rtol=1e-12, atol=0)


def test_variant_rows_match_reference_when_gap_free():
    d, aln, _, _ = refio.within_case("w_cfg1")
    _, vrows = refio.read_tsv(os.path.join(d, "within_group_variant_results.txt"))
    gold = {(r["product_name"], int(r["codon_num"])): r for r in vrows}
    seen = 0
    for p in hostio.read_gtf_products(os.path.join(d, "cds.gtf")):
        pos = orc.product_positions(p.starts, p.stops, p.strand, aln.shape[1])
        raw, codes = orc.extract(aln, pos, p.strand)
        h = orc.hist(codes)
        res = orc.within_hist(h, orc.other_distinct(raw, codes))
        for c in np.flatnonzero(res["poly"]):
            row = report.variant_row(h[c])
            g = gold[(p.name, c + 1)]
            assert g["consensus_codon"] == row["consensus_codon"]
            assert g["codons"].split(",") == row["codons"] and g["amino_acids"].split(",") == row["amino_acids"]
            assert g["variant_type"] == row["variant_type"]
            assert [int(x) for x in g["codon_counts"].split(",")] == row["codon_counts"]
            assert g["codon_freqs"] == ",".join("%.3f" % f for f in row["codon_freqs"])
            assert (int(g["count_total"]), int(g["max_codon_count"]), int(g["min_codon_count"]), int(g["num_defined_seqs"])) == \
                (row["count_total"], row["max_codon_count"], row["min_codon_count"], row["num_defined_seqs"])
            seen += 1
    assert seen == len(gold) > 100


def test_amino_acid_and_variant_types():
    assert report.amino_acid("ATG") == "M" and report.amino_acid("TAA") == "*" and report.amino_acid("ARG") == "?"
    h = np.zeros(66, dtype=np.uint32)
    h[report.CODONS.index("CGA")] = 5
    h[report.CODONS.index("AGA")] = 2            # both R
    assert report.variant_row(h)["variant_type"] == "Synonymous"
    h[report.CODONS.index("GGA")] = 1            # G: now R-R syn and R-G nonsyn pairs
    assert report.variant_row(h)["variant_type"] == "Ambiguous"
    row = report.variant_row(h, {"CGR": 3})
    assert row["codons"] == ["AGA", "CGA", "CGR", "GGA"] and row["amino_acids"] == ["R", "R", "?", "G"]
    assert row["consensus_codon"] == "CGA" and row["num_defined_seqs"] == 11 and row["min_codon_count"] == 6

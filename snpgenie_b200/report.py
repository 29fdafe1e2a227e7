"""Host-side derivations from the device histograms for the reference's two secondary tables.

  within_group_variant_results.txt  (snpgenie_within_group.pl:1061-1174): per polymorphic codon the
      alleles, their counts, amino acids and variant type -- everything follows from the codon's
      66-bin histogram (plus the strings of any 'other' alleles).  Counts here are the TRUE integer
      counts; the reference's are distorted when gaps are present (its n_defined is truncated by
      an early loop exit).
  within_group_site_results.txt     (:1203-1295): per alignment column A/C/G/T counts, majority
      nucleotide, coverage and pi = sum_{x<y} c_x c_y / C(coverage, 2).
"""
from __future__ import annotations

import numpy as np

CODONS = [a + b + c for a in "ACGT" for b in "ACGT" for c in "ACGT"]
_AA = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF"
AMINO_ACID = dict(zip(CODONS, _AA))


def amino_acid(codon: str) -> str:
    """get_amino_acid (wg:1400-1516): '?' for anything that is not a pure ACGT codon."""
    return AMINO_ACID.get(codon, "?")


def variant_row(hist_row: np.ndarray, other_alleles: dict | None = None) -> dict:
    """Fields of one within_group_variant_results.txt row from a codon's histogram."""
    alleles = {CODONS[b]: int(hist_row[b]) for b in range(64) if hist_row[b]}
    if other_alleles:
        alleles.update(other_alleles)
    names = sorted(alleles)                                  # `sort keys` (wg:1080)
    counts = [alleles[k] for k in names]
    aas = [amino_acid(k) for k in names]
    n_def = sum(counts)
    consensus, best = "", 0
    for k, c in zip(names, counts):                           # strict '>' keeps the first maximum (wg:1091-1096)
        if c > best:
            consensus, best = k, c
    syn = non = 0
    for i in range(len(aas)):
        for j in range(i + 1, len(aas)):
            if aas[i] == aas[j]:
                syn += 1
            else:
                non += 1
    vtype = "Ambiguous" if (non and syn) else ("Nonsynonymous" if non else "Synonymous")
    return dict(consensus_codon=consensus, num_alleles=len(names), codons=names, amino_acids=aas, variant_type=vtype,
                codon_counts=counts, count_total=n_def, codon_freqs=[c / n_def for c in counts],
                max_codon_count=best, min_codon_count=n_def - best, num_defined_seqs=n_def)


def site_table(counts: np.ndarray) -> dict:
    """Columns of within_group_site_results.txt from [L,4] A,C,G,T counts."""
    c = counts.astype(np.int64)
    cov = c.sum(axis=1)
    maj = np.argmax(c, axis=1)                                # first maximum in A,C,G,T order == the reference's strict '>' chain
    pw = c[:, 0] * (c[:, 1] + c[:, 2] + c[:, 3]) + c[:, 1] * (c[:, 2] + c[:, 3]) + c[:, 2] * c[:, 3]
    comps = (cov * cov - cov) / 2
    with np.errstate(divide="ignore", invalid="ignore"):
        pi = np.where(comps > 0, pw / comps, np.nan)          # 'NA' in the reference when coverage < 2
    return dict(maj_nt=np.array(list("ACGT"))[maj], maj_nt_count=c[np.arange(len(c)), maj], pi=pi, coverage=cov,
                A=c[:, 0], C=c[:, 1], G=c[:, 2], T=c[:, 3])

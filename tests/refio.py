"""Parsers for the reference's TSV outputs frozen under tests/golden/."""
from __future__ import annotations

import os
import re

import numpy as np

from snpgenie_b200 import hostio

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STATS = ("N_sites", "S_sites", "N_diffs", "S_diffs")


def read_tsv(path):
    with open(path) as fh:
        header = fh.readline().rstrip("\n").split("\t")
        rows = [dict(zip(header, line.rstrip("\n").split("\t"))) for line in fh if line.strip()]
    return header, rows


def within_case(case):
    d = os.path.join(GOLD, case)
    _, aln = hostio.read_fasta(os.path.join(d, "aln.fasta"))
    _, rows = read_tsv(os.path.join(d, "within_group_codon_results.txt"))
    codon = {(r["product"], int(r["codon"])): r for r in rows}
    _, prows = read_tsv(os.path.join(d, "within_group_product_results.txt"))
    return d, aln, codon, {r["product"]: r for r in prows}


def codon_arrays(codon_rows, product, ncodon):
    out = {k: np.zeros(ncodon) for k in STATS}
    poly = np.zeros(ncodon, dtype=np.uint8)
    comps = [""] * ncodon
    for c in range(ncodon):
        r = codon_rows[(product, c + 1)]
        for k in STATS:
            out[k][c] = float(r[k])
        poly[c] = r["variability"] == "polymorphic"
        comps[c] = r["comparisons"]
    out["poly"] = poly
    out["comparisons_str"] = comps
    return out


def stats4_of(arrays):
    return np.stack([arrays[k] for k in STATS], axis=1)


def pair_set(comparisons: str):
    """'(AAA-AAG)(AAG-AAA)' -> set of unordered codon pairs."""
    return {frozenset((a, b)) for a, b in re.findall(r"\(([^()-]*)-([^()-]*)\)", comparisons)}


def between_label_maps(stdout_log):
    """Undo quirk q2: per product, map the TSV's group label -> the true group."""
    blocks, order = {}, []
    cur = None
    for line in open(stdout_log):
        m = re.match(r"#+ Analyzing (\S+)", line)
        if m:
            cur = m.group(1)
            blocks[cur] = []
            order.append(cur)
        m = re.match(r"group index (\d+) is group (\S+)", line)
        if m and cur is not None:
            blocks[cur].append(m.group(2))
    first = blocks[order[0]]
    return {p: dict(zip(first, blocks[p])) for p in order}

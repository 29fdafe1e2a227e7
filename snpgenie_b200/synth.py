"""Seeded synthetic alignments of the shapes BASELINE.json names.

Ancestral sequence: uniform ACGT outside CDS; inside each '+' CDS, codons drawn
from the 61 sense codons with ATG first and a STOP last.  Each codon column is
polymorphic with probability p_poly: 2-4 alleles differing from the ancestor at
1-3 positions, minor-allele frequencies ~ Beta(0.5, 5).  Optional damage:
codon-aligned '-' runs, 'N', IUPAC R/Y, lower case and RNA 'U'.
"""
from __future__ import annotations

import numpy as np

from .hostio import Product

_NT = np.frombuffer(b"ACGT", dtype=np.uint8)
_STOPS = {"TAA", "TAG", "TGA"}
_SENSE = np.array(
    [[ord(a), ord(b), ord(c)] for a in "ACGT" for b in "ACGT" for c in "ACGT" if a + b + c not in _STOPS],
    dtype=np.uint8,
)

SARS2_CDS = [
    (266, 13483), (13468, 21555), (21563, 25384), (25393, 26220), (26245, 26472),
    (26523, 27191), (27202, 27387), (27394, 27759), (27894, 28259), (28274, 29533),
]


def sars2_products() -> list[Product]:
    """Config 2/4/5 annotation: 10 '+' CDS, 9,677 codons on a 29,903-nt genome."""
    return [Product(f"CDS{i + 1}", 1, [s], [e]) for i, (s, e) in enumerate(SARS2_CDS)]


def hiv_products() -> list[Product]:
    """Config 3 annotation: 9,719 nt, overlapping frames, a 2-segment product, two '-' CDS."""
    return [
        Product("gag", 1, [790], [2292]),
        Product("pol", 1, [2085], [5096]),           # overlaps gag in another frame
        Product("vif", 1, [5041], [5619]),
        Product("vpr", 1, [5559], [5849]),
        Product("tat", 1, [5831, 8379], [6046, 8468]),  # two segments: 216 + 90
        Product("vpu", 1, [6062], [6307]),
        Product("env", 1, [6225], [8795]),
        Product("asp", -1, [7373], [7942]),           # antisense, inside env
        Product("nefr", -1, [8797], [9417]),
    ]


def make_alignment(n: int, length: int, products: list[Product], seed: int, *, p_poly: float = 0.3, max_changes: int = 3, max_alleles: int = 3,
                   gap_frac: float = 0.0, n_frac: float = 0.0, iupac_frac: float = 0.0,
                   lower_frac: float = 0.0, u_frac: float = 0.0) -> np.ndarray:
    """uint8 matrix [n, length] of ASCII residues."""
    rng = np.random.default_rng(seed)
    anc = _NT[rng.integers(0, 4, size=length)]
    plus = [p for p in products if p.strand > 0]
    for p in plus:
        pos = np.concatenate([np.arange(s - 1, e) for s, e in zip(p.starts, p.stops)])
        nc = len(pos) // 3
        cod = _SENSE[rng.integers(0, len(_SENSE), size=nc)]
        cod[0] = (65, 84, 71)      # ATG
        cod[-1] = (84, 65, 65)     # TAA
        anc[pos] = cod.reshape(-1)
    aln = np.repeat(anc[None, :], n, axis=0)
    # polymorphic codon columns, taken in every product's own frame
    for p in products:
        pos = np.concatenate([np.arange(s - 1, e) for s, e in zip(p.starts, p.stops)])
        nc = len(pos) // 3
        for c in np.flatnonzero(rng.random(nc) < p_poly):
            cols = pos[3 * c:3 * c + 3]
            k = min(int(rng.integers(1, 4)), max_alleles)  # minor alleles
            freqs = rng.beta(0.5, 5.0, size=k)
            assign = rng.random(n)
            lo = 0.0
            for f in freqs:
                rows = np.flatnonzero((assign >= lo) & (assign < lo + f))
                lo += f
                if rows.size == 0:
                    continue
                npos = int(rng.integers(1, 4))
                if npos > max_changes:      # max_changes = 1: every allele is a single-nucleotide change (side experiments)
                    npos = max_changes
                which = rng.choice(3, size=npos, replace=False)
                for w in which:
                    old = aln[rows[0], cols[w]]
                    choices = _NT[_NT != old]
                    aln[rows, cols[w]] = choices[rng.integers(0, len(choices))]
    total = n * length
    if gap_frac > 0:   # codon-aligned runs of '-' (1..7 codons) in a random product's frame
        n_runs = int(gap_frac * total / 12)
        pos_of = [np.concatenate([np.arange(s - 1, e) for s, e in zip(p.starts, p.stops)]) for p in products]
        which_p = rng.integers(0, len(products), size=n_runs)
        rows = rng.integers(0, n, size=n_runs)
        lens = rng.integers(1, 8, size=n_runs)
        for r, ip, l in zip(rows, which_p, lens):
            pos = pos_of[ip]
            c = int(rng.integers(0, len(pos) // 3))
            aln[r, pos[3 * c:3 * (c + l)]] = ord("-")
    def _sprinkle(frac, symbols):
        m = int(frac * total)
        if m:
            idx = rng.integers(0, total, size=m)
            aln.reshape(-1)[idx] = np.frombuffer(symbols, dtype=np.uint8)[rng.integers(0, len(symbols), size=m)]
    _sprinkle(n_frac, b"N")
    _sprinkle(iupac_frac, b"RY")
    if u_frac > 0:
        flat = aln.reshape(-1)
        idx = rng.integers(0, total, size=int(u_frac * total))
        idx = idx[flat[idx] == ord("T")]
        flat[idx] = ord("U")
    if lower_frac > 0:
        flat = aln.reshape(-1)
        idx = rng.integers(0, total, size=int(lower_frac * total))
        idx = idx[(flat[idx] >= 65) & (flat[idx] <= 90)]
        flat[idx] += 32
    return aln


def write_fasta(path: str, aln: np.ndarray, prefix: str = "seq", width: int = 0) -> None:
    with open(path, "wb") as fh:
        for i, row in enumerate(aln):
            fh.write(b">%s_%d\n" % (prefix.encode(), i + 1))
            data = row.tobytes()
            if width:
                for k in range(0, len(data), width):
                    fh.write(data[k:k + width] + b"\n")
            else:
                fh.write(data + b"\n")


def write_gtf(path: str, products: list[Product]) -> None:
    with open(path, "w") as fh:
        for p in products:
            for s, e in zip(p.starts, p.stops):
                fh.write(f"synthetic\tb200\tCDS\t{s}\t{e}\t.\t{'+' if p.strand > 0 else '-'}\t0\tgene_id \"{p.name}\";\n")

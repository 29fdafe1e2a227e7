"""Host-side ingest: aligned FASTA and CDS GTF, with the reference's semantics.

These are the L4 "input model" steps that stay on the host (SURVEY.md section 8
rows a1/a2).  They only produce the plain buffers the C-ABI takes: a row-major
byte matrix for the alignment and a CSR list of CDS segments per product.

Reference behaviour mirrored here:
  * FASTA (snpgenie_within_group.pl:204-253): any line containing '>' starts a
    record; other lines are chomp'ed (LF only) and concatenated; all records
    must have equal length.  Upper-casing and U->T are NOT done here -- the
    device packing kernel owns them so codon indices stay bit-exact in one place.
  * GTF (snpgenie_within_group.pl:1368-1392, :1530-1621): a line is a CDS record
    iff it matches /CDS\\t\\d+\\t\\d+\\t[\\.\\d+]\\t\\+/; product name from gene_id
    (three regex forms, tried in the reference's order); segments keep file
    order and are then sorted by start; total length must be a multiple of 3.
    The reference ignores '-' strand records; with include_minus=True they are
    kept as strand -1 products (natively equivalent to the reference's
    fasta2revcom.pl + gtf2revcom.pl route).
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field

import numpy as np

_CDS_PLUS = re.compile(r"CDS\t\d+\t\d+\t[\.\d+]\t\+")
_CDS_MINUS = re.compile(r"CDS\t\d+\t\d+\t[\.\d+]\t\-")
_GENE_ID_FORMS = (
    re.compile(r"gene_id \"gene\:([\w\s\.\-\:']+)\""),
    re.compile(r"gene_id \"([\w\s\.\-\:']+ [\w\s\.\-\:']+)\""),
    re.compile(r"gene_id \"([\w\s\.\-\:']+)\""),
)
_COORDS = re.compile(r"CDS\t(\d+)\t(\d+)")


class InputError(ValueError):
    """Raised where the reference would die() on malformed input."""


@dataclass
class Product:
    name: str
    strand: int  # +1 / -1
    starts: list = field(default_factory=list)  # 1-based inclusive, sorted by start
    stops: list = field(default_factory=list)

    @property
    def n_codons(self) -> int:
        return sum(e - s + 1 for s, e in zip(self.starts, self.stops)) // 3


def read_fasta(path: str) -> tuple[list[str], np.ndarray]:
    """Return (headers, uint8 matrix [n_seqs, aln_len]) for an aligned FASTA."""
    headers: list[str] = []
    chunks: list[list[bytes]] = []
    with open(path, "rb") as fh:
        for line in fh:
            if line.endswith(b"\n"):
                line = line[:-1]  # chomp: LF only, a CR stays in the sequence
            if b">" in line:
                headers.append(line.decode("latin-1"))
                chunks.append([])
            elif chunks:
                chunks[-1].append(line)
    seqs = [b"".join(c) for c in chunks]
    if not seqs:
        raise InputError(f"no FASTA records in {path}")
    length = len(seqs[0])
    for s in seqs:
        if len(s) != length:
            raise InputError("The sequences must be aligned, i.e., must be the same length.")
    aln = np.frombuffer(b"".join(seqs), dtype=np.uint8).reshape(len(seqs), length)
    return headers, aln


def _gene_id(line: str) -> str:
    for rx in _GENE_ID_FORMS:
        m = rx.search(line)
        if m:
            return m.group(1)
    raise InputError("CDS annotation(s) in the GTF file do not have a gene_id.")


def read_gtf_products(path: str, include_minus: bool = False) -> list[Product]:
    """Products in first-appearance order; '+' records only unless include_minus."""
    products: dict[tuple[str, int], Product] = {}
    with open(path, "r", newline="") as fh:
        for line in fh:
            if _CDS_PLUS.search(line):
                strand = 1
            elif include_minus and _CDS_MINUS.search(line):
                strand = -1
            else:
                continue
            name = _gene_id(line)
            m = _COORDS.search(line)
            p = products.setdefault((name, strand), Product(name, strand))
            p.starts.append(int(m.group(1)))
            p.stops.append(int(m.group(2)))
    out = []
    for p in products.values():
        order = sorted(range(len(p.starts)), key=lambda i: (p.starts[i], i))
        p.starts = [p.starts[i] for i in order]
        p.stops = [p.stops[i] for i in order]
        total = sum(e - s + 1 for s, e in zip(p.starts, p.stops))
        if total % 3 != 0:
            raise InputError(
                f"The CDS coordinates for gene {p.name} in the gtf file do not yield a set of complete codons"
            )
        out.append(p)
    return out


def products_to_csr(products: list[Product]):
    """CSR arrays (seg_ofs, seg_start, seg_stop, strand) for snpg_set_products."""
    seg_ofs = np.zeros(len(products) + 1, dtype=np.int64)
    starts, stops = [], []
    for i, p in enumerate(products):
        starts += p.starts
        stops += p.stops
        seg_ofs[i + 1] = len(starts)
    return (
        seg_ofs,
        np.asarray(starts, dtype=np.int64),
        np.asarray(stops, dtype=np.int64),
        np.asarray([p.strand for p in products], dtype=np.int8),
    )

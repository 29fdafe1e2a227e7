#!/usr/bin/env python
"""Generate tests/golden/ by running the UNMODIFIED reference scripts.

Run in the authoring container only (needs /root/reference and perl); the GPU
box never runs this -- it reads the committed fixtures.  Every case is a seeded
synthetic alignment (snpgenie_b200.synth) so the inputs can be regenerated, but
inputs are committed too so tests do not depend on numpy's RNG stream.

    python tools/make_goldens.py [case ...]

Reference quirks handled here (SURVEY.md section 8 "quirks"):
  * outputs are opened in append mode -> every run gets a fresh directory;
  * Parallel::ForkManager is not installed -> tools/perl_standin on -I;
  * between-group labels are permuted after the first product (q2) -> the
    "group index N is group X" log lines are saved so tests can undo it.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from snpgenie_b200 import synth  # noqa: E402
from snpgenie_b200.hostio import Product  # noqa: E402

REF = "/root/reference"
STANDIN = os.path.join(ROOT, "tools", "perl_standin")
GOLD = os.path.join(ROOT, "tests", "golden")
SEED = 20261019


def perl(script, args, cwd, log="stdout.log"):
    cmd = ["perl", "-I", STANDIN, os.path.join(REF, script)] + args
    with open(os.path.join(cwd, log), "w") as fh:
        subprocess.run(cmd, cwd=cwd, stdout=fh, stderr=subprocess.STDOUT, check=True)


def small_products():
    """600-nt toy genome: overlap in another frame, a 2-segment product."""
    return [
        Product("geneA", 1, [10], [189]),             # 60 codons
        Product("geneB", 1, [150, 300], [260, 389]),  # 37 + 30 codons, overlaps geneA (frame +2)
        Product("geneC", 1, [400], [579]),            # 60 codons
    ]


def keep(tmp, case, names):
    out = os.path.join(GOLD, case)
    os.makedirs(out, exist_ok=True)
    for nm in names:
        shutil.copy(os.path.join(tmp, nm), os.path.join(out, nm))


def example_alignment(n, seed):
    """Config 1: n sequences mutated from EXAMPLE_files/REFERENCE_EXAMPLE.fasta."""
    ref = "".join(l.strip() for l in open(os.path.join(REF, "EXAMPLE_files", "REFERENCE_EXAMPLE.fasta")) if not l.startswith(">"))
    anc = np.frombuffer(ref.encode(), dtype=np.uint8)
    rng = np.random.default_rng(seed)
    aln = np.repeat(anc[None, :], n, axis=0).copy()
    nt = np.frombuffer(b"ACGT", dtype=np.uint8)
    # ~4% of sites segregate, minor-allele frequency Beta(0.5,5)
    for site in np.flatnonzero(rng.random(len(anc)) < 0.04):
        rows = np.flatnonzero(rng.random(n) < max(rng.beta(0.5, 5.0), 1.0 / n))
        aln[rows, site] = rng.choice(nt[nt != anc[site]])
    return aln


def case_w_cfg1():
    with tempfile.TemporaryDirectory() as tmp:
        aln = example_alignment(50, SEED + 1)
        synth.write_fasta(os.path.join(tmp, "aln.fasta"), aln, width=80)
        shutil.copy(os.path.join(REF, "EXAMPLE_files", "CDS_EXAMPLE.gtf"), os.path.join(tmp, "cds.gtf"))
        perl("snpgenie_within_group.pl", ["--fasta_file_name=aln.fasta", "--gtf_file_name=cds.gtf", "--num_bootstraps=100", "--procs_per_node=8"], tmp)
        keep(tmp, "w_cfg1", ["aln.fasta", "cds.gtf", "within_group_codon_results.txt", "within_group_product_results.txt",
                             "within_group_variant_results.txt", "within_group_site_results.txt"])


def _within_small(case, seed, **damage):
    with tempfile.TemporaryDirectory() as tmp:
        prods = small_products()
        aln = synth.make_alignment(24, 600, prods, seed, p_poly=0.35, **damage)
        if case == "w_heavy":
            aln[:, 30:36] = ord("-")        # two all-undefined codon columns in geneA
            aln[1:, 45:48] = ord("N")       # a single-defined column
            aln[:, 420] = ord("n")          # lower-case n -> undefined after uc
        synth.write_fasta(os.path.join(tmp, "aln.fasta"), aln)
        synth.write_gtf(os.path.join(tmp, "cds.gtf"), prods)
        perl("snpgenie_within_group.pl", ["--fasta_file_name=aln.fasta", "--gtf_file_name=cds.gtf", "--num_bootstraps=200", "--procs_per_node=8"], tmp)
        names = ["aln.fasta", "cds.gtf", "within_group_codon_results.txt", "within_group_product_results.txt",
                 "within_group_variant_results.txt", "within_group_site_results.txt"]
        if case == "w_gaps":
            perl("snpgenie_within_group_processor.pl", ["--codon_file=within_group_codon_results.txt", "--sliding_window_size=10",
                                                        "--sliding_window_step=3", "--num_bootstraps=200"], tmp, log="processor.log")
            names.append("sw_10codons_results.txt")
        keep(tmp, case, names)


def case_w_gaps():
    _within_small("w_gaps", SEED + 2, gap_frac=0.03, n_frac=0.02, iupac_frac=0.005, lower_frac=0.05, u_frac=0.02)


def case_w_heavy():
    _within_small("w_heavy", SEED + 3, gap_frac=0.15, n_frac=0.10, iupac_frac=0.01)


def case_w_minus():
    """'-' strand oracle route: fasta2revcom.pl + gtf2revcom.pl, then the within driver."""
    with tempfile.TemporaryDirectory() as tmp:
        prods = [Product("geneA", 1, [10], [189]), Product("revB", -1, [150, 300], [260, 389]), Product("revC", -1, [400], [579])]
        aln = synth.make_alignment(16, 600, prods, SEED + 4, p_poly=0.35, gap_frac=0.03, n_frac=0.01, iupac_frac=0.005, lower_frac=0.05, u_frac=0.02)
        synth.write_fasta(os.path.join(tmp, "aln.fasta"), aln)
        synth.write_gtf(os.path.join(tmp, "cds.gtf"), prods)
        perl("fasta2revcom.pl", ["aln.fasta"], tmp, log="revcom1.log")
        perl("gtf2revcom.pl", ["cds.gtf", "600"], tmp, log="revcom2.log")
        perl("snpgenie_within_group.pl", ["--fasta_file_name=aln_revcom.fasta", "--gtf_file_name=cds_revcom.gtf", "--num_bootstraps=2", "--procs_per_node=8"], tmp)
        keep(tmp, "w_minus", ["aln.fasta", "cds.gtf", "cds_revcom.gtf", "within_group_codon_results.txt", "within_group_product_results.txt"])


def case_b_3groups():
    with tempfile.TemporaryDirectory() as tmp:
        prods = small_products()
        aln = synth.make_alignment(19, 600, prods, SEED + 5, p_poly=0.4, gap_frac=0.04, n_frac=0.02, iupac_frac=0.004)
        # make the groups differ: group-specific substitutions
        rng = np.random.default_rng(SEED + 6)
        groups = {"grpA.fasta": aln[:8].copy(), "grpB.fasta": aln[8:14].copy(), "grpC.fasta": aln[14:].copy()}
        nt = np.frombuffer(b"ACGT", dtype=np.uint8)
        for g in groups.values():
            for site in rng.integers(9, 580, size=25):
                g[:, site] = nt[rng.integers(0, 4)]
        groups["grpB.fasta"][:, 60:63] = ord("-")   # one group entirely undefined at a codon of geneA
        groups["grpC.fasta"][:, 405:408] = ord("N")
        groups["grpA.fasta"][:, 405:408] = ord("N")  # both groups of a pair undefined -> '---'
        for nm, g in groups.items():
            synth.write_fasta(os.path.join(tmp, nm), g, prefix=nm[:4])
        synth.write_gtf(os.path.join(tmp, "cds.gtf"), prods)
        perl("snpgenie_between_group.pl", ["--gtf_file=cds.gtf", "--procs_per_node=8"], tmp)
        perl("snpgenie_between_group_processor.pl", ["--between_group_codon_file=between_group_codon_results.txt", "--sliding_window_size=8",
                                                     "--sliding_window_step=2", "--num_bootstraps=200", "--codon_min_taxa_total=9",
                                                     "--codon_min_taxa_group=4"], tmp, log="processor.log")
        keep(tmp, "b_3groups", list(groups) + ["cds.gtf", "stdout.log", "between_group_codon_results.txt", "between_group_product_results.txt",
                                               "between_group_sw_8codons_results.txt", "between_group_product_results_filtered.txt"])


CASES = {f[5:]: v for f, v in list(globals().items()) if f.startswith("case_")}

if __name__ == "__main__":
    todo = sys.argv[1:] or list(CASES)
    for c in todo:
        print("generating", c, flush=True)
        CASES[c]()
    print("done")

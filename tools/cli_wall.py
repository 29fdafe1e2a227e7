#!/usr/bin/env python
"""Wall time of the Perl command-line host on a config-2-sized FASTA (10,000 x 29,903, 10 CDS, B = 10,000).

    python tools/cli_wall.py [--n 10000] [--boot 10000] [--keep-dir DIR]
Writes the FASTA/GTF to a scratch directory, runs perl/snpgenie_within_group_b200.pl there and prints the
wall time of the whole program (FASTA read, H2D, kernels, table writing).
"""
import argparse
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from snpgenie_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=10000)
ap.add_argument("--boot", type=int, default=10000)
ap.add_argument("--width", type=int, default=70, help="FASTA line width (0 = one line per sequence)")
ap.add_argument("--gpus", type=int, default=1)
a = ap.parse_args()
products = synth.sars2_products()
with tempfile.TemporaryDirectory() as d:
    t0 = time.perf_counter()
    aln = synth.make_alignment(a.n, 29903, products, 20261021)
    synth.write_fasta(os.path.join(d, "aln.fasta"), aln, width=a.width)
    synth.write_gtf(os.path.join(d, "cds.gtf"), products)
    size = os.path.getsize(os.path.join(d, "aln.fasta"))
    print(f"generated {size / 1e6:.0f} MB FASTA in {time.perf_counter() - t0:.1f} s", flush=True)
    for rep in range(2):
        for f in os.listdir(d):
            if f.endswith(".txt"):
                os.remove(os.path.join(d, f))
        t0 = time.perf_counter()
        p = subprocess.run(["perl", os.path.join(ROOT, "perl", "snpgenie_within_group_b200.pl"), "--fasta_file_name=aln.fasta",
                            "--gtf_file_name=cds.gtf", f"--num_bootstraps={a.boot}", "--procs_per_node=16", "--no_bootstrap_files",
                            f"--gpus={a.gpus}"],
                           cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=dict(os.environ, SNPG_TRACE="1"))
        wall = time.perf_counter() - t0
        if p.returncode:
            print(p.stdout[-2000:])
            sys.exit(1)
        rows = sum(1 for _ in open(os.path.join(d, "within_group_codon_results.txt"))) - 1
        print(f"run {rep}: within-group CLI wall {wall:.2f} s ({rows} codon rows, {a.n} seqs, B={a.boot}, {a.gpus} GPU(s))", flush=True)
        for line in p.stdout.splitlines():           # where the time goes: the library's traced entry points, the script's phases
            if line.startswith(("[snpg]", "[perl]")):
                print("    " + line)
    print(open(os.path.join(d, "within_group_product_results.txt")).read()[:600])

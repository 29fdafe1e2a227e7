"""The Perl command-line host (perl/*.pl + XS glue) against the reference's own outputs.

CPU part: the glue builds, loads, exposes the ABI and the scripts validate flags like the reference.
GPU part: the four programs run end to end on the golden inputs; every numeric column is compared with
the tables the unmodified reference wrote (1e-12; bootstrap columns statistically).
"""
import os
import shutil
import subprocess

import numpy as np
import pytest

from tests import refio
from tests.refio import STATS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PERL = os.path.join(ROOT, "perl")


def close(a, b, rtol=1e-12, atol=1e-15):
    return np.allclose(a, b, rtol=rtol, atol=atol)


def run(script, args, cwd, check=True):
    p = subprocess.run(["perl", os.path.join(PERL, script)] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if check and p.returncode != 0:
        raise AssertionError(f"{script} failed:\n{p.stdout}")
    return p


@pytest.fixture(scope="module", autouse=True)
def built_glue():
    if not os.path.exists(os.path.join(PERL, "SNPGenieCUDA", "lib", "auto", "SNPGenieCUDA", "SNPGenieCUDA.so")):
        subprocess.check_call(["make", "-C", os.path.join(PERL, "SNPGenieCUDA")], stdout=subprocess.DEVNULL)


def test_glue_loads_and_tables_match_header_contract():
    code = ('use SNPGenieCUDA; my ($s,$d)=SNPGenieCUDA::tables(); my @s=unpack("d*",$s); my @d=unpack("d*",$d);'
            'printf("%d %d %.17g %.17g\\n", scalar(@s), scalar(@d), $s[0], $d[0*4096+0*64+5]); print SNPGenieCUDA::version(),"\\n";')
    out = subprocess.check_output(["perl", "-I", os.path.join(PERL, "SNPGenieCUDA", "lib"), "-e", code], text=True).split("\n")
    n_s, n_d, aaa_n, aaa_acc = out[0].split()
    assert (int(n_s), int(n_d)) == (128, 8192)
    assert float(aaa_n) == 2 + 2 / 3 and float(aaa_acc) == 1.5          # AAA sites; AAA->ACC N_diffs
    assert "sm_100a" in out[1]


@pytest.mark.parametrize("script", ["snpgenie_within_group_b200.pl", "snpgenie_between_group_b200.pl",
                                    "snpgenie_within_group_processor_b200.pl", "snpgenie_between_group_processor_b200.pl",
                                    "SNPGenie_sliding_windows_b200.pl"])
def test_scripts_compile(script):
    p = subprocess.run(["perl", "-c", os.path.join(PERL, script)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 0, p.stdout


def test_flag_validation_matches_reference_messages(tmp_path):
    p = run("snpgenie_within_group_b200.pl", ["--fasta_file_name=x.txt", "--gtf_file_name=a.gtf"], tmp_path, check=False)
    assert p.returncode != 0 and "--fasta_file_name option must be a file with a .fa or .fasta extension" in p.stdout
    p = run("snpgenie_within_group_b200.pl", ["--fasta_file_name=x.fa", "--gtf_file_name=a.txt"], tmp_path, check=False)
    assert p.returncode != 0 and "--gtf_file_name option" in p.stdout
    p = run("snpgenie_between_group_processor_b200.pl", ["--between_group_codon_file=nope.txt", "--num_bootstraps=50"], tmp_path, check=False)
    assert p.returncode != 0 and "integer >=100" in p.stdout
    (tmp_path / "bad.gtf").write_text('x\ty\tCDS\t1\t10\t.\t+\t0\tgene_id "g";\n')
    (tmp_path / "a.fa").write_text(">s\nACGTACGTACGT\n")
    p = run("snpgenie_within_group_b200.pl", ["--fasta_file_name=a.fa", "--gtf_file_name=bad.gtf"], tmp_path, check=False)
    assert p.returncode != 0 and "do not yield a set of complete codons" in p.stdout


# ------------------------------------------------------------------------------- GPU

def _compare_within(tmp, case, products_filter=None):
    d, aln, gold_codon, gold_prod = refio.within_case(case)
    _, rows = refio.read_tsv(os.path.join(tmp, "within_group_codon_results.txt"))
    mine = {(r["product"], int(r["codon"])): r for r in rows}
    keys = [k for k in gold_codon if products_filter is None or k[0] in products_filter]
    assert set(mine) == set(keys)
    for k in keys:
        g, m = gold_codon[k], mine[k]
        assert g["variability"] == m["variability"], k
        for s in STATS:
            assert close(float(m[s]), float(g[s])), (k, s)
        if g["variability"] == "conserved":
            assert g["comparisons"] == m["comparisons"], k
        else:
            assert refio.pair_set(g["comparisons"]) == refio.pair_set(m["comparisons"]), k
        assert m["codon"] != m["codon"].strip()                       # trailing space kept (quirk q7)
    _, prow = refio.read_tsv(os.path.join(tmp, "within_group_product_results.txt"))
    for r in prow:
        g = gold_prod[r["product"]]
        for col in ("N_sites", "S_sites", "N_diffs", "S_diffs", "dN", "dS", "dN_minus_dS", "dN_over_dS"):
            assert close(float(r[col]), float(g[col])), (r["product"], col)
        if "SE_dN" in g and "SE_dN" in r:
            for col in ("SE_dN", "SE_dS", "SE_dN_minus_dS"):
                assert abs(float(r[col]) / float(g[col]) - 1) < 0.25, (r["product"], col)
            assert close(float(r["Z_value"]), float(r["dN_minus_dS"]) / float(r["SE_dN_minus_dS"]), rtol=1e-9)
    return prow


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["w_cfg1", "w_gaps", "w_heavy"])
def test_within_cli_matches_reference_tables(tmp_path, case):
    d = os.path.join(refio.GOLD, case)
    for f in ("aln.fasta", "cds.gtf"):
        shutil.copy(os.path.join(d, f), tmp_path / f)
    run("snpgenie_within_group_b200.pl", ["--fasta_file_name=aln.fasta", "--gtf_file_name=cds.gtf", "--num_bootstraps=4000",
                                          "--procs_per_node=8", "--seed=5"], tmp_path)
    prow = _compare_within(tmp_path, case)
    hdr, _ = refio.read_tsv(os.path.join(tmp_path, "within_group_product_results.txt"))
    ghdr, _ = refio.read_tsv(os.path.join(d, "within_group_product_results.txt"))
    assert hdr == ghdr
    for r in prow:
        _, b = refio.read_tsv(os.path.join(tmp_path, f"{r['product']}_bootstrap_results.txt"))
        assert len(b) == 4000 and list(b[0])[:3] == ["file", "product_name", "bootstrap_num"]
    # "next" rows f1: the site table is identical, the variant table too wherever the reference's counts are sane
    gh, gs = refio.read_tsv(os.path.join(d, "within_group_site_results.txt"))
    mh, ms = refio.read_tsv(os.path.join(tmp_path, "within_group_site_results.txt"))
    assert gh == mh and len(gs) == len(ms)
    for g, m in zip(gs, ms):
        for col in ("site", "maj_nt", "maj_nt_count", "coverage", "A", "C", "G", "T"):
            assert g[col] == m[col], (g["site"], col)
        assert g["pi"] == m["pi"] or close(float(m["pi"]), float(g["pi"]))
    gh, gv = refio.read_tsv(os.path.join(d, "within_group_variant_results.txt"))
    mh, mv = refio.read_tsv(os.path.join(tmp_path, "within_group_variant_results.txt"))
    assert gh == mh and len(gv) == len(mv)
    mvd = {(r["product_name"], r["codon_num"]): r for r in mv}
    for g in gv:
        m = mvd[(g["product_name"], g["codon_num"])]
        cols = ["num_alleles", "codons", "amino_acids", "variant_type"]
        if case == "w_cfg1":          # no gaps: the reference's counts are exact (quirk q1 does not bite)
            cols += ["consensus_codon", "codon_counts", "count_total", "codon_freqs", "max_codon_count", "min_codon_count", "num_defined_seqs"]
        for col in cols:
            assert g[col] == m[col], (g["product_name"], g["codon_num"], col)


@pytest.mark.gpu
def test_within_cli_minus_strand_native(tmp_path):
    d = os.path.join(refio.GOLD, "w_minus")
    for f in ("aln.fasta", "cds.gtf"):
        shutil.copy(os.path.join(d, f), tmp_path / f)
    run("snpgenie_within_group_b200.pl", ["--fasta_file_name=aln.fasta", "--gtf_file_name=cds.gtf", "--minus_strand"], tmp_path)
    _, rows = refio.read_tsv(os.path.join(tmp_path, "within_group_codon_results.txt"))
    _, _, gold_codon, _ = refio.within_case("w_minus")
    mine = {(r["product"], int(r["codon"])): r for r in rows if r["product"].startswith("rev")}
    assert set(mine) == set(gold_codon)
    for k, g in gold_codon.items():
        assert g["variability"] == mine[k]["variability"]
        for s in STATS:
            assert close(float(mine[k][s]), float(g[s])), (k, s)
        if g["variability"] == "conserved":
            assert g["comparisons"] == mine[k]["comparisons"]


@pytest.mark.gpu
def test_between_cli_and_processor_match_reference_tables(tmp_path):
    d = os.path.join(refio.GOLD, "b_3groups")
    for f in ("grpA.fasta", "grpB.fasta", "grpC.fasta", "cds.gtf"):
        shutil.copy(os.path.join(d, f), tmp_path / f)
    run("snpgenie_between_group_b200.pl", ["--gtf_file=cds.gtf", "--procs_per_node=8"], tmp_path)
    ghdr, grow = refio.read_tsv(os.path.join(d, "between_group_codon_results.txt"))
    hdr, rows = refio.read_tsv(os.path.join(tmp_path, "between_group_codon_results.txt"))
    assert hdr == ghdr and len(rows) == len(grow)
    maps = refio.between_label_maps(os.path.join(d, "stdout.log"))
    mine = {(r["product"], frozenset((r["group_1"], r["group_2"])), r["codon"]): r for r in rows}
    n_checked = 0
    for g in grow:
        true_pair = frozenset((maps[g["product"]][g["group_1"]], maps[g["product"]][g["group_2"]]))
        m = mine[(g["product"], true_pair, g["codon"])]
        swapped = m["group_1"] != maps[g["product"]][g["group_1"]]
        a, b = ("num_defined_codons_g2", "num_defined_codons_g1") if swapped else ("num_defined_codons_g1", "num_defined_codons_g2")
        assert int(g["num_defined_codons_g1"] or 0) == int(m[a]) and int(g["num_defined_codons_g2"] or 0) == int(m[b])
        assert g["variability"] == m["variability"]
        ambiguous = g["variability"] == "conserved" and min(int(m[a]), int(m[b])) == 0 and g["comparisons"] != m["comparisons"]
        if not ambiguous:
            for s in STATS:
                assert close(float(m[s]), float(g[s])), (g["product"], g["codon"], s)
            if g["variability"] == "conserved":
                assert g["comparisons"] == m["comparisons"]
            n_checked += 1
    assert n_checked >= len(grow) - 6
    # product table: same numbers for the same true group pair
    _, gp = refio.read_tsv(os.path.join(d, "between_group_product_results.txt"))
    _, mp = refio.read_tsv(os.path.join(tmp_path, "between_group_product_results.txt"))
    mpd = {(r["product"], frozenset((r["group_1"], r["group_2"]))): r for r in mp}
    for g in gp:
        m = mpd[(g["product"], frozenset((maps[g["product"]][g["group_1"]], maps[g["product"]][g["group_2"]])))]
        for col in ("dN", "dS", "dN/dS"):
            assert abs(float(m[col]) / float(g[col]) - 1) < 2e-2      # only the ambiguous first-codon rows can differ
    # the processor on the REFERENCE's codon file (labels as the reference wrote them)
    shutil.copy(os.path.join(d, "between_group_codon_results.txt"), tmp_path / "ref_codons.txt")
    run("snpgenie_between_group_processor_b200.pl", ["--between_group_codon_file=ref_codons.txt", "--sliding_window_size=8",
                                                     "--sliding_window_step=2", "--num_bootstraps=4000", "--codon_min_taxa_total=9",
                                                     "--codon_min_taxa_group=4", "--seed=3"], tmp_path)
    for name, keycols, exact, stat in (
        ("between_group_product_results_filtered.txt", ("product", "group_1", "group_2"),
         ("N_sites", "S_sites", "N_diffs", "S_diffs", "dN", "dS", "JC_dN", "JC_dS", "dN_minus_dS", "dNdS", "JC_dN_minus_dS", "JC_dNdS"),
         ("SE_dN", "SE_dS", "SE_JC_dN", "SE_JC_dS", "SE_dN_minus_dS", "SE_JC_dN_minus_dS")),
        ("between_group_sw_8codons_results.txt", ("product", "group_1", "group_2", "window"),
         ("first_codon", "last_codon", "num_defined_codons_g1", "num_defined_codons_g2", "mean_num_defined_codons",
          "min_num_defined_codons", "N_sites", "S_sites", "N_diffs", "S_diffs", "dN", "dS", "dN_minus_dS", "dN_over_dS"), ()),
    ):
        gh, gr = refio.read_tsv(os.path.join(d, name))
        mh, mr = refio.read_tsv(os.path.join(tmp_path, name))
        assert gh == mh and len(gr) == len(mr)
        md = {tuple(r[k] for k in keycols): r for r in mr}
        for g in gr:
            m = md[tuple(g[k] for k in keycols)]
            for col in exact:
                if g[col] in ("NA", "*", ""):
                    assert m[col] == g[col], (name, col)
                else:
                    assert close(float(m[col]), float(g[col])), (name, col, m[col], g[col])
            for col in stat:
                assert abs(float(m[col]) / float(g[col]) - 1) < 0.25, (name, col)


@pytest.mark.gpu
def test_within_processor_cli_matches_reference(tmp_path):
    d = os.path.join(refio.GOLD, "w_gaps")
    shutil.copy(os.path.join(d, "within_group_codon_results.txt"), tmp_path / "within_group_codon_results.txt")
    run("snpgenie_within_group_processor_b200.pl", ["--codon_file=within_group_codon_results.txt", "--sliding_window_size=10",
                                                    "--sliding_window_step=3", "--num_bootstraps=4000", "--seed=9"], tmp_path)
    gh, gr = refio.read_tsv(os.path.join(d, "sw_10codons_results.txt"))
    mh, mr = refio.read_tsv(os.path.join(tmp_path, "sw_10codons_results.txt"))
    assert gh == mh and len(gr) == len(mr)
    md = {(r["product"], r["window"]): r for r in mr}
    ratios = []
    for g in gr:
        m = md[(g["product"], g["window"])]
        for col in ("first_codon", "last_codon", "N_sites", "S_sites", "N_diffs", "S_diffs", "dN", "dS", "dN_minus_dS", "dN_over_dS"):
            if g[col] in ("*", ""):
                assert m[col] == g[col]
            else:
                assert close(float(m[col]), float(g[col])), (g["product"], g["window"], col)
        assert m["dN_gt_total_dS"] == g["dN_gt_total_dS"]
        if float(g["SE_dN_minus_dS"]) > 0:
            ratios.append(float(m["SE_dN_minus_dS"]) / float(g["SE_dN_minus_dS"]))
    assert 0.85 < np.median(ratios) < 1.2


def test_sliding_windows_cli_validates_arguments_like_the_r_script(tmp_path):
    p = subprocess.run(["perl", os.path.join(PERL, "SNPGenie_sliding_windows_b200.pl"), "x.tsv", "N", "S"], cwd=tmp_path, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 1 and "Command line Error 1" in p.stdout and "there must be 5-10 command line arguments" in p.stdout
    p = subprocess.run(["perl", os.path.join(PERL, "SNPGenie_sliding_windows_b200.pl"), "x.tsv", "N", "S", "1", "1"], cwd=tmp_path,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 1 and "Command line Error 5" in p.stdout


@pytest.mark.gpu
def test_sliding_windows_cli_on_a_between_group_table(tmp_path):
    """The drop-in for SNPGenie_sliding_windows.R on one product x one group pair of the reference's own between-group
    codon table: window numbering, sums and ratios against the oracle's restatement of the script; the table layout
    (input columns + num_defined_seqs + codon_num + sw_* columns, NA outside window starts) as write_tsv leaves it."""
    from oracle import oracle as orc
    d = os.path.join(refio.GOLD, "b_3groups")
    hdr, rows = refio.read_tsv(os.path.join(d, "between_group_codon_results.txt"))
    first = rows[0]
    sel = [r for r in rows if (r["product"], r["group_1"], r["group_2"]) == (first["product"], first["group_1"], first["group_2"])]
    with open(tmp_path / "one.txt", "w") as fh:
        fh.write("\t".join(hdr) + "\n")
        for r in sel:
            fh.write("\t".join(r[h] for h in hdr) + "\n")
    env = dict(os.environ, SNPG_SEED="5")
    p = subprocess.run(["perl", os.path.join(PERL, "SNPGenie_sliding_windows_b200.pl"), "one.txt", "N", "S", "8", "2", "3000", "4", "NONE", "1"],
                       cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    assert p.returncode == 0, p.stdout
    assert "BETWEEN-GROUP FILE DETECTED" in p.stdout and "DONE" in p.stdout
    oh, orows = refio.read_tsv(os.path.join(tmp_path, "one_WINDOWS_dNdS.tsv"))
    assert oh[:len(hdr)] == hdr and oh[len(hdr):len(hdr) + 3] == ["num_defined_seqs", "codon_num", "sw_ratio"] and oh[-1] == "sw_ASL_dNdS_P"
    assert len(orows) == len(sel)
    t = np.array([[float(r[k]) for k in ("N_sites", "S_sites", "N_diffs", "S_diffs")] for r in sel])
    nd = np.array([min(int(r["num_defined_codons_g1"]), int(r["num_defined_codons_g2"])) for r in sel])
    cn = np.array([int(r["codon"].replace("codon_", "")) for r in sel])
    want = orc.sliding_windows_r(t, 8, 2, B=100, n_defined=nd, min_defined=4, codon_num=cn)
    starts = {int(s): k for k, s in enumerate(want["sw_start"])}
    seen = 0
    for r in orows:
        k = starts.get(int(r["codon_num"]))
        if k is None:
            assert r["sw_start"] == "NA" and r["sw_boot_dN_SE"] == "NA"
            continue
        seen += 1
        assert r["sw_ratio"] == "dNdS" and int(r["sw_num_replicates"]) == 3000
        for col in ("sw_start", "sw_center", "sw_end", "sw_N_diffs", "sw_S_diffs", "sw_N_sites", "sw_S_sites", "sw_dN", "sw_dS", "sw_dNdS", "sw_dN_m_dS"):
            w = want[col][k]
            assert (r[col] == "NaN") if w != w else close(float(r[col]), w, rtol=1e-12), (col, r[col], w)
        if r["sw_boot_dN_gt_dS_count"] != "NaN":
            assert int(float(r["sw_boot_dN_gt_dS_count"])) + int(float(r["sw_boot_dN_eq_dS_count"])) + int(float(r["sw_boot_dN_lt_dS_count"])) == 3000
    assert seen == len(starts) > 0

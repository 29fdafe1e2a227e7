#!/usr/bin/env perl
# Within-group dN/dS for one aligned FASTA + CDS GTF, computed on a B200 through
# libsnpgenie_cuda.so.  Drop-in for the reference's snpgenie_within_group.pl: same flags,
# same per-codon / per-product / bootstrap tables (appended to, as the reference does).
#
#   snpgenie_within_group_b200.pl --fasta_file_name=<aligned>.fa --gtf_file_name=<cds>.gtf \
#       --num_bootstraps=10000 --procs_per_node=16
#
# --procs_per_node is accepted and ignored: the reference's fork-per-codon has no counterpart
# (CUDA contexts do not survive fork()).  Extra flags: --gpus=N (split the codon axis and the
# bootstrap replicates over N GPUs of this node; the ranks exchange over NVLink), --device,
# --seed, --minus_strand (analyse '-' strand CDS natively instead of ignoring them),
# --no_bootstrap_files.
use strict;
use warnings;
use FindBin;
use lib "$FindBin::Bin/lib", "$FindBin::Bin/SNPGenieCUDA/lib";
use Getopt::Long;
use SNPGenieCUDA;
use SNPGenieHost qw(read_gtf pack_products ratio stars codon_of other_from_strings fmt_pairs variant_line site_lines);

STDOUT->autoflush(1);
my $time1 = time;
use Time::HiRes ();
my $trace_t = Time::HiRes::time();
sub trace {      # SNPG_TRACE=1: where the wall time goes (tools/cli_wall.py)
    return unless $ENV{SNPG_TRACE};
    my $now = Time::HiRes::time();
    printf STDERR "[perl] %-24s %9.3f ms\n", $_[0], 1e3 * ($now - $trace_t);
    $trace_t = $now;
}
my ($fasta_file_name, $gtf_file_name, $procs_per_node, $num_bootstraps);
my ($device, $seed, $minus_strand, $no_boot_files, $gpus) = (0, time ^ ($$ << 15), 0, 0, 1);
GetOptions("fasta_file_name=s" => \$fasta_file_name, "gtf_file_name=s" => \$gtf_file_name,
           "procs_per_node=i" => \$procs_per_node, "num_bootstraps=i" => \$num_bootstraps,
           "device=i" => \$device, "seed=i" => \$seed, "minus_strand" => \$minus_strand,
           "no_bootstrap_files" => \$no_boot_files, "gpus=i" => \$gpus)
    or die "\n### WARNING: Error in command line arguments. Script terminated.\n\n";
unless (defined $fasta_file_name && $fasta_file_name =~ /.fa/) {
    die "\n### WARNING: The --fasta_file_name option must be a file with a .fa or .fasta extension\n### SNPGenie terminated.\n\n";
}
unless (defined $gtf_file_name && $gtf_file_name =~ /.gtf/) {
    die "\n### WARNING: The --gtf_file_name option must be a file with a .gtf or .gtf extension\n### SNPGenie terminated.\n\n";
}
if (defined $procs_per_node && $procs_per_node < 1) {
    die "\n### WARNING: The --procs_per_node option must be an integer >=1\n### SNPGenie terminated.\n\n";
}
if ($gpus < 1 || $gpus > 8) {
    die "\n### WARNING: The --gpus option must be an integer from 1 to 8\n### SNPGenie terminated.\n\n";
}
if (defined $num_bootstraps && $num_bootstraps != 0 && $num_bootstraps < 2) {
    die "\n### WARNING: The --num_bootstraps option must be an integer >=2\n### SNPGenie terminated.\n\n";
}
$num_bootstraps ||= 0;

print "\n################################################################################"
    . "\n##                 Within-Group SNPGenie (B200 engine) Initiated!             ##"
    . "\n################################################################################\n";
print "\n# " . SNPGenieCUDA::version() . "; --procs_per_node is ignored on this path.\n";

trace('startup');
my $products = read_gtf($gtf_file_name, $minus_strand);
print "\nRecording coding sequence data for $fasta_file_name...\n";
my $ctx = $gpus > 1 ? SNPGenieCUDA::create_multi($gpus, $seed) : SNPGenieCUDA::create($device, $seed);
# native reader: same record rules as wg:204-253, every host thread, straight into pinned memory; the rows are still being
# copied in the background while the job below starts to move the first row blocks to the GPU
SNPGenieCUDA::set_option($ctx, SNPGenieCUDA::OPT_INGEST_ASYNC, 1);
my ($bases_ptr, $n_seqs, $aln_len) = SNPGenieCUDA::read_fasta($ctx, 0, $fasta_file_name);
SNPGenieCUDA::set_option($ctx, SNPGenieCUDA::OPT_KEEP_CODES, 0);        # histograms straight from the alignment bytes
# per-site A/C/G/T in the same load on one GPU; several GPUs stage only the columns of their codon ranges, so the
# site table then comes from one more pass over the pinned alignment on the first GPU
SNPGenieCUDA::set_option($ctx, SNPGenieCUDA::OPT_SITE_COUNTS, 1) if $gpus == 1;
SNPGenieCUDA::set_products($ctx, scalar @$products, pack_products($products), $aln_len);
my $total_codons = 0;
$total_codons += $_->{n_codons} for @$products;
# the whole job in one call: load, per-codon engine, product sums and bootstrap SEs (the per-replicate files, when wanted,
# come from a bootstrap of their own per product below)
my $job_B = $no_boot_files ? $num_bootstraps : 0;
my ($job_sums_s, $job_se_s, $job_poly_s, $job_ndef_s, $job_cmp_s, $job_stats_s, $job_cons_s) =
    SNPGenieCUDA::within_job($ctx, 0, $bases_ptr, $n_seqs, $aln_len, $job_B, $total_codons, scalar @$products);
my @job_se = $job_B > 1 ? unpack('d*', $job_se_s) : ();
trace('create + read + job');
print "\n# " . ($gpus > 1 ? "$gpus GPUs" : "1 GPU") . ": $n_seqs sequences x $aln_len sites, $total_codons codons.\n";

# headers, append mode like the reference (wg:337-361)
open(my $codon_fh, '>>', 'within_group_codon_results.txt') or die "cannot write codon results: $!";
print $codon_fh "file\tproduct\tcodon\tvariability\tcomparisons\tN_sites\tS_sites\tN_diffs\tS_diffs\n";
open(my $var_fh, '>>', 'within_group_variant_results.txt') or die "cannot write variant results: $!";
print $var_fh "file\tproduct_name\tcodon_num\tconsensus_codon\tnum_alleles\tcodons\tamino_acids\tvariant_type\tcodon_counts\t"
            . "count_total\tcodon_freqs\tmax_codon_count\tmin_codon_count\tnum_defined_seqs\n";
open(my $prod_fh, '>>', 'within_group_product_results.txt') or die "cannot write product results: $!";
print $prod_fh "file\tproduct\tN_sites\tS_sites\tN_diffs\tS_diffs\tdN\tdS\t";
print $prod_fh "SE_dN\tSE_dS\t" if $num_bootstraps > 1;
print $prod_fh "dN_minus_dS\tdN_over_dS";
print $prod_fh($num_bootstraps > 1 ? "\tSE_dN_minus_dS\tZ_value\tsignificance\n" : "\n");

for my $pi (0 .. $#$products) {
    my $p = $products->[$pi];
    my $name = $p->{name};
    print "\n################################################################################\nANALYZING $name:\n";
    my $C = $p->{n_codons};
    my $c0 = 0;
    $c0 += $products->[$_]{n_codons} for 0 .. $pi - 1;                  # the product's place on the global codon axis
    my @poly = unpack('C*', substr($job_poly_s, $c0, $C));
    my @cons = unpack('C*', substr($job_cons_s, $c0, $C));
    my $stats_s = substr($job_stats_s, 32 * $c0, 32 * $C);
    my @stats = unpack('d*', $stats_s);
    my ($hist_s) = SNPGenieCUDA::hist($ctx, 0, $pi);
    my ($sum_Ns, $sum_Ss, $sum_Nd, $sum_Sd) = (0, 0, 0, 0);
    for my $c (0 .. $C - 1) {
        my ($Ns, $Ss, $Nd, $Sd) = @stats[4 * $c .. 4 * $c + 3];
        print $codon_fh "$fasta_file_name\t$name\t" . ($c + 1) . " \t";
        if ($poly[$c]) {
            my @h = unpack('L*', substr($hist_s, 4 * 66 * $c, 4 * 66));
            my @alleles = map { [ codon_of($_), $h[$_] ] } grep { $h[$_] > 0 } 0 .. 63;
            if ($h[65] > 0) {
                my $others = other_from_strings(SNPGenieCUDA::codon_strings($ctx, 0, $pi, $c, $n_seqs));
                push @alleles, map { [ $_, $others->{$_} ] } sort keys %$others;
            }
            print $codon_fh "polymorphic\t" . fmt_pairs(\@alleles) . "\t$Ns\t$Ss\t$Nd\t$Sd\n";
            print $var_fh variant_line($fasta_file_name, $name, $c + 1, \@alleles);
            # product sums accumulate codon by codon like wg:690-693
            $sum_Nd += $Nd; $sum_Sd += $Sd; $sum_Ns += $Ns; $sum_Ss += $Ss;
        } else {
            my $codon = '';
            if ($cons[$c] < 64) { $codon = codon_of($cons[$c]); }
            elsif ($cons[$c] == 65) { ($codon) = sort keys %{ other_from_strings(SNPGenieCUDA::codon_strings($ctx, 0, $pi, $c, $n_seqs)) }; }
            print $codon_fh "conserved\t$codon\t$Ns\t$Ss\t0\t0\n";
            $sum_Ns += $Ns; $sum_Ss += $Ss;
        }
    }

    my ($SE_dN, $SE_dS, $SE_diff, $Z) = ('', '', 0, 'NA');
    my $dN = ratio($sum_Nd, $sum_Ns);
    my $dS = ratio($sum_Sd, $sum_Ss);
    if ($num_bootstraps > 1) {
        print "\nBootstrapping for standard error...\n";
        if ($no_boot_files) {
            ($SE_dN, $SE_dS, $SE_diff) = @job_se[6 * $pi .. 6 * $pi + 2];      # from the whole-job call
        } else {
            my ($se_s, $sums_s) = SNPGenieCUDA::bootstrap_product($ctx, $stats_s, undef, $C, $num_bootstraps, 1);
            my @se = unpack('d*', $se_s);
            ($SE_dN, $SE_dS, $SE_diff) = @se[0 .. 2];
            open(my $bfh, '>>', "${name}_bootstrap_results.txt") or die "cannot write bootstrap file: $!";
            print $bfh "file\tproduct_name\tbootstrap_num\tsim_product_N_sites_sum\tsim_product_S_sites_sum\t"
                     . "sim_product_N_diffs_sum\tsim_product_S_diffs_sum\n";
            my @sums = unpack('d*', $sums_s);
            for my $b (0 .. $num_bootstraps - 1) {
                print $bfh join("\t", $fasta_file_name, $name, $b + 1, @sums[4 * $b .. 4 * $b + 3]) . "\n";
            }
            close $bfh;
        }
        no warnings 'numeric';       # '*' reads as 0 in the reference's arithmetic (wg:920-932)
        my $diff = ($dN >= 0 && $dS >= 0) ? $dN - $dS : '*';
        $Z = $diff / $SE_diff if $SE_diff > 0;
        $SE_diff = 'NA' if $SE_diff == 0;
    }
    my ($dN_minus_dS, $dN_over_dS);
    {
        no warnings 'numeric';
        $dN_minus_dS = $dN - $dS;                                   # wg:998
        $dN_over_dS = ($dS > 0) ? $dN / $dS : '*';                  # wg:1000-1004
    }
    my $line = "$fasta_file_name\t$name\t$sum_Ns\t$sum_Ss\t$sum_Nd\t$sum_Sd\t$dN\t$dS\t";
    $line .= "$SE_dN\t$SE_dS\t" if $num_bootstraps > 1;
    $line .= "$dN_minus_dS\t$dN_over_dS";
    {
        no warnings 'numeric';
        printf "\ndN=%.5f\ndS=%.5f\ndN/dS=%s\n", $dN, $dS, ($dN_over_dS eq '*' ? '*' : sprintf('%.5f', $dN_over_dS));
    }
    if ($num_bootstraps > 1) {
        no warnings 'numeric';
        printf "\nSE(dN-dS) = %.5f\nZ-value = %.5f\n\n", $SE_diff, $Z;
        $line .= "\t$SE_diff\t$Z\t" . stars($Z, $SE_diff);
    }
    print $prod_fh "$line\n";
}
close $codon_fh;
close $prod_fh;
close $var_fh;
trace('codon/product tables');
# per-site nucleotide counts and pi (wg:1203-1295)
open(my $site_fh, '>>', 'within_group_site_results.txt') or die "cannot write site results: $!";
print $site_fh "file\tsite\tmaj_nt\tmaj_nt_count\tpi\tcoverage\tA\tC\tG\tT\n";
if ($gpus > 1) {
    my $one = SNPGenieCUDA::create(0, $seed);
    SNPGenieCUDA::set_option($one, SNPGenieCUDA::OPT_SITE_COUNTS, 1);
    SNPGenieCUDA::set_products($one, scalar @$products, pack_products($products), $aln_len);
    SNPGenieCUDA::load_group_ptr($one, 0, $bases_ptr, $n_seqs, $aln_len);
    site_lines($fasta_file_name, SNPGenieCUDA::site_counts($one, 0, $aln_len), $site_fh);
    SNPGenieCUDA::destroy($one);
} else {
    site_lines($fasta_file_name, SNPGenieCUDA::site_counts($ctx, 0, $aln_len), $site_fh);
}
close $site_fh;
trace('site table');
SNPGenieCUDA::destroy($ctx);
trace('destroy');
printf "\nSNPGenie completed in %d seconds.\n\n", time - $time1;

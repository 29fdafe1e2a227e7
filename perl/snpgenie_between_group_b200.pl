#!/usr/bin/env perl
# Between-group dN/dS for every pair of aligned FASTA files (*.fa, *.fasta) in the working
# directory + one CDS GTF, computed on a B200 through libsnpgenie_cuda.so.  Drop-in for the
# reference's snpgenie_between_group.pl: same flags, same between_group_codon_results.txt and
# between_group_product_results.txt tables.  Group labels are the true file names (the
# reference permutes them after the first product; SURVEY.md quirk q2).
#
#   snpgenie_between_group_b200.pl --gtf_file=<cds>.gtf --procs_per_node=16
use strict;
use warnings;
use FindBin;
use lib "$FindBin::Bin/lib", "$FindBin::Bin/SNPGenieCUDA/lib";
use Getopt::Long;
use SNPGenieCUDA;
use SNPGenieHost qw(read_gtf pack_products ratio codon_of other_from_strings fmt_pairs);

STDOUT->autoflush(1);
my $time1 = time;
my ($gtf_file, $procs_per_node, $num_bootstraps);
my ($device, $seed, $minus_strand) = (0, 1, 0);
GetOptions("gtf_file=s" => \$gtf_file, "procs_per_node=i" => \$procs_per_node, "num_bootstraps=i" => \$num_bootstraps,
           "device=i" => \$device, "seed=i" => \$seed, "minus_strand" => \$minus_strand)
    or die "\n### WARNING: Error in command line arguments. Script terminated.\n\n";
unless (defined $gtf_file && $gtf_file =~ /.gtf/) {
    die "\n### WARNING: The --gtf_file option must be a file with a .gtf or .gtf extension\n### SNPGenie terminated.\n\n";
}
if (defined $procs_per_node && $procs_per_node < 1) {
    die "\n### WARNING: The --procs_per_node option must be an integer >=1\n### SNPGenie terminated.\n\n";
}
# all bootstrap code in the reference driver is commented out (bg:1050-1446): use the processor script

my @fasta_files = sort(glob("*.fa"), glob("*.fasta"));      # bg:1657-1668
die "\n\n## WARNING: There are no .fa or .fasta files. SNPGenie terminated.\n\n" unless @fasta_files;
die "\n\n## WARNING: Between-group analysis needs at least two FASTA files. SNPGenie terminated.\n\n" if @fasta_files < 2;
die "\n\n## WARNING: at most 64 groups are supported. SNPGenie terminated.\n\n" if @fasta_files > 64;

print "\n################################################################################"
    . "\n##                Between-Group SNPGenie (B200 engine) Initiated!             ##"
    . "\n################################################################################\n";

my $products = read_gtf($gtf_file, $minus_strand);
my $ctx = SNPGenieCUDA::create($device, $seed);
# histograms straight from the alignment bytes (no code matrix); the "first defined codon" rule (bg:853-890) needs sequence
# order only at conserved codons where one group is entirely undefined: a small kernel notes it while the rows go by
SNPGenieCUDA::set_option($ctx, SNPGenieCUDA::OPT_KEEP_CODES, 0);
SNPGenieCUDA::set_option($ctx, SNPGenieCUDA::OPT_FIRST_DEFINED, 1);
my (@ptr, @n_seqs, $aln_len);
for my $g (0 .. $#fasta_files) {
    my ($ptr, $n, $len) = SNPGenieCUDA::read_fasta($ctx, $g, $fasta_files[$g]);      # native reader, pinned memory
    die "\n\nDIE: The sequences must be aligned, i.e., must be the same length. TERMINATED.\n\n" if defined $aln_len && $len != $aln_len;
    $aln_len = $len;
    push @ptr, $ptr;
    push @n_seqs, $n;
    print "\ngroup index $g is group $fasta_files[$g]\n";
}
SNPGenieCUDA::set_products($ctx, scalar @$products, pack_products($products), $aln_len);
SNPGenieCUDA::load_group_ptr($ctx, $_, $ptr[$_], $n_seqs[$_], $aln_len) for 0 .. $#fasta_files;

open(my $codon_fh, '>>', 'between_group_codon_results.txt') or die "cannot write codon results: $!";
print $codon_fh "analysis\tproduct\tgroup_1\tgroup_2\tcodon\tvariability\tnum_defined_codons_g1\tnum_defined_codons_g2\t"
              . "comparisons\tN_sites\tS_sites\tN_diffs\tS_diffs\n";
open(my $prod_fh, '>>', 'between_group_product_results.txt') or die "cannot write product results: $!";
print $prod_fh "analysis\tproduct\tgroup_1\tgroup_2\tN_sites\tS_sites\tN_diffs\tS_diffs\tdN\tdS\tdN-dS\tdN/dS\n";
open(my $poly_fh, '>>', 'between_group_polymorphism.txt') or die "cannot write polymorphism file: $!";

my @tot = (0, 0, 0, 0);
for my $pi (0 .. $#$products) {
    my $p = $products->[$pi];
    my $name = $p->{name};
    my $C = $p->{n_codons};
    print "\n#########################################################\n################ Analyzing $name\n"
        . "#########################################################\n";
    my @hist = map { (SNPGenieCUDA::hist($ctx, $_, $pi))[0] } 0 .. $#fasta_files;
    for my $i (0 .. $#fasta_files) {
        for my $j ($i + 1 .. $#fasta_files) {
            print "\nComparing $C codons in $name between groups $i and $j...\n";
            my ($poly_s, $nda_s, $ndb_s, $cmp_s, $stats_s, $cons_s) = SNPGenieCUDA::between($ctx, $i, $j, $pi);
            my @poly = unpack('C*', $poly_s);
            my @cons = unpack('C*', $cons_s);
            my @nda = unpack('L*', $nda_s);
            my @ndb = unpack('L*', $ndb_s);
            my @stats = unpack('d*', $stats_s);
            print $poly_fh "\nDone recording polymorphism in $name between groups $i and $j\nIt is: @poly\n\n";
            my @sum = (0, 0, 0, 0);
            for my $c (0 .. $C - 1) {
                my ($Ns, $Ss, $Nd, $Sd) = @stats[4 * $c .. 4 * $c + 3];
                my $row = "ANALYSIS\t$name\t$fasta_files[$i]\t$fasta_files[$j]\tcodon_" . ($c + 1) . " \t";
                if ($poly[$c]) {
                    # cross pairs (gi allele - gj allele), bg:649-667
                    my @side;
                    for my $g ($i, $j) {
                        my @h = unpack('L*', substr($hist[$g], 4 * 66 * $c, 4 * 66));
                        my @al = map { codon_of($_) } grep { $h[$_] > 0 } 0 .. 63;
                        push @al, sort keys %{ other_from_strings(SNPGenieCUDA::codon_strings($ctx, $g, $pi, $c, $n_seqs[$g])) } if $h[65] > 0;
                        push @side, \@al;
                    }
                    my $pairs = join('', map { my $a = $_; map { "($a-$_)" } @{ $side[1] } } @{ $side[0] });
                    $row .= "polymorphic\t$nda[$c]\t$ndb[$c]\t$pairs\t$Ns\t$Ss\t$Nd\t$Sd";
                    $sum[$_] += ($Ns, $Ss, $Nd, $Sd)[$_] for 0 .. 3;
                } else {
                    my $codon = '---';                                   # bg:893-899
                    if ($cons[$c] < 64) { $codon = codon_of($cons[$c]); }
                    elsif ($cons[$c] == 65) {
                        my $g = $nda[$c] > 0 ? $i : $j;
                        ($codon) = sort keys %{ other_from_strings(SNPGenieCUDA::codon_strings($ctx, $g, $pi, $c, $n_seqs[$g])) };
                    }
                    $row .= "conserved\t$nda[$c]\t$ndb[$c]\t$codon\t$Ns\t$Ss\t0\t0";
                    $sum[0] += $Ns; $sum[1] += $Ss;
                }
                print $codon_fh "$row\n";
            }
            my $dN = ratio($sum[2], $sum[0]);
            my $dS = ratio($sum[3], $sum[1]);
            my ($w, $diff);
            {
                no warnings 'numeric';
                $w = ($dS > 0 && $dS ne '*') ? $dN / $dS : '*';           # bg:1497-1501
                $diff = ($dN >= 0 && $dS >= 0) ? $dN - $dS : '';
            }
            $tot[$_] += $sum[$_] for 0 .. 3;
            print $prod_fh "ANALYSIS\t$name\t$fasta_files[$i]\t$fasta_files[$j]\t$sum[0]\t$sum[1]\t$sum[2]\t$sum[3]\t$dN\t$dS\t$diff\t$w\n";
        }
    }
}
close $_ for $codon_fh, $prod_fh, $poly_fh;
SNPGenieCUDA::destroy($ctx);
print "\nTotals:\nN_sites: $tot[0]\nS_sites: $tot[1]\nN_diffs: $tot[2]\nS_diffs: $tot[3]\n\n";
printf "SNPGenie completed in %d seconds.\n\n", time - $time1;

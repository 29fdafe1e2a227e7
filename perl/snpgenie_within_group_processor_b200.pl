#!/usr/bin/env perl
# Sliding windows + per-window bootstrap SE over within_group_codon_results.txt, on a B200.
# Drop-in for the reference's snpgenie_within_group_processor.pl (same flags, same
# sw_<W>codons_results.txt table).
#
#   snpgenie_within_group_processor_b200.pl --codon_file=within_group_codon_results.txt \
#       --sliding_window_size=100 --sliding_window_step=10 --num_bootstraps=1000
use strict;
use warnings;
use FindBin;
use lib "$FindBin::Bin/lib", "$FindBin::Bin/SNPGenieCUDA/lib";
use Getopt::Long;
use SNPGenieCUDA;
use SNPGenieHost qw(ratio stars);

my ($codon_file, $W, $step, $B, $procs);
my ($device, $seed) = (0, time ^ ($$ << 15));
GetOptions("codon_file=s" => \$codon_file, "sliding_window_size=i" => \$W, "sliding_window_step=i" => \$step,
           "num_bootstraps=i" => \$B, "procs_per_node=i" => \$procs, "device=i" => \$device, "seed=i" => \$seed)
    or die "\n### WARNING: Error in command line arguments. Script terminated.\n\n";
die "\n### WARNING: The --codon_file option must be provided\n### SNPGenie terminated.\n\n" unless $codon_file;
$W = 10 unless defined $W;          # wgp:93-118 defaults
$step = 1 unless defined $step;
$B = 0 unless defined $B;
die "\n### WARNING: The --sliding_window_size option must be an integer >=1\n### SNPGenie terminated.\n\n" if $W < 1;
die "\n### WARNING: The --sliding_window_step option must be an integer >=1\n### SNPGenie terminated.\n\n" if $step < 1;
die "\n### WARNING: The --num_bootstraps option must be an integer >=2\n### SNPGenie terminated.\n\n" if $B != 0 && $B < 2;

open(my $in, '<', $codon_file) or die "\nCould not open the file $codon_file - $!\n\n";
my $header = <$in>;
chomp $header;
my @names = split(/\t/, $header, -1);
my %col = map { $names[$_] => $_ } 0 .. $#names;
for my $need (qw(product codon variability comparisons N_sites S_sites N_diffs S_diffs)) {
    die "### DIE: the header name $need is not present in the codon results file.\n\n" unless exists $col{$need};
}
my (%data, %sums);
while (my $line = <$in>) {
    chomp $line;
    next unless length $line;
    my @f = split(/\t/, $line, -1);
    next if $f[$col{product}] eq 'product';        # a second header from an appended run
    my $codon = $f[$col{codon}];
    $codon =~ s/codon_//;
    $codon =~ s/\s//g;
    $data{ $f[$col{product}] }{$codon} = [ map { $f[$col{$_}] } qw(N_sites S_sites N_diffs S_diffs) ];
    $sums{ $f[$col{product}] }[$_] += $f[$col{ (qw(N_sites S_sites N_diffs S_diffs))[$_] }] for 0 .. 3;
}
close $in;

my $out_file = "sw_${W}codons_results.txt";
open(my $out, '>>', $out_file) or die "cannot write $out_file: $!";
print $out "product\twindow\tfirst_codon\tlast_codon\tN_sites\tS_sites\tN_diffs\tS_diffs\tdN\tdS\t";
print $out "SE_dN\tSE_dS\t" if $B > 1;
print $out "dN_minus_dS\tdN_over_dS\tdN_gt_total_dS";
print $out($B > 1 ? "\tSE_dN_minus_dS\tZ_value\tsignificance\n" : "\n");

my $ctx = SNPGenieCUDA::create($device, $seed);
for my $product (sort keys %data) {
    my @codons = sort { $a <=> $b } keys %{ $data{$product} };
    my $C = scalar @codons;
    if ($W > $C) {
        warn "\n\n### WARNING:\n## For partition $product,\n## there are not enough codons to perform a sliding window of $W.\n### Product skipped.\n\n";
        next;
    }
    my $stats = pack('d*', map { @{ $data{$product}{$_} } } @codons);
    my ($nwin, $sums_s, $se_s) = SNPGenieCUDA::windows($ctx, $stats, undef, $C, $W, $step, $B);
    my @ws = unpack('d*', $sums_s);
    my @se = $B > 1 ? unpack('d*', $se_s) : ();
    my $product_dS = ratio($sums{$product}[3], $sums{$product}[1]);
    for my $k (0 .. $nwin - 1) {
        my ($Ns, $Ss, $Nd, $Sd) = @ws[4 * $k .. 4 * $k + 3];
        my $first = $k * $step + $codons[0];
        my $last = $first + $W - 1;
        my $dN = ratio($Nd, $Ns);
        my $dS = ratio($Sd, $Ss);
        no warnings 'numeric';
        my $w = ($dS > 0 && $dS ne '*') ? $dN / $dS : '*';
        my $diff = ($dN >= 0 && $dS >= 0) ? $dN - $dS : '';
        my $gt = ($dN > $product_dS) ? '*' : '';
        my $line = "$product\t" . ($k + 1) . "\t$first\t$last\t$Ns\t$Ss\t$Nd\t$Sd\t$dN\t$dS\t";
        if ($B > 1) {
            my ($se_dN, $se_dS, $se_diff) = @se[3 * $k .. 3 * $k + 2];
            my $Z = $se_diff > 0 ? $diff / $se_diff : 'NA';
            $line .= "$se_dN\t$se_dS\t$diff\t$w\t$gt\t$se_diff\t$Z\t" . stars($Z, $se_diff);
        } else {
            $line .= "$diff\t$w\t$gt";
        }
        print $out "$line\n";
    }
}
close $out;
SNPGenieCUDA::destroy($ctx);
print "\n\n";

#!/usr/bin/env perl
# Min-taxa filtering, sliding windows, whole-product bootstraps and Jukes-Cantor correction
# over between_group_codon_results.txt, on a B200.  Drop-in for the reference's
# snpgenie_between_group_processor.pl: same flags, same between_group_sw_<W>codons_results.txt
# and between_group_product_results_filtered.txt tables.
#
#   snpgenie_between_group_processor_b200.pl --between_group_codon_file=between_group_codon_results.txt \
#       --sliding_window_size=10 --sliding_window_step=1 --num_bootstraps=1000 \
#       [--codon_min_taxa_total=k] [--codon_min_taxa_group=k]
use strict;
use warnings;
use FindBin;
use lib "$FindBin::Bin/lib", "$FindBin::Bin/SNPGenieCUDA/lib";
use Getopt::Long;
use SNPGenieCUDA;
use SNPGenieHost qw(ratio stars);

my ($codon_file, $W, $step, $min_total, $min_group, $B, $procs);
my ($device, $seed, $boot_files) = (0, time ^ ($$ << 15), 0);
GetOptions("between_group_codon_file=s" => \$codon_file, "sliding_window_size=i" => \$W, "sliding_window_step=i" => \$step,
           "codon_min_taxa_total=i" => \$min_total, "codon_min_taxa_group=i" => \$min_group, "num_bootstraps=i" => \$B,
           "procs_per_node:i" => \$procs, "device=i" => \$device, "seed=i" => \$seed, "bootstrap_files" => \$boot_files)
    or die "\n### WARNING: Error in command line arguments. SNPGenie terminated.\n\n";
die "\n### WARNING: The --between_group_codon_file option must be provided\n### SNPGenie terminated.\n\n" unless $codon_file;
$W = 10 unless defined $W;
$step = 1 unless defined $step;
die "\n### WARNING: The --sliding_window_size option must be an integer >=1\n### SNPGenie terminated.\n\n" if $W < 1;
die "\n### WARNING: The --sliding_window_step option must be an integer >=1\n### SNPGenie terminated.\n\n" if $step < 1;
$min_total ||= 0;
$min_group ||= 0;
$B ||= 0;
die "\n### WARNING: The --num_bootstraps option must be an integer >=100\n### SNPGenie terminated.\n\n" if $B && $B < 100;   # bgp:158

my $sw_file = "between_group_sw_${W}codons_results.txt";
my $prod_file = 'between_group_product_results_filtered.txt';
die "\n###\tThe input file doesn't exist.\n\n" unless -f $codon_file;
for my $f ($sw_file, $prod_file) {            # bgp:196-201, :227-232: old outputs are replaced
    if (-f $f) { warn "\n### WARNING: $f already exists in this directory.\n### REMOVING AND REPLACING using current criteria.\n"; unlink $f; }
}

open(my $in, '<', $codon_file) or die "\nCould not open the file $codon_file - $!\n\n";
my $header = <$in>;
chomp $header;
my @names = split(/\t/, $header, -1);
my %col = map { $names[$_] => $_ } 0 .. $#names;
for my $need (qw(analysis product group_1 group_2 codon variability num_defined_codons_g1 num_defined_codons_g2 comparisons
                 N_sites S_sites N_diffs S_diffs)) {
    die "### DIE: the header name $need is not present in the between-group codon results file.\n\n" unless exists $col{$need};
}
my %data;    # {analysis}{g1}{g2}{product}{codon} = [Ns, Ss, Nd, Sd, ndef1, ndef2]
while (my $line = <$in>) {
    chomp $line;
    next unless length $line;
    my @f = split(/\t/, $line, -1);
    next if $f[$col{analysis}] eq 'analysis';
    my $codon = $f[$col{codon}];
    $codon =~ s/codon_//;
    $codon =~ s/\s//g;
    my ($g1, $g2) = sort ($f[$col{group_1}], $f[$col{group_2}]);          # bgp:326
    $data{ $f[$col{analysis}] }{$g1}{$g2}{ $f[$col{product}] }{$codon} =
        [ (map { $f[$col{$_}] } qw(N_sites S_sites N_diffs S_diffs)), $f[$col{num_defined_codons_g1}] || 0, $f[$col{num_defined_codons_g2}] || 0 ];
}
close $in;

open(my $sw, '>>', $sw_file) or die "cannot write $sw_file: $!";
print $sw "analysis\tproduct\tgroup_1\tgroup_2\twindow\tfirst_codon\tlast_codon\tnum_defined_codons_g1\tnum_defined_codons_g2\t"
        . "mean_num_defined_codons\tmin_num_defined_codons\tN_sites\tS_sites\tN_diffs\tS_diffs\tdN\tdS\t";
print $sw "SE_dN\tSE_dS\t" if $B > 1;
print $sw "dN_minus_dS\tdN_over_dS";
print $sw($B > 1 ? "\tSE_dN_minus_dS\tZ_value\tsignificance\n" : "\n");
open(my $pf, '>>', $prod_file) or die "cannot write $prod_file: $!";
print $pf "analysis\tproduct\tgroup_1\tgroup_2\tN_sites\tS_sites\tN_diffs\tS_diffs\tdN\tdS\tJC_dN\tJC_dS\t";
print $pf "SE_dN\tSE_dS\tSE_JC_dN\tSE_JC_dS\t" if $B > 1;
print $pf "dN_minus_dS\tdNdS\tJC_dN_minus_dS\tJC_dNdS";
print $pf($B > 1 ? "\tSE_dN_minus_dS\tZ_value\tsignificance\tSE_JC_dN_minus_dS\tJC_Z_value\tJC_significance\n" : "\n");

sub jc { my ($d) = @_; no warnings 'numeric'; return -3 / 4 * log(1 - ((4 / 3) * $d)); }     # bgp:483-485

my $ctx = SNPGenieCUDA::create($device, $seed);
for my $analysis (sort keys %data) {
  for my $g1 (sort keys %{ $data{$analysis} }) {
    for my $g2 (sort keys %{ $data{$analysis}{$g1} }) {
      for my $product (sort keys %{ $data{$analysis}{$g1}{$g2} }) {
        my $D = $data{$analysis}{$g1}{$g2}{$product};
        my @codons = sort { $a <=> $b } keys %$D;
        my $C = scalar @codons;
        die "\n\n### WARNING:\n## For product $product,\n## there are not enough codons to perform a sliding window of $W.\n\n" if $W > $C;
        # min-taxa mask (bgp:412-467): a codon that fails contributes nothing anywhere
        my @keep = map {
            my ($n1, $n2) = @{ $D->{$_} }[4, 5];
            my $min = $n1 < $n2 ? $n1 : $n2;
            (($min_total > 0 ? $n1 + $n2 >= $min_total : 1) && ($min_group > 0 ? $min >= $min_group : 1)) ? 1 : 0;
        } @codons;
        my $stats = pack('d*', map { @{ $D->{$_} }[0 .. 3] } @codons);
        my $keep = pack('C*', @keep);
        my @tot = (0, 0, 0, 0);
        for my $i (0 .. $#codons) { next unless $keep[$i]; $tot[$_] += $D->{ $codons[$i] }[$_] for 0 .. 3; }
        my $dN = ratio($tot[2], $tot[0]);
        my $dS = ratio($tot[3], $tot[1]);
        my ($JC_dN, $JC_dS) = (jc($dN), jc($dS));

        # sliding windows (bgp:501-837)
        my ($nwin, $sums_s, $se_s) = SNPGenieCUDA::windows($ctx, $stats, $keep, $C, $W, $step, $B);
        my @ws = unpack('d*', $sums_s);
        my @wse = $B > 1 ? unpack('d*', $se_s) : ();
        for my $k (0 .. $nwin - 1) {
            my ($Ns, $Ss, $Nd, $Sd) = @ws[4 * $k .. 4 * $k + 3];
            my $first = $k * $step + $codons[0];
            my $last = $first + $W - 1;
            my ($d1, $d2, $min_def) = (0, 0, 1000);
            for my $c ($first .. $last) {
                my ($n1, $n2) = @{ $D->{$c} }[4, 5];
                $d1 += $n1; $d2 += $n2;
                $min_def = $n1 if $n1 < $min_def;
                $min_def = $n2 if $n2 < $min_def;
            }
            my $wdN = ratio($Nd, $Ns);
            my $wdS = ratio($Sd, $Ss);
            no warnings 'numeric';
            my $w = ($wdS > 0 && $wdS ne '*') ? $wdN / $wdS : '*';
            my $diff = ($wdN >= 0 && $wdS >= 0) ? $wdN - $wdS : '';
            ($Ns, $Ss, $Nd, $Sd) = ('NA') x 4 if !$Ns && !$Ss && !$Nd && !$Sd;        # bgp:778-785
            my $line = "$analysis\t$product\t$g1\t$g2\t" . ($k + 1) . "\t$first\t$last\t$d1\t$d2\t" . (($d1 + $d2) / 2)
                     . "\t$min_def\t$Ns\t$Ss\t$Nd\t$Sd\t$wdN\t$wdS\t";
            if ($B > 1) {
                my ($se_dN, $se_dS, $se_diff) = @wse[3 * $k .. 3 * $k + 2];
                my $Z = $se_diff > 0 ? $diff / $se_diff : 'NA';
                my $sig = $se_diff > 0 ? stars($Z, $se_diff) : 'NA';
                $se_diff = 'NA' if $se_diff == 0;
                ($se_diff, $Z, $sig) = ('NA', 'NA', 'NA') if $W == 1;                   # bgp:819-823
                $line .= "$se_dN\t$se_dS\t$diff\t$w\t$se_diff\t$Z\t$sig";
            } else {
                $line .= "$diff\t$w";
            }
            print $sw "$line\n";
        }

        # whole-product bootstrap (bgp:869-1102)
        my @se;
        my ($Z, $JC_Z) = ('NA', 'NA');
        my ($diff, $JC_diff);
        {
            no warnings 'numeric';
            $diff = $dN - $dS;
            $JC_diff = $JC_dN - $JC_dS;
        }
        if ($B > 1) {
            print "\n#########################################################################################\n"
                . "Bootstrapping $product between-group $g1 vs. $g2 for dN/dS standard error...\n";
            my ($se_s2, $sums_s2) = SNPGenieCUDA::bootstrap_product($ctx, $stats, $keep, $C, $B, $boot_files ? 1 : 0);
            @se = unpack('d*', $se_s2);
            if ($boot_files) {
                open(my $bfh, '>>', "${product}_bootstrap_results.txt") or die "cannot write bootstrap file: $!";
                print $bfh "analysis\tgroup1\tgroup2\tproduct_name\tbootstrap_num\tsim_product_N_sites_sum\tsim_product_S_sites_sum\t"
                         . "sim_product_N_diffs_sum\tsim_product_S_diffs_sum\n";
                my @sums = unpack('d*', $sums_s2);
                print $bfh join("\t", $analysis, $g1, $g2, $product, $_ + 1, @sums[4 * $_ .. 4 * $_ + 3]) . "\n" for 0 .. $B - 1;
                close $bfh;
            }
            no warnings 'numeric';
            $Z = $diff / $se[2] if $se[2] > 0;
            $JC_Z = $JC_diff / $se[5] if $se[5] > 0;
        }
        my ($w, $JC_w);
        {
            no warnings 'numeric';
            $w = $dS > 0 ? $dN / $dS : '*';
            $JC_w = $JC_dS > 0 ? $JC_dN / $JC_dS : '*';
        }
        my $line = "$analysis\t$product\t$g1\t$g2\t$tot[0]\t$tot[1]\t$tot[2]\t$tot[3]\t$dN\t$dS\t$JC_dN\t$JC_dS\t";
        $line .= "$se[0]\t$se[1]\t$se[3]\t$se[4]\t" if $B > 1;
        $line .= "$diff\t$w\t$JC_diff\t$JC_w";
        if ($B > 1) {
            my ($se_d, $se_j) = ($se[2] == 0 ? 'NA' : $se[2], $se[5] == 0 ? 'NA' : $se[5]);
            $line .= "\t$se_d\t$Z\t" . stars($Z, $se_d) . "\t$se_j\t$JC_Z\t" . stars($JC_Z, $se_j);
        }
        print $pf "$line\n";
      }
    }
  }
}
close $sw;
close $pf;
SNPGenieCUDA::destroy($ctx);
print "\n\n";

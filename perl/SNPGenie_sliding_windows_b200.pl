#!/usr/bin/env perl
# Sliding windows with bootstrap standard errors, Z tests and achieved significance levels over a per-codon results
# table, on a B200.  Drop-in for the reference's SNPGenie_sliding_windows.R (the script its README recommends over the Perl
# window processors): the same positional arguments, the same <input>_WINDOWS_dNdS.tsv table (the input's columns followed
# by the sw_* columns on the row of each window's first codon, NA elsewhere).
#
#   SNPGenie_sliding_windows_b200.pl within_group_codon_results.txt N S 40 1 1000 6 NONE 1 [prefix]
#
# Arguments (SNPGenie_sliding_windows.R:49-60): (1) codon results file, (2) numerator site type N, (3) denominator site
# type S or N, (4) window size >= 2, (5) step >= 1, (6) bootstrap replicates (default 1000), (7) minimum number of defined
# codons (default 6), (8) NONE or JC, (9) CPUs (accepted, unused: the GPU does the resampling), (10) output line prefix.
# Extra environment: SNPG_DEVICE, SNPG_SEED.
use strict;
use warnings;
use FindBin;
use lib "$FindBin::Bin/lib", "$FindBin::Bin/SNPGenieCUDA/lib";
use SNPGenieCUDA;

sub usage {
    print "\n\n### WARNING: there must be 5-10 command line arguments, in this order:\n";
    print "    (1) CODON RESULTS FILE\n    (2) NUMERATOR SITE TYPE (\"N\"; e.g., \"N\" for dN)\n";
    print "    (3) DENOMINATOR SITE TYPE (\"S\" or \"N\"; e.g., \"S\" for dS)\n    (4) SLIDING WINDOW SIZE (CODONS; >=2; >=10 recommended)\n";
    print "    (5) SLIDING WINDOW STEP SIZE (CODONS; >=1)\n    (6) NUMBER OF BOOTSTRAP REPLICATES PER WINDOW (OPTIONAL; >=2; DEFAULT=1000)\n";
    print "    (7) MINIMUM NUMBER OF DEFINED CODONS PER CODON POSITION (OPTIONAL; >=2; DEFAULT=6)\n";
    print "    (8) MULTIPLE HITS CORRECTION (OPTIONAL; \"NONE\" or \"JC\"; DEFAULT=NONE)\n    (9) NUMBER OF CPUS (OPTIONAL; unused here)\n";
    print "    (10) STRING TO PREPEND TO OUTPUT LINES (OPTIONAL; DEFAULT=\"\")\n\n";
    exit 1;
}
my @A = @ARGV;
if (@A < 5) { print "\n### Command line Error 1.\n"; usage(); }
if ($A[0] !~ /\w/) { print "\n### Command line Error 2.\n"; usage(); }
if ($A[1] ne 'N') { print "\n### Command line Error 3.\n"; usage(); }                       # (the tables carry N and S columns only)
if ($A[2] ne 'S' && $A[2] ne 'N') { print "\n### Command line Error 4.\n"; usage(); }
if ($A[3] !~ /^\d+$/ || $A[3] < 2) { print "\n### Command line Error 5.\n"; usage(); }
if ($A[4] !~ /^\d+$/ || $A[4] < 1) { print "\n### Command line Error 6.\n"; usage(); }
my ($file, $num, $den, $W, $step) = @A[0 .. 4];
my $B = (defined $A[5] && $A[5] =~ /^\d+$/) ? $A[5] : 1000;
my $min_def = (defined $A[6] && $A[6] =~ /^\d+$/) ? $A[6] : 6;
if (!($B >= 2 && $B <= 1000000)) { print "\n### WARNING: NBOOTSTRAPS must be in the range [2,1000000]. Using: 1000.\n"; $B = 1000; }
if (!($min_def >= 2)) { print "\n### WARNING: MIN_DEFINED_CODONS must be >=2. Using: 6.\n"; $min_def = 6; }
my $correction = 'NONE';
if (defined $A[7] && $A[7] eq 'JC') { $correction = 'JC'; }
elsif (!defined $A[7] || $A[7] ne 'NONE') { print "\n### WARNING: unrecognized CORRECTION supplied. Using: \"NONE\".\n"; }
my $prefix = defined $A[9] ? $A[9] : '';

open(my $in, '<', $file) or die "\nCould not open the file $file - $!\n\n";
my $header = <$in>;
defined $header or die "\n### $file is empty.\n\n";
chomp $header;
my @names = split(/\t/, $header, -1);
my %col = map { $names[$_] => $_ } 0 .. $#names;
for my $need (qw(product N_sites S_sites N_diffs S_diffs)) {
    die "### DIE: the header name $need is not present in the codon results file.\n\n" unless exists $col{$need};
}
my $pooled = exists $col{N_diffs_vs_ref};                 # R :144-177: a classic snpgenie.pl table (codons numbered from their sites)
die "### DIE: the codon results file has neither a codon_num nor a codon column.\n\n" unless $pooled || exists $col{codon_num} || exists $col{codon};
my $between = exists $col{num_defined_codons_g1};         # R :181-194
my $nd_col = $between ? undef : (exists $col{num_defined_seqs} ? 'num_defined_seqs' : exists $col{num_defined_codons} ? 'num_defined_codons' : undef);
if ($between) {
    print "\n###############################################################################\n";
    print "BETWEEN-GROUP FILE DETECTED. SNPGenie will require both groups to have at least\n";
    print "    the user-specified number of MIN_DEFINED_CODONS: $min_def .\n";
}
my (@rows, %products, %seen);
while (my $line = <$in>) {
    chomp $line;
    next unless length $line;
    my @f = split(/\t/, $line, -1);
    my $cn = $pooled ? $f[$col{site}] : exists $col{codon_num} && !$between ? $f[$col{codon_num}] : $f[$col{codon}];
    $cn =~ s/codon_//;
    $cn =~ s/\s//g;
    my $nd = $pooled ? 100 : $between ? ($f[$col{num_defined_codons_g1}] < $f[$col{num_defined_codons_g2}] ? $f[$col{num_defined_codons_g1}] : $f[$col{num_defined_codons_g2}])
           : defined $nd_col ? $f[$col{$nd_col}] : undef;
    push @rows, { f => \@f, cn => $cn, nd => $nd };
    $products{ $f[$col{product}] } = 1;
    $seen{$cn}++;
}
close $in;
if ($pooled) {
    # number the codons by their starting sites (R :146-166); a minimum read depth of 100 is assumed (R :169)
    my ($min) = sort { $a <=> $b } map { $_->{cn} } @rows;
    if (grep { ($_->{cn} - $min + 3) % 3 } @rows) {
        print "\n###############################################################################\n";
        print "WARNING: the input seems to contain codons for more than one gene/product, a frameshift,\n    or multiple exons. Codons will be numbered by their relative starting site position.\n";
        @rows = sort { $a->{cn} <=> $b->{cn} } @rows;
        $rows[$_]{cn} = $_ + 1 for 0 .. $#rows;
    } else {
        $_->{cn} = ($_->{cn} - $min + 3) / 3 for @rows;
    }
    %seen = ();
    $seen{ $_->{cn} }++ for @rows;
    print "\n###############################################################################\n";
    print "ALERT: bootstrapping assumes reliable estimates of N and S differences for all codons.\n";
}
if (keys(%products) != 1 || grep { $_ > 1 } values %seen) {                                  # R :197-202
    print "\n\n### WARNING: there must be results for ONLY ONE GENE PRODUCT PER FILE, i.e., a unique set of codons per frame.\n";
    print "### Additionally, for between-group comparisons, there must be results for ONLY ONE GROUP-TO-GROUP COMPARISON.\n### TERMINATED.\n\n";
    exit 1;
}
if ($W < 10) {
    print "\n###############################################################################\n";
    print "ALERT: user-specified WINDOW_SIZE=$W\n    With rare exceptions, windows should be at least 10 codons in width.\n";
}
print "\n###############################################################################\nSNPGenie sliding window analysis LOG\n\nPARAMETERS:\n";
print "(1) CODON_RESULTS_FILE=$file\n(2) NUMERATOR=$num\n(3) DENOMINATOR=$den\n(4) WINDOW_SIZE=$W\n(5) STEP_SIZE=$step\n(6) NBOOTSTRAPS=$B\n";
print "(7) MIN_DEFINED_CODONS=$min_def\n(8) CORRECTION=$correction\n(9) NCPUS=GPU\n(10) PREPEND_TO_OUTPUT=$prefix\n\n";

my $C = scalar @rows;
my $stats = pack('d*', map { my $r = $_; map { $r->{f}[$col{$_}] } qw(N_sites S_sites N_diffs S_diffs) } @rows);
my $have_nd = !grep { !defined $_->{nd} } @rows;
my $nd = $have_nd ? pack('L*', map { $_->{nd} } @rows) : undef;
my $cn = pack('q*', map { $_->{cn} } @rows);
my $ctx = SNPGenieCUDA::create($ENV{SNPG_DEVICE} // 0, $ENV{SNPG_SEED} // (time ^ ($$ << 15)));
print "Performing sliding window analysis...\n\n";
my ($nwin, $packed) = SNPGenieCUDA::sliding_windows_r($ctx, $stats, $nd, $cn, $C, 0, $den eq 'S' ? 1 : 0, $correction eq 'JC' ? 1 : 0, $W, $step, $B, $min_def);
SNPGenieCUDA::destroy($ctx);
if ($nwin == 0) {
    print "\n\n### WARNING: the number of codons with num_defined_seqs>=$min_def is less than the WINDOW_SIZE=$W.\n";
    print "### Thus, a sliding window using these parameters is not possible.\n### TERMINATED.\n\n";
    exit 1;
}
my @v = unpack('d*', $packed);
my $ncol = 24;
my %by_start;                                             # window values by the codon number of the window's first row
$by_start{ int($v[$ncol * $_]) } = [ @v[$ncol * $_ .. $ncol * $_ + $ncol - 1] ] for 0 .. $nwin - 1;

my $ratio = "d${num}d${den}";
(my $out_file = $file) =~ s/\.txt//;
$out_file =~ s/\.tsv//;
$out_file .= "_WINDOWS_${ratio}.tsv";
print "Writing output to: $out_file\n\n";
my @sw = qw(sw_ratio sw_start sw_center sw_end sw_num_replicates sw_N_diffs sw_S_diffs sw_N_sites sw_S_sites sw_dN sw_dS sw_dNdS sw_dN_m_dS
            sw_boot_dN_SE sw_boot_dS_SE sw_boot_dN_over_dS_SE sw_boot_dN_over_dS_P sw_boot_dN_m_dS_SE sw_boot_dN_m_dS_P sw_boot_dN_gt_dS_count
            sw_boot_dN_eq_dS_count sw_boot_dN_lt_dS_count sw_ASL_dN_gt_dS_P sw_ASL_dN_lt_dS_P sw_ASL_dNdS_P);
sub fmt {                                                 # R prints NaN / Inf / NA; integers without a fraction
    my ($x) = @_;
    return 'NaN' if $x != $x;
    return $x > 0 ? 'Inf' : '-Inf' if $x == 9**9**9 || $x == -9**9**9;
    return sprintf('%.15g', $x);
}
open(my $out, '>', $out_file) or die "cannot write $out_file: $!";
my @extra = $between ? qw(num_defined_seqs codon_num) : $pooled ? qw(codon_num num_defined_seqs) : ();
print $out $prefix . join("\t", @names, @extra, @sw) . "\n";
for my $r (@rows) {
    my @line = @{ $r->{f} };
    push @line, $r->{nd}, $r->{cn} if $between;
    push @line, $r->{cn}, $r->{nd} if $pooled;
    my $w = $by_start{ $r->{cn} };
    push @line, $ratio, ($w ? map { fmt($_) } @$w : ('NA') x $ncol);
    print $out $prefix . join("\t", @line) . "\n";
}
close $out;
print "DONE\n\n";

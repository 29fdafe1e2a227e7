#!/usr/bin/env perl
# Dump the reference's Nei-Gojobori tables by CALLING the reference's own subs.
#
# The two subs are lifted at run time out of the read-only reference script
# (nothing is copied into this repo): get_number_of_sites
# (snpgenie_within_group.pl:1625-1721) and return_avg_diffs (:1728-17996).
# Output: one TSV on STDOUT, %.17g so every double round-trips exactly.
#   S <codon> <pos> <N> <S>           64*3 rows
#   D <codon1> <codon2> <N> <S>       64*64 rows (diagonal/undef printed as 0)
# Usage: perl tools/dump_ref_tables.pl [/root/reference/snpgenie_within_group.pl] > tests/golden/ng_tables.tsv
use strict;
use warnings;

my $src = $ARGV[0] // '/root/reference/snpgenie_within_group.pl';
open(my $fh, '<', $src) or die "cannot open $src: $!";
my @lines = <$fh>;
close $fh;

sub grab_sub {
    my ($name) = @_;
    my $start;
    for my $i (0 .. $#lines) {
        if ($lines[$i] =~ /^sub \Q$name\E\b/) { $start = $i; last; }
    }
    die "sub $name not found" unless defined $start;
    my $end;
    for my $i ($start + 1 .. $#lines) {
        if ($lines[$i] =~ /^}/) { $end = $i; last; }
    }
    return join('', @lines[$start .. $end]);
}

my $code = "no strict; no warnings;\n" . grab_sub('get_number_of_sites') . grab_sub('return_avg_diffs') . "1;\n";
eval $code or die "eval failed: $@";

my @nt = qw(A C G T);
my @codons;
for my $a (@nt) { for my $b (@nt) { for my $c (@nt) { push @codons, "$a$b$c"; } } }

for my $c (@codons) {
    for my $p (1 .. 3) {
        my @s = get_number_of_sites($c, $p);
        printf("S\t%s\t%d\t%.17g\t%.17g\n", $c, $p, $s[0] // 0, $s[1] // 0);
    }
}
# return_avg_diffs rebuilds its hash per call (~2 ms); build it once by
# re-evaluating the sub body with the hash hoisted is not possible without
# editing reference code, so just call it 4,096 times (~10 s).
for my $c1 (@codons) {
    for my $c2 (@codons) {
        my @d = (0, 0);
        @d = return_avg_diffs($c1, $c2) if $c1 ne $c2;
        # two literals (TAA<->TGG) are the string '*', which Perl numifies to 0
        # wherever the reference multiplies it by a weight; dump what it computes with.
        my @v = map { (defined $_ && $_ ne '*') ? $_ : 0 } @d[0, 1];
        printf("D\t%s\t%s\t%.17g\t%.17g\n", $c1, $c2, $v[0], $v[1]);
    }
}

package SNPGenieHost;
# Host-side helpers shared by the four *_b200.pl command-line programs: the parts of
# SNPGenie's within-/between-group drivers that stay in Perl (flags, GTF and FASTA
# ingest, output tables).  The per-codon engine, bootstraps and windows are calls
# into SNPGenieCUDA (libsnpgenie_cuda.so).
#
# Behaviour follows the reference scripts (cited as wg = snpgenie_within_group.pl,
# bg = snpgenie_between_group.pl); nothing here is copied from them.
use strict;
use warnings;
use Exporter 'import';
our @EXPORT_OK = qw(read_gtf read_fasta pack_products ratio stars codon_of other_alleles fmt_pairs amino_acid variant_line site_lines other_from_strings);

my @NT = qw(A C G T);
sub codon_of { my ($c) = @_; return $NT[$c >> 4] . $NT[($c >> 2) & 3] . $NT[$c & 3]; }

# GTF model (wg:1368-1392, :1530-1621): a CDS record matches
# /CDS\t\d+\t\d+\t[\.\d+]\t\+/; the product is its gene_id (three forms, tried in
# this order); segments are sorted by start; the total must be a multiple of 3.
# With $minus_too, '-' strand records become strand -1 products (the reference
# ignores them and asks the user to reverse-complement the inputs instead).
sub read_gtf {
    my ($file, $minus_too) = @_;
    my (%seen, @order);
    open(my $fh, '<', $file) or die "\n## Cannot open $file. TERMINATED.\n\n";
    while (my $line = <$fh>) {
        my $strand;
        if ($line =~ /CDS\t\d+\t\d+\t[\.\d+]\t\+/) { $strand = 1; }
        elsif ($minus_too && $line =~ /CDS\t\d+\t\d+\t[\.\d+]\t\-/) { $strand = -1; }
        else { next; }
        my $name;
        if ($line =~ /gene_id \"gene\:([\w\s\.\-\:']+)\"/) { $name = $1; }
        elsif ($line =~ /gene_id \"([\w\s\.\-\:']+ [\w\s\.\-\:']+)\"/) { $name = $1; }
        elsif ($line =~ /gene_id \"([\w\s\.\-\:']+)\"/) { $name = $1; }
        else { die "\n\n## WARNING: CDS annotation(s) in $file does not have a gene_id. SNPGenie terminated.\n\n"; }
        my ($start, $stop) = $line =~ /CDS\t(\d+)\t(\d+)/;
        my $key = "$name\t$strand";
        unless ($seen{$key}) { $seen{$key} = { name => $name, strand => $strand, segs => [] }; push @order, $key; }
        push @{ $seen{$key}{segs} }, [ $start, $stop, scalar @{ $seen{$key}{segs} } ];
    }
    close $fh;
    my @products;
    for my $key (@order) {
        my $p = $seen{$key};
        my @segs = sort { $a->[0] <=> $b->[0] || $a->[2] <=> $b->[2] } @{ $p->{segs} };
        my $total = 0;
        $total += $_->[1] - $_->[0] + 1 for @segs;
        if ($total % 3 != 0) {
            die "\n\n## WARNING: The CDS coordinates for gene $p->{name} in the gtf file do not yield a set of complete codons,\n"
              . "## or are absent from the file. The number of nucleotides must be a multiple of 3.\n## SNPGenie terminated.\n\n";
        }
        push @products, { name => $p->{name}, strand => $p->{strand}, starts => [ map { $_->[0] } @segs ],
                          stops => [ map { $_->[1] } @segs ], n_codons => $total / 3 };
    }
    die "\n## No CDS records found in $file. TERMINATED.\n\n" unless @products;
    # the reference sorts names with a numeric comparison (wg:189), i.e. arbitrarily for text names;
    # a plain string sort keeps runs reproducible
    @products = sort { $a->{name} cmp $b->{name} || $b->{strand} <=> $a->{strand} } @products;
    return \@products;
}

# Aligned FASTA (wg:204-253): lines containing '>' start a record, other lines are
# chomp'ed and concatenated; all records must have one length.  Returns the
# residues of all records as ONE string (row-major) plus count and length.  Case
# and U are left alone: the device applies uc and U->T.
sub read_fasta {
    my ($file) = @_;
    open(my $fh, '<', $file) or die "Could not open file $file\n";
    my ($all, $cur, $n, $len) = ('', '', 0, undef);
    my $started = 0;
    my $flush = sub {
        if (defined $len && length($cur) != $len) {
            die "\n\nDIE: The sequences must be aligned, i.e., must be the same length. TERMINATED.\n\n";
        }
        $len = length($cur);
        $all .= $cur;
        $n++;
        $cur = '';
    };
    while (my $line = <$fh>) {
        chomp $line;
        if ($line =~ />/) { $flush->() if $started; $started = 1; }
        elsif ($started) { $cur .= $line; }
    }
    close $fh;
    $flush->() if $started;
    die "\n## No sequences in $file. TERMINATED.\n\n" unless $n;
    return ($all, $n, $len);
}

# CSR arrays for SNPGenieCUDA::set_products
sub pack_products {
    my ($products) = @_;
    my (@ofs, @st, @sp, @strand);
    push @ofs, 0;
    for my $p (@$products) {
        push @st, @{ $p->{starts} };
        push @sp, @{ $p->{stops} };
        push @ofs, scalar @st;
        push @strand, $p->{strand};
    }
    return (pack('q*', @ofs), pack('q*', @st), pack('q*', @sp), pack('c*', @strand));
}

# d = diffs/sites with the reference's '*' for an empty denominator (wg:985-996)
sub ratio { my ($num, $den) = @_; return $den > 0 ? $num / $den : '*'; }

# significance stars (wg:1036-1046)
sub stars {
    my ($z, $se) = @_;
    return '' unless $se ne 'NA' && $se > 0;
    return abs($z) > 2.81 ? '***' : abs($z) > 1.96 ? '**' : abs($z) > 1.64 ? '*' : '';
}

# 'other' alleles (defined codons that are not pure ACGT) need their strings, which the
# histogram does not carry: scan that one column of the host copy.  Returns {string => count}.
sub other_alleles {
    my ($seqs_ref, $n, $len, $product, $codon_index) = @_;
    my @pos;
    for my $i (0 .. $#{ $product->{starts} }) { push @pos, ($product->{starts}[$i] - 1) .. ($product->{stops}[$i] - 1); }
    @pos = reverse @pos if $product->{strand} < 0;
    my @p3 = @pos[3 * $codon_index .. 3 * $codon_index + 2];
    my %count;
    for my $s (0 .. $n - 1) {
        my $codon = join('', map { substr($$seqs_ref, $s * $len + $_, 1) } @p3);
        $codon = uc $codon;
        $codon =~ tr/ACGT/TGCA/ if $product->{strand} < 0;
        $codon =~ tr/U/T/;
        next if $codon =~ /N/ || $codon =~ /-/;
        next if $codon =~ /^[ACGT]{3}$/;
        $count{$codon}++;
    }
    return \%count;
}

# same from the packed [n][3] normalised characters SNPGenieCUDA::codon_strings returns
sub other_from_strings {
    my ($packed) = @_;
    my %count;
    for my $s (0 .. length($packed) / 3 - 1) {
        my $codon = substr($packed, 3 * $s, 3);
        next if $codon =~ /N/ || $codon =~ /-/ || $codon =~ /^[ACGT]{3}$/;
        $count{$codon}++;
    }
    return \%count;
}

# the "comparisons" column of a polymorphic codon: every unordered allele pair once, plus
# (X-X) for alleles seen at least twice (the reference lists ordered pairs in hash order, wg:618-632)
sub fmt_pairs {
    my ($alleles) = @_;    # [ [string, count], ... ]
    my $out = '';
    for my $i (0 .. $#$alleles) {
        $out .= "($alleles->[$i][0]-$alleles->[$i][0])" if $alleles->[$i][1] > 1;
        for my $j ($i + 1 .. $#$alleles) { $out .= "($alleles->[$i][0]-$alleles->[$j][0])"; }
    }
    return $out;
}

# get_amino_acid (wg:1400-1516): standard code, '?' for anything that is not a pure ACGT codon
my @AA = split //, 'KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF';
my %AA_OF = map { codon_of($_) => $AA[$_] } 0 .. 63;
sub amino_acid { my ($c) = @_; return exists $AA_OF{$c} ? $AA_OF{$c} : '?'; }

# one within_group_variant_results.txt row (wg:1061-1167) from a codon's alleles [[string, count], ...];
# counts are the true integer counts (the reference's are distorted by gaps, quirk q1)
sub variant_line {
    my ($file, $product, $codon_num, $alleles) = @_;
    my @al = sort { $a->[0] cmp $b->[0] } @$alleles;
    my ($consensus, $best, $total) = ('', 0, 0);
    for my $a (@al) { $total += $a->[1]; if ($a->[1] > $best) { $best = $a->[1]; $consensus = $a->[0]; } }
    my @aas = map { amino_acid($_->[0]) } @al;
    my ($syn, $non) = (0, 0);
    for my $i (0 .. $#aas) { for my $j ($i + 1 .. $#aas) { if ($aas[$i] eq $aas[$j]) { $syn++ } else { $non++ } } }
    my $type = ($non > 0 && $syn > 0) ? 'Ambiguous' : $non > 0 ? 'Nonsynonymous' : 'Synonymous';
    return join("\t", $file, $product, $codon_num, $consensus, scalar(@al), join(',', map { $_->[0] } @al), join(',', @aas), $type,
                join(',', map { $_->[1] } @al), $total, join(',', map { sprintf('%.3f', $_->[1] / $total) } @al), $best,
                $total - $best, $total) . "\n";
}

# within_group_site_results.txt rows (wg:1242-1293) from packed [L][4] u32 counts
sub site_lines {
    my ($file, $counts_packed, $fh) = @_;
    my @c = unpack('L*', $counts_packed);
    for my $pos (0 .. @c / 4 - 1) {
        my ($A, $C, $G, $T) = @c[4 * $pos .. 4 * $pos + 3];
        my ($maj, $maj_count) = ('A', $A);
        ($maj, $maj_count) = ('C', $C) if $C > $maj_count;
        ($maj, $maj_count) = ('G', $G) if $G > $maj_count;
        ($maj, $maj_count) = ('T', $T) if $T > $maj_count;
        my $pw = $A * $C + $A * $G + $A * $T + $C * $G + $C * $T + $G * $T;
        my $cov = $A + $C + $G + $T;
        my $comps = ($cov**2 - $cov) / 2;
        my $pi = $comps > 0 ? $pw / $comps : 'NA';
        print $fh join("\t", $file, $pos + 1, $maj, $maj_count, $pi, $cov, $A, $C, $G, $T) . "\n";
    }
}

1;

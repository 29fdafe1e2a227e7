package SNPGenieCUDA;
# Perl binding of libsnpgenie_cuda.so (the B200 engine behind the within-/between-group
# SNPGenie drivers).  Thin on purpose: the XS layer maps one-to-one onto
# include/snpgenie_cuda.h; arrays travel as packed strings.
use strict;
use warnings;
require XSLoader;
our $VERSION = '0.1';
XSLoader::load('SNPGenieCUDA', $VERSION);

use constant { CODE_UNDEF => 64, CODE_OTHER => 65, HIST_BINS => 66, OPT_KEEP_CODES => 1, OPT_PROFILE => 2, OPT_SITE_COUNTS => 3,
               OPT_FUSED_KERNEL => 4, OPT_BOOTSTRAP_KERNEL => 5, OPT_BOOT_SLOTS => 6, OPT_INGEST_ASYNC => 7, OPT_FIRST_DEFINED => 8, BOOT_GATHER => 0, BOOT_WEIGHTS => 1, BOOT_CLASSES => 2 };

my @NT = qw(A C G T);
# codon string of a code 0..63
sub codon_of { my ($c) = @_; return $NT[$c >> 4] . $NT[($c >> 2) & 3] . $NT[$c & 3]; }

1;

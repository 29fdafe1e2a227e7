/*
 * SNPGenieCUDA.xs -- Perl XS glue over the C ABI of libsnpgenie_cuda.so
 * (include/snpgenie_cuda.h).  Arrays cross the boundary as packed Perl strings
 * (pack "q*", "d*", ...): the caller packs inputs, the glue sizes output SVs and
 * lets the library write straight into them.  Errors croak with the library's
 * message, which is what the reference's die() calls look like to a user.
 */
#define PERL_NO_GET_CONTEXT
#include "EXTERN.h"
#include "perl.h"
#include "XSUB.h"

#include "snpgenie_cuda.h"

static snpg_ctx *ctx_of(pTHX_ IV h) {
    if (!h) croak("SNPGenieCUDA: null context");
    return INT2PTR(snpg_ctx *, h);
}

static void check(pTHX_ snpg_ctx *ctx, int rc, const char *what) {
    if (rc != SNPG_OK) croak("\n### SNPGenieCUDA %s failed (%d): %s\n", what, rc, snpg_last_error(ctx));
}

/* a fresh mortal string SV with room for n bytes, length already set */
static SV *out_buf(pTHX_ STRLEN n) {
    SV *sv = sv_2mortal(newSV(n ? n : 1));
    SvPOK_only(sv);
    SvCUR_set(sv, n);
    if (n) memset(SvPVX(sv), 0, n);
    *SvEND(sv) = '\0';
    return sv;
}

static const void *in_buf(pTHX_ SV *sv, STRLEN need, const char *name) {
    STRLEN len;
    const char *p = SvPVbyte(sv, len);
    if (len < need) croak("SNPGenieCUDA: %s holds %lu bytes, %lu needed", name, (unsigned long)len, (unsigned long)need);
    return p;
}

MODULE = SNPGenieCUDA    PACKAGE = SNPGenieCUDA

PROTOTYPES: DISABLE

const char *
version()
  CODE:
    RETVAL = snpg_version();
  OUTPUT:
    RETVAL

IV
create(device, seed)
    int device
    UV seed
  CODE:
    snpg_ctx *ctx = NULL;
    int rc = snpg_create(&ctx, device, (uint64_t)seed);
    if (rc != SNPG_OK) croak("\n### SNPGenieCUDA could not start (%d): %s\n", rc, snpg_last_error(NULL));
    RETVAL = PTR2IV(ctx);
  OUTPUT:
    RETVAL

IV
create_multi(n_gpus, seed)
    int n_gpus
    UV seed
  CODE:
    /* one ctx over GPUs 0 .. n_gpus - 1 of this process (codon ranges and bootstrap replicates split over them) */
    snpg_ctx *ctx = NULL;
    int rc = snpg_create_multi(&ctx, n_gpus, NULL, (uint64_t)seed);
    if (rc != SNPG_OK) croak("\n### SNPGenieCUDA could not start on %d GPUs (%d): %s\n", n_gpus, rc, snpg_last_error(NULL));
    RETVAL = PTR2IV(ctx);
  OUTPUT:
    RETVAL

int
team_size(h)
    IV h
  CODE:
    RETVAL = snpg_team_size(ctx_of(aTHX_ h));
  OUTPUT:
    RETVAL

void
destroy(h)
    IV h
  CODE:
    if (h) snpg_destroy(INT2PTR(snpg_ctx *, h));

void
within_job(h, group, ptr, n_seqs, aln_len, B, total_codons, n_products)
    IV h
    int group
    IV ptr
    UV n_seqs
    UV aln_len
    IV B
    IV total_codons
    int n_products
  PPCODE:
    /* the whole within-group job in ONE call (snpg_within_job_ex, host memory): load, per-codon engine, product sums and
     * bootstrap SEs.  Returns prod_sums [P][4], prod_se [P][6], poly, n_defined, comparisons, stats4, conserved over the
     * global codon axis (products concatenated in set_products order), all as packed strings. */
    snpg_ctx *ctx = ctx_of(aTHX_ h);
    STRLEN T = (STRLEN)total_codons, P = (STRLEN)n_products;
    SV *sums = out_buf(aTHX_ 32 * P), *se = out_buf(aTHX_ 48 * P), *poly = out_buf(aTHX_ T), *ndef = out_buf(aTHX_ 4 * T);
    SV *cmp = out_buf(aTHX_ 8 * T), *stats = out_buf(aTHX_ 32 * T), *cons = out_buf(aTHX_ T);
    snpg_job job;
    memset(&job, 0, sizeof job);
    job.struct_size = sizeof job;
    job.n_bootstraps = (int64_t)B;
    job.prod_sums = (double *)SvPVX(sums); job.prod_se = B >= 2 ? (double *)SvPVX(se) : NULL;
    job.poly = (uint8_t *)SvPVX(poly); job.n_defined = (uint32_t *)SvPVX(ndef); job.comparisons = (uint64_t *)SvPVX(cmp);
    job.stats4 = (double *)SvPVX(stats); job.conserved = (uint8_t *)SvPVX(cons);
    check(aTHX_ ctx, snpg_within_job_ex(ctx, group, INT2PTR(const uint8_t *, ptr), 0, n_seqs, aln_len, aln_len, &job), "within_job");
    EXTEND(SP, 7);
    PUSHs(sums); PUSHs(se); PUSHs(poly); PUSHs(ndef); PUSHs(cmp); PUSHs(stats); PUSHs(cons);

void
set_option(h, option, value)
    IV h
    int option
    IV value
  CODE:
    snpg_ctx *ctx = ctx_of(aTHX_ h);
    check(aTHX_ ctx, snpg_set_option(ctx, option, (int64_t)value), "set_option");

void
set_products(h, n_products, seg_ofs, seg_start, seg_stop, strand, aln_len)
    IV h
    int n_products
    SV *seg_ofs
    SV *seg_start
    SV *seg_stop
    SV *strand
    IV aln_len
  CODE:
    snpg_ctx *ctx = ctx_of(aTHX_ h);
    const int64_t *ofs = (const int64_t *)in_buf(aTHX_ seg_ofs, sizeof(int64_t) * (n_products + 1), "seg_ofs");
    STRLEN nseg = (STRLEN)ofs[n_products];
    check(aTHX_ ctx, snpg_set_products(ctx, n_products, ofs,
                                       (const int64_t *)in_buf(aTHX_ seg_start, sizeof(int64_t) * nseg, "seg_start"),
                                       (const int64_t *)in_buf(aTHX_ seg_stop, sizeof(int64_t) * nseg, "seg_stop"),
                                       (const int8_t *)in_buf(aTHX_ strand, n_products, "strand"), (int64_t)aln_len),
          "set_products");

IV
n_codons(h, product)
    IV h
    int product
  CODE:
    int64_t n = 0;
    snpg_ctx *ctx = ctx_of(aTHX_ h);
    check(aTHX_ ctx, snpg_n_codons(ctx, product, &n), "n_codons");
    RETVAL = (IV)n;
  OUTPUT:
    RETVAL

void
load_group(h, group, bases, n_seqs, aln_len)
    IV h
    int group
    SV *bases
    UV n_seqs
    UV aln_len
  CODE:
    snpg_ctx *ctx = ctx_of(aTHX_ h);
    const uint8_t *p = (const uint8_t *)in_buf(aTHX_ bases, (STRLEN)(n_seqs * aln_len), "bases");
    check(aTHX_ ctx, snpg_load_group(ctx, group, p, n_seqs, aln_len, aln_len), "load_group");

void
read_fasta(h, slot, path)
    IV h
    int slot
    const char *path
  PPCODE:
    snpg_ctx *ctx = ctx_of(aTHX_ h);
    const uint8_t *bases = NULL;
    uint64_t n = 0, len = 0;
    check(aTHX_ ctx, snpg_read_fasta(ctx, slot, path, &bases, &n, &len), "read_fasta");
    EXTEND(SP, 3);
    PUSHs(sv_2mortal(newSViv(PTR2IV(bases))));
    PUSHs(sv_2mortal(newSVuv((UV)n)));
    PUSHs(sv_2mortal(newSVuv((UV)len)));

void
load_group_ptr(h, group, ptr, n_seqs, aln_len)
    IV h
    int group
    IV ptr
    UV n_seqs
    UV aln_len
  CODE:
    snpg_ctx *ctx = ctx_of(aTHX_ h);
    check(aTHX_ ctx, snpg_load_group(ctx, group, INT2PTR(const uint8_t *, ptr), n_seqs, aln_len, aln_len), "load_group");

SV *
codon_strings(h, slot, product, codon, n_seqs)
    IV h
    int slot
    int product
    IV codon
    UV n_seqs
  CODE:
    snpg_ctx *ctx = ctx_of(aTHX_ h);
    SV *out = newSV(3 * (STRLEN)n_seqs + 1);
    SvPOK_only(out);
    SvCUR_set(out, 3 * (STRLEN)n_seqs);
    *SvEND(out) = '\0';
    int rc = snpg_codon_strings(ctx, slot, product, (int64_t)codon, (uint8_t *)SvPVX(out));
    if (rc != SNPG_OK) { SvREFCNT_dec(out); check(aTHX_ ctx, rc, "codon_strings"); }
    RETVAL = out;
  OUTPUT:
    RETVAL

void
hist(h, group, product)
    IV h
    int group
    int product
  PPCODE:
    snpg_ctx *ctx = ctx_of(aTHX_ h);
    int64_t C = 0;
    check(aTHX_ ctx, snpg_n_codons(ctx, product, &C), "n_codons");
    SV *hs = out_buf(aTHX_ sizeof(uint32_t) * SNPG_HIST_BINS * (STRLEN)C), *od = out_buf(aTHX_ (STRLEN)C);
    check(aTHX_ ctx, snpg_get_hist(ctx, group, product, (uint32_t *)SvPVX(hs), (uint8_t *)SvPVX(od)), "get_hist");
    EXTEND(SP, 2);
    PUSHs(hs);
    PUSHs(od);

SV *
site_counts(h, group, aln_len)
    IV h
    int group
    UV aln_len
  CODE:
    snpg_ctx *ctx = ctx_of(aTHX_ h);
    SV *out = newSV(16 * (STRLEN)aln_len + 1);
    SvPOK_only(out);
    SvCUR_set(out, 16 * (STRLEN)aln_len);
    *SvEND(out) = '\0';
    int rc = snpg_get_site_counts(ctx, group, (uint32_t *)SvPVX(out));
    if (rc != SNPG_OK) { SvREFCNT_dec(out); check(aTHX_ ctx, rc, "site_counts"); }
    RETVAL = out;
  OUTPUT:
    RETVAL

void
within(h, group, product)
    IV h
    int group
    int product
  PPCODE:
    snpg_ctx *ctx = ctx_of(aTHX_ h);
    int64_t C = 0;
    check(aTHX_ ctx, snpg_n_codons(ctx, product, &C), "n_codons");
    SV *poly = out_buf(aTHX_ (STRLEN)C), *ndef = out_buf(aTHX_ 4 * (STRLEN)C), *cmp = out_buf(aTHX_ 8 * (STRLEN)C),
       *st = out_buf(aTHX_ 32 * (STRLEN)C), *cons = out_buf(aTHX_ (STRLEN)C);
    check(aTHX_ ctx, snpg_within(ctx, group, product, (uint8_t *)SvPVX(poly), (uint32_t *)SvPVX(ndef), (uint64_t *)SvPVX(cmp),
                                 (double *)SvPVX(st), (uint8_t *)SvPVX(cons)), "within");
    EXTEND(SP, 5);
    PUSHs(poly); PUSHs(ndef); PUSHs(cmp); PUSHs(st); PUSHs(cons);

void
between(h, group_a, group_b, product)
    IV h
    int group_a
    int group_b
    int product
  PPCODE:
    snpg_ctx *ctx = ctx_of(aTHX_ h);
    int64_t C = 0;
    check(aTHX_ ctx, snpg_n_codons(ctx, product, &C), "n_codons");
    SV *poly = out_buf(aTHX_ (STRLEN)C), *nda = out_buf(aTHX_ 4 * (STRLEN)C), *ndb = out_buf(aTHX_ 4 * (STRLEN)C),
       *cmp = out_buf(aTHX_ 8 * (STRLEN)C), *st = out_buf(aTHX_ 32 * (STRLEN)C), *cons = out_buf(aTHX_ (STRLEN)C);
    check(aTHX_ ctx, snpg_between(ctx, group_a, group_b, product, (uint8_t *)SvPVX(poly), (uint32_t *)SvPVX(nda),
                                  (uint32_t *)SvPVX(ndb), (uint64_t *)SvPVX(cmp), (double *)SvPVX(st), (uint8_t *)SvPVX(cons)),
          "between");
    EXTEND(SP, 6);
    PUSHs(poly); PUSHs(nda); PUSHs(ndb); PUSHs(cmp); PUSHs(st); PUSHs(cons);

void
bootstrap_product(h, stats4, keep, C, B, want_sums)
    IV h
    SV *stats4
    SV *keep
    IV C
    IV B
    int want_sums
  PPCODE:
    snpg_ctx *ctx = ctx_of(aTHX_ h);
    const double *st = (const double *)in_buf(aTHX_ stats4, 32 * (STRLEN)C, "stats4");
    const uint8_t *kp = SvOK(keep) ? (const uint8_t *)in_buf(aTHX_ keep, (STRLEN)C, "keep") : NULL;
    SV *se = out_buf(aTHX_ 6 * sizeof(double)), *sums = out_buf(aTHX_ want_sums ? 32 * (STRLEN)B : 0);
    check(aTHX_ ctx, snpg_bootstrap_product(ctx, st, kp, (int64_t)C, (int64_t)B, want_sums ? (double *)SvPVX(sums) : NULL,
                                            (double *)SvPVX(se)), "bootstrap_product");
    EXTEND(SP, 2);
    PUSHs(se);
    PUSHs(sums);

void
windows(h, stats4, keep, C, W, step, B)
    IV h
    SV *stats4
    SV *keep
    IV C
    IV W
    IV step
    IV B
  PPCODE:
    snpg_ctx *ctx = ctx_of(aTHX_ h);
    const double *st = (const double *)in_buf(aTHX_ stats4, 32 * (STRLEN)C, "stats4");
    const uint8_t *kp = SvOK(keep) ? (const uint8_t *)in_buf(aTHX_ keep, (STRLEN)C, "keep") : NULL;
    int64_t nwin = 0;
    check(aTHX_ ctx, snpg_windows(ctx, st, kp, (int64_t)C, (int64_t)W, (int64_t)step, 0, NULL, NULL, &nwin), "windows");
    SV *sums = out_buf(aTHX_ 32 * (STRLEN)nwin), *se = out_buf(aTHX_ B >= 2 ? 24 * (STRLEN)nwin : 0);
    check(aTHX_ ctx, snpg_windows(ctx, st, kp, (int64_t)C, (int64_t)W, (int64_t)step, (int64_t)B, (double *)SvPVX(sums),
                                  B >= 2 ? (double *)SvPVX(se) : NULL, &nwin), "windows");
    EXTEND(SP, 3);
    PUSHs(sv_2mortal(newSViv((IV)nwin)));
    PUSHs(sums);
    PUSHs(se);

void
sliding_windows_r(h, stats4, n_defined, codon_num, C, numerator, denominator, jukes_cantor, W, step, B, min_defined)
    IV h
    SV *stats4
    SV *n_defined
    SV *codon_num
    IV C
    int numerator
    int denominator
    int jukes_cantor
    IV W
    IV step
    IV B
    IV min_defined
  PPCODE:
    /* SNPGenie_sliding_windows.R on one product's codon table: (number of windows, packed doubles [n][SNPG_RWIN_COLS]) */
    snpg_ctx *ctx = ctx_of(aTHX_ h);
    const double *st = (const double *)in_buf(aTHX_ stats4, 32 * (STRLEN)C, "stats4");
    const uint32_t *nd = SvOK(n_defined) ? (const uint32_t *)in_buf(aTHX_ n_defined, 4 * (STRLEN)C, "n_defined") : NULL;
    const int64_t *cn = SvOK(codon_num) ? (const int64_t *)in_buf(aTHX_ codon_num, 8 * (STRLEN)C, "codon_num") : NULL;
    snpg_rwindows opt;
    memset(&opt, 0, sizeof opt);
    opt.struct_size = sizeof opt; opt.numerator = numerator; opt.denominator = denominator; opt.jukes_cantor = jukes_cantor;
    opt.window = (int64_t)W; opt.step = (int64_t)step; opt.n_bootstraps = (int64_t)B; opt.min_defined = (int64_t)min_defined;
    int64_t nwin = 0;
    check(aTHX_ ctx, snpg_sliding_windows_r(ctx, st, nd, cn, (int64_t)C, &opt, &nwin, NULL, 0), "sliding_windows_r");
    SV *out = out_buf(aTHX_ sizeof(double) * SNPG_RWIN_COLS * (STRLEN)nwin);
    if (nwin) check(aTHX_ ctx, snpg_sliding_windows_r(ctx, st, nd, cn, (int64_t)C, &opt, &nwin, (double *)SvPVX(out), nwin), "sliding_windows_r");
    EXTEND(SP, 2);
    PUSHs(sv_2mortal(newSViv((IV)nwin)));
    PUSHs(out);

void
tables()
  PPCODE:
    SV *sites = out_buf(aTHX_ 128 * sizeof(double)), *diffs = out_buf(aTHX_ 8192 * sizeof(double));
    if (snpg_get_tables((double *)SvPVX(sites), (double *)SvPVX(diffs)) != SNPG_OK) croak("SNPGenieCUDA: get_tables failed");
    EXTEND(SP, 2);
    PUSHs(sites);
    PUSHs(diffs);

/*
 * snpgenie_cuda.h -- C ABI of libsnpgenie_cuda.so
 *
 * B200-native engine for SNPGenie's within-/between-group Nei-Gojobori dN/dS
 * path.  The reference (chasewnelson/SNPGenie) has no FFI seam: the engine is
 * inline in the Perl product loop.  Each entry point below names the reference
 * block it replaces ("wg" = snpgenie_within_group.pl, "bg" =
 * snpgenie_between_group.pl, "wgp"/"bgp" = the two *_processor.pl scripts).
 * INTEGRATION.md shows the XS binding a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ or torch types cross this boundary
 *   - every call returns SNPG_OK (0) or a negative SNPG_E_* code and never
 *     aborts, throws or exits; text via snpg_last_error()
 *   - the caller owns every buffer; inputs are copied to the device before the
 *     call returns, outputs are complete when it returns (calls synchronise)
 *   - one ctx = one CUDA device, used from one host thread at a time
 *   - do not fork() after snpg_create(): CUDA contexts do not survive fork, so
 *     the reference's fork-per-codon (wg:470-541) has no counterpart here
 *   - there is NO CPU fallback: without a usable device snpg_create() fails
 *
 * Codon codes (1 byte per sequence per codon): 0..63 = 16*b0+4*b1+b2 with
 * A,C,G,T = 0..3 after the reference's normalisation (uc, U->T; wg:231-232);
 * 64 = undefined (any N or '-'; wg:493,498); 65 = other (defined but not pure
 * ACGT, e.g. IUPAC -- misses both tables, contributes 0; wg:1715, :17990).
 */
#ifndef SNPGENIE_CUDA_H
#define SNPGENIE_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNPG_CODE_UNDEF 64
#define SNPG_CODE_OTHER 65
#define SNPG_HIST_BINS 66 /* 64 codons, undefined, other */

#define SNPG_OK 0
#define SNPG_E_ARG (-1)     /* bad argument / out of range */
#define SNPG_E_STATE (-2)   /* call order (e.g. no products set) */
#define SNPG_E_INPUT (-3)   /* input the reference would die() on (CDS not %3 ...) */
#define SNPG_E_CUDA (-4)    /* CUDA runtime failure, incl. "no device" */
#define SNPG_E_NOMEM (-5)

typedef struct snpg_ctx snpg_ctx;

/* ---- lifecycle --------------------------------------------------------- */
int snpg_create(snpg_ctx **out, int device, uint64_t seed);
void snpg_destroy(snpg_ctx *ctx);
const char *snpg_last_error(const snpg_ctx *ctx); /* ctx may be NULL: error of the last failed snpg_create */
const char *snpg_version(void);

/* ---- Nei-Gojobori tables (host only, no device needed) ------------------
 * Replaces get_number_of_sites (wg:1625-1721; positions summed 1+2+3 as at
 * wg:636-646) and return_avg_diffs (wg:1728-17996).
 * sites [64][2] = N_sites,S_sites; diffs [2][64][64] = N_diffs, S_diffs. */
int snpg_get_tables(double *sites, double *diffs);

/* ---- annotation ---------------------------------------------------------
 * Replaces get_product_coordinates (wg:1530-1621) + product extraction
 * (wg:286-329; bg:238-298) + codon split (wg:417-454; bg:351-388).
 * Products as CSR: segments seg_ofs[p]..seg_ofs[p+1], 1-based inclusive
 * coordinates on the '+' strand; strand[p] = +1/-1.  Segments are sorted by
 * start inside (wg:1608).  A '-' product is the reverse complement of the
 * concatenation (== fasta2revcom.pl + gtf2revcom.pl:319-320).
 * SNPG_E_INPUT if a product's length is not a multiple of 3 (wg:1596). */
int snpg_set_products(snpg_ctx *ctx, int n_products, const int64_t *seg_ofs,
                      const int64_t *seg_start, const int64_t *seg_stop,
                      const int8_t *strand, int64_t aln_len);
int snpg_n_products(const snpg_ctx *ctx);
int snpg_n_codons(const snpg_ctx *ctx, int product, int64_t *out);

/* Multi-GPU: this ctx owns the slice [first, first+count) of the global codon
 * axis (all products concatenated in set_products order).  rank/world as in
 * torch.distributed; default 0/1.  Call before snpg_set_products. */
int snpg_set_shard(snpg_ctx *ctx, int rank, int world);
/* The split itself (pure, no ctx, no device): rank r of w owns [first, first+count) of `total` units.
 * Used for codon ranges (snpg_set_shard) and for bootstrap replicate ranges. */
int snpg_shard_plan(int64_t total, int rank, int world, int64_t *first, int64_t *count);
int snpg_shard_range(const snpg_ctx *ctx, int64_t *first, int64_t *count, int64_t *total);

/* ---- alignment ingest: pack + per-codon histogram ------------------------
 * Replaces FASTA normalisation (wg:231-232) and, through the histogram, the
 * polymorphism pre-check (wg:469-541; bg:472-544) and the O(n^2) pair counting
 * (wg:596-609; bg:649-667).  bases: row-major, one row per sequence, raw FASTA
 * residues (any case, U allowed).  _host copies from host memory (pinned or
 * pageable) in row blocks overlapped with packing; _device takes a device
 * pointer.  Up to SNPG_MAX_GROUPS groups; reloading a group replaces it. */
#define SNPG_MAX_GROUPS 64
int snpg_load_group(snpg_ctx *ctx, int group, const uint8_t *bases, uint64_t n_seqs,
                    uint64_t aln_len, uint64_t row_stride);
int snpg_load_group_device(snpg_ctx *ctx, int group, const uint8_t *d_bases, uint64_t n_seqs,
                           uint64_t aln_len, uint64_t row_stride);
int snpg_drop_group(snpg_ctx *ctx, int group);
int snpg_group_size(const snpg_ctx *ctx, int group, uint64_t *n_seqs);

/* ---- FASTA ingest ("next" row f3) -----------------------------------------
 * Native reader with the reference's semantics (wg:204-253; bg:133-187): every line
 * that contains '>' starts a record, other lines are chomp'ed (LF only: a CR stays
 * in the sequence) and concatenated, all records must have one length
 * (SNPG_E_INPUT otherwise).  The residues land row-major in a PINNED host buffer
 * owned by the ctx (slot = group id; replaced by the next read into the same slot),
 * so snpg_load_group() on it streams at full PCIe rate.  No device needed for
 * the parse itself, but a ctx is (it owns the buffer). */
int snpg_read_fasta(snpg_ctx *ctx, int slot, const char *path, const uint8_t **bases, uint64_t *n_seqs, uint64_t *aln_len);
/* The reader works with every host thread (SNPG_INGEST_THREADS in the environment caps them): one pass finds the records
 * in the memory-mapped file while the pinned buffer is allocated, a second copies them into their rows.  With
 * SNPG_OPT_INGEST_ASYNC = 1 snpg_read_fasta returns while the second pass is still running in the background; library
 * calls that read the buffer (snpg_load_group, whole jobs, snpg_codon_strings) wait for the rows they need -- the
 * host-to-device copy of one row block then overlaps the parse of the next -- and snpg_fasta_wait makes the whole buffer
 * safe to read for the caller itself. */
int snpg_fasta_wait(snpg_ctx *ctx, int slot);
/* The three normalised characters of one codon for every sequence of a FASTA slot
 * (out [n_seqs][3]; uc, U->T, complemented on the '-' strand) -- what the host needs to
 * print the strings of 'other' alleles.  Product/codon as in snpg_within. */
int snpg_codon_strings(snpg_ctx *ctx, int slot, int product, int64_t codon, uint8_t *out);

/* Options: SNPG_OPT_KEEP_CODES (default 0).  With 0 the histograms are built straight
 * from the alignment bytes by the fused kernel and no code matrix exists.  With 1 the
 * packed codon matrix is materialised on the device (K1 then K2): needed by
 * snpg_get_codon_indices (a test hook); the between-group "first defined codon" rule
 * (bg:853-890) is served by SNPG_OPT_FIRST_DEFINED on the fast route.
 * Set before loading a group. */
#define SNPG_OPT_KEEP_CODES 1
/* SNPG_OPT_SITE_COUNTS (default 0): while a group is loaded also count A, C, G, T per
 * alignment column (snpg_get_site_counts) -- the whole row is then staged, not only the
 * CDS columns.  Set before snpg_set_products. */
#define SNPG_OPT_SITE_COUNTS 3
/* SNPG_OPT_PROFILE (default 0): bit mask over SNPG_K_* kinds (-1 = all).  Launches of a
 * selected kind are bracketed with CUDA events on the launch stream and their device
 * time is accumulated (snpg_get_profile).  Each pair of events costs a few
 * microseconds of stream time, so time whole jobs with a narrow mask. */
#define SNPG_OPT_PROFILE 2
/* SNPG_OPT_FUSED_KERNEL (default SNPG_FUSED_LDG_SPANS): which kernel builds the histograms when no code matrix is
 * kept.  All three give identical histograms.
 * LDG_SPANS: a lane owns a 16-byte column span and compares every row of it with three reference rows; the rows are read
 * with 128-bit global loads (the fastest measured on every diversity tried, see profiles/).
 * TMA_SPANS: the same arithmetic with the rows arriving through TMA tensor loads (512 bytes x 4 rows per stage) into a
 * per-warp shared-memory pipeline, so that the bytes in flight are bounded by shared memory, not by registers.  Falls
 * back to LDG_SPANS when a tensor map cannot describe the buffer.
 * TMA_TILES: persistent CTAs, frame-aligned 208-byte x 128-row TMA tensor loads, one lane per group of four codons;
 * falls back to LDG_SPANS like TMA_SPANS. */
#define SNPG_OPT_FUSED_KERNEL 4
#define SNPG_FUSED_TMA_TILES 0
#define SNPG_FUSED_LDG_SPANS 1
#define SNPG_FUSED_TMA_SPANS 2
/* SNPG_OPT_BOOTSTRAP_KERNEL (default SNPG_BOOT_CLASSES): how the whole-product bootstrap (wg:854-916) sums a
 * replicate.  CLASSES: a replicate is C independent uniform draws, so the numbers of draws per slot are multinomial
 * and the draws that land on codons with identical rows only count through their total.  The conserved codons of a
 * product (N_diffs = S_diffs = 0) fall into a handful of classes of identical (N_sites, S_sites); a chain of exact
 * binomial variates (inversion below a mean of 10, Hoermann's BTRS rejection sampler above) gives every replicate its
 * number of draws per class and the number of draws on the remaining codons, and only those are drawn one by one, as
 * WEIGHTS does.  The same distribution of replicate sums as the other two forms from a different use of the Philox
 * stream; on a typical alignment (most codons conserved) a fraction of the draws.  Sliding windows and tables beyond
 * shared memory take WEIGHTS / GATHER.  WEIGHTS: every draw bumps a 1-byte counter of its slot in shared memory -- a warp turns its
 * replicate into a row of the multinomial weight matrix -- and the block contracts its rows with the product's
 * (N_sites, S_sites, N_diffs, S_diffs) table as FP64 tensor-core tiles.  GATHER: every draw gathers its codon's
 * row from a shared-memory table and adds it.  Same Philox stream draw for draw: the replicate sums of the two
 * agree to rounding (tested).  Tables too large for shared memory use GATHER from global memory either way. */
#define SNPG_OPT_BOOTSTRAP_KERNEL 5
#define SNPG_BOOT_GATHER 0
#define SNPG_BOOT_WEIGHTS 1
#define SNPG_BOOT_CLASSES 2
/* SNPG_OPT_BOOT_SLOTS (default SNPG_SLOTS_FAITHFUL): the whole-product bootstrap of the reference draws
 * int(rand(C+1)) (wg:885; bgp:919), i.e. C+1 slots of which slot 0 is a codon that does not exist and adds
 * nothing -- every replicate is on average C/(C+1) of a product.  FAITHFUL reproduces that; CORRECTED draws from
 * the C real codons.  (Window bootstraps always draw from the window's W codons, wgp:406.) */
#define SNPG_OPT_BOOT_SLOTS 6
#define SNPG_SLOTS_CORRECTED 0
#define SNPG_SLOTS_FAITHFUL 1
/* SNPG_OPT_INGEST_ASYNC (default 0): see snpg_read_fasta. */
#define SNPG_OPT_INGEST_ASYNC 7
/* SNPG_OPT_FIRST_DEFINED (default 0): while a group is loaded WITHOUT a code matrix, also note the first defined codon
 * of every codon column in file order -- what the between-group "first defined codon" rule (bg:853-890) takes when the
 * other group has no defined codon at a conserved position -- so that between-group analyses keep the fast histogram
 * route (without it that rare case takes the lowest allele code).  Set before loading a group. */
#define SNPG_OPT_FIRST_DEFINED 8
int snpg_set_option(snpg_ctx *ctx, int option, int64_t value);

/* ---- test hooks for the bit-exact checks ---------------------------------
 * codes out: [n_seqs][C] row-major (sequence-major, like the reference's
 * $products_seqs_hh); hist out: [C][SNPG_HIST_BINS]; other_distinct out: [C],
 * 1 if >= 2 distinct 'other' codon strings occur (affects only the
 * polymorphic flag).  Only the codons of this ctx's shard are filled when
 * world > 1 (C = shard slice of the product). */
int snpg_get_codon_indices(snpg_ctx *ctx, int group, int product, uint8_t *out);
int snpg_get_hist(snpg_ctx *ctx, int group, int product, uint32_t *out, uint8_t *other_distinct);

/* Per-site nucleotide counts out [aln_len][4] = A, C, G, T after uc and U->T; everything else is
 * not counted (replaces the site pass wg:1203-1237; pi and majority follow from the counts). */
int snpg_get_site_counts(snpg_ctx *ctx, int group, uint32_t *out);

/* ---- per-codon engine ------------------------------------------------------
 * Replaces wg:586-751 (within) and bg:573-955 (between): per codon the
 * variability flag, number of defined sequences, number of comparisons, mean
 * N/S sites and mean N/S differences, and for conserved codons the code of the
 * conserved codon (SNPG_CODE_UNDEF when none is defined: wg prints an empty
 * codon, bg prints '---').  All outputs have length C (product's codons in this
 * shard); any output pointer may be NULL.  stats4 [C][4] = N_sites, S_sites,
 * N_diffs, S_diffs (the layout the resampling calls take). */
int snpg_within(snpg_ctx *ctx, int group, int product, uint8_t *poly, uint32_t *n_defined,
                uint64_t *comparisons, double *stats4, uint8_t *conserved_code);
int snpg_between(snpg_ctx *ctx, int group_a, int group_b, int product, uint8_t *poly,
                 uint32_t *n_defined_a, uint32_t *n_defined_b, uint64_t *comparisons,
                 double *stats4, uint8_t *conserved_code);

/* ---- resampling -------------------------------------------------------------
 * snpg_bootstrap_product replaces wg:831-978 and bgp:869-1102: B replicates of C
 * draws from C+1 slots (slot 0 = the reference's nonexistent codon 0, wg:885).
 * keep (or NULL) is the min-taxa mask of bgp:412-467 (0 = codon contributes 0).
 * sums_out [B][4] or NULL (the <product>_bootstrap_results.txt rows).
 * se_out [6] = sample SD (n-1, two-pass; wg:1351-1365) over replicates of dN,
 * dS, dN-dS and their Jukes-Cantor forms (bgp:1047-1071), with the reference's
 * '*' -> 0 coercion.  Streams are Philox4x32-10 keyed by the ctx seed; results
 * agree with the reference statistically, not bitwise. */
int snpg_bootstrap_product(snpg_ctx *ctx, const double *stats4, const uint8_t *keep, int64_t C,
                           int64_t B, double *sums_out, double *se_out);

/* snpg_windows replaces wgp:293-465 and bgp:501-736.  Window k covers codons
 * [k*step, k*step+W) for k*step <= C-W.  win_sums [nwin][4] accumulate in
 * ascending codon order (bitwise the reference's order); win_se [nwin][3] = SD
 * of dN, dS, dN-dS over B replicates of W uniform draws inside the window
 * (NULL or B < 2: skipped).  *nwin_out receives the number of windows. */
int snpg_windows(snpg_ctx *ctx, const double *stats4, const uint8_t *keep, int64_t C, int64_t W,
                 int64_t step, int64_t B, double *win_sums, double *win_se, int64_t *nwin_out);

/* ---- SNPGenie_sliding_windows.R ("next" row f2) ------------------------------
 * The sliding-window statistics of the R script the reference's README recommends over the Perl window processor
 * (README.md:281), for ONE product's (or one group pair's) per-codon table: SNPGenie_sliding_windows.R:441-548.
 * stats4 [C][4] = N_sites, S_sites, N_diffs, S_diffs per codon (snpg_within / snpg_between / a codon results file);
 * n_defined [C] or NULL = num_defined_seqs (between-group: the smaller of the two groups' counts, R :187-190) -- codons
 * below opt->min_defined are dropped first (R :441); codon_num [C] or NULL (1..C in table order).
 * Windows: rows i .. i + window - 1 of the kept, codon_num-sorted table for i = min(codon_num), + step, ... (R :444-449:
 * the script indexes ROWS by codon numbers); they end with the last complete window.  Call with out = NULL for the
 * number of windows.  out [n_windows][SNPG_RWIN_COLS], in the script's column order (sw_start ... sw_ASL_dNdS_P);
 * undefined values are NaN / Inf as R computes them (its NA prints as NaN here).  Each of the script's four boot()
 * calls draws its own resamples; here dN, dS, dN/dS and dN - dS come from one set of replicates per window. */
typedef struct snpg_rwindows {
    uint32_t struct_size;         /* sizeof(snpg_rwindows) */
    int32_t numerator;            /* 0 = N (the script's "N"); 1 = S */
    int32_t denominator;          /* 1 = S (the script's "S"); 0 = N */
    int32_t jukes_cantor;         /* CORRECTION = "JC" (R :326-361) */
    int64_t window, step;         /* WINDOW_SIZE >= 2, STEP_SIZE >= 1 */
    int64_t n_bootstraps;         /* NBOOTSTRAPS, 2 .. 1,000,000 */
    int64_t min_defined;          /* MIN_DEFINED_CODONS (ignored when n_defined is NULL) */
} snpg_rwindows;
#define SNPG_RWIN_COLS 24
/* columns: 0 sw_start, 1 sw_center, 2 sw_end, 3 sw_num_replicates, 4 sw_N_diffs, 5 sw_S_diffs, 6 sw_N_sites, 7 sw_S_sites,
 * 8 sw_dN, 9 sw_dS, 10 sw_dNdS, 11 sw_dN_m_dS, 12 sw_boot_dN_SE, 13 sw_boot_dS_SE, 14 sw_boot_dN_over_dS_SE,
 * 15 sw_boot_dN_over_dS_P, 16 sw_boot_dN_m_dS_SE, 17 sw_boot_dN_m_dS_P, 18 sw_boot_dN_gt_dS_count, 19 .._eq_.., 20 .._lt_..,
 * 21 sw_ASL_dN_gt_dS_P, 22 sw_ASL_dN_lt_dS_P, 23 sw_ASL_dNdS_P */
int snpg_sliding_windows_r(snpg_ctx *ctx, const double *stats4, const uint32_t *n_defined, const int64_t *codon_num, int64_t C,
                           const snpg_rwindows *opt, int64_t *n_windows, double *out, int64_t out_cap_windows);

/* ---- pooled-sample codon contraction ("next" row f4) ---------------------------
 * snpgenie.pl's pooled mode (:6631-6760) knows, per codon position, the FREQUENCIES of the codons seen in the pool and
 * the mean coverage; its mean numbers of differences between two reads are the same table contraction as the sequence
 * engine's, on frequencies: sum over i < j of (f_i cov)(f_j cov) T[i][j] / ((cov^2 - cov) / 2).
 * freq [C][64] (codon code order, 0 where a codon is absent), coverage [C] -> n_diffs [C], s_diffs [C].
 * SNPG_E_INPUT where the script dies: several codons at a position whose coverage leaves no pair of reads. */
int snpg_pooled_diffs(snpg_ctx *ctx, const double *freq, const double *coverage, int64_t C, double *n_diffs, double *s_diffs);

/* ---- whole-job entry (what bench.py times) ---------------------------------
 * For every product of a loaded group, entirely on the device: per-codon stats
 * (as snpg_within), product sums (wg:690-693) and bootstrap SEs (wg:831-978).
 * prod_sums [P][4]; prod_se [P][6] (NULL or B < 2: no bootstrap).
 * Under a shard (world > 1) this covers the local codon slice only and callers
 * combine slices (see snpg_stats_device / bench.py). */
int snpg_within_all(snpg_ctx *ctx, int group, int64_t B, double *prod_sums, double *prod_se);

/* Between-group analysis of every pair of the listed groups and every product in ONE call with a single
 * synchronisation: the device-resident counterpart of the between driver's group-pair loop
 * (snpgenie_between_group.pl:410-423, per-codon values :680-947) followed by the between processor's
 * filtered product sums (snpgenie_between_group_processor.pl:412-467) and whole-product bootstrap
 * (:869-1102).  Pairs are ordered (g[0],g[1]), (g[0],g[2]), ..., (g[n-2],g[n-1]).  A codon is kept when
 * n_def_a + n_def_b >= min_taxa_total (if > 0) and min(n_def_a, n_def_b) >= min_taxa_group (if > 0); any other
 * codon contributes nothing to the sums or to a bootstrap draw.
 * pair_sums [n_pairs][P][4] (N_sites, S_sites, N_diffs, S_diffs); pair_se [n_pairs][P][6] (SE of dN, dS,
 * dN-dS and their Jukes-Cantor forms; may be NULL, and is not computed for B < 2). */
int snpg_between_all(snpg_ctx *ctx, const int *groups, int n_groups, int64_t min_taxa_total, int64_t min_taxa_group, int64_t B,
                     double *pair_sums, double *pair_se);

/* The between-group analysis as ONE call with a single synchronisation: loads n_groups alignments (all in host memory
 * or all in device memory; arguments as snpg_load_group / snpg_load_group_device) as groups 0 .. n_groups-1 and runs
 * snpg_between_all on them -- the between driver's outer loop over the group files (snpgenie_between_group.pl:133-187,
 * :410-423) and the processor's sums and bootstraps (snpgenie_between_group_processor.pl:412-467, :869-1102).  The
 * loads' kernels stay queued behind each other and the pair chains run on parallel streams; results equal the separate
 * calls bit for bit.  Device buffers must stay valid until the call returns. */
int snpg_between_job(snpg_ctx *ctx, int n_groups, const uint8_t *const *bases, int bases_on_device, const uint64_t *n_seqs, uint64_t aln_len,
                     const uint64_t *row_stride, int64_t min_taxa_total, int64_t min_taxa_group, int64_t B, double *pair_sums, double *pair_se);

/* Load + snpg_within_all as ONE call with a single synchronisation at the end: the
 * whole within-group job for a product-level answer (what the within driver's product
 * table needs).  bases is a host pointer (bases_on_device = 0; must stay valid until
 * the call returns) or a device pointer (1). */
int snpg_within_job(snpg_ctx *ctx, int group, const uint8_t *bases, int bases_on_device, uint64_t n_seqs,
                    uint64_t aln_len, uint64_t row_stride, int64_t B, double *prod_sums, double *prod_se);

/* The whole job with everything the drivers' tables need, still ONE call and one synchronisation: load, per-codon
 * engine (wg:586-751), product sums and whole-product bootstrap (wg:690-693, :831-978), sliding windows with their
 * bootstraps (wgp:293-465).  Outputs are host pointers, any may be NULL; pinned memory (cudaHostAlloc/cudaHostRegister)
 * is written by asynchronous copies that overlap the bootstrap, other memory goes through a staging buffer.
 * Per-codon outputs cover the GLOBAL codon axis: all products concatenated in snpg_set_products order (product p at
 * sum of snpg_n_codons of the products before it), `total` of snpg_shard_range entries -- on several GPUs as well
 * (every rank receives every codon range through the exchange).  Window outputs cover the windows of all products in
 * product order: product p owns windows [win_ofs[p], win_ofs[p+1]) of snpg_window_plan; a product shorter than the
 * window has none (wgp:299-305). */
typedef struct snpg_job {
    uint32_t struct_size;        /* sizeof(snpg_job) */
    uint32_t reserved;
    int64_t n_bootstraps;        /* whole-product bootstrap replicates; < 2: none */
    int64_t window, step;        /* sliding windows of `window` codons every `step` codons; window = 0: none */
    int64_t n_window_bootstraps; /* < 2: window sums only */
    double *prod_sums;           /* [P][4] N_sites, S_sites, N_diffs, S_diffs */
    double *prod_se;             /* [P][6] SE of dN, dS, dN-dS and their Jukes-Cantor forms */
    uint32_t *prod_undefined;    /* [P][6] replicates whose value is undefined (Jukes-Cantor of d >= 3/4; the reference dies, bgp:1047) */
    uint8_t *poly;               /* [total] 1 = polymorphic */
    uint32_t *n_defined;         /* [total] */
    uint64_t *comparisons;       /* [total] */
    double *stats4;              /* [total][4] */
    uint8_t *conserved;          /* [total] code of the conserved codon (SNPG_CODE_UNDEF: polymorphic or none defined) */
    double *win_sums;            /* [n_windows][4] */
    double *win_se;              /* [n_windows][3] SE of dN, dS, dN-dS over n_window_bootstraps replicates */
} snpg_job;
int snpg_window_plan(snpg_ctx *ctx, int64_t window, int64_t step, int64_t *win_ofs /*[P+1]*/);
int snpg_within_job_ex(snpg_ctx *ctx, int group, const uint8_t *bases, int bases_on_device, uint64_t n_seqs, uint64_t aln_len,
                       uint64_t row_stride, const snpg_job *job);

/* ---- several GPUs -----------------------------------------------------------------------------------------
 * The path shards by codon range (histograms, per-codon statistics) and by bootstrap replicate (SURVEY.md 8e;
 * north_star: "sharded across the 8 B200s by product/codon range and by bootstrap replicate").  The exchange is the
 * reference's gather of its forked children's results (wg:508-538; bg:949-1017): here every rank's kernels store
 * their codon range / replicate range straight into every other rank's memory over NVLink (peer access inside a
 * process, CUDA IPC across processes) and signal arrival with flags the consuming kernels wait on -- no host
 * round trip, no separate collective launch.  Whole jobs (snpg_within_job, snpg_within_job_ex) then return the same
 * results on every rank as a single GPU would, bit for bit.
 *
 * One process, N GPUs (what the Perl command lines use with --gpus N): snpg_create_multi returns ONE ctx that fans
 * every call out to a child ctx per GPU.  Supported on it: snpg_set_option, snpg_set_products, snpg_n_products,
 * snpg_n_codons, snpg_shard_range, snpg_read_fasta, snpg_codon_strings, snpg_load_group (host memory),
 * snpg_drop_group, snpg_group_size, snpg_get_hist, snpg_get_site_counts (not supported: E_STATE), snpg_within,
 * snpg_within_all, snpg_within_job / _ex (host memory), snpg_window_plan, snpg_launch_count, snpg_last_error,
 * snpg_destroy; snpg_bootstrap_product, snpg_windows and snpg_test_philox run on the first GPU.  Anything else
 * returns SNPG_E_STATE.  devices = NULL means GPUs 0..n_gpus-1. */
int snpg_create_multi(snpg_ctx **out, int n_gpus, const int *devices, uint64_t seed);
int snpg_team_size(const snpg_ctx *ctx);                 /* 1 for an ordinary ctx */
/* One process per GPU (torchrun, MPI ...): every rank calls snpg_set_shard, snpg_set_products, then
 * snpg_peer_export; the host program gathers the handles of all ranks in rank order (any transport: a file,
 * torch.distributed, MPI) and hands the array to snpg_peer_connect on every rank.  max_bootstraps bounds
 * n_bootstraps of later jobs.  Ranks of one process are connected by pointer, ranks of other processes through
 * CUDA IPC.  snpg_set_products afterwards disconnects (export and connect again). */
#define SNPG_PEER_HANDLE_BYTES 128
int snpg_peer_export(snpg_ctx *ctx, int64_t max_bootstraps, void *handle);
int snpg_peer_connect(snpg_ctx *ctx, const void *handles, int n_handles);
int snpg_peer_connected(const snpg_ctx *ctx);            /* 1 when whole jobs exchange with the peers */

/* Device-pointer variants for multi-GPU plumbing done by the CALLER (e.g. NCCL through torch.distributed -- the
 * baseline the peer-memory exchange above replaces): d_stats4 is a DEVICE buffer [count][4] for this shard's codons. */
int snpg_within_stats_device(snpg_ctx *ctx, int group, double *d_stats4);
/* Bootstrap all products from a full-length DEVICE stats buffer [total][4];
 * replicates [b_first, b_first+b_count) only; d_rep_out DEVICE [P][b_count][4]. */
int snpg_bootstrap_all_device(snpg_ctx *ctx, const double *d_stats4_full, int64_t b_first,
                              int64_t b_count, double *d_rep_out);

/* Product sums (wg:690-693) from a full-length DEVICE stats buffer -> host prod_sums [P][4]. */
int snpg_product_sums_device(snpg_ctx *ctx, const double *d_stats4_full, double *prod_sums);
/* SD over replicates from a DEVICE buffer rep [n_items][B][4] -> host se_out [n_items][6]. */
int snpg_rep_moments_device(snpg_ctx *ctx, const double *d_rep, int64_t n_items, int64_t B, double *se_out);

/* Launch all kernels on a caller-owned CUDA stream (a cudaStream_t, e.g. torch's
 * current stream) instead of the ctx's own; NULL restores an internal stream.  The legacy default stream has the
 * handle 0 in most bindings (torch.cuda.default_stream().cuda_stream == 0): pass cudaStreamLegacy, (void *)1, for it --
 * NULL does NOT mean "the default stream" here. */
int snpg_set_stream(snpg_ctx *ctx, void *cuda_stream);

/* Per-kernel device time measured with CUDA events (needs SNPG_OPT_PROFILE=1). */
#define SNPG_K_PACK 0
#define SNPG_K_HIST 1
#define SNPG_K_STATS 2
#define SNPG_K_SUMS 3
#define SNPG_K_BOOT 4
#define SNPG_K_MOMENTS 5
#define SNPG_K_WINDOW 6
#define SNPG_K_OTHER 7
#define SNPG_K_FUSED 8   /* pack + histogram in one pass (no code matrix kept) */
#define SNPG_K_SITES 9   /* per-site A/C/G/T counts */
#define SNPG_N_KERNEL_KINDS 10
int snpg_get_profile(snpg_ctx *ctx, double *ms /*[10]*/, int64_t *calls /*[10]*/, int reset);

/* Test hook: one Philox4x32-10 block computed by the device code the bootstraps use
 * (checked against the Random123 known-answer vectors). */
int snpg_test_philox(snpg_ctx *ctx, const uint32_t *ctr4, const uint32_t *key2, uint32_t *out4);

/* The columns of every row a load stages from a host buffer (this ctx's codon range: from its first CDS column
 * rounded down to 16 to its last): what bench.py counts as h2d bytes. */
int snpg_staged_columns(const snpg_ctx *ctx, int64_t *first_col, int64_t *n_cols);
/* Peaks the resampling kernel is measured against, probed on this GPU: out[0] = Philox4x32-10 words generated per
 * second by a kernel that does nothing else (the issue bound of a sampler that spends one word per draw), out[1] =
 * FP64 tensor TFLOP/s of mma.sync.m8n8k4.f64; out[2..3] reserved. */
int snpg_probe_peaks(snpg_ctx *ctx, double *out /*[4]*/);

/* Kernel launches issued by this ctx since creation (bench.py's gpu_launches). */
int64_t snpg_launch_count(const snpg_ctx *ctx);
/* Of those, launches of the TMA-tiled histogram kernel (0 means every fused histogram went
 * through the vector-load kernel: option, or a buffer a tensor map cannot describe). */
int64_t snpg_tile_launch_count(const snpg_ctx *ctx);
/* The same for the TMA-fed span kernel (SNPG_FUSED_TMA_SPANS). */
int64_t snpg_span_tma_launch_count(const snpg_ctx *ctx);
/* snpg_within_job on a device-resident alignment is captured into a CUDA graph on the second identical call
 * in a row and replayed from then on (same kernels, same results, fewer launch gaps; SNPG_NO_GRAPH=1 in the
 * environment or SNPG_OPT_PROFILE != 0 turns this off).  Number of replays so far: */
int64_t snpg_graph_replays(const snpg_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* SNPGENIE_CUDA_H */

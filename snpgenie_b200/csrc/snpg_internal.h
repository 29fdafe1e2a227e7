// Internal declarations shared by the kernels and the C-ABI layer.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace snpg {

constexpr int kHistBins = 66;      // 64 codons, undefined, other
constexpr int kCodeUndef = 64;
constexpr int kCodeOther = 65;
constexpr int kCodePad = 127;      // fills rows >= n_seqs up to the padded row count; never counted
constexpr int kRowPad = 128;       // packed matrix row count is padded to a multiple of this
constexpr uint32_t kOtherNone = 0xFFFFFFFFu;

// One entry of the codon index map: the alignment columns of the codon's three
// nucleotides and the strand (minus = 1 for '-': complement while classifying).
// Built on the host from the CDS segments (frame, multi-segment, strand).
struct CodonMap { int32_t p0, p1, p2, minus; };

// Fused path: the codons that START inside one 16-byte column span, as up to
// kSpanRuns runs of consecutive codons of one product (start offsets off,
// off+3, ...; codon index codon0 + dir*j).  cnt == 0 ends the list.
constexpr int kSpanRuns = 4;
struct SpanRun { int32_t codon0; int8_t off, cnt, dir, minus; };

// Fused path, TMA-tiled variant (tile_hist.cu): up to 64 consecutive codons of one product that occupy
// 3*ncodons contiguous alignment bytes starting at byte column x0 (relative to the staged block's first
// column; any alignment).  The codon in the j-th 3-byte slot (ascending address) has index codon0 + dir*j.
constexpr int kTileRows = 128;     // rows per TMA box
struct CodonTile { int32_t x0, codon0; int16_t ncodons; int8_t dir, minus; int32_t pad; };

// Device view of one loaded group, offset to the first codon a kernel works on.
struct GroupView {
    const uint32_t *hist;        // [codons][66]
    const uint32_t *other_min;   // [codons] min packed chars over 'other' codons (kOtherNone if none)
    const uint32_t *other_max;   // [codons] max ...
    const uint8_t *codes;        // [codons][n_pad] or null
    const uint8_t *first_def;    // [codons] code of the first defined codon in row order (SNPG_OPT_FIRST_DEFINED) or null
    int64_t n, n_pad;
};

// ---- multi-GPU exchange over peer memory (snpg_peer.cuh, exchange.cu) --------------------------------
constexpr int kMaxPeers = 8;       // ranks of one NVSwitch domain
constexpr int kStageStats = 0;     // per-codon statistics of every codon range have arrived
constexpr int kStageVals = 1;      // per-replicate values of every replicate range have arrived
// First bytes of every rank's exchange region.  flag[stage][src] = epoch of the last job whose stage-`stage` data
// from rank `src` is complete in THIS region; done[] = blocks of the running producer kernel that have finished.
struct XHeader { uint32_t flag[2][kMaxPeers]; uint32_t done[2]; uint32_t err; uint32_t pad[13]; };
static_assert(sizeof(XHeader) == 128, "exchange header is one 128-byte line");
// world <= 1: no exchange, every peer_* call is a no-op
struct PeerSync { int world = 1, rank = 0; XHeader *hdr[kMaxPeers] = {}; const uint32_t *epoch = nullptr; };
// Per-codon outputs replicated into every rank's region, indexed on the GLOBAL codon axis (first = this shard's
// first codon there).
struct PeerOut {
    int world = 1; int64_t first = 0;
    double *stats4[kMaxPeers] = {}; unsigned long long *cmp[kMaxPeers] = {}; uint32_t *ndef[kMaxPeers] = {};
    uint8_t *poly[kMaxPeers] = {}, *cons[kMaxPeers] = {};
};
// Per-replicate values vals[which][item][B] replicated into every rank's region (global replicate index).
struct PeerVals { int world = 1; double *vals[kMaxPeers] = {}; };

// host (ng_tables.cu)
void build_tables(double *sites, double *diffs);

// ---- kernel launch wrappers; all asynchronous on `stream` -------------------

// K1 pack: raw alignment rows -> 1-byte codon codes, TRANSPOSED: codes[c][row],
// row pitch n_pad (multiple of kRowPad; rows >= n hold kCodePad).  row0 (multiple
// of kRowPad) is the position of d_aln's first row inside the group.
void launch_pack(const uint8_t *d_aln, int64_t row_stride, int64_t n_rows, int64_t row0, const CodonMap *d_map,
                 int64_t n_codons, uint8_t *d_codes, int64_t n_pad, uint32_t *d_other_min, uint32_t *d_other_max,
                 cudaStream_t stream);

// The code of the first row (in file order) whose codon is defined, per codon: what the between-group "first defined
// codon" rule (bg:853-890) needs when no code matrix is kept.  d_first [n_codons], 0xFF = not found yet (set by the
// caller before the first row block); every row block of a group goes through this in order.
void launch_first_defined(const uint8_t *d_aln, int64_t row_stride, int64_t n_rows, const CodonMap *d_map, int64_t n_codons,
                          uint8_t *d_first, cudaStream_t stream);

// K2 histogram: codes[c][0..n_pad) -> hist[c][66] (u32, zeroed by the caller)
void launch_hist(const uint8_t *d_codes, int64_t n_pad, int64_t n_codons, uint32_t *d_hist, cudaStream_t stream);

// Reference rows of the compare-with-reference kernels from the group's first n_sample (<= 32) rows:
// d_cons[column] (n_spans*16 + 4 bytes written) and d_alt (see below).  Every row needs
// n_spans*16 + 4 readable bytes.
void launch_vote_refs(const uint8_t *d_aln, int64_t pitch, int n_sample, int64_t n_spans, uint8_t *d_cons, uint32_t *d_alt,
                      uint32_t *d_zero_hist, uint32_t *d_other_min, uint32_t *d_other_max, int64_t n_codons, uint32_t *d_sched,
                      cudaStream_t stream);
// d_alt: the spans' second and third reference patterns and the counts of the rows that equalled them, as planes over
// the span axis (RefPlanes, snpg_span.cuh): alt_words(n_spans) u32, written by launch_vote_refs, counts raised by the
// histogram kernels, spent by the fix-up.  d_sched: four zeroed u32 (chunk counter, finished blocks, the vote's count of
// sampled (span, row) pairs that match no reference -- launch_vote_refs adds to it, or null --, spare); the kernel leaves
// them zeroed.  Every row needs n_spans*16 + 16 readable bytes.
__host__ __device__ constexpr int64_t alt_stride(int64_t n_spans) { return (n_spans + 1 + 3) & ~(int64_t)3; }
constexpr int64_t alt_words(int64_t n_spans) { return 11 * alt_stride(n_spans); }
void launch_fused_hist(const uint8_t *d_aln, int64_t pitch, int64_t n_rows, const uint8_t *d_cons, uint32_t *d_alt, int64_t n_spans,
                       const SpanRun *d_runs, uint32_t *d_hist, uint32_t *d_other_min, uint32_t *d_other_max, uint32_t *d_sched,
                       int n_sms, cudaStream_t stream);
void launch_fused_fixup(const uint8_t *d_cons, int64_t cons_shift, uint32_t *d_alt, int64_t n_spans, const CodonMap *d_map, const uint8_t *d_regular,
                        int64_t n_codons, uint32_t n_rows, uint32_t *d_hist, uint32_t *d_other_min, uint32_t *d_other_max,
                        cudaStream_t stream);
// TMA-tiled variant of launch_fused_hist (same contract: counts of the non-consensus codons of the regular
// codons are added to hist).  make_tile_tensor_map fills a 128-byte CUtensorMap (64-byte aligned storage)
// for a block of rows; it returns false when the geometry cannot be expressed (base or pitch not 16-byte
// aligned) or the driver lacks the encoder -- the caller then takes launch_fused_hist.
bool make_tile_tensor_map(void *out_map, const uint8_t *d_aln, int64_t pitch, int64_t rows, int64_t row_bytes, int l2_promotion);
// d_sched: two zeroed u32 (chunk counter, finished CTAs) owned by the ctx; the kernel leaves them zeroed.
cudaError_t launch_tile_hist(const void *tensor_map, int64_t n_rows, const uint8_t *d_cons, const CodonTile *d_tiles, int64_t n_tiles,
                             int n_sms, uint32_t *d_sched, uint32_t *d_hist, uint32_t *d_other_min, uint32_t *d_other_max,
                             cudaStream_t stream);
// TMA-fed span kernel (span_tma.cu; an option, measured slower than k_fused_hist: see profiles/): k_fused_hist's arithmetic with the rows staged in shared memory by
// cp.async.bulk.tensor loads.  Same contract as launch_fused_hist.  make_span_tensor_map as make_tile_tensor_map (words of
// 4 bytes, boxes of 512 bytes x span_stage_rows() rows, zero fill outside).  launch_span_hist returns
// cudaErrorInvalidValue when the chunk index would not fit 32 bits (the caller then takes launch_fused_hist).
int span_stage_rows();
bool make_span_tensor_map(void *out_map, const uint8_t *d_aln, int64_t pitch, int64_t rows, int64_t row_bytes, int l2_promotion);
cudaError_t launch_span_hist(const void *tensor_map, int64_t n_rows, const uint8_t *d_cons, uint32_t *d_alt, int64_t n_spans,
                             const SpanRun *d_runs, uint32_t *d_hist, uint32_t *d_other_min, uint32_t *d_other_max, uint32_t *d_sched,
                             int n_sms, cudaStream_t stream);
void launch_scatter_irregular(const uint32_t *d_tmp_hist, const uint32_t *d_tmp_min, const uint32_t *d_tmp_max,
                              const int32_t *d_idx, int64_t n_irr, uint32_t *d_hist, uint32_t *d_other_min,
                              uint32_t *d_other_max, cudaStream_t stream);

// Per-site A/C/G/T counts: sites[col][5] (A, C, G, T, other-but-recorded), zeroed before the
// first row block; d_aln points at column 0 (16-byte aligned, pitch covers 16*ceil(n_cols/16));
// d_cons = the group's first row.  launch_site_fixup once after the last block.
void launch_site_hist(const uint8_t *d_aln, int64_t pitch, int64_t n_rows, const uint8_t *d_cons, int64_t n_cols, uint32_t *d_sites,
                      cudaStream_t stream);
void launch_site_fixup(const uint8_t *d_cons, int64_t n_cols, uint32_t n_rows, uint32_t *d_sites, cudaStream_t stream);

// transposed codes -> sequence-major [n][C] (test hook)
void launch_untranspose(const uint8_t *d_codes, int64_t n_pad, int64_t n, int64_t c0, int64_t C, uint8_t *d_out,
                        cudaStream_t stream);

// K3 per-codon Nei-Gojobori statistics from histograms.
// d_tables: [64*2 sites][2*64*64 diffs] doubles.  Any output may be null.
struct StatsOut {
    uint8_t *poly; uint32_t *ndef_a; uint32_t *ndef_b; unsigned long long *comparisons;
    double *stats4; uint8_t *conserved;
};
// po/ps (or null): multi-GPU -- the outputs also go into every peer's exchange region and stage kStageStats is signalled.
void launch_within_stats(GroupView A, int64_t n_codons, const double *d_tables, StatsOut out, const PeerOut *po, const PeerSync *ps,
                         cudaStream_t stream);
// launch_fused_fixup (every codon regular) and launch_within_stats in one kernel, codon by codon: same arithmetic.
void launch_fused_fixup_stats(const uint8_t *d_cons, int64_t cons_shift, uint32_t *d_alt, int64_t n_spans, const CodonMap *d_map, int64_t n_codons,
                              uint32_t n_rows, uint32_t *d_hist, uint32_t *d_other_min, uint32_t *d_other_max, GroupView A,
                              const double *d_tables, StatsOut out, const PeerOut *po, const PeerSync *ps, cudaStream_t stream);
// a rank whose codon range is empty still has to signal its stage
void launch_peer_signal(const PeerSync &ps, int stage, cudaStream_t stream);
// everything enqueued on `stream` after this sees the stage's data of every rank
void launch_peer_wait(const PeerSync &ps, int stage, cudaStream_t stream);
void launch_between_stats(GroupView A, GroupView B, int64_t n_codons, const double *d_tables, StatsOut out,
                          cudaStream_t stream);

// between-group min-taxa filters (bgp:412-467): zero the rows of the codons that fail (0 = criterion off)
void launch_taxa_filter(double *d_stats4, const uint32_t *d_ndef_a, const uint32_t *d_ndef_b, int64_t n_codons, uint32_t min_total,
                        uint32_t min_group, cudaStream_t stream);

// product sums: sums[p][4] = sum over codons [ofs[p], ofs[p+1]) of stats4
void launch_product_sums(const double *d_stats4, const int64_t *d_prod_ofs, int n_products, double *d_sums,
                         cudaStream_t stream);

// K4 resampling.  Item i (a product, a sliding window ...) resamples the codons [beg[i], end[i]) of stats4: every
// replicate is C = end - beg draws from C + slot0 slots (slot0 = 1: the reference's whole-product bootstrap,
// int(rand(C+1)) with a nonexistent codon 0, wg:885; slot0 = 0: its window bootstrap, wgp:406).  This launch covers
// the replicates [b_first, b_first + b_count).  order (or null) lists the items largest first.  Outputs (any may be
// null): rep[item][b][4] replicate sums; vals[(which * n_items + item) * vals_stride + vals_b0 + b] per-replicate
// dN, dS, dN-dS and their Jukes-Cantor forms (vals_stride <= 0: b_count); prod_sums[item][4] as launch_product_sums
// writes them, bit for bit.  Philox4x32-10: key = seed, the counter carries (draw quad, replicate, item, kind, call +
// *salt): kind separates the callers' streams, salt (or null) is a device word a replayed CUDA graph refreshes.
// pv/ps (world > 1): the values go into every rank's exchange region instead of vals, the kernel waits for stage
// kStageStats before it reads stats4 and signals kStageVals when it is done.
// SNPG_BOOT_CLASSES: the codons of an item whose row of the table is (N_sites, S_sites, 0, 0) with the same two numbers,
// bit for bit, form a class (every conserved codon of the same amino-acid degeneracy; the reference's empty slot 0 and
// all-zero rows are class 0).  A replicate's draws are multinomial over the slots, so the number that land in a class
// is all its sum needs from them: k_boot_split draws (general draws, class 0, class 1, ...) by a chain of exact
// binomials and only the draws on the n_general other codons are made one by one.
constexpr int kBootMaxClasses = 16;
struct BootClasses {
    uint32_t n_general, n_classes;            // codons drawn one by one (first in the item's rows of perm); classes in use
    uint32_t count[kBootMaxClasses];          // slots of each class
    double vn[kBootMaxClasses], vs[kBootMaxClasses];     // its N_sites, S_sites
    // the chain of binomials, node 0 = the general codons, node 1 + k = class k (the last class takes what is left):
    // success probability given the slots still open, folded to <= 1/2 (flip: count the failures), with log(p / (1 - p))
    // and log(1 - p)
    double node_p[kBootMaxClasses + 1], node_lpq[kBootMaxClasses + 1], node_l1q[kBootMaxClasses + 1];
    uint32_t node_flip[kBootMaxClasses + 1];
    uint32_t mma_blocks;                      // blocks of k_bootstrap_mma's grid row that have replicates of this item (BootArgs::mma_*)
};
struct BootArgs {
    const double *stats4 = nullptr; const int64_t *beg = nullptr, *end = nullptr; const int *order = nullptr;
    int n_items = 0, slot0 = 1;
    uint32_t item0 = 0;            // added to the item index in the Philox counter (batches of windows, group pairs)
    int64_t b_first = 0, b_count = 0, vals_stride = 0, vals_b0 = 0;
    uint2 key = {0, 0}; uint32_t kind = 0, call = 0; const uint32_t *salt = nullptr;
    double *rep = nullptr, *vals = nullptr, *prod_sums = nullptr;
    PeerVals pv; PeerSync ps;
    // SNPG_BOOT_CLASSES (items that do not overlap): scratch of boot_scratch_bytes(), carved by boot_scratch_carve()
    double *perm = nullptr; BootClasses *cls = nullptr; uint32_t *split_m = nullptr; double2 *split_ns = nullptr;
    int n_which = 6;               // per-replicate values written: 6 = dN, dS, dN-dS and their Jukes-Cantor forms, 3 = the first three (window SEs)
    int cls_shift = 6;             // a group of conserved codons becomes a class from max(32, C >> cls_shift) members on
    uint32_t mma_row_words = 0, mma_nb = 0, mma_split = 0;    // the shape launch_bootstrap gives k_bootstrap_mma (0: not that kernel), for k_boot_classify
};
cudaError_t prepare_device_kernels();      // once per ctx on its device: see kernels.cu
// Scratch of the class-aggregated bootstrap for items inside stats4[0, total) and b_count replicates of each.
size_t boot_scratch_bytes(int64_t total, int n_items, int64_t b_count);
void boot_scratch_carve(void *scratch, int64_t total, int n_items, int64_t b_count, BootArgs &a);
enum BootKind : uint32_t { kBootProduct = 1, kBootWindow = 2, kBootAllDevice = 3, kBootJob = 4, kBootBetween = 5, kBootJobWindow = 6, kBootRWindow = 7 };
// form: SNPG_BOOT_CLASSES (class counts by binomials, the other draws as WEIGHTS; needs a.perm etc., else WEIGHTS),
// SNPG_BOOT_WEIGHTS (slot counters + tensor-core contraction) or SNPG_BOOT_GATHER (per-draw row gathers);
// max_codons = the longest item (sizes the shared-memory table; beyond the limit: gathers from global memory).
void launch_bootstrap(int form, BootArgs a, int64_t max_codons, cudaStream_t stream);
// one Philox4x32-10 block on the device (known-answer test hook)
void launch_philox_kat(const uint32_t *ctr, const uint32_t *key, uint32_t *d_out, cudaStream_t stream);
// K6: rep[item][B][4] -> se[item][6]: sample SD (n-1) of the per-replicate dN, dS, dN-dS and their Jukes-Cantor forms
void launch_moments(const double *d_rep, int64_t n_items, int64_t B, double *d_se, cudaStream_t stream);
// the same from d_vals[which][item][B] (what launch_bootstrap's epilogue writes).  d_undefined (or null): [item][6]
// replicates whose value is not finite (Jukes-Cantor of a distance >= 3/4: the reference dies there, bgp:1047).
// ps (or null): wait for stage kStageVals first.
void launch_moments_vals(const double *d_vals, int64_t n_items, int64_t B, double *d_se, uint32_t *d_undefined, const PeerSync *ps,
                         cudaStream_t stream, int n_which = 6);     // n_which: only the first statistics (se[item][0 .. n_which))

// Row f4 (snpgenie.pl:6631-6760): per codon position, freq[c][64] codon frequencies and cov[c] mean coverage -> mean numbers of
// nonsynonymous and synonymous differences between two reads of the pool.
void launch_pooled_diffs(const double *d_freq, const double *d_cov, int64_t n, const double *d_tables, double *d_n_diffs, double *d_s_diffs,
                         cudaStream_t stream);

// Row f2 (SNPGenie_sliding_windows.R:241-402): rep[window][B][4] replicate sums + win_sums[window][4] -> out[window][16]
// (dN, dS, dN/dS, dN-dS, their bootstrap SEs and Z-test P values, the >/==/< 0 counts of dN-dS and the achieved
// significance levels).  num / den: 0 = N, 1 = S columns; jc != 0: Jukes-Cantor distances.
void launch_window_rstats(const double *d_rep, const double *d_win_sums, int64_t nwin, int64_t B, int num, int den, int jc, double *d_out,
                          cudaStream_t stream);

// K5 windows: window k sums the codons [beg[k], end[k]) in ascending order; its bootstrap is launch_bootstrap over the same items
void launch_window_sums(const double *d_stats4, const int64_t *d_beg, const int64_t *d_end, int64_t nwin, double *d_win_sums,
                        cudaStream_t stream);

}  // namespace snpg

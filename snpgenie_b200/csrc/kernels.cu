// sm_100a kernels for SNPGenie's within-/between-group Nei-Gojobori path.
//
// K1 k_pack        alignment bytes -> 1-byte codon codes (transposed [codon][seq])
// K2 k_hist        per-codon 66-bin histogram over sequences
// K3 k_stats       per-codon N/S sites and differences from histograms
// K4 k_bootstrap_mma  whole-product codon resampling (Philox4x32-10): slot counters x table on the FP64 tensor cores
//    k_bootstrap      the same with per-draw row gathers (option / tables beyond shared memory)
// K5 k_window_*    sliding-window sums and per-window resampling
// K6 k_rep_moments two-pass SD of per-replicate dN, dS, dN-dS (+ Jukes-Cantor)
//
// The reference blocks each kernel replaces are cited at the kernels.  All
// arithmetic that reaches an output table is FP64; K1/K2 are byte/integer work
// and HBM-bound (see DESIGN.md for the algorithmic bytes of each).
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>

#include <cooperative_groups.h>

#include "snpg_device.cuh"
#include "../../include/snpgenie_cuda.h"
#include "snpg_internal.h"
#include "snpg_peer.cuh"
#include "snpg_span.cuh"

namespace snpg {

namespace {

__device__ __forceinline__ uint4 ld_stream_v4(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ uint4 ld_nc_v4(const uint4 *p) {          // may hit a line a prefetch brought into L1
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ uint32_t warp_sum_u32(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// ---------------------------------------------------------------------------
// K1: codon packing.  Replaces product extraction + codon split
// (wg:286-329, :417-454; bg:238-298, :351-388) and the per-comparison
// definedness regexes (wg:493,498,601,604).  A block handles 32 codons x 128
// sequences: lanes run along codons when reading (a '+' CDS makes the three byte
// loads of a warp fall in one or two 128-B lines of the row), the tile is
// turned in shared memory, and each codon's 128 codes leave as 16-B stores.
// ---------------------------------------------------------------------------
constexpr int kPackCodons = 32;
constexpr int kPackRows = 128;
constexpr int kPackPitch = 144;  // bytes, multiple of 16

__global__ void __launch_bounds__(256)
k_pack(const uint8_t *__restrict__ aln, int64_t stride, int64_t n_rows, int64_t row0,
       const CodonMap *__restrict__ map, int64_t n_codons, uint8_t *__restrict__ codes, int64_t n_pad,
       uint32_t *__restrict__ other_min, uint32_t *__restrict__ other_max)
{
    __shared__ uint8_t lut_chr[2][256];
    __shared__ uint8_t lut_cls[2][256];
    __shared__ __align__(16) uint8_t tile[kPackCodons][kPackPitch];

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int s = 0; s < 2; s++) {
        uint8_t n = norm_char((uint8_t)t, s);
        lut_chr[s][t] = n;
        lut_cls[s][t] = class_of(n);
    }
    __syncthreads();

    const int64_t c = (int64_t)blockIdx.x * kPackCodons + lane;
    const int64_t tile_row0 = (int64_t)blockIdx.y * kPackRows;
    const bool live = c < n_codons;
    CodonMap m = {0, 0, 0, 0};
    if (live) m = map[c];
    const int s = m.minus ? 1 : 0;

#pragma unroll 4
    for (int i = 0; i < 16; i++) {
        const int lr = warp * 16 + i;
        const int64_t r = tile_row0 + lr;
        uint8_t code = kCodePad;
        if (live && r < n_rows) {
            const uint8_t *row = aln + r * stride;
            const uint8_t b0 = row[m.p0], b1 = row[m.p1], b2 = row[m.p2];
            const uint8_t k0 = lut_cls[s][b0], k1 = lut_cls[s][b1], k2 = lut_cls[s][b2];
            const uint8_t worst = max(k0, max(k1, k2));
            if (worst < 4) {
                code = (uint8_t)(k0 * 16 + k1 * 4 + k2);
            } else if (k0 == 4 || k1 == 4 || k2 == 4) {
                code = kCodeUndef;
            } else {
                code = kCodeOther;
                const uint32_t key = ((uint32_t)lut_chr[s][b0] << 16) | ((uint32_t)lut_chr[s][b1] << 8) | lut_chr[s][b2];
                atomicMin(&other_min[c], key);
                atomicMax(&other_max[c], key);
            }
        }
        tile[lane][lr] = code;
    }
    __syncthreads();

    const int j = t >> 3, v = t & 7;
    const int64_t cj = (int64_t)blockIdx.x * kPackCodons + j;
    if (cj < n_codons) {
        const uint4 val = *reinterpret_cast<const uint4 *>(&tile[j][v * 16]);
        *reinterpret_cast<uint4 *>(codes + cj * n_pad + row0 + tile_row0 + v * 16) = val;
    }
}

// The first defined codon of every codon column in row order (bg:853-890 picks it when the other group has no defined
// codon at all).  One thread per codon; nearly every thread is done after row 0.
__global__ void k_first_defined(const uint8_t *__restrict__ aln, int64_t stride, int64_t n_rows, const CodonMap *__restrict__ map,
                                int64_t n_codons, uint8_t *__restrict__ first)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_codons || first[c] != 0xFF) return;
    const CodonMap m = map[c];
    for (int64_t r = 0; r < n_rows; r++) {
        const uint8_t *row = aln + r * stride;
        const uint8_t k0 = class_of(norm_char(row[m.p0], m.minus)), k1 = class_of(norm_char(row[m.p1], m.minus)),
                      k2 = class_of(norm_char(row[m.p2], m.minus));
        if (k0 == 4 || k1 == 4 || k2 == 4) continue;             // undefined: N or '-' (wg:493,498)
        first[c] = max(k0, max(k1, k2)) < 4 ? (uint8_t)(k0 * 16 + k1 * 4 + k2) : (uint8_t)kCodeOther;
        return;
    }
}

// ---------------------------------------------------------------------------
// K2: per-codon histogram.  Replaces the polymorphism pre-check (wg:469-541;
// bg:472-544) and the O(n^2) pair counting (wg:596-609; bg:649-667): the 66
// counts of a column determine every per-codon output.
// One warp owns (column, row chunk).  Lanes stream 16-byte vectors of the
// column; each 4-code word is XORed with the chunk's first code replicated,
// and only the bytes that differ go to the warp's private shared-memory
// histogram.  The matching code's count is rows - mismatches, so a conserved
// column costs three ALU ops per 4 sequences and no atomics at all.
// Algorithmic traffic: 1 byte per (sequence, codon) read + 264 B per codon written.
// ---------------------------------------------------------------------------
constexpr int kHistWarps = 8;
constexpr int kHistChunkRows = 8192;

__device__ __forceinline__ void hist_word(uint32_t w, uint32_t g4, uint32_t *wh, uint32_t &mism) {
    const uint32_t d = w ^ g4;                                   // codes < 128, so no byte has bit 7
    uint32_t m = (d + 0x7F7F7F7Fu) & 0x80808080u;                // bit 7 of every byte that differs
    if (m) {
        mism += __popc(m);
        do {
            const int bit = __ffs(m) - 1;                        // 7, 15, 23 or 31
            atomicAdd(&wh[(w >> (bit - 7)) & 0xFFu], 1u);
            m &= m - 1;
        } while (m);
    }
}

__global__ void __launch_bounds__(kHistWarps * 32)
k_hist(const uint8_t *__restrict__ codes, int64_t n_pad, int64_t n_codons, int chunks, int64_t chunk_rows,
       uint32_t *__restrict__ hist)
{
    __shared__ uint32_t wh_all[kHistWarps][128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *wh = wh_all[warp];
    for (int i = lane; i < 128; i += 32) wh[i] = 0;
    __syncwarp();

    const int64_t item = (int64_t)blockIdx.x * kHistWarps + warp;
    if (item >= n_codons * chunks) return;
    const int64_t col = item / chunks;
    const int64_t r0 = (item % chunks) * chunk_rows;
    const int64_t r1 = min(n_pad, r0 + chunk_rows);
    if (r0 >= r1) return;
    const uint8_t *base = codes + col * n_pad + r0;
    const uint4 *p = reinterpret_cast<const uint4 *>(base);
    const int nvec = (int)((r1 - r0) >> 4);
    const uint32_t guess = base[0];
    const uint32_t g4 = guess * 0x01010101u;
    uint32_t mism = 0;

    for (int v0 = 0; v0 < nvec; v0 += 128) {
        uint4 x[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int v = v0 + k * 32 + lane;
            x[k] = v < nvec ? ld_stream_v4(p + v) : make_uint4(g4, g4, g4, g4);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            hist_word(x[k].x, g4, wh, mism);
            hist_word(x[k].y, g4, wh, mism);
            hist_word(x[k].z, g4, wh, mism);
            hist_word(x[k].w, g4, wh, mism);
        }
    }
    mism = warp_sum_u32(mism);
    __syncwarp();
    const uint32_t matched = (uint32_t)(r1 - r0) - mism;
    for (int b = lane; b < kHistBins; b += 32) {
        const uint32_t cnt = wh[b] + ((uint32_t)b == guess ? matched : 0u);
        if (cnt) atomicAdd(&hist[col * kHistBins + b], cnt);
    }
}

// ---------------------------------------------------------------------------
// K1+K2 fused: histogram straight from the alignment bytes, no code matrix.
// Used when the packed codon matrix is not kept (SNPG_OPT_KEEP_CODES = 0).
//
// A thread owns one 16-byte column span of the alignment and walks down rows;
// a warp covers 31 spans plus the 16 bytes after them, 512 contiguous bytes of
// each row.  Every row's span is XORed against three reference rows
// (k_vote_refs): the per-column consensus and the span's two most frequent
// other patterns.  If the 16 bytes (plus the 2 bytes a codon may reach into the
// next span) equal the consensus -- the common case in real alignments -- every
// codon starting in the span equals the consensus codon and nothing is recorded;
// a row equal to one of the other two only bumps a counter (k_fused_fixup decodes
// the pattern once, weighted).  Otherwise the bytes are parked and only the codons
// that contain a byte differing from the consensus are decoded, 32 parked spans
// at a time, and counted with a global reduction (RED) into hist[codon][code].
// k_fused_fixup then credits each codon's consensus code with (rows - recorded).
// Bytes are compared raw, so a lower-case or 'U' residue simply takes the decode
// path.
// Algorithmic traffic: every alignment byte of the CDS span read once.
// ---------------------------------------------------------------------------
constexpr int kFusedWarps = 8;
constexpr int kFusedEntryBytes = 32;                                  // 5 raw words + span index + pad
constexpr int kFusedSpans = 31;                                       // spans per warp (lane 31 fetches the look-ahead)
constexpr int kFusedTailRows = 4;                                     // rows per chunk in the fine-grained tail region
static_assert(kHistBins % 3 == 0, "the decode steps through hist[] by kHistBins / 3 per byte of codon offset");

// One parked (span, row): decode and count exactly the codons that contain a byte
// differing from the consensus.  e = shared-memory address of 5 raw words + span index.
__device__ __forceinline__ void fused_decode(uint32_t e, const uint8_t *__restrict__ cons,
                                             const SpanRun *__restrict__ runs, uint32_t *__restrict__ hist,
                                             uint32_t *__restrict__ other_min, uint32_t *__restrict__ other_max,
                                             uint32_t lut_cls, uint32_t lut_chr)
{
    uint4 v;
    uint2 t;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(e) : "memory");
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(t.x), "=r"(t.y) : "r"(e + 16) : "memory");
    const uint32_t sp = t.y;
    const uint2 *rp = reinterpret_cast<const uint2 *>(runs + (size_t)sp * kSpanRuns);
    const uint4 c = __ldg(reinterpret_cast<const uint4 *>(cons + (size_t)sp * 16));
    const uint32_t c4 = __ldg(reinterpret_cast<const uint32_t *>(cons + (size_t)sp * 16 + 16));
    uint2 rw = __ldg(rp);                                             // the first run, asked for with the consensus
    const uint32_t m20 = diff_mask20(v, t.x, c, c4);
#pragma unroll 1
    for (int l = 0;;) {
        const uint32_t cnt = (rw.y >> 8) & 0xFFu;
        if (cnt == 0) break;
        const uint32_t off = rw.y & 0xFFu, mi = (rw.y >> 24) & 1u;
        const int dir = (int)(int8_t)((rw.y >> 16) & 0xFF);
        const uint32_t w = m20 >> off;
        uint32_t cm = (w | (w >> 1) | (w >> 2)) & 0x9249u & ((1u << (3 * cnt)) - 1u);
        if (cm) {
            const uint32_t cls = lut_cls + (mi << 8);
            const uint32_t rb = e + off;
            const uint32_t s_lo = mi ? 0u : 4u, s_hi = mi ? 4u : 0u;     // '-' strand reads the codon right to left
            const uint32_t base = rw.x * (uint32_t)kHistBins;
            const int step = dir * (kHistBins / 3);                      // per byte of offset: j = 3 * codon ordinal
            do {
                const int j = __ffs(cm) - 1;
                cm &= cm - 1;
                const uint32_t r0 = lds_u8(rb + j), r1 = lds_u8(rb + j + 1), r2 = lds_u8(rb + j + 2);
                const uint32_t k0 = lds_u8(cls + r0), k1 = lds_u8(cls + r1), k2 = lds_u8(cls + r2);
                uint32_t code = (k0 << s_lo) + k1 * 4 + (k2 << s_hi);
                if (max(k0, max(k1, k2)) >= 4) {
                    code = (k0 == 4 || k1 == 4 || k2 == 4) ? kCodeUndef : kCodeOther;
                    if (code == kCodeOther) {
                        const uint32_t chr = lut_chr + (mi << 8);
                        const uint32_t n0 = lds_u8(chr + r0), n1 = lds_u8(chr + r1), n2 = lds_u8(chr + r2);
                        const uint32_t key = mi ? (n2 << 16) | (n1 << 8) | n0 : (n0 << 16) | (n1 << 8) | n2;
                        const int cidx = (int)rw.x + dir * (int)((uint32_t)j * 0xABu >> 9);   // j / 3 for j < 48
                        atomicMin(&other_min[cidx], key);
                        atomicMax(&other_max[cidx], key);
                    }
                }
                atomicAdd(&hist[base + (uint32_t)(step * j) + code], 1u);
            } while (cm);
        }
        if (++l == kSpanRuns) break;
        rw = __ldg(rp + l);
    }
}

// Persistent warps.  A warp works on a block of kFusedSpans = 31 spans: lanes 0-30 own one span each and lane 31
// only fetches the next 16 bytes, whose first word is lane 30's look-ahead (every lane's look-ahead is its
// right-hand neighbour's first word).  The (span block) x (rows) plane is cut into chunks that warps claim
// from a global counter (sched[0]): chunks of chunk_a rows over rows [0, rows_a), then chunks of
// kFusedTailRows rows over the rest, so the last claims are small and every SM finishes at about the same
// time whatever the local mutation density.  The parked-span queue carries the span index, so it lives across
// chunks; the counts of rows equal to a span's second and third reference go to alt[span][..] (one RED each per
// chunk) and k_fused_fixup decodes those patterns once, weighted.  The last block to finish leaves sched[]
// zeroed for the next launch.
// UNROLL rows are loaded at a time; parked spans are decoded after every DRAIN_EVERY rows, so the ring needs
// 31 + 32 * DRAIN_EVERY entries (QUEUE, a power of two).  inv_sb = floor(2^32 / n_sb).
template <int UNROLL, int DRAIN_EVERY, int QUEUE, int MIN_BLOCKS, int PREFETCH = 0>
__global__ void __launch_bounds__(kFusedWarps * 32, MIN_BLOCKS)
k_fused_hist(const uint8_t *__restrict__ aln, int64_t pitch, uint32_t n_rows, const uint8_t *__restrict__ cons,
             uint32_t *__restrict__ alt, uint32_t n_spans, const SpanRun *__restrict__ runs, uint32_t *__restrict__ hist,
             uint32_t *__restrict__ other_min, uint32_t *__restrict__ other_max, uint32_t *__restrict__ sched,
             uint32_t n_sb, uint32_t inv_sb, uint32_t chunk_a, uint32_t rows_a, uint32_t k_a, uint32_t k_total,
             uint32_t dense_min, uint32_t chunk_d, uint32_t rows_a_d, uint32_t k_a_d, uint32_t k_total_d)
{
    static_assert(QUEUE >= 31 + 32 * DRAIN_EVERY && (QUEUE & (QUEUE - 1)) == 0 && UNROLL % DRAIN_EVERY == 0, "ring too small");
    constexpr int kFusedQueueBytes = QUEUE * kFusedEntryBytes;        // the ring is aligned to its size
    extern __shared__ __align__(16) uint8_t fused_smem[];             // kFusedWarps rings + one ring of slack for the alignment
    __shared__ uint8_t lut_chr[2 * 256];
    __shared__ uint8_t lut_cls[2 * 256];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int s = 0; s < 2; s++) {
        const uint8_t n = norm_char((uint8_t)t, s);
        lut_chr[s * 256 + t] = n;
        lut_cls[s * 256 + t] = class_of(n);
    }
    __syncthreads();

    // Rows that match none of the three references are expensive (parked, decoded); where the vote's sample says they are
    // many (sched[2] = unmatched (span, sample row) pairs, left by k_vote_refs, zeroed by the last block below), small
    // chunks even out the warps' shares; where rows are cheap, a chunk's fixed cost (claim, references, first loads) wants
    // 32 rows.  Both geometries come from the host (launch_fused_hist).
    if (sched[2] >= dense_min) { chunk_a = chunk_d; rows_a = rows_a_d; k_a = k_a_d; k_total = k_total_d; }

    const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(fused_smem);
    const uint32_t pad = (kFusedQueueBytes - (smem0 & (kFusedQueueBytes - 1))) & (kFusedQueueBytes - 1);
    uint32_t q_base = smem0 + pad + warp * kFusedQueueBytes;
    uint32_t a_cls = (uint32_t)__cvta_generic_to_shared(lut_cls), a_chr = (uint32_t)__cvta_generic_to_shared(lut_chr);
    uint32_t q_head = 0, q_tail = 0;    // byte offsets into the ring, warp-uniform, free-running
    unsigned lt = (1u << lane) - 1;
    const uint32_t lane_ofs = lane * kFusedEntryBytes;
    uint32_t one = 1u + (uint32_t)lane;          // 1 in lane 0, the only lane that claims
    const RefPlanes refs = ref_planes(alt, n_spans);

    auto drain = [&]() {     // at most 31 + 32 * DRAIN_EVERY parked entries between drains
        while (q_tail - q_head >= 32u * kFusedEntryBytes) {
            __syncwarp();
            fused_decode(q_base | ((q_head + lane_ofs) & (kFusedQueueBytes - 1)), cons, runs, hist, other_min, other_max, a_cls, a_chr);
            __syncwarp();
            q_head += 32u * kFusedEntryBytes;
        }
    };

    uint32_t k_next = 0;
    if (lane == 0) k_next = claim_async(sched, one);
    uint32_t k = __shfl_sync(kFull, k_next, 0);
    while (k < k_total) {
        const bool in_a = k < k_a;
        uint32_t rb, sb;
        div_mod(in_a ? k : k - k_a, n_sb, inv_sb, rb, sb);
        const uint32_t row0 = in_a ? rb * chunk_a : rows_a + rb * kFusedTailRows;
        const uint32_t n_my = in_a ? chunk_a : min((uint32_t)kFusedTailRows, n_rows - row0);
        const uint32_t span = sb * kFusedSpans + lane;
        uint32_t live = lane < kFusedSpans && span < n_spans ? 0xFFFFFFFFu : 0u;
        uint32_t sp = min(span, n_spans - 1);        // the references of a dead lane are never used
        // pinned in registers: ptxas otherwise re-derives them inside the row loop (S2R, LDC, compares, LEA)
        asm volatile("" : "+r"(lt), "+r"(q_base), "+r"(sp), "+r"(live));
        const uint4 c = *reinterpret_cast<const uint4 *>(cons + (size_t)sp * 16);
        const uint32_t c4 = *reinterpret_cast<const uint32_t *>(cons + (size_t)sp * 16 + 16);
        // the span's second and third reference: its most common non-consensus bytes among the sampled rows (the
        // consensus itself when the sample had none).  Rows equal to one of them are only counted here.
        const uint4 a = refs.a16[sp], b = refs.b16[sp];
        const uint32_t tails = refs.tails[sp], a4 = tails & 0xFFFFu, b4 = tails >> 16;
        uint32_t n_alt = 0, n_alt2 = 0;
        const uint8_t *p = aln + (size_t)min(span, n_spans) * 16 + (size_t)row0 * pitch;

        // one row of the warp: compare with the three references, count or park
        auto row = [&](const uint4 &v) {
            const uint32_t v4 = __shfl_down_sync(kFull, v.x, 1);
            const uint32_t d = ((v.x ^ c.x) | (v.y ^ c.y) | (v.z ^ c.z) | (v.w ^ c.w) | ((v4 ^ c4) & 0xFFFFu)) & live;
            const uint32_t da = (v.x ^ a.x) | (v.y ^ a.y) | (v.z ^ a.z) | (v.w ^ a.w) | ((v4 ^ a4) & 0xFFFFu);
            const uint32_t db = (v.x ^ b.x) | (v.y ^ b.y) | (v.z ^ b.z) | (v.w ^ b.w) | ((v4 ^ b4) & 0xFFFFu);
            n_alt += da == 0 ? 1u : 0u;          // the three references are pairwise distinct: at most one matches
            n_alt2 += db == 0 ? 1u : 0u;
            const bool differs = d != 0 && da != 0 && db != 0;
            const unsigned bal = __ballot_sync(kFull, differs);
            if (bal) {           // nothing to park on most rows of a realistic alignment
                if (differs) {   // park the bytes; a full warp decodes 32 parked spans at a time
                    const uint32_t at = q_base | ((q_tail + __popc(bal & lt) * kFusedEntryBytes) & (kFusedQueueBytes - 1));
                    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(at), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
                    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(at + 16), "r"(v4), "r"(sp) : "memory");
                }
                q_tail += __popc(bal) * kFusedEntryBytes;
            }
        };

        // the next chunk is claimed while the last full group of rows of this one is processed: early enough to
        // hide most of the atomic's latency, late enough that its result does not sit in a register for long
        const uint32_t claim_at = n_my >= UNROLL ? (n_my / UNROLL - 1) * UNROLL : 0xFFFFFFFFu;
        uint32_t i = 0;
        for (; i + UNROLL <= n_my; i += UNROLL) {
            if (i == claim_at && lane == 0) k_next = claim_async(sched, one);
            uint4 v[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                v[u] = PREFETCH ? ld_nc_v4(reinterpret_cast<const uint4 *>(p)) : ld_stream_v4(reinterpret_cast<const uint4 *>(p));
                p += pitch;
            }
            if (PREFETCH) {      // the next group of rows into L1 while this one is compared (experiment)
                const uint8_t *q = p;
#pragma unroll
                for (int u = 0; u < UNROLL; u++) {
                    if (PREFETCH == 1 || (lane & 7) == 0 || lane == 31) prefetch_l1(q);
                    q += pitch;
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                row(v[u]);
                if ((u + 1) % DRAIN_EVERY == 0) drain();
            }
        }
        if (claim_at == 0xFFFFFFFFu && lane == 0) k_next = claim_async(sched, one);
        for (; i < n_my; i++, p += pitch) {          // the last rows of the group
            const uint4 v = ld_stream_v4(reinterpret_cast<const uint4 *>(p));
            row(v);
            drain();
        }
        if (live && n_alt) atomicAdd(&refs.cnt_a[sp], n_alt);
        if (live && n_alt2) atomicAdd(&refs.cnt_b[sp], n_alt2);
        k = __shfl_sync(kFull, k_next, 0);
    }
    __syncwarp();
    if ((uint32_t)lane_ofs < q_tail - q_head)
        fused_decode(q_base | ((q_head + lane_ofs) & (kFusedQueueBytes - 1)), cons, runs, hist, other_min, other_max, a_cls, a_chr);
    __syncthreads();
    if (t == 0 && atomicAdd(&sched[1], 1u) == gridDim.x - 1) { sched[0] = 0; sched[1] = 0; sched[2] = 0; }
}

// ---------------------------------------------------------------------------
// k_fused_hist_db: k_fused_hist with DOUBLE-BUFFERED row groups.  While a warp compares, parks and decodes the G rows
// of one group, the loads of the next group are in flight (in k_fused_hist a warp has loads in flight only while it
// waits for them: 39 % of its stall samples sit on the first use of a row group).  ptxas puts every row load on ONE
// scoreboard, so waiting for any row waits for all loads issued so far: the next group's loads must be issued AFTER the
// first use of the current group, and that order is forced by a data dependency ptxas cannot fold (the address of the
// next loads adds `word & pitch & 15`, always 0 because the pitch is a multiple of 16; it does fold `(x*x+x)&1`).
// Same arithmetic, chunk geometry, ring and decode as k_fused_hist; 16 more registers (3 blocks per SM).
// Measured and not kept on the way (profiles/README.md): refilling the four registers of a row as soon as the row is
// done (the newest load is waited for at every row: 85 us against 77); a ring of two conflict-free arrays with the
// decode reading consensus and runs from per-warp shared-memory copies (fewer wavefronts, but 58 M instructions
// instead of 42 M at the 80-register cap: 81 us); the same copies with the ring left as it is (47 M instructions, but
// 18 KB more shared memory per block takes the L1 space those loads were hitting: 79.6 us).
// ---------------------------------------------------------------------------
template <int G, int QUEUE, int MIN_BLOCKS, int WARPS = kFusedWarps, int DRAIN_EVERY = (G <= 3 ? G : 2)>
__global__ void __launch_bounds__(WARPS * 32, MIN_BLOCKS)
k_fused_hist_db(const uint8_t *__restrict__ aln, int64_t pitch, uint32_t n_rows, const uint8_t *__restrict__ cons,
                uint32_t *__restrict__ alt, uint32_t n_spans, const SpanRun *__restrict__ runs, uint32_t *__restrict__ hist,
                uint32_t *__restrict__ other_min, uint32_t *__restrict__ other_max, uint32_t *__restrict__ sched,
                uint32_t n_sb, uint32_t inv_sb, uint32_t chunk_a, uint32_t rows_a, uint32_t k_a, uint32_t k_total,
             uint32_t dense_min, uint32_t chunk_d, uint32_t rows_a_d, uint32_t k_a_d, uint32_t k_total_d)
{
    static_assert(QUEUE >= 31 + 32 * DRAIN_EVERY && (QUEUE & (QUEUE - 1)) == 0 && G % DRAIN_EVERY == 0, "ring too small");
    constexpr int kFusedQueueBytes = QUEUE * kFusedEntryBytes;        // the ring is aligned to its size
    extern __shared__ __align__(16) uint8_t fused_smem[];             // kFusedWarps rings + one ring of slack for the alignment
    __shared__ uint8_t lut_chr[2 * 256];
    __shared__ uint8_t lut_cls[2 * 256];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int i = t; i < 512; i += WARPS * 32) {
        const uint8_t n = norm_char((uint8_t)i, i >> 8);
        lut_chr[i] = n;
        lut_cls[i] = class_of(n);
    }
    __syncthreads();

    // Rows that match none of the three references are expensive (parked, decoded); where the vote's sample says they are
    // many (sched[2] = unmatched (span, sample row) pairs, left by k_vote_refs, zeroed by the last block below), small
    // chunks even out the warps' shares; where rows are cheap, a chunk's fixed cost (claim, references, first loads) wants
    // 32 rows.  Both geometries come from the host (launch_fused_hist).
    if (sched[2] >= dense_min) { chunk_a = chunk_d; rows_a = rows_a_d; k_a = k_a_d; k_total = k_total_d; }

    const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(fused_smem);
    const uint32_t pad = (kFusedQueueBytes - (smem0 & (kFusedQueueBytes - 1))) & (kFusedQueueBytes - 1);
    uint32_t q_base = smem0 + pad + warp * kFusedQueueBytes;
    uint32_t a_cls = (uint32_t)__cvta_generic_to_shared(lut_cls), a_chr = (uint32_t)__cvta_generic_to_shared(lut_chr);
    uint32_t q_head = 0, q_tail = 0;    // byte offsets into the ring, warp-uniform, free-running
    unsigned lt = (1u << lane) - 1;
    const uint32_t lane_ofs = lane * kFusedEntryBytes;
    uint32_t one = 1u + (uint32_t)lane;          // 1 in lane 0, the only lane that claims
    const RefPlanes refs = ref_planes(alt, n_spans);

    auto drain = [&]() {     // at most 31 + 32 * DRAIN_EVERY parked entries between drains
        while (q_tail - q_head >= 32u * kFusedEntryBytes) {
            __syncwarp();
            fused_decode(q_base | ((q_head + lane_ofs) & (kFusedQueueBytes - 1)), cons, runs, hist, other_min, other_max, a_cls, a_chr);
            __syncwarp();
            q_head += 32u * kFusedEntryBytes;
        }
    };
    auto flush_ring = [&]() {   // fewer than 32 entries are left after a drain
        __syncwarp();
        if ((uint32_t)lane_ofs < q_tail - q_head)
            fused_decode(q_base | ((q_head + lane_ofs) & (kFusedQueueBytes - 1)), cons, runs, hist, other_min, other_max, a_cls, a_chr);
        __syncwarp();
        q_head = q_tail;
    };

    uint32_t k_next = 0;
    if (lane == 0) k_next = claim_async(sched, one);
    uint32_t k = __shfl_sync(kFull, k_next, 0);
    while (k < k_total) {
        const bool in_a = k < k_a;
        uint32_t rb, sb;
        div_mod(in_a ? k : k - k_a, n_sb, inv_sb, rb, sb);
        const uint32_t row0 = in_a ? rb * chunk_a : rows_a + rb * kFusedTailRows;
        const uint32_t n_my = in_a ? chunk_a : min((uint32_t)kFusedTailRows, n_rows - row0);
        const uint32_t span = sb * kFusedSpans + lane;
        uint32_t live = lane < kFusedSpans && span < n_spans ? 0xFFFFFFFFu : 0u;
        uint32_t sp = min(span, n_spans - 1);        // the references of a dead lane are never used
        asm volatile("" : "+r"(lt), "+r"(q_base), "+r"(sp), "+r"(live));
        const uint8_t *p = aln + (size_t)min(span, n_spans) * 16 + (size_t)row0 * pitch;
        uint4 v[G], w[G];
        auto fetch = [&](uint4 (&x)[G]) {
#pragma unroll
            for (int u = 0; u < G; u++) {
                x[u] = ld_stream_v4(reinterpret_cast<const uint4 *>(p));
                p += pitch;
            }
        };
        const uint32_t n_grp = n_my / G;             // a chunk is whole groups, but for the last rows of the alignment
        if (n_grp) fetch(v);
        const uint4 c = *reinterpret_cast<const uint4 *>(cons + (size_t)sp * 16);
        const uint32_t c4 = *reinterpret_cast<const uint32_t *>(cons + (size_t)sp * 16 + 16);
        const uint4 a = refs.a16[sp], b = refs.b16[sp];
        const uint32_t tails = refs.tails[sp], a4 = tails & 0xFFFFu, b4 = tails >> 16;
        uint32_t n_alt = 0, n_alt2 = 0;

        // one row of the warp: compare with the three references, count or park
        auto row = [&](const uint4 &x) {
            const uint32_t x4 = __shfl_down_sync(kFull, x.x, 1);
            const uint32_t d = ((x.x ^ c.x) | (x.y ^ c.y) | (x.z ^ c.z) | (x.w ^ c.w) | ((x4 ^ c4) & 0xFFFFu)) & live;
            const uint32_t da = (x.x ^ a.x) | (x.y ^ a.y) | (x.z ^ a.z) | (x.w ^ a.w) | ((x4 ^ a4) & 0xFFFFu);
            const uint32_t db = (x.x ^ b.x) | (x.y ^ b.y) | (x.z ^ b.z) | (x.w ^ b.w) | ((x4 ^ b4) & 0xFFFFu);
            n_alt += da == 0 ? 1u : 0u;          // the three references are pairwise distinct: at most one matches
            n_alt2 += db == 0 ? 1u : 0u;
            const bool differs = d != 0 && da != 0 && db != 0;
            const unsigned bal = __ballot_sync(kFull, differs);
            if (bal) {
                if (differs) {
                    const uint32_t at = q_base | ((q_tail + __popc(bal & lt) * kFusedEntryBytes) & (kFusedQueueBytes - 1));
                    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(at), "r"(x.x), "r"(x.y), "r"(x.z), "r"(x.w) : "memory");
                    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(at + 16), "r"(x4), "r"(sp) : "memory");
                }
                q_tail += __popc(bal) * kFusedEntryBytes;
            }
        };
        auto rows = [&](uint4 (&x)[G]) {
#pragma unroll
            for (int u = 0; u < G; u++) {
                row(x[u]);
                if ((u + 1) % DRAIN_EVERY == 0) drain();
            }
        };

        for (uint32_t g = 0; g < n_grp; g += 2) {
            if (g + 2 >= n_grp && lane == 0) k_next = claim_async(sched, one);    // during the last one or two groups
            p += v[0].x & (uint32_t)pitch & 15u;     // always 0: orders the next loads after the first use of this group
            if (g + 1 < n_grp) fetch(w);
            rows(v);
            if (g + 1 >= n_grp) break;
            p += w[0].x & (uint32_t)pitch & 15u;
            if (g + 2 < n_grp) fetch(v);
            rows(w);
        }
        if (n_grp == 0 && lane == 0) k_next = claim_async(sched, one);
        for (uint32_t i = n_grp * G; i < n_my; i++, p += pitch) {     // the last rows of the last chunk
            const uint4 x = ld_stream_v4(reinterpret_cast<const uint4 *>(p));
            row(x);
            drain();
        }
        if (live && n_alt) atomicAdd(&refs.cnt_a[sp], n_alt);
        if (live && n_alt2) atomicAdd(&refs.cnt_b[sp], n_alt2);
        k = __shfl_sync(kFull, k_next, 0);
    }
    flush_ring();
    __syncthreads();
    if (t == 0 && atomicAdd(&sched[1], 1u) == gridDim.x - 1) { sched[0] = 0; sched[1] = 0; sched[2] = 0; }
}

// The three reference rows of the compare-with-reference kernels, from the first n_sample (<= 32) rows of
// the group.  One warp per 16-byte span, one lane per sampled row.
//   cons[column]   the most frequent byte of every column (lowest row on ties)
//   alt            (RefPlanes, snpg_span.cuh) the span's two most frequent 18-byte patterns (16 bytes + the 2 a codon may
//                  reach into the next span) that differ from the consensus.  A span whose sample has fewer patterns gets
//                  made-up ones (the consensus with bit 7 of byte 0 or byte 1 flipped), so that the three references are
//                  always pairwise distinct and a row can equal at most one of them.
// ANY rows are valid references (the kernels are exact for every choice); frequent ones just keep more
// rows off the decode path (a first sequence carrying minor alleles would make most rows differ).
// Span n_spans only supplies the look-ahead bytes of the last real span.
__device__ __forceinline__ uint32_t plurality_byte(uint32_t b, unsigned valid, int lane) {
    const unsigned same = __match_any_sync(kFull, b) & valid;
    uint32_t key = ((valid >> lane) & 1u) ? ((uint32_t)__popc(same) << 13) | ((uint32_t)(31 - lane) << 8) | b : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) key = max(key, __shfl_xor_sync(kFull, key, o));
    return key & 0xFFu;
}

__global__ void __launch_bounds__(256)
k_vote_refs(const uint8_t *__restrict__ aln, int64_t pitch, int n_sample, int64_t n_spans, uint8_t *__restrict__ cons,
            uint32_t *__restrict__ alt, uint4 *__restrict__ zero_hist, int64_t zero_vec, uint32_t *__restrict__ other_min,
            uint32_t *__restrict__ other_max, int64_t n_codons, uint32_t *__restrict__ unmatched_total)
{
    // the group's histograms start from zero: cleared here (when asked to) instead of by three memsets ahead of the kernel
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < zero_vec; i += (int64_t)gridDim.x * blockDim.x)
        zero_hist[i] = make_uint4(0, 0, 0, 0);
    if (other_min)
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_codons; i += (int64_t)gridDim.x * blockDim.x) {
            other_min[i] = 0xFFFFFFFFu;
            other_max[i] = 0;
        }
    const int lane = threadIdx.x & 31;
    const int64_t sp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (sp > n_spans) return;
    const unsigned valid = n_sample >= 32 ? kFull : ((1u << n_sample) - 1u);
    const uint8_t *pi = aln + (size_t)min(lane, n_sample - 1) * pitch + sp * 16;     // lanes past the sample repeat its last row
    uint32_t x[5] = {0, 0, 0, 0, 0};
    x[0] = *reinterpret_cast<const uint32_t *>(pi);
    if (sp < n_spans) {
        const uint4 v = *reinterpret_cast<const uint4 *>(pi);
        x[1] = v.y; x[2] = v.z; x[3] = v.w;
        x[4] = *reinterpret_cast<const uint32_t *>(pi + 16) & 0xFFFFu;
    }
    // consensus: per column plurality (18 columns; the last two repeat the next span's first two)
    uint32_t c[5] = {0, 0, 0, 0, 0};
    const int n_cols = sp < n_spans ? 18 : 4;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        // most words are the same in every sampled row (lanes past the sample repeat its last row): one vote for the word
        // then; the per-column plurality only where the rows differ
        int unanimous = 0;
        __match_all_sync(kFull, x[k], &unanimous);
        if (unanimous) { c[k] = x[k]; continue; }
#pragma unroll
        for (int j = 4 * k; j < 4 * k + 4; j++)
            if (j < n_cols) c[k] |= plurality_byte((x[k] >> (8 * (j & 3))) & 0xFFu, valid, lane) << (8 * (j & 3));
    }
    if (lane == 0) {
        if (sp < n_spans) *reinterpret_cast<uint4 *>(cons + sp * 16) = make_uint4(c[0], c[1], c[2], c[3]);
        else *reinterpret_cast<uint32_t *>(cons + sp * 16) = c[0];
    }
    if (sp >= n_spans) return;
    // second and third reference: the two most frequent patterns that differ from the consensus
    const bool differs = ((valid >> lane) & 1u) && ((x[0] ^ c[0]) | (x[1] ^ c[1]) | (x[2] ^ c[2]) | (x[3] ^ c[3]) | (x[4] ^ c[4])) != 0;
    const unsigned same = __match_any_sync(kFull, x[0]) & __match_any_sync(kFull, x[1]) & __match_any_sync(kFull, x[2]) &
                          __match_any_sync(kFull, x[3]) & __match_any_sync(kFull, x[4]) & valid;
    unsigned taken = 0;                                             // the sampled rows that carry an already chosen pattern
    uint32_t pat[2][5];
    bool real[2];
    for (int r = 0; r < 2; r++) {
        uint32_t key = differs && !((taken >> lane) & 1u) ? ((uint32_t)__popc(same) << 5) | (uint32_t)(31 - lane) : 0u;   // most frequent, lowest row on ties
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) key = max(key, __shfl_xor_sync(kFull, key, o));
        const int src = 31 - (int)(key & 31u);
#pragma unroll
        for (int k = 0; k < 5; k++) pat[r][k] = __shfl_sync(kFull, x[k], src);
        taken |= __shfl_sync(kFull, same, src);
        real[r] = key != 0;
    }
    // sampled rows of this span that equal none of its three references: what the histogram kernel will have to park
    const unsigned unmatched = __ballot_sync(kFull, differs && !((taken >> lane) & 1u));
    if (lane == 0) {
        if (unmatched && unmatched_total) atomicAdd(unmatched_total, (uint32_t)__popc(unmatched));
        // made-up patterns where the sample had none: distinct from the consensus and from each other
        if (!real[0]) {
#pragma unroll
            for (int k = 0; k < 5; k++) pat[0][k] = c[k];
            pat[0][0] ^= 0x80u;
        }
        if (!real[1]) {
            const bool clash = pat[0][0] == (c[0] ^ 0x8000u) && pat[0][1] == c[1] && pat[0][2] == c[2] && pat[0][3] == c[3] && pat[0][4] == c[4];
#pragma unroll
            for (int k = 0; k < 5; k++) pat[1][k] = c[k];
            pat[1][0] ^= real[0] && clash ? 0x80u : 0x8000u;
        }
        const RefPlanes refs = ref_planes(alt, n_spans);
        refs.a16[sp] = make_uint4(pat[0][0], pat[0][1], pat[0][2], pat[0][3]);
        refs.b16[sp] = make_uint4(pat[1][0], pat[1][1], pat[1][2], pat[1][3]);
        refs.tails[sp] = (pat[0][4] & 0xFFFFu) | (pat[1][4] << 16);
        refs.cnt_a[sp] = 0;        // rows equal to the pattern, counted by the histogram kernels, spent by the fix-up
        refs.cnt_b[sp] = 0;
    }
}

// After all row blocks: credit the consensus codon of every regular codon with the rows that never reached
// the decode path.  The consensus codon's code (and, for an 'other' codon, its character key) comes straight
// from the consensus row and the codon index map.
// Byte of the span's second (r = 0) or third (r = 1) reference at column q (relative to the first span): the pattern planes
// are byte rows over the columns; the two bytes a codon reaches into the next span come from the span's tails word.
__device__ __forceinline__ uint8_t ref_byte(const RefPlanes &refs, int r, int64_t sp, int64_t q) {
    if ((q >> 4) == sp) return reinterpret_cast<const uint8_t *>(r ? refs.b16 : refs.a16)[q];
    return (uint8_t)(refs.tails[sp] >> (16 * r + 8 * (int)(q & 15)));
}

__device__ __forceinline__ void fixup_codon(const uint8_t *__restrict__ cons, int64_t cons_shift, const RefPlanes &refs,
                                            const CodonMap *__restrict__ map, int64_t c, uint32_t n_rows, uint32_t *__restrict__ hist,
                                            uint32_t *__restrict__ other_min, uint32_t *__restrict__ other_max, int lane)
{
    // one warp per codon: the 66 bins are summed across lanes
    uint32_t *h = hist + c * kHistBins;
    const CodonMap m = map[c];
    auto credit = [&](uint8_t r0, uint8_t r1, uint8_t r2, uint32_t rows) {     // lane 0: rows copies of the codon r0 r1 r2
        const uint8_t n0 = norm_char(r0, m.minus), n1 = norm_char(r1, m.minus), n2 = norm_char(r2, m.minus);
        const uint8_t k0 = class_of(n0), k1 = class_of(n1), k2 = class_of(n2);
        uint32_t code;
        if (max(k0, max(k1, k2)) < 4) code = k0 * 16 + k1 * 4 + k2;
        else if (k0 == 4 || k1 == 4 || k2 == 4) code = kCodeUndef;
        else code = kCodeOther;
        h[code] += rows;
        if (code == kCodeOther) {
            const uint32_t key = ((uint32_t)n0 << 16) | ((uint32_t)n1 << 8) | n2;
            other_min[c] = min(other_min[c], key);
            other_max[c] = max(other_max[c], key);
        }
    };
    const int64_t q0 = m.p0 + cons_shift, q1 = m.p1 + cons_shift, q2 = m.p2 + cons_shift;     // columns relative to the first span
    const uint8_t c0 = cons[q0], c1 = cons[q1], c2 = cons[q2];
    // the rows that equalled the second or third reference of the span this codon starts in: one decode each, weighted
    const int64_t sp = min(q0, q2) >> 4;
    for (int r = 0; r < 2; r++) {
        const uint32_t n_alt = (r ? refs.cnt_b : refs.cnt_a)[sp];
        if (n_alt) {
            const uint8_t a0 = ref_byte(refs, r, sp, q0), a1 = ref_byte(refs, r, sp, q1), a2 = ref_byte(refs, r, sp, q2);
            if ((a0 != c0 || a1 != c1 || a2 != c2) && lane == 0) credit(a0, a1, a2, n_alt);
            __syncwarp();
        }
    }
    const uint32_t seen = warp_sum_u32(h[lane] + h[lane + 32] + (lane < kHistBins - 64 ? h[lane + 64] : 0u));
    const uint32_t rest = n_rows - seen;       // rows that equalled the consensus codon and were never recorded
    if (rest && lane == 0) credit(c0, c1, c2, rest);
}

// The same fix-up with every load issued up front and the credits applied to the warp's register copy of the bins
// (lane l holds bins l and l + 32, lanes 0 and 1 also bins 64 and 65): one chain of dependent loads instead of four.
// bins[0..2] on return: this lane's bins l, l + 32 and the 'other' bin, as k_stats wants them.
__device__ __forceinline__ void fixup_codon_regs(const uint8_t *__restrict__ cons, int64_t cons_shift, const RefPlanes &refs,
                                                 const CodonMap *__restrict__ map, int64_t c, uint32_t n_rows,
                                                 uint32_t *__restrict__ hist, uint32_t *__restrict__ other_min,
                                                 uint32_t *__restrict__ other_max, int lane, uint32_t *bins)
{
    uint32_t *h = hist + c * kHistBins;
    const CodonMap m = map[c];
    const int64_t q0 = m.p0 + cons_shift, q1 = m.p1 + cons_shift, q2 = m.p2 + cons_shift;     // columns relative to the first span
    const int64_t sp = min(q0, q2) >> 4;
    uint32_t h0 = h[lane], h1 = h[lane + 32], h2 = lane < kHistBins - 64 ? h[lane + 64] : 0u;
    const uint8_t c0 = cons[q0], c1 = cons[q1], c2 = cons[q2];
    const uint32_t n1 = refs.cnt_a[sp], n2 = refs.cnt_b[sp];
    const uint8_t x0 = ref_byte(refs, 0, sp, q0), x1 = ref_byte(refs, 0, sp, q1), x2 = ref_byte(refs, 0, sp, q2);
    const uint8_t y0 = ref_byte(refs, 1, sp, q0), y1 = ref_byte(refs, 1, sp, q1), y2 = ref_byte(refs, 1, sp, q2);
    uint32_t omin = 0xFFFFFFFFu, omax = 0;         // keys of 'other' codons credited here
    bool any_other = false;
    auto credit = [&](uint8_t r0, uint8_t r1, uint8_t r2, uint32_t rows) {     // every lane: rows copies of the codon r0 r1 r2
        const uint8_t k0n = norm_char(r0, m.minus), k1n = norm_char(r1, m.minus), k2n = norm_char(r2, m.minus);
        const uint8_t k0 = class_of(k0n), k1 = class_of(k1n), k2 = class_of(k2n);
        uint32_t code;
        if (max(k0, max(k1, k2)) < 4) code = k0 * 16 + k1 * 4 + k2;
        else if (k0 == 4 || k1 == 4 || k2 == 4) code = kCodeUndef;
        else code = kCodeOther;
        if (code == (uint32_t)lane) { h0 += rows; h[code] = h0; }
        if (code == (uint32_t)lane + 32) { h1 += rows; h[code] = h1; }
        if (code == (uint32_t)lane + 64) { h2 += rows; h[code] = h2; }
        if (code == kCodeOther) {
            const uint32_t key = ((uint32_t)k0n << 16) | ((uint32_t)k1n << 8) | k2n;
            omin = min(omin, key);
            omax = max(omax, key);
            any_other = true;
        }
    };
    if (n1 && (x0 != c0 || x1 != c1 || x2 != c2)) credit(x0, x1, x2, n1);
    if (n2 && (y0 != c0 || y1 != c1 || y2 != c2)) credit(y0, y1, y2, n2);
    const uint32_t seen = warp_sum_u32(h0 + h1 + h2);
    const uint32_t rest = n_rows - seen;           // rows that equalled the consensus codon and were never recorded
    if (rest) credit(c0, c1, c2, rest);
    if (any_other && lane == 0) {
        other_min[c] = min(other_min[c], omin);
        other_max[c] = max(other_max[c], omax);
    }
    bins[0] = h0;
    bins[1] = h1;
    bins[2] = __shfl_sync(kFull, h2, kCodeOther - 64);
}

__global__ void __launch_bounds__(256)
k_fused_fixup(const uint8_t *__restrict__ cons, int64_t cons_shift, const RefPlanes refs, const CodonMap *__restrict__ map,
              const uint8_t *__restrict__ regular, int64_t n_codons, uint32_t n_rows,
              uint32_t *__restrict__ hist, uint32_t *__restrict__ other_min, uint32_t *__restrict__ other_max)
{
    const int lane = threadIdx.x & 31;
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= n_codons || !regular[c]) return;
    fixup_codon(cons, cons_shift, refs, map, c, n_rows, hist, other_min, other_max, lane);
}

// irregular codons (split across CDS segments, or beyond the run capacity of a span) go
// through K1+K2 on a compact list; this adds their histograms into place
__global__ void k_scatter_irregular(const uint32_t *__restrict__ tmp_hist, const uint32_t *__restrict__ tmp_min,
                                    const uint32_t *__restrict__ tmp_max, const int32_t *__restrict__ idx, int64_t n_irr,
                                    uint32_t *__restrict__ hist, uint32_t *__restrict__ other_min, uint32_t *__restrict__ other_max)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_irr * kHistBins) return;
    const int64_t k = i / kHistBins;
    const int b = (int)(i % kHistBins);
    const int64_t c = idx[k];
    hist[c * kHistBins + b] = tmp_hist[i];
    if (b == 0) { other_min[c] = tmp_min[k]; other_max[c] = tmp_max[k]; }
}

// ---------------------------------------------------------------------------
// Per-site nucleotide counts (the "next" row f1: within_group_site_results.txt,
// snpgenie_within_group.pl:1203-1295).  For every alignment column the number of
// A, C, G, T after the reference's normalisation (uc, U->T; anything else is not
// counted).  Same plan as the fused histogram: a thread owns a 16-byte column span
// and walks the rows; rows equal to the consensus row cost one compare, differing
// bytes are classified and RED-added into sites[col][class] (class 4 = not ACGT),
// and k_site_fixup credits the consensus residue with (rows - recorded).
// Algorithmic traffic: 1 byte per (sequence, site) read + 20 B per site written.
// ---------------------------------------------------------------------------
constexpr int kSiteWarps = 8;
constexpr int kSiteRowsPerWarp = 32;

__global__ void __launch_bounds__(kSiteWarps * 32)
k_site_hist(const uint8_t *__restrict__ aln, int64_t pitch, int64_t n_rows, const uint8_t *__restrict__ cons, int64_t n_spans,
            int64_t n_cols, uint32_t *__restrict__ sites)
{
    __shared__ uint8_t lut[256];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    {
        const uint8_t k = class_of(norm_char((uint8_t)t, 0));
        lut[t] = k < 4 ? k : 4;
    }
    __syncthreads();
    const int64_t span = (int64_t)blockIdx.x * 32 + lane;
    if (span >= n_spans) return;
    const int64_t row_begin = ((int64_t)blockIdx.y * kSiteWarps + warp) * kSiteRowsPerWarp;
    const int n_my = (int)max((int64_t)0, min((int64_t)kSiteRowsPerWarp, n_rows - row_begin));
    const uint4 c = *reinterpret_cast<const uint4 *>(cons + span * 16);
    const uint32_t cw[4] = {c.x, c.y, c.z, c.w};
    const uint8_t *p = aln + span * 16 + row_begin * pitch;
    const int64_t col0 = span * 16;
    for (int i = 0; i < n_my; i += 4) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = ld_stream_v4(reinterpret_cast<const uint4 *>(p + (size_t)(i + u < n_my ? i + u : i) * pitch));
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (i + u >= n_my) break;
            const uint32_t vw[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int w = 0; w < 4; w++) {
                uint32_t d = vw[w] ^ cw[w];
                if (d == 0) continue;
                uint32_t m = (((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u;   // bit 7 of every differing byte
                while (m) {
                    const int bit = __ffs(m) - 1;
                    m &= m - 1;
                    const int64_t col = col0 + 4 * w + (bit >> 3);
                    if (col < n_cols) atomicAdd(&sites[col * 5 + lut[(vw[w] >> (bit - 7)) & 0xFF]], 1u);
                }
            }
        }
    }
}

__global__ void k_site_fixup(const uint8_t *__restrict__ cons, int64_t n_cols, uint32_t n_rows, uint32_t *__restrict__ sites)
{
    const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= n_cols) return;
    uint32_t *s = sites + col * 5;
    const uint32_t seen = s[0] + s[1] + s[2] + s[3] + s[4];
    const uint8_t k = class_of(norm_char(cons[col], 0));
    s[k < 4 ? k : 4] += n_rows - seen;
}

// transposed codes -> sequence-major block [n][C] for product columns [c0, c0+C)
__global__ void k_untranspose(const uint8_t *__restrict__ codes, int64_t n_pad, int64_t n, int64_t c0, int64_t C,
                              uint8_t *__restrict__ out)
{
    __shared__ uint8_t tile[32][33];
    const int64_t cb = (int64_t)blockIdx.x * 32, rb = (int64_t)blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    if (cb + ty < C && rb + tx < n) tile[ty][tx] = codes[(c0 + cb + ty) * n_pad + rb + tx];
    __syncthreads();
    if (rb + ty < n && cb + tx < C) out[(rb + ty) * C + cb + tx] = tile[tx][ty];
}

// ---------------------------------------------------------------------------
// K3: per-codon statistics from histograms (one warp per codon).
// within  (wg:611-744):  comparisons = n(n-1)/2, sites = h.s/n,
//                        diffs = (1/2) h'Th / comparisons   
// between (bg:680-947):  comparisons = na*nb, sites = (ha.s/na + hb.s/nb)/2,
//                        diffs = ha'T hb / comparisons
// Conserved codons take the table sites of the conserved codon and 0 diffs
// (wg:706-744; bg:834-947).  'other' codons (IUPAC...) are defined alleles with
// 0 sites and 0 diffs; two distinct 'other' strings make a column polymorphic
// (tracked by K1's min/max signature).
// ---------------------------------------------------------------------------
constexpr int kStatsWarps = 8;

__device__ __forceinline__ int first_defined_code(const uint8_t *col, int64_t n, int lane) {
    // lowest sequence index whose codon is defined (bg:853-890 scan order)
    for (int64_t r0 = 0; r0 < n; r0 += 32) {
        const int64_t r = r0 + lane;
        const int code = r < n ? col[r] : kCodeUndef;
        const unsigned hit = __ballot_sync(kFull, code != kCodeUndef && code != kCodePad);
        if (hit) return __shfl_sync(kFull, code, __ffs(hit) - 1);
    }
    return kCodeUndef;
}

// One codon, one warp.  s_bin/s_ca/s_cb: kStatsWarps x 64 entries of shared scratch, row `warp` is this warp's.
template <bool BETWEEN>
__device__ __forceinline__ void stats_codon(const GroupView &A, const GroupView &B, int64_t c, const double *__restrict__ tables,
                                            const StatsOut &out, const PeerOut &po, int lane, int warp, int (*s_bin)[64], double (*s_ca)[64],
                                            double (*s_cb)[64], const uint32_t *bins = nullptr)
{
    const double *sites = tables;            // [64][2]
    const double *TN = tables + 128;         // [64][64]
    const double *TS = tables + 128 + 4096;

    uint32_t a0, a1, ao;
    if (bins) { a0 = bins[0]; a1 = bins[1]; ao = bins[2]; }            // the caller holds them (k_fused_fixup_stats)
    else {
        const uint32_t *ha = A.hist + c * kHistBins;
        a0 = ha[lane]; a1 = ha[lane + 32]; ao = ha[kCodeOther];
    }
    uint32_t b0 = a0, b1 = a1, bo = ao;
    if (BETWEEN) {
        const uint32_t *hb = B.hist + c * kHistBins;
        b0 = hb[lane]; b1 = hb[lane + 32]; bo = hb[kCodeOther];
    }
    const uint32_t na_u = warp_sum_u32(a0 + a1) + ao;
    const uint32_t nb_u = BETWEEN ? warp_sum_u32(b0 + b1) + bo : na_u;
    const double na = (double)na_u, nb = (double)nb_u;

    const unsigned m0 = __ballot_sync(kFull, (a0 | b0) != 0), m1 = __ballot_sync(kFull, (a1 | b1) != 0);
    const int n_acgt = __popc(m0) + __popc(m1);
    int n_other = 0;
    if (ao | bo) {
        uint32_t lo = kOtherNone, hi = 0;
        if (ao) { lo = min(lo, A.other_min[c]); hi = max(hi, A.other_max[c]); }
        if (BETWEEN && bo) { lo = min(lo, B.other_min[c]); hi = max(hi, B.other_max[c]); }
        n_other = lo != hi ? 2 : 1;
    }
    const int alleles = n_acgt + n_other;
    const bool poly = BETWEEN ? (na_u > 0 && nb_u > 0 && alleles >= 2) : (alleles >= 2);

    double r_ns = 0, r_ss = 0, r_nd = 0, r_sd = 0;
    unsigned long long cmp = 0;
    int conserved = kCodeUndef;

    if (poly) {
        const double sn0 = sites[2 * lane], ss0 = sites[2 * lane + 1];
        const double sn1 = sites[2 * (lane + 32)], ss1 = sites[2 * (lane + 32) + 1];
        const double hsn_a = warp_sum((double)a0 * sn0 + (double)a1 * sn1);
        const double hss_a = warp_sum((double)a0 * ss0 + (double)a1 * ss1);
        // compact the occupied ACGT bins
        const unsigned lt = (1u << lane) - 1;
        if ((a0 | b0) != 0) { const int k = __popc(m0 & lt); s_bin[warp][k] = lane; s_ca[warp][k] = a0; s_cb[warp][k] = b0; }
        if ((a1 | b1) != 0) { const int k = __popc(m0) + __popc(m1 & lt); s_bin[warp][k] = lane + 32; s_ca[warp][k] = a1; s_cb[warp][k] = b1; }
        __syncwarp();
        double qn = 0, qs = 0;
        if (BETWEEN) {
            for (int i = 0; i < n_acgt; i++) {
                const double wi = s_ca[warp][i];
                if (wi == 0) continue;
                const int bi = s_bin[warp][i];
                for (int j = lane; j < n_acgt; j += 32) {
                    if (j == i) continue;
                    const double w = wi * s_cb[warp][j];
                    const int bj = s_bin[warp][j];
                    qn += w * TN[bi * 64 + bj];
                    qs += w * TS[bi * 64 + bj];
                }
            }
        } else {
            for (int i = 0; i < n_acgt; i++) {
                const double wi = s_ca[warp][i];
                const int bi = s_bin[warp][i];
                for (int j = i + 1 + lane; j < n_acgt; j += 32) {
                    const double w = wi * s_ca[warp][j];
                    const int bj = s_bin[warp][j];
                    qn += w * TN[bi * 64 + bj];
                    qs += w * TS[bi * 64 + bj];
                }
            }
        }
        qn = warp_sum(qn);
        qs = warp_sum(qs);
        if (BETWEEN) {
            const double hsn_b = warp_sum((double)b0 * sn0 + (double)b1 * sn1);
            const double hss_b = warp_sum((double)b0 * ss0 + (double)b1 * ss1);
            cmp = (unsigned long long)na_u * (unsigned long long)nb_u;
            r_ns = (hsn_a / na + hsn_b / nb) / 2;
            r_ss = (hss_a / na + hss_b / nb) / 2;
        } else {
            cmp = (unsigned long long)na_u * (unsigned long long)(na_u - 1) / 2;
            r_ns = hsn_a / na;
            r_ss = hss_a / na;
        }
        const double dc = (double)cmp;
        r_nd = qn / dc;
        r_sd = qs / dc;
    } else {
        // conserved: which codon?
        if (!BETWEEN || (na_u > 0 && nb_u > 0)) {
            if (n_acgt == 1) conserved = m0 ? __ffs(m0) - 1 : 32 + __ffs(m1) - 1;
            else if (n_other) conserved = kCodeOther;
        } else if (na_u > 0 || nb_u > 0) {
            // one group has no defined codon here; the reference takes the first
            // defined codon of the other group in sequence order (bg:853-890)
            const bool use_a = na_u > 0;
            const GroupView &G = use_a ? A : B;
            const uint8_t *codes = G.codes;
            if (alleles == 1) {
                conserved = n_acgt == 1 ? (m0 ? __ffs(m0) - 1 : 32 + __ffs(m1) - 1) : kCodeOther;
            } else if (codes) {
                conserved = first_defined_code(codes + c * G.n_pad, G.n, lane);
            } else if (G.first_def) {
                conserved = G.first_def[c];
            } else {
                conserved = n_acgt ? (m0 ? __ffs(m0) - 1 : 32 + __ffs(m1) - 1) : kCodeOther;
            }
        }
        if (conserved < 64) { r_ns = sites[2 * conserved]; r_ss = sites[2 * conserved + 1]; }
    }

    if (lane == 0) {
        if (out.poly) out.poly[c] = poly ? 1 : 0;
        if (out.ndef_a) out.ndef_a[c] = na_u;
        if (out.ndef_b) out.ndef_b[c] = nb_u;
        if (out.comparisons) out.comparisons[c] = cmp;
        if (out.conserved) out.conserved[c] = (uint8_t)(poly ? kCodeUndef : conserved);
        if (out.stats4) {
            double2 *o = reinterpret_cast<double2 *>(out.stats4 + 4 * c);
            o[0] = make_double2(r_ns, r_ss);
            o[1] = make_double2(r_nd, r_sd);
        }
    }
    // multi-GPU: the same values straight into every rank's exchange region (the all-gather of SURVEY.md 8e as the
    // tail of the producing kernel); lane r serves rank r
    if (po.world > 1 && lane < po.world) {
        const int64_t g = po.first + c;
        double2 *o = reinterpret_cast<double2 *>(po.stats4[lane] + 4 * g);
        o[0] = make_double2(r_ns, r_ss);
        o[1] = make_double2(r_nd, r_sd);
        po.cmp[lane][g] = cmp;
        po.ndef[lane][g] = na_u;
        po.poly[lane][g] = poly ? 1 : 0;
        po.cons[lane][g] = (uint8_t)(poly ? kCodeUndef : conserved);
    }
}

template <bool BETWEEN>
__global__ void __launch_bounds__(kStatsWarps * 32)
k_stats(GroupView A, GroupView B, int64_t n_codons, const double *__restrict__ tables, StatsOut out,
        const __grid_constant__ PeerOut po, const __grid_constant__ PeerSync ps)
{
    __shared__ int s_bin[kStatsWarps][64];
    __shared__ double s_ca[kStatsWarps][64];
    __shared__ double s_cb[kStatsWarps][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t c = (int64_t)blockIdx.x * kStatsWarps + warp;
    if (c < n_codons) stats_codon<BETWEEN>(A, B, c, tables, out, po, lane, warp, s_bin, s_ca, s_cb);
    peer_signal(ps, kStageStats, c < n_codons && lane < po.world);
}

// k_fused_fixup followed, codon by codon, by the within-group statistics (K3): one launch less for a whole job
// (snpg_within_job), the same arithmetic as the two kernels.  Only when every codon is regular.
__global__ void __launch_bounds__(kStatsWarps * 32)
k_fused_fixup_stats(const uint8_t *__restrict__ cons, int64_t cons_shift, const RefPlanes refs, const CodonMap *__restrict__ map,
                    int64_t n_codons, uint32_t n_rows, uint32_t *__restrict__ hist, uint32_t *__restrict__ other_min,
                    uint32_t *__restrict__ other_max, GroupView A, const double *__restrict__ tables, StatsOut out,
                    const __grid_constant__ PeerOut po, const __grid_constant__ PeerSync ps)
{
    static_assert(kStatsWarps * 32 == 256, "the fix-up addresses codons by 256-thread blocks");
    __shared__ int s_bin[kStatsWarps][64];
    __shared__ double s_ca[kStatsWarps][64];
    __shared__ double s_cb[kStatsWarps][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t c = (int64_t)blockIdx.x * kStatsWarps + warp;
    if (c < n_codons) {
        uint32_t bins[3];
        fixup_codon_regs(cons, cons_shift, refs, map, c, n_rows, hist, other_min, other_max, lane, bins);
        __syncwarp();                              // lane 0's 'other' keys are visible to the warp
        stats_codon<false>(A, A, c, tables, out, po, lane, warp, s_bin, s_ca, s_cb, bins);
    }
    peer_signal(ps, kStageStats, c < n_codons && lane < po.world);
}

__global__ void k_peer_signal(const __grid_constant__ PeerSync ps, int stage) { peer_signal(ps, stage); }
// for consumers that cannot wait themselves (copies, kernels that know nothing about peers): everything behind this
// kernel in its stream sees the stage's data
__global__ void k_peer_wait(const __grid_constant__ PeerSync ps, int stage) { peer_wait(ps, stage); }

// min-taxa codon filters of the between-group processor (bgp:412-467, :929-970): a codon that fails them
// contributes nothing to the sums, the windows or a bootstrap draw -- its row is zeroed in place.
__global__ void k_taxa_filter(double *__restrict__ stats4, const uint32_t *__restrict__ ndef_a, const uint32_t *__restrict__ ndef_b,
                              int64_t n_codons, uint32_t min_total, uint32_t min_group)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_codons) return;
    const uint32_t a = ndef_a[c], b = ndef_b[c];
    const bool keep = (min_total == 0 || a + b >= min_total) && (min_group == 0 || min(a, b) >= min_group);
    if (!keep) {
        double2 *o = reinterpret_cast<double2 *>(stats4 + 4 * c);
        o[0] = make_double2(0.0, 0.0);
        o[1] = make_double2(0.0, 0.0);
    }
}

// product sums (wg:690-693; bg:984-1005): fixed-shape tree, deterministic
constexpr int kSumThreads = 1024;
__global__ void __launch_bounds__(kSumThreads)
k_product_sums(const double *__restrict__ stats4, const int64_t *__restrict__ ofs, double *__restrict__ sums)
{
    __shared__ double red[4][kSumThreads];
    const int p = blockIdx.x, t = threadIdx.x;
    const int64_t c0 = ofs[p], c1 = ofs[p + 1];
    double s[4] = {0, 0, 0, 0};
    for (int64_t c = c0 + t; c < c1; c += kSumThreads) {
        const double2 x = reinterpret_cast<const double2 *>(stats4 + 4 * c)[0];
        const double2 y = reinterpret_cast<const double2 *>(stats4 + 4 * c)[1];
        s[0] += x.x; s[1] += x.y; s[2] += y.x; s[3] += y.y;
    }
    for (int q = 0; q < 4; q++) red[q][t] = s[q];
    __syncthreads();
    for (int h = kSumThreads / 2; h > 0; h >>= 1) {
        if (t < h) for (int q = 0; q < 4; q++) red[q][t] += red[q][t + h];
        __syncthreads();
    }
    if (t < 4) sums[4 * p + t] = red[t][0];
}

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter-based: any (replicate, draw) pair
// regenerates its numbers independently of launch geometry.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}

__global__ void k_philox_kat(uint4 ctr, uint2 key, uint32_t *out)
{
    const uint4 r = philox4x32_10(ctr, key);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

// ---------------------------------------------------------------------------
// Per-replicate values.  dN = Nd/Ns if Ns > 0, else the reference's '*' which its
// arithmetic reads as 0 (wg:942-961); Jukes-Cantor as bgp:1047-1071.
// which: 0 dN, 1 dS, 2 dN-dS, 3 JC dN, 4 JC dS, 5 JC dN-dS.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double rep_value(double ns, double ss, double nd, double sd, int which) {
    const double dn = ns > 0 ? nd / ns : 0.0;
    const double ds = ss > 0 ? sd / ss : 0.0;
    if (which == 0) return dn;
    if (which == 1) return ds;
    if (which == 2) return dn - ds;
    const double jn = -0.75 * log(1.0 - (4.0 / 3.0) * dn);
    const double js = -0.75 * log(1.0 - (4.0 / 3.0) * ds);
    if (which == 3) return jn;
    if (which == 4) return js;
    return (jn >= 0 && js >= 0) ? jn - js : 0.0;
}

// ---------------------------------------------------------------------------
// K4: whole-product bootstrap (wg:854-916; bgp:893-989).  One warp per
// replicate; lanes share the C draws.  A draw is int(rand(C+1)) (wg:885): slot 0
// is the reference's nonexistent codon 0 and contributes nothing.
// The product's table is staged in shared memory as two 16-byte arrays,
// (N_sites, S_sites) and (N_diffs, S_diffs); the sign bit of N_sites marks the
// codons that have differences at all, so a draw on a conserved codon (most of
// an alignment) costs one 16-byte gather instead of two.
// Outputs (either may be null): rep[p][b][4] replicate sums and vals[which][p][b], written by a dense
// block-level epilogue.
// ---------------------------------------------------------------------------
constexpr int kBootWarps = 32;
constexpr int kBootRepsPerWarp = 2;
constexpr int kBootSmemLimit = 200 * 1024;

// Philox counter of one quad of draws: x = quad index inside the replicate, y and 4 bits of z = replicate number
// (< 2^36), z also carries the item (product or window, < 2^24) and the KIND of call (product bootstrap, window
// bootstrap, between-group pair ...) in its top four bits, w = the ctx's call counter.  Calls of different kinds can
// therefore never share a stream, whatever the number of calls.
__device__ __forceinline__ uint4 boot_counter(uint32_t q, uint64_t b, uint32_t item, uint32_t kind, uint32_t call) {
    return make_uint4(q, (uint32_t)b, (kind << 28) | (((uint32_t)(b >> 32) & 0xFu) << 24) | (item & 0xFFFFFFu), call);
}

// One per-replicate value into vals[which][item][B] -- on several GPUs into every rank's exchange region (the gather
// of the replicate ranges is the tail of the kernel that draws them).
__device__ __forceinline__ void put_val(const BootArgs &a, int which, int item, int64_t bl, double v) {
    const size_t idx = ((size_t)which * a.n_items + item) * a.vals_stride + a.vals_b0 + bl;
    if (a.pv.world > 1) {
        for (int r = 0; r < a.pv.world; r++) a.pv.vals[r][idx] = v;
    } else {
        a.vals[idx] = v;
    }
}

// (A banked two-stage sampler -- class counts per shared-memory bank group first, then conflict-free draws inside
// the lane's class; exact, 49 M -> 30 M wavefronts -- was measured at 226 us against 212 us for the direct draws
// below: its extra instructions cost more than the conflicts.  So did a two-table form -- a 4-byte class word per
// draw that carries a conserved codon's sites in whole sixths, summed in integers, the other codons queued per
// warp and gathered 32 at a time from compact 32-byte rows: 274 us -- and unrolling the draw loop twice (223 us).
// See profiles/README.md and the git history.)
template <bool SMEM>
__global__ void __launch_bounds__(kBootWarps * 32)
k_bootstrap(const __grid_constant__ BootArgs a)
{
    extern __shared__ double2 s_tab[];
    const uint32_t call = a.call + (a.salt ? *a.salt : 0u);      // a replayed CUDA graph gets its per-call number from memory
    __shared__ double2 s_sums[kBootWarps * kBootRepsPerWarp][2];     // the block's replicate sums, for the dense epilogue
    __shared__ double s_red[SMEM ? kBootWarps * 32 : 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p = a.order ? a.order[blockIdx.y] : blockIdx.y;
    peer_wait(a.ps, kStageStats);
    const int64_t c0 = a.beg[p];
    const uint32_t C = (uint32_t)(a.end[p] - c0);
    const uint32_t s0 = (uint32_t)a.slot0, n_slots = C + s0;
    const double2 *G = reinterpret_cast<const double2 *>(a.stats4 + 4 * c0);
    // shared table rows 0..n_slots-1: with slot0 row 0 is the empty slot and row r is codon r - 1
    double2 *A = s_tab, *D = s_tab + (C + 1);
    if (SMEM) {
        if (threadIdx.x == 0) { A[0] = make_double2(0.0, 0.0); D[0] = make_double2(0.0, 0.0); }
        for (uint32_t i = threadIdx.x; i < C; i += blockDim.x) {
            double2 x = G[2 * i];
            const double2 d = G[2 * i + 1];
            if (d.x != 0.0 || d.y != 0.0) x.x = -x.x;        // flag: this codon has differences (-0.0 also carries it)
            A[i + s0] = x;
            D[i + s0] = d;
        }
        __syncthreads();
        if (a.prod_sums && blockIdx.x == 0) {
            // the product's sums (wg:690-693), from the staged table: the additions and the tree of k_product_sums
            double s[4] = {0, 0, 0, 0};
            for (uint32_t c = threadIdx.x; c < C; c += blockDim.x) {
                const double2 x = A[c + s0], y = D[c + s0];
                s[0] += fabs(x.x); s[1] += x.y; s[2] += y.x; s[3] += y.y;
            }
            for (int q = 0; q < 4; q++) {
                s_red[threadIdx.x] = s[q];
                __syncthreads();
                for (int h = kBootWarps * 32 / 2; h > 0; h >>= 1) {
                    if ((int)threadIdx.x < h) s_red[threadIdx.x] += s_red[threadIdx.x + h];
                    __syncthreads();
                }
                if (threadIdx.x == 0) a.prod_sums[4 * p + q] = s_red[0];
                __syncthreads();
            }
        }
    }
    const uint32_t nquad = (C + 3) >> 2, full_quads = C >> 2;
    for (int k = 0; k < kBootRepsPerWarp; k++) {
        const int64_t bl = ((int64_t)blockIdx.x * kBootWarps + warp) * kBootRepsPerWarp + k;
        if (bl >= a.b_count) continue;
        const uint64_t b = (uint64_t)(a.b_first + bl);
        double s0_ = 0, s1 = 0, s2 = 0, s3 = 0;
        for (uint32_t q = lane; q < nquad; q += 32) {
            const uint4 u = philox4x32_10(boot_counter(q, b, (uint32_t)p + a.item0, a.kind, call), a.key);
            const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
            const int nd = q < full_quads ? 4 : (int)(C & 3);
            uint32_t r[4];
            bool live[4];
#pragma unroll
            for (int j = 0; j < 4; j++) { live[j] = j < nd; r[j] = live[j] ? __umulhi(uu[j], n_slots) : 0u; }
            if (SMEM) {
                // all gathers of the quad are issued before any is consumed
                double2 x[4], d[4];
#pragma unroll
                for (int j = 0; j < 4; j++) x[j] = live[j] ? A[r[j]] : make_double2(0.0, 0.0);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    d[j] = make_double2(0.0, 0.0);
                    if (__double2hiint(x[j].x) < 0) d[j] = D[r[j]];
                }
                s0_ += (fabs(x[0].x) + fabs(x[1].x)) + (fabs(x[2].x) + fabs(x[3].x));
                s1 += (x[0].y + x[1].y) + (x[2].y + x[3].y);
                s2 += (d[0].x + d[1].x) + (d[2].x + d[3].x);
                s3 += (d[0].y + d[1].y) + (d[2].y + d[3].y);
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (live[j] && r[j] >= s0) {
                        const double2 x = __ldg(G + 2 * (size_t)(r[j] - s0));
                        const double2 d = __ldg(G + 2 * (size_t)(r[j] - s0) + 1);
                        s0_ += x.x; s1 += x.y; s2 += d.x; s3 += d.y;
                    }
                }
            }
        }
        s0_ = warp_sum(s0_); s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3);
        if (lane == 0) {
            s_sums[warp * kBootRepsPerWarp + k][0] = make_double2(s0_, s1);
            s_sums[warp * kBootRepsPerWarp + k][1] = make_double2(s2, s3);
        }
    }
    // dense epilogue: one thread per (replicate of the block, statistic) derives dN, dS, ... (divisions,
    // logarithms); each of 12 warps runs one statistic for 32 replicates.  Done by six lanes of every warp
    // right after its reduction, this was 44 % of the kernel's instructions.
    __syncthreads();
    constexpr int kReps = kBootWarps * kBootRepsPerWarp;
    const int i = threadIdx.x, r = i % kReps, which = i / kReps;
    const int64_t bl = (int64_t)blockIdx.x * kReps + r;
    const bool mine = which < a.n_which && bl < a.b_count;
    if (mine) {
        const double2 x = s_sums[r][0], y = s_sums[r][1];
        if (a.rep && which == 0) {
            double2 *o = reinterpret_cast<double2 *>(a.rep + 4 * ((size_t)p * a.b_count + bl));
            o[0] = x;
            o[1] = y;
        }
        if (a.vals || a.pv.world > 1) put_val(a, which, p, bl, rep_value(x.x, x.y, y.x, y.y, which));
    }
    peer_signal(a.ps, kStageVals, mine);
}

// ---------------------------------------------------------------------------
// K4, class-aggregated form (SNPG_BOOT_CLASSES; snpg_internal.h: BootClasses).  Two small kernels ahead of the
// weight-matrix kernel:
//   k_boot_classify  one block per item: the conserved codons (N_diffs = S_diffs = 0) are grouped by the bit patterns of
//                    (N_sites, S_sites) in a 64-entry shared-memory table; the kBootMaxClasses largest groups (class 0 =
//                    all-zero rows + the reference's empty slot, always kept) become classes, every other codon is
//                    copied, in order, to the front of the item's rows of perm.  Which codons end in which class does
//                    not depend on the order in which threads reach the table.
//   k_boot_split     one thread per (item, replicate): the C draws of a replicate fall into (general codons, class 0,
//                    class 1, ...) multinomially; the counts are drawn as a chain of conditional binomials and the
//                    classes' contribution to the replicate sums (count x the class's N_sites, S_sites) is added up.
// k_bootstrap_mma<.., true> then makes only the draws on the general codons.
// ---------------------------------------------------------------------------
constexpr int kBootKSplit = 8;                     // warps sharing the slots of one group of 8 replicates (one row per warp)
constexpr int kBootMaxRows = 8;                    // replicates a warp takes at a time when the product is short
// Several blocks of kMmaWarps warps share an SM (five for config 2's largest product, at 48 registers a thread): while
// some contract (FP64 tensor pipe) the others sample (integer pipes).  Measured: 32 warps x 1 block 178 us, 16 x 2 159,
// 8 x 4 (64 registers) 158, 8 x 5 (48 registers, spills only in the prologue) 151 -> 146.
constexpr int kMmaWarps = 8;
static_assert(kMmaWarps % kBootKSplit == 0, "a batch is kMmaWarps / 8 groups of 8 replicates, each contracted by 8 warps");
constexpr uint32_t kLogMmaWarps = kMmaWarps == 16 ? 4 : kMmaWarps == 8 ? 3 : 5;
static_assert((1u << kLogMmaWarps) == kMmaWarps, "kMmaWarps must be 8, 16 or 32");
static_assert(kBootMaxRows == 8, "mma_log_rows_per_warp adds up three comparisons");
// Shape of a k_bootstrap_mma block for an item of n_slots slots (k_boot_classify uses the same arithmetic to tell the kernel
// how many blocks of the item have replicates at all).  A warp takes m = 1 << lm replicates at a time: the largest power of two
// up to kBootMaxRows whose rows (rw words each) fit in the warp's row_words counters.
__device__ __forceinline__ uint32_t mma_row_words_of(uint32_t n_slots) { return ((n_slots + 3) >> 2) | 1; }
__device__ __forceinline__ uint32_t mma_log_rows_per_warp(uint32_t rw, uint32_t row_words) {
    return (2 * rw <= row_words ? 1u : 0u) + (4 * rw <= row_words ? 1u : 0u) + (8 * rw <= row_words ? 1u : 0u);
}
// replicates one block takes: (1 << lrows) per batch x n_batch batches (nb, split: the launch's arguments)
__device__ __forceinline__ uint32_t mma_reps_per_block(uint32_t n_slots, uint32_t row_words, uint32_t nb, uint32_t split) {
    const uint32_t lm = mma_log_rows_per_warp(mma_row_words_of(n_slots), row_words), m = 1u << lm;
    const uint32_t ls = lm == 0 ? split : 0u, lrows = kLogMmaWarps + lm - ls, n_batch = m >= nb ? 1u : nb >> lm;
    return (1u << lrows) * n_batch;
}

constexpr int kLogFactMax = 8192;                    // longest item of the class-aggregated form (table of log-factorials)
constexpr int kClsSlots = 64;
constexpr int kClsCtas = 8;                          // thread blocks (one cluster) per item
constexpr int kClassifyThreads = 1024;               // one codon per thread: items of up to kClsCtas * 1024 codons
constexpr unsigned long long kClsEmpty = ~0ull;
static_assert(kClsCtas * kClassifyThreads >= kLogFactMax, "the classes are used for items of up to kLogFactMax codons");

__device__ __forceinline__ uint32_t cls_hash(unsigned long long n_bits) {
    return (((uint32_t)(n_bits >> 32) * 0x9E3779B1u) ^ ((uint32_t)n_bits * 0x85EBCA6Bu)) >> 26;
}

// One cluster of kClsCtas blocks per item (a single block spends 35 us on config 2's longest product: 4,406 codons
// through four passes on one SM).  Block r holds the codons [1024 r, 1024 r + 1024), one per thread, and builds a table of
// its own keys; after one cluster barrier every block reads all eight tables through distributed shared memory and
// merges them into its own copy of the item's table, after a second one it adds up the eight blocks' member counts.
// (Cluster barriers cost about 2.5 us each here: a first version with one shared table and five barriers took 17 us.)
__global__ void __cluster_dims__(kClsCtas, 1, 1) __launch_bounds__(kClassifyThreads)
k_boot_classify(const __grid_constant__ BootArgs a)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ unsigned long long loc_n[kClsSlots], loc_s[kClsSlots];     // this block's keys (read by its peers)
    __shared__ unsigned long long key_n[kClsSlots], key_s[kClsSlots];     // the item's keys (every block its own copy)
    __shared__ uint32_t my_cnt[kClsSlots], cnt[kClsSlots];                // members among this block's codons (read by its peers) / in the item
    __shared__ int rank_of[kClsSlots];
    __shared__ uint32_t s_scan[kClassifyThreads / 32];
    __shared__ uint32_t overflow, s_classes, members[kClsCtas];
    const uint32_t r = cluster.block_rank();
    const int p = blockIdx.x / kClsCtas, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int64_t c0 = a.beg[p];
    const uint32_t C = (uint32_t)(a.end[p] - c0);
    const double2 *G = reinterpret_cast<const double2 *>(a.stats4 + 4 * c0);
    double2 *O = reinterpret_cast<double2 *>(a.perm + 4 * c0);
    const uint32_t zero_slot = cls_hash(0ull);
    if (t < kClsSlots) { loc_n[t] = kClsEmpty; loc_s[t] = kClsEmpty; key_n[t] = kClsEmpty; key_s[t] = kClsEmpty; my_cnt[t] = 0; cnt[t] = 0; rank_of[t] = -1; }
    if (t < kClsCtas) members[t] = 0;
    if (t == 0) overflow = 0;
    __syncthreads();
    if (t == 0) { key_n[zero_slot] = 0ull; key_s[zero_slot] = 0ull; }    // class 0 exists in every item: zero rows + the empty slot
    // this thread's codon
    const uint32_t i = r * kClassifyThreads + t;
    double2 x = make_double2(0.0, 0.0), d = make_double2(1.0, 1.0);
    if (i < C) { x = G[2 * i]; d = G[2 * i + 1]; }
    const unsigned long long nb = (unsigned long long)__double_as_longlong(x.x), sb = (unsigned long long)__double_as_longlong(x.y);
    const bool cons = i < C && d.x == 0.0 && d.y == 0.0 && nb != kClsEmpty && sb != kClsEmpty;
    // the table entry of a key in `keys`, or -1
    auto find = [&](const unsigned long long *keys, unsigned long long key) {
        uint32_t h = cls_hash(key);
        for (int probe = 0; probe < kClsSlots; probe++, h = (h + 1) & (kClsSlots - 1)) {
            const unsigned long long k = keys[h];
            if (k == key) return (int)h;
            if (k == kClsEmpty) return -1;
        }
        return -1;
    };
    auto insert = [&](unsigned long long *keys, unsigned long long key) {       // -1: the table is full
        uint32_t h = cls_hash(key);
        for (int probe = 0; probe < kClsSlots; probe++, h = (h + 1) & (kClsSlots - 1)) {
            const unsigned long long old = atomicCAS(&keys[h], kClsEmpty, key);
            if (old == kClsEmpty || old == key) return (int)h;
        }
        return -1;
    };
    // A warp's lanes mostly hold a handful of distinct rows: one lane per distinct key goes to the table.
    const unsigned grp = __match_any_sync(kFull, cons ? nb : kClsEmpty);
    const unsigned grp_s = __match_any_sync(kFull, cons ? sb : kClsEmpty);
    const bool first = cons && lane == __ffs(grp) - 1;
    int lh = -1;
    if (first) {                                         // pass 1: the N_sites keys of this block
        lh = insert(loc_n, nb);
        if (lh < 0) overflow = 1;
    }
    __syncthreads();
    if (cons && (first || (grp & ~grp_s) != 0)) {        // pass 2: per key the smallest S_sites pattern (a group that agrees: its first lane)
        if (!first) lh = find(loc_n, nb);
        if (lh >= 0) atomicMin(&loc_s[lh], sb);
    }
    cluster.sync();                                      // ---- every block's keys are in place
    unsigned long long peer_key = kClsEmpty;             // the table entry of a peer this thread merges: kept for the counts below
    uint32_t peer = 0, peer_slot = 0;
    if (t < kClsCtas * kClsSlots) {
        peer = t / kClsSlots; peer_slot = t % kClsSlots;
        peer_key = cluster.map_shared_rank(loc_n, peer)[peer_slot];
        if (peer_key != kClsEmpty) {
            const int h = insert(key_n, peer_key);
            if (h < 0) overflow = 1;
            else atomicMin(&key_s[h], cluster.map_shared_rank(loc_s, peer)[peer_slot]);
        }
        if (cluster.map_shared_rank(&overflow, peer)[0]) overflow = 1;
    }
    __syncthreads();
    int h = -1;                                          // pass 3: this block's members, one atomic per entry and warp
    if (cons) { h = find(key_n, nb); if (h >= 0 && key_s[h] != sb) h = -1; }
    {
        const unsigned same = __match_any_sync(kFull, h);
        if (h >= 0 && lane == __ffs(same) - 1) atomicAdd(&my_cnt[h], (uint32_t)__popc(same));
    }
    if (r == 0 && t == 0) atomicAdd(&my_cnt[zero_slot], (uint32_t)a.slot0);
    cluster.sync();                                      // ---- every block's member counts are in place
    // The item's counts: every block adds up all eight.  A peer's entries sit at the peer's own table positions, so they
    // are matched by key.
    uint32_t peer_cnt = 0;
    int peer_h = -1;
    if (t < kClsCtas * kClsSlots) {
        const unsigned long long k = cluster.map_shared_rank(key_n, peer)[peer_slot];
        peer_cnt = k != kClsEmpty ? cluster.map_shared_rank(my_cnt, peer)[peer_slot] : 0u;
        if (peer_cnt) { peer_h = find(key_n, k); if (peer_h >= 0) atomicAdd(&cnt[peer_h], peer_cnt); }
    }
    cluster.barrier_arrive();                            // (the last access to the peers' shared memory; waited for at the end)
    __syncthreads();
    // Classes in a canonical order: class 0, then by size, then by key.  A class pays for its binomial variate per
    // replicate only when it takes a fair share of the draws: small groups (and all groups of a short item) are drawn one
    // by one like the codons with differences.  (Every block ranks; block 0 writes the table.)
    const uint32_t min_count = max(32u, C >> a.cls_shift);
    if (t < kClsSlots) {
        const bool present = t == (int)zero_slot || (cnt[t] >= min_count && !overflow);
        int rank = 0;
        if (present && t != (int)zero_slot) {
            rank = 1;
            for (int e = 0; e < kClsSlots; e++) {
                if (e == t || e == (int)zero_slot || cnt[e] < min_count) continue;
                const bool before = cnt[e] > cnt[t] || (cnt[e] == cnt[t] && (key_n[e] < key_n[t] || (key_n[e] == key_n[t] && key_s[e] < key_s[t])));
                rank += before ? 1 : 0;
            }
        }
        if (present && rank < kBootMaxClasses) {
            rank_of[t] = rank;
            if (r == 0) {
                a.cls[p].count[rank] = cnt[t];
                a.cls[p].vn[rank] = __longlong_as_double((long long)key_n[t]);
                a.cls[p].vs[rank] = __longlong_as_double((long long)key_s[t]);
            }
        }
        const unsigned present_mask = __ballot_sync(kFull, present);     // (the two warps of the 64 entries)
        if (lane == 0) s_scan[warp] = __popc(present_mask);
    }
    __syncthreads();
    if (t == 0) { s_classes = min((uint32_t)kBootMaxClasses, s_scan[0] + s_scan[1]); if (r == 0) a.cls[p].n_classes = s_classes; }
    // codons of every block that are in a class (the empty slot is not a codon): the blocks before this one fix where its
    // general codons go
    if (peer_h >= 0 && rank_of[peer_h] >= 0) atomicAdd(&members[peer], peer_cnt);
    __syncthreads();
    // pass 4: the codons outside the classes, in order, to the front of perm
    const bool general = i < C && !(h >= 0 && rank_of[h] >= 0);
    const unsigned bal = __ballot_sync(kFull, general);
    if (lane == 0) s_scan[warp] = __popc(bal);
    __syncthreads();
    uint32_t before;
    {
        const uint32_t v = s_scan[lane];                 // kClassifyThreads / 32 = 32 warps: one count per lane
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(kFull, inc, o); if (lane >= o) inc += n; }
        before = __shfl_sync(kFull, inc - v, warp);
    }
    uint32_t ahead = 0, all = 0;
    for (uint32_t q = 0; q < kClsCtas; q++) {
        const uint32_t codons_q = min((uint32_t)kClassifyThreads, C - min(C, q * kClassifyThreads));
        const uint32_t v = codons_q - (members[q] - (q == 0 && rank_of[zero_slot] >= 0 ? (uint32_t)a.slot0 : 0u));
        ahead += q < r ? v : 0u;
        all += v;
    }
    if (general) {
        const uint32_t pos = ahead + before + __popc(bal & ((1u << lane) - 1));
        O[2 * pos] = x;
        O[2 * pos + 1] = d;
    }
    // the chain's constants: node 0 = the general codons against all slots, node 1 + k = class k against the slots still open
    if (r == 0 && t == 0) {
        a.cls[p].n_general = all;
        // blocks of k_bootstrap_mma that have replicates of this item (its grid is sized for the longest item: the others' last
        // blocks leave at once instead of working out their shape first)
        const uint32_t per = a.mma_row_words ? mma_reps_per_block(all, a.mma_row_words, a.mma_nb, a.mma_split) : 1u;
        a.cls[p].mma_blocks = a.mma_row_words ? (uint32_t)((a.b_count + per - 1) / per) : 0xFFFFFFFFu;
    }
    if (r == 0 && t <= (int)s_classes) {
        uint32_t mass = C + (uint32_t)a.slot0, num = all;
        if (t > 0) {
            mass -= all;
            for (int e = 0; e < kClsSlots; e++) {
                const int rk = rank_of[e];
                if (rk >= 0 && rk < t - 1) mass -= cnt[e];
                if (rk == t - 1) num = cnt[e];
            }
        }
        const bool flip = 2ull * num > mass;
        const double pr = mass ? (double)(flip ? mass - num : num) / (double)mass : 0.0;
        a.cls[p].node_p[t] = pr;
        a.cls[p].node_flip[t] = flip ? 1u : 0u;
        a.cls[p].node_lpq[t] = pr > 0.0 ? log(pr / (1.0 - pr)) : 0.0;
        a.cls[p].node_l1q[t] = log1p(-pr);
    }
    cluster.barrier_wait();                              // this block's shared memory stays until every peer has read it
}

// Uniform numbers of one replicate's class counts: Philox blocks whose quad index starts at kSplitQuad0, far above any
// quad of draws (a replicate has fewer than 2^31 draws).
constexpr uint32_t kSplitQuad0 = 0xC0000000u;
struct SplitRng {
    uint4 ctr; uint2 key; uint4 buf; int have;
    __device__ __forceinline__ uint32_t next() {
        if (have == 0) { buf = philox4x32_10(ctr, key); ctr.x++; have = 4; }
        const uint32_t r = have == 4 ? buf.x : have == 3 ? buf.y : have == 2 ? buf.z : buf.w;
        have--;
        return r;
    }
    __device__ __forceinline__ double uniform() { return ((double)next() + 0.5) * 0x1p-32; }        // in (0, 1)
};

// log(k!) for k <= kLogFactMax, filled once per device (launch_bootstrap) with lgamma
__device__ double g_logfact[kLogFactMax + 1];
__global__ void k_fill_logfact() {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k <= kLogFactMax) g_logfact[k] = lgamma((double)k + 1.0);
}

// Binomial(n, p) variate for p <= 1/2 with lpq = log(p / (1 - p)), l1q = log(1 - p); n <= kLogFactMax.  Mean below 10:
// inversion of the cumulative distribution, term by term from P(0) = (1 - p)^n.  Otherwise BTRS (Hoermann, "The generation
// of binomial random variates", 1993): transformed rejection with a squeeze, about 1.15 pairs of uniforms per variate;
// the exact test compares with the logarithm of the probability ratio f(k) / f(mode) from the table of log-factorials.
// Both are exact up to the resolution of the uniforms and of double arithmetic.
__device__ __forceinline__ uint32_t binomial_variate(SplitRng &rng, uint32_t n, double p, double lpq, double l1q) {
    if (n == 0 || p <= 0.0) return 0;
    const double dn = (double)n;
    uint32_t k;
    if (dn * p < 10.0) {
        const double s = exp(lpq), a = (dn + 1.0) * s, p0 = exp(dn * l1q);
        double u = rng.uniform(), px = p0;
        k = 0;
        while (u > px) {
            u -= px;
            k++;
            px *= a / (double)k - s;
            if (k > n || px <= 0.0) { k = 0; px = p0; u = rng.uniform(); }     // rounding left u past the last term: again
        }
    } else {
        const double q = 1.0 - p, sd = sqrt(dn * p * q), b = 1.15 + 2.53 * sd, a = -0.0873 + 0.0248 * b + 0.01 * p, c = dn * p + 0.5;
        const double rb = 1.0 / b, v_r = 0.92 - 4.2 * rb, alpha = (2.83 + 5.1 * rb) * sd;
        const uint32_t m = (uint32_t)((dn + 1.0) * p);
        const double h = g_logfact[m] + g_logfact[n - m];
        for (;;) {
            const double u = rng.uniform() - 0.5;
            const double v = rng.uniform();
            const double us = 0.5 - fabs(u), kd = floor((2.0 * a / us + b) * u + c);
            if (kd < 0.0 || kd > dn) continue;
            k = (uint32_t)kd;
            if (us >= 0.07 && v <= v_r) break;
            if (log(v * alpha / (a / (us * us) + b)) <= h - g_logfact[k] - g_logfact[n - k] + (kd - (double)m) * lpq) break;
        }
    }
    return k;
}

constexpr int kSplitThreads = 64;
__global__ void __launch_bounds__(kSplitThreads)
k_boot_split(const __grid_constant__ BootArgs a)
{
    __shared__ BootClasses s_cls;
    const uint32_t call = a.call + (a.salt ? *a.salt : 0u);
    const int p = blockIdx.y;
    for (uint32_t i = threadIdx.x; i < sizeof(BootClasses) / 4; i += kSplitThreads)
        reinterpret_cast<uint32_t *>(&s_cls)[i] = reinterpret_cast<const uint32_t *>(a.cls + p)[i];
    __syncthreads();
    const int64_t bl = (int64_t)blockIdx.x * kSplitThreads + threadIdx.x;
    if (bl >= a.b_count) return;
    const uint64_t b = (uint64_t)(a.b_first + bl);
    const uint32_t C = (uint32_t)(a.end[p] - a.beg[p]);
    SplitRng rng;
    rng.ctr = boot_counter(kSplitQuad0, b, (uint32_t)p + a.item0, a.kind, call);
    rng.key = a.key;
    rng.have = 0;
    auto node = [&](uint32_t n, int i) {
        const uint32_t k = binomial_variate(rng, n, s_cls.node_p[i], s_cls.node_lpq[i], s_cls.node_l1q[i]);
        return s_cls.node_flip[i] ? n - k : k;
    };
    uint32_t n = C;
    const uint32_t m_general = node(n, 0);
    n -= m_general;
    double sn = 0.0, ss = 0.0;
    const uint32_t Q = s_cls.n_classes;
    for (uint32_t k = 0; k < Q; k++) {
        const uint32_t hits = k + 1 == Q ? n : node(n, 1 + (int)k);
        n -= hits;
        sn += (double)hits * s_cls.vn[k];
        ss += (double)hits * s_cls.vs[k];
    }
    const size_t at = (size_t)p * a.b_count + bl;
    a.split_m[at] = m_general;
    a.split_ns[at] = make_double2(sn, ss);
}

// K4, weight-matrix form (the default when a warp's counts fit in shared memory): the contraction
// north_star describes.  The shared-memory pipe bounds the direct form above -- a 16-byte gather at 32 random
// addresses costs ~11.5 wavefronts, and every draw pays one or two.  Here a draw only bumps a 1-byte counter
// of its slot (a 4-byte shared-memory atomic at a random word: ~3.4 wavefronts per 32 draws): a warp turns its
// replicate into the row K[slot] of the multinomial weight matrix.  The block then contracts its 32 rows with
// the product's table, sums[rep][0..3] = sum over slots of K[rep][slot] * stats4[slot], as FP64 tensor-core tiles
// (mma.m8n8k4: A = 8 replicates x 4 slots of counts, B = 4 slots x the 4 statistics read straight from global
// memory/L2, 8 warps splitting the slots of each group of 8 replicates, partials summed in a fixed order).
// Same Philox stream, draw for draw, as the direct form; replicate sums agree to rounding.  A counter would wrap at
// 256 draws of one slot in one replicate: impossible up to C = 255 and of probability < 1e-400 beyond
// (each of the C draws picks the slot with probability 1/(C+1)).
// Items are products (beg = ofs, end = ofs + 1; slot0 = 1: the reference's whole-product bootstrap draws
// int(rand(C+1)), wg:885, and its codon 0 does not exist) or sliding windows (overlapping ranges, slot0 = 0:
// int(rand(W)) + first_codon, wgp:406).  On several GPUs the sampling of the first batch runs before the block waits
// for the peers' per-codon statistics -- only the contraction reads them.

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// CLS (SNPG_BOOT_CLASSES): the slots are the item's general codons (the first rows of perm, no empty slot), a replicate
// makes split_m draws on them and the classes' share of its sums comes from split_ns.
template <uint32_t split, bool CLS>
__global__ void __launch_bounds__(kMmaWarps * 32, 5)
k_bootstrap_mma(const __grid_constant__ BootArgs a, uint32_t row_words, uint32_t nb)
{
    // nb (1 or 2): replicates a warp takes one after the other when it takes one at a time; split (0 or 1): log2 of the
    // warps that share one replicate of a long item.  Both only shape the blocks -- few replicates per GPU (a rank of
    // eight) want many short blocks -- and never what a replicate draws.
    extern __shared__ uint32_t s_cnt[];            // [kMmaWarps][row_words]: 1-byte counters, four to a word
    const uint32_t call = a.call + (a.salt ? *a.salt : 0u);      // a replayed CUDA graph gets its per-call number from memory
    __shared__ double2 s_sums[kMmaWarps * kBootMaxRows][2];         // a batch's replicate sums, for the dense epilogue
    __shared__ double s_part[kBootKSplit][kMmaWarps][4];            // per K-split partial sums of a batch; also the reduction scratch
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p = a.order ? a.order[blockIdx.y] : blockIdx.y;
    if (CLS && blockIdx.x >= a.cls[p].mma_blocks) {      // no replicates of this item left for this block (k_boot_classify; block 0 always has some)
        peer_signal(a.ps, kStageVals, false);
        return;
    }
    const int64_t c0 = a.beg[p];
    const uint32_t C = (uint32_t)(a.end[p] - c0);
    const uint32_t s0 = CLS ? 0u : (uint32_t)a.slot0, n_slots = CLS ? a.cls[p].n_general : C + s0;
    const double *G = (CLS ? a.perm : a.stats4) + 4 * c0;
    const uint32_t nslot4 = (n_slots + 3) >> 2;    // groups of 4 slots
    // A short product does not fill a warp's counters: its warps then take m = 2, 4 or 8 replicates at a time (32 / m
    // lanes and one row of rw words each), so that the fixed cost of a batch is spread over 32 m replicates.
    const uint32_t rw = nslot4 | 1;                // words per row: odd, so the eight rows of a tile fall in different banks
    const uint32_t lm = mma_log_rows_per_warp(rw, row_words);       // m = 1 << lm: every count below is a power of two, divisions are shifts
    const uint32_t m = 1u << lm;
    constexpr uint32_t kLogWarps = kLogMmaWarps;
    const uint32_t ls = lm == 0 ? split : 0u;      // 1 << ls warps draw one replicate together (its row takes their REDs)
    const uint32_t lrows = kLogWarps + lm - ls;    // log2 of the replicates per batch
    const uint32_t rows = 1u << lrows;             // replicates per batch
    const uint32_t n_batch = m >= nb ? 1 : nb >> lm;
    const int64_t block_first = (int64_t)blockIdx.x * rows * n_batch;
    const bool active = block_first < a.b_count;   // the grid is sized for m = 1
    const uint32_t lanes_per = (32u >> lm) << ls, sub = lm ? lane >> (5 - lm) : 0u;
    const uint32_t sl = (lane & ((32u >> lm) - 1)) + ((warp & ((1u << ls) - 1)) << 5);
    const uint32_t my_row = ls ? warp >> ls : warp * m + sub;        // the replicate of the batch this lane draws for
    uint32_t *warp_rows = s_cnt + (size_t)warp * m * rw;
    const uint32_t mine_s = (uint32_t)__cvta_generic_to_shared(s_cnt + (size_t)my_row * rw);
    double *part = &s_part[0][0][0];               // [ksplit][rows][4], ksplit * rows <= 8 * kMmaWarps
    bool wrote = false;
    for (uint32_t k = 0; active && k < n_batch; k++) {
        const int64_t batch_first = block_first + (int64_t)k * rows;
        // --- sampling: one replicate per group of 32 / m lanes (or per 1 << ls warps) -> its row of counts
        if (ls) {
            for (uint32_t i = threadIdx.x; i < rows * rw; i += blockDim.x) s_cnt[i] = 0;
            __syncthreads();                       // a row is bumped by several warps
        } else {
            for (uint32_t i = lane; i < m * rw; i += 32) warp_rows[i] = 0;
            __syncwarp();
        }
        const int64_t bl = batch_first + my_row;
        if (bl < a.b_count) {
            const uint64_t b = (uint64_t)(a.b_first + bl);
            const uint32_t n_draws = CLS ? a.split_m[(size_t)p * a.b_count + bl] : C;
            const uint32_t nquad = (n_draws + 3) >> 2, full_quads = n_draws >> 2;     // the partial quad, if any, is quad full_quads
            auto bump = [&](uint32_t u) {          // one draw: slot = floor(u * n_slots / 2^32); its byte of the row goes up by one
                const uint32_t r = __umulhi(u, n_slots);
                // 1 << 8 * (r & 3): a funnel shift takes its count modulo 32, so r << 3 needs no mask
                asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(mine_s + (r & ~3u)), "r"(__funnelshift_l(0u, 1u, r << 3)) : "memory");
            };
            uint32_t q = sl;
            for (; q < full_quads; q += lanes_per) {
                const uint4 u = philox4x32_10(boot_counter(q, b, (uint32_t)p + a.item0, a.kind, call), a.key);
                bump(u.x); bump(u.y); bump(u.z); bump(u.w);
            }
            if (q < nquad) {                       // the last, partial quad (one lane)
                const uint4 u = philox4x32_10(boot_counter(q, b, (uint32_t)p + a.item0, a.kind, call), a.key);
                const uint32_t nd = n_draws & 3;
                bump(u.x);
                if (nd > 1) bump(u.y);
                if (nd > 2) bump(u.z);
            }
        }
        if (k == 0) peer_wait(a.ps, kStageStats);  // the table below comes from every rank's codon range
        __syncthreads();
        // --- contraction: group gq of 8 replicates x the slots of K-split ks, 4 slots per tensor-core tile
        {
            const uint32_t lks = 3 - lm, ksplit = 1u << lks;         // kBootKSplit / m warps per group of 8 replicates
            const uint32_t gq = warp >> lks, ks = warp & (ksplit - 1);
            const uint32_t per = (nslot4 + ksplit - 1) >> lks;
            const uint32_t t0 = min(nslot4, ks * per), t1 = min(nslot4, t0 + per);
            const uint32_t row = gq * 8 + (lane >> 2);
            // A: row = replicate lane/4 (fewer than 8 replicates in the batch: the tile's other rows repeat them and are dropped)
            const uint32_t *cnt_row = s_cnt + (size_t)(row & (rows - 1)) * rw;
            const uint32_t col = lane & 3;                           //    column = slot 4t + lane%4
            const uint32_t stat = lane >> 2;                         // B: column lane/4 = statistic (columns 4-7: padding)
            const uint32_t sel = 0x4440u + col;                      // byte col of a word, zero-extended
            // + 16 t: stats4[slot 4t + col][stat].  Columns 4-7 of B only feed columns 4-7 of D, which nobody reads: their
            // lanes load the same (finite) table entries as columns 0-3 instead of a predicated zero
            const double *gb = G + 4 * ((int64_t)col - (int64_t)s0) + (stat & 3);
            double acc0 = 0.0, acc1 = 0.0;
            auto tile = [&](uint32_t t, bool guard) {
                const double av = (double)__byte_perm(cnt_row[t], 0u, sel);
                double bv = 0.0;
                if (!guard || (4 * t + col >= s0 && 4 * t + col < n_slots)) bv = __ldg(gb + 16 * (size_t)t);
                dmma_m8n8k4(acc0, acc1, av, bv);
            };
            uint32_t t = t0;
            if (t == 0 && t < t1) tile(t++, true);                  // may hold slot 0, the empty one
            const uint32_t t_inner = min(t1, n_slots / 4);          // tiles below hold table rows only
#pragma unroll 4
            for (; t < t_inner; t++) tile(t, false);
            for (; t < t1; t++) tile(t, true);                      // may reach past the last slot
            // D: lane holds row lane/4, columns 2*(lane%4), +1; statistics are columns 0..3
            if (col < 2 && row < rows) {
                double *o = part + (((ks << lrows) + row) << 2) + 2 * col;
                o[0] = acc0;
                o[1] = acc1;
            }
        }
        __syncthreads();
        if (threadIdx.x < rows * 4) {              // (replicate of the batch, statistic): K-split partials in a fixed order
            const uint32_t r = threadIdx.x >> 2, q = threadIdx.x & 3, ksplit = kBootKSplit >> lm;
            double v = 0.0;
            for (uint32_t ks = 0; ks < ksplit; ks++) v += part[(((ks << lrows) + r) << 2) + q];
            if (CLS && q < 2 && batch_first + r < a.b_count)      // the classes' share of N_sites, S_sites
                v += reinterpret_cast<const double *>(a.split_ns + (size_t)p * a.b_count + batch_first + r)[q];
            reinterpret_cast<double *>(&s_sums[r][0])[q] = v;
        }
        __syncthreads();
        // dense epilogue: one thread per (replicate of the batch, statistic), as in k_bootstrap
        for (uint32_t i = threadIdx.x; i < rows * (uint32_t)a.n_which; i += blockDim.x) {
            const uint32_t r = i & (rows - 1), which = i >> lrows;
            const int64_t bl2 = batch_first + r;
            if (bl2 < a.b_count) {
                const double2 x = s_sums[r][0], y = s_sums[r][1];
                if (a.rep && which == 0) {
                    double2 *o = reinterpret_cast<double2 *>(a.rep + 4 * ((size_t)p * a.b_count + bl2));
                    o[0] = x;
                    o[1] = y;
                }
                if (a.vals || a.pv.world > 1) put_val(a, (int)which, p, bl2, rep_value(x.x, x.y, y.x, y.y, (int)which));
                wrote = true;
            }
        }
        // (the next batch's first barrier comes after its sampling: s_sums and part are not written before it)
    }
    if (a.prod_sums && blockIdx.x == 0) {
        // the product's sums (wg:690-693) with the additions and the tree of k_product_sums, bit for bit: its
        // kSumThreads partial sums, held kSumThreads / blockDim.x to a thread, and the same halving tree
        constexpr int kV = kSumThreads / (kMmaWarps * 32);          // partial sums of k_product_sums' threads held by one thread
        static_assert(kV >= 1 && (kV & (kV - 1)) == 0 && kV * kMmaWarps * 32 == kSumThreads, "block size must divide kSumThreads");
        __syncthreads();
        double *red = &s_part[0][0][0];            // kBootKSplit * kMmaWarps * 4 = kMmaWarps * 32 doubles: one per thread
        const double2 *G2 = reinterpret_cast<const double2 *>(a.stats4 + 4 * c0);
        double s[kV][4];
        for (int v = 0; v < kV; v++) {
            for (int q = 0; q < 4; q++) s[v][q] = 0.0;
            for (uint32_t c = threadIdx.x + v * (kMmaWarps * 32); c < C; c += kSumThreads) {
                const double2 x = G2[2 * c], y = G2[2 * c + 1];
                s[v][0] += x.x; s[v][1] += x.y; s[v][2] += y.x; s[v][3] += y.y;
            }
        }
        for (int h = kV / 2; h > 0; h >>= 1)       // the tree's first levels: t += t + h * blockDim.x
            for (int v = 0; v < h; v++)
                for (int q = 0; q < 4; q++) s[v][q] += s[v + h][q];
        for (int q = 0; q < 4; q++) {
            red[threadIdx.x] = s[0][q];
            __syncthreads();
            for (int h = kMmaWarps * 32 / 2; h > 0; h >>= 1) {
                if ((int)threadIdx.x < h) red[threadIdx.x] += red[threadIdx.x + h];
                __syncthreads();
            }
            if (threadIdx.x == 0) a.prod_sums[4 * p + q] = red[0];
            __syncthreads();
        }
    }
    peer_signal(a.ps, kStageVals, wrote);
}

// ---------------------------------------------------------------------------
// K6: SD over replicates in the reference's two-pass form with n-1
// (wg:1351-1365).  One block per (item, statistic) over rep[item][B][4]; the
// statistic of a replicate is derived on the fly (rep_value); se[item][6].
// ---------------------------------------------------------------------------
constexpr int kMomentThreads = 1024;
constexpr int kMomentCached = 10;        // values a thread keeps between the two passes
__device__ __forceinline__ double block_sum(double v, double *red) {     // fixed shape (butterflies): deterministic
    const int t = threadIdx.x;
    v = warp_sum(v);
    __syncthreads();
    if ((t & 31) == 0) red[t >> 5] = v;
    __syncthreads();
    return warp_sum(red[t & 31]);          // kMomentThreads / 32 == 32 partial sums, every warp adds them the same way
}
static_assert(kMomentThreads == 1024, "block_sum adds exactly 32 warp sums");

// CACHED (the launchers take it for B > kMomentThreads): see below; the other form keeps 30 registers, two blocks per SM -- what
// the window bootstraps' hundreds of small (item, statistic) blocks want.
template <bool FROM_REP, bool CACHED>
__global__ void __launch_bounds__(kMomentThreads)
k_moments(const double *__restrict__ src, int64_t n_items, int64_t B, double *__restrict__ se, uint32_t *__restrict__ undefined,
          const __grid_constant__ PeerSync ps)
{
    // FROM_REP: src = rep[item][B][4], the statistic is derived on the fly; else src = vals[which][item][B]
    __shared__ double red[kMomentThreads];
    const int p = blockIdx.x, which = blockIdx.y, t = threadIdx.x;
    peer_wait(ps, kStageVals);             // several GPUs: every rank's replicate range has arrived
    const double2 *R = reinterpret_cast<const double2 *>(src + 4 * (size_t)p * B);
    const double *V = src + ((size_t)which * n_items + p) * B;
    auto value = [&](int64_t b) {
        if (!FROM_REP) return V[b];
        const double2 x = R[2 * b], y = R[2 * b + 1];
        return rep_value(x.x, x.y, y.x, y.y, which);
    };
    double s = 0;
    int differs = 0;                       // replicates that are all identical have SD exactly 0, whatever
    int bad = 0;                           // replicates whose value is undefined (Jukes-Cantor of d >= 3/4: the reference dies, bgp:1047)
    const double v0 = value(0);            // rounding the mean of B equal numbers picks up
    double mean, ss = 0;
    if (CACHED) {
        // A thread's first kMomentCached values stay in registers for the second pass (B = 10,000: all ten of them): their loads
        // go out together and only once -- the kernel is the last of a job, all of it latency.  Same additions in the same order.
        double c[kMomentCached];
#pragma unroll
        for (int j = 0; j < kMomentCached; j++) {
            const int64_t b = t + (int64_t)j * kMomentThreads;
            c[j] = b < B ? value(b) : 0.0;
        }
#pragma unroll
        for (int j = 0; j < kMomentCached; j++)
            if (t + (int64_t)j * kMomentThreads < B) { const double v = c[j]; s += v; differs |= (v != v0); bad += !isfinite(v); }
        for (int64_t b = t + (int64_t)kMomentCached * kMomentThreads; b < B; b += kMomentThreads) { const double v = value(b); s += v; differs |= (v != v0); bad += !isfinite(v); }
        mean = block_sum(s, red) / (double)B;
#pragma unroll
        for (int j = 0; j < kMomentCached; j++)
            if (t + (int64_t)j * kMomentThreads < B) { const double d = c[j] - mean; ss += d * d; }
        for (int64_t b = t + (int64_t)kMomentCached * kMomentThreads; b < B; b += kMomentThreads) { const double d = value(b) - mean; ss += d * d; }
    } else {
        for (int64_t b = t; b < B; b += kMomentThreads) { const double v = value(b); s += v; differs |= (v != v0); bad += !isfinite(v); }
        mean = block_sum(s, red) / (double)B;
        for (int64_t b = t; b < B; b += kMomentThreads) { const double d = value(b) - mean; ss += d * d; }
    }
    const double tot = block_sum(ss, red);
    const int any = __syncthreads_or(differs);
    const int n_bad = __syncthreads_count(bad) ? (int)block_sum((double)bad, red) : 0;
    if (t == 0) {
        se[6 * p + which] = any ? sqrt(tot / (double)(B - 1)) : 0.0;
        if (undefined) undefined[6 * p + which] = (uint32_t)n_bad;
    }
}

// ---------------------------------------------------------------------------
// Row f2: the per-window statistics of SNPGenie_sliding_windows.R (:241-402) from the replicate sums of the window
// bootstrap.  One block per window.  R's arithmetic is kept as it is: dN = sum(diffs) / sum(sites) with IEEE results
// (0/0 = NaN, x/0 = Inf -- no '*' -> 0 coercion here), optionally Jukes-Cantor (:326-361), sd() with n - 1 (two passes),
// Z = estimate / SE, P = 2 * pnorm(-|Z|) (:303-312), the counts of replicates with dN - dS >, ==, < 0 and the achieved
// significance levels (:314-319; a NaN replicate makes R's sums NA: NaN here).  The script bootstraps each of its four
// statistics with its own boot() call; here the four are derived from ONE set of replicates (the same marginal
// distributions; nothing downstream uses their joint distribution).
// rep[window][B][4] = replicate sums (N_sites, S_sites, N_diffs, S_diffs); win_sums[window][4]; num / den: 0 = N, 1 = S.
// out[window][16] = dN, dS, dN/dS, dN-dS, SE(dN), SE(dS), SE(dN/dS), P(dN/dS), SE(dN-dS), P(dN-dS), gt, eq, lt,
// ASL(dN>dS), ASL(dN<dS), ASL two-sided (:492-500).
// ---------------------------------------------------------------------------
__device__ __forceinline__ double jc_distance(double d) { return -0.75 * log(1.0 - (4.0 / 3.0) * d); }
__device__ __forceinline__ double two_sided_p(double z) { return erfc(fabs(z) * 0.70710678118654752440); }   // 2 * pnorm(-|z|)

__global__ void __launch_bounds__(kMomentThreads)
k_window_rstats(const double *__restrict__ rep, const double *__restrict__ win_sums, int64_t B, int num, int den, int jc,
                double *__restrict__ out)
{
    __shared__ double red[kMomentThreads];
    const int w = blockIdx.x, t = threadIdx.x;
    const double *R = rep + 4 * (size_t)w * B;
    auto values = [&](const double *x, double v[4]) {       // dN, dS, dN/dS, dN-dS of one set of four sums
        double dn = x[2 + num] / x[num], ds = x[2 + den] / x[den];
        if (jc) { dn = jc_distance(dn); ds = jc_distance(ds); }
        v[0] = dn; v[1] = ds; v[2] = dn / ds; v[3] = dn - ds;
    };
    double first[4], s[4] = {0, 0, 0, 0}, gt = 0, eq = 0, lt = 0, nan_d = 0;
    values(R, first);
    int differs = 0;                                         // bit q: some replicate's statistic q differs from the first replicate's
    for (int64_t b = t; b < B; b += kMomentThreads) {
        double v[4];
        values(R + 4 * b, v);
#pragma unroll
        for (int q = 0; q < 4; q++) { s[q] += v[q]; differs |= (v[q] != first[q] && !(v[q] != v[q] && first[q] != first[q])) << q; }
        gt += v[3] > 0; eq += v[3] == 0; lt += v[3] < 0; nan_d += v[3] != v[3];
    }
    double mean[4];
#pragma unroll
    for (int q = 0; q < 4; q++) mean[q] = block_sum(s[q], red) / (double)B;
    double ss[4] = {0, 0, 0, 0};
    for (int64_t b = t; b < B; b += kMomentThreads) {
        double v[4];
        values(R + 4 * b, v);
#pragma unroll
        for (int q = 0; q < 4; q++) { const double d = v[q] - mean[q]; ss[q] += d * d; }
    }
    double sd[4];
#pragma unroll
    for (int q = 0; q < 4; q++) sd[q] = sqrt(block_sum(ss[q], red) / (double)(B - 1));
    int any = 0;                                             // (__syncthreads_or is a logical OR: one call per statistic)
#pragma unroll
    for (int q = 0; q < 4; q++) any |= (__syncthreads_or((differs >> q) & 1) ? 1 : 0) << q;
    gt = block_sum(gt, red); eq = block_sum(eq, red); lt = block_sum(lt, red); nan_d = block_sum(nan_d, red);
    if (t != 0) return;
    // identical replicates have SD exactly 0 (R's mean() is exact for them; a rounded mean here would leave 1e-19)
#pragma unroll
    for (int q = 0; q < 4; q++)
        if (!((any >> q) & 1) && first[q] == first[q] && !isinf(first[q])) sd[q] = 0.0;
    double est[4];
    values(win_sums + 4 * (size_t)w, est);
    double *o = out + 16 * (size_t)w;
    o[0] = est[0]; o[1] = est[1]; o[2] = est[2]; o[3] = est[3];
    o[4] = sd[0]; o[5] = sd[1];
    o[6] = sd[2]; o[7] = two_sided_p(est[2] / sd[2]);
    o[8] = sd[3]; o[9] = two_sided_p(est[3] / sd[3]);
    const double nan = __longlong_as_double(0x7FF8000000000000ll);
    const bool na = nan_d > 0;                               // R: sum(t > 0) is NA when a replicate is NaN
    const double tot = gt + eq + lt;
    const double p_gt = na ? nan : lt / tot, p_lt = na ? nan : gt / tot;
    o[10] = na ? nan : gt; o[11] = na ? nan : eq; o[12] = na ? nan : lt;
    o[13] = p_gt; o[14] = p_lt;
    double asl = na ? 1.0 : 2.0 * (p_gt < p_lt ? p_gt : p_lt);
    if (asl == 0) asl = 1.0 / (double)B;
    o[15] = asl;
}

// ---------------------------------------------------------------------------
// Row f4: the pooled-mode contraction f'Tf of snpgenie.pl (:6631-6760).  A pooled sample has no sequences, only codon
// FREQUENCIES per codon position and a mean coverage; the mean numbers of differences are
//   sum over i < j of (f_i cov)(f_j cov) T[i][j] / ((cov^2 - cov) / 2)
// with the same tables as the sequence engine (stop codons contribute 0: they are 0 in the tables).  One thread per
// codon position; the products and the additions are those of the script, in its order (codons sorted alphabetically =
// ascending code), so the values are the script's to the bit.  A position with several codons but no pair of reads
// (cov^2 - cov = 0) makes the script die: NaN here, reported by the host.
// ---------------------------------------------------------------------------
__global__ void k_pooled_diffs(const double *__restrict__ freq, const double *__restrict__ cov, int64_t n, const double *__restrict__ tables,
                               double *__restrict__ n_diffs, double *__restrict__ s_diffs)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const double *TN = tables + 128, *TS = tables + 128 + 4096;
    const double *f = freq + 64 * c;
    const double cv = cov[c];
    // (explicitly rounded products and sums: a fused multiply-add would differ from the script's doubles in the last bit)
    const double comps = __dsub_rn(__dmul_rn(cv, cv), cv) / 2;      // :6647-6649
    int present[64], m = 0;
    for (int i = 0; i < 64; i++)
        if (f[i] > 0) present[m++] = i;
    double dn = 0, ds = 0;
    for (int a = 0; a < m; a++)
        for (int b = a + 1; b < m; b++) {
            const int i = present[a], j = present[b];
            const double num = __dmul_rn(__dmul_rn(f[i], cv), __dmul_rn(f[j], cv));     // :6683-6685
            const double w = num / comps;                                                // :6752
            dn = __dadd_rn(dn, __dmul_rn(TN[i * 64 + j], w));                            // :6753, :6760
            ds = __dadd_rn(ds, __dmul_rn(TS[i * 64 + j], w));
        }
    n_diffs[c] = dn;                                         // (NaN or Inf when comps == 0 and m > 1)
    s_diffs[c] = ds;
}

// ---------------------------------------------------------------------------
// K5: sliding windows (wgp:314-362; bgp:501-592).  Sums run codon by codon in
// ascending order, one thread per (window, statistic): the same additions in
// the same order as the reference, so window sums are bitwise reproducible and
// an all-zero window stays exactly 0 (a prefix-sum difference would not).
// ---------------------------------------------------------------------------
__global__ void k_window_sums(const double *__restrict__ stats4, const int64_t *__restrict__ beg, const int64_t *__restrict__ end,
                              int64_t nwin, double *__restrict__ win_sums)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nwin * 4) return;
    const int64_t k = i >> 2;
    const int q = (int)(i & 3);
    const double *p = stats4 + 4 * beg[k] + q;
    const int64_t W = end[k] - beg[k];
    double s = 0;
    int64_t j = 0;
    for (; j + 16 <= W; j += 16) {         // sixteen loads in flight, the additions in ascending order as before
        double v[16];
#pragma unroll
        for (int u = 0; u < 16; u++) v[u] = __ldg(p + 4 * (j + u));
#pragma unroll
        for (int u = 0; u < 16; u++) s += v[u];
    }
    for (; j < W; j++) s += p[4 * j];
    win_sums[i] = s;
}

// (The per-window bootstrap, wgp:384-427 and bgp:604-736, is k_bootstrap_mma / k_bootstrap over window items: W draws
// uniform over the window's W codons, slot0 = 0.)

}  // namespace

// ------------------------------------------------------------------ launchers

void launch_pack(const uint8_t *d_aln, int64_t row_stride, int64_t n_rows, int64_t row0, const CodonMap *d_map,
                 int64_t n_codons, uint8_t *d_codes, int64_t n_pad, uint32_t *d_other_min, uint32_t *d_other_max,
                 cudaStream_t stream)
{
    if (n_codons <= 0 || n_rows <= 0) return;
    dim3 grid((unsigned)((n_codons + kPackCodons - 1) / kPackCodons), (unsigned)((n_rows + kPackRows - 1) / kPackRows));
    k_pack<<<grid, 256, 0, stream>>>(d_aln, row_stride, n_rows, row0, d_map, n_codons, d_codes, n_pad, d_other_min,
                                     d_other_max);
}

void launch_hist(const uint8_t *d_codes, int64_t n_pad, int64_t n_codons, uint32_t *d_hist, cudaStream_t stream)
{
    if (n_codons <= 0 || n_pad <= 0) return;
    int chunks = (int)((n_pad + kHistChunkRows - 1) / kHistChunkRows);
    int64_t chunk_rows = ((n_pad + chunks - 1) / chunks + 15) & ~(int64_t)15;
    const int64_t items = n_codons * chunks;
    k_hist<<<(unsigned)((items + kHistWarps - 1) / kHistWarps), kHistWarps * 32, 0, stream>>>(d_codes, n_pad, n_codons, chunks,
                                                                                               chunk_rows, d_hist);
}

void launch_first_defined(const uint8_t *d_aln, int64_t row_stride, int64_t n_rows, const CodonMap *d_map, int64_t n_codons,
                          uint8_t *d_first, cudaStream_t stream)
{
    if (n_codons <= 0 || n_rows <= 0) return;
    k_first_defined<<<(unsigned)((n_codons + 127) / 128), 128, 0, stream>>>(d_aln, row_stride, n_rows, d_map, n_codons, d_first);
}

void launch_vote_refs(const uint8_t *d_aln, int64_t pitch, int n_sample, int64_t n_spans, uint8_t *d_cons, uint32_t *d_alt,
                      uint32_t *d_zero_hist, uint32_t *d_other_min, uint32_t *d_other_max, int64_t n_codons, uint32_t *d_sched,
                      cudaStream_t stream)
{
    if (n_spans <= 0 || n_sample <= 0) return;
    // d_zero_hist (or null): hist[n_codons][kHistBins] to clear, 16-byte aligned and with room for whole 16-byte vectors
    // (up to 12 bytes past the last bin), together with other_min (all ones) and other_max (zero)
    const int64_t words = d_zero_hist ? n_codons * kHistBins : 0;
    k_vote_refs<<<(unsigned)((n_spans + 1 + 7) / 8), 256, 0, stream>>>(d_aln, pitch, min(n_sample, 32), n_spans, d_cons, d_alt,
                                                                         reinterpret_cast<uint4 *>(d_zero_hist), (words + 3) / 4,
                                                                         d_zero_hist ? d_other_min : nullptr, d_other_max, n_codons,
                                                                         d_sched ? d_sched + 2 : nullptr);
}

void launch_fused_hist(const uint8_t *d_aln, int64_t pitch, int64_t n_rows, const uint8_t *d_cons, uint32_t *d_alt,
                       int64_t n_spans, const SpanRun *d_runs, uint32_t *d_hist, uint32_t *d_other_min, uint32_t *d_other_max,
                       uint32_t *d_sched, int n_sms, cudaStream_t stream)
{
    if (n_spans <= 0 || n_rows <= 0) return;
    using Kernel = void (*)(const uint8_t *, int64_t, uint32_t, const uint8_t *, uint32_t *, uint32_t, const SpanRun *, uint32_t *, uint32_t *,
                            uint32_t *, uint32_t *, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t,
                            uint32_t, uint32_t, uint32_t);
    // variant (SNPG_SPAN_VARIANT): 0 (default) = 4 blocks/SM at <= 64 registers; 1 = 3 blocks/SM, no register cap
    // 2 / 3 = variant 0 with the next four rows prefetched into L1 by every lane / by one lane per 128-byte line
    // k_fused_hist_db (double-buffered row groups): 4 = groups of 2 rows, 4 blocks/SM; 5 = groups of 4, 3 blocks/SM; 6 = groups of 2, 3 blocks/SM
    // (groups of 4 in four blocks of 7 warps per SM, 72 registers: 99 us, the spills sit in the row loop)
    // (rings of 64 entries drained after every row -- half the shared memory, more L1 -- were measured with both kernels: 78 us)
    static const int variant = [] { const char *v = getenv("SNPG_SPAN_VARIANT"); return v ? max(0, min(6, atoi(v))) : 0; }();
    static const Kernel kernel = variant == 1 ? (Kernel)k_fused_hist<4, 2, 128, 3> : variant == 2 ? (Kernel)k_fused_hist<4, 2, 128, 4, 1>
                                 : variant == 3 ? (Kernel)k_fused_hist<4, 2, 128, 4, 2> : variant == 4 ? (Kernel)k_fused_hist_db<2, 128, 4>
                                 : variant == 5 ? (Kernel)k_fused_hist_db<4, 128, 3> : variant == 6 ? (Kernel)k_fused_hist_db<2, 128, 3>
                                 : (Kernel)k_fused_hist<4, 2, 128, 4>;
    constexpr int warps = kFusedWarps;
    static const size_t smem = (size_t)(warps + 1) * 128 * kFusedEntryBytes;
    static const int blocks_per_sm = [] {
        int nb = 0;
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, warps * 32, smem) != cudaSuccess || nb < 1) nb = 1;
        if (const char *v = getenv("SNPG_SPAN_BLOCKS")) nb = max(1, min(nb, atoi(v)));
        return nb;
    }();
    static const int env_rows = [] { const char *v = getenv("SNPG_SPAN_ROWS"); return v ? max(4, atoi(v) & ~3) : 0; }();
    static const int env_tail = [] { const char *v = getenv("SNPG_SPAN_TAIL"); return v ? max(0, min(100, atoi(v))) : 2; }();
    const int64_t n_sb = (n_spans + kFusedSpans - 1) / kFusedSpans;
    // chunk indices are 32-bit: very long alignments go out in several launches over row ranges
    const int64_t rows_per_launch = std::max<int64_t>(kFusedTailRows, (((int64_t)1 << 31) / n_sb) & ~(int64_t)3);
    for (int64_t r0 = 0; r0 < n_rows; r0 += rows_per_launch) {
        const int64_t rows = std::min<int64_t>(rows_per_launch, n_rows - r0);
        int64_t grid = (int64_t)n_sms * blocks_per_sm;
        // chunk of rows a warp claims at a time (measured on config 2: 12 rows 94 us, 16 rows 90, 32 rows 82, 64 rows
        // 84, 128 rows 91).  When every warp gets at least two chunks, the last env_tail % of the rows go out in chunks
        // of kFusedTailRows to even out the finish; a small launch (a staged block of a host buffer, a rank's eighth of
        // the columns) is one chunk per warp and only the rows left over by the chunk size go out small.
        // Where the vote's sample says that many rows will be parked (a quarter of the (span, sample row) pairs match no
        // reference: config 3's damaged bases), the kernel takes the second geometry instead: chunks of half a warp's
        // share, at least 8 rows.  Measured (profiles/r2_chunk_sweep.txt): config 3 79 us at 32 rows, 55 at 8; a rank's
        // eighth of config 2 (cheap rows) 24 us at 32 rows, 30 at 8, 39 at 4.
        const int64_t share = rows * n_sb / (grid * warps);
        struct Geo { int64_t chunk, rows_a, k_a, k_b; };
        auto geometry = [&](int64_t chunk) {
            Geo g;
            g.chunk = chunk;
            g.rows_a = (share >= 2 * chunk ? rows - rows * env_tail / 100 : rows) / chunk * chunk;
            g.k_a = g.rows_a / chunk * n_sb;
            g.k_b = (rows - g.rows_a + kFusedTailRows - 1) / kFusedTailRows * n_sb;
            return g;
        };
        const Geo cheap = geometry(env_rows ? env_rows : 32);
        const Geo dense = geometry(env_rows ? env_rows : std::max<int64_t>(8, std::min<int64_t>(32, (share / 2) & ~(int64_t)3)));
        const uint32_t dense_min = (uint32_t)std::max<int64_t>(1, std::min<int64_t>(32, rows) * n_spans / 4);
        grid = std::max<int64_t>(1, std::min<int64_t>(grid, (std::max(cheap.k_a + cheap.k_b, dense.k_a + dense.k_b) + warps - 1) / warps));
        kernel<<<(unsigned)grid, warps * 32, smem, stream>>>(
            d_aln + (size_t)r0 * pitch, pitch, (uint32_t)rows, d_cons, d_alt, (uint32_t)n_spans, d_runs, d_hist, d_other_min, d_other_max,
            d_sched, (uint32_t)n_sb, n_sb == 1 ? 0xFFFFFFFFu : (uint32_t)(((uint64_t)1 << 32) / (uint64_t)n_sb), (uint32_t)cheap.chunk,
            (uint32_t)cheap.rows_a, (uint32_t)cheap.k_a, (uint32_t)(cheap.k_a + cheap.k_b), dense_min, (uint32_t)dense.chunk,
            (uint32_t)dense.rows_a, (uint32_t)dense.k_a, (uint32_t)(dense.k_a + dense.k_b));
    }
}

void launch_fused_fixup_stats(const uint8_t *d_cons, int64_t cons_shift, uint32_t *d_alt, int64_t n_spans, const CodonMap *d_map, int64_t n_codons,
                              uint32_t n_rows, uint32_t *d_hist, uint32_t *d_other_min, uint32_t *d_other_max, GroupView A,
                              const double *d_tables, StatsOut out, const PeerOut *po, const PeerSync *ps, cudaStream_t stream)
{
    if (n_codons <= 0) { if (ps) launch_peer_signal(*ps, kStageStats, stream); return; }
    k_fused_fixup_stats<<<(unsigned)((n_codons + kStatsWarps - 1) / kStatsWarps), kStatsWarps * 32, 0, stream>>>(
        d_cons, cons_shift, ref_planes(d_alt, n_spans), d_map, n_codons, n_rows, d_hist, d_other_min, d_other_max, A, d_tables, out, po ? *po : PeerOut(),
        ps ? *ps : PeerSync());
}

void launch_peer_signal(const PeerSync &ps, int stage, cudaStream_t stream)
{
    if (ps.world > 1) k_peer_signal<<<1, 32, 0, stream>>>(ps, stage);
}

void launch_peer_wait(const PeerSync &ps, int stage, cudaStream_t stream)
{
    if (ps.world > 1) k_peer_wait<<<1, 32, 0, stream>>>(ps, stage);
}

void launch_fused_fixup(const uint8_t *d_cons, int64_t cons_shift, uint32_t *d_alt, int64_t n_spans, const CodonMap *d_map, const uint8_t *d_regular,
                        int64_t n_codons, uint32_t n_rows, uint32_t *d_hist, uint32_t *d_other_min, uint32_t *d_other_max,
                        cudaStream_t stream)
{
    if (n_codons <= 0) return;
    k_fused_fixup<<<(unsigned)((n_codons + 7) / 8), 256, 0, stream>>>(d_cons, cons_shift, ref_planes(d_alt, n_spans), d_map, d_regular, n_codons, n_rows,
                                                                       d_hist, d_other_min, d_other_max);
}

void launch_scatter_irregular(const uint32_t *d_tmp_hist, const uint32_t *d_tmp_min, const uint32_t *d_tmp_max,
                              const int32_t *d_idx, int64_t n_irr, uint32_t *d_hist, uint32_t *d_other_min,
                              uint32_t *d_other_max, cudaStream_t stream)
{
    if (n_irr <= 0) return;
    k_scatter_irregular<<<(unsigned)((n_irr * kHistBins + 255) / 256), 256, 0, stream>>>(d_tmp_hist, d_tmp_min, d_tmp_max, d_idx,
                                                                                          n_irr, d_hist, d_other_min, d_other_max);
}

void launch_site_hist(const uint8_t *d_aln, int64_t pitch, int64_t n_rows, const uint8_t *d_cons, int64_t n_cols, uint32_t *d_sites,
                      cudaStream_t stream)
{
    if (n_cols <= 0 || n_rows <= 0) return;
    const int64_t n_spans = (n_cols + 15) / 16;
    const int64_t rows_per_block = (int64_t)kSiteWarps * kSiteRowsPerWarp;
    dim3 grid((unsigned)((n_spans + 31) / 32), (unsigned)((n_rows + rows_per_block - 1) / rows_per_block));
    k_site_hist<<<grid, kSiteWarps * 32, 0, stream>>>(d_aln, pitch, n_rows, d_cons, n_spans, n_cols, d_sites);
}

void launch_site_fixup(const uint8_t *d_cons, int64_t n_cols, uint32_t n_rows, uint32_t *d_sites, cudaStream_t stream)
{
    if (n_cols <= 0) return;
    k_site_fixup<<<(unsigned)((n_cols + 127) / 128), 128, 0, stream>>>(d_cons, n_cols, n_rows, d_sites);
}

void launch_untranspose(const uint8_t *d_codes, int64_t n_pad, int64_t n, int64_t c0, int64_t C, uint8_t *d_out,
                        cudaStream_t stream)
{
    if (C <= 0 || n <= 0) return;
    dim3 grid((unsigned)((C + 31) / 32), (unsigned)((n + 31) / 32)), block(32, 32);
    k_untranspose<<<grid, block, 0, stream>>>(d_codes, n_pad, n, c0, C, d_out);
}

void launch_within_stats(GroupView A, int64_t n_codons, const double *d_tables, StatsOut out, const PeerOut *po, const PeerSync *ps,
                         cudaStream_t stream)
{
    if (n_codons <= 0) { if (ps) launch_peer_signal(*ps, kStageStats, stream); return; }
    k_stats<false><<<(unsigned)((n_codons + kStatsWarps - 1) / kStatsWarps), kStatsWarps * 32, 0, stream>>>(
        A, A, n_codons, d_tables, out, po ? *po : PeerOut(), ps ? *ps : PeerSync());
}

void launch_between_stats(GroupView A, GroupView B, int64_t n_codons, const double *d_tables, StatsOut out,
                          cudaStream_t stream)
{
    if (n_codons <= 0) return;
    k_stats<true><<<(unsigned)((n_codons + kStatsWarps - 1) / kStatsWarps), kStatsWarps * 32, 0, stream>>>(A, B, n_codons, d_tables, out,
                                                                                                          PeerOut(), PeerSync());
}

void launch_taxa_filter(double *d_stats4, const uint32_t *d_ndef_a, const uint32_t *d_ndef_b, int64_t n_codons, uint32_t min_total,
                        uint32_t min_group, cudaStream_t stream)
{
    if (n_codons <= 0 || (min_total == 0 && min_group == 0)) return;
    k_taxa_filter<<<(unsigned)((n_codons + 255) / 256), 256, 0, stream>>>(d_stats4, d_ndef_a, d_ndef_b, n_codons, min_total, min_group);
}

void launch_product_sums(const double *d_stats4, const int64_t *d_prod_ofs, int n_products, double *d_sums,
                         cudaStream_t stream)
{
    if (n_products <= 0) return;
    k_product_sums<<<n_products, kSumThreads, 0, stream>>>(d_stats4, d_prod_ofs, d_sums);
}

void launch_bootstrap(int form, BootArgs a, int64_t max_codons, cudaStream_t stream)
{
    if (a.n_items <= 0 || a.b_count <= 0) { launch_peer_signal(a.ps, kStageVals, stream); return; }
    if (a.vals_stride <= 0) a.vals_stride = a.b_count;
    const int64_t per_block = (int64_t)kBootWarps * kBootRepsPerWarp;
    dim3 grid((unsigned)((a.b_count + per_block - 1) / per_block), (unsigned)a.n_items);
    const size_t smem = (size_t)(max_codons + 1) * 32;
    // a warp's counters: one byte per slot, rows padded to an odd number of words (no two of the eight rows a
    // tensor-core tile reads share a bank)
    // (at least 16 KB a block even for a short table, so that short items take up to eight replicates per warp: sliding
    // windows of 100 codons at 127 words took four -- config 3 whole job 0.260 ms, at 255 words 0.243, at 511 0.239)
    static const size_t min_row_words = [] { const char *v = getenv("SNPG_BOOT_MINROW"); return v ? (size_t)(std::max(31, std::min(1023, atoi(v))) | 1) : (size_t)511; }();
    const size_t row_words = std::max<size_t>((((size_t)max_codons + 1 + 3) / 4) | 1, min_row_words);
    const bool classes = form == SNPG_BOOT_CLASSES && a.perm && a.cls && a.split_m && a.split_ns && max_codons <= kLogFactMax;
    if ((form == SNPG_BOOT_WEIGHTS || form == SNPG_BOOT_CLASSES) && row_words * 4 * 32 <= (size_t)kBootSmemLimit) {
        // five blocks of kMmaWarps warps per SM; the grid is sized for one replicate per warp.  A block takes two batches
        // of replicates when that still leaves every SM a few blocks per slot (measured on config 2: one batch 157 us, two
        // 147, four 168); with few replicates (a rank of several GPUs) one batch, then two warps per replicate.
        static const int n_sms = [] { int d = 0, n = 0; cudaGetDevice(&d); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); return n > 0 ? n : 148; }();
        static const int env_nb = [] { const char *v = getenv("SNPG_BOOT_NB"); return v ? atoi(v) : 0; }();
        static const int env_split = [] { const char *v = getenv("SNPG_BOOT_SPLIT"); return v ? atoi(v) : -1; }();
        // What counts is the longest item (scheduled first, and on config 2 nearly half of all draws): its blocks alone should
        // fill most of the 5 * n_sms block slots, otherwise the SMs end the launch with one or two long blocks each.
        const int64_t slots = 5ll * n_sms;
        auto blocks = [&](int nb, int split) { const int64_t per = (int64_t)(kMmaWarps >> split) * nb; return (a.b_count + per - 1) / per; };
        int nb = 2, split = 0;
        if (blocks(2, 0) * 5 < slots * 4) { nb = 1; if (blocks(1, 0) * 2 < slots) split = 1; }
        if (env_nb == 1 || env_nb == 2) nb = env_nb;
        if (env_split == 0 || env_split == 1) split = env_split;
        // replicates a block takes: the kernel's own arithmetic for the LONGEST item (shorter items take more per warp and
        // leave their last blocks idle).  Sized for one replicate per warp, a launch over sliding windows (100 slots: four
        // replicates per warp) had seven idle blocks in eight: 60,000 blocks of which 7,680 worked, 114 us for config 3's
        // window bootstraps.
        const uint32_t rw_max = (uint32_t)(((size_t)max_codons + (size_t)(classes ? 0 : a.slot0) + 3) / 4) | 1u;
        uint32_t lm = 0;
        while ((2u << lm) <= (uint32_t)kBootMaxRows && (size_t)(2u << lm) * rw_max <= row_words) lm++;
        const int64_t per_mma = lm ? (int64_t)kMmaWarps << lm : (int64_t)(kMmaWarps >> split) * nb;
        dim3 grid_mma((unsigned)((a.b_count + per_mma - 1) / per_mma), (unsigned)a.n_items);
        if (classes) {
            static const int env_shift = [] { const char *v = getenv("SNPG_BOOT_CLASS_SHIFT"); return v ? std::max(1, std::min(12, atoi(v))) : 0; }();
            if (env_shift) a.cls_shift = env_shift;
            a.mma_row_words = (uint32_t)row_words; a.mma_nb = (uint32_t)nb; a.mma_split = (uint32_t)split;     // -> BootClasses::mma_blocks
            launch_peer_wait(a.ps, kStageStats, stream);         // the classes come from every rank's statistics
            k_boot_classify<<<(unsigned)a.n_items * kClsCtas, kClassifyThreads, 0, stream>>>(a);
            k_boot_split<<<dim3((unsigned)((a.b_count + kSplitThreads - 1) / kSplitThreads), (unsigned)a.n_items), kSplitThreads, 0, stream>>>(a);
        }
        auto kernel = classes ? (split ? k_bootstrap_mma<1, true> : k_bootstrap_mma<0, true>) : (split ? k_bootstrap_mma<1, false> : k_bootstrap_mma<0, false>);
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBootSmemLimit * kMmaWarps / 32);  // per device
        kernel<<<grid_mma, kMmaWarps * 32, row_words * 4 * kMmaWarps, stream>>>(a, (uint32_t)row_words, (uint32_t)nb);
    } else if (smem <= (size_t)kBootSmemLimit) {
        cudaFuncSetAttribute(k_bootstrap<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBootSmemLimit);  // per device
        k_bootstrap<true><<<grid, kBootWarps * 32, smem, stream>>>(a);
    } else {
        // tables beyond shared memory: gathers from global memory; the sums need the exchanged table, so they follow the kernel
        double *sums = a.prod_sums;
        a.prod_sums = nullptr;
        k_bootstrap<false><<<grid, kBootWarps * 32, 0, stream>>>(a);
        if (sums) launch_product_sums(a.stats4, a.beg, a.n_items, sums, stream);
    }
}

// Once per ctx, before any job (snpg_create): loads every kernel a job may launch AFTER a kernel that waits for peer
// GPUs -- with CUDA's lazy module loading the first launch of a kernel loads it, which waits for the device to drain; a
// rank spinning for peers whose work the same host thread has not enqueued yet (one process, several GPUs) would never
// get there -- and fills the table of log-factorials of the current device.
cudaError_t prepare_device_kernels()
{
    cudaFuncAttributes fa;
    const void *kernels[] = {(const void *)k_boot_classify, (const void *)k_boot_split, (const void *)k_bootstrap_mma<0, false>,
                             (const void *)k_bootstrap_mma<1, false>, (const void *)k_bootstrap_mma<0, true>, (const void *)k_bootstrap_mma<1, true>,
                             (const void *)k_bootstrap<true>, (const void *)k_bootstrap<false>, (const void *)k_moments<false, false>, (const void *)k_moments<true, false>,
                             (const void *)k_moments<false, true>,
                             (const void *)k_window_sums, (const void *)k_product_sums, (const void *)k_peer_signal, (const void *)k_peer_wait,
                             (const void *)k_stats<false>, (const void *)k_stats<true>, (const void *)k_fused_fixup_stats, (const void *)k_fused_fixup,
                             (const void *)k_vote_refs, (const void *)k_fill_logfact, (const void *)k_window_rstats, (const void *)k_taxa_filter};
    for (const void *k : kernels) {
        const cudaError_t e = cudaFuncGetAttributes(&fa, k);
        if (e != cudaSuccess) return e;
    }
    k_fill_logfact<<<(kLogFactMax + 256) / 256, 256>>>();
    return cudaDeviceSynchronize();
}

size_t boot_scratch_bytes(int64_t total, int n_items, int64_t b_count)
{
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    return up(sizeof(double) * 4 * (size_t)std::max<int64_t>(1, total)) + up(sizeof(BootClasses) * (size_t)std::max(1, n_items)) +
           up(sizeof(uint32_t) * (size_t)n_items * (size_t)b_count) + up(sizeof(double2) * (size_t)n_items * (size_t)b_count);
}

void boot_scratch_carve(void *scratch, int64_t total, int n_items, int64_t b_count, BootArgs &a)
{
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    uint8_t *p = static_cast<uint8_t *>(scratch);
    a.perm = reinterpret_cast<double *>(p); p += up(sizeof(double) * 4 * (size_t)std::max<int64_t>(1, total));
    a.cls = reinterpret_cast<BootClasses *>(p); p += up(sizeof(BootClasses) * (size_t)std::max(1, n_items));
    a.split_m = reinterpret_cast<uint32_t *>(p); p += up(sizeof(uint32_t) * (size_t)n_items * (size_t)b_count);
    a.split_ns = reinterpret_cast<double2 *>(p);
}

void launch_philox_kat(const uint32_t *ctr, const uint32_t *key, uint32_t *d_out, cudaStream_t stream)
{
    k_philox_kat<<<1, 1, 0, stream>>>(make_uint4(ctr[0], ctr[1], ctr[2], ctr[3]), make_uint2(key[0], key[1]), d_out);
}

void launch_moments(const double *d_rep, int64_t n_items, int64_t B, double *d_se, cudaStream_t stream)
{
    if (n_items <= 0) return;
    // (the form that derives the statistic from the replicate sums keeps the plain two passes: with ten cached values it sits at the
    // 64-register limit of a 1,024-thread block, and only snpg_rep_moments_device -- a caller doing its own plumbing -- comes here)
    k_moments<true, false><<<dim3((unsigned)n_items, 6), kMomentThreads, 0, stream>>>(d_rep, n_items, B, d_se, nullptr, PeerSync());
}

void launch_moments_vals(const double *d_vals, int64_t n_items, int64_t B, double *d_se, uint32_t *d_undefined, const PeerSync *ps,
                         cudaStream_t stream, int n_which)
{
    if (n_items <= 0) return;
    auto kernel = B > kMomentThreads ? k_moments<false, true> : k_moments<false, false>;
    kernel<<<dim3((unsigned)n_items, (unsigned)n_which), kMomentThreads, 0, stream>>>(d_vals, n_items, B, d_se, d_undefined, ps ? *ps : PeerSync());
}

void launch_pooled_diffs(const double *d_freq, const double *d_cov, int64_t n, const double *d_tables, double *d_n_diffs, double *d_s_diffs,
                         cudaStream_t stream)
{
    if (n <= 0) return;
    k_pooled_diffs<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(d_freq, d_cov, n, d_tables, d_n_diffs, d_s_diffs);
}

void launch_window_rstats(const double *d_rep, const double *d_win_sums, int64_t nwin, int64_t B, int num, int den, int jc, double *d_out,
                          cudaStream_t stream)
{
    if (nwin <= 0) return;
    k_window_rstats<<<(unsigned)nwin, kMomentThreads, 0, stream>>>(d_rep, d_win_sums, B, num, den, jc, d_out);
}

void launch_window_sums(const double *d_stats4, const int64_t *d_beg, const int64_t *d_end, int64_t nwin, double *d_win_sums,
                        cudaStream_t stream)
{
    if (nwin <= 0) return;
    k_window_sums<<<(unsigned)((nwin * 4 + 127) / 128), 128, 0, stream>>>(d_stats4, d_beg, d_end, nwin, d_win_sums);
}

}  // namespace snpg

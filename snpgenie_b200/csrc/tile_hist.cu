// K1+K2 fused, TMA-tiled: per-codon histograms straight from the alignment bytes.
//
// Replaces, like k_fused_hist, the codon split (snpgenie_within_group.pl:417-454), the
// polymorphism pre-check (wg:469-541; bg:472-544) and the O(n^2) pair counting
// (wg:596-609; bg:649-667): the 66 counts of a codon column determine every per-codon output.
//
// Layout of the work.  The host cuts every run of consecutive, contiguous codons of one product
// into *tiles* of up to 64 codons = 192 alignment bytes that start at the run's own frame (any
// byte column).  A work item is (tile, block of 128 rows); consecutive items of a tile form a
// chunk.  Persistent CTAs (two per SM) claim chunks from a global counter.  A producer warp issues
// one TMA tensor load per item (cp.async.bulk.tensor.2d, box 208 B x 128 rows, starting at the
// tile's byte column rounded down to 16 -- TMA faults on an innermost coordinate that is not
// 16-byte aligned, measured with tools/tma_probe.cu) into a 3-stage shared-memory ring: no load
// instruction or global address arithmetic is spent by the consumers and the loads never occupy
// registers (tools/tma_stream.cu: this access pattern alone streams at 5.8 TB/s).
//
// Sixteen consumer warps share the rows of a stage.  A lane owns one group of four codons (12
// bytes: four aligned 32-bit shared-memory words funnel-shifted by the tile's sub-word offset into
// three; lanes 0-15 take row r, lanes 16-31 row r+4, which makes the loads bank-conflict free at
// the 52-word row pitch).  For each of its four codons the lane keeps a tiny allele cache in
// registers: the consensus codon's bytes plus up to two alternate byte patterns with counters.
// A row costs three masked XOR-compares and two predicated adds per codon -- no ballots, no
// queues, no divergence, and a cost that does not depend on how polymorphic the column is.  A row
// whose codon matches none of the three patterns sets a bit in the lane's miss mask; after the
// item the lane revisits those rows: it installs the pattern in a free cache slot or, when both
// are taken, counts that single codon directly.  At the end of a chunk every (codon, alternate)
// with a non-zero counter is classified once through the 256-entry LUT and added to
// hist[codon][code] with one RED.  k_fused_fixup then credits each codon's consensus code with
// (rows - recorded), exactly as for k_fused_hist.  Bytes are compared raw, so a lower-case or 'U'
// residue is simply another pattern that classifies to the same code.
//
// Algorithmic traffic: the CDS bytes of every row once (3 B per (sequence, codon)) + 264 B per
// codon of histogram.  Non-coding columns are never read; overlapping products read their
// shared bytes once per product (L2 hits when the tiles are processed close in time).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>

#include "snpg_device.cuh"
#include "snpg_internal.h"

namespace snpg {

namespace {

constexpr int kTileCodons = 64;
constexpr int kTileBytes = 3 * kTileCodons + 16;         // 208: box width; the tile's 192 bytes start 0..15 bytes in
constexpr int kTileStages = 3;                           // per CTA; two CTAs share an SM
constexpr int kTileConsumers = 16;                       // consumer warps; warp kTileConsumers is the producer
constexpr int kTileRowsPerWarp = kTileRows / kTileConsumers;   // 8 rows = 4 row pairs per stage
constexpr int kTileStageBytes = kTileRows * kTileBytes;  // 26,624
constexpr int kTileCtasPerSm = 2;
constexpr int kTileChunkRbs = 8;                         // blocks of rows per scheduling chunk (1,024 rows)

static_assert(kTileRowsPerWarp == 8, "the consumer loop is written for four row pairs per stage");
static_assert(kTileBytes % 16 == 0 && kTileBytes <= 256 && (4 * kTileBytes / 4) % 32 == 16, "box width: TMA limits and the bank layout");

// What the producer tells the consumers about the stage it filled (written before its arrive on the
// stage's full barrier, read after the wait): the tile's descriptor, the block of rows, and whether the
// stage also carries the tile's consensus bytes (first stage of a chunk).  tile < 0 ends the kernel.
struct StageMeta { int32_t x0, codon0, info, rb; };     // info = ncodons | dir << 8 | minus << 16 | first << 24 (dir as int8)

struct TileSmem {
    uint8_t stage[kTileStages][kTileStageBytes];         // TMA destinations (128-byte aligned)
    uint8_t cons_row[kTileStages][256];                  // consensus bytes of the tile's 208-byte window (first stage of a chunk)
    uint8_t lut_chr[2][256];
    uint8_t lut_cls[2][256];
    StageMeta meta[kTileStages];
    uint64_t full[kTileStages];
    uint64_t empty[kTileStages];
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared (16-byte aligned, size multiple of 16), completing on the same kind of barrier
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int32_t x, int32_t y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y)
                 : "memory");
}

// Adds `count` sequences whose codon `cidx` has the raw bytes pat (byte 0 = lowest alignment column).
__device__ __forceinline__ void count_pattern(uint32_t pat, uint32_t cidx, uint32_t count, int minus, uint32_t *__restrict__ hist,
                                              uint32_t *__restrict__ other_min, uint32_t *__restrict__ other_max,
                                              const uint8_t (*lut_cls)[256], const uint8_t (*lut_chr)[256])
{
    const uint32_t r0 = pat & 0xFF, r1 = (pat >> 8) & 0xFF, r2 = (pat >> 16) & 0xFF;
    const uint32_t b0 = minus ? r2 : r0, b2 = minus ? r0 : r2;          // '-' strand reads the codon right to left
    const uint8_t *cls = lut_cls[minus];
    const uint32_t k0 = cls[b0], k1 = cls[r1], k2 = cls[b2];
    uint32_t code = k0 * 16 + k1 * 4 + k2;
    if (max(k0, max(k1, k2)) >= 4) {
        code = (k0 == 4 || k1 == 4 || k2 == 4) ? kCodeUndef : kCodeOther;
        if (code == kCodeOther) {
            const uint8_t *chr = lut_chr[minus];
            const uint32_t key = ((uint32_t)chr[b0] << 16) | ((uint32_t)chr[r1] << 8) | chr[b2];
            atomicMin(&other_min[cidx], key);
            atomicMax(&other_max[cidx], key);
        }
    }
    atomicAdd(&hist[cidx * (uint32_t)kHistBins + code], count);          // 32-bit index: codons * 66 < 2^32
}

// The lane's four codons of one row, each in its own register: codons 0..2 in bits 0-23, codon 3 in bits 8-31
// (patterns and masks for codon 3 are kept in that position, so no shift is spent on it).
__device__ __forceinline__ void split_codons(const uint32_t *row, uint32_t sh, uint32_t (&y)[4]) {
    const uint32_t w0 = row[0], w1 = row[1], w2 = row[2], w3 = row[3];
    const uint32_t x0 = __funnelshift_r(w0, w1, sh), x1 = __funnelshift_r(w1, w2, sh), x2 = __funnelshift_r(w2, w3, sh);
    y[0] = x0;
    y[1] = __byte_perm(x0, x1, 0x0543);
    y[2] = __byte_perm(x1, x2, 0x0432);
    y[3] = x2;
}

// Per-lane allele cache of the current chunk: for codon k the consensus pattern pc[k], up to two alternates
// a1[k], a2[k] (an empty slot holds the consensus pattern, so it can never turn a mismatch into a hit) and
// their counters packed in cnt[k] (a1 low half, a2 high half; a chunk gives a lane at most a few hundred
// rows).  nalt holds two bits per codon: how many alternates are installed.
struct AlleleCache {
    uint32_t pc[4], pm[4], a1[4], a2[4], cnt[4], nalt;
};

template <bool TAIL>
__device__ __forceinline__ uint32_t scan_rows(const uint32_t *sp, uint32_t sh, AlleleCache &C, int rows_left)
{
    uint32_t miss = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t y[4];
        split_codons(sp + i * (kTileBytes / 4), sh, y);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            bool p0 = ((y[k] ^ C.pc[k]) & C.pm[k]) == 0;
            bool p1 = ((y[k] ^ C.a1[k]) & C.pm[k]) == 0;
            bool p2 = ((y[k] ^ C.a2[k]) & C.pm[k]) == 0;
            if (TAIL) { const bool real = i < rows_left; p0 |= !real; p1 &= real; p2 &= real; }
            if (p1) C.cnt[k] += 1u;
            if (p2) C.cnt[k] += 0x10000u;
            if (!(p0 || p1 || p2)) miss |= 1u << (4 * i + k);
        }
    }
    return miss;
}

__global__ void __launch_bounds__((kTileConsumers + 1) * 32, kTileCtasPerSm)
k_tile_hist(const __grid_constant__ CUtensorMap tmap, int n_rows, const uint8_t *__restrict__ cons,
            const CodonTile *__restrict__ tiles, int n_tiles, int n_rb, int chunk_rbs, int row_major, int debug_skip, uint32_t *__restrict__ sched,
            uint32_t *__restrict__ hist, uint32_t *__restrict__ other_min, uint32_t *__restrict__ other_max)
{
    extern __shared__ uint8_t smem_raw[];
    // TMA destinations want 128-byte alignment; the launch reserves the slack
    TileSmem &S = *reinterpret_cast<TileSmem *>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;

    if (t < 256) {
        for (int s = 0; s < 2; s++) {
            const uint8_t n = norm_char((uint8_t)t, s);
            S.lut_chr[s][t] = n;
            S.lut_cls[s][t] = class_of(n);
        }
    }
    if (t == 0) {
        for (int s = 0; s < kTileStages; s++) {
            mbar_init(&S.full[s], 1);
            mbar_init(&S.empty[s], kTileConsumers);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    uint32_t stage = 0, phase = 0;

    if (warp == kTileConsumers) {
        // ---------------- producer: one lane claims chunks and feeds the ring ----------------
        // A chunk is (tile, chunk_rbs consecutive blocks of rows).  Chunks are claimed from a global counter,
        // so an SM that drew tiles with many differing codons simply claims fewer of them.
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
            const uint32_t chunks_per_tile = (uint32_t)((n_rb + chunk_rbs - 1) / chunk_rbs);
            const uint32_t n_chunks = (uint32_t)n_tiles * chunks_per_tile;
            uint32_t next = atomicAdd(&sched[0], 1u);
            while (next < n_chunks) {
                const uint32_t chunk = next;
                next = atomicAdd(&sched[0], 1u);          // claimed early: its latency hides behind this chunk
                // row-major order: CTAs that run at the same time read neighbouring tiles of the same rows
                const uint32_t tile = row_major ? chunk % (uint32_t)n_tiles : chunk / chunks_per_tile;
                const int rb0 = (int)(row_major ? chunk / (uint32_t)n_tiles : chunk % chunks_per_tile) * chunk_rbs, rb1 = min(n_rb, rb0 + chunk_rbs);
                const int4 td = __ldg(reinterpret_cast<const int4 *>(tiles + tile));     // x0, codon0, ncodons|dir|minus, pad
                const int xa = td.x & ~15;
                const int info = (td.z & 0xFF) | (((td.z >> 16) & 0xFF) << 8) | (((td.z >> 24) & 1) << 16);   // ncodons, dir, minus
                for (int rb = rb0; rb < rb1; rb++) {
                    const bool first = rb == rb0;
                    mbar_wait(&S.empty[stage], phase ^ 1);
                    S.meta[stage] = StageMeta{td.x, td.y, info | (first ? 1 << 24 : 0), rb};
                    mbar_arrive_expect_tx(&S.full[stage], kTileStageBytes + (first ? kTileBytes : 0));
                    tma_load_2d(S.stage[stage], &tmap, &S.full[stage], xa, rb * kTileRows);
                    if (first) bulk_load(S.cons_row[stage], cons + xa, kTileBytes, &S.full[stage]);
                    if (++stage == kTileStages) { stage = 0; phase ^= 1; }
                }
            }
            // end marker, then leave the scheduler clean for the next launch (the last CTA to finish resets it)
            mbar_wait(&S.empty[stage], phase ^ 1);
            S.meta[stage] = StageMeta{0, 0, 0, -1};
            mbar_arrive(&S.full[stage]);
            __threadfence();
            if (atomicAdd(&sched[1], 1u) == gridDim.x - 1) { sched[0] = 0; sched[1] = 0; __threadfence(); }
        }
        return;
    }

    // -------------------- consumers --------------------
    const int gl = lane & 15, half = lane >> 4;
    int codon0 = 0, dir = 1, minus = 0;
    uint32_t lane_ofs = 0, sh = 0;               // byte offset of the lane's first aligned word in its row; funnel shift in bits
    AlleleCache C;
#pragma unroll
    for (int k = 0; k < 4; k++) { C.pc[k] = 0; C.pm[k] = 0; C.a1[k] = 0; C.a2[k] = 0; C.cnt[k] = 0; }
    C.nalt = 0;

    // every installed alternate with a non-zero counter goes to the histogram once
    auto flush = [&]() {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t n = (C.nalt >> (2 * k)) & 3u;
            const uint32_t cidx = (uint32_t)(codon0 + dir * (4 * gl + k));
            const uint32_t c1 = C.cnt[k] & 0xFFFFu, c2 = C.cnt[k] >> 16;
            if (n >= 1 && c1) count_pattern(k == 3 ? C.a1[k] >> 8 : C.a1[k], cidx, c1, minus, hist, other_min, other_max, S.lut_cls, S.lut_chr);
            if (n >= 2 && c2) count_pattern(k == 3 ? C.a2[k] >> 8 : C.a2[k], cidx, c2, minus, hist, other_min, other_max, S.lut_cls, S.lut_chr);
        }
    };

    for (;;) {
        mbar_wait(&S.full[stage], phase);
        const int4 mt = *reinterpret_cast<const int4 *>(&S.meta[stage]);
        if (mt.w < 0) break;
        if (mt.z >> 24) {
            // a new chunk: the counters belong to the previous tile
            flush();
            codon0 = mt.y; dir = (int)(int8_t)(mt.z >> 8); minus = (mt.z >> 16) & 1;
            const int cnt = min(4, max(0, (mt.z & 0xFF) - 4 * gl));
            lane_ofs = (uint32_t)(((mt.x & 15) >> 2) + 3 * gl) * 4u;
            sh = (uint32_t)(mt.x & 3) * 8u;
            uint32_t y[4];
            split_codons(reinterpret_cast<const uint32_t *>(S.cons_row[stage] + lane_ofs), sh, y);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                C.pm[k] = k < cnt ? (k == 3 ? 0xFFFFFF00u : 0x00FFFFFFu) : 0u;      // an absent codon matches everything
                C.pc[k] = y[k] & C.pm[k];
                C.a1[k] = C.pc[k];
                C.a2[k] = C.pc[k];
                C.cnt[k] = 0;
            }
            C.nalt = 0;
        }

        // lanes 0-15 read row i of the warp's eight, lanes 16-31 row i + 4
        const uint32_t *sp = reinterpret_cast<const uint32_t *>(S.stage[stage] + (warp * kTileRowsPerWarp + 4 * half) * kTileBytes + lane_ofs);
        const int rows_left = n_rows - (mt.w * kTileRows + warp * kTileRowsPerWarp + 4 * half);   // row i is real iff i < rows_left
        const bool tail = n_rows - mt.w * kTileRows < kTileRows;                                   // warp-uniform: last block of rows
        uint32_t miss = 0;
        if (!debug_skip) miss = tail ? scan_rows<true>(sp, sh, C, rows_left) : scan_rows<false>(sp, sh, C, rows_left);

        if (miss) {
            // rare and lane-local: a pattern the cache does not hold yet.  Revisit those rows in order.
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint32_t mk = miss & (0x1111u << k);
                while (mk) {
                    const int i = (__ffs(mk) - 1) >> 2;
                    mk &= mk - 1;
                    uint32_t y[4];
                    split_codons(sp + i * (kTileBytes / 4), sh, y);
                    const uint32_t v = y[k] & C.pm[k];
                    const uint32_t n = (C.nalt >> (2 * k)) & 3u;
                    if (n >= 1 && v == C.a1[k]) C.cnt[k] += 1u;
                    else if (n >= 2 && v == C.a2[k]) C.cnt[k] += 0x10000u;
                    else if (n == 0) { C.a1[k] = v; C.cnt[k] = (C.cnt[k] & 0xFFFF0000u) | 1u; C.nalt += 1u << (2 * k); }
                    else if (n == 1) { C.a2[k] = v; C.cnt[k] = (C.cnt[k] & 0x0000FFFFu) | 0x10000u; C.nalt += 1u << (2 * k); }
                    else count_pattern(k == 3 ? v >> 8 : v, (uint32_t)(codon0 + dir * (4 * gl + k)), 1u, minus, hist, other_min, other_max,
                                       S.lut_cls, S.lut_chr);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.empty[stage]);      // this warp is done with the stage's bytes
        if (++stage == kTileStages) { stage = 0; phase ^= 1; }
    }
    flush();
}

}  // namespace

size_t tile_hist_smem_bytes() { return sizeof(TileSmem) + 1024; }

// The tensor map views `rows` rows of `row_bytes` readable bytes each, `pitch` apart, as a u8
// matrix; boxes are kTileBytes x kTileRows with zero fill outside.  Returns false when the driver
// entry point is missing or rejects the geometry (the caller then uses k_fused_hist).
bool make_tile_tensor_map(void *out_map, const uint8_t *d_aln, int64_t pitch, int64_t rows, int64_t row_bytes, int l2_promotion)
{
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeFn>(p);
    }();
    if (!encode) return false;
    if (((uintptr_t)d_aln & 15) || (pitch & 15) || pitch <= 0 || rows <= 0 || row_bytes <= 0 || pitch >= ((int64_t)1 << 40)) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)row_bytes, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)pitch};
    const cuuint32_t box[2] = {(cuuint32_t)kTileBytes, (cuuint32_t)kTileRows};
    const cuuint32_t estride[2] = {1, 1};
    const CUtensorMapL2promotion promo = l2_promotion == 0   ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                         : l2_promotion == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                         : l2_promotion == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                                             : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    const CUresult r = encode(static_cast<CUtensorMap *>(out_map), CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t *>(d_aln), gdim,
                              gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

cudaError_t launch_tile_hist(const void *tensor_map, int64_t n_rows, const uint8_t *d_cons, const CodonTile *d_tiles, int64_t n_tiles,
                             int n_sms, uint32_t *d_sched, uint32_t *d_hist, uint32_t *d_other_min, uint32_t *d_other_max,
                             cudaStream_t stream)
{
    if (n_tiles <= 0 || n_rows <= 0) return cudaSuccess;
    const size_t smem = tile_hist_smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(k_tile_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // per device
    if (e != cudaSuccess) return e;
    const int64_t n_rb = (n_rows + kTileRows - 1) / kTileRows;
    // chunk = kTileChunkRbs blocks of rows of one tile; shorter when there are too few chunks to balance the SMs
    int64_t chunk_rbs = kTileChunkRbs;
    int row_major = 1;
    if (const char *v = getenv("SNPG_TILE_CHUNK")) chunk_rbs = std::max(1, atoi(v));        // tuning knobs
    if (const char *v = getenv("SNPG_TILE_ORDER")) row_major = atoi(v);
    int debug_skip = 0;
    if (const char *v = getenv("SNPG_TILE_SKIP")) debug_skip = atoi(v);                       // timing experiments only: wrong results
    while (chunk_rbs > 1 && n_tiles * ((n_rb + chunk_rbs - 1) / chunk_rbs) < 8 * (int64_t)kTileCtasPerSm * n_sms) chunk_rbs >>= 1;
    const int64_t chunks = n_tiles * ((n_rb + chunk_rbs - 1) / chunk_rbs);
    const unsigned grid = (unsigned)std::min<int64_t>(chunks, (int64_t)kTileCtasPerSm * std::max(1, n_sms));
    k_tile_hist<<<grid, (kTileConsumers + 1) * 32, smem, stream>>>(*static_cast<const CUtensorMap *>(tensor_map), (int)n_rows, d_cons,
                                                                   d_tiles, (int)n_tiles, (int)n_rb, (int)chunk_rbs, row_major, debug_skip, d_sched, d_hist,
                                                                   d_other_min, d_other_max);
    return cudaGetLastError();
}

}  // namespace snpg

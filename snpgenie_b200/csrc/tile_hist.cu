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
// the 52-word row pitch), XORs it with the group's words of the consensus row and, when anything
// differs, parks the three XOR words in a per-warp ring.  When 32 groups are parked the full warp
// decodes them densely: which of the four codons differ, their residues (XOR back with the
// consensus), classification through the 256-entry LUT, and one RED.ADD per differing codon into
// hist[codon][code].  k_fused_fixup then credits each codon's consensus code with
// (rows - recorded), exactly as for k_fused_hist.  Bytes are compared raw, so a lower-case or 'U'
// residue simply takes the decode path.
//
// (A density-independent variant -- a per-lane cache of the consensus and two alternate patterns
// per codon with register counters, no queues -- was measured too: 98 M warp instructions, ALU
// pipe 78 % busy, 122 us on config 2; see profiles/.)
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
#include "snpg_tma.cuh"

namespace snpg {

namespace {

constexpr int kTileCodons = 64;
constexpr int kTileBytes = 3 * kTileCodons + 16;         // 208: box width; the tile's 192 bytes start 0..15 bytes in
constexpr int kTileStages = 3;                           // per CTA; two CTAs share an SM
constexpr int kTileConsumers = 16;                       // consumer warps; warp kTileConsumers is the producer
constexpr int kTileRowsPerWarp = kTileRows / kTileConsumers;   // 8 rows = 4 row pairs per stage
constexpr int kTileStageBytes = kTileRows * kTileBytes;  // 26,624
constexpr int kTileQueue = 64;                           // parked groups per warp (ring, power of two)
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
    uint4 queue[kTileConsumers][kTileQueue];             // parked (d0, d1, d2, group)
    uint4 cons[kTileConsumers][16];                      // per warp: the current tile's consensus words per group
    uint8_t lut_chr[2][256];
    uint8_t lut_cls[2][256];
    StageMeta meta[kTileStages];
    uint64_t full[kTileStages];
    uint64_t empty[kTileStages];
};

// One parked group: d0..d2 = (row ^ consensus) of the group's 12 bytes (masked to its codons),
// g = the group's index inside the tile.  Codon k of the group sits in bytes 3k..3k+2.
__device__ __forceinline__ void tile_decode(uint4 e, const uint4 *__restrict__ wcons, int codon0, int dir, int minus,
                                            uint32_t *__restrict__ hist, uint32_t *__restrict__ other_min,
                                            uint32_t *__restrict__ other_max, const uint8_t (*lut_cls)[256],
                                            const uint8_t (*lut_chr)[256])
{
    const uint32_t d0 = e.x, d1 = e.y, d2 = e.z, g = e.w;
    const uint4 c = wcons[g];
    uint32_t cm = ((d0 & 0x00FFFFFFu) ? 1u : 0u) | ((((d0 >> 24) | (d1 << 8)) & 0x00FFFFFFu) ? 2u : 0u) |
                  ((((d1 >> 16) | (d2 << 16)) & 0x00FFFFFFu) ? 4u : 0u) | ((d2 >> 8) ? 8u : 0u);
    const uint32_t x0 = d0 ^ c.x, x1 = d1 ^ c.y, x2 = d2 ^ c.z;        // the row's own bytes
    // byte selectors (PRMT) of codon k out of (x0, x1) for k < 2 and (x1, x2) for k >= 2; the '-' strand
    // reads the codon right to left, so its selectors are reversed
    const unsigned long long sel_all = minus ? 0x0567023403450012ull : 0x0765043205430210ull;
    const uint8_t *cls = lut_cls[minus];
    const uint8_t *chr = lut_chr[minus];
    const int cbase = codon0 + dir * 4 * (int)g;
    while (cm) {
        const int k = __ffs(cm) - 1;
        cm &= cm - 1;
        const uint32_t v = __byte_perm(k < 2 ? x0 : x1, k < 2 ? x1 : x2, (uint32_t)(sel_all >> (16 * k)));
        const uint32_t b0 = v & 0xFF, b1 = __byte_perm(v, 0, 0x4441), b2 = __byte_perm(v, 0, 0x4442);
        const uint32_t k0 = cls[b0], k1 = cls[b1], k2 = cls[b2];
        const uint32_t cidx = (uint32_t)(cbase + dir * k);
        uint32_t code = k0 * 16 + k1 * 4 + k2;
        if (max(k0, max(k1, k2)) >= 4) {
            code = (k0 == 4 || k1 == 4 || k2 == 4) ? kCodeUndef : kCodeOther;
            if (code == kCodeOther) {
                const uint32_t key = ((uint32_t)chr[b0] << 16) | ((uint32_t)chr[b1] << 8) | chr[b2];
                atomicMin(&other_min[cidx], key);
                atomicMax(&other_max[cidx], key);
            }
        }
        atomicAdd(&hist[cidx * (uint32_t)kHistBins + code], 1u);     // 32-bit index: codons * 66 < 2^32
    }
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
            if (atomicAdd(&sched[1], 1u) == gridDim.x - 1) { sched[0] = 0; sched[1] = 0; sched[2] = 0; __threadfence(); }
        }
        return;
    }

    // -------------------- consumers --------------------
    uint4 *queue = S.queue[warp];
    uint4 *wcons = S.cons[warp];
    uint32_t q_head = 0, q_count = 0;            // warp-uniform
    unsigned lt = (1u << lane) - 1;
    asm volatile("" : "+r"(lt));                 // keep it in a register (ptxas otherwise re-reads %tid through the MIO queue)
    const int gl = lane & 15, half = lane >> 4;
    int codon0 = 0, dir = 1, minus = 0;
    uint32_t c0 = 0, c1 = 0, c2 = 0, m0 = 0, m1 = 0, m2 = 0;
    uint32_t lane_ofs = 0, sh = 0;               // byte offset of the lane's first aligned word in its row; funnel shift in bits

    for (;;) {
        mbar_wait(&S.full[stage], phase);
        const int4 mt = *reinterpret_cast<const int4 *>(&S.meta[stage]);
        if (mt.w < 0) break;
        if (mt.z >> 24) {
            // a new chunk: the parked groups belong to the previous tile, decode them before its constants go
            if (q_count) {
                __syncwarp();
                if ((uint32_t)lane < q_count)
                    tile_decode(queue[(q_head + lane) & (kTileQueue - 1)], wcons, codon0, dir, minus, hist, other_min, other_max,
                                S.lut_cls, S.lut_chr);
                __syncwarp();
                q_head = 0; q_count = 0;
            }
            codon0 = mt.y; dir = (int)(int8_t)(mt.z >> 8); minus = (mt.z >> 16) & 1;
            const int cnt = min(4, max(0, (mt.z & 0xFF) - 4 * gl));
            const int nb = 3 * cnt;
            m0 = nb >= 4 ? 0xFFFFFFFFu : (nb == 3 ? 0x00FFFFFFu : 0u);
            m1 = nb >= 8 ? 0xFFFFFFFFu : (nb == 6 ? 0x0000FFFFu : 0u);
            m2 = nb >= 12 ? 0xFFFFFFFFu : (nb == 9 ? 0x000000FFu : 0u);
            lane_ofs = (uint32_t)(((mt.x & 15) >> 2) + 3 * gl) * 4u;
            sh = (uint32_t)(mt.x & 3) * 8u;
            const uint32_t *cw = reinterpret_cast<const uint32_t *>(S.cons_row[stage] + lane_ofs);
            const uint32_t w0 = cw[0], w1 = cw[1], w2 = cw[2], w3 = cw[3];
            c0 = __funnelshift_r(w0, w1, sh) & m0;
            c1 = __funnelshift_r(w1, w2, sh) & m1;
            c2 = __funnelshift_r(w2, w3, sh) & m2;
            if (half == 0) wcons[gl] = make_uint4(c0, c1, c2, 0);
            __syncwarp();
        }

        // lanes 0-15 read row i of the warp's eight, lanes 16-31 row i + 4
        const uint32_t *sp = reinterpret_cast<const uint32_t *>(S.stage[stage] + (warp * kTileRowsPerWarp + 4 * half) * kTileBytes + lane_ofs);
        const int rows_left = debug_skip ? 0 : n_rows - (mt.w * kTileRows + warp * kTileRowsPerWarp + 4 * half);   // row i is real iff i < rows_left
        uint32_t w[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
#pragma unroll
            for (int k = 0; k < 4; k++) w[i][k] = sp[i * (kTileBytes / 4) + k];
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t x0 = __funnelshift_r(w[i][0], w[i][1], sh), x1 = __funnelshift_r(w[i][1], w[i][2], sh),
                           x2 = __funnelshift_r(w[i][2], w[i][3], sh);
            const uint32_t d0 = (x0 ^ c0) & m0, d1 = (x1 ^ c1) & m1, d2 = (x2 ^ c2) & m2;
            const bool differs = (d0 | d1 | d2) != 0 && i < rows_left;
            const unsigned bal = __ballot_sync(kFull, differs);
            if (bal) {       // on realistic alignments most row pairs have nothing to park
                if (differs) queue[(q_head + q_count + __popc(bal & lt)) & (kTileQueue - 1)] = make_uint4(d0, d1, d2, (uint32_t)gl);
                q_count += __popc(bal);
                while (q_count >= 32) {      // at most 31 + 32 parked groups at any time
                    __syncwarp();
                    tile_decode(queue[(q_head + lane) & (kTileQueue - 1)], wcons, codon0, dir, minus, hist, other_min, other_max,
                                S.lut_cls, S.lut_chr);
                    __syncwarp();
                    q_head += 32;
                    q_count -= 32;
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.empty[stage]);      // this warp is done with the stage's bytes
        if (++stage == kTileStages) { stage = 0; phase ^= 1; }
    }
    __syncwarp();
    if ((uint32_t)lane < q_count)
        tile_decode(queue[(q_head + lane) & (kTileQueue - 1)], wcons, codon0, dir, minus, hist, other_min, other_max, S.lut_cls,
                    S.lut_chr);
}

}  // namespace

size_t tile_hist_smem_bytes() { return sizeof(TileSmem) + 1024; }

// The tensor map views `rows` rows of `row_bytes` readable bytes each, `pitch` apart, as a u8
// matrix; boxes are kTileBytes x kTileRows with zero fill outside.  Returns false when the driver
// entry point is missing or rejects the geometry (the caller then uses k_fused_hist).
bool make_tile_tensor_map(void *out_map, const uint8_t *d_aln, int64_t pitch, int64_t rows, int64_t row_bytes, int l2_promotion)
{
    const TensorMapEncodeFn encode = tensor_map_encoder();
    if (!encode) return false;
    if (((uintptr_t)d_aln & 15) || (pitch & 15) || pitch <= 0 || rows <= 0 || row_bytes <= 0 || pitch >= ((int64_t)1 << 40)) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)row_bytes, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)pitch};
    const cuuint32_t box[2] = {(cuuint32_t)kTileBytes, (cuuint32_t)kTileRows};
    const cuuint32_t estride[2] = {1, 1};
    const CUtensorMapL2promotion promo = tensor_map_promotion(l2_promotion);
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

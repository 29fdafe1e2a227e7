// K1+K2 fused, TMA-fed span kernel: per-codon histograms straight from the alignment bytes (SNPG_FUSED_TMA_SPANS).
//
// Replaces, like k_fused_hist (kernels.cu), the codon split (snpgenie_within_group.pl:417-454), the polymorphism
// pre-check (wg:469-541; bg:472-544) and the O(n^2) pair counting (wg:596-609; bg:649-667): the 66 counts of a codon
// column determine every per-codon output.
//
// The arithmetic is k_fused_hist's: a lane owns one 16-byte column span, compares every row of it (plus the two bytes a
// codon may reach into the next span) with three reference rows -- consensus, second and third pattern of the span
// (k_vote_refs) --, counts the rows that equal the second or third, parks the rows that equal none, and a full warp
// decodes 32 parked spans at a time into RED.ADDs on hist[codon][code].  What differs is how the rows arrive.  In
// k_fused_hist a warp issues four 128-bit global loads and stalls on the first use (31 % of its stall samples;
// registers bound the bytes in flight).  Here every warp owns a small pipeline of shared-memory stages that its lane 0
// keeps filled with TMA tensor loads (cp.async.bulk.tensor.2d: one 512-byte x R-row box per stage = 31 spans + the
// look-ahead of the last), STAGES - 1 stages ahead of the rows being compared; the warp reads a stage with conflict-free
// 128-bit shared-memory loads.  Warps never wait for each other (a first version with a producer warp per block and
// stages shared by eight consumer warps lost 15 % of its time waiting for the slowest warp of a stage).
//
// The parked spans live in a per-warp ring of two arrays (16-byte vectors; a 4-byte tag = the two look-ahead bytes, the
// lane and the low bits of the span block): parking stores and the decode's 32-entry fetches are conflict-free.
//
// Measured on config 2 (profiles/): 84.5 us against 75.9 us for k_fused_hist -- IPC rises from 2.1 to 2.4-2.7, but the
// per-stage bookkeeping costs more instructions than the load stalls it removes, so this kernel is an option, not the
// default.
// Algorithmic traffic: every alignment byte of the CDS span read once (3 B per (sequence, codon)) + 264 B per codon.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstddef>
#include <cstdlib>

#include "snpg_device.cuh"
#include "snpg_internal.h"
#include "snpg_span.cuh"
#include "snpg_tma.cuh"

namespace snpg {

namespace {

constexpr int kSpanWarps = 8;                 // warps per block, each with its own TMA pipeline and ring
constexpr int kSpanSpans = 31;                // spans per block of spans (lane 31 only supplies the look-ahead)
constexpr int kSpanBoxBytes = 512;            // 31 spans + 16 bytes
constexpr int kSpanBoxWords = kSpanBoxBytes / 4;
constexpr uint32_t kSbTagBits = 11;           // low bits of the span block carried by a parked entry
constexpr int kSpanTailRows = 4;              // rows per chunk in the fine-grained tail region

// What a warp's lane 0 notes about the stage it asked for (read back by the whole warp when the bytes have landed).
struct SpanStageMeta { uint32_t sb, n_rows, last, pad; };     // n_rows = 0 ends the warp's work; last: the chunk ends with this stage

template <int STAGES, int R, int QUEUE>
struct SpanSmem {
    struct Warp {
        uint8_t stage[STAGES][R * kSpanBoxBytes];              // TMA destinations (multiples of 128 bytes)
        uint4 ring_a[QUEUE];                                   // parked spans: their 16 bytes
        uint32_t ring_b[QUEUE];                                // look-ahead bytes | lane << 16 | (span block & 2047) << 21
        SpanStageMeta meta[STAGES];
        uint64_t full[STAGES];
        uint8_t pad[128 - (STAGES * 24 + QUEUE * 4) % 128];
    } w[kSpanWarps];
    uint8_t lut_chr[2 * 256];
    uint8_t lut_cls[2 * 256];
};

// One parked (span, row): decode and count exactly the codons that contain a byte differing from the consensus.
// ea / eb = shared-memory addresses of the entry's vector and tag; sb_hi = the high bits of its span block.
__device__ __forceinline__ void span_decode(uint32_t ea, uint32_t eb, uint32_t sb_hi, const uint8_t *__restrict__ cons,
                                            const SpanRun *__restrict__ runs, uint32_t *__restrict__ hist,
                                            uint32_t *__restrict__ other_min, uint32_t *__restrict__ other_max,
                                            uint32_t lut_cls, uint32_t lut_chr)
{
    uint4 v;
    uint32_t tag;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(ea) : "memory");
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tag) : "r"(eb) : "memory");
    const uint32_t sp = (sb_hi | (tag >> 21)) * (uint32_t)kSpanSpans + ((tag >> 16) & 31u);
    const uint2 *rp = reinterpret_cast<const uint2 *>(runs + (size_t)sp * kSpanRuns);
    const uint4 c = __ldg(reinterpret_cast<const uint4 *>(cons + (size_t)sp * 16));
    const uint32_t c4 = __ldg(reinterpret_cast<const uint32_t *>(cons + (size_t)sp * 16 + 16));
    uint2 rw = __ldg(rp);                                             // the first run, asked for with the consensus
    const uint32_t m20 = diff_mask20(v, tag & 0xFFFFu, c, c4 & 0xFFFFu);
    // byte x of the 18-byte window sits at ea + x for x < 16 and at eb + x - 16 beyond
    const uint32_t hop = eb - ea - 16u;
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
            const uint32_t s_lo = mi ? 0u : 4u, s_hi = mi ? 4u : 0u;     // '-' strand reads the codon right to left
            const uint32_t base = rw.x * (uint32_t)kHistBins;
            const int step = dir * (kHistBins / 3);                      // per byte of offset: j = 3 * codon ordinal
            do {
                const int j = __ffs(cm) - 1;
                cm &= cm - 1;
                const uint32_t x = off + (uint32_t)j;                    // first byte of the codon, <= 15
                const uint32_t r0 = lds_u8(ea + x);
                const uint32_t r1 = lds_u8(ea + x + 1 + ((x + 1) >> 4) * hop), r2 = lds_u8(ea + x + 2 + ((x + 2) >> 4) * hop);
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

// mbarrier / TMA operations on 32-bit shared-memory addresses (the kernel keeps ONE base address per warp in a register and
// adds compile-time offsets: with generic pointers ptxas re-derives the shared window from %cgaid for every use)
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_s(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_s(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d_s(uint32_t dst, const CUtensorMap *map, uint32_t bar, int32_t x, int32_t y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}

// Every warp runs its own pipeline: lane 0 claims chunks -- (block of 31 spans) x (chunk_a rows, or kSpanTailRows rows
// over the last rows) -- from the global counter sched[0] and keeps STAGES - 1 TMA loads of R rows in flight ahead of the
// stage the warp is working on; no warp ever waits for another.  Parked spans are decoded after every DRAIN_EVERY rows,
// so the ring needs 31 + 32 * DRAIN_EVERY entries (QUEUE, a power of two).  inv_sb = floor(2^32 / n_sb).
template <int STAGES, int R, int DRAIN_EVERY, int QUEUE, int MIN_BLOCKS>
__global__ void __launch_bounds__(kSpanWarps * 32, MIN_BLOCKS)
k_span_hist(const __grid_constant__ CUtensorMap tmap, uint32_t n_rows, const uint8_t *__restrict__ cons, uint32_t *__restrict__ alt,
            uint32_t n_spans, const SpanRun *__restrict__ runs, uint32_t *__restrict__ hist, uint32_t *__restrict__ other_min,
            uint32_t *__restrict__ other_max, uint32_t *__restrict__ sched, uint32_t n_sb, uint32_t inv_sb, uint32_t chunk_a,
            uint32_t rows_a, uint32_t k_a, uint32_t k_total)
{
    static_assert(QUEUE >= 31 + 32 * DRAIN_EVERY && (QUEUE & (QUEUE - 1)) == 0 && R % DRAIN_EVERY == 0, "ring too small");
    static_assert(kSpanTailRows % R == 0 || R % kSpanTailRows == 0, "tail chunks are whole stages or fractions of one");
    using Smem = SpanSmem<STAGES, R, QUEUE>;
    using WarpSmem = typename Smem::Warp;
    constexpr uint32_t kStageBytes = R * kSpanBoxBytes;
    constexpr uint32_t kOfsRingA = offsetof(WarpSmem, ring_a), kOfsRingB = offsetof(WarpSmem, ring_b), kOfsMeta = offsetof(WarpSmem, meta);
    constexpr uint32_t kOfsFull = offsetof(WarpSmem, full);
    static_assert(sizeof(WarpSmem) % 128 == 0 && kOfsRingA % 16 == 0, "per-warp blocks keep the TMA destinations aligned");
    extern __shared__ uint8_t span_smem_raw[];
    Smem &S = *reinterpret_cast<Smem *>(span_smem_raw + ((128u - (smem_u32(span_smem_raw) & 127u)) & 127u));
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;

    for (int s = 0; s < 2; s++) {
        const uint8_t n = norm_char((uint8_t)t, s);
        S.lut_chr[s * 256 + t] = n;
        S.lut_cls[s * 256 + t] = class_of(n);
    }
    if (lane == 0) {
        for (int s = 0; s < STAGES; s++) mbar_init(&S.w[warp].full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // every shared-memory address below is w0 (this warp's block) or lut0 + a compile-time offset + a stage offset
    uint32_t w0 = smem_u32(&S.w[warp]);
    const uint32_t lut0 = smem_u32(S.lut_chr);                   // classes follow 512 bytes further on
    uint32_t q_head = 0, q_tail = 0;             // entry counts, warp-uniform, free-running
    unsigned lt = (1u << lane) - 1;
    uint32_t cur_sb = 0xFFFFFFFFu;               // the block of spans the references in registers belong to
    bool live = false;                           // this lane owns a span of it
    uint4 c = make_uint4(0, 0, 0, 0), a = c, b = c;
    uint32_t c4 = 0, a4 = 0, b4 = 0;             // (the look-ahead words carry the lane number in bits 16..20)
    uint32_t hits = 0;                           // bit u: row u of the stage equalled the second reference; bit 16 + u: the third
    uint32_t n_a = 0, n_b = 0;
    const uint32_t lane16 = (uint32_t)lane << 16;
    uint32_t one = 1u + (uint32_t)lane;          // 1 in lane 0, the only lane that claims

    // ---- the issuing side (lane 0): a stream of stages over the claimed chunks
    uint32_t iss_sb = 0, iss_row = 0, iss_left = 0;      // the chunk being issued: its span block, next row, rows left
    uint32_t k_next = 0;                                 // the chunk claimed ahead (>= k_total: none left)
    if (lane == 0) k_next = claim_async(sched, one);
    auto issue = [&](uint32_t stage) {       // lane 0: ask for the next rows of the stream into `stage` (or post the end marker)
        const uint32_t bar = w0 + kOfsFull + stage * 8u;
        if (iss_left == 0 && k_next < k_total) {
            const uint32_t k = k_next;
            k_next = claim_async(sched, one);
            const bool in_a = k < k_a;
            uint32_t rb;
            div_mod(in_a ? k : k - k_a, n_sb, inv_sb, rb, iss_sb);
            iss_row = in_a ? rb * chunk_a : rows_a + rb * kSpanTailRows;
            iss_left = in_a ? chunk_a : min((uint32_t)kSpanTailRows, n_rows - iss_row);
        }
        if (iss_left == 0) {     // nothing left: a stage of no rows ends the consuming side
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(w0 + kOfsMeta + stage * 16u), "r"(0u), "r"(0u), "r"(1u), "r"(0u) : "memory");
            mbar_arrive_s(bar);
        } else {
            const uint32_t nr = min((uint32_t)R, iss_left);
            iss_left -= nr;
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(w0 + kOfsMeta + stage * 16u), "r"(iss_sb), "r"(nr), "r"(iss_left == 0 ? 1u : 0u), "r"(0u) : "memory");
            // (the stage was read through the generic proxy: order those reads before the asynchronous proxy's writes)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx_s(bar, kStageBytes);
            tma_load_2d_s(w0 + stage * kStageBytes, &tmap, bar, (int32_t)(iss_sb * (kSpanSpans * 4)), (int32_t)iss_row);
            iss_row += nr;
        }
    };

    auto decode_at = [&](uint32_t idx) {
        const uint32_t slot = idx & (QUEUE - 1);
        // (the parked entries belong to span blocks with cur_sb's high bits: the ring is flushed before those change)
        span_decode(w0 + kOfsRingA + slot * 16u, w0 + kOfsRingB + slot * 4u, cur_sb & ~((1u << kSbTagBits) - 1u), cons, runs, hist, other_min,
                    other_max, lut0 + 512u, lut0);
    };
    auto drain = [&]() {     // at most 31 + 32 * DRAIN_EVERY parked entries between drains
        while (q_tail - q_head >= 32u) {
            __syncwarp();
            decode_at(q_head + lane);
            __syncwarp();
            q_head += 32u;
        }
    };
    auto flush_ring = [&]() {
        drain();
        __syncwarp();
        if ((uint32_t)lane < q_tail - q_head) decode_at(q_head + lane);
        __syncwarp();
        q_head = q_tail;
    };
    // the counts of the rows that equalled the second / third reference of the chunk just finished
    auto flush_counts = [&]() {
        if (live && (n_a | n_b)) {
            uint32_t *ap = alt;
            uint32_t ns = n_spans;
            asm volatile("" : "+l"(ap), "+r"(ns));   // (the plane addresses are derived here, not kept across the row loop)
            const RefPlanes refs = ref_planes(ap, ns);
            const uint32_t sp = cur_sb * kSpanSpans + lane;
            if (n_a) atomicAdd(&refs.cnt_a[sp], n_a);
            if (n_b) atomicAdd(&refs.cnt_b[sp], n_b);
        }
        n_a = 0; n_b = 0;
    };
    // one row of the warp: compare with the three references (pairwise distinct: at most one matches), note or park
    auto row = [&](const uint4 &v, int u) {
        const uint32_t v4 = (__shfl_down_sync(kFull, v.x, 1) & 0xFFFFu) | lane16;
        const uint32_t dc = (v.x ^ c.x) | (v.y ^ c.y) | (v.z ^ c.z) | (v.w ^ c.w) | (v4 ^ c4);
        const uint32_t da = (v.x ^ a.x) | (v.y ^ a.y) | (v.z ^ a.z) | (v.w ^ a.w) | (v4 ^ a4);
        const uint32_t db = (v.x ^ b.x) | (v.y ^ b.y) | (v.z ^ b.z) | (v.w ^ b.w) | (v4 ^ b4);
        if (da == 0) hits |= 1u << u;
        if (db == 0) hits |= 0x10000u << u;
        const bool differs = live && dc != 0 && da != 0 && db != 0;
        const unsigned bal = __ballot_sync(kFull, differs);
        // park the bytes (predicated stores, no branch: some lane parks on nearly every row of the benchmark's diversity); a
        // full warp decodes 32 parked spans at a time
        const uint32_t slot = (q_tail + __popc(bal & lt)) & (QUEUE - 1);
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.u32 p, %0, 0;\n\t"
            "@p st.shared.v4.u32 [%1 + %2], {%3, %4, %5, %6};\n\t"
            "@p st.shared.u32 [%7 + %8], %9;\n\t"
            "}" ::"r"((uint32_t)differs), "r"(w0 + slot * 16u), "n"(kOfsRingA), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(w0 + slot * 4u),
            "n"(kOfsRingB), "r"(v4 | (cur_sb << 21)) : "memory");
        q_tail += __popc(bal);
    };

    if (lane == 0) {
#pragma unroll 1
        for (uint32_t s = 0; s < STAGES - 1; s++) issue(s);
    }
    for (uint32_t it = 0;; it++) {
        // pinned in registers: ptxas otherwise re-derives them (from %tid, %cgaid) for every row
        asm volatile("" : "+r"(lt), "+r"(w0));
        const uint32_t stage = it % STAGES, phase = (it / STAGES) & 1u;
        if (lane == 0) issue((it + STAGES - 1) % STAGES);        // STAGES - 1 stages ahead, into the stage consumed last
        mbar_wait_s(w0 + kOfsFull + stage * 8u, phase);
        const uint4 mt = lds_v4(w0 + kOfsMeta + stage * 16u);
        if (mt.y == 0) break;
        if (mt.x != cur_sb) {
            // another block of spans: its references come in (the ring keeps entries of span blocks with the same high bits)
            if ((mt.x ^ cur_sb) >> kSbTagBits) flush_ring();
            cur_sb = mt.x;
            const uint32_t span = cur_sb * kSpanSpans + lane;
            uint32_t *ap = alt;
            uint32_t ns = n_spans;
            asm volatile("" : "+l"(ap), "+r"(ns));
            live = lane < kSpanSpans && span < ns;
            const uint32_t sp = min(span, ns - 1);                 // the references of a dead lane are never used
            const RefPlanes refs = ref_planes(ap, ns);
            c = *reinterpret_cast<const uint4 *>(cons + (size_t)sp * 16);
            c4 = (*reinterpret_cast<const uint32_t *>(cons + (size_t)sp * 16 + 16) & 0xFFFFu) | lane16;
            a = refs.a16[sp];
            b = refs.b16[sp];
            const uint32_t tails = refs.tails[sp];
            a4 = (tails & 0xFFFFu) | lane16;
            b4 = (tails >> 16) | lane16;
        }
        const uint32_t sbase = w0 + stage * kStageBytes + (uint32_t)lane * 16u;
        if (mt.y == R) {
#pragma unroll
            for (int u0 = 0; u0 < R; u0 += DRAIN_EVERY) {
                uint4 v[DRAIN_EVERY];
#pragma unroll
                for (int u = 0; u < DRAIN_EVERY; u++) v[u] = lds_v4(sbase + (u0 + u) * kSpanBoxBytes);
#pragma unroll
                for (int u = 0; u < DRAIN_EVERY; u++) row(v[u], u0 + u);
                drain();
            }
        } else {
            for (uint32_t u = 0; u < mt.y; u++) {
                row(lds_v4(sbase + u * kSpanBoxBytes), (int)u);
                drain();
            }
        }
        n_a += __popc(hits & 0xFFFFu);
        n_b += __popc(hits >> 16);
        hits = 0;
        if (mt.z) flush_counts();                                // the chunk is complete
        __syncwarp();                                            // every lane has read the stage: lane 0 may refill it
    }
    flush_ring();
    __syncthreads();
    if (t == 0 && atomicAdd(&sched[1], 1u) == gridDim.x - 1) { sched[0] = 0; sched[1] = 0; sched[2] = 0; }
}

template <int STAGES, int R, int DRAIN_EVERY, int QUEUE, int MIN_BLOCKS>
struct SpanConfig {
    static constexpr int stage_rows = R;
    static size_t smem() { return sizeof(SpanSmem<STAGES, R, QUEUE>) + 128; }
    static auto kernel() { return k_span_hist<STAGES, R, DRAIN_EVERY, QUEUE, MIN_BLOCKS>; }
};

template <class Cfg>
cudaError_t launch_span_cfg(const CUtensorMap &tmap, int64_t n_rows, const uint8_t *d_cons, uint32_t *d_alt, int64_t n_spans, const SpanRun *d_runs,
                            uint32_t *d_hist, uint32_t *d_other_min, uint32_t *d_other_max, uint32_t *d_sched, int n_sms, int chunk_rows_env,
                            int tail_env, cudaStream_t stream)
{
    auto kernel = Cfg::kernel();
    const size_t smem = Cfg::smem();
    static const int blocks_per_sm = [&] {
        int nb = 0;
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, kSpanWarps * 32, smem) != cudaSuccess || nb < 1) nb = 1;
        if (const char *v = getenv("SNPG_SPAN_BLOCKS")) nb = std::max(1, std::min(nb, atoi(v)));
        return nb;
    }();
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // per device
    if (e != cudaSuccess) return e;
    const int64_t n_sb = (n_spans + kSpanSpans - 1) / kSpanSpans;
    int64_t grid = (int64_t)n_sms * blocks_per_sm;
    // chunk of rows a warp claims at a time (the vector-load kernel's sweep: 32 rows).  When every warp gets at least two
    // chunks, the last tail % of the rows go out in chunks of kSpanTailRows to even out the finish; a small launch (a staged
    // block of a host buffer) is one chunk per warp and only the rows left over by the chunk size go out small.
    const int64_t share = n_rows * n_sb / (grid * kSpanWarps);
    const int64_t chunk_a = chunk_rows_env > 0 ? chunk_rows_env : 32;
    const int64_t tail_pct = tail_env >= 0 ? tail_env : 2;
    const int64_t rows_a = (share >= 2 * chunk_a ? n_rows - n_rows * tail_pct / 100 : n_rows) / chunk_a * chunk_a;
    const int64_t k_a = rows_a / chunk_a * n_sb;
    const int64_t k_b = (n_rows - rows_a + kSpanTailRows - 1) / kSpanTailRows * n_sb;
    if (k_a + k_b >= ((int64_t)1 << 32) - 65536) return cudaErrorInvalidValue;      // (the caller takes the vector-load kernel)
    grid = std::max<int64_t>(1, std::min<int64_t>(grid, (k_a + k_b + kSpanWarps - 1) / kSpanWarps));
    kernel<<<(unsigned)grid, kSpanWarps * 32, smem, stream>>>(
        tmap, (uint32_t)n_rows, d_cons, d_alt, (uint32_t)n_spans, d_runs, d_hist, d_other_min, d_other_max, d_sched, (uint32_t)n_sb,
        n_sb == 1 ? 0xFFFFFFFFu : (uint32_t)(((uint64_t)1 << 32) / (uint64_t)n_sb), (uint32_t)chunk_a, (uint32_t)rows_a, (uint32_t)k_a,
        (uint32_t)(k_a + k_b));
    return cudaGetLastError();
}

}  // namespace

// The tensor map views `rows` rows of `row_bytes` readable bytes each, `pitch` apart, as a matrix of 32-bit words (a box
// is 128 words wide: the element count of a box dimension is limited to 256); boxes reach past the last row / column with
// zero fill.  Returns false when the driver entry point is missing or rejects the geometry (the caller then uses
// k_fused_hist).
bool make_span_tensor_map(void *out_map, const uint8_t *d_aln, int64_t pitch, int64_t rows, int64_t row_bytes, int l2_promotion)
{
    const TensorMapEncodeFn encode = tensor_map_encoder();
    if (!encode) return false;
    if (((uintptr_t)d_aln & 15) || (pitch & 15) || pitch <= 0 || rows <= 0 || row_bytes <= 0 || pitch >= ((int64_t)1 << 40)) return false;
    const int64_t words = std::min<int64_t>((row_bytes + 3) / 4, pitch / 4);
    if (words >= ((int64_t)1 << 32) || rows >= ((int64_t)1 << 32)) return false;
    static const int box_rows = span_stage_rows();
    const cuuint64_t gdim[2] = {(cuuint64_t)words, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)pitch};
    const cuuint32_t box[2] = {(cuuint32_t)kSpanBoxWords, (cuuint32_t)box_rows};
    const cuuint32_t estride[2] = {1, 1};
    const CUresult r = encode(static_cast<CUtensorMap *>(out_map), CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint8_t *>(d_aln), gdim,
                              gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              tensor_map_promotion(l2_promotion), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// Kernel shape (SNPG_SPAN_SHAPE, experiments): stages x rows per stage of a warp's pipeline.  0 (default) = 2 x 4;
// 1 = 4 x 2; 2 = 3 x 2; 3 = 4 x 2 at 3 blocks/SM.
static int span_shape() {
    static const int shape = [] { const char *v = getenv("SNPG_SPAN_SHAPE"); return v ? std::max(0, std::min(3, atoi(v))) : 0; }();
    return shape;
}
using SpanShape0 = SpanConfig<2, 4, 2, 128, 4>;
using SpanShape1 = SpanConfig<4, 2, 2, 128, 4>;
using SpanShape2 = SpanConfig<3, 2, 2, 128, 4>;
using SpanShape3 = SpanConfig<4, 2, 2, 128, 3>;

int span_stage_rows() {
    switch (span_shape()) {
    case 1: return SpanShape1::stage_rows;
    case 2: return SpanShape2::stage_rows;
    case 3: return SpanShape3::stage_rows;
    default: return SpanShape0::stage_rows;
    }
}

cudaError_t launch_span_hist(const void *tensor_map, int64_t n_rows, const uint8_t *d_cons, uint32_t *d_alt, int64_t n_spans,
                             const SpanRun *d_runs, uint32_t *d_hist, uint32_t *d_other_min, uint32_t *d_other_max, uint32_t *d_sched,
                             int n_sms, cudaStream_t stream)
{
    if (n_spans <= 0 || n_rows <= 0) return cudaSuccess;
    static const int env_chunk = [] { const char *v = getenv("SNPG_SPAN_ROWS"); return v ? std::max(4, atoi(v) & ~3) : 0; }();
    static const int env_tail = [] { const char *v = getenv("SNPG_SPAN_TAIL"); return v ? std::max(0, std::min(100, atoi(v))) : -1; }();
    const CUtensorMap &tmap = *static_cast<const CUtensorMap *>(tensor_map);
    switch (span_shape()) {
    case 1: return launch_span_cfg<SpanShape1>(tmap, n_rows, d_cons, d_alt, n_spans, d_runs, d_hist, d_other_min, d_other_max, d_sched, n_sms, env_chunk, env_tail, stream);
    case 2: return launch_span_cfg<SpanShape2>(tmap, n_rows, d_cons, d_alt, n_spans, d_runs, d_hist, d_other_min, d_other_max, d_sched, n_sms, env_chunk, env_tail, stream);
    case 3: return launch_span_cfg<SpanShape3>(tmap, n_rows, d_cons, d_alt, n_spans, d_runs, d_hist, d_other_min, d_other_max, d_sched, n_sms, env_chunk, env_tail, stream);
    default: return launch_span_cfg<SpanShape0>(tmap, n_rows, d_cons, d_alt, n_spans, d_runs, d_hist, d_other_min, d_other_max, d_sched, n_sms, env_chunk, env_tail, stream);
    }
}

}  // namespace snpg

// Device helpers shared by the two span histogram kernels (k_fused_hist in kernels.cu: rows through 128-bit global
// loads; k_span_hist in span_tma.cu: rows through TMA into shared memory) and by the kernels around them (k_vote_refs,
// the fix-up kernels).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "snpg_device.cuh"
#include "snpg_internal.h"

namespace snpg {

// The second and third reference pattern of every 16-byte span as planes over the span axis (launch_vote_refs writes
// them; a warp's 31 lanes read 16 consecutive bytes each: 4-5 cache lines per load, where span-major records cost 16):
//   a16[span], b16[span]   bytes 0..15 of the two patterns
//   tails[span]            bytes 16..17 (the two a codon may reach into the next span): second reference in the low half,
//                          third in the high half
//   cnt_a[span], cnt_b[span]  rows that equalled the pattern (zeroed by the vote, raised by the histogram kernels, spent by
//                          the fix-up).
// The three references of a span (consensus, a, b) are pairwise distinct, so a row equals at most one of them.
struct RefPlanes {
    uint4 *a16, *b16;
    uint32_t *tails, *cnt_a, *cnt_b;
};
__host__ __device__ __forceinline__ RefPlanes ref_planes(uint32_t *alt, int64_t n_spans) {
    const int64_t s = alt_stride(n_spans);
    RefPlanes r;
    r.a16 = reinterpret_cast<uint4 *>(alt);
    r.b16 = reinterpret_cast<uint4 *>(alt + 4 * s);
    r.tails = alt + 8 * s;
    r.cnt_a = alt + 9 * s;
    r.cnt_b = alt + 10 * s;
    return r;
}

// Flags of the nonzero bytes of d, gathered into bits 28..31 (byte 0 -> bit 28).
__device__ __forceinline__ uint32_t nonzero_bytes_hi(uint32_t d) {
    const uint32_t t = (((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u;
    return t * 0x00204081u;                                           // bits 7,15,23,31 -> 28..31, no carries
}

__device__ __forceinline__ uint32_t lds_u8(uint32_t a) { uint32_t r; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(r) : "r"(a) : "memory"); return r; }

// The chunk counter is bumped by one lane.  one = 1 comes in a register ptxas cannot prove warp-uniform:
// with a constant it turns the add into a warp-aggregated atomic whose result is shuffled at once, which
// makes the warp wait for the claim it meant to hide behind the current chunk.
__device__ __forceinline__ uint32_t claim_async(uint32_t *ctr, uint32_t one) {
    uint32_t r;
    asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(r) : "l"(ctr), "r"(one) : "memory");
    return r;
}
// k / d and k % d from m = floor(2^32 / d) (0xFFFFFFFF for d = 1): the estimate is at most one short.
__device__ __forceinline__ void div_mod(uint32_t k, uint32_t d, uint32_t m, uint32_t &q, uint32_t &r) {
    q = __umulhi(k, m);
    r = k - q * d;
    if (r >= d) { q++; r -= d; }
    if (r >= d) { q++; r -= d; }
}

// Byte-difference mask of a parked span against the consensus: bit b set when byte b of the 20-byte window (16 bytes of
// the span + 4 of the next; only 18 are ever looked at) differs.
__device__ __forceinline__ uint32_t diff_mask20(const uint4 &v, uint32_t v4, const uint4 &c, uint32_t c4) {
    uint32_t m20 = nonzero_bytes_hi(v4 ^ c4) >> 28;
    m20 = __funnelshift_l(nonzero_bytes_hi(v.w ^ c.w), m20, 4);
    m20 = __funnelshift_l(nonzero_bytes_hi(v.z ^ c.z), m20, 4);
    m20 = __funnelshift_l(nonzero_bytes_hi(v.y ^ c.y), m20, 4);
    m20 = __funnelshift_l(nonzero_bytes_hi(v.x ^ c.x), m20, 4);
    return m20;
}

}  // namespace snpg

// Peaks the resampling kernel is measured against, probed on the GPU the ctx runs on ("not measured --
// builder must probe"): the rate at which the SMs generate Philox4x32-10 blocks and nothing else (the integer-multiply
// issue bound of any sampler that spends one 32-bit Philox word per draw), and the FP64 tensor-core rate of
// mma.sync.m8n8k4.f64 (the contraction's instruction).  Timed with CUDA events, best of five launches.
#include <algorithm>

#include "snpg_ctx.h"
#include "snpg_team.h"

using namespace snpg;

namespace {

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

__global__ void __launch_bounds__(256) k_probe_philox(uint32_t iters, uint2 key, uint32_t *out)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (uint32_t i = 0; i < iters; i++) {
        const uint4 u = philox4x32_10(make_uint4(i, t, 7u, 11u), key);
        acc ^= u.x ^ u.y ^ u.z ^ u.w;
    }
    if (acc == 0x12345678u) out[0] = acc;      // keeps the loop alive; practically never true
}

__global__ void __launch_bounds__(256) k_probe_dmma(uint32_t iters, double *out)
{
    double c[8][2];
#pragma unroll
    for (int k = 0; k < 8; k++) { c[k][0] = 0.0; c[k][1] = 0.0; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += c[k][0] + c[k][1];
    if (s == 123.456) out[0] = s;
}

}  // namespace

extern "C" int snpg_probe_peaks(snpg_ctx *ctx, double *out) {
    if (!ctx || !out) return SNPG_E_ARG;
    TEAM_FORWARD0(ctx, snpg_probe_peaks(c, out));
    if (int rc = bind(ctx)) return rc;
    CU(ctx, ctx->s_misc.ensure(64));
    cudaEvent_t e0, e1;
    CU(ctx, cudaEventCreate(&e0));
    CU(ctx, cudaEventCreate(&e1));
    const int blocks = ctx->n_sms * 8;
    double best_philox = 0, best_dmma = 0;
    for (int rep = 0; rep < 6; rep++) {
        const uint32_t it_p = 2048, it_d = 4096;
        float ms = 0;
        cudaEventRecord(e0, ctx->stream);
        k_probe_philox<<<blocks, 256, 0, ctx->stream>>>(it_p, make_uint2(1u, 2u), ctx->s_misc.as<uint32_t>());
        cudaEventRecord(e1, ctx->stream);
        CU(ctx, cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep) best_philox = std::max(best_philox, 4.0 * blocks * 256.0 * it_p / (ms * 1e-3));
        cudaEventRecord(e0, ctx->stream);
        k_probe_dmma<<<blocks, 256, 0, ctx->stream>>>(it_d, ctx->s_misc.as<double>());
        cudaEventRecord(e1, ctx->stream);
        CU(ctx, cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep) best_dmma = std::max(best_dmma, 512.0 * 8 * it_d * (blocks * 8.0) / (ms * 1e-3));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    CU(ctx, cudaGetLastError());
    ctx->launches += 12;
    out[0] = best_philox;            // Philox draws (32-bit words) per second, generation only
    out[1] = best_dmma / 1e12;       // FP64 tensor TFLOP/s, all 8 x 8 x 4 x 2 flops of every tile counted
    out[2] = 0;
    out[3] = 0;
    return SNPG_OK;
}

// Probe for the TMA tile load used by tile_hist.cu: loads one 192 x 128 u8 box at an arbitrary byte column
// and checks the bytes.  Usage: tma_probe <variant>
//   0 descriptor as __grid_constant__ param + prefetch; 1 same without prefetch; 2 descriptor in global memory;
//   3 box 128 wide (x multiple of 16); 4 box 192 wide but x multiple of 16
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int W>
__global__ void k_probe(const __grid_constant__ CUtensorMap tmap, const CUtensorMap *gmap, int use_global, int prefetch, int x, int y,
                        uint8_t *out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    const CUtensorMap *m = use_global ? gmap : &tmap;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (prefetch) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(W * 128) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(&bar)), "r"(x), "r"(y) : "memory");
    }
    asm volatile(
        "{\n\t.reg .pred P1;\n\tW_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0, 0x989680;\n\t@P1 bra W_DONE;\n\tbra W_LOOP;\n\tW_DONE:\n\t}"
        ::"r"(smem_u32(&bar)) : "memory");
    for (int i = threadIdx.x; i < W * 128; i += blockDim.x) out[i] = smem[i];
}

int main(int argc, char **argv)
{
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int W = variant == 3 ? 128 : 192;
    const int64_t pitch = 29952, rows = 1000, row_bytes = 29903;
    std::vector<uint8_t> h((size_t)pitch * rows);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t *d = nullptr, *out = nullptr;
    cudaMalloc(&d, h.size());
    cudaMalloc(&out, 192 * 128);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { printf("no entry point\n"); return 2; }
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    alignas(64) CUtensorMap map;
    const cuuint64_t gdim[2] = {(cuuint64_t)row_bytes, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)pitch};
    const cuuint32_t box[2] = {(cuuint32_t)W, 128};
    const cuuint32_t es[2] = {1, 1};
    CUresult r = ((EncodeFn)fp)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, gdim, gstride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d encode rc=%d\n", variant, (int)r);
    CUtensorMap *gmap = nullptr;
    cudaMalloc(&gmap, sizeof(CUtensorMap));
    cudaMemcpy(gmap, &map, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    const int x = (variant == 3 || variant == 4) ? 272 : 265, y = 37;
    const int use_global = variant == 2, prefetch = variant == 0;
    if (W == 128) k_probe<128><<<1, 128, 128 * 128>>>(map, gmap, use_global, prefetch, x, y, out);
    else k_probe<192><<<1, 128, 192 * 128>>>(map, gmap, use_global, prefetch, x, y, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("variant %d kernel: %s\n", variant, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<uint8_t> o((size_t)W * 128);
    cudaMemcpy(o.data(), out, o.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int rr = 0; rr < 128; rr++)
        for (int c = 0; c < W; c++) {
            const uint8_t want = (x + c < row_bytes && y + rr < rows) ? h[(size_t)(y + rr) * pitch + x + c] : 0;
            if (o[(size_t)rr * W + c] != want) bad++;
        }
    printf("variant %d mismatches: %d\n", variant, bad);
    return bad != 0;
}

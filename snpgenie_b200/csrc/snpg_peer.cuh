// Multi-GPU exchange over peer memory (NVLink 5 / NVSwitch): device side.
//
// Every rank owns one exchange region in its own HBM.  Its peers map it (cudaDeviceEnablePeerAccess inside one
// process, CUDA IPC handles across processes) and the producing kernels store their results straight into every
// rank's copy -- the per-codon statistics of a codon range (k_stats / k_fused_fixup_stats), the per-replicate values
// of a replicate range (k_bootstrap*) -- so the "all-gather" of SURVEY.md 8(e) is the tail of the kernel that
// computes the data, not a separate collective.  Arrival is signalled per (stage, source rank) with a flag that
// carries the job's epoch: the last block of a producer kernel to finish releases it at system scope into every
// rank's header; the first thread of every consumer block acquires all of them before the block reads the data.
// Flags only ever grow, so nothing is reset between jobs; the data halves alternate with the epoch's parity
// (chosen on the host: a rank can be at most one job ahead of a peer).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "snpg_internal.h"

namespace snpg {

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// End of a producer kernel; EVERY thread of EVERY block calls it once, after its last store into peer memory
// (wrote = this thread stored something there: only those threads pay for the system-wide fence).
__device__ __forceinline__ void peer_signal(const PeerSync &ps, int stage, bool wrote = true) {
    if (ps.world <= 1) return;
    if (wrote) __threadfence_system();           // this thread's peer stores are visible system-wide before the flag can be
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        XHeader *me = ps.hdr[ps.rank];
        const unsigned blocks = gridDim.x * gridDim.y * gridDim.z;
        if (atomicAdd(&me->done[stage], 1u) == blocks - 1) {
            me->done[stage] = 0;                 // for the next launch
            __threadfence_system();
            const uint32_t e = *ps.epoch;
            for (int r = 0; r < ps.world; r++) st_release_sys(&ps.hdr[r]->flag[stage][ps.rank], e);
        }
    }
}

// Start of the consuming part of a kernel; every thread of the block calls it.  Gives up after kPeerTimeoutNs
// (a peer died or the ranks disagree about the job sequence): the error word is what the host reports.
constexpr uint64_t kPeerTimeoutNs = 8000000000ull;
__device__ __forceinline__ void peer_wait(const PeerSync &ps, int stage) {
    if (ps.world <= 1) return;
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        XHeader *me = ps.hdr[ps.rank];
        const uint32_t e = *ps.epoch;
        const uint64_t t0 = global_timer_ns();
        for (int r = 0; r < ps.world; r++) {
            while ((int32_t)(ld_acquire_sys(&me->flag[stage][r]) - e) < 0) {
                __nanosleep(64);
                if (global_timer_ns() - t0 > kPeerTimeoutNs) { me->err = 1u + (uint32_t)stage; break; }
            }
        }
    }
    __syncthreads();
}

}  // namespace snpg

// Several GPUs behind the C ABI.
//
// The path shards by codon range and by bootstrap replicate; what has to travel is small -- per-codon
// statistics (46 bytes a codon) and per-replicate values (48 bytes a replicate and product) -- so the exchange is
// latency, not bandwidth.  Every rank owns an exchange region in its HBM; its peers map it (peer access inside one
// process, CUDA IPC between processes) and the kernels that produce the data store it into every region themselves
// and flag its arrival (snpg_peer.cuh).  This file is the host side: laying the region out, exporting and mapping it,
// and the team of per-GPU contexts behind snpg_create_multi.  The reference's counterpart is the parent process
// collecting its forked children's results (snpgenie_within_group.pl:508-538; snpgenie_between_group.pl:949-1017).
#include <unistd.h>

#include <algorithm>

#include "snpg_ctx.h"
#include "snpg_team.h"

using namespace snpg;

namespace {

struct PeerHandle {                 // SNPG_PEER_HANDLE_BYTES, plain bytes for the host program to carry
    uint32_t magic; int32_t pid; int32_t device; int32_t rank; int32_t world; int32_t P;
    int64_t total, max_B; uint64_t bytes; uint64_t ptr;
    cudaIpcMemHandle_t ipc;         // 64 bytes
    uint32_t pci;                   // domain << 16 | bus << 8 | device: which physical GPU (ordinals differ between processes)
    uint8_t pad[SNPG_PEER_HANDLE_BYTES - 4 * 6 - 8 * 4 - 64 - 4];
};
uint32_t pci_of(int device) {
    int dom = 0, bus = 0, dev = 0;
    cudaDeviceGetAttribute(&dom, cudaDevAttrPciDomainId, device);
    cudaDeviceGetAttribute(&bus, cudaDevAttrPciBusId, device);
    cudaDeviceGetAttribute(&dev, cudaDevAttrPciDeviceId, device);
    return ((uint32_t)dom << 16) | ((uint32_t)bus << 8) | (uint32_t)dev;
}
static_assert(sizeof(PeerHandle) == SNPG_PEER_HANDLE_BYTES, "handle size is part of the ABI");
constexpr uint32_t kHandleMagic = 0x58504E53u;   // "SNPX"

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

void layout(Exchange &X, int64_t total, int P, int64_t max_B) {
    X.total = total; X.P = P; X.max_B = max_B;
    size_t o = 0;
    X.off_stats = o; o = align_up(o + sizeof(double) * 4 * (size_t)total, 256);
    X.off_cmp = o;   o = align_up(o + sizeof(uint64_t) * (size_t)total, 256);
    X.off_ndef = o;  o = align_up(o + sizeof(uint32_t) * (size_t)total, 256);
    X.off_poly = o;  o = align_up(o + (size_t)total, 256);
    X.off_cons = o;  o = align_up(o + (size_t)total, 256);
    X.off_vals = o;  o = align_up(o + sizeof(double) * 6 * (size_t)P * (size_t)std::max<int64_t>(1, max_B), 256);
    X.half = o;
    X.bytes = sizeof(XHeader) + 2 * X.half;
}

}  // namespace

namespace snpg {

PeerSync peer_sync_of(const snpg_ctx *ctx) {
    PeerSync ps;
    if (ctx->world > 1 && ctx->xchg.ready) {
        ps.world = ctx->world;
        ps.rank = ctx->rank;
        for (int r = 0; r < ctx->world; r++) ps.hdr[r] = reinterpret_cast<XHeader *>(ctx->xchg.base[r]);
    }
    return ps;
}

void exchange_release(snpg_ctx *ctx) {
    Exchange &X = ctx->xchg;
    if (!X.region.p && !X.ready) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (int r = 0; r < kMaxPeers; r++) {
        if (X.ipc_opened[r] && X.base[r]) cudaIpcCloseMemHandle(X.base[r]);
        X.ipc_opened[r] = false;
        X.base[r] = nullptr;
    }
    X.region.release();
    X.ready = false;
    X.epoch = 0;
    invalidate_job_graph(ctx);
}

int exchange_check_error(snpg_ctx *ctx) {
    if (!(ctx->world > 1 && ctx->xchg.ready)) return SNPG_OK;
    uint32_t err = 0;
    CU(ctx, cudaMemcpy(&err, ctx->xchg.base[ctx->rank] + offsetof(XHeader, err), sizeof err, cudaMemcpyDeviceToHost));
    if (err) {
        cudaMemset(ctx->xchg.base[ctx->rank] + offsetof(XHeader, err), 0, sizeof err);
        return fail(ctx, SNPG_E_CUDA, "rank %d waited more than 8 s for its peers' %s (a peer failed, or the ranks do not run the same jobs)",
                    ctx->rank, err == 1 ? "per-codon statistics" : "bootstrap replicates");
    }
    return SNPG_OK;
}

int team_fail(snpg_ctx *team, snpg_ctx *child, int rc) {
    if (child) team->err = "GPU " + std::to_string(child->device) + ": " + child->err;
    return rc;
}

int team_wire(snpg_ctx *team, int64_t max_bootstraps) {
    const size_t n = team->team.size();
    std::vector<uint8_t> handles(n * SNPG_PEER_HANDLE_BYTES);
    for (size_t i = 0; i < n; i++)
        if (int rc = snpg_peer_export(team->team[i], max_bootstraps, handles.data() + i * SNPG_PEER_HANDLE_BYTES)) return team_fail(team, team->team[i], rc);
    for (size_t i = 0; i < n; i++)
        if (int rc = snpg_peer_connect(team->team[i], handles.data(), (int)n)) return team_fail(team, team->team[i], rc);
    return SNPG_OK;
}

}  // namespace snpg

extern "C" {

int snpg_peer_export(snpg_ctx *ctx, int64_t max_bootstraps, void *handle) {
    if (!ctx || !handle) return SNPG_E_ARG;
    TEAM_UNSUPPORTED(ctx, "snpg_peer_export");
    NEED(ctx, ctx->n_products > 0, SNPG_E_STATE, "set the shard and the products before exporting the exchange region");
    NEED(ctx, ctx->world > 1 && ctx->world <= kMaxPeers, SNPG_E_STATE, "the exchange is for 2..%d ranks (snpg_set_shard), this ctx has world = %d", kMaxPeers, ctx->world);
    NEED(ctx, max_bootstraps >= 0 && max_bootstraps < ((int64_t)1 << 36), SNPG_E_ARG, "bad max_bootstraps");
    if (int rc = bind(ctx)) return rc;
    exchange_release(ctx);
    Exchange &X = ctx->xchg;
    layout(X, ctx->full_ofs[ctx->n_products], ctx->n_products, max_bootstraps);
    CU(ctx, X.region.ensure(X.bytes));
    CU(ctx, cudaMemset(X.region.p, 0, X.bytes));
    CU(ctx, cudaDeviceSynchronize());
    X.base[ctx->rank] = X.region.as<uint8_t>();
    PeerHandle h;
    memset(&h, 0, sizeof h);
    h.magic = kHandleMagic; h.pid = (int32_t)getpid(); h.device = ctx->device; h.rank = ctx->rank; h.world = ctx->world; h.P = X.P;
    h.total = X.total; h.max_B = X.max_B; h.bytes = X.bytes; h.ptr = (uint64_t)(uintptr_t)X.region.p; h.pci = pci_of(ctx->device);
    CU(ctx, cudaIpcGetMemHandle(&h.ipc, X.region.p));
    memcpy(handle, &h, sizeof h);
    return SNPG_OK;
}

int snpg_peer_connect(snpg_ctx *ctx, const void *handles, int n_handles) {
    if (!ctx || !handles) return SNPG_E_ARG;
    TEAM_UNSUPPORTED(ctx, "snpg_peer_connect");
    Exchange &X = ctx->xchg;
    NEED(ctx, X.region.p && X.base[ctx->rank], SNPG_E_STATE, "call snpg_peer_export first");
    NEED(ctx, n_handles == ctx->world, SNPG_E_ARG, "%d handles for %d ranks", n_handles, ctx->world);
    if (int rc = bind(ctx)) return rc;
    const PeerHandle *H = static_cast<const PeerHandle *>(handles);
    X.peer_on_my_gpu = false;
    for (int r = 0; r < ctx->world; r++) {
        PeerHandle h;
        memcpy(&h, H + r, sizeof h);
        NEED(ctx, h.magic == kHandleMagic && h.rank == r && h.world == ctx->world, SNPG_E_ARG, "handle %d is not rank %d's export", r, r);
        NEED(ctx, h.total == X.total && h.P == X.P && h.max_B == X.max_B && h.bytes == X.bytes, SNPG_E_ARG,
             "rank %d exported a different layout (annotation or max_bootstraps differ)", r);
        if (r == ctx->rank) continue;
        // a peer on this very GPU (tests, a one-GPU box): kernels of the two ranks then compete for the same SMs, and a
        // grid of blocks that all wait for the peer could keep the peer's producer from ever starting -- the job puts its
        // waits into one-block kernels ahead of the consumers instead (job.cu)
        if (h.pci == pci_of(ctx->device)) X.peer_on_my_gpu = true;
        if (h.pid == (int32_t)getpid()) {
            // same process: the pointer itself, once this device may dereference it
            if (h.device != ctx->device) {
                int can = 0;
                CU(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, h.device));
                NEED(ctx, can, SNPG_E_CUDA, "GPU %d cannot access GPU %d's memory (no NVLink/PCIe peer path)", ctx->device, h.device);
                const cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return fail(ctx, SNPG_E_CUDA, "cudaDeviceEnablePeerAccess(%d) failed: %s", h.device, cudaGetErrorString(e));
                cudaGetLastError();
            }
            X.base[r] = reinterpret_cast<uint8_t *>((uintptr_t)h.ptr);
        } else {
            void *p = nullptr;
            const cudaError_t e = cudaIpcOpenMemHandle(&p, h.ipc, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                return fail(ctx, SNPG_E_CUDA, "cudaIpcOpenMemHandle for rank %d failed: %s", r, cudaGetErrorString(e));
            }
            X.base[r] = static_cast<uint8_t *>(p);
            X.ipc_opened[r] = true;
        }
    }
    X.ready = true;
    X.epoch = 0;
    invalidate_job_graph(ctx);
    return SNPG_OK;
}

int snpg_peer_connected(const snpg_ctx *ctx) {
    if (!ctx) return 0;
    if (is_team(ctx)) return ctx->team[0]->xchg.ready ? 1 : 0;
    return ctx->world > 1 && ctx->xchg.ready ? 1 : 0;
}

int snpg_team_size(const snpg_ctx *ctx) { return !ctx ? 0 : is_team(ctx) ? (int)ctx->team.size() : 1; }

int snpg_create_multi(snpg_ctx **out, int n_gpus, const int *devices, uint64_t seed) {
    if (!out) return fail(nullptr, SNPG_E_ARG, "out is NULL");
    *out = nullptr;
    if (n_gpus < 1 || n_gpus > kMaxPeers) return fail(nullptr, SNPG_E_ARG, "n_gpus must be 1..%d", kMaxPeers);
    if (n_gpus == 1) return snpg_create(out, devices ? devices[0] : 0, seed);
    snpg_ctx *team = new snpg_ctx();
    team->seed = seed;
    for (int i = 0; i < n_gpus; i++) {
        snpg_ctx *c = nullptr;
        int rc = snpg_create(&c, devices ? devices[i] : i, seed);
        if (rc == SNPG_OK) rc = snpg_set_shard(c, i, n_gpus);
        if (rc != SNPG_OK) {
            if (c) { fail(nullptr, rc, "%s", c->err.c_str()); snpg_destroy(c); }
            for (snpg_ctx *d : team->team) snpg_destroy(d);
            delete team;
            return rc;
        }
        c->team_parent = team;
        team->team.push_back(c);
    }
    team->device = team->team[0]->device;
    *out = team;
    return SNPG_OK;
}

}  // extern "C"

// The context behind the C ABI (include/snpgenie_cuda.h) and the helpers its translation units share:
// abi.cu (lifecycle, annotation, loading, call-by-call engine), job.cu (whole jobs, CUDA graphs, windows),
// exchange.cu (multi-GPU: peer-memory exchange, teams of GPUs in one process), ingest.cu (FASTA reader).
#pragma once

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include "../../include/snpgenie_cuda.h"
#include "snpg_internal.h"

namespace snpg {

extern std::atomic<uint64_t> g_alloc_epoch;   // bumped whenever a device buffer moves: captured CUDA graphs hold raw pointers

struct DevBuf {  // grow-only device scratch
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        g_alloc_epoch++;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) { cudaFree(p); g_alloc_epoch++; } p = nullptr; cap = 0; }
    template <class T> T *as() const { return static_cast<T *>(p); }
};

struct Group {
    bool loaded = false;
    uint64_t n = 0;
    int64_t n_pad = 0;
    DevBuf codes, hist, other;   // other: [2][count] min then max
    DevBuf sites;                // [aln_len][5] per-site A,C,G,T,other (SNPG_OPT_SITE_COUNTS)
    DevBuf first_def;            // [count] first defined codon in row order (SNPG_OPT_FIRST_DEFINED)
    bool has_first_def = false;
    bool has_sites = false;
    bool has_codes = false;
    bool clear_in_vote = false;  // hist/other are cleared by the load's first kernel, not by memsets
};

// Multi-GPU exchange (exchange.cu): this rank's region and the peers' regions as mapped into this device.
// Region = XHeader, then two halves (jobs alternate between them with the parity of their epoch) of
//   stats4 [total][4] f64 | comparisons [total] u64 | n_defined [total] u32 | poly [total] u8 | conserved [total] u8 |
//   vals [6][P][max_B] f64
struct Exchange {
    bool ready = false;
    int64_t total = 0, max_B = 0; int P = 0;
    DevBuf region;
    size_t bytes = 0, half = 0, off_stats = 0, off_cmp = 0, off_ndef = 0, off_poly = 0, off_cons = 0, off_vals = 0;
    uint8_t *base[kMaxPeers] = {};
    bool ipc_opened[kMaxPeers] = {};
    bool peer_on_my_gpu = false;         // some peer runs on this rank's GPU: waits go into one-block kernels (exchange.cu)
    uint32_t epoch = 0;                  // jobs run so far (all ranks run the same sequence of jobs)
    template <class T> T *at(int r, int parity, size_t off) const { return reinterpret_cast<T *>(base[r] + sizeof(XHeader) + (size_t)parity * half + off); }
};

}  // namespace snpg

struct snpg_ctx {
    int device = 0;
    uint64_t seed = 0;
    std::string err;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_packed[2] = {nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;     // result copies run beside the bootstrap (job.cu)
    cudaStream_t win_stream = nullptr;                    // a whole job's sliding windows run beside the product bootstraps
    cudaEvent_t ev_win = nullptr;
    cudaEvent_t ev_salt_fork = nullptr, ev_salt = nullptr;     // a job's call number is copied beside its first kernels (job.cu)
    // snpg_between_all: the group pairs are independent chains of small kernels; they run on kPairLanes streams side by
    // side, each lane with its own slice of the scratch buffers (created on first use)
    static constexpr int kPairLanes = 6;
    cudaStream_t pair_stream[kPairLanes] = {};
    cudaEvent_t pair_done[kPairLanes] = {};
    int rank = 0, world = 1;
    bool keep_codes = false;        // default route: fused histogram, no code matrix
    bool site_counts = false;       // also count A/C/G/T per alignment column while loading (whole rows are staged)
    bool ingest_async = false;      // snpg_read_fasta returns while its second pass is still copying rows
    bool first_defined = false;     // loads without a code matrix also note every codon's first defined codon (between-group rule)
    bool defer_load_sync = false;   // set inside a whole job: the load's kernels stay queued behind the analysis
    bool load_no_sync = false;      // set inside snpg_between_job: a load returns with its kernels queued (nothing else changes)
    // annotation
    int n_products = 0;
    int64_t aln_len = 0;
    std::vector<int64_t> full_ofs;    // [P+1] offsets on the global codon axis
    std::vector<int64_t> local_ofs;   // [P+1] offsets inside this shard (clamped)
    int64_t shard_first = 0, shard_count = 0;
    int64_t col_lo = 0, col_hi = 0;   // alignment columns this shard touches, [lo, hi)
    snpg::DevBuf d_map, d_local_ofs, d_full_ofs, d_tables;
    // fused path (no code matrix): 16-byte column spans starting at base_al
    int64_t base_al = 0, n_spans = 0;     // base_al = col_lo rounded down to 16
    int64_t n_irr = 0;                    // codons the span runs cannot express
    snpg::DevBuf d_runs, d_regular, d_irr_map, d_irr_idx, d_cons, d_alt;
    snpg::DevBuf t_codes, t_hist, t_other;      // irregular codons: compact K1+K2 scratch
    // TMA-tiled variant of the fused path: frame-aligned tiles of <= 64 codons over the regular codons
    snpg::DevBuf d_tiles, d_sched;              // d_sched: the histogram kernels' chunk counter and finished-CTA counter
    int64_t n_tiles = 0;
    int fused_kernel = SNPG_FUSED_LDG_SPANS;
    int boot_kernel = SNPG_BOOT_CLASSES;
    int boot_slot0 = 1;                   // SNPG_OPT_BOOT_SLOTS: 1 = the reference's C+1 slots (wg:885), 0 = C slots
    const snpg::Group *stats_of = nullptr;      // the group whose per-codon statistics sit in the job's stats buffers (written by its load)
    int n_sms = 148;
    int tma_l2_promotion = 2;             // CU_TENSOR_MAP_L2_PROMOTION_L2_128B (SNPG_TMA_L2PROMO overrides: 0..3)
    // groups
    snpg::Group groups[SNPG_MAX_GROUPS];
    // FASTA slots: pinned host copies [n][len] read by snpg_read_fasta
    struct HostAln { uint8_t *p = nullptr; size_t cap = 0; uint64_t n = 0, len = 0; } host_aln[SNPG_MAX_GROUPS];
    std::vector<snpg::CodonMap> host_map;   // full codon map (absolute columns), for snpg_codon_strings
    // scratch
    snpg::DevBuf s_vals, d_order_local, d_order_full;
    int64_t max_codons_local = 0, max_codons_full = 0;
    snpg::DevBuf s_boot;              // SNPG_BOOT_CLASSES scratch (boot_scratch_bytes)
    snpg::DevBuf staging[2], s_stats, s_poly, s_nda, s_ndb, s_cmp, s_cons, s_sums, s_rep, s_se, s_undef, s_misc, s_in;
    // sliding windows of a whole job: items over the (global) codon axis, cached per (W, step)
    struct WindowPlan { int64_t W = 0, step = 0, n = 0; std::vector<int64_t> ofs; snpg::DevBuf beg, end; } wplan;
    snpg::DevBuf w_sums, w_vals, w_se;
    // where the statistics a job hands to its analysis live (set by job.cu before the load: local scratch on one GPU,
    // every rank's exchange region on several)
    snpg::StatsOut job_out = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    snpg::PeerOut job_peer_out;
    snpg::PeerSync job_peer_sync;
    // a whole job on a device-resident alignment, replayed as a CUDA graph from the third identical call on
    struct TimedSpan { cudaEvent_t e0, e1; int kind; };
    struct PendingCopy { int field; const void *src; size_t bytes; };   // field: which output of the snpg_job (job.cu)
    bool capturing = false;
    struct JobKey {
        uint32_t profile_mask = 0;
        int group = -1; const uint8_t *bases = nullptr; uint64_t n = 0, len = 0, stride = 0;
        bool keep_codes = false, site_counts = false; int fused_kernel = 0, boot_kernel = 0, boot_slot0 = 1, parity = 0;
        snpg_job job = {};
        bool operator==(const JobKey &o) const {
            return profile_mask == o.profile_mask && group == o.group && bases == o.bases && n == o.n && len == o.len && stride == o.stride &&
                   keep_codes == o.keep_codes && site_counts == o.site_counts && fused_kernel == o.fused_kernel && boot_kernel == o.boot_kernel &&
                   boot_slot0 == o.boot_slot0 && parity == o.parity && std::memcmp(&job, &o.job, sizeof job) == 0;
        }
    };
    struct JobGraph {
        cudaGraphExec_t exec = nullptr;
        JobKey key;
        uint64_t epoch = 0;
        int64_t n_launches = 0, n_tile_launches = 0, n_span_tma_launches = 0;
        std::vector<TimedSpan> spans;   // event pairs recorded inside the graph (profiling on)
        std::vector<PendingCopy> pending;   // host copies that follow every replay
    } job_graph[2];                     // one per exchange parity (a single GPU only uses [0])
    JobKey job_candidate; uint64_t job_candidate_epoch = ~0ull; bool job_graphs_disabled = false;
    JobGraph *capturing_into = nullptr;
    int64_t graph_replays = 0;
    uint32_t *h_salt = nullptr;         // pinned: [0] the job's call number, [1] its exchange epoch
    uint8_t *h_out = nullptr; size_t h_out_cap = 0;   // pinned staging of a job's outputs
    std::vector<PendingCopy> pending;   // staging -> caller's (pageable) arrays, after the job's synchronisation
    snpg::DevBuf d_salt;
    int64_t launches = 0;
    int64_t tile_launches = 0;      // launches of k_tile_hist
    int64_t span_tma_launches = 0;  // launches of k_span_hist (the rest of SNPG_K_FUSED went to k_fused_hist)
    uint32_t call_id = 0;
    // optional per-kernel timing (CUDA events on the launch stream)
    uint32_t profile_mask = 0;      // bit k set: bracket launches of kernel kind k with timing events
    bool own_stream = true;
    typedef TimedSpan Span;
    std::vector<Span> spans;        // recorded, not yet resolved
    std::vector<cudaEvent_t> free_events;
    double kernel_ms[SNPG_N_KERNEL_KINDS] = {0};
    int64_t kernel_calls[SNPG_N_KERNEL_KINDS] = {0};
    // several GPUs
    snpg::Exchange xchg;
    std::vector<snpg_ctx *> team;   // snpg_create_multi: this ctx only fans calls out to one child per GPU
    snpg_ctx *team_parent = nullptr;
};

namespace snpg {

int fail(snpg_ctx *ctx, int code, const char *fmt, ...);
cudaEvent_t take_event(snpg_ctx *ctx);

// Brackets one kernel launch with events when profiling is on; counts it always.
struct Launch {
    snpg_ctx *ctx; int kind; cudaEvent_t e0 = nullptr;
    Launch(snpg_ctx *c, int k) : ctx(c), kind(k) {
        ctx->launches++;
        if ((ctx->profile_mask >> kind) & 1u) {
            e0 = take_event(ctx);
            // inside a graph capture the records become external event-record nodes: the same pair of events is
            // re-recorded by every replay and read after it
            if (ctx->capturing) cudaEventRecordWithFlags(e0, ctx->stream, cudaEventRecordExternal);
            else cudaEventRecord(e0, ctx->stream);
        }
    }
    ~Launch() {
        if (!e0) return;
        cudaEvent_t e1 = take_event(ctx);
        if (ctx->capturing && ctx->capturing_into) { cudaEventRecordWithFlags(e1, ctx->stream, cudaEventRecordExternal); ctx->capturing_into->spans.push_back({e0, e1, kind}); }
        else { cudaEventRecord(e1, ctx->stream); ctx->spans.push_back({e0, e1, kind}); }
    }
};

// SNPG_BOOT_CLASSES: the scratch of a bootstrap over items inside stats4[0, total) (call before the Launch scope; a no-op
// for the other kernels).  Grows ctx->s_boot when needed, so whole jobs size it in job_prepare.
int boot_classes_scratch(snpg_ctx *ctx, BootArgs &a, int64_t total);
void resolve_spans(snpg_ctx *ctx);           // after a stream synchronise: recorded spans -> accumulated milliseconds
void invalidate_job_graph(snpg_ctx *ctx);    // a captured job graph holds pointers, options and event pairs: anything that changes them drops it
cudaError_t sync_stream(snpg_ctx *ctx);
int bind(snpg_ctx *ctx);
int check_group(snpg_ctx *ctx, int g);
int check_product(snpg_ctx *ctx, int p);
GroupView view_of(const snpg_ctx *ctx, const Group &g, int64_t c0);
void product_slice(const snpg_ctx *ctx, int product, int64_t *c0, int64_t *C);
PeerSync peer_sync_of(const snpg_ctx *ctx);  // world <= 1 unless the exchange is connected
void exchange_release(snpg_ctx *ctx);        // unmaps the peers, frees the region
int exchange_check_error(snpg_ctx *ctx);     // the region's error word (a peer wait timed out) -> SNPG_E_CUDA

// team fan-out (exchange.cu): every entry point of the C ABI that a multi-GPU ctx supports starts with one of these
bool is_team(const snpg_ctx *ctx);

// FASTA slots are parsed in the background (ingest.cu): whoever reads the buffer waits for the rows it needs
void ingest_wait_rows(const uint8_t *p, uint64_t rows);   // rows [0, rows) counted from the row p points into; no-op for other memory
void ingest_forget(const uint8_t *buffer);                // before a slot's buffer is reused or freed

// SNPG_TRACE=1 in the environment: wall time of the traced entry points on stderr (tools/cli_wall.py reads them)
struct Trace {
    const char *what; double t0;
    static bool on() { static const bool v = [] { const char *e = getenv("SNPG_TRACE"); return e && *e && *e != '0'; }(); return v; }
    static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
    explicit Trace(const char *w) : what(w), t0(on() ? now() : 0) {}
    ~Trace() { if (on()) fprintf(stderr, "[snpg] %-24s %9.3f ms\n", what, now() - t0); }
};

}  // namespace snpg

#define CU(ctx, call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return snpg::fail(ctx, SNPG_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

#define NEED(ctx, cond, code, ...) do { if (!(cond)) return snpg::fail(ctx, code, __VA_ARGS__); } while (0)

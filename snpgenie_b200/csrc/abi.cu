// C-ABI of libsnpgenie_cuda.so: context, device memory, streams, and the
// sequencing of kernels.  See include/snpgenie_cuda.h for the contract and the
// reference blocks each entry point replaces.  There is no CPU fallback here:
// every compute entry point runs the CUDA kernels or fails with SNPG_E_CUDA.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/snpgenie_cuda.h"
#include "snpg_internal.h"

using namespace snpg;

namespace {

std::string g_create_error;

uint64_t g_alloc_epoch = 0;   // bumped whenever a device buffer moves: captured CUDA graphs hold raw pointers

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
    bool has_sites = false;
    bool has_codes = false;
    bool clear_in_vote = false;  // hist/other are cleared by the load's first kernel, not by memsets
};

}  // namespace

struct snpg_ctx {
    int device = 0;
    uint64_t seed = 0;
    std::string err;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_packed[2] = {nullptr, nullptr};
    int rank = 0, world = 1;
    bool keep_codes = false;        // default route: fused histogram, no code matrix
    bool site_counts = false;       // also count A/C/G/T per alignment column while loading (whole rows are staged)
    bool defer_load_sync = false;   // set inside snpg_within_job*: the load's kernels stay queued behind the analysis
    // annotation
    int n_products = 0;
    int64_t aln_len = 0;
    std::vector<int64_t> full_ofs;    // [P+1] offsets on the global codon axis
    std::vector<int64_t> local_ofs;   // [P+1] offsets inside this shard (clamped)
    int64_t shard_first = 0, shard_count = 0;
    int64_t col_lo = 0, col_hi = 0;   // alignment columns this shard touches, [lo, hi)
    DevBuf d_map, d_local_ofs, d_full_ofs, d_tables;
    // fused path (no code matrix): 16-byte column spans starting at base_al
    int64_t base_al = 0, n_spans = 0;     // base_al = col_lo rounded down to 16
    int64_t n_irr = 0;                    // codons the span runs cannot express
    DevBuf d_runs, d_regular, d_irr_map, d_irr_idx, d_cons, d_alt;
    DevBuf t_codes, t_hist, t_other;      // irregular codons: compact K1+K2 scratch
    // TMA-tiled variant of the fused path: frame-aligned tiles of <= 64 codons over the regular codons
    DevBuf d_tiles, d_sched;              // d_sched: the tile kernel's chunk counter and finished-CTA counter
    int64_t n_tiles = 0;
    int fused_kernel = SNPG_FUSED_LDG_SPANS;
    int boot_kernel = SNPG_BOOT_WEIGHTS;
    const Group *stats_of = nullptr;      // the group whose per-codon statistics sit in s_stats (written by its load; see finish_group)
    int n_sms = 148;
    int tma_l2_promotion = 2;             // CU_TENSOR_MAP_L2_PROMOTION_L2_128B (SNPG_TMA_L2PROMO overrides: 0..3)
    // groups
    Group groups[SNPG_MAX_GROUPS];
    // FASTA slots: pinned host copies [n][len] read by snpg_read_fasta
    struct HostAln { uint8_t *p = nullptr; size_t cap = 0; uint64_t n = 0, len = 0; } host_aln[SNPG_MAX_GROUPS];
    std::vector<CodonMap> host_map;   // full codon map (absolute columns), for snpg_codon_strings
    // scratch
    DevBuf s_vals, d_order_local, d_order_full;
    int64_t max_codons_local = 0, max_codons_full = 0;
    DevBuf staging[2], s_stats, s_poly, s_nda, s_ndb, s_cmp, s_cons, s_sums, s_rep, s_se, s_misc, s_in;
    // snpg_within_job on a device-resident alignment, replayed as a CUDA graph from the third identical call on
    struct TimedSpan { cudaEvent_t e0, e1; int kind; };
    bool capturing = false;
    struct JobKey {
        uint32_t profile_mask = 0;
        int group = -1; const uint8_t *bases = nullptr; uint64_t n = 0, len = 0, stride = 0; int64_t B = 0;
        bool keep_codes = false, site_counts = false, want_se = false; int fused_kernel = 0;
        bool operator==(const JobKey &o) const {
            return profile_mask == o.profile_mask && group == o.group && bases == o.bases && n == o.n && len == o.len && stride == o.stride && B == o.B &&
                   keep_codes == o.keep_codes && site_counts == o.site_counts && want_se == o.want_se && fused_kernel == o.fused_kernel;
        }
    };
    struct JobGraph {
        cudaGraphExec_t exec = nullptr;
        JobKey key, candidate;
        uint64_t epoch = 0, candidate_epoch = ~0ull;
        int64_t n_launches = 0, n_tile_launches = 0;
        bool disabled = false;
        std::vector<TimedSpan> spans;   // event pairs recorded inside the graph (profiling on)
        uint32_t *h_salt = nullptr;     // pinned
        double *h_out = nullptr;        // pinned: sums [P][4] then se [P][6]
        int h_out_products = 0;
        int64_t replays = 0;
    } job;
    DevBuf d_salt;
    int64_t launches = 0;
    int64_t tile_launches = 0;      // launches of k_tile_hist (the rest of SNPG_K_FUSED went to k_fused_hist)
    uint32_t call_id = 0;
    // optional per-kernel timing (CUDA events on the launch stream)
    uint32_t profile_mask = 0;      // bit k set: bracket launches of kernel kind k with timing events
    bool own_stream = true;
    typedef TimedSpan Span;
    std::vector<Span> spans;        // recorded, not yet resolved
    std::vector<cudaEvent_t> free_events;
    double kernel_ms[SNPG_N_KERNEL_KINDS] = {0};
    int64_t kernel_calls[SNPG_N_KERNEL_KINDS] = {0};
};

namespace {

int fail(snpg_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return code;
}

cudaEvent_t take_event(snpg_ctx *ctx) {
    if (!ctx->free_events.empty()) { cudaEvent_t e = ctx->free_events.back(); ctx->free_events.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

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
        if (ctx->capturing) { cudaEventRecordWithFlags(e1, ctx->stream, cudaEventRecordExternal); ctx->job.spans.push_back({e0, e1, kind}); }
        else { cudaEventRecord(e1, ctx->stream); ctx->spans.push_back({e0, e1, kind}); }
    }
};

// after a stream synchronise: turn recorded spans into accumulated milliseconds
void resolve_spans(snpg_ctx *ctx) {
    for (auto &sp : ctx->spans) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, sp.e0, sp.e1) == cudaSuccess) { ctx->kernel_ms[sp.kind] += ms; ctx->kernel_calls[sp.kind]++; }
        ctx->free_events.push_back(sp.e0);
        ctx->free_events.push_back(sp.e1);
    }
    ctx->spans.clear();
}

// a captured snpg_within_job graph holds pointers, options and event pairs: anything that changes them drops it
void invalidate_job_graph(snpg_ctx *ctx) {
    if (ctx->job.exec) cudaGraphExecDestroy(ctx->job.exec);
    ctx->job.exec = nullptr;
    for (auto &sp : ctx->job.spans) { ctx->free_events.push_back(sp.e0); ctx->free_events.push_back(sp.e1); }
    ctx->job.spans.clear();
    ctx->job.candidate_epoch = ~0ull;
}

cudaError_t sync_stream(snpg_ctx *ctx) {
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess && !ctx->spans.empty()) resolve_spans(ctx);
    return e;
}

#define CU(ctx, call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return fail(ctx, SNPG_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

#define NEED(ctx, cond, code, ...) do { if (!(cond)) return fail(ctx, code, __VA_ARGS__); } while (0)

int bind(snpg_ctx *ctx) {
    CU(ctx, cudaSetDevice(ctx->device));
    return SNPG_OK;
}

int check_group(snpg_ctx *ctx, int g) {
    NEED(ctx, g >= 0 && g < SNPG_MAX_GROUPS, SNPG_E_ARG, "group %d out of range", g);
    NEED(ctx, ctx->groups[g].loaded, SNPG_E_STATE, "group %d is not loaded", g);
    return SNPG_OK;
}
int check_product(snpg_ctx *ctx, int p) {
    NEED(ctx, ctx->n_products > 0, SNPG_E_STATE, "no products set");
    NEED(ctx, p >= 0 && p < ctx->n_products, SNPG_E_ARG, "product %d out of range", p);
    return SNPG_OK;
}

GroupView view_of(const snpg_ctx *ctx, const Group &g, int64_t c0) {
    GroupView v;
    v.hist = g.hist.as<uint32_t>() + c0 * kHistBins;
    v.other_min = g.other.as<uint32_t>() + c0;
    v.other_max = g.other.as<uint32_t>() + ctx->shard_count + c0;
    v.codes = g.has_codes ? g.codes.as<uint8_t>() + c0 * g.n_pad : nullptr;
    v.n = (int64_t)g.n;
    v.n_pad = g.n_pad;
    return v;
}

int prepare_group(snpg_ctx *ctx, Group &g, uint64_t n_seqs) {
    g.loaded = false;
    ctx->stats_of = nullptr;
    g.n = n_seqs;
    g.n_pad = (int64_t)((n_seqs + kRowPad - 1) / kRowPad) * kRowPad;
    const int64_t cnt = ctx->shard_count;
    if (ctx->keep_codes) CU(ctx, g.codes.ensure((size_t)std::max<int64_t>(1, cnt * g.n_pad)));
    CU(ctx, g.hist.ensure(sizeof(uint32_t) * (size_t)std::max<int64_t>(1, cnt * kHistBins) + 16));
    CU(ctx, g.other.ensure(sizeof(uint32_t) * (size_t)std::max<int64_t>(1, 2 * cnt)));
    // the fused route clears them in its first kernel (launch_vote_refs, first row block) instead of three memsets
    g.clear_in_vote = !ctx->keep_codes && cnt > 0 && ctx->n_spans > 0 && n_seqs > 0;
    if (!g.clear_in_vote) {
        CU(ctx, cudaMemsetAsync(g.hist.p, 0, sizeof(uint32_t) * (size_t)(cnt * kHistBins), ctx->stream));
        CU(ctx, cudaMemsetAsync(g.other.p, 0xFF, sizeof(uint32_t) * (size_t)cnt, ctx->stream));
        CU(ctx, cudaMemsetAsync(g.other.as<uint32_t>() + cnt, 0, sizeof(uint32_t) * (size_t)cnt, ctx->stream));
    }
    g.has_codes = ctx->keep_codes;
    g.has_sites = ctx->site_counts;
    if (ctx->site_counts) {
        CU(ctx, g.sites.ensure(sizeof(uint32_t) * 5 * (size_t)ctx->aln_len));
        CU(ctx, cudaMemsetAsync(g.sites.p, 0, sizeof(uint32_t) * 5 * (size_t)ctx->aln_len, ctx->stream));
    }
    if (!ctx->keep_codes && ctx->n_irr) {
        CU(ctx, ctx->t_codes.ensure((size_t)(ctx->n_irr * g.n_pad)));
        CU(ctx, ctx->t_hist.ensure(sizeof(uint32_t) * (size_t)(ctx->n_irr * kHistBins)));
        CU(ctx, ctx->t_other.ensure(sizeof(uint32_t) * (size_t)(2 * ctx->n_irr)));
        CU(ctx, cudaMemsetAsync(ctx->t_hist.p, 0, sizeof(uint32_t) * (size_t)(ctx->n_irr * kHistBins), ctx->stream));
        CU(ctx, cudaMemsetAsync(ctx->t_other.p, 0xFF, sizeof(uint32_t) * (size_t)ctx->n_irr, ctx->stream));
        CU(ctx, cudaMemsetAsync(ctx->t_other.as<uint32_t>() + ctx->n_irr, 0, sizeof(uint32_t) * (size_t)ctx->n_irr, ctx->stream));
    }
    return SNPG_OK;
}

// One block of rows [r0, r0+rows) whose bytes start at column base_al: d_block points at
// (row r0, column base_al); row_bytes = readable bytes per row from there.
int pack_block(snpg_ctx *ctx, Group &g, const uint8_t *d_block, int64_t pitch, int64_t rows, int64_t r0, int64_t row_bytes) {
    uint32_t *omin = g.other.as<uint32_t>(), *omax = g.other.as<uint32_t>() + ctx->shard_count;
    if (r0 == 0 && (!ctx->keep_codes || ctx->site_counts)) {
        // the rows every row is compared with: per-column plurality and per-span runner-up over the group's first rows
        Launch l(ctx, SNPG_K_OTHER);
        launch_vote_refs(d_block, pitch, (int)std::min<int64_t>(rows, 32), ctx->n_spans, ctx->d_cons.as<uint8_t>(), ctx->d_alt.as<uint32_t>(),
                         g.clear_in_vote ? g.hist.as<uint32_t>() : nullptr, omin, omax, ctx->shard_count, ctx->stream);
        g.clear_in_vote = false;
    }
    if (ctx->site_counts) {
        Launch l(ctx, SNPG_K_SITES);
        launch_site_hist(d_block, pitch, rows, ctx->d_cons.as<uint8_t>(), ctx->aln_len, g.sites.as<uint32_t>(), ctx->stream);
    }
    if (ctx->keep_codes) {
        Launch l(ctx, SNPG_K_PACK);
        launch_pack(d_block + (ctx->col_lo - ctx->base_al), pitch, rows, r0, ctx->d_map.as<CodonMap>(), ctx->shard_count,
                    g.codes.as<uint8_t>(), g.n_pad, omin, omax, ctx->stream);
    } else {
        {
            Launch l(ctx, SNPG_K_FUSED);
            alignas(64) unsigned char tmap[128];
            if (ctx->fused_kernel == SNPG_FUSED_TMA_TILES && ctx->n_tiles > 0 &&
                make_tile_tensor_map(tmap, d_block, pitch, rows, row_bytes, ctx->tma_l2_promotion)) {
                CU(ctx, launch_tile_hist(tmap, rows, ctx->d_cons.as<uint8_t>(), ctx->d_tiles.as<CodonTile>(), ctx->n_tiles, ctx->n_sms,
                                         ctx->d_sched.as<uint32_t>(), g.hist.as<uint32_t>(), omin, omax, ctx->stream));
                ctx->tile_launches++;
            } else {
                launch_fused_hist(d_block, pitch, rows, ctx->d_cons.as<uint8_t>(), ctx->d_alt.as<uint32_t>(), ctx->n_spans, ctx->d_runs.as<SpanRun>(),
                                  g.hist.as<uint32_t>(), omin, omax, ctx->d_sched.as<uint32_t>(), ctx->n_sms, ctx->stream);
            }
        }
        if (ctx->n_irr) {
            Launch l(ctx, SNPG_K_PACK);
            launch_pack(d_block, pitch, rows, r0, ctx->d_irr_map.as<CodonMap>(), ctx->n_irr, ctx->t_codes.as<uint8_t>(), g.n_pad,
                        ctx->t_other.as<uint32_t>(), ctx->t_other.as<uint32_t>() + ctx->n_irr, ctx->stream);
        }
    }
    CU(ctx, cudaGetLastError());
    return SNPG_OK;
}

int finish_group(snpg_ctx *ctx, Group &g) {
    uint32_t *omin = g.other.as<uint32_t>(), *omax = g.other.as<uint32_t>() + ctx->shard_count;
    if (ctx->site_counts) {
        Launch l(ctx, SNPG_K_SITES);
        launch_site_fixup(ctx->d_cons.as<uint8_t>(), ctx->aln_len, (uint32_t)g.n, g.sites.as<uint32_t>(), ctx->stream);
    }
    if (ctx->keep_codes) {
        Launch l(ctx, SNPG_K_HIST);
        launch_hist(g.codes.as<uint8_t>(), g.n_pad, ctx->shard_count, g.hist.as<uint32_t>(), ctx->stream);
    } else {
        // a whole job (snpg_within_job: the analysis follows the load at once) also takes the per-codon statistics here
        const bool with_stats = ctx->defer_load_sync && ctx->n_irr == 0 && ctx->shard_count > 0 && !ctx->site_counts &&
                                ctx->s_stats.ensure(sizeof(double) * 4 * (size_t)ctx->shard_count) == cudaSuccess;
        ctx->stats_of = nullptr;
        if (with_stats) {
            Launch l(ctx, SNPG_K_OTHER);
            StatsOut o{nullptr, nullptr, nullptr, nullptr, ctx->s_stats.as<double>(), nullptr};
            launch_fused_fixup_stats(ctx->d_cons.as<uint8_t>(), ctx->col_lo - ctx->base_al, ctx->d_alt.as<uint32_t>(), ctx->d_map.as<CodonMap>(),
                                     ctx->shard_count, (uint32_t)g.n, g.hist.as<uint32_t>(), omin, omax, view_of(ctx, g, 0),
                                     ctx->d_tables.as<double>(), o, ctx->stream);
            ctx->stats_of = &g;
        } else {
            Launch l(ctx, SNPG_K_OTHER);
            launch_fused_fixup(ctx->d_cons.as<uint8_t>(), ctx->col_lo - ctx->base_al, ctx->d_alt.as<uint32_t>(), ctx->d_map.as<CodonMap>(), ctx->d_regular.as<uint8_t>(),
                               ctx->shard_count, (uint32_t)g.n, g.hist.as<uint32_t>(), omin, omax, ctx->stream);
        }
        if (ctx->n_irr) {
            { Launch l(ctx, SNPG_K_HIST); launch_hist(ctx->t_codes.as<uint8_t>(), g.n_pad, ctx->n_irr, ctx->t_hist.as<uint32_t>(), ctx->stream); }
            Launch l(ctx, SNPG_K_OTHER);
            launch_scatter_irregular(ctx->t_hist.as<uint32_t>(), ctx->t_other.as<uint32_t>(), ctx->t_other.as<uint32_t>() + ctx->n_irr,
                                     ctx->d_irr_idx.as<int32_t>(), ctx->n_irr, g.hist.as<uint32_t>(), omin, omax, ctx->stream);
        }
    }
    CU(ctx, cudaGetLastError());
    if (!ctx->defer_load_sync) CU(ctx, sync_stream(ctx));
    g.loaded = true;
    return SNPG_OK;
}

}  // namespace

extern "C" {

const char *snpg_version(void) { return "snpgenie-b200 0.1 (sm_100a)"; }

const char *snpg_last_error(const snpg_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int snpg_get_tables(double *sites, double *diffs) {
    if (!sites || !diffs) return SNPG_E_ARG;
    build_tables(sites, diffs);
    return SNPG_OK;
}

int snpg_create(snpg_ctx **out, int device, uint64_t seed) {
    if (!out) return fail(nullptr, SNPG_E_ARG, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, SNPG_E_CUDA, "no CUDA device available (%s); libsnpgenie_cuda has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(nullptr, SNPG_E_ARG, "device %d out of range (have %d)", device, count);
    snpg_ctx *ctx = new snpg_ctx();
    ctx->device = device;
    ctx->seed = seed;
    auto bail = [&](const char *what, cudaError_t ce) {
        fail(nullptr, SNPG_E_CUDA, "%s failed: %s", what, cudaGetErrorString(ce));
        delete ctx;
        return SNPG_E_CUDA;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    for (int i = 0; i < 2; i++) {
        if ((e = cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
        if ((e = cudaEventCreateWithFlags(&ctx->ev_packed[i], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    }
    if (cudaDeviceGetAttribute(&ctx->n_sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || ctx->n_sms <= 0) ctx->n_sms = 148;
    if (const char *v = getenv("SNPG_TMA_L2PROMO")) ctx->tma_l2_promotion = atoi(v);
    if ((e = ctx->d_sched.ensure(2 * sizeof(uint32_t))) != cudaSuccess) return bail("cudaMalloc(sched)", e);
    if ((e = cudaMemset(ctx->d_sched.p, 0, 2 * sizeof(uint32_t))) != cudaSuccess) return bail("cudaMemset(sched)", e);
    std::vector<double> tab(128 + 8192);
    build_tables(tab.data(), tab.data() + 128);
    if ((e = ctx->d_tables.ensure(sizeof(double) * tab.size())) != cudaSuccess) return bail("cudaMalloc(tables)", e);
    if ((e = cudaMemcpy(ctx->d_tables.p, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice)) != cudaSuccess)
        return bail("cudaMemcpy(tables)", e);
    *out = ctx;
    return SNPG_OK;
}

void snpg_destroy(snpg_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    for (auto &g : ctx->groups) { g.codes.release(); g.hist.release(); g.other.release(); g.sites.release(); }
    for (auto &h : ctx->host_aln) if (h.p) cudaFreeHost(h.p);
    invalidate_job_graph(ctx);
    if (ctx->job.h_salt) cudaFreeHost(ctx->job.h_salt);
    if (ctx->job.h_out) cudaFreeHost(ctx->job.h_out);
    DevBuf *bufs[] = {&ctx->d_map, &ctx->d_local_ofs, &ctx->d_full_ofs, &ctx->d_tables, &ctx->staging[0], &ctx->staging[1],
                      &ctx->d_runs, &ctx->d_tiles, &ctx->d_sched, &ctx->d_salt, &ctx->d_regular, &ctx->d_irr_map, &ctx->d_irr_idx, &ctx->d_cons, &ctx->d_alt, &ctx->t_codes, &ctx->t_hist,
                      &ctx->t_other,
                      &ctx->s_stats, &ctx->s_poly, &ctx->s_nda, &ctx->s_ndb, &ctx->s_cmp, &ctx->s_cons, &ctx->s_sums,
                      &ctx->s_rep, &ctx->s_se, &ctx->s_misc, &ctx->s_in, &ctx->s_vals, &ctx->d_order_local, &ctx->d_order_full};
    for (DevBuf *b : bufs) b->release();
    for (int i = 0; i < 2; i++) {
        if (ctx->ev_copied[i]) cudaEventDestroy(ctx->ev_copied[i]);
        if (ctx->ev_packed[i]) cudaEventDestroy(ctx->ev_packed[i]);
    }
    resolve_spans(ctx);
    for (cudaEvent_t e : ctx->free_events) cudaEventDestroy(e);
    if (ctx->stream && ctx->own_stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
}

int snpg_set_option(snpg_ctx *ctx, int option, int64_t value) {
    if (!ctx) return SNPG_E_ARG;
    invalidate_job_graph(ctx);
    if (option == SNPG_OPT_KEEP_CODES) { ctx->keep_codes = value != 0; return SNPG_OK; }
    if (option == SNPG_OPT_SITE_COUNTS) {
        NEED(ctx, ctx->n_products == 0, SNPG_E_STATE, "SNPG_OPT_SITE_COUNTS must be set before snpg_set_products");
        ctx->site_counts = value != 0;
        return SNPG_OK;
    }
    if (option == SNPG_OPT_FUSED_KERNEL) {
        NEED(ctx, value == SNPG_FUSED_TMA_TILES || value == SNPG_FUSED_LDG_SPANS, SNPG_E_ARG, "unknown fused kernel %lld", (long long)value);
        ctx->fused_kernel = (int)value;
        return SNPG_OK;
    }
    if (option == SNPG_OPT_BOOTSTRAP_KERNEL) {
        NEED(ctx, value == SNPG_BOOT_GATHER || value == SNPG_BOOT_WEIGHTS, SNPG_E_ARG, "unknown bootstrap kernel %lld", (long long)value);
        ctx->boot_kernel = (int)value;
        return SNPG_OK;
    }
    if (option == SNPG_OPT_PROFILE) { ctx->profile_mask = value < 0 ? 0xFFFFFFFFu : (uint32_t)value; return SNPG_OK; }
    return fail(ctx, SNPG_E_ARG, "unknown option %d", option);
}

int snpg_shard_plan(int64_t total, int rank, int world, int64_t *first, int64_t *count) {
    if (total < 0 || world < 1 || rank < 0 || rank >= world || !first || !count) return SNPG_E_ARG;
    *first = total * rank / world;
    *count = total * (rank + 1) / world - *first;
    return SNPG_OK;
}

int snpg_set_shard(snpg_ctx *ctx, int rank, int world) {
    if (!ctx) return SNPG_E_ARG;
    NEED(ctx, world >= 1 && rank >= 0 && rank < world, SNPG_E_ARG, "bad shard %d/%d", rank, world);
    NEED(ctx, ctx->n_products == 0, SNPG_E_STATE, "set_shard must precede set_products");
    ctx->rank = rank;
    ctx->world = world;
    return SNPG_OK;
}

int snpg_set_products(snpg_ctx *ctx, int n_products, const int64_t *seg_ofs, const int64_t *seg_start,
                      const int64_t *seg_stop, const int8_t *strand, int64_t aln_len) {
    if (!ctx) return SNPG_E_ARG;
    NEED(ctx, n_products > 0 && seg_ofs && seg_start && seg_stop && strand, SNPG_E_ARG, "bad product arguments");
    NEED(ctx, aln_len > 0 && aln_len < (int64_t)1 << 31, SNPG_E_ARG, "alignment length %lld unsupported", (long long)aln_len);
    if (int rc = bind(ctx)) return rc;
    for (auto &g : ctx->groups) g.loaded = false;  // maps change: loaded groups are stale
    invalidate_job_graph(ctx);

    std::vector<CodonMap> map_all;
    std::vector<int64_t> full_ofs(n_products + 1, 0);
    for (int p = 0; p < n_products; p++) {
        const int64_t s0 = seg_ofs[p], s1 = seg_ofs[p + 1];
        NEED(ctx, s1 > s0, SNPG_E_INPUT, "product %d has no CDS segment", p);
        std::vector<int> order((size_t)(s1 - s0));
        for (size_t i = 0; i < order.size(); i++) order[i] = (int)(s0 + i);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return seg_start[a] < seg_start[b]; });  // wg:1608
        std::vector<int32_t> pos;
        for (int s : order) {
            NEED(ctx, seg_start[s] >= 1 && seg_stop[s] >= seg_start[s] && seg_stop[s] <= aln_len, SNPG_E_INPUT,
                 "product %d: segment %lld..%lld outside the alignment (length %lld)", p, (long long)seg_start[s],
                 (long long)seg_stop[s], (long long)aln_len);
            for (int64_t x = seg_start[s]; x <= seg_stop[s]; x++) pos.push_back((int32_t)(x - 1));
        }
        NEED(ctx, pos.size() % 3 == 0, SNPG_E_INPUT,
             "product %d: the CDS coordinates do not yield a set of complete codons (%zu nucleotides)", p, pos.size());  // wg:1596
        const bool minus = strand[p] < 0;
        if (minus) std::reverse(pos.begin(), pos.end());
        for (size_t k = 0; k < pos.size(); k += 3) map_all.push_back(CodonMap{pos[k], pos[k + 1], pos[k + 2], minus ? 1 : 0});
        full_ofs[p + 1] = (int64_t)map_all.size();
    }
    const int64_t total = (int64_t)map_all.size();
    ctx->host_map = map_all;
    snpg_shard_plan(total, ctx->rank, ctx->world, &ctx->shard_first, &ctx->shard_count);
    ctx->n_products = n_products;
    ctx->aln_len = aln_len;
    ctx->full_ofs = full_ofs;
    ctx->local_ofs.assign(n_products + 1, 0);
    for (int p = 0; p <= n_products; p++)
        ctx->local_ofs[p] = std::min(std::max(full_ofs[p] - ctx->shard_first, (int64_t)0), ctx->shard_count);

    // columns this shard reads; positions become relative to col_lo
    int64_t lo = aln_len, hi = 0;
    for (int64_t c = ctx->shard_first; c < ctx->shard_first + ctx->shard_count; c++) {
        const CodonMap &m = map_all[(size_t)c];
        lo = std::min<int64_t>(lo, std::min(m.p0, std::min(m.p1, m.p2)));
        hi = std::max<int64_t>(hi, std::max(m.p0, std::max(m.p1, m.p2)) + 1);
    }
    if (ctx->shard_count == 0) { lo = 0; hi = 0; }
    if (ctx->site_counts) { lo = 0; hi = aln_len; }     // every column is read: stage whole rows
    ctx->col_lo = lo;
    ctx->col_hi = hi;
    std::vector<CodonMap> local(map_all.begin() + ctx->shard_first, map_all.begin() + ctx->shard_first + ctx->shard_count);
    for (auto &m : local) { m.p0 -= (int32_t)lo; m.p1 -= (int32_t)lo; m.p2 -= (int32_t)lo; }

    // Fused-path description: which codons start in which 16-byte span.
    {
        const int64_t base = lo & ~(int64_t)15;
        const int64_t n_spans = (hi - base + 15) / 16;
        ctx->base_al = base;
        ctx->n_spans = n_spans;
        std::vector<SpanRun> runs((size_t)std::max<int64_t>(1, n_spans) * kSpanRuns, SpanRun{0, 0, 0, 0, 0});
        std::vector<int> used((size_t)std::max<int64_t>(1, n_spans), 0);
        std::vector<uint8_t> regular(std::max<size_t>(1, local.size()), 0);
        std::vector<CodonMap> irr_map;
        std::vector<int32_t> irr_idx;
        const int64_t shift = lo - base;            // map positions are relative to lo
        for (size_t c = 0; c < local.size(); c++) {
            const CodonMap &m = local[c];
            // contiguous bytes q, q+1, q+2: ascending for '+', descending for '-'
            const bool contiguous = m.minus ? (m.p1 == m.p0 - 1 && m.p2 == m.p0 - 2) : (m.p1 == m.p0 + 1 && m.p2 == m.p0 + 2);
            bool placed = false;
            if (contiguous) {
                const int64_t q = (m.minus ? m.p2 : m.p0) + shift;   // lowest byte, relative to base
                const int64_t sp = q / 16;
                const int off = (int)(q % 16);
                const int dir = m.minus ? -1 : 1;
                SpanRun *r = &runs[(size_t)sp * kSpanRuns];
                // extend a run of the same product if this codon is its next element
                for (int l = 0; l < used[(size_t)sp] && !placed; l++) {
                    if (r[l].minus == m.minus && r[l].dir == dir && r[l].off + 3 * r[l].cnt == off &&
                        (int64_t)r[l].codon0 + (int64_t)dir * r[l].cnt == (int64_t)c) { r[l].cnt++; placed = true; }
                    // '-' strand codons arrive in descending position order: prepend
                    else if (r[l].minus == m.minus && r[l].dir == dir && r[l].off - 3 == off && (int64_t)r[l].codon0 - dir == (int64_t)c) {
                        r[l].off = (int8_t)off; r[l].codon0 = (int32_t)c; r[l].cnt++; placed = true;
                    }
                }
                if (!placed && used[(size_t)sp] < kSpanRuns) {
                    r[used[(size_t)sp]++] = SpanRun{(int32_t)c, (int8_t)off, 1, (int8_t)dir, (int8_t)m.minus};
                    placed = true;
                }
            }
            if (placed) regular[c] = 1;
            else {
                CodonMap mm = m;
                mm.p0 += (int32_t)shift; mm.p1 += (int32_t)shift; mm.p2 += (int32_t)shift;   // relative to base
                irr_map.push_back(mm);
                irr_idx.push_back((int32_t)c);
            }
        }
        ctx->n_irr = (int64_t)irr_map.size();
        // TMA tiles: maximal runs of regular codons whose bytes advance by exactly one codon per codon index
        // ('+': ascending columns, '-': descending), cut into pieces of <= 64 codons in ascending-address order.
        {
            std::vector<CodonTile> tiles;
            auto emit = [&](int64_t c_first, int64_t len, int64_t q_first, int minus) {
                // '+': slot j of the run is codon c_first + j at byte q_first + 3j;
                // '-': the lowest address holds the LAST codon of the run and indices fall as addresses rise
                const int64_t q_low = minus ? q_first - 3 * (len - 1) : q_first;
                const int64_t c_low = minus ? c_first + len - 1 : c_first;
                const int dir = minus ? -1 : 1;
                for (int64_t j = 0; j < len; j += 64) {
                    CodonTile t;
                    t.x0 = (int32_t)(q_low + 3 * j);
                    t.codon0 = (int32_t)(c_low + dir * j);
                    t.ncodons = (int16_t)std::min<int64_t>(64, len - j);
                    t.dir = (int8_t)dir;
                    t.minus = (int8_t)minus;
                    t.pad = 0;
                    tiles.push_back(t);
                }
            };
            int64_t run_c = -1, run_len = 0, run_q = 0, run_last_q = 0;
            int run_minus = 0;
            for (size_t c = 0; c < local.size(); c++) {
                if (!regular[c]) { if (run_len) emit(run_c, run_len, run_q, run_minus); run_len = 0; continue; }
                const CodonMap &m = local[c];
                const int64_t q = (m.minus ? m.p2 : m.p0) + shift;
                if (run_len && run_minus == m.minus && q == run_last_q + (m.minus ? -3 : 3)) { run_len++; run_last_q = q; continue; }
                if (run_len) emit(run_c, run_len, run_q, run_minus);
                run_c = (int64_t)c; run_len = 1; run_q = q; run_last_q = q; run_minus = m.minus;
            }
            if (run_len) emit(run_c, run_len, run_q, run_minus);
            ctx->n_tiles = (int64_t)tiles.size();
            if (!tiles.empty()) {
                CU(ctx, ctx->d_tiles.ensure(sizeof(CodonTile) * tiles.size()));
                CU(ctx, cudaMemcpy(ctx->d_tiles.p, tiles.data(), sizeof(CodonTile) * tiles.size(), cudaMemcpyHostToDevice));
            }
        }
        CU(ctx, ctx->d_runs.ensure(sizeof(SpanRun) * runs.size()));
        CU(ctx, ctx->d_regular.ensure(regular.size()));
        CU(ctx, ctx->d_cons.ensure((size_t)(n_spans * 16 + 512)));     // the tile kernel copies 208-byte windows of it
        CU(ctx, ctx->d_alt.ensure(sizeof(uint32_t) * kAltWords * (size_t)std::max<int64_t>(1, n_spans)));
        CU(ctx, cudaMemcpy(ctx->d_runs.p, runs.data(), sizeof(SpanRun) * runs.size(), cudaMemcpyHostToDevice));
        CU(ctx, cudaMemcpy(ctx->d_regular.p, regular.data(), regular.size(), cudaMemcpyHostToDevice));
        CU(ctx, cudaMemset(ctx->d_cons.p, 0, (size_t)(n_spans * 16 + 512)));
        if (ctx->n_irr) {
            CU(ctx, ctx->d_irr_map.ensure(sizeof(CodonMap) * irr_map.size()));
            CU(ctx, ctx->d_irr_idx.ensure(sizeof(int32_t) * irr_idx.size()));
            CU(ctx, cudaMemcpy(ctx->d_irr_map.p, irr_map.data(), sizeof(CodonMap) * irr_map.size(), cudaMemcpyHostToDevice));
            CU(ctx, cudaMemcpy(ctx->d_irr_idx.p, irr_idx.data(), sizeof(int32_t) * irr_idx.size(), cudaMemcpyHostToDevice));
        }
    }

    // bootstrap schedules the largest products first
    {
        std::vector<int> ord_l(n_products), ord_f(n_products);
        for (int p = 0; p < n_products; p++) ord_l[p] = ord_f[p] = p;
        auto len_l = [&](int p) { return ctx->local_ofs[p + 1] - ctx->local_ofs[p]; };
        auto len_f = [&](int p) { return full_ofs[p + 1] - full_ofs[p]; };
        std::stable_sort(ord_l.begin(), ord_l.end(), [&](int a, int b) { return len_l(a) > len_l(b); });
        std::stable_sort(ord_f.begin(), ord_f.end(), [&](int a, int b) { return len_f(a) > len_f(b); });
        ctx->max_codons_local = len_l(ord_l[0]);
        ctx->max_codons_full = len_f(ord_f[0]);
        CU(ctx, ctx->d_order_local.ensure(sizeof(int) * n_products));
        CU(ctx, ctx->d_order_full.ensure(sizeof(int) * n_products));
        CU(ctx, cudaMemcpy(ctx->d_order_local.p, ord_l.data(), sizeof(int) * n_products, cudaMemcpyHostToDevice));
        CU(ctx, cudaMemcpy(ctx->d_order_full.p, ord_f.data(), sizeof(int) * n_products, cudaMemcpyHostToDevice));
    }

    CU(ctx, ctx->d_map.ensure(sizeof(CodonMap) * std::max<size_t>(1, local.size())));
    CU(ctx, ctx->d_local_ofs.ensure(sizeof(int64_t) * (n_products + 1)));
    CU(ctx, ctx->d_full_ofs.ensure(sizeof(int64_t) * (n_products + 1)));
    if (!local.empty()) CU(ctx, cudaMemcpy(ctx->d_map.p, local.data(), sizeof(CodonMap) * local.size(), cudaMemcpyHostToDevice));
    CU(ctx, cudaMemcpy(ctx->d_local_ofs.p, ctx->local_ofs.data(), sizeof(int64_t) * (n_products + 1), cudaMemcpyHostToDevice));
    CU(ctx, cudaMemcpy(ctx->d_full_ofs.p, full_ofs.data(), sizeof(int64_t) * (n_products + 1), cudaMemcpyHostToDevice));
    return SNPG_OK;
}

int snpg_n_products(const snpg_ctx *ctx) { return ctx ? ctx->n_products : SNPG_E_ARG; }

int snpg_n_codons(const snpg_ctx *ctx, int product, int64_t *out) {
    if (!ctx || !out || product < 0 || product >= ctx->n_products) return SNPG_E_ARG;
    *out = ctx->full_ofs[product + 1] - ctx->full_ofs[product];
    return SNPG_OK;
}

int snpg_shard_range(const snpg_ctx *ctx, int64_t *first, int64_t *count, int64_t *total) {
    if (!ctx || ctx->n_products == 0) return SNPG_E_STATE;
    if (first) *first = ctx->shard_first;
    if (count) *count = ctx->shard_count;
    if (total) *total = ctx->full_ofs[ctx->n_products];
    return SNPG_OK;
}

// Local slice [c0, c0+C) of a product inside this shard.
static void product_slice(const snpg_ctx *ctx, int product, int64_t *c0, int64_t *C) {
    *c0 = ctx->local_ofs[product];
    *C = ctx->local_ofs[product + 1] - ctx->local_ofs[product];
}

int snpg_load_group(snpg_ctx *ctx, int group, const uint8_t *bases, uint64_t n_seqs, uint64_t aln_len, uint64_t row_stride) {
    if (!ctx) return SNPG_E_ARG;
    NEED(ctx, ctx->n_products > 0, SNPG_E_STATE, "set products before loading a group");
    NEED(ctx, group >= 0 && group < SNPG_MAX_GROUPS, SNPG_E_ARG, "group %d out of range", group);
    NEED(ctx, bases && n_seqs > 0 && n_seqs < ((uint64_t)1 << 31), SNPG_E_ARG, "bad alignment arguments");
    NEED(ctx, (int64_t)aln_len == ctx->aln_len, SNPG_E_INPUT,
         "alignment length %llu differs from the annotation's %lld", (unsigned long long)aln_len, (long long)ctx->aln_len);
    NEED(ctx, row_stride >= aln_len, SNPG_E_ARG, "row_stride < aln_len");
    if (int rc = bind(ctx)) return rc;
    Group &g = ctx->groups[group];
    if (int rc = prepare_group(ctx, g, n_seqs)) return rc;

    // Row blocks: H2D of block i+1 overlaps packing of block i (two staging buffers).
    // Staged columns start at base_al (16-byte aligned); the pitch leaves room for the
    // fused kernel's 16-byte vectors and look-ahead past the last CDS column.
    const int64_t width = ctx->col_hi - ctx->base_al;
    if (ctx->shard_count > 0 && width > 0) {
        const int64_t pitch = (ctx->n_spans * 16 + 16 + 127) / 128 * 128;
        int64_t rows_per = std::max<int64_t>(kRowPad, ((int64_t)48 << 20) / pitch / kRowPad * kRowPad);
        rows_per = std::min<int64_t>(rows_per, (int64_t)((n_seqs + kRowPad - 1) / kRowPad * kRowPad));
        for (int i = 0; i < 2; i++) {
            const size_t before = ctx->staging[i].cap;
            CU(ctx, ctx->staging[i].ensure((size_t)(rows_per * pitch)));
            if (ctx->staging[i].cap != before) CU(ctx, cudaMemset(ctx->staging[i].p, 0, ctx->staging[i].cap));  // pad bytes defined
        }
        int64_t blk = 0;
        for (int64_t r0 = 0; r0 < (int64_t)n_seqs; r0 += rows_per, blk++) {
            const int b = (int)(blk & 1);
            const int64_t rows = std::min<int64_t>(rows_per, (int64_t)n_seqs - r0);
            if (blk >= 2) CU(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_packed[b], 0));
            CU(ctx, cudaMemcpy2DAsync(ctx->staging[b].p, (size_t)pitch, bases + (size_t)r0 * row_stride + ctx->base_al,
                                      (size_t)row_stride, (size_t)width, (size_t)rows, cudaMemcpyHostToDevice, ctx->copy_stream));
            CU(ctx, cudaEventRecord(ctx->ev_copied[b], ctx->copy_stream));
            CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[b], 0));
            if (int rc = pack_block(ctx, g, ctx->staging[b].as<uint8_t>(), pitch, rows, r0, pitch)) return rc;
            CU(ctx, cudaEventRecord(ctx->ev_packed[b], ctx->stream));
        }
    }
    return finish_group(ctx, g);
}

int snpg_load_group_device(snpg_ctx *ctx, int group, const uint8_t *d_bases, uint64_t n_seqs, uint64_t aln_len,
                           uint64_t row_stride) {
    if (!ctx) return SNPG_E_ARG;
    NEED(ctx, ctx->n_products > 0, SNPG_E_STATE, "set products before loading a group");
    NEED(ctx, group >= 0 && group < SNPG_MAX_GROUPS, SNPG_E_ARG, "group %d out of range", group);
    NEED(ctx, d_bases && n_seqs > 0 && n_seqs < ((uint64_t)1 << 31), SNPG_E_ARG, "bad alignment arguments");
    NEED(ctx, (int64_t)aln_len == ctx->aln_len, SNPG_E_INPUT,
         "alignment length %llu differs from the annotation's %lld", (unsigned long long)aln_len, (long long)ctx->aln_len);
    NEED(ctx, row_stride >= aln_len, SNPG_E_ARG, "row_stride < aln_len");
    if (int rc = bind(ctx)) return rc;
    Group &g = ctx->groups[group];
    if (int rc = prepare_group(ctx, g, n_seqs)) return rc;
    // The fused path reads whole 16-byte vectors: it needs an aligned base and a row
    // pitch that covers the last span; anything else keeps the code matrix path.
    const bool fused_ok = ((uintptr_t)d_bases % 16 == 0) && (row_stride % 16 == 0) &&
                          ((int64_t)row_stride >= ctx->base_al + ctx->n_spans * 16 + 16);
    NEED(ctx, !ctx->site_counts || fused_ok, SNPG_E_ARG,
         "site counts from a device buffer need a 16-byte aligned base and a row pitch that is a multiple of 16 and covers the last 16-byte span");
    const bool saved_keep = ctx->keep_codes;
    if (!saved_keep && !fused_ok) {
        ctx->keep_codes = true;                                    // fall back to K1+K2 for this load
        if (int rc = prepare_group(ctx, g, n_seqs)) { ctx->keep_codes = saved_keep; return rc; }
    }
    const int64_t rows_per = (int64_t)1 << 20;  // keeps gridDim.y < 65536 for both paths
    int rc = SNPG_OK;
    for (int64_t r0 = 0; r0 < (int64_t)n_seqs && ctx->shard_count > 0 && rc == SNPG_OK; r0 += rows_per) {
        const int64_t rows = std::min<int64_t>(rows_per, (int64_t)n_seqs - r0);
        rc = pack_block(ctx, g, d_bases + (size_t)r0 * row_stride + ctx->base_al, (int64_t)row_stride, rows, r0,
                        (int64_t)aln_len - ctx->base_al);
    }
    if (rc == SNPG_OK) rc = finish_group(ctx, g);
    if (!saved_keep && !fused_ok) { ctx->keep_codes = saved_keep; g.codes.release(); g.has_codes = false; }
    return rc;
}

int snpg_read_fasta(snpg_ctx *ctx, int slot, const char *path, const uint8_t **bases, uint64_t *n_seqs, uint64_t *aln_len) {
    if (!ctx || !path || !bases || !n_seqs || !aln_len) return SNPG_E_ARG;
    NEED(ctx, slot >= 0 && slot < SNPG_MAX_GROUPS, SNPG_E_ARG, "slot %d out of range", slot);
    if (int rc = bind(ctx)) return rc;
    const int fd = open(path, O_RDONLY);
    NEED(ctx, fd >= 0, SNPG_E_INPUT, "Could not open file %s", path);
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size == 0) { close(fd); return fail(ctx, SNPG_E_INPUT, "%s is empty or unreadable", path); }
    const size_t size = (size_t)st.st_size;
    const char *f = static_cast<const char *>(mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0));
    close(fd);
    NEED(ctx, f != MAP_FAILED, SNPG_E_INPUT, "cannot map %s", path);
    madvise(const_cast<char *>(f), size, MADV_SEQUENTIAL);
    struct Unmap { const char *p; size_t n; ~Unmap() { munmap(const_cast<char *>(p), n); } } unmap{f, size};

    // pass 1: count records and measure the first one
    uint64_t n = 0, first_len = 0;
    bool in_first = false;
    for (size_t pos = 0; pos < size;) {
        const char *nl = static_cast<const char *>(memchr(f + pos, '\n', size - pos));
        const size_t end = nl ? (size_t)(nl - f) : size;
        if (memchr(f + pos, '>', end - pos)) { n++; in_first = (n == 1); }
        else if (in_first) first_len += end - pos;
        pos = end + 1;
    }
    NEED(ctx, n > 0 && first_len > 0, SNPG_E_INPUT, "no FASTA records in %s", path);
    auto &h = ctx->host_aln[slot];
    const size_t need = (size_t)n * first_len;
    if (need > h.cap) {
        if (h.p) cudaFreeHost(h.p);
        h.p = nullptr; h.cap = 0;
        CU(ctx, cudaHostAlloc(reinterpret_cast<void **>(&h.p), need, cudaHostAllocDefault));
        h.cap = need;
    }
    // pass 2: copy the data lines of each record into its row
    uint64_t rec = 0, filled = 0;
    bool started = false;
    for (size_t pos = 0; pos < size;) {
        const char *nl = static_cast<const char *>(memchr(f + pos, '\n', size - pos));
        const size_t end = nl ? (size_t)(nl - f) : size;
        if (memchr(f + pos, '>', end - pos)) {
            if (started) {
                NEED(ctx, filled == first_len, SNPG_E_INPUT, "The sequences must be aligned, i.e., must be the same length.");
                rec++;
            }
            started = true;
            filled = 0;
        } else if (started) {
            const size_t len = end - pos;
            NEED(ctx, filled + len <= first_len, SNPG_E_INPUT, "The sequences must be aligned, i.e., must be the same length.");
            memcpy(h.p + rec * first_len + filled, f + pos, len);
            filled += len;
        }
        pos = end + 1;
    }
    NEED(ctx, filled == first_len, SNPG_E_INPUT, "The sequences must be aligned, i.e., must be the same length.");
    h.n = n;
    h.len = first_len;
    *bases = h.p;
    *n_seqs = n;
    *aln_len = first_len;
    return SNPG_OK;
}

int snpg_codon_strings(snpg_ctx *ctx, int slot, int product, int64_t codon, uint8_t *out) {
    if (!ctx || !out) return SNPG_E_ARG;
    NEED(ctx, slot >= 0 && slot < SNPG_MAX_GROUPS && ctx->host_aln[slot].p, SNPG_E_STATE, "no FASTA was read into slot %d", slot);
    if (int rc = check_product(ctx, product)) return rc;
    const int64_t C = ctx->full_ofs[product + 1] - ctx->full_ofs[product];
    NEED(ctx, codon >= 0 && codon < C, SNPG_E_ARG, "codon %lld out of range", (long long)codon);
    const auto &h = ctx->host_aln[slot];
    NEED(ctx, (int64_t)h.len == ctx->aln_len, SNPG_E_INPUT, "slot %d has alignment length %llu, the annotation %lld", slot,
         (unsigned long long)h.len, (long long)ctx->aln_len);
    const CodonMap m = ctx->host_map[(size_t)(ctx->full_ofs[product] + codon)];
    const int32_t pos[3] = {m.p0, m.p1, m.p2};
    for (uint64_t s = 0; s < h.n; s++)
        for (int k = 0; k < 3; k++) {
            uint8_t c = h.p[s * h.len + (size_t)pos[k]];
            if (c >= 'a' && c <= 'z') c = (uint8_t)(c - 32);
            if (m.minus) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c;   // fasta2revcom.pl:79
            if (c == 'U') c = 'T';
            out[3 * s + k] = c;
        }
    return SNPG_OK;
}

int snpg_drop_group(snpg_ctx *ctx, int group) {
    if (!ctx || group < 0 || group >= SNPG_MAX_GROUPS) return SNPG_E_ARG;
    if (int rc = bind(ctx)) return rc;
    Group &g = ctx->groups[group];
    g.codes.release(); g.hist.release(); g.other.release(); g.sites.release();
    g.loaded = false; g.has_codes = false; g.has_sites = false;
    return SNPG_OK;
}

int snpg_group_size(const snpg_ctx *ctx, int group, uint64_t *n_seqs) {
    if (!ctx || group < 0 || group >= SNPG_MAX_GROUPS || !n_seqs || !ctx->groups[group].loaded) return SNPG_E_ARG;
    *n_seqs = ctx->groups[group].n;
    return SNPG_OK;
}

int snpg_get_codon_indices(snpg_ctx *ctx, int group, int product, uint8_t *out) {
    if (!ctx || !out) return SNPG_E_ARG;
    if (int rc = check_group(ctx, group)) return rc;
    if (int rc = check_product(ctx, product)) return rc;
    if (int rc = bind(ctx)) return rc;
    Group &g = ctx->groups[group];
    NEED(ctx, g.has_codes, SNPG_E_STATE, "codon matrix was not kept (SNPG_OPT_KEEP_CODES=0)");
    int64_t c0, C;
    product_slice(ctx, product, &c0, &C);
    if (C == 0) return SNPG_OK;
    const size_t bytes = (size_t)g.n * (size_t)C;
    CU(ctx, ctx->s_misc.ensure(bytes));
    { Launch l(ctx, SNPG_K_OTHER); launch_untranspose(g.codes.as<uint8_t>(), g.n_pad, (int64_t)g.n, c0, C, ctx->s_misc.as<uint8_t>(), ctx->stream); }
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(out, ctx->s_misc.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, sync_stream(ctx));
    return SNPG_OK;
}

int snpg_get_hist(snpg_ctx *ctx, int group, int product, uint32_t *out, uint8_t *other_distinct) {
    if (!ctx || !out) return SNPG_E_ARG;
    if (int rc = check_group(ctx, group)) return rc;
    if (int rc = check_product(ctx, product)) return rc;
    if (int rc = bind(ctx)) return rc;
    Group &g = ctx->groups[group];
    int64_t c0, C;
    product_slice(ctx, product, &c0, &C);
    if (C == 0) return SNPG_OK;
    CU(ctx, cudaMemcpyAsync(out, g.hist.as<uint32_t>() + c0 * kHistBins, sizeof(uint32_t) * (size_t)(C * kHistBins),
                            cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<uint32_t> lo, hi;
    if (other_distinct) {
        lo.resize((size_t)C); hi.resize((size_t)C);
        CU(ctx, cudaMemcpyAsync(lo.data(), g.other.as<uint32_t>() + c0, sizeof(uint32_t) * (size_t)C, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaMemcpyAsync(hi.data(), g.other.as<uint32_t>() + ctx->shard_count + c0, sizeof(uint32_t) * (size_t)C,
                                cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(ctx, sync_stream(ctx));
    if (other_distinct)
        for (int64_t c = 0; c < C; c++) other_distinct[c] = (lo[(size_t)c] != kOtherNone && lo[(size_t)c] != hi[(size_t)c]) ? 1 : 0;
    return SNPG_OK;
}

int snpg_get_site_counts(snpg_ctx *ctx, int group, uint32_t *out) {
    if (!ctx || !out) return SNPG_E_ARG;
    if (int rc = check_group(ctx, group)) return rc;
    if (int rc = bind(ctx)) return rc;
    Group &g = ctx->groups[group];
    NEED(ctx, g.has_sites, SNPG_E_STATE, "site counts were not requested (SNPG_OPT_SITE_COUNTS) when the group was loaded");
    const int64_t L = ctx->aln_len;
    std::vector<uint32_t> tmp((size_t)L * 5);
    CU(ctx, cudaMemcpyAsync(tmp.data(), g.sites.p, sizeof(uint32_t) * tmp.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, sync_stream(ctx));
    for (int64_t c = 0; c < L; c++)
        for (int k = 0; k < 4; k++) out[4 * c + k] = tmp[(size_t)(5 * c + k)];
    return SNPG_OK;
}

// shared by snpg_within / snpg_between: run K3 on [c0, c0+C) and copy what was asked for
static int run_stats(snpg_ctx *ctx, bool between, const Group &ga, const Group &gb, int64_t c0, int64_t C, uint8_t *poly,
                     uint32_t *nda, uint32_t *ndb, uint64_t *cmp, double *stats4, uint8_t *cons) {
    if (C == 0) return SNPG_OK;
    CU(ctx, ctx->s_poly.ensure((size_t)C));
    CU(ctx, ctx->s_cons.ensure((size_t)C));
    CU(ctx, ctx->s_nda.ensure(sizeof(uint32_t) * (size_t)C));
    CU(ctx, ctx->s_ndb.ensure(sizeof(uint32_t) * (size_t)C));
    CU(ctx, ctx->s_cmp.ensure(sizeof(uint64_t) * (size_t)C));
    ctx->stats_of = nullptr;
    CU(ctx, ctx->s_stats.ensure(sizeof(double) * 4 * (size_t)C));
    StatsOut o{ctx->s_poly.as<uint8_t>(), ctx->s_nda.as<uint32_t>(), ctx->s_ndb.as<uint32_t>(),
               ctx->s_cmp.as<unsigned long long>(), ctx->s_stats.as<double>(), ctx->s_cons.as<uint8_t>()};
    {
        Launch l(ctx, SNPG_K_STATS);
        if (between) launch_between_stats(view_of(ctx, ga, c0), view_of(ctx, gb, c0), C, ctx->d_tables.as<double>(), o, ctx->stream);
        else launch_within_stats(view_of(ctx, ga, c0), C, ctx->d_tables.as<double>(), o, ctx->stream);
    }
    CU(ctx, cudaGetLastError());
    if (poly) CU(ctx, cudaMemcpyAsync(poly, o.poly, (size_t)C, cudaMemcpyDeviceToHost, ctx->stream));
    if (cons) CU(ctx, cudaMemcpyAsync(cons, o.conserved, (size_t)C, cudaMemcpyDeviceToHost, ctx->stream));
    if (nda) CU(ctx, cudaMemcpyAsync(nda, o.ndef_a, sizeof(uint32_t) * (size_t)C, cudaMemcpyDeviceToHost, ctx->stream));
    if (ndb) CU(ctx, cudaMemcpyAsync(ndb, o.ndef_b, sizeof(uint32_t) * (size_t)C, cudaMemcpyDeviceToHost, ctx->stream));
    if (cmp) CU(ctx, cudaMemcpyAsync(cmp, o.comparisons, sizeof(uint64_t) * (size_t)C, cudaMemcpyDeviceToHost, ctx->stream));
    if (stats4) CU(ctx, cudaMemcpyAsync(stats4, o.stats4, sizeof(double) * 4 * (size_t)C, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, sync_stream(ctx));
    return SNPG_OK;
}

int snpg_within(snpg_ctx *ctx, int group, int product, uint8_t *poly, uint32_t *n_defined, uint64_t *comparisons,
                double *stats4, uint8_t *conserved_code) {
    if (!ctx) return SNPG_E_ARG;
    if (int rc = check_group(ctx, group)) return rc;
    if (int rc = check_product(ctx, product)) return rc;
    if (int rc = bind(ctx)) return rc;
    int64_t c0, C;
    product_slice(ctx, product, &c0, &C);
    const Group &g = ctx->groups[group];
    return run_stats(ctx, false, g, g, c0, C, poly, n_defined, nullptr, comparisons, stats4, conserved_code);
}

int snpg_between(snpg_ctx *ctx, int group_a, int group_b, int product, uint8_t *poly, uint32_t *n_defined_a,
                 uint32_t *n_defined_b, uint64_t *comparisons, double *stats4, uint8_t *conserved_code) {
    if (!ctx) return SNPG_E_ARG;
    if (int rc = check_group(ctx, group_a)) return rc;
    if (int rc = check_group(ctx, group_b)) return rc;
    if (int rc = check_product(ctx, product)) return rc;
    if (int rc = bind(ctx)) return rc;
    int64_t c0, C;
    product_slice(ctx, product, &c0, &C);
    return run_stats(ctx, true, ctx->groups[group_a], ctx->groups[group_b], c0, C, poly, n_defined_a, n_defined_b, comparisons,
                     stats4, conserved_code);
}

// host stats4 (+ optional keep mask) -> device scratch s_in, masked
static int upload_stats(snpg_ctx *ctx, const double *stats4, const uint8_t *keep, int64_t C) {
    CU(ctx, ctx->s_in.ensure(sizeof(double) * 4 * (size_t)C));
    if (keep) {
        std::vector<double> m((size_t)C * 4);
        for (int64_t c = 0; c < C; c++)
            for (int q = 0; q < 4; q++) m[(size_t)(4 * c + q)] = keep[c] ? stats4[4 * c + q] : 0.0;
        CU(ctx, cudaMemcpy(ctx->s_in.p, m.data(), sizeof(double) * m.size(), cudaMemcpyHostToDevice));
    } else {
        CU(ctx, cudaMemcpy(ctx->s_in.p, stats4, sizeof(double) * 4 * (size_t)C, cudaMemcpyHostToDevice));
    }
    return SNPG_OK;
}

int snpg_bootstrap_product(snpg_ctx *ctx, const double *stats4, const uint8_t *keep, int64_t C, int64_t B, double *sums_out,
                           double *se_out) {
    if (!ctx) return SNPG_E_ARG;
    NEED(ctx, stats4 && C > 0 && C < ((int64_t)1 << 31) - 1 && B >= 2, SNPG_E_ARG, "bad bootstrap arguments (C=%lld, B=%lld)",
         (long long)C, (long long)B);
    if (int rc = bind(ctx)) return rc;
    if (int rc = upload_stats(ctx, stats4, keep, C)) return rc;
    const int64_t ofs[2] = {0, C};
    CU(ctx, ctx->s_misc.ensure(sizeof ofs));
    CU(ctx, cudaMemcpy(ctx->s_misc.p, ofs, sizeof ofs, cudaMemcpyHostToDevice));
    CU(ctx, ctx->s_rep.ensure(sizeof(double) * 4 * (size_t)B));
    CU(ctx, ctx->s_se.ensure(sizeof(double) * 6));
    const uint32_t sid = 0x10000u + ctx->call_id++;
    CU(ctx, ctx->s_vals.ensure(sizeof(double) * 6 * (size_t)B));
    {
        Launch l(ctx, SNPG_K_BOOT);
        launch_bootstrap(ctx->boot_kernel, ctx->s_in.as<double>(), ctx->s_misc.as<int64_t>(), nullptr, 1, C, 0, B, ctx->seed, sid, nullptr,
                         sums_out ? ctx->s_rep.as<double>() : nullptr, ctx->s_vals.as<double>(), nullptr, ctx->stream);
    }
    { Launch l(ctx, SNPG_K_MOMENTS); launch_moments_vals(ctx->s_vals.as<double>(), 1, B, ctx->s_se.as<double>(), ctx->stream); }
    CU(ctx, cudaGetLastError());
    if (sums_out) CU(ctx, cudaMemcpyAsync(sums_out, ctx->s_rep.p, sizeof(double) * 4 * (size_t)B, cudaMemcpyDeviceToHost, ctx->stream));
    if (se_out) CU(ctx, cudaMemcpyAsync(se_out, ctx->s_se.p, sizeof(double) * 6, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, sync_stream(ctx));
    return SNPG_OK;
}

int snpg_windows(snpg_ctx *ctx, const double *stats4, const uint8_t *keep, int64_t C, int64_t W, int64_t step, int64_t B,
                 double *win_sums, double *win_se, int64_t *nwin_out) {
    if (!ctx) return SNPG_E_ARG;
    NEED(ctx, stats4 && C > 0 && W >= 1 && step >= 1, SNPG_E_ARG, "bad window arguments");
    NEED(ctx, W <= C, SNPG_E_INPUT, "not enough codons (%lld) for a sliding window of %lld", (long long)C, (long long)W);  // wgp:300
    NEED(ctx, W < ((int64_t)1 << 31), SNPG_E_ARG, "window too large");
    if (int rc = bind(ctx)) return rc;
    const int64_t nwin = (C - W) / step + 1;
    if (nwin_out) *nwin_out = nwin;
    if (!win_sums && !win_se) return SNPG_OK;
    if (int rc = upload_stats(ctx, stats4, keep, C)) return rc;
    if (win_sums) {
        CU(ctx, ctx->s_sums.ensure(sizeof(double) * 4 * (size_t)nwin));
        { Launch l(ctx, SNPG_K_WINDOW); launch_window_sums(ctx->s_in.as<double>(), W, step, nwin, ctx->s_sums.as<double>(), ctx->stream); }
        CU(ctx, cudaGetLastError());
        CU(ctx, cudaMemcpyAsync(win_sums, ctx->s_sums.p, sizeof(double) * 4 * (size_t)nwin, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (win_se && B >= 2) {
        const int64_t batch = std::max<int64_t>(1, std::min<int64_t>(nwin, ((int64_t)256 << 20) / (32 * B)));
        CU(ctx, ctx->s_rep.ensure(sizeof(double) * 4 * (size_t)(batch * B)));
        CU(ctx, ctx->s_se.ensure(sizeof(double) * 6 * (size_t)batch));
        std::vector<double> se6((size_t)batch * 6);
        const uint32_t sid = 0x20000u + ctx->call_id++;
        for (int64_t w0 = 0; w0 < nwin; w0 += batch) {
            const int64_t nb = std::min(batch, nwin - w0);
            { Launch l(ctx, SNPG_K_WINDOW); launch_window_bootstrap(ctx->s_in.as<double>(), W, step, w0, nb, B, ctx->seed, sid, ctx->s_rep.as<double>(), ctx->stream); }
            { Launch l(ctx, SNPG_K_MOMENTS); launch_moments(ctx->s_rep.as<double>(), nb, B, ctx->s_se.as<double>(), ctx->stream); }
            CU(ctx, cudaGetLastError());
            CU(ctx, cudaMemcpyAsync(se6.data(), ctx->s_se.p, sizeof(double) * 6 * (size_t)nb, cudaMemcpyDeviceToHost, ctx->stream));
            CU(ctx, sync_stream(ctx));
            for (int64_t k = 0; k < nb; k++)
                for (int q = 0; q < 3; q++) win_se[3 * (w0 + k) + q] = se6[(size_t)(6 * k + q)];
        }
    }
    CU(ctx, sync_stream(ctx));
    return SNPG_OK;
}

int snpg_within_stats_device(snpg_ctx *ctx, int group, double *d_stats4) {
    if (!ctx || !d_stats4) return SNPG_E_ARG;
    if (int rc = check_group(ctx, group)) return rc;
    if (int rc = bind(ctx)) return rc;
    const Group &g = ctx->groups[group];
    StatsOut o{nullptr, nullptr, nullptr, nullptr, d_stats4, nullptr};
    { Launch l(ctx, SNPG_K_STATS); launch_within_stats(view_of(ctx, g, 0), ctx->shard_count, ctx->d_tables.as<double>(), o, ctx->stream); }
    CU(ctx, cudaGetLastError());
    CU(ctx, sync_stream(ctx));
    return SNPG_OK;
}

int snpg_bootstrap_all_device(snpg_ctx *ctx, const double *d_stats4_full, int64_t b_first, int64_t b_count, double *d_rep_out) {
    if (!ctx || !d_stats4_full || !d_rep_out) return SNPG_E_ARG;
    NEED(ctx, ctx->n_products > 0, SNPG_E_STATE, "no products set");
    NEED(ctx, b_first >= 0 && b_count > 0, SNPG_E_ARG, "bad replicate range");
    if (int rc = bind(ctx)) return rc;
    {
        Launch l(ctx, SNPG_K_BOOT);
        launch_bootstrap(ctx->boot_kernel, d_stats4_full, ctx->d_full_ofs.as<int64_t>(), ctx->d_order_full.as<int>(), ctx->n_products, ctx->max_codons_full,
                         b_first, b_count, ctx->seed, 0x30000u, nullptr, d_rep_out, nullptr, nullptr, ctx->stream);
    }
    CU(ctx, cudaGetLastError());
    CU(ctx, sync_stream(ctx));
    return SNPG_OK;
}

int snpg_product_sums_device(snpg_ctx *ctx, const double *d_stats4_full, double *prod_sums) {
    if (!ctx || !d_stats4_full || !prod_sums) return SNPG_E_ARG;
    NEED(ctx, ctx->n_products > 0, SNPG_E_STATE, "no products set");
    if (int rc = bind(ctx)) return rc;
    CU(ctx, ctx->s_sums.ensure(sizeof(double) * 4 * (size_t)ctx->n_products));
    { Launch l(ctx, SNPG_K_SUMS); launch_product_sums(d_stats4_full, ctx->d_full_ofs.as<int64_t>(), ctx->n_products, ctx->s_sums.as<double>(), ctx->stream); }
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(prod_sums, ctx->s_sums.p, sizeof(double) * 4 * (size_t)ctx->n_products, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, sync_stream(ctx));
    return SNPG_OK;
}

int snpg_rep_moments_device(snpg_ctx *ctx, const double *d_rep, int64_t n_items, int64_t B, double *se_out) {
    if (!ctx || !d_rep || !se_out || n_items <= 0 || B < 2) return SNPG_E_ARG;
    if (int rc = bind(ctx)) return rc;
    CU(ctx, ctx->s_se.ensure(sizeof(double) * 6 * (size_t)n_items));
    { Launch l(ctx, SNPG_K_MOMENTS); launch_moments(d_rep, n_items, B, ctx->s_se.as<double>(), ctx->stream); }
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(se_out, ctx->s_se.p, sizeof(double) * 6 * (size_t)n_items, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, sync_stream(ctx));
    return SNPG_OK;
}

int snpg_between_all(snpg_ctx *ctx, const int *groups, int n_groups, int64_t min_taxa_total, int64_t min_taxa_group, int64_t B,
                     double *pair_sums, double *pair_se) {
    if (!ctx) return SNPG_E_ARG;
    NEED(ctx, groups && n_groups >= 2 && n_groups <= SNPG_MAX_GROUPS && pair_sums, SNPG_E_ARG, "bad between_all arguments");
    NEED(ctx, min_taxa_total >= 0 && min_taxa_group >= 0 && min_taxa_total < ((int64_t)1 << 32) && min_taxa_group < ((int64_t)1 << 32),
         SNPG_E_ARG, "bad min-taxa criteria");
    for (int i = 0; i < n_groups; i++)
        if (int rc = check_group(ctx, groups[i])) return rc;
    if (int rc = bind(ctx)) return rc;
    const int P = ctx->n_products;
    const int64_t cnt = ctx->shard_count;
    const int n_pairs = n_groups * (n_groups - 1) / 2;
    const bool boot = pair_se && B >= 2;
    ctx->stats_of = nullptr;
    CU(ctx, ctx->s_stats.ensure(sizeof(double) * 4 * (size_t)std::max<int64_t>(1, cnt)));
    CU(ctx, ctx->s_nda.ensure(sizeof(uint32_t) * (size_t)std::max<int64_t>(1, cnt)));
    CU(ctx, ctx->s_ndb.ensure(sizeof(uint32_t) * (size_t)std::max<int64_t>(1, cnt)));
    CU(ctx, ctx->s_sums.ensure(sizeof(double) * 4 * (size_t)(P * n_pairs)));
    if (boot) {
        CU(ctx, ctx->s_vals.ensure(sizeof(double) * 6 * (size_t)(B * P)));
        CU(ctx, ctx->s_se.ensure(sizeof(double) * 6 * (size_t)(P * n_pairs)));
    }
    int pair = 0;
    for (int i = 0; i < n_groups; i++) {
        for (int j = i + 1; j < n_groups; j++, pair++) {
            const Group &ga = ctx->groups[groups[i]], &gb = ctx->groups[groups[j]];
            StatsOut o{nullptr, ctx->s_nda.as<uint32_t>(), ctx->s_ndb.as<uint32_t>(), nullptr, ctx->s_stats.as<double>(), nullptr};
            { Launch l(ctx, SNPG_K_STATS); launch_between_stats(view_of(ctx, ga, 0), view_of(ctx, gb, 0), cnt, ctx->d_tables.as<double>(), o, ctx->stream); }
            if (min_taxa_total > 0 || min_taxa_group > 0) {
                Launch l(ctx, SNPG_K_OTHER);
                launch_taxa_filter(ctx->s_stats.as<double>(), ctx->s_nda.as<uint32_t>(), ctx->s_ndb.as<uint32_t>(), cnt,
                                   (uint32_t)min_taxa_total, (uint32_t)min_taxa_group, ctx->stream);
            }
            double *d_pair_sums = ctx->s_sums.as<double>() + 4 * (size_t)P * pair;
            if (!boot) { Launch l(ctx, SNPG_K_SUMS); launch_product_sums(ctx->s_stats.as<double>(), ctx->d_local_ofs.as<int64_t>(), P, d_pair_sums, ctx->stream); }
            if (boot) {
                {
                    Launch l(ctx, SNPG_K_BOOT);      // also writes the product sums
                    launch_bootstrap(ctx->boot_kernel, ctx->s_stats.as<double>(), ctx->d_local_ofs.as<int64_t>(), ctx->d_order_local.as<int>(), P,
                                     ctx->max_codons_local, 0, B, ctx->seed, 0x50000u + ctx->call_id++, nullptr, nullptr,
                                     ctx->s_vals.as<double>(), d_pair_sums, ctx->stream);
                }
                { Launch l(ctx, SNPG_K_MOMENTS); launch_moments_vals(ctx->s_vals.as<double>(), P, B, ctx->s_se.as<double>() + 6 * (size_t)P * pair, ctx->stream); }
            }
        }
    }
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(pair_sums, ctx->s_sums.p, sizeof(double) * 4 * (size_t)(P * n_pairs), cudaMemcpyDeviceToHost, ctx->stream));
    if (boot) CU(ctx, cudaMemcpyAsync(pair_se, ctx->s_se.p, sizeof(double) * 6 * (size_t)(P * n_pairs), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, sync_stream(ctx));
    return SNPG_OK;
}

// The whole within-group analysis of a loaded group, enqueued on ctx->stream.  final_sync = false and
// d_salt != null is the form a CUDA graph captures: no host synchronisation, outputs into pinned memory, the
// bootstrap's per-call stream number read from device memory when the graph runs.
static int within_all_impl(snpg_ctx *ctx, int group, int64_t B, double *prod_sums, double *prod_se, bool final_sync,
                           const uint32_t *d_salt) {
    const Group &g = ctx->groups[group];
    const int P = ctx->n_products;
    const int64_t cnt = ctx->shard_count;
    CU(ctx, ctx->s_stats.ensure(sizeof(double) * 4 * (size_t)std::max<int64_t>(1, cnt)));
    CU(ctx, ctx->s_sums.ensure(sizeof(double) * 4 * (size_t)P));
    StatsOut o{nullptr, nullptr, nullptr, nullptr, ctx->s_stats.as<double>(), nullptr};
    if (ctx->stats_of != &g) { Launch l(ctx, SNPG_K_STATS); launch_within_stats(view_of(ctx, g, 0), cnt, ctx->d_tables.as<double>(), o, ctx->stream); }
    ctx->stats_of = nullptr;         // (set by the load of a whole job: its fix-up kernel already wrote s_stats)
    const bool boot = prod_se && B >= 2;
    if (!boot) { Launch l(ctx, SNPG_K_SUMS); launch_product_sums(ctx->s_stats.as<double>(), ctx->d_local_ofs.as<int64_t>(), P, ctx->s_sums.as<double>(), ctx->stream); }
    if (boot) {
        CU(ctx, ctx->s_vals.ensure(sizeof(double) * 6 * (size_t)(B * P)));
        CU(ctx, ctx->s_se.ensure(sizeof(double) * 6 * (size_t)P));
        {
            Launch l(ctx, SNPG_K_BOOT);
            const uint32_t sid = d_salt ? 0x40000u : 0x40000u + ctx->call_id++;
            launch_bootstrap(ctx->boot_kernel, ctx->s_stats.as<double>(), ctx->d_local_ofs.as<int64_t>(), ctx->d_order_local.as<int>(), P,
                             ctx->max_codons_local, 0, B, ctx->seed, sid, d_salt, nullptr, ctx->s_vals.as<double>(),
                             ctx->s_sums.as<double>(), ctx->stream);      // also writes the product sums
        }
        { Launch l(ctx, SNPG_K_MOMENTS); launch_moments_vals(ctx->s_vals.as<double>(), P, B, ctx->s_se.as<double>(), ctx->stream); }
    }
    CU(ctx, cudaGetLastError());
    if (prod_sums) CU(ctx, cudaMemcpyAsync(prod_sums, ctx->s_sums.p, sizeof(double) * 4 * (size_t)P, cudaMemcpyDeviceToHost, ctx->stream));
    if (boot) CU(ctx, cudaMemcpyAsync(prod_se, ctx->s_se.p, sizeof(double) * 6 * (size_t)P, cudaMemcpyDeviceToHost, ctx->stream));
    if (final_sync) CU(ctx, sync_stream(ctx));
    return SNPG_OK;
}

int snpg_within_all(snpg_ctx *ctx, int group, int64_t B, double *prod_sums, double *prod_se) {
    if (!ctx) return SNPG_E_ARG;
    if (int rc = check_group(ctx, group)) return rc;
    if (int rc = bind(ctx)) return rc;
    return within_all_impl(ctx, group, B, prod_sums, prod_se, true, nullptr);
}

static void drop_job_graph(snpg_ctx *ctx) { invalidate_job_graph(ctx); }

// Captures the job (load + analysis) into a CUDA graph and runs it once.  Any failure leaves no graph behind
// and reports SNPG_E_STATE so that the caller runs the job the ordinary way.
static int capture_job_graph(snpg_ctx *ctx, const snpg_ctx::JobKey &key, int64_t B) {
    auto &J = ctx->job;
    const int P = ctx->n_products;
    if (!J.h_salt && cudaMallocHost(reinterpret_cast<void **>(&J.h_salt), sizeof(uint32_t)) != cudaSuccess) return SNPG_E_STATE;
    if (J.h_out_products < P) {
        if (J.h_out) cudaFreeHost(J.h_out);
        J.h_out = nullptr;
        if (cudaMallocHost(reinterpret_cast<void **>(&J.h_out), sizeof(double) * 10 * (size_t)P) != cudaSuccess) return SNPG_E_STATE;
        J.h_out_products = P;
    }
    if (ctx->d_salt.ensure(sizeof(uint32_t)) != cudaSuccess) return SNPG_E_STATE;
    const uint64_t epoch0 = g_alloc_epoch;
    const int64_t l0 = ctx->launches, t0 = ctx->tile_launches;
    if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); return SNPG_E_STATE; }
    ctx->capturing = true;
    int rc = SNPG_OK;
    if (cudaMemcpyAsync(ctx->d_salt.p, J.h_salt, sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) rc = SNPG_E_CUDA;
    if (rc == SNPG_OK) {
        ctx->defer_load_sync = true;
        rc = snpg_load_group_device(ctx, key.group, key.bases, key.n, key.len, key.stride);
        ctx->defer_load_sync = false;
    }
    if (rc == SNPG_OK) rc = within_all_impl(ctx, key.group, B, J.h_out, key.want_se ? J.h_out + 4 * P : nullptr, false, ctx->d_salt.as<uint32_t>());
    ctx->capturing = false;
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
    cudaGraphExec_t exec = nullptr;
    if (rc == SNPG_OK && e == cudaSuccess && graph && epoch0 == g_alloc_epoch && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
        J.exec = exec;
        J.key = key;
        J.epoch = g_alloc_epoch;
        J.n_launches = ctx->launches - l0;
        J.n_tile_launches = ctx->tile_launches - t0;
    } else {
        rc = SNPG_E_STATE;
        cudaGetLastError();
    }
    if (graph) cudaGraphDestroy(graph);
    ctx->launches = l0;                 // nothing ran yet: the replay below counts them
    ctx->tile_launches = t0;
    return rc;
}

int snpg_within_job(snpg_ctx *ctx, int group, const uint8_t *bases, int bases_on_device, uint64_t n_seqs, uint64_t aln_len,
                    uint64_t row_stride, int64_t B, double *prod_sums, double *prod_se) {
    if (!ctx) return SNPG_E_ARG;
    auto &J = ctx->job;
    snpg_ctx::JobKey key;
    key.group = group; key.bases = bases; key.n = n_seqs; key.len = aln_len; key.stride = row_stride; key.B = B;
    key.keep_codes = ctx->keep_codes; key.site_counts = ctx->site_counts; key.want_se = prod_se != nullptr; key.fused_kernel = ctx->fused_kernel;
    static const bool graphs_on = getenv("SNPG_NO_GRAPH") == nullptr;
    key.profile_mask = ctx->profile_mask;
    const bool eligible = graphs_on && bases_on_device && bases && prod_sums && !J.disabled &&
                          group >= 0 && group < SNPG_MAX_GROUPS && ctx->n_products > 0;
    if (eligible) {
        if (int rc = bind(ctx)) return rc;
        if (J.exec && (!(J.key == key) || J.epoch != g_alloc_epoch)) drop_job_graph(ctx);
        if (!J.exec && J.candidate == key && J.candidate_epoch == g_alloc_epoch) {
            // the same job for the second time in a row with every buffer already in place: capture it
            if (capture_job_graph(ctx, key, B) != SNPG_OK) { drop_job_graph(ctx); J.disabled = true; }
        }
        if (J.exec) {
            const int P = ctx->n_products;
            *J.h_salt = ctx->call_id++;
            CU(ctx, cudaGraphLaunch(J.exec, ctx->stream));
            CU(ctx, sync_stream(ctx));
            ctx->launches += J.n_launches;
            ctx->tile_launches += J.n_tile_launches;
            J.replays++;
            for (auto &sp : J.spans) {
                float ms = 0;
                if (cudaEventElapsedTime(&ms, sp.e0, sp.e1) == cudaSuccess) { ctx->kernel_ms[sp.kind] += ms; ctx->kernel_calls[sp.kind]++; }
            }
            memcpy(prod_sums, J.h_out, sizeof(double) * 4 * (size_t)P);
            if (prod_se) memcpy(prod_se, J.h_out + 4 * P, sizeof(double) * 6 * (size_t)P);
            return SNPG_OK;
        }
    }
    ctx->defer_load_sync = true;     // one synchronisation for the whole job, at the end of the analysis
    int rc = bases_on_device ? snpg_load_group_device(ctx, group, bases, n_seqs, aln_len, row_stride)
                             : snpg_load_group(ctx, group, bases, n_seqs, aln_len, row_stride);
    ctx->defer_load_sync = false;
    if (rc != SNPG_OK) { ctx->stats_of = nullptr; sync_stream(ctx); return rc; }
    rc = snpg_within_all(ctx, group, B, prod_sums, prod_se);
    if (eligible && rc == SNPG_OK) { J.candidate = key; J.candidate_epoch = g_alloc_epoch; }
    return rc;
}

int64_t snpg_graph_replays(const snpg_ctx *ctx) { return ctx ? ctx->job.replays : 0; }

int snpg_test_philox(snpg_ctx *ctx, const uint32_t *ctr4, const uint32_t *key2, uint32_t *out4) {
    if (!ctx || !ctr4 || !key2 || !out4) return SNPG_E_ARG;
    if (int rc = bind(ctx)) return rc;
    CU(ctx, ctx->s_misc.ensure(16));
    { Launch l(ctx, SNPG_K_OTHER); launch_philox_kat(ctr4, key2, ctx->s_misc.as<uint32_t>(), ctx->stream); }
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(out4, ctx->s_misc.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, sync_stream(ctx));
    return SNPG_OK;
}

int64_t snpg_launch_count(const snpg_ctx *ctx) { return ctx ? ctx->launches : 0; }
int64_t snpg_tile_launch_count(const snpg_ctx *ctx) { return ctx ? ctx->tile_launches : 0; }

int snpg_set_stream(snpg_ctx *ctx, void *cuda_stream) {
    if (!ctx) return SNPG_E_ARG;
    if (int rc = bind(ctx)) return rc;
    CU(ctx, sync_stream(ctx));
    invalidate_job_graph(ctx);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (cuda_stream) { ctx->stream = static_cast<cudaStream_t>(cuda_stream); ctx->own_stream = false; }
    else { CU(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->own_stream = true; }
    return SNPG_OK;
}

int snpg_get_profile(snpg_ctx *ctx, double *ms, int64_t *calls, int reset) {
    if (!ctx) return SNPG_E_ARG;
    for (int k = 0; k < SNPG_N_KERNEL_KINDS; k++) {
        if (ms) ms[k] = ctx->kernel_ms[k];
        if (calls) calls[k] = ctx->kernel_calls[k];
        if (reset) { ctx->kernel_ms[k] = 0; ctx->kernel_calls[k] = 0; }
    }
    return SNPG_OK;
}

}  // extern "C"

// C-ABI of libsnpgenie_cuda.so: context, device memory, streams, and the
// sequencing of kernels.  See include/snpgenie_cuda.h for the contract and the
// reference blocks each entry point replaces.  There is no CPU fallback here:
// every compute entry point runs the CUDA kernels or fails with SNPG_E_CUDA.
// (Whole jobs: job.cu; several GPUs: exchange.cu; the FASTA reader: ingest.cu.)
#include <algorithm>
#include <cstdlib>

#include "snpg_ctx.h"
#include "snpg_team.h"

using namespace snpg;

namespace snpg {

static std::string g_create_error;
std::atomic<uint64_t> g_alloc_epoch{0};

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

void resolve_spans(snpg_ctx *ctx) {
    for (auto &sp : ctx->spans) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, sp.e0, sp.e1) == cudaSuccess) { ctx->kernel_ms[sp.kind] += ms; ctx->kernel_calls[sp.kind]++; }
        ctx->free_events.push_back(sp.e0);
        ctx->free_events.push_back(sp.e1);
    }
    ctx->spans.clear();
}

void invalidate_job_graph(snpg_ctx *ctx) {
    for (auto &J : ctx->job_graph) {
        if (J.exec) cudaGraphExecDestroy(J.exec);
        J.exec = nullptr;
        for (auto &sp : J.spans) { ctx->free_events.push_back(sp.e0); ctx->free_events.push_back(sp.e1); }
        J.spans.clear();
    }
    ctx->job_candidate_epoch = ~0ull;
}

cudaError_t sync_stream(snpg_ctx *ctx) {
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess && !ctx->spans.empty()) resolve_spans(ctx);
    return e;
}

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
    v.first_def = g.has_first_def ? g.first_def.as<uint8_t>() + c0 : nullptr;
    v.n = (int64_t)g.n;
    v.n_pad = g.n_pad;
    return v;
}

// Local slice [c0, c0+C) of a product inside this shard.
void product_slice(const snpg_ctx *ctx, int product, int64_t *c0, int64_t *C) {
    *c0 = ctx->local_ofs[product];
    *C = ctx->local_ofs[product + 1] - ctx->local_ofs[product];
}

}  // namespace snpg

namespace {

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
    g.has_first_def = ctx->first_defined && !ctx->keep_codes && cnt > 0;
    if (g.has_first_def) {
        CU(ctx, g.first_def.ensure((size_t)cnt));
        CU(ctx, cudaMemsetAsync(g.first_def.p, 0xFF, (size_t)cnt, ctx->stream));
    }
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
                         g.clear_in_vote ? g.hist.as<uint32_t>() : nullptr, omin, omax, ctx->shard_count, ctx->d_sched.as<uint32_t>(), ctx->stream);
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
            bool done = false;
            if (ctx->fused_kernel == SNPG_FUSED_TMA_SPANS && make_span_tensor_map(tmap, d_block, pitch, rows, row_bytes, ctx->tma_l2_promotion)) {
                const cudaError_t e = launch_span_hist(tmap, rows, ctx->d_cons.as<uint8_t>(), ctx->d_alt.as<uint32_t>(), ctx->n_spans, ctx->d_runs.as<SpanRun>(),
                                                       g.hist.as<uint32_t>(), omin, omax, ctx->d_sched.as<uint32_t>(), ctx->n_sms, ctx->stream);
                if (e == cudaSuccess) { done = true; ctx->span_tma_launches++; }
                else if (e != cudaErrorInvalidValue) CU(ctx, e);
                else cudaGetLastError();
            }
            if (done) {
            } else if (ctx->fused_kernel == SNPG_FUSED_TMA_TILES && ctx->n_tiles > 0 &&
                make_tile_tensor_map(tmap, d_block, pitch, rows, row_bytes, ctx->tma_l2_promotion)) {
                CU(ctx, launch_tile_hist(tmap, rows, ctx->d_cons.as<uint8_t>(), ctx->d_tiles.as<CodonTile>(), ctx->n_tiles, ctx->n_sms,
                                         ctx->d_sched.as<uint32_t>(), g.hist.as<uint32_t>(), omin, omax, ctx->stream));
                ctx->tile_launches++;
            } else {
                launch_fused_hist(d_block, pitch, rows, ctx->d_cons.as<uint8_t>(), ctx->d_alt.as<uint32_t>(), ctx->n_spans, ctx->d_runs.as<SpanRun>(),
                                  g.hist.as<uint32_t>(), omin, omax, ctx->d_sched.as<uint32_t>(), ctx->n_sms, ctx->stream);
            }
        }
        if (g.has_first_def) {
            Launch l(ctx, SNPG_K_OTHER);
            launch_first_defined(d_block + (ctx->col_lo - ctx->base_al), pitch, rows, ctx->d_map.as<CodonMap>(), ctx->shard_count,
                                 g.first_def.as<uint8_t>(), ctx->stream);
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
        // a whole job (job.cu: the analysis follows the load at once) also takes the per-codon statistics here, into the
        // buffers the job chose (ctx->job_out: local scratch, or every rank's exchange region through job_peer_out)
        const bool with_stats = ctx->defer_load_sync && ctx->n_irr == 0 && !ctx->site_counts;
        ctx->stats_of = nullptr;
        if (with_stats) {
            Launch l(ctx, SNPG_K_OTHER);
            launch_fused_fixup_stats(ctx->d_cons.as<uint8_t>(), ctx->col_lo - ctx->base_al, ctx->d_alt.as<uint32_t>(), ctx->n_spans, ctx->d_map.as<CodonMap>(),
                                     ctx->shard_count, (uint32_t)g.n, g.hist.as<uint32_t>(), omin, omax, view_of(ctx, g, 0),
                                     ctx->d_tables.as<double>(), ctx->job_out, &ctx->job_peer_out, &ctx->job_peer_sync, ctx->stream);
            ctx->stats_of = &g;
        } else {
            Launch l(ctx, SNPG_K_OTHER);
            launch_fused_fixup(ctx->d_cons.as<uint8_t>(), ctx->col_lo - ctx->base_al, ctx->d_alt.as<uint32_t>(), ctx->n_spans, ctx->d_map.as<CodonMap>(), ctx->d_regular.as<uint8_t>(),
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
    if (!ctx->defer_load_sync && !ctx->load_no_sync) CU(ctx, sync_stream(ctx));
    g.loaded = true;
    return SNPG_OK;
}

}  // namespace

namespace snpg {
int boot_classes_scratch(snpg_ctx *ctx, BootArgs &a, int64_t total) {
    if (ctx->boot_kernel != SNPG_BOOT_CLASSES) return SNPG_OK;
    CU(ctx, ctx->s_boot.ensure(boot_scratch_bytes(total, a.n_items, a.b_count)));
    boot_scratch_carve(ctx->s_boot.p, total, a.n_items, a.b_count, a);
    ctx->launches += 2;        // k_boot_classify, k_boot_split
    return SNPG_OK;
}
}  // namespace snpg

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
    Trace trace("snpg_create");
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
    if ((e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->win_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_win, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_salt_fork, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_salt, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if (cudaDeviceGetAttribute(&ctx->n_sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || ctx->n_sms <= 0) ctx->n_sms = 148;
    if (const char *v = getenv("SNPG_TMA_L2PROMO")) ctx->tma_l2_promotion = atoi(v);
    if ((e = ctx->d_sched.ensure(4 * sizeof(uint32_t))) != cudaSuccess) return bail("cudaMalloc(sched)", e);
    if ((e = cudaMemset(ctx->d_sched.p, 0, 4 * sizeof(uint32_t))) != cudaSuccess) return bail("cudaMemset(sched)", e);
    if ((e = prepare_device_kernels()) != cudaSuccess) return bail("loading the kernels", e);
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
    if (is_team(ctx)) {
        for (snpg_ctx *c : ctx->team) { c->team_parent = nullptr; snpg_destroy(c); }
        delete ctx;
        return;
    }
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    exchange_release(ctx);
    for (auto &g : ctx->groups) { g.codes.release(); g.hist.release(); g.other.release(); g.sites.release(); g.first_def.release(); }
    for (auto &h : ctx->host_aln) if (h.p) { ingest_forget(h.p); cudaFreeHost(h.p); }
    invalidate_job_graph(ctx);
    if (ctx->h_salt) cudaFreeHost(ctx->h_salt);
    if (ctx->h_out) cudaFreeHost(ctx->h_out);
    DevBuf *bufs[] = {&ctx->d_map, &ctx->d_local_ofs, &ctx->d_full_ofs, &ctx->d_tables, &ctx->staging[0], &ctx->staging[1],
                      &ctx->d_runs, &ctx->d_tiles, &ctx->d_sched, &ctx->d_salt, &ctx->d_regular, &ctx->d_irr_map, &ctx->d_irr_idx, &ctx->d_cons, &ctx->d_alt, &ctx->t_codes, &ctx->t_hist,
                      &ctx->t_other, &ctx->wplan.beg, &ctx->wplan.end, &ctx->w_sums, &ctx->w_vals, &ctx->w_se, &ctx->s_undef,
                      &ctx->s_stats, &ctx->s_poly, &ctx->s_nda, &ctx->s_ndb, &ctx->s_cmp, &ctx->s_cons, &ctx->s_sums,
                      &ctx->s_boot, &ctx->s_rep, &ctx->s_se, &ctx->s_misc, &ctx->s_in, &ctx->s_vals, &ctx->d_order_local, &ctx->d_order_full};
    for (DevBuf *b : bufs) b->release();
    for (int i = 0; i < 2; i++) {
        if (ctx->ev_copied[i]) cudaEventDestroy(ctx->ev_copied[i]);
        if (ctx->ev_packed[i]) cudaEventDestroy(ctx->ev_packed[i]);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    resolve_spans(ctx);
    for (cudaEvent_t e : ctx->free_events) cudaEventDestroy(e);
    if (ctx->stream && ctx->own_stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->ev_win) cudaEventDestroy(ctx->ev_win);
    if (ctx->ev_salt_fork) cudaEventDestroy(ctx->ev_salt_fork);
    if (ctx->ev_salt) cudaEventDestroy(ctx->ev_salt);
    if (ctx->win_stream) cudaStreamDestroy(ctx->win_stream);
    for (int l = 0; l < snpg_ctx::kPairLanes; l++) {
        if (ctx->pair_done[l]) cudaEventDestroy(ctx->pair_done[l]);
        if (ctx->pair_stream[l]) cudaStreamDestroy(ctx->pair_stream[l]);
    }
    delete ctx;
}

int snpg_set_option(snpg_ctx *ctx, int option, int64_t value) {
    if (!ctx) return SNPG_E_ARG;
    TEAM_EACH(ctx, snpg_set_option(c, option, value));
    invalidate_job_graph(ctx);
    if (option == SNPG_OPT_KEEP_CODES) { ctx->keep_codes = value != 0; return SNPG_OK; }
    if (option == SNPG_OPT_BOOT_SLOTS) {
        NEED(ctx, value == SNPG_SLOTS_CORRECTED || value == SNPG_SLOTS_FAITHFUL, SNPG_E_ARG, "unknown slot rule %lld", (long long)value);
        ctx->boot_slot0 = (int)value;
        return SNPG_OK;
    }
    if (option == SNPG_OPT_SITE_COUNTS) {
        NEED(ctx, ctx->n_products == 0, SNPG_E_STATE, "SNPG_OPT_SITE_COUNTS must be set before snpg_set_products");
        ctx->site_counts = value != 0;
        return SNPG_OK;
    }
    if (option == SNPG_OPT_FUSED_KERNEL) {
        NEED(ctx, value == SNPG_FUSED_TMA_TILES || value == SNPG_FUSED_LDG_SPANS || value == SNPG_FUSED_TMA_SPANS, SNPG_E_ARG, "unknown fused kernel %lld", (long long)value);
        ctx->fused_kernel = (int)value;
        return SNPG_OK;
    }
    if (option == SNPG_OPT_BOOTSTRAP_KERNEL) {
        NEED(ctx, value == SNPG_BOOT_GATHER || value == SNPG_BOOT_WEIGHTS || value == SNPG_BOOT_CLASSES, SNPG_E_ARG, "unknown bootstrap kernel %lld", (long long)value);
        ctx->boot_kernel = (int)value;
        return SNPG_OK;
    }
    if (option == SNPG_OPT_PROFILE) { ctx->profile_mask = value < 0 ? 0xFFFFFFFFu : (uint32_t)value; return SNPG_OK; }
    if (option == SNPG_OPT_INGEST_ASYNC) { ctx->ingest_async = value != 0; return SNPG_OK; }
    if (option == SNPG_OPT_FIRST_DEFINED) { ctx->first_defined = value != 0; return SNPG_OK; }
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
    TEAM_UNSUPPORTED(ctx, "snpg_set_shard");
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
    if (is_team(ctx)) {
        for (snpg_ctx *c : ctx->team)
            if (int rc = snpg_set_products(c, n_products, seg_ofs, seg_start, seg_stop, strand, aln_len)) return team_fail(ctx, c, rc);
        return team_wire(ctx, 10000);      // grows on demand (a job with more replicates re-wires)
    }
    if (int rc = bind(ctx)) return rc;
    for (auto &g : ctx->groups) g.loaded = false;  // maps change: loaded groups are stale
    invalidate_job_graph(ctx);
    exchange_release(ctx);                         // the exchange region is laid out for one annotation
    ctx->wplan = snpg_ctx::WindowPlan();

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
        CU(ctx, ctx->d_alt.ensure(sizeof(uint32_t) * (size_t)alt_words(std::max<int64_t>(1, n_spans))));
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

int snpg_n_products(const snpg_ctx *ctx) { return !ctx ? SNPG_E_ARG : is_team(ctx) ? ctx->team[0]->n_products : ctx->n_products; }

int snpg_n_codons(const snpg_ctx *ctx, int product, int64_t *out) {
    if (is_team(ctx)) ctx = ctx->team[0];
    if (!ctx || !out || product < 0 || product >= ctx->n_products) return SNPG_E_ARG;
    *out = ctx->full_ofs[product + 1] - ctx->full_ofs[product];
    return SNPG_OK;
}

int snpg_shard_range(const snpg_ctx *ctx, int64_t *first, int64_t *count, int64_t *total) {
    if (is_team(ctx)) {            // the team as a whole owns the whole axis
        const snpg_ctx *c = ctx->team[0];
        if (c->n_products == 0) return SNPG_E_STATE;
        if (first) *first = 0;
        if (count) *count = c->full_ofs[c->n_products];
        if (total) *total = c->full_ofs[c->n_products];
        return SNPG_OK;
    }
    if (!ctx || ctx->n_products == 0) return SNPG_E_STATE;
    if (first) *first = ctx->shard_first;
    if (count) *count = ctx->shard_count;
    if (total) *total = ctx->full_ofs[ctx->n_products];
    return SNPG_OK;
}

int snpg_load_group(snpg_ctx *ctx, int group, const uint8_t *bases, uint64_t n_seqs, uint64_t aln_len, uint64_t row_stride) {
    if (!ctx) return SNPG_E_ARG;
    TEAM_EACH(ctx, snpg_load_group(c, group, bases, n_seqs, aln_len, row_stride));
    NEED(ctx, ctx->n_products > 0, SNPG_E_STATE, "set products before loading a group");
    NEED(ctx, group >= 0 && group < SNPG_MAX_GROUPS, SNPG_E_ARG, "group %d out of range", group);
    NEED(ctx, bases && n_seqs > 0 && n_seqs < ((uint64_t)1 << 31), SNPG_E_ARG, "bad alignment arguments");
    NEED(ctx, (int64_t)aln_len == ctx->aln_len, SNPG_E_INPUT,
         "alignment length %llu differs from the annotation's %lld", (unsigned long long)aln_len, (long long)ctx->aln_len);
    NEED(ctx, row_stride >= aln_len, SNPG_E_ARG, "row_stride < aln_len");
    if (int rc = bind(ctx)) return rc;
    Trace trace("snpg_load_group");
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
            ingest_wait_rows(bases, (uint64_t)(r0 + rows));       // a FASTA slot still being parsed: the copy of this block overlaps the parse of the next
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
    TEAM_UNSUPPORTED(ctx, "snpg_load_group_device (a device pointer belongs to one GPU; use snpg_team_ctx)");
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

int snpg_drop_group(snpg_ctx *ctx, int group) {
    if (!ctx || group < 0 || group >= SNPG_MAX_GROUPS) return SNPG_E_ARG;
    TEAM_EACH(ctx, snpg_drop_group(c, group));
    if (int rc = bind(ctx)) return rc;
    Group &g = ctx->groups[group];
    g.codes.release(); g.hist.release(); g.other.release(); g.sites.release(); g.first_def.release();
    g.loaded = false; g.has_codes = false; g.has_sites = false; g.has_first_def = false;
    return SNPG_OK;
}

int snpg_group_size(const snpg_ctx *ctx, int group, uint64_t *n_seqs) {
    if (is_team(ctx)) ctx = ctx->team[0];
    if (!ctx || group < 0 || group >= SNPG_MAX_GROUPS || !n_seqs || !ctx->groups[group].loaded) return SNPG_E_ARG;
    *n_seqs = ctx->groups[group].n;
    return SNPG_OK;
}

int snpg_get_codon_indices(snpg_ctx *ctx, int group, int product, uint8_t *out) {
    if (!ctx || !out) return SNPG_E_ARG;
    TEAM_UNSUPPORTED(ctx, "snpg_get_codon_indices");
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
    if (is_team(ctx)) {            // every child fills the codons of its range; product-local offsets follow the global axis
        for (snpg_ctx *c : ctx->team) {
            if (product < 0 || product >= c->n_products) return fail(ctx, SNPG_E_ARG, "product %d out of range", product);
            const int64_t at = std::max<int64_t>(0, c->shard_first - c->full_ofs[product]);
            if (int rc = snpg_get_hist(c, group, product, out + at * kHistBins, other_distinct ? other_distinct + at : nullptr)) return team_fail(ctx, c, rc);
        }
        return SNPG_OK;
    }
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
    TEAM_UNSUPPORTED(ctx, "snpg_get_site_counts");
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
        else launch_within_stats(view_of(ctx, ga, c0), C, ctx->d_tables.as<double>(), o, nullptr, nullptr, ctx->stream);
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
    if (is_team(ctx)) {
        for (snpg_ctx *c : ctx->team) {
            if (product < 0 || product >= c->n_products) return fail(ctx, SNPG_E_ARG, "product %d out of range", product);
            const int64_t at = std::max<int64_t>(0, c->shard_first - c->full_ofs[product]);
            if (int rc = snpg_within(c, group, product, poly ? poly + at : nullptr, n_defined ? n_defined + at : nullptr,
                                     comparisons ? comparisons + at : nullptr, stats4 ? stats4 + 4 * at : nullptr,
                                     conserved_code ? conserved_code + at : nullptr)) return team_fail(ctx, c, rc);
        }
        return SNPG_OK;
    }
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
    TEAM_UNSUPPORTED(ctx, "snpg_between");
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

static uint2 philox_key(const snpg_ctx *ctx) { return make_uint2((uint32_t)ctx->seed, (uint32_t)(ctx->seed >> 32)); }

int snpg_bootstrap_product(snpg_ctx *ctx, const double *stats4, const uint8_t *keep, int64_t C, int64_t B, double *sums_out,
                           double *se_out) {
    if (!ctx) return SNPG_E_ARG;
    TEAM_FORWARD0(ctx, snpg_bootstrap_product(c, stats4, keep, C, B, sums_out, se_out));
    NEED(ctx, stats4 && C > 0 && C < ((int64_t)1 << 31) - 1 && B >= 2 && B < ((int64_t)1 << 36), SNPG_E_ARG, "bad bootstrap arguments (C=%lld, B=%lld)",
         (long long)C, (long long)B);
    if (int rc = bind(ctx)) return rc;
    if (int rc = upload_stats(ctx, stats4, keep, C)) return rc;
    const int64_t ofs[2] = {0, C};
    CU(ctx, ctx->s_misc.ensure(sizeof ofs));
    CU(ctx, cudaMemcpy(ctx->s_misc.p, ofs, sizeof ofs, cudaMemcpyHostToDevice));
    CU(ctx, ctx->s_rep.ensure(sizeof(double) * 4 * (size_t)B));
    CU(ctx, ctx->s_se.ensure(sizeof(double) * 6));
    CU(ctx, ctx->s_vals.ensure(sizeof(double) * 6 * (size_t)B));
    {
        Launch l(ctx, SNPG_K_BOOT);
        BootArgs a;
        a.stats4 = ctx->s_in.as<double>(); a.beg = ctx->s_misc.as<int64_t>(); a.end = a.beg + 1; a.n_items = 1; a.slot0 = ctx->boot_slot0;
        a.b_count = B; a.key = philox_key(ctx); a.kind = kBootProduct; a.call = ctx->call_id++;
        a.rep = sums_out ? ctx->s_rep.as<double>() : nullptr; a.vals = ctx->s_vals.as<double>();
        if (int rc = boot_classes_scratch(ctx, a, C)) return rc;
        launch_bootstrap(ctx->boot_kernel, a, C, ctx->stream);
    }
    { Launch l(ctx, SNPG_K_MOMENTS); launch_moments_vals(ctx->s_vals.as<double>(), 1, B, ctx->s_se.as<double>(), nullptr, nullptr, ctx->stream); }
    CU(ctx, cudaGetLastError());
    if (sums_out) CU(ctx, cudaMemcpyAsync(sums_out, ctx->s_rep.p, sizeof(double) * 4 * (size_t)B, cudaMemcpyDeviceToHost, ctx->stream));
    if (se_out) CU(ctx, cudaMemcpyAsync(se_out, ctx->s_se.p, sizeof(double) * 6, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, sync_stream(ctx));
    return SNPG_OK;
}

int snpg_windows(snpg_ctx *ctx, const double *stats4, const uint8_t *keep, int64_t C, int64_t W, int64_t step, int64_t B,
                 double *win_sums, double *win_se, int64_t *nwin_out) {
    if (!ctx) return SNPG_E_ARG;
    TEAM_FORWARD0(ctx, snpg_windows(c, stats4, keep, C, W, step, B, win_sums, win_se, nwin_out));
    NEED(ctx, stats4 && C > 0 && W >= 1 && step >= 1, SNPG_E_ARG, "bad window arguments");
    NEED(ctx, W <= C, SNPG_E_INPUT, "not enough codons (%lld) for a sliding window of %lld", (long long)C, (long long)W);  // wgp:300
    NEED(ctx, W < ((int64_t)1 << 31) && B < ((int64_t)1 << 36), SNPG_E_ARG, "window or replicate count too large");
    if (int rc = bind(ctx)) return rc;
    const int64_t nwin = (C - W) / step + 1;
    NEED(ctx, nwin < ((int64_t)1 << 24), SNPG_E_ARG, "too many windows (%lld)", (long long)nwin);
    if (nwin_out) *nwin_out = nwin;
    if (!win_sums && !win_se) return SNPG_OK;
    if (int rc = upload_stats(ctx, stats4, keep, C)) return rc;
    // window k = the item [k * step, k * step + W) of the uploaded table
    std::vector<int64_t> be((size_t)(2 * nwin));
    for (int64_t k = 0; k < nwin; k++) { be[(size_t)k] = k * step; be[(size_t)(nwin + k)] = k * step + W; }
    CU(ctx, ctx->s_misc.ensure(sizeof(int64_t) * be.size()));
    CU(ctx, cudaMemcpy(ctx->s_misc.p, be.data(), sizeof(int64_t) * be.size(), cudaMemcpyHostToDevice));
    const int64_t *d_beg = ctx->s_misc.as<int64_t>(), *d_end = d_beg + nwin;
    if (win_sums) {
        CU(ctx, ctx->s_sums.ensure(sizeof(double) * 4 * (size_t)nwin));
        { Launch l(ctx, SNPG_K_WINDOW); launch_window_sums(ctx->s_in.as<double>(), d_beg, d_end, nwin, ctx->s_sums.as<double>(), ctx->stream); }
        CU(ctx, cudaGetLastError());
        CU(ctx, cudaMemcpyAsync(win_sums, ctx->s_sums.p, sizeof(double) * 4 * (size_t)nwin, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (win_se && B >= 2) {
        // the window bootstrap (wgp:384-427; bgp:604-736) is the resampling kernel over window items: W draws uniform over
        // the window's W codons.  Windows go out in batches that keep the per-replicate values under 256 MB.
        const int64_t batch = std::max<int64_t>(1, std::min<int64_t>(nwin, ((int64_t)256 << 20) / (48 * B)));
        CU(ctx, ctx->s_vals.ensure(sizeof(double) * 6 * (size_t)(batch * B)));
        CU(ctx, ctx->s_se.ensure(sizeof(double) * 6 * (size_t)batch));
        std::vector<double> se6((size_t)batch * 6);
        const uint32_t call = ctx->call_id++;
        for (int64_t w0 = 0; w0 < nwin; w0 += batch) {
            const int64_t nb = std::min(batch, nwin - w0);
            {
                Launch l(ctx, SNPG_K_WINDOW);
                BootArgs a;
                a.stats4 = ctx->s_in.as<double>(); a.beg = d_beg + w0; a.end = d_end + w0; a.n_items = (int)nb; a.slot0 = 0;
                a.b_count = B; a.key = philox_key(ctx); a.kind = kBootWindow; a.call = call;
                a.vals = ctx->s_vals.as<double>();
                a.item0 = (uint32_t)w0;             // windows keep their own Philox streams whatever the batching
                launch_bootstrap(ctx->boot_kernel, a, W, ctx->stream);
            }
            { Launch l(ctx, SNPG_K_MOMENTS); launch_moments_vals(ctx->s_vals.as<double>(), nb, B, ctx->s_se.as<double>(), nullptr, nullptr, ctx->stream); }
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
    TEAM_UNSUPPORTED(ctx, "snpg_within_stats_device");
    if (int rc = check_group(ctx, group)) return rc;
    if (int rc = bind(ctx)) return rc;
    const Group &g = ctx->groups[group];
    StatsOut o{nullptr, nullptr, nullptr, nullptr, d_stats4, nullptr};
    { Launch l(ctx, SNPG_K_STATS); launch_within_stats(view_of(ctx, g, 0), ctx->shard_count, ctx->d_tables.as<double>(), o, nullptr, nullptr, ctx->stream); }
    CU(ctx, cudaGetLastError());
    CU(ctx, sync_stream(ctx));
    return SNPG_OK;
}

// (The Philox stream of this call does not depend on the ctx's call counter: ranks that split the replicates of one
// analysis must draw from one stream whatever else they called before.  Repeated calls repeat the replicates.)
int snpg_bootstrap_all_device(snpg_ctx *ctx, const double *d_stats4_full, int64_t b_first, int64_t b_count, double *d_rep_out) {
    if (!ctx || !d_stats4_full || !d_rep_out) return SNPG_E_ARG;
    TEAM_UNSUPPORTED(ctx, "snpg_bootstrap_all_device");
    NEED(ctx, ctx->n_products > 0, SNPG_E_STATE, "no products set");
    NEED(ctx, b_first >= 0 && b_count > 0 && b_first + b_count < ((int64_t)1 << 36), SNPG_E_ARG, "bad replicate range");
    if (int rc = bind(ctx)) return rc;
    {
        Launch l(ctx, SNPG_K_BOOT);
        BootArgs a;
        a.stats4 = d_stats4_full; a.beg = ctx->d_full_ofs.as<int64_t>(); a.end = a.beg + 1; a.order = ctx->d_order_full.as<int>();
        a.n_items = ctx->n_products; a.slot0 = ctx->boot_slot0; a.b_first = b_first; a.b_count = b_count;
        a.key = philox_key(ctx); a.kind = kBootAllDevice; a.call = 0; a.rep = d_rep_out;
        if (int rc = boot_classes_scratch(ctx, a, ctx->full_ofs[(size_t)ctx->n_products])) return rc;
        launch_bootstrap(ctx->boot_kernel, a, ctx->max_codons_full, ctx->stream);
    }
    CU(ctx, cudaGetLastError());
    CU(ctx, sync_stream(ctx));
    return SNPG_OK;
}

int snpg_product_sums_device(snpg_ctx *ctx, const double *d_stats4_full, double *prod_sums) {
    if (!ctx || !d_stats4_full || !prod_sums) return SNPG_E_ARG;
    TEAM_UNSUPPORTED(ctx, "snpg_product_sums_device");
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
    TEAM_UNSUPPORTED(ctx, "snpg_rep_moments_device");
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
    TEAM_UNSUPPORTED(ctx, "snpg_between_all");
    NEED(ctx, groups && n_groups >= 2 && n_groups <= SNPG_MAX_GROUPS && pair_sums, SNPG_E_ARG, "bad between_all arguments");
    NEED(ctx, min_taxa_total >= 0 && min_taxa_group >= 0 && min_taxa_total < ((int64_t)1 << 32) && min_taxa_group < ((int64_t)1 << 32),
         SNPG_E_ARG, "bad min-taxa criteria");
    NEED(ctx, B < ((int64_t)1 << 36), SNPG_E_ARG, "too many replicates");
    for (int i = 0; i < n_groups; i++)
        if (int rc = check_group(ctx, groups[i])) return rc;
    if (int rc = bind(ctx)) return rc;
    const int P = ctx->n_products;
    const int64_t cnt = ctx->shard_count;
    const int n_pairs = n_groups * (n_groups - 1) / 2;
    const bool boot = pair_se && B >= 2;
    ctx->stats_of = nullptr;
    // The pairs are independent chains of small, latency-bound kernels (cross-term statistics, filter, classes, binomial
    // split, bootstrap, moments: 80 us a pair on config 4, 1.3 % of the SM time busy): up to kPairLanes of them run side by
    // side, each on its own stream with its own slice of the scratch.  Same launches, same Philox streams, same results.
    // With per-kernel timing on (SNPG_OPT_PROFILE) they stay on the ctx stream, one after the other.
    const int lanes = (ctx->profile_mask || n_pairs < 2) ? 1 : std::min<int>(snpg_ctx::kPairLanes, n_pairs);
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t stats_b = up(sizeof(double) * 4 * (size_t)std::max<int64_t>(1, cnt)), nd_b = up(sizeof(uint32_t) * (size_t)std::max<int64_t>(1, cnt));
    const size_t vals_b = boot ? up(sizeof(double) * 6 * (size_t)(B * P)) : 0;
    const size_t boot_b = boot && ctx->boot_kernel == SNPG_BOOT_CLASSES ? up(boot_scratch_bytes(cnt, P, B)) : 0;
    CU(ctx, ctx->s_stats.ensure(stats_b * lanes));
    CU(ctx, ctx->s_nda.ensure(nd_b * lanes));
    CU(ctx, ctx->s_ndb.ensure(nd_b * lanes));
    CU(ctx, ctx->s_sums.ensure(sizeof(double) * 4 * (size_t)(P * n_pairs)));
    if (boot) {
        CU(ctx, ctx->s_vals.ensure(vals_b * lanes));
        CU(ctx, ctx->s_se.ensure(sizeof(double) * 6 * (size_t)(P * n_pairs)));
        if (boot_b) CU(ctx, ctx->s_boot.ensure(boot_b * lanes));
    }
    if (lanes > 1) {
        for (int l = 0; l < lanes; l++) {
            if (!ctx->pair_stream[l]) CU(ctx, cudaStreamCreateWithFlags(&ctx->pair_stream[l], cudaStreamNonBlocking));
            if (!ctx->pair_done[l]) CU(ctx, cudaEventCreateWithFlags(&ctx->pair_done[l], cudaEventDisableTiming));
        }
        CU(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));          // the groups' histograms are ready on the ctx stream
        for (int l = 0; l < lanes; l++) CU(ctx, cudaStreamWaitEvent(ctx->pair_stream[l], ctx->ev_fork, 0));
    }
    const uint32_t call = ctx->call_id++;
    int pair = 0;
    for (int i = 0; i < n_groups; i++) {
        for (int j = i + 1; j < n_groups; j++, pair++) {
            const Group &ga = ctx->groups[groups[i]], &gb = ctx->groups[groups[j]];
            const int lane = pair % lanes;
            cudaStream_t st = lanes > 1 ? ctx->pair_stream[lane] : ctx->stream;
            double *l_stats = reinterpret_cast<double *>(static_cast<uint8_t *>(ctx->s_stats.p) + stats_b * lane);
            uint32_t *l_nda = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(ctx->s_nda.p) + nd_b * lane);
            uint32_t *l_ndb = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(ctx->s_ndb.p) + nd_b * lane);
            StatsOut o{nullptr, l_nda, l_ndb, nullptr, l_stats, nullptr};
            { Launch l(ctx, SNPG_K_STATS); launch_between_stats(view_of(ctx, ga, 0), view_of(ctx, gb, 0), cnt, ctx->d_tables.as<double>(), o, st); }
            if (min_taxa_total > 0 || min_taxa_group > 0) {
                Launch l(ctx, SNPG_K_OTHER);
                launch_taxa_filter(l_stats, l_nda, l_ndb, cnt, (uint32_t)min_taxa_total, (uint32_t)min_taxa_group, st);
            }
            double *d_pair_sums = ctx->s_sums.as<double>() + 4 * (size_t)P * pair;
            if (!boot) { Launch l(ctx, SNPG_K_SUMS); launch_product_sums(l_stats, ctx->d_local_ofs.as<int64_t>(), P, d_pair_sums, st); }
            if (boot) {
                double *l_vals = reinterpret_cast<double *>(static_cast<uint8_t *>(ctx->s_vals.p) + vals_b * lane);
                {
                    Launch l(ctx, SNPG_K_BOOT);      // also writes the product sums
                    BootArgs a;
                    a.stats4 = l_stats; a.beg = ctx->d_local_ofs.as<int64_t>(); a.end = a.beg + 1; a.order = ctx->d_order_local.as<int>();
                    a.n_items = P; a.slot0 = ctx->boot_slot0; a.b_count = B; a.key = philox_key(ctx); a.kind = kBootBetween; a.call = call;
                    a.item0 = (uint32_t)(pair * P);  // every (pair, product) its own stream
                    a.vals = l_vals; a.prod_sums = d_pair_sums;
                    if (boot_b) {
                        boot_scratch_carve(static_cast<uint8_t *>(ctx->s_boot.p) + boot_b * lane, cnt, P, B, a);
                        ctx->launches += 2;          // k_boot_classify, k_boot_split
                    }
                    launch_bootstrap(ctx->boot_kernel, a, ctx->max_codons_local, st);
                }
                { Launch l(ctx, SNPG_K_MOMENTS); launch_moments_vals(l_vals, P, B, ctx->s_se.as<double>() + 6 * (size_t)P * pair, nullptr, nullptr, st); }
            }
        }
    }
    if (lanes > 1)
        for (int l = 0; l < lanes; l++) {
            CU(ctx, cudaEventRecord(ctx->pair_done[l], ctx->pair_stream[l]));
            CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->pair_done[l], 0));
        }
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(pair_sums, ctx->s_sums.p, sizeof(double) * 4 * (size_t)(P * n_pairs), cudaMemcpyDeviceToHost, ctx->stream));
    if (boot) CU(ctx, cudaMemcpyAsync(pair_se, ctx->s_se.p, sizeof(double) * 6 * (size_t)(P * n_pairs), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, sync_stream(ctx));
    return SNPG_OK;
}

int snpg_between_job(snpg_ctx *ctx, int n_groups, const uint8_t *const *bases, int bases_on_device, const uint64_t *n_seqs, uint64_t aln_len,
                     const uint64_t *row_stride, int64_t min_taxa_total, int64_t min_taxa_group, int64_t B, double *pair_sums, double *pair_se) {
    if (!ctx) return SNPG_E_ARG;
    TEAM_UNSUPPORTED(ctx, "snpg_between_job");
    NEED(ctx, bases && n_seqs && row_stride && n_groups >= 2 && n_groups <= SNPG_MAX_GROUPS, SNPG_E_ARG, "bad between_job arguments");
    int groups[SNPG_MAX_GROUPS];
    int rc = SNPG_OK;
    ctx->load_no_sync = true;
    for (int g = 0; g < n_groups && rc == SNPG_OK; g++) {
        groups[g] = g;
        rc = bases_on_device ? snpg_load_group_device(ctx, g, bases[g], n_seqs[g], aln_len, row_stride[g])
                             : snpg_load_group(ctx, g, bases[g], n_seqs[g], aln_len, row_stride[g]);
    }
    ctx->load_no_sync = false;
    if (rc != SNPG_OK) { sync_stream(ctx); return rc; }     // nothing stays queued behind a failed call
    return snpg_between_all(ctx, groups, n_groups, min_taxa_total, min_taxa_group, B, pair_sums, pair_se);
}

int snpg_test_philox(snpg_ctx *ctx, const uint32_t *ctr4, const uint32_t *key2, uint32_t *out4) {
    if (!ctx || !ctr4 || !key2 || !out4) return SNPG_E_ARG;
    TEAM_FORWARD0(ctx, snpg_test_philox(c, ctr4, key2, out4));
    if (int rc = bind(ctx)) return rc;
    CU(ctx, ctx->s_misc.ensure(16));
    { Launch l(ctx, SNPG_K_OTHER); launch_philox_kat(ctr4, key2, ctx->s_misc.as<uint32_t>(), ctx->stream); }
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(out4, ctx->s_misc.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, sync_stream(ctx));
    return SNPG_OK;
}

int snpg_staged_columns(const snpg_ctx *ctx, int64_t *first_col, int64_t *n_cols) {
    if (!ctx || is_team(ctx) || ctx->n_products == 0) return SNPG_E_STATE;
    if (first_col) *first_col = ctx->base_al;
    if (n_cols) *n_cols = ctx->shard_count > 0 ? ctx->col_hi - ctx->base_al : 0;
    return SNPG_OK;
}

int64_t snpg_launch_count(const snpg_ctx *ctx) {
    if (!ctx) return 0;
    int64_t n = ctx->launches;
    for (const snpg_ctx *c : ctx->team) n += c->launches;
    return n;
}
int64_t snpg_tile_launch_count(const snpg_ctx *ctx) {
    if (!ctx) return 0;
    int64_t n = ctx->tile_launches;
    for (const snpg_ctx *c : ctx->team) n += c->tile_launches;
    return n;
}

int64_t snpg_span_tma_launch_count(const snpg_ctx *ctx) {
    if (!ctx) return 0;
    int64_t n = ctx->span_tma_launches;
    for (const snpg_ctx *c : ctx->team) n += c->span_tma_launches;
    return n;
}

int snpg_set_stream(snpg_ctx *ctx, void *cuda_stream) {
    if (!ctx) return SNPG_E_ARG;
    TEAM_UNSUPPORTED(ctx, "snpg_set_stream");
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
    TEAM_FORWARD0(ctx, snpg_get_profile(c, ms, calls, reset));
    for (int k = 0; k < SNPG_N_KERNEL_KINDS; k++) {
        if (ms) ms[k] = ctx->kernel_ms[k];
        if (calls) calls[k] = ctx->kernel_calls[k];
        if (reset) { ctx->kernel_ms[k] = 0; ctx->kernel_calls[k] = 0; }
    }
    return SNPG_OK;
}

}  // extern "C"

// Whole jobs behind the C ABI: load -> per-codon statistics -> product sums -> whole-product bootstrap -> moments ->
// sliding windows -> window bootstraps, enqueued without a host round trip and, on a device-resident alignment,
// replayed as a CUDA graph.  On several GPUs (exchange.cu) the same sequence runs on every rank over its codon range
// and its replicate range; the kernels exchange through peer memory (snpg_peer.cuh), so the launch list is the same.
#include <algorithm>
#include <cstdlib>

#include "snpg_ctx.h"
#include "snpg_team.h"

using namespace snpg;

namespace {

bool exchanging(const snpg_ctx *ctx) { return ctx->world > 1 && ctx->xchg.ready; }

uint2 philox_key(const snpg_ctx *ctx) { return make_uint2((uint32_t)ctx->seed, (uint32_t)(ctx->seed >> 32)); }

bool is_pinned(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// The job's codon axis: every codon of the annotation on one GPU or with a connected exchange; the local slice for a
// rank of a caller-managed shard (snpg_set_shard without snpg_peer_connect).
int64_t axis_codons(const snpg_ctx *ctx) { return exchanging(ctx) ? ctx->full_ofs[ctx->n_products] : ctx->shard_count; }

// sliding windows of every product over the job's codon axis (wgp:293-312: a product shorter than the window is skipped)
int plan_windows(snpg_ctx *ctx, int64_t W, int64_t step) {
    auto &wp = ctx->wplan;
    if (wp.W == W && wp.step == step && !wp.ofs.empty()) return SNPG_OK;
    NEED(ctx, ctx->world == 1 || exchanging(ctx), SNPG_E_STATE, "sliding windows need the whole codon axis: connect the ranks (snpg_peer_connect)");
    const int P = ctx->n_products;
    std::vector<int64_t> ofs((size_t)P + 1, 0), be;
    std::vector<int64_t> beg, end;
    for (int p = 0; p < P; p++) {
        const int64_t c0 = ctx->full_ofs[p], C = ctx->full_ofs[p + 1] - c0;
        const int64_t nw = C >= W ? (C - W) / step + 1 : 0;
        for (int64_t k = 0; k < nw; k++) { beg.push_back(c0 + k * step); end.push_back(c0 + k * step + W); }
        ofs[(size_t)p + 1] = ofs[(size_t)p] + nw;
    }
    const int64_t n = ofs[(size_t)P];
    NEED(ctx, n < ((int64_t)1 << 24), SNPG_E_ARG, "too many windows (%lld)", (long long)n);
    if (n) {
        CU(ctx, wp.beg.ensure(sizeof(int64_t) * (size_t)n));
        CU(ctx, wp.end.ensure(sizeof(int64_t) * (size_t)n));
        CU(ctx, cudaMemcpy(wp.beg.p, beg.data(), sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice));
        CU(ctx, cudaMemcpy(wp.end.p, end.data(), sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice));
    }
    wp.W = W; wp.step = step; wp.n = n; wp.ofs = ofs;
    return SNPG_OK;
}

int validate_job(snpg_ctx *ctx, const snpg_job &job) {
    NEED(ctx, job.struct_size == sizeof(snpg_job), SNPG_E_ARG, "snpg_job.struct_size is %u, this library expects %zu", job.struct_size, sizeof(snpg_job));
    NEED(ctx, ctx->n_products > 0, SNPG_E_STATE, "no products set");
    NEED(ctx, job.n_bootstraps >= 0 && job.n_bootstraps < ((int64_t)1 << 36) && job.n_window_bootstraps >= 0 && job.n_window_bootstraps < ((int64_t)1 << 36),
         SNPG_E_ARG, "bad replicate counts");
    NEED(ctx, job.window >= 0 && (job.window == 0 || job.step >= 1) && job.window < ((int64_t)1 << 31), SNPG_E_ARG, "bad window arguments");
    NEED(ctx, ctx->world == 1 || exchanging(ctx) || job.window == 0, SNPG_E_STATE, "sliding windows need connected ranks");
    if (exchanging(ctx)) NEED(ctx, job.n_bootstraps <= ctx->xchg.max_B, SNPG_E_ARG, "%lld bootstrap replicates, but the exchange was exported for at most %lld",
                              (long long)job.n_bootstraps, (long long)ctx->xchg.max_B);
    return SNPG_OK;
}

// Every buffer the job will touch, sized before its first launch: nothing is allocated between launches (inside one
// process a cudaMalloc on one GPU must never wait behind a kernel that waits for a peer not launched yet), and a
// captured graph needs fixed addresses anyway.
int job_prepare(snpg_ctx *ctx, const snpg_job &job) {
    const int P = ctx->n_products;
    const int64_t cnt = std::max<int64_t>(1, ctx->shard_count), axis = std::max<int64_t>(1, axis_codons(ctx));
    const int64_t B = job.n_bootstraps;
    if (!exchanging(ctx)) {
        CU(ctx, ctx->s_stats.ensure(sizeof(double) * 4 * (size_t)cnt));
        CU(ctx, ctx->s_poly.ensure((size_t)cnt));
        CU(ctx, ctx->s_cons.ensure((size_t)cnt));
        CU(ctx, ctx->s_nda.ensure(sizeof(uint32_t) * (size_t)cnt));
        CU(ctx, ctx->s_cmp.ensure(sizeof(uint64_t) * (size_t)cnt));
        if (B >= 2) CU(ctx, ctx->s_vals.ensure(sizeof(double) * 6 * (size_t)(B * P)));
    }
    if (B >= 2 && ctx->boot_kernel == SNPG_BOOT_CLASSES) CU(ctx, ctx->s_boot.ensure(boot_scratch_bytes(axis, P, B)));
    // a job's product-level results sit next to each other -- sums [P][4] | SEs [P][6] | undefined counts [P][6] -- and leave
    // with one copy (three copy nodes at the end of the step's critical path cost about 4 us more)
    CU(ctx, ctx->s_sums.ensure((sizeof(double) * 10 + sizeof(uint32_t) * 6) * (size_t)P));
    CU(ctx, ctx->d_salt.ensure(2 * sizeof(uint32_t)));
    if (!ctx->h_salt) CU(ctx, cudaHostAlloc(reinterpret_cast<void **>(&ctx->h_salt), 2 * sizeof(uint32_t), cudaHostAllocPortable));
    size_t stage = (size_t)P * (4 * 8 + 6 * 8 + 6 * 4) + 64;
    if (job.window > 0) {
        if (int rc = plan_windows(ctx, job.window, job.step)) return rc;
        const int64_t nw = std::max<int64_t>(1, ctx->wplan.n);
        CU(ctx, ctx->w_sums.ensure(sizeof(double) * 4 * (size_t)nw));
        if (job.n_window_bootstraps >= 2) {
            CU(ctx, ctx->w_vals.ensure(sizeof(double) * 6 * (size_t)(nw * job.n_window_bootstraps)));
            CU(ctx, ctx->w_se.ensure(sizeof(double) * 6 * (size_t)nw));
        }
        stage += (size_t)nw * (4 * 8 + 6 * 8) + 64;
    }
    // staging for the outputs that are not in pinned memory (and always for the product-level ones)
    if (job.poly && !is_pinned(job.poly)) stage += (size_t)axis + 16;
    if (job.conserved && !is_pinned(job.conserved)) stage += (size_t)axis + 16;
    if (job.n_defined && !is_pinned(job.n_defined)) stage += 4 * (size_t)axis + 16;
    if (job.comparisons && !is_pinned(job.comparisons)) stage += 8 * (size_t)axis + 16;
    if (job.stats4 && !is_pinned(job.stats4)) stage += 32 * (size_t)axis + 16;
    if (stage > ctx->h_out_cap) {
        if (ctx->h_out) cudaFreeHost(ctx->h_out);
        ctx->h_out = nullptr; ctx->h_out_cap = 0;
        g_alloc_epoch++;
        CU(ctx, cudaHostAlloc(reinterpret_cast<void **>(&ctx->h_out), stage, cudaHostAllocPortable));
        ctx->h_out_cap = stage;
    }
    return SNPG_OK;
}

// The outputs of a snpg_job, by number (a replayed graph copies into the CURRENT call's arrays).
enum Field { F_PROD_SUMS, F_PROD_SE, F_PROD_UNDEF, F_POLY, F_NDEF, F_CMP, F_STATS, F_CONS, F_WIN_SUMS, F_WIN_SE, F_COUNT };
void *field_ptr(const snpg_job &j, int f) {
    switch (f) {
    case F_PROD_SUMS: return j.prod_sums; case F_PROD_SE: return j.prod_se; case F_PROD_UNDEF: return j.prod_undefined;
    case F_POLY: return j.poly; case F_NDEF: return j.n_defined; case F_CMP: return j.comparisons; case F_STATS: return j.stats4;
    case F_CONS: return j.conserved; case F_WIN_SUMS: return j.win_sums; case F_WIN_SE: return j.win_se;
    }
    return nullptr;
}
bool staged_field(int f) { return f == F_PROD_SUMS || f == F_PROD_SE || f == F_PROD_UNDEF || f == F_WIN_SUMS || f == F_WIN_SE; }

// What identifies a job for graph replay: sizes and options, and of the outputs only what the captured copies depend
// on -- whether an output is wanted, and the address of those written directly (pinned per-codon tables).  Everything
// else goes through the ctx's staging buffer and may live anywhere from call to call.
snpg_job job_signature(const snpg_job &job) {
    snpg_job k = job;
    void **fields[] = {(void **)&k.prod_sums, (void **)&k.prod_se, (void **)&k.prod_undefined, (void **)&k.poly, (void **)&k.n_defined,
                       (void **)&k.comparisons, (void **)&k.stats4, (void **)&k.conserved, (void **)&k.win_sums, (void **)&k.win_se};
    for (int f = 0; f < F_COUNT; f++) {
        void *p = *fields[f];
        if (p && (staged_field(f) || !is_pinned(p))) *fields[f] = reinterpret_cast<void *>(1);
    }
    return k;
}

// device -> host copy of one output: straight into pinned memory, else through the ctx's pinned staging (the host
// copy to the caller's array follows the job's synchronisation)
struct Stager {
    snpg_ctx *ctx; const snpg_job *job; size_t used = 0;
    int put(int field, const void *src, size_t bytes, cudaStream_t stream) {
        void *dst = field_ptr(*job, field);
        if (!dst || bytes == 0) return SNPG_OK;
        if (!staged_field(field) && is_pinned(dst)) {
            CU(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream));
            return SNPG_OK;
        }
        const size_t at = (used + 15) & ~(size_t)15;
        NEED(ctx, at + bytes <= ctx->h_out_cap, SNPG_E_NOMEM, "output staging too small");
        CU(ctx, cudaMemcpyAsync(ctx->h_out + at, src, bytes, cudaMemcpyDeviceToHost, stream));
        ctx->pending.push_back({field, ctx->h_out + at, bytes});
        used = at + bytes;
        return SNPG_OK;
    }
};

// Enqueues the whole job on ctx->stream (and the result copies on ctx->copy_stream): no host synchronisation.  bases =
// null: the group is already loaded (snpg_within_all).  This is also what a CUDA graph capture records.
int job_record(snpg_ctx *ctx, int group, const uint8_t *bases, int on_device, uint64_t n_seqs, uint64_t aln_len, uint64_t row_stride,
               const snpg_job &job, int parity) {
    Group &g = ctx->groups[group];
    const int P = ctx->n_products;
    const int64_t cnt = ctx->shard_count;
    const bool multi = exchanging(ctx);
    const Exchange &X = ctx->xchg;
    const int64_t axis = axis_codons(ctx);
    ctx->pending.clear();
    // The job's call number and exchange epoch.  On one GPU only the bootstraps read them (Philox counters): the 8-byte copy
    // goes out on the copy stream beside the load's kernels instead of ahead of them (a copy node of its own at the head of
    // the replayed graph: about 2 us of the step's critical path); several GPUs need the epoch in the first kernel that
    // signals its peers.
    const bool salt_beside = !multi && !ctx->profile_mask;
    if (salt_beside) {
        CU(ctx, cudaEventRecord(ctx->ev_salt_fork, ctx->stream));
        CU(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_salt_fork, 0));
        CU(ctx, cudaMemcpyAsync(ctx->d_salt.p, ctx->h_salt, 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->copy_stream));
        CU(ctx, cudaEventRecord(ctx->ev_salt, ctx->copy_stream));
    } else {
        CU(ctx, cudaMemcpyAsync(ctx->d_salt.p, ctx->h_salt, 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    }
    const uint32_t *d_salt = ctx->d_salt.as<uint32_t>();

    // where the per-codon statistics go: local scratch, or every rank's exchange region (this job's half)
    PeerSync ps;
    PeerOut po;
    if (multi) {
        ps = peer_sync_of(ctx);
        ps.epoch = d_salt + 1;
        po.world = ctx->world;
        po.first = ctx->shard_first;
        for (int r = 0; r < ctx->world; r++) {
            po.stats4[r] = X.at<double>(r, parity, X.off_stats);
            po.cmp[r] = X.at<unsigned long long>(r, parity, X.off_cmp);
            po.ndef[r] = X.at<uint32_t>(r, parity, X.off_ndef);
            po.poly[r] = X.at<uint8_t>(r, parity, X.off_poly);
            po.cons[r] = X.at<uint8_t>(r, parity, X.off_cons);
        }
        ctx->job_out = StatsOut{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    } else {
        ctx->job_out = StatsOut{ctx->s_poly.as<uint8_t>(), ctx->s_nda.as<uint32_t>(), nullptr, ctx->s_cmp.as<unsigned long long>(),
                                ctx->s_stats.as<double>(), ctx->s_cons.as<uint8_t>()};
    }
    ctx->job_peer_out = po;
    ctx->job_peer_sync = ps;

    if (bases) {
        ctx->defer_load_sync = true;     // the load's kernels stay queued; its fix-up kernel also takes the statistics
        const int rc = on_device ? snpg_load_group_device(ctx, group, bases, n_seqs, aln_len, row_stride)
                                 : snpg_load_group(ctx, group, bases, n_seqs, aln_len, row_stride);
        ctx->defer_load_sync = false;
        if (rc != SNPG_OK) { ctx->stats_of = nullptr; return rc; }
    }
    if (ctx->stats_of != &g) {
        Launch l(ctx, SNPG_K_STATS);
        launch_within_stats(view_of(ctx, g, 0), cnt, ctx->d_tables.as<double>(), ctx->job_out, &po, &ps, ctx->stream);
    }
    ctx->stats_of = nullptr;

    const double *stats = multi ? X.at<double>(ctx->rank, parity, X.off_stats) : ctx->s_stats.as<double>();
    const int64_t *ofs = multi ? ctx->d_full_ofs.as<int64_t>() : ctx->d_local_ofs.as<int64_t>();
    const int *order = multi ? ctx->d_order_full.as<int>() : ctx->d_order_local.as<int>();
    const int64_t max_codons = multi ? ctx->max_codons_full : ctx->max_codons_local;
    Stager st{ctx, &job};

    // the per-codon tables leave on the copy stream while the bootstrap runs; on one GPU the sliding windows (sums,
    // bootstraps, moments: they read the statistics only) run on a third stream beside the product bootstraps -- config 3:
    // two chains of about 75 us each.  With per-kernel timing on, everything stays on the ctx stream.
    const bool codon_out = job.poly || job.n_defined || job.comparisons || job.stats4 || job.conserved;
    const int64_t nwin = job.window > 0 ? ctx->wplan.n : 0;
    const bool win_boot = nwin > 0 && job.n_window_bootstraps >= 2 && job.win_se;
    const bool win_beside = nwin > 0 && (job.win_sums || win_boot) && job.n_bootstraps >= 2 && !multi && !ctx->profile_mask;
    cudaStream_t ws = win_beside ? ctx->win_stream : ctx->stream;
    if (codon_out || win_beside) CU(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
    if (win_beside) CU(ctx, cudaStreamWaitEvent(ctx->win_stream, ctx->ev_fork, 0));
    if (codon_out) {
        CU(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_fork, 0));
        if (multi) launch_peer_wait(ps, kStageStats, ctx->copy_stream);
        const uint8_t *d_poly = multi ? X.at<uint8_t>(ctx->rank, parity, X.off_poly) : ctx->s_poly.as<uint8_t>();
        const uint8_t *d_cons = multi ? X.at<uint8_t>(ctx->rank, parity, X.off_cons) : ctx->s_cons.as<uint8_t>();
        const uint32_t *d_ndef = multi ? X.at<uint32_t>(ctx->rank, parity, X.off_ndef) : ctx->s_nda.as<uint32_t>();
        const uint64_t *d_cmp = multi ? X.at<uint64_t>(ctx->rank, parity, X.off_cmp) : ctx->s_cmp.as<uint64_t>();
        if (int rc = st.put(F_STATS, stats, sizeof(double) * 4 * (size_t)axis, ctx->copy_stream)) return rc;
        if (int rc = st.put(F_CMP, d_cmp, sizeof(uint64_t) * (size_t)axis, ctx->copy_stream)) return rc;
        if (int rc = st.put(F_NDEF, d_ndef, sizeof(uint32_t) * (size_t)axis, ctx->copy_stream)) return rc;
        if (int rc = st.put(F_POLY, d_poly, (size_t)axis, ctx->copy_stream)) return rc;
        if (int rc = st.put(F_CONS, d_cons, (size_t)axis, ctx->copy_stream)) return rc;
        CU(ctx, cudaEventRecord(ctx->ev_join, ctx->copy_stream));
    }

    // sliding windows over the whole axis (every rank computes all of them: 4 W adds and B_win W draws per window)
    auto enqueue_windows = [&](bool stats_arrived) -> int {
        if (nwin > 0 && (job.win_sums || win_boot)) {
            if (!stats_arrived) { Launch l(ctx, SNPG_K_OTHER); launch_peer_wait(ps, kStageStats, ws); }
            if (job.win_sums) {
                Launch l(ctx, SNPG_K_WINDOW);
                launch_window_sums(stats, ctx->wplan.beg.as<int64_t>(), ctx->wplan.end.as<int64_t>(), nwin, ctx->w_sums.as<double>(), ws);
            }
            if (win_boot) {
                if (salt_beside) CU(ctx, cudaStreamWaitEvent(ws, ctx->ev_salt, 0));
                {
                    Launch l(ctx, SNPG_K_WINDOW);
                    BootArgs a;
                    a.stats4 = stats; a.beg = ctx->wplan.beg.as<int64_t>(); a.end = ctx->wplan.end.as<int64_t>(); a.n_items = (int)nwin; a.slot0 = 0;
                    a.b_count = job.n_window_bootstraps; a.key = philox_key(ctx); a.kind = kBootJobWindow; a.call = 0; a.salt = d_salt;
                    a.vals = ctx->w_vals.as<double>();
                    a.n_which = 3;                  // win_se holds the SEs of dN, dS, dN-dS (wgp:406-465): no Jukes-Cantor logarithms
                    launch_bootstrap(ctx->boot_kernel, a, job.window, ws);
                }
                { Launch l(ctx, SNPG_K_MOMENTS); launch_moments_vals(ctx->w_vals.as<double>(), nwin, job.n_window_bootstraps, ctx->w_se.as<double>(), nullptr, nullptr, ws, 3); }
            }
        }
        if (win_beside) CU(ctx, cudaEventRecord(ctx->ev_win, ctx->win_stream));
        return SNPG_OK;
    };
    if (win_beside)                       // on their own stream, enqueued first: the product bootstraps below run beside them
        if (int rc = enqueue_windows(true)) return rc;
    double *r_sums = ctx->s_sums.as<double>(), *r_se = r_sums + 4 * (size_t)P;
    uint32_t *r_undef = reinterpret_cast<uint32_t *>(r_se + 6 * (size_t)P);
    const int64_t B = job.n_bootstraps;
    const bool boot = B >= 2;             // (a rank's replicate range is needed by its peers whether or not it reports anything itself)
    bool stats_arrived = !multi;          // every rank's statistics are known to be in this rank's region
    if (!boot) {
        if (job.prod_sums || job.window > 0) {
            if (multi) { Launch l(ctx, SNPG_K_OTHER); launch_peer_wait(ps, kStageStats, ctx->stream); }
            stats_arrived = true;
        }
        if (job.prod_sums) { Launch l(ctx, SNPG_K_SUMS); launch_product_sums(stats, ofs, P, r_sums, ctx->stream); }
    } else {
        int64_t b_first = 0, b_count = B;
        if (multi) snpg_shard_plan(B, ctx->rank, ctx->world, &b_first, &b_count);
        // (ranks that share a GPU: a full grid of waiting blocks could keep the peer's producer off the SMs)
        if (multi && X.peer_on_my_gpu) { Launch l(ctx, SNPG_K_OTHER); launch_peer_wait(ps, kStageStats, ctx->stream); }
        if (salt_beside) CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_salt, 0));
        {
            Launch l(ctx, SNPG_K_BOOT);
            BootArgs a;
            a.stats4 = stats; a.beg = ofs; a.end = ofs + 1; a.order = order; a.n_items = P; a.slot0 = ctx->boot_slot0;
            a.b_first = b_first; a.b_count = b_count; a.vals_stride = B; a.vals_b0 = b_first;
            a.key = philox_key(ctx); a.kind = kBootJob; a.call = 0; a.salt = d_salt;
            a.prod_sums = r_sums;                                   // also writes the product sums
            if (multi) {
                a.pv.world = ctx->world;
                for (int r = 0; r < ctx->world; r++) a.pv.vals[r] = X.at<double>(r, parity, X.off_vals);
                a.ps = ps;
            } else {
                a.vals = ctx->s_vals.as<double>();
            }
            if (int rc = boot_classes_scratch(ctx, a, axis_codons(ctx))) return rc;
            launch_bootstrap(ctx->boot_kernel, a, max_codons, ctx->stream);
        }
        stats_arrived = true;             // every block of the bootstrap waited for them
        if (multi && X.peer_on_my_gpu) { Launch l(ctx, SNPG_K_OTHER); launch_peer_wait(ps, kStageVals, ctx->stream); }
        {
            Launch l(ctx, SNPG_K_MOMENTS);
            const double *vals = multi ? X.at<double>(ctx->rank, parity, X.off_vals) : ctx->s_vals.as<double>();
            launch_moments_vals(vals, P, B, r_se, r_undef, multi ? &ps : nullptr, ctx->stream);
        }
    }
    if (win_beside) CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_win, 0));
    else if (int rc = enqueue_windows(stats_arrived)) return rc;
    if (salt_beside) CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_salt, 0));      // (a job without bootstraps: the copy still joins the stream)
    CU(ctx, cudaGetLastError());
    // the small tables always go through the staging buffer (one pinned line each)
    {
        const size_t b_sums = sizeof(double) * 4 * (size_t)P, b_se = boot ? sizeof(double) * 6 * (size_t)P : 0, b_undef = boot ? sizeof(uint32_t) * 6 * (size_t)P : 0;
        const size_t at = (st.used + 15) & ~(size_t)15;
        NEED(ctx, at + b_sums + b_se + b_undef <= ctx->h_out_cap, SNPG_E_NOMEM, "output staging too small");
        CU(ctx, cudaMemcpyAsync(ctx->h_out + at, r_sums, b_sums + b_se + b_undef, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->pending.push_back({F_PROD_SUMS, ctx->h_out + at, b_sums});
        if (boot) {
            ctx->pending.push_back({F_PROD_SE, ctx->h_out + at + b_sums, b_se});
            ctx->pending.push_back({F_PROD_UNDEF, ctx->h_out + at + b_sums + b_se, b_undef});
        }
        st.used = at + b_sums + b_se + b_undef;
    }
    if (nwin > 0) {
        if (int rc = st.put(F_WIN_SUMS, ctx->w_sums.p, sizeof(double) * 4 * (size_t)nwin, ctx->stream)) return rc;
        // win_se is [n][3]; the device holds [n][6]: the staged copy is compacted on the host (apply_pending)
        if (win_boot)
            if (int rc = st.put(F_WIN_SE, ctx->w_se.p, sizeof(double) * 6 * (size_t)nwin, ctx->stream)) return rc;
    }
    if (codon_out) CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    return SNPG_OK;
}

void apply_pending(snpg_ctx *ctx, const snpg_job &job) {
    for (const auto &c : ctx->pending) {
        void *dst = field_ptr(job, c.field);
        if (!dst) continue;
        if (c.field != F_WIN_SE) { memcpy(dst, c.src, c.bytes); continue; }
        const size_t n = c.bytes / (6 * sizeof(double));                    // window SEs: the first three of six per window
        const double *src = static_cast<const double *>(c.src);
        double *d = static_cast<double *>(dst);
        for (size_t k = 0; k < n; k++)
            for (int q = 0; q < 3; q++) d[3 * k + q] = src[6 * k + q];
    }
    ctx->pending.clear();
}

// Captures the job into a CUDA graph.  Any failure leaves no graph behind and reports SNPG_E_STATE so that the
// caller runs the job the ordinary way.
int capture_job_graph(snpg_ctx *ctx, snpg_ctx::JobGraph &J, const snpg_ctx::JobKey &key, const snpg_job &job) {
    const uint64_t epoch0 = g_alloc_epoch;
    const int64_t l0 = ctx->launches, t0 = ctx->tile_launches, t1 = ctx->span_tma_launches;
    if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); return SNPG_E_STATE; }
    ctx->capturing = true;
    ctx->capturing_into = &J;
    int rc = job_record(ctx, key.group, key.bases, 1, key.n, key.len, key.stride, job, key.parity);
    ctx->capturing = false;
    ctx->capturing_into = nullptr;
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
    cudaGraphExec_t exec = nullptr;
    if (rc == SNPG_OK && e == cudaSuccess && graph && epoch0 == g_alloc_epoch && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
        J.exec = exec;
        J.key = key;
        J.epoch = g_alloc_epoch;
        J.n_launches = ctx->launches - l0;
        J.n_tile_launches = ctx->tile_launches - t0;
        J.n_span_tma_launches = ctx->span_tma_launches - t1;
        J.pending = ctx->pending;
    } else {
        rc = SNPG_E_STATE;
        cudaGetLastError();
    }
    if (graph) cudaGraphDestroy(graph);
    ctx->pending.clear();
    ctx->launches = l0;                 // nothing ran yet: the replay counts them
    ctx->tile_launches = t0;
    ctx->span_tma_launches = t1;
    return rc;
}

void drop_graph(snpg_ctx *ctx, snpg_ctx::JobGraph &J) {
    if (J.exec) cudaGraphExecDestroy(J.exec);
    J.exec = nullptr;
    for (auto &sp : J.spans) { ctx->free_events.push_back(sp.e0); ctx->free_events.push_back(sp.e1); }
    J.spans.clear();
    J.pending.clear();
}

// Launches the job without waiting for it.  *replayed (or null) receives the graph that ran, if one did.
int job_enqueue(snpg_ctx *ctx, int group, const uint8_t *bases, int on_device, uint64_t n_seqs, uint64_t aln_len, uint64_t row_stride,
                const snpg_job &job, snpg_ctx::JobGraph **replayed) {
    if (replayed) *replayed = nullptr;
    const bool multi = exchanging(ctx);
    ctx->h_salt[0] = ctx->call_id++;
    ctx->h_salt[1] = multi ? ++ctx->xchg.epoch : 0;
    const int parity = multi ? (int)(ctx->xchg.epoch & 1u) : 0;
    static const bool graphs_on = getenv("SNPG_NO_GRAPH") == nullptr;
    snpg_ctx::JobKey key;
    key.profile_mask = ctx->profile_mask; key.group = group; key.bases = bases; key.n = n_seqs; key.len = aln_len; key.stride = row_stride;
    key.keep_codes = ctx->keep_codes; key.site_counts = ctx->site_counts; key.fused_kernel = ctx->fused_kernel; key.boot_kernel = ctx->boot_kernel;
    key.boot_slot0 = ctx->boot_slot0; key.parity = parity; key.job = job_signature(job);
    const bool eligible = graphs_on && bases && on_device && !ctx->job_graphs_disabled;
    if (eligible) {
        auto &J = ctx->job_graph[parity];
        if (J.exec && (!(J.key == key) || J.epoch != g_alloc_epoch)) drop_graph(ctx, J);
        snpg_ctx::JobKey cand = key;
        cand.parity = 0;
        if (!J.exec && ctx->job_candidate == cand && ctx->job_candidate_epoch == g_alloc_epoch) {
            // the same job again with every buffer already in place: capture it
            if (capture_job_graph(ctx, J, key, job) != SNPG_OK) { drop_graph(ctx, J); ctx->job_graphs_disabled = true; }
        }
        if (J.exec) {
            CU(ctx, cudaGraphLaunch(J.exec, ctx->stream));
            ctx->launches += J.n_launches;
            ctx->tile_launches += J.n_tile_launches;
            ctx->span_tma_launches += J.n_span_tma_launches;
            ctx->graph_replays++;
            ctx->pending = J.pending;
            // the load inside the graph bypassed the host-side bookkeeping of the group
            Group &g = ctx->groups[group];
            g.n = n_seqs; g.n_pad = (int64_t)((n_seqs + kRowPad - 1) / kRowPad) * kRowPad;
            g.loaded = true; g.has_sites = ctx->site_counts; g.clear_in_vote = false;
            if (replayed) *replayed = &J;
            return SNPG_OK;
        }
    }
    const int rc = job_record(ctx, group, bases, on_device, n_seqs, aln_len, row_stride, job, parity);
    if (rc == SNPG_OK && eligible) {
        ctx->job_candidate = key;
        ctx->job_candidate.parity = 0;
        ctx->job_candidate_epoch = g_alloc_epoch;
    }
    return rc;
}

int job_finish(snpg_ctx *ctx, snpg_ctx::JobGraph *replayed, const snpg_job &job) {
    CU(ctx, sync_stream(ctx));
    if (replayed)
        for (auto &sp : replayed->spans) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, sp.e0, sp.e1) == cudaSuccess) { ctx->kernel_ms[sp.kind] += ms; ctx->kernel_calls[sp.kind]++; }
        }
    apply_pending(ctx, job);
    return exchange_check_error(ctx);
}

int check_job_args(snpg_ctx *ctx, int group, const uint8_t *bases, const snpg_job *job) {
    NEED(ctx, job, SNPG_E_ARG, "job is NULL");
    NEED(ctx, group >= 0 && group < SNPG_MAX_GROUPS, SNPG_E_ARG, "group %d out of range", group);
    if (!bases)
        if (int rc = check_group(ctx, group)) return rc;
    return validate_job(ctx, *job);
}

// one job on every GPU of a team: all launched before any is waited for (their kernels wait for each other)
int team_job(snpg_ctx *team, int group, const uint8_t *bases, int on_device, uint64_t n_seqs, uint64_t aln_len, uint64_t row_stride,
             const snpg_job *job) {
    NEED(team, job && job->struct_size == sizeof(snpg_job), SNPG_E_ARG, "bad job");
    NEED(team, !(bases && on_device), SNPG_E_STATE, "a device pointer belongs to one GPU: a multi-GPU ctx loads from host memory");
    if (job->n_bootstraps > team->team[0]->xchg.max_B)
        if (int rc = team_wire(team, job->n_bootstraps)) return rc;
    snpg_job quiet = *job;              // the children behind the first compute their share and exchange it, but report nothing
    quiet.prod_sums = nullptr; quiet.prod_se = nullptr; quiet.prod_undefined = nullptr; quiet.poly = nullptr; quiet.n_defined = nullptr;
    quiet.comparisons = nullptr; quiet.stats4 = nullptr; quiet.conserved = nullptr; quiet.win_sums = nullptr; quiet.win_se = nullptr;
    quiet.window = 0; quiet.step = 0; quiet.n_window_bootstraps = 0;
    for (size_t i = 0; i < team->team.size(); i++) {
        snpg_ctx *c = team->team[i];
        if (int rc = check_job_args(c, group, bases, i == 0 ? job : &quiet)) return team_fail(team, c, rc);
        if (int rc = bind(c)) return team_fail(team, c, rc);
        if (int rc = job_prepare(c, i == 0 ? *job : quiet)) return team_fail(team, c, rc);
    }
    int first_rc = SNPG_OK;
    snpg_ctx *failed = nullptr;
    for (size_t i = 0; i < team->team.size(); i++) {
        snpg_ctx *c = team->team[i];
        bind(c);
        const int rc = job_enqueue(c, group, bases, on_device, n_seqs, aln_len, row_stride, i == 0 ? *job : quiet, nullptr);
        if (rc && !first_rc) { first_rc = rc; failed = c; }
    }
    for (snpg_ctx *c : team->team) {
        bind(c);
        const int rc = job_finish(c, nullptr, c == team->team[0] ? *job : quiet);
        if (rc && !first_rc) { first_rc = rc; failed = c; }
    }
    return first_rc ? team_fail(team, failed, first_rc) : SNPG_OK;
}

}  // namespace

extern "C" {

int snpg_window_plan(snpg_ctx *ctx, int64_t window, int64_t step, int64_t *win_ofs) {
    if (!ctx) return SNPG_E_ARG;
    TEAM_FORWARD0(ctx, snpg_window_plan(c, window, step, win_ofs));
    NEED(ctx, ctx->n_products > 0, SNPG_E_STATE, "no products set");
    NEED(ctx, window >= 1 && step >= 1 && window < ((int64_t)1 << 31), SNPG_E_ARG, "bad window arguments");
    if (int rc = bind(ctx)) return rc;
    if (int rc = plan_windows(ctx, window, step)) return rc;
    if (win_ofs)
        for (int p = 0; p <= ctx->n_products; p++) win_ofs[p] = ctx->wplan.ofs[(size_t)p];
    return SNPG_OK;
}

int snpg_within_job_ex(snpg_ctx *ctx, int group, const uint8_t *bases, int bases_on_device, uint64_t n_seqs, uint64_t aln_len,
                       uint64_t row_stride, const snpg_job *job) {
    if (!ctx) return SNPG_E_ARG;
    Trace trace(is_team(ctx) ? "snpg_within_job_ex(team)" : "snpg_within_job_ex");
    if (is_team(ctx)) return team_job(ctx, group, bases, bases_on_device, n_seqs, aln_len, row_stride, job);
    if (int rc = check_job_args(ctx, group, bases, job)) return rc;
    if (int rc = bind(ctx)) return rc;
    if (int rc = job_prepare(ctx, *job)) return rc;
    snpg_ctx::JobGraph *replayed = nullptr;
    const int rc = job_enqueue(ctx, group, bases, bases_on_device, n_seqs, aln_len, row_stride, *job, &replayed);
    const int rc2 = job_finish(ctx, replayed, *job);      // also after a failed enqueue: the stream must be drained
    return rc ? rc : rc2;
}

int snpg_within_job(snpg_ctx *ctx, int group, const uint8_t *bases, int bases_on_device, uint64_t n_seqs, uint64_t aln_len,
                    uint64_t row_stride, int64_t B, double *prod_sums, double *prod_se) {
    if (!ctx) return SNPG_E_ARG;
    NEED(ctx, bases, SNPG_E_ARG, "bases is NULL");
    snpg_job job = {};
    job.struct_size = sizeof job;
    job.n_bootstraps = prod_se ? B : 0;
    job.prod_sums = prod_sums;
    job.prod_se = prod_se;
    return snpg_within_job_ex(ctx, group, bases, bases_on_device, n_seqs, aln_len, row_stride, &job);
}

// the analysis of an already loaded group
int snpg_within_all(snpg_ctx *ctx, int group, int64_t B, double *prod_sums, double *prod_se) {
    if (!ctx) return SNPG_E_ARG;
    snpg_job job = {};
    job.struct_size = sizeof job;
    job.n_bootstraps = prod_se ? B : 0;
    job.prod_sums = prod_sums;
    job.prod_se = prod_se;
    return snpg_within_job_ex(ctx, group, nullptr, 0, 0, 0, 0, &job);
}

int64_t snpg_graph_replays(const snpg_ctx *ctx) { return ctx ? ctx->graph_replays : 0; }

}  // extern "C"

// "Next" row f2: the sliding-window statistics of SNPGenie_sliding_windows.R -- the script the reference's README
// recommends over the Perl window processor (README.md:281) -- on the engine's per-codon tables.
//
// The script (file:line of /root/reference/SNPGenie_sliding_windows.R):
//   :441-442  keeps the codons with num_defined_seqs >= MIN_DEFINED_CODONS and sorts them by codon_num
//   :444-445  if at least WINDOW_SIZE codons are left: for i from min(codon_num) to max(codon_num) - WINDOW_SIZE + 1 in
//             steps of STEP_SIZE the window is ROWS i .. i + WINDOW_SIZE - 1 of the kept table (a codon NUMBER used as
//             a row index: the two agree when the table starts at codon 1 and nothing was filtered out)
//   :241-321  per window: dN, dS, dN/dS, dN - dS from the summed sites and differences, bootstrap SEs over NBOOTSTRAPS
//             resamples of the window's codons (boot(): WINDOW_SIZE draws with replacement), Z-test P values, ASL counts
//   :492-500  the two-sided ASL P value
// A window whose rows reach past the kept table makes the script stop with an R error (a row of NAs cannot be assigned
// back); here the windows simply end with the last complete one.
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

#include "snpg_ctx.h"
#include "snpg_team.h"

using namespace snpg;

extern "C" {

int snpg_sliding_windows_r(snpg_ctx *ctx, const double *stats4, const uint32_t *n_defined, const int64_t *codon_num, int64_t C,
                           const snpg_rwindows *opt, int64_t *n_windows, double *out, int64_t out_cap) {
    if (!ctx || !opt) return SNPG_E_ARG;
    TEAM_FORWARD0(ctx, snpg_sliding_windows_r(c, stats4, n_defined, codon_num, C, opt, n_windows, out, out_cap));
    NEED(ctx, opt->struct_size == sizeof(snpg_rwindows), SNPG_E_ARG, "snpg_rwindows.struct_size does not match this library");
    NEED(ctx, stats4 && C > 0 && C < ((int64_t)1 << 31), SNPG_E_ARG, "bad codon table");
    NEED(ctx, opt->window >= 2 && opt->step >= 1, SNPG_E_ARG, "the window must be at least 2 codons and the step at least 1");       // R :40-46
    NEED(ctx, opt->n_bootstraps >= 2 && opt->n_bootstraps <= 1000000, SNPG_E_ARG, "NBOOTSTRAPS must be in [2, 1000000]");            // R :76-79
    NEED(ctx, (opt->numerator == 0 || opt->numerator == 1) && (opt->denominator == 0 || opt->denominator == 1), SNPG_E_ARG,
         "numerator / denominator: 0 = N, 1 = S");
    const int64_t W = opt->window, step = opt->step, B = opt->n_bootstraps;
    if (int rc = bind(ctx)) return rc;

    // :441-442 -- the kept codons in codon_num order (stable: the file order breaks ties, as dplyr::arrange does)
    std::vector<int64_t> rows((size_t)C);
    std::iota(rows.begin(), rows.end(), (int64_t)0);
    auto num_of = [&](int64_t r) { return codon_num ? codon_num[r] : r + 1; };
    if (n_defined) rows.erase(std::remove_if(rows.begin(), rows.end(), [&](int64_t r) { return (int64_t)n_defined[r] < opt->min_defined; }), rows.end());
    std::stable_sort(rows.begin(), rows.end(), [&](int64_t x, int64_t y) { return num_of(x) < num_of(y); });
    const int64_t kept = (int64_t)rows.size();
    // :444-445 -- window starts (1-based row i = a codon number), while the window's rows exist
    std::vector<int64_t> begs;
    if (kept >= W) {
        const int64_t lo = num_of(rows.front()), hi = num_of(rows.back()) - W + 1;
        for (int64_t i = lo; i <= hi; i += step) {
            if (i < 1 || i - 1 + W > kept) break;            // (the script fails here)
            begs.push_back(i - 1);
        }
    }
    const int64_t nwin = (int64_t)begs.size();
    NEED(ctx, nwin < ((int64_t)1 << 24), SNPG_E_ARG, "too many windows (%lld)", (long long)nwin);
    if (n_windows) *n_windows = nwin;
    if (!out || nwin == 0) return SNPG_OK;
    NEED(ctx, out_cap >= nwin, SNPG_E_ARG, "room for %lld windows, %lld needed", (long long)out_cap, (long long)nwin);

    // the kept table and the windows, on the device
    std::vector<double> tab((size_t)kept * 4);
    for (int64_t k = 0; k < kept; k++)
        for (int q = 0; q < 4; q++) tab[(size_t)(4 * k + q)] = stats4[4 * rows[(size_t)k] + q];
    std::vector<int64_t> be((size_t)(2 * nwin));
    for (int64_t k = 0; k < nwin; k++) { be[(size_t)k] = begs[(size_t)k]; be[(size_t)(nwin + k)] = begs[(size_t)k] + W; }
    CU(ctx, ctx->s_in.ensure(sizeof(double) * tab.size()));
    CU(ctx, cudaMemcpyAsync(ctx->s_in.p, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, ctx->s_misc.ensure(sizeof(int64_t) * be.size()));
    CU(ctx, cudaMemcpyAsync(ctx->s_misc.p, be.data(), sizeof(int64_t) * be.size(), cudaMemcpyHostToDevice, ctx->stream));
    const int64_t *d_beg = ctx->s_misc.as<int64_t>(), *d_end = d_beg + nwin;
    CU(ctx, ctx->s_sums.ensure(sizeof(double) * 4 * (size_t)nwin));
    { Launch l(ctx, SNPG_K_WINDOW); launch_window_sums(ctx->s_in.as<double>(), d_beg, d_end, nwin, ctx->s_sums.as<double>(), ctx->stream); }
    std::vector<double> sums((size_t)nwin * 4), st((size_t)nwin * 16);
    CU(ctx, cudaMemcpyAsync(sums.data(), ctx->s_sums.p, sizeof(double) * sums.size(), cudaMemcpyDeviceToHost, ctx->stream));

    // the window bootstrap: replicate sums of W draws uniform over the window's W codons (boot(): ordinary resampling),
    // in batches of windows that keep the replicate sums under 256 MB
    const int64_t batch = std::max<int64_t>(1, std::min<int64_t>(nwin, ((int64_t)256 << 20) / (32 * B)));
    CU(ctx, ctx->s_vals.ensure(sizeof(double) * 4 * (size_t)(batch * B)));
    CU(ctx, ctx->s_se.ensure(sizeof(double) * 16 * (size_t)batch));
    const uint32_t call = ctx->call_id++;
    const uint2 key = make_uint2((uint32_t)ctx->seed, (uint32_t)(ctx->seed >> 32));
    for (int64_t w0 = 0; w0 < nwin; w0 += batch) {
        const int64_t nb = std::min(batch, nwin - w0);
        {
            Launch l(ctx, SNPG_K_WINDOW);
            BootArgs a;
            a.stats4 = ctx->s_in.as<double>(); a.beg = d_beg + w0; a.end = d_end + w0; a.n_items = (int)nb; a.slot0 = 0;
            a.b_count = B; a.key = key; a.kind = kBootRWindow; a.call = call;
            a.rep = ctx->s_vals.as<double>();
            a.item0 = (uint32_t)w0;             // windows keep their own Philox streams whatever the batching
            launch_bootstrap(ctx->boot_kernel, a, W, ctx->stream);
        }
        {
            Launch l(ctx, SNPG_K_MOMENTS);
            launch_window_rstats(ctx->s_vals.as<double>(), ctx->s_sums.as<double>() + 4 * w0, nb, B, opt->numerator, opt->denominator,
                                 opt->jukes_cantor != 0, ctx->s_se.as<double>(), ctx->stream);
        }
        CU(ctx, cudaGetLastError());
        CU(ctx, cudaMemcpyAsync(st.data() + 16 * w0, ctx->s_se.p, sizeof(double) * 16 * (size_t)nb, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, sync_stream(ctx));              // the next batch reuses the replicate buffer
    }
    CU(ctx, sync_stream(ctx));

    const int nu = opt->numerator, de = opt->denominator;
    for (int64_t k = 0; k < nwin; k++) {
        double *o = out + SNPG_RWIN_COLS * k;
        const int64_t b = begs[(size_t)k];
        int64_t lo = num_of(rows[(size_t)b]), hi = lo;           // :450-451 min / max codon_num of the window's rows
        for (int64_t j = b; j < b + W; j++) { lo = std::min(lo, num_of(rows[(size_t)j])); hi = std::max(hi, num_of(rows[(size_t)j])); }
        o[0] = (double)lo; o[1] = (double)(lo + hi) / 2.0; o[2] = (double)hi; o[3] = (double)B;
        const double *s4 = &sums[(size_t)(4 * k)];
        o[4] = s4[2 + nu]; o[5] = s4[2 + de]; o[6] = s4[nu]; o[7] = s4[de];              // :512-515
        for (int q = 0; q < 16; q++) o[8 + q] = st[(size_t)(16 * k + q)];
    }
    return SNPG_OK;
}

// "Next" row f4: the pooled-mode codon contraction of snpgenie.pl (:6631-6760) -- see k_pooled_diffs.
int snpg_pooled_diffs(snpg_ctx *ctx, const double *freq, const double *coverage, int64_t C, double *n_diffs, double *s_diffs) {
    if (!ctx) return SNPG_E_ARG;
    TEAM_FORWARD0(ctx, snpg_pooled_diffs(c, freq, coverage, C, n_diffs, s_diffs));
    NEED(ctx, freq && coverage && n_diffs && s_diffs && C > 0 && C < ((int64_t)1 << 31), SNPG_E_ARG, "bad pooled-codon arguments");
    if (int rc = bind(ctx)) return rc;
    CU(ctx, ctx->s_in.ensure(sizeof(double) * 65 * (size_t)C));
    CU(ctx, ctx->s_sums.ensure(sizeof(double) * 2 * (size_t)C));
    double *d_freq = ctx->s_in.as<double>(), *d_cov = d_freq + 64 * C, *d_out = ctx->s_sums.as<double>();
    CU(ctx, cudaMemcpyAsync(d_freq, freq, sizeof(double) * 64 * (size_t)C, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(d_cov, coverage, sizeof(double) * (size_t)C, cudaMemcpyHostToDevice, ctx->stream));
    { Launch l(ctx, SNPG_K_STATS); launch_pooled_diffs(d_freq, d_cov, C, ctx->d_tables.as<double>(), d_out, d_out + C, ctx->stream); }
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(n_diffs, d_out, sizeof(double) * (size_t)C, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(s_diffs, d_out + C, sizeof(double) * (size_t)C, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, sync_stream(ctx));
    for (int64_t c = 0; c < C; c++)                      // the script dies here (:6727-6744)
        NEED(ctx, std::isfinite(n_diffs[c]) && std::isfinite(s_diffs[c]), SNPG_E_INPUT,
             "an impossible number of sequenced alleles given the allele frequencies reported at codon %lld (coverage %g)", (long long)(c + 1), coverage[c]);
    return SNPG_OK;
}

}  // extern "C"

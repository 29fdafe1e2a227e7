// FASTA ingest ("next" row f3; snpgenie_within_group.pl:204-253, snpgenie_between_group.pl:133-187).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "snpg_ctx.h"
#include "snpg_team.h"

using namespace snpg;

extern "C" {

int snpg_read_fasta(snpg_ctx *ctx, int slot, const char *path, const uint8_t **bases, uint64_t *n_seqs, uint64_t *aln_len) {
    if (!ctx || !path || !bases || !n_seqs || !aln_len) return SNPG_E_ARG;
    TEAM_FORWARD0(ctx, snpg_read_fasta(c, slot, path, bases, n_seqs, aln_len));
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
        CU(ctx, cudaHostAlloc(reinterpret_cast<void **>(&h.p), need, cudaHostAllocPortable));
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
    TEAM_FORWARD0(ctx, snpg_codon_strings(c, slot, product, codon, out));
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

}  // extern "C"

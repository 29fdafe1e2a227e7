// FASTA ingest ("next" row f3; snpgenie_within_group.pl:204-253, snpgenie_between_group.pl:133-187).
//
// The reference reads the file line by line: a line that contains '>' starts a record, every other line is chomp'ed (LF
// only) and appended to the current record; all records must have one length.  Here:
//   pass 1 (all host threads, one byte range of the memory-mapped file each): where the header lines are and how many
//          residues lie between them; a short sequential merge turns that into one offset per record and checks the lengths;
//          the pinned destination (at most the file's size) is allocated meanwhile;
//   pass 2 (the same threads, in the background): records are copied into their rows in chunks handed out in row order.
// snpg_read_fasta returns as soon as pass 2 is running.  Whoever reads the buffer first waits for the rows it needs
// (ingest_wait_rows): snpg_load_group asks row block by row block, so the host-to-device copy of block i runs while the
// threads still parse block i + 1.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <memory>
#include <mutex>
#include <thread>

#include "snpg_ctx.h"
#include "snpg_team.h"

using namespace snpg;

namespace snpg {

struct ParseJob {
    const char *file = nullptr; size_t size = 0;          // the mapping, released by the last worker
    uint8_t *dst = nullptr; uint64_t n = 0, len = 0;      // rows of len residues
    std::vector<uint64_t> data_begin, data_end;           // per record: its data lines lie in [begin, end) of the file
    uint64_t rows_per_chunk = 1, n_chunks = 0;
    std::atomic<uint64_t> next_chunk{0};
    std::unique_ptr<std::atomic<uint8_t>[]> done;         // per chunk
    uint64_t watermark = 0;                               // chunks [0, watermark) are known complete (one reader at a time)
    std::mutex wait_lock;
    std::vector<std::thread> workers;
    ~ParseJob() {
        for (auto &t : workers) if (t.joinable()) t.join();
        if (file) munmap(const_cast<char *>(file), size);
    }
    void work() {
        for (;;) {
            const uint64_t c = next_chunk.fetch_add(1);
            if (c >= n_chunks) return;
            const uint64_t r1 = std::min(n, (c + 1) * rows_per_chunk);
            for (uint64_t r = c * rows_per_chunk; r < r1; r++) {
                uint8_t *row = dst + r * len;
                uint64_t filled = 0;
                for (size_t pos = data_begin[r]; pos < data_end[r];) {
                    const char *nl = static_cast<const char *>(memchr(file + pos, '\n', data_end[r] - pos));
                    const size_t end = nl ? (size_t)(nl - file) : data_end[r];
                    memcpy(row + filled, file + pos, end - pos);      // (pass 1 checked that the lines add up to len)
                    filled += end - pos;
                    pos = end + 1;
                }
            }
            done[c].store(1, std::memory_order_release);
        }
    }
    void wait_rows(uint64_t rows) {                       // rows [0, rows) are in place when this returns
        const uint64_t need = std::min(n_chunks, (std::min(rows, n) + rows_per_chunk - 1) / rows_per_chunk);
        std::lock_guard<std::mutex> g(wait_lock);
        while (watermark < need) {
            if (done[watermark].load(std::memory_order_acquire)) watermark++;
            else std::this_thread::sleep_for(std::chrono::microseconds(20));
        }
    }
};

namespace {
std::mutex g_jobs_lock;
std::vector<std::shared_ptr<ParseJob>> g_jobs;           // running or finished parses whose buffers are still alive
}

// Blocks until rows [0, rows) of the FASTA buffer that contains p are parsed (no-op for any other memory).
void ingest_wait_rows(const uint8_t *p, uint64_t rows) {
    std::shared_ptr<ParseJob> job;
    {
        std::lock_guard<std::mutex> g(g_jobs_lock);
        for (auto &j : g_jobs)
            if (p >= j->dst && p < j->dst + j->n * j->len) { job = j; break; }
    }
    if (job) job->wait_rows(rows + (uint64_t)(p - job->dst) / job->len);
}
void ingest_forget(const uint8_t *buffer) {               // the buffer is about to be reused or freed
    std::shared_ptr<ParseJob> job;
    {
        std::lock_guard<std::mutex> g(g_jobs_lock);
        for (size_t i = 0; i < g_jobs.size(); i++)
            if (g_jobs[i]->dst == buffer) { job = g_jobs[i]; g_jobs.erase(g_jobs.begin() + (long)i); break; }
    }
    // (the job's destructor joins its workers when the last reference goes)
}

}  // namespace snpg

namespace {

// what one thread learns about its byte range [b, e) (both at line starts)
struct RangeScan {
    std::vector<uint64_t> hdr_begin, hdr_next;            // header lines: their first byte, the first byte after them
    std::vector<uint64_t> seg_bytes;                      // residues before the first header, then after each header
};

void scan_range(const char *f, size_t b, size_t e, RangeScan &out) {
    out.seg_bytes.push_back(0);
    for (size_t pos = b; pos < e;) {
        const char *nl = static_cast<const char *>(memchr(f + pos, '\n', e - pos));
        const size_t end = nl ? (size_t)(nl - f) : e;
        if (memchr(f + pos, '>', end - pos)) {
            out.hdr_begin.push_back(pos);
            out.hdr_next.push_back(std::min(e, end + 1));
            out.seg_bytes.push_back(0);
        } else {
            out.seg_bytes.back() += end - pos;
        }
        pos = end + 1;
    }
}

}  // namespace

extern "C" {

int snpg_read_fasta(snpg_ctx *ctx, int slot, const char *path, const uint8_t **bases, uint64_t *n_seqs, uint64_t *aln_len) {
    if (!ctx || !path || !bases || !n_seqs || !aln_len) return SNPG_E_ARG;
    TEAM_FORWARD0(ctx, snpg_read_fasta(c, slot, path, bases, n_seqs, aln_len));
    NEED(ctx, slot >= 0 && slot < SNPG_MAX_GROUPS, SNPG_E_ARG, "slot %d out of range", slot);
    if (int rc = bind(ctx)) return rc;
    Trace trace("snpg_read_fasta");
    const int fd = open(path, O_RDONLY);
    NEED(ctx, fd >= 0, SNPG_E_INPUT, "Could not open file %s", path);
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size == 0) { close(fd); return fail(ctx, SNPG_E_INPUT, "%s is empty or unreadable", path); }
    const size_t size = (size_t)st.st_size;
    const char *f = static_cast<const char *>(mmap(nullptr, size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0));
    close(fd);
    NEED(ctx, f != MAP_FAILED, SNPG_E_INPUT, "cannot map %s", path);
    auto job = std::make_shared<ParseJob>();
    job->file = f; job->size = size;                          // unmapped by the job's destructor from here on

    // the destination: at most `size` residues; allocated (pinning is slow) while pass 1 runs
    auto &h = ctx->host_aln[slot];
    if (h.p) ingest_forget(h.p);
    cudaError_t alloc_err = cudaSuccess;
    std::thread alloc;
    if (size > h.cap) {
        alloc = std::thread([&] {
            cudaSetDevice(ctx->device);
            if (h.p) cudaFreeHost(h.p);
            h.p = nullptr; h.cap = 0;
            alloc_err = cudaHostAlloc(reinterpret_cast<void **>(&h.p), size, cudaHostAllocPortable);
            if (alloc_err == cudaSuccess) h.cap = size;
        });
    }

    // pass 1: header lines and residue counts per byte range
    static const unsigned hw = [] { const char *v = getenv("SNPG_INGEST_THREADS"); const unsigned n = v ? (unsigned)atoi(v) : std::thread::hardware_concurrency(); return std::max(1u, std::min(64u, n)); }();
    const unsigned T = (unsigned)std::max<size_t>(1, std::min<size_t>(hw, size / ((size_t)1 << 20) + 1));
    std::vector<size_t> cut(T + 1, size);
    cut[0] = 0;
    for (unsigned t = 1; t < T; t++) {
        const size_t at = std::max(cut[t - 1], size / T * t);
        const char *nl = at < size ? static_cast<const char *>(memchr(f + at, '\n', size - at)) : nullptr;
        cut[t] = nl ? (size_t)(nl - f) + 1 : size;
    }
    std::vector<RangeScan> scans(T);
    {
        std::vector<std::thread> th;
        for (unsigned t = 1; t < T; t++) th.emplace_back(scan_range, f, cut[t], cut[t + 1], std::ref(scans[t]));
        scan_range(f, cut[0], cut[1], scans[0]);
        for (auto &x : th) x.join();
    }
    // merge: one (begin, end) per record, all of one length
    uint64_t first_len = 0, cur_len = 0;
    bool open_rec = false, ragged = false;
    auto close_rec = [&](uint64_t end) {
        if (!open_rec) return;
        if (job->data_begin.size() == 1) first_len = cur_len;
        else if (cur_len != first_len) ragged = true;
        job->data_end.push_back(end);
    };
    for (unsigned t = 0; t < T; t++) {
        const RangeScan &r = scans[t];
        if (open_rec) cur_len += r.seg_bytes[0];                       // (lines before the first header are ignored: wg:226)
        for (size_t k = 0; k < r.hdr_begin.size(); k++) {
            close_rec(r.hdr_begin[k]);
            open_rec = true;
            job->data_begin.push_back(r.hdr_next[k]);
            cur_len = r.seg_bytes[k + 1];
        }
    }
    close_rec(size);
    if (alloc.joinable()) alloc.join();
    const uint64_t n = job->data_begin.size();
    NEED(ctx, n > 0 && first_len > 0, SNPG_E_INPUT, "no FASTA records in %s", path);
    NEED(ctx, !ragged, SNPG_E_INPUT, "The sequences must be aligned, i.e., must be the same length.");
    if (alloc_err != cudaSuccess) return fail(ctx, SNPG_E_NOMEM, "cudaHostAlloc(%zu bytes) failed: %s", size, cudaGetErrorString(alloc_err));

    // pass 2 in the background: chunks of about 2 MB of rows, handed out in row order
    job->dst = h.p; job->n = n; job->len = first_len;
    job->rows_per_chunk = std::max<uint64_t>(1, ((uint64_t)2 << 20) / first_len);
    job->n_chunks = (n + job->rows_per_chunk - 1) / job->rows_per_chunk;
    job->done.reset(new std::atomic<uint8_t>[job->n_chunks]);
    for (uint64_t c = 0; c < job->n_chunks; c++) job->done[c].store(0, std::memory_order_relaxed);
    const unsigned W = (unsigned)std::min<uint64_t>(T, job->n_chunks);
    ParseJob *raw = job.get();
    for (unsigned t = 0; t < W; t++) job->workers.emplace_back([raw] { raw->work(); });
    {
        std::lock_guard<std::mutex> g(g_jobs_lock);
        g_jobs.push_back(job);
    }
    h.n = n;
    h.len = first_len;
    *bases = h.p;
    *n_seqs = n;
    *aln_len = first_len;
    if (!ctx->ingest_async) job->wait_rows(n);
    return SNPG_OK;
}

int snpg_fasta_wait(snpg_ctx *ctx, int slot) {
    if (!ctx) return SNPG_E_ARG;
    TEAM_FORWARD0(ctx, snpg_fasta_wait(c, slot));
    NEED(ctx, slot >= 0 && slot < SNPG_MAX_GROUPS && ctx->host_aln[slot].p, SNPG_E_STATE, "no FASTA was read into slot %d", slot);
    ingest_wait_rows(ctx->host_aln[slot].p, ctx->host_aln[slot].n);
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
    ingest_wait_rows(h.p, h.n);
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

// hostlogic.cpp -- the O(sketch size) host steps around the GPU path, and the callers of it.
//
//   fb2_filter_counts        <- FilterParams::filter_counts      (lib/src/filtering.rs:60-87)
//     strand filter          <- filter_strands                   (filtering.rs:413-432)
//     error cutoff           <- guess_filter_threshold + hist    (filtering.rs:154-195, statistics.rs:30-47)
//     abundance filter       <- filter_abundance                 (filtering.rs:329-343)
//   fb2_process_post_filter  <- SketchParams::process_post_filter (sketch_schemes/mod.rs:115-128)
//   fb2_sketch_stream/files  <- sketch_stream / sketch_files     (lib/src/lib.rs:29-94)
//   fb2_distance_finish      <- raw_distance tail + distance     (distance.rs:117-125, :35-41)
// These stay on the host exactly as SURVEY 8a (row S14, D2) scopes them: a few f64 compares over
// <= kmers_to_sketch entries.  Written independently of oracle/ (which is test infrastructure).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#include <zlib.h>

#include "../../include/finch_b200.h"

int fb2_fail(int code, const std::string &msg);  // engine.cu
void fb2_sketcher_hint_finish(fb2_sketcher *s, uint64_t final_size, int filter_on);   // engine.cu (internal)

// ---- filters ---------------------------------------------------------------------------------
static void compact(fb2_result *r, const std::vector<uint8_t> &keep) {
    uint64_t m = 0;
    const size_t st = r->kmer_stride;
    for (uint64_t i = 0; i < r->n; ++i) {
        if (!keep[i]) continue;
        if (m != i) {
            r->hashes[m] = r->hashes[i]; r->counts[m] = r->counts[i]; r->extras[m] = r->extras[i];
            memmove(r->kmers + m * st, r->kmers + i * st, st);
        }
        ++m;
    }
    r->n = m;
}

extern "C" uint32_t fb2_guess_filter_threshold(const uint32_t *counts, size_t n, double filter_level) {
    // histogram of counts: hist[c-1] = number of k-mers seen c times (statistics.rs:30-47)
    uint32_t max_count = 0;
    for (size_t i = 0; i < n; ++i) max_count = std::max(max_count, counts[i]);
    std::vector<uint64_t> hist(max_count, 0);
    for (size_t i = 0; i < n; ++i) if (counts[i]) hist[counts[i] - 1]++;
    uint64_t total = 0;
    for (size_t c = 0; c < hist.size(); ++c) total += (uint64_t)(c + 1) * hist[c];
    const double cutoff_amt = filter_level * (double)total;
    // coverage index below which `filter_level` of the weighted data lies
    size_t wgt_cutoff = 0;
    uint64_t cum = 0;
    for (size_t c = 0; c < hist.size(); ++c) {
        cum += (uint64_t)wgt_cutoff * hist[c];
        if ((double)cum > cutoff_amt) break;
        ++wgt_cutoff;
    }
    if (wgt_cutoff == 0) return 1;
    // minimum of the sliding window sum left of the cutoff (ties: the right-most window)
    const size_t win = std::max<size_t>(1, wgt_cutoff / 20);
    uint64_t sum = 0;
    for (size_t c = 0; c < win; ++c) sum += hist[c];
    uint64_t lowest = sum;
    size_t lowest_idx = win - 1;
    for (size_t lo = 0, hi = win; hi < wgt_cutoff; ++lo, ++hi) {
        if (sum <= lowest) { lowest = sum; lowest_idx = hi; }
        sum -= hist[lo];
        sum += hist[hi];
    }
    return (uint32_t)lowest_idx + 1;
}

// Core of FilterParams::filter_counts on bare (count, extra) columns: writes the indices that
// survive, ascending, and updates `f` like the reference (filter_on resolved, abun_low raised).
int fb2_filter_select(const uint32_t *counts, const uint32_t *extras, size_t n, fb2_filter *f, int format,
                      std::vector<uint32_t> &keep) {
    if (f->filter_on < 0) {  // lib.rs:71-76
        if (format == FB2_FORMAT_FASTA) f->filter_on = 0;
        else if (format == FB2_FORMAT_FASTQ) f->filter_on = 1;
        else return fb2_fail(FB2_EEMPTY, "Should have got a type");
    }
    keep.resize(n);
    for (size_t i = 0; i < n; ++i) keep[i] = (uint32_t)i;
    const bool on = f->filter_on == 1;
    if (on && f->strand_filter > 0.0) {  // filter_strands (filtering.rs:413-432)
        size_t m = 0;
        for (size_t j = 0; j < keep.size(); ++j) {
            const uint32_t i = keep[j], c = counts[i];
            bool ok = true;
            if (c >= 16) {  // fewer observations are too noisy to call an adapter
                const uint32_t lowest = std::min(extras[i], c - extras[i]);
                ok = ((double)lowest / (double)c) >= f->strand_filter;
            }
            if (ok) keep[m++] = i;
        }
        keep.resize(m);
    }
    if (on && f->err_filter > 0.0) {     // guess_filter_threshold on what is left (filtering.rs:68-79)
        std::vector<uint32_t> c(keep.size());
        for (size_t j = 0; j < keep.size(); ++j) c[j] = counts[keep[j]];
        const uint32_t cutoff = fb2_guess_filter_threshold(c.data(), c.size(), f->err_filter);
        if (f->has_abun_low) { if (cutoff > f->abun_low) f->abun_low = cutoff; }
        else { f->has_abun_low = 1; f->abun_low = cutoff; }
    }
    if (on && (f->has_abun_low || f->has_abun_high)) {  // filter_abundance (filtering.rs:329-343)
        const uint32_t lo = f->has_abun_low ? f->abun_low : 0u;
        const uint32_t hi = f->has_abun_high ? f->abun_high : UINT32_MAX;
        size_t m = 0;
        for (size_t j = 0; j < keep.size(); ++j) {
            const uint32_t c = counts[keep[j]];
            if (lo <= c && c <= hi) keep[m++] = keep[j];
        }
        keep.resize(m);
    }
    return FB2_OK;
}

extern "C" int fb2_filter_counts(fb2_result *r, fb2_filter *f) {
    if (!r || !f) return fb2_fail(FB2_EINVAL, "null argument");
    std::vector<uint32_t> keep;
    const int rc = fb2_filter_select(r->counts, r->extras, (size_t)r->n, f, r->format, keep);
    if (rc != FB2_OK) return rc;
    std::vector<uint8_t> flag((size_t)r->n, 0);
    for (uint32_t i : keep) flag[i] = 1;
    compact(r, flag);
    r->filters = *f;
    return FB2_OK;
}

extern "C" int fb2_process_post_filter(fb2_result *r, const fb2_params *p, const char *name) {
    if (!r || !p) return fb2_fail(FB2_EINVAL, "null argument");
    if (p->kind == FB2_KIND_MASH) {
        if (r->n > p->final_size) r->n = p->final_size;
        if (!p->no_strict && r->n < p->final_size)
            return fb2_fail(FB2_ETOOFEW, std::string(name ? name : "") + " had too few kmers (" +
                                             std::to_string(r->n) + ") to sketch");
    }
    return FB2_OK;
}

// ---- sketch_stream / sketch_files ----------------------------------------------------------------
static int finish_sketch(fb2_sketcher *s, const char *name, const fb2_params *p, const fb2_filter *f,
                         fb2_result *out) {
    return fb2_sketcher_sketch(s, name, p, f, out);
}

extern "C" int fb2_sketch_stream(const uint8_t *bytes, size_t len, const char *name, const fb2_params *p,
                                 const fb2_filter *f, fb2_result *out) {
    if (!p || !f || !out || (!bytes && len)) return fb2_fail(FB2_EINVAL, "null argument");
    memset(out, 0, sizeof(*out));
    fb2_sketcher *s = nullptr;
    int rc = fb2_sketcher_create(p, &s);
    if (rc != FB2_OK) return rc;
    if (p->kind == FB2_KIND_MASH) fb2_sketcher_hint_finish(s, p->final_size, f->filter_on);
    rc = fb2_sketcher_feed_fastx(s, bytes, len, 1);
    if (rc == FB2_OK) rc = finish_sketch(s, name, p, f, out);
    fb2_sketcher_destroy(s);
    return rc;
}

// One file through handle `s` (lib.rs:51-94 per file): read in pieces into the worker's pinned buffer,
// feed the raw bytes, finish the sketch.  gzip input (needletail sniffs the 1f 8b magic, lib.rs:60 via
// parse_fastx_reader) is inflated on the host by this worker thread, concatenated members included;
// bz2 / xz stay unsupported (no headers for them in this image) and are reported by the engine.
static int sketch_one_file(fb2_sketcher *s, bool reuse, const char *path, uint8_t *buf, size_t piece,
                           const fb2_params *p, const fb2_filter *f, fb2_result *out) {
    const bool is_stdin = strcmp(path, "-") == 0;  // lib.rs:38-40
    FILE *fp = is_stdin ? stdin : fopen(path, "rb");
    if (!fp) return fb2_fail(FB2_EIO, std::string(path) + ": No such file or directory");
    int rc = reuse ? fb2_sketcher_reset(s) : FB2_OK;
    if (p->kind == FB2_KIND_MASH) fb2_sketcher_hint_finish(s, p->final_size, f->filter_on);
    bool any = false;
    // sniff the first two bytes
    unsigned char magic[2] = {0, 0};
    const size_t nmagic = rc == FB2_OK ? fread(magic, 1, 2, fp) : 0;
    if (rc == FB2_OK && nmagic == 2 && magic[0] == 0x1f && magic[1] == 0x8b) {
        std::vector<unsigned char> in(1u << 20);
        memcpy(in.data(), magic, 2);
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, 15 + 32) != Z_OK) rc = fb2_fail(FB2_EIO, std::string(path) + ": zlib initialisation failed");
        bool zopen = rc == FB2_OK;
        zs.next_in = in.data(); zs.avail_in = 2;
        size_t fill = 0;
        bool eof = false, member_done = false;
        while (rc == FB2_OK) {
            if (zs.avail_in == 0 && !eof) {
                const size_t got = fread(in.data(), 1, in.size(), fp);
                if (got == 0) eof = true;
                zs.next_in = in.data(); zs.avail_in = (uInt)got;
            }
            if (zs.avail_in == 0 && eof) {
                if (!member_done) rc = fb2_fail(FB2_EIO, std::string(path) + ": truncated gzip stream");
                break;
            }
            if (member_done) {                      // another gzip member follows (MultiGzDecoder semantics)
                if (inflateReset(&zs) != Z_OK) { rc = fb2_fail(FB2_EIO, std::string(path) + ": zlib reset failed"); break; }
                member_done = false;
            }
            zs.next_out = buf + fill; zs.avail_out = (uInt)std::min<size_t>(piece - fill, 1u << 30);
            const int zr = inflate(&zs, Z_NO_FLUSH);
            fill = (size_t)(zs.next_out - buf);
            if (zr == Z_STREAM_END) member_done = true;
            else if (zr != Z_OK && zr != Z_BUF_ERROR) { rc = fb2_fail(FB2_EIO, std::string(path) + ": corrupt gzip stream"); break; }
            if (fill == piece) { any = true; rc = fb2_sketcher_feed_fastx(s, buf, fill, 0); fill = 0; }
        }
        if (rc == FB2_OK && fill) { any = true; rc = fb2_sketcher_feed_fastx(s, buf, fill, 0); }
        if (zopen) inflateEnd(&zs);
    } else {
        size_t fill = nmagic;
        if (nmagic) memcpy(buf, magic, nmagic);
        while (rc == FB2_OK) {
            const size_t got = fread(buf + fill, 1, piece - fill, fp) + fill;
            fill = 0;
            if (got) { any = true; rc = fb2_sketcher_feed_fastx(s, buf, got, 0); }
            if (got < piece) break;
        }
    }
    if (!is_stdin) fclose(fp);
    if (rc == FB2_OK && !any) rc = fb2_fail(FB2_EEMPTY, std::string(path) + ": empty input");
    if (rc == FB2_OK) rc = fb2_sketcher_feed_fastx(s, nullptr, 0, 1);
    if (rc == FB2_OK) rc = finish_sketch(s, path, p, f, out);
    return rc;
}

// Idle worker handles (sketcher + pinned read buffer) kept between sketch_files calls: creating a handle
// allocates device buffers and a pinned read buffer (tens of ms, serialised by the driver), which would otherwise
// dominate batches of small files.  The reference API has no release call, so retention is bounded: at most
// FB2_POOL_MAX handles (default 2; 0 disables pooling) stay behind a call, whatever the worker count was, and a
// handle is only re-used under the chunk / log settings it was created with.  fb2_sketch_files_release_pool()
// frees them (the Python mirror registers it with atexit).
struct PoolEntry { fb2_params p; fb2_sketcher *s; uint8_t *buf; std::string env; };
static std::mutex g_pool_mu;
static std::vector<PoolEntry> g_pool;
static size_t pool_max() {
    if (const char *e = getenv("FB2_POOL_MAX")) { const long v = atol(e); if (v >= 0 && v <= 64) return (size_t)v; }
    return 2;
}
static std::string pool_env_key() {
    const char *a = getenv("FB2_CHUNK_MB"), *b = getenv("FB2_LOG_M"), *c = getenv("FB2_TABLE_MULT");
    return std::string(a ? a : "") + "/" + (b ? b : "") + "/" + (c ? c : "");
}
static bool same_sketcher_params(const fb2_params &a, const fb2_params &b) {
    return a.kind == b.kind && a.kmers_to_sketch == b.kmers_to_sketch && a.kmer_length == b.kmer_length &&
           a.hash_seed == b.hash_seed && a.scale == b.scale && a.device == b.device && a.stream == nullptr &&
           b.stream == nullptr;
}
static bool pool_acquire(const fb2_params *p, fb2_sketcher **s, uint8_t **buf) {
    std::lock_guard<std::mutex> g(g_pool_mu);
    const std::string key = pool_env_key();
    for (size_t i = 0; i < g_pool.size(); ++i)
        if (same_sketcher_params(g_pool[i].p, *p) && g_pool[i].env == key) {
            *s = g_pool[i].s; *buf = g_pool[i].buf;
            g_pool.erase(g_pool.begin() + (long)i);
            return true;
        }
    return false;
}
static bool pool_release(const fb2_params *p, fb2_sketcher *s, uint8_t *buf) {
    std::lock_guard<std::mutex> g(g_pool_mu);
    if (g_pool.size() >= pool_max() || p->stream) return false;
    g_pool.push_back(PoolEntry{*p, s, buf, pool_env_key()});
    return true;
}
extern "C" void fb2_sketch_files_release_pool(void) {
    std::vector<PoolEntry> old;
    { std::lock_guard<std::mutex> g(g_pool_mu); old.swap(g_pool); }
    for (auto &e : old) { cudaFreeHost(e.buf); fb2_sketcher_destroy(e.s); }
}

// sketch_files (lib.rs:29-49): the reference fans the files out over a rayon pool, one sketcher per task.
// Here a few host threads each own one sketcher handle (its own CUDA streams and buffers) and pull
// file indices from a shared counter, so the small kernels and host<->device round trips of different
// files overlap on the GPU; results land in input order.  FB2_FILE_WORKERS overrides the thread count.
extern "C" int fb2_sketch_files(const char *const *paths, size_t n, const fb2_params *p, const fb2_filter *f,
                                fb2_result *outs) {
    if (!p || !f || (n && (!paths || !outs))) return fb2_fail(FB2_EINVAL, "null argument");
    for (size_t i = 0; i < n; ++i) memset(&outs[i], 0, sizeof(fb2_result));
    if (!n) return FB2_OK;
    size_t workers = 8;
    if (const char *e = getenv("FB2_FILE_WORKERS")) { const long v = atol(e); if (v >= 1 && v <= 64) workers = (size_t)v; }
    workers = std::min(workers, n);
    for (size_t i = 0; i < n; ++i) if (strcmp(paths[i], "-") == 0) workers = 1;   // stdin is consumed in order
    const size_t piece = 32u << 20;
    std::atomic<size_t> next{0};
    std::atomic<int> first_rc{FB2_OK};
    std::mutex mu;
    std::string first_msg;
    auto fail = [&](int rc) {
        std::lock_guard<std::mutex> g(mu);
        if (first_rc.load() == FB2_OK) { first_rc.store(rc); first_msg = fb2_last_error(); }
    };
    auto work = [&]() {
        fb2_sketcher *s = nullptr;
        uint8_t *buf = nullptr;
        bool reuse = pool_acquire(p, &s, &buf);     // an idle handle of an earlier call, same parameters
        int rc = FB2_OK;
        if (!reuse) {
            rc = fb2_sketcher_create(p, &s);        // selects the device for this thread
            if (rc == FB2_OK && cudaHostAlloc((void **)&buf, piece, cudaHostAllocDefault) != cudaSuccess)
                rc = fb2_fail(FB2_ECUDA, "cudaHostAlloc failed (file read buffer)");
        }
        while (rc == FB2_OK && first_rc.load() == FB2_OK) {
            const size_t i = next.fetch_add(1);
            if (i >= n) break;
            rc = sketch_one_file(s, reuse, paths[i], buf, piece, p, f, &outs[i]);
            reuse = true;
        }
        if (rc != FB2_OK) fail(rc);
        if (rc == FB2_OK && s && buf && pool_release(p, s, buf)) return;   // kept for the next call
        if (buf) cudaFreeHost(buf);
        if (s) fb2_sketcher_destroy(s);
    };
    if (workers <= 1) work();
    else {
        std::vector<std::thread> th;
        for (size_t w = 0; w < workers; ++w) th.emplace_back(work);
        for (auto &t : th) t.join();
    }
    const int rc = first_rc.load();
    if (rc != FB2_OK) {
        for (size_t i = 0; i < n; ++i) fb2_result_free(&outs[i]);
        return fb2_fail(rc, first_msg);
    }
    return FB2_OK;
}

// ---- distance epilogue ---------------------------------------------------------------------------
extern "C" void fb2_distance_finish(const fb2_pair_out *p, uint8_t kmer_length, double *containment,
                                    double *jaccard, double *mash_distance, uint64_t *common_hashes,
                                    uint64_t *total_hashes) {
    const uint64_t common = p->common, i = p->i, j = p->j;
    const double cont = j == 0 ? 0.0 : (double)common / (double)j;
    const uint64_t total = i - common + j;
    const double jac = total == 0 ? 1.0 : (double)common / (double)total;
    double md = -1.0 * std::log((2.0 * jac) / (1.0 + jac)) / (double)kmer_length;
    md = (md != md) ? 0.0 : (md > 0.0 ? md : 0.0);  // f64::max(0, md)
    md = md < 1.0 ? md : 1.0;                       // f64::min(1, md)
    if (containment) *containment = cont;
    if (jaccard) *jaccard = jac;
    if (mash_distance) *mash_distance = md;
    if (common_hashes) *common_hashes = common;
    if (total_hashes) *total_hashes = total;
}

// old_distance (distance.rs:136-157) from the integers the GPU returns: for sorted distinct lists its pointer
// walk counts exactly |Q n R| (the `common` of fb2_dist_batch with scale 0) over all |R| reference hashes.
// An empty query with a non-empty reference panics in the reference (index out of bounds): FB2_EINVAL here.
extern "C" int fb2_old_distance_finish(uint64_t common, uint64_t query_len, uint64_t ref_len, uint8_t kmer_length,
                                       double *containment, double *jaccard, double *mash_distance,
                                       uint64_t *common_hashes, uint64_t *total_hashes) {
    if (query_len == 0 && ref_len != 0) return fb2_fail(FB2_EINVAL, "old_distance: index out of bounds: the query sketch is empty");
    const uint64_t total = ref_len;
    const double cont = (double)common / (double)total;
    const double jac = (double)common / (double)(common + 2 * (total - common));
    double md = -1.0 * std::log((2.0 * jac) / (1.0 + jac)) / (double)kmer_length;
    md = (md != md) ? 0.0 : (md > 0.0 ? md : 0.0);  // f64::max(0, md)
    md = md < 1.0 ? md : 1.0;                       // f64::min(1, md)
    if (containment) *containment = cont;
    if (jaccard) *jaccard = jac;
    if (mash_distance) *mash_distance = md;
    if (common_hashes) *common_hashes = common;
    if (total_hashes) *total_hashes = total;
    return FB2_OK;
}

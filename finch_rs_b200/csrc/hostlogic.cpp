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
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include "common.cuh"

#include "../../include/finch_b200.h"

int fb2_fail(int code, const std::string &msg);  // engine.cu
// engine.cu (internal API, not exported in the header)
void fb2_sketcher_hint_finish(fb2_sketcher *s, uint64_t final_size, int filter_on);
int fb2_sketcher_begin_range(fb2_sketcher *s, int format, uint32_t state, uint32_t prev1, uint32_t prev2,
                             const uint8_t *tail_syms, uint64_t raw_base, uint64_t ord_base, int strip);
int fb2_sketcher_end_range(fb2_sketcher *s, uint32_t *end_state, uint32_t *last_byte, uint64_t *first_bad_pos, uint64_t *len_bad_pos);
uint32_t fb2_sketcher_halo(const fb2_sketcher *s);
void fb2_sketcher_set_polite_copy(fb2_sketcher *s, unsigned piece_mb);
void fb2_sketcher_set_polite_sync(fb2_sketcher *s, int on);   // host waits yield / sleep instead of spinning
void fb2_sketcher_set_force_strip(fb2_sketcher *s, int on);   // the next FASTQ stream is framed on the host (FB2_HOST_STRIP=1's mode)
int fb2_sketcher_sketch_small(fb2_sketcher *s, const uint8_t *bytes, size_t len, const char *name, const fb2_params *p,
                              const fb2_filter *f, fb2_result *out);
void fb2_sketcher_set_link_flag(fb2_sketcher *s, std::atomic<int> *flag, int owner);
int fb2_sketcher_merge_from(fb2_sketcher *dst, fb2_sketcher *src);
size_t fb2_sketcher_device_bytes(const fb2_sketcher *s);

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

uint32_t fb2_threshold_from_hist(const std::vector<uint64_t> &hist, double filter_level);
extern "C" uint32_t fb2_guess_filter_threshold(const uint32_t *counts, size_t n, double filter_level) {
    // histogram of counts: hist[c-1] = number of k-mers seen c times (statistics.rs:30-47)
    uint32_t max_count = 0;
    for (size_t i = 0; i < n; ++i) max_count = std::max(max_count, counts[i]);
    std::vector<uint64_t> hist(max_count, 0);
    for (size_t i = 0; i < n; ++i) if (counts[i]) hist[counts[i] - 1]++;
    return fb2_threshold_from_hist(hist, filter_level);
}

// guess_filter_threshold (filtering.rs:154-195) from the histogram of counts: hist[c-1] = number of k-mers seen c times.
uint32_t fb2_threshold_from_hist(const std::vector<uint64_t> &hist, double filter_level) {
    uint64_t total = 0;
    for (size_t c = 0; c < hist.size(); ++c) total += (uint64_t)(c + 1) * hist[c];
    const double cutoff_amt = filter_level * (double)total;
    size_t wgt_cutoff = 0;
    uint64_t cum = 0;
    for (size_t c = 0; c < hist.size(); ++c) {
        cum += (uint64_t)wgt_cutoff * hist[c];
        if ((double)cum > cutoff_amt) break;
        ++wgt_cutoff;
    }
    if (wgt_cutoff == 0) return 1;
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

// Core of FilterParams::filter_counts on bare (count, extra) columns: writes the indices that survive, ascending,
// and updates `f` like the reference (filter_on resolved, abun_low raised).  `limit`: stop after that many survivors
// (the caller truncates to final_size anyway, mod.rs:115-128; when fewer survive, all of them are returned, so the
// "too few" count is exact).  Two light passes over the columns: strand decision + histogram, then the abundance cut.
int fb2_filter_select(const uint32_t *counts, const uint32_t *extras, size_t n, fb2_filter *f, int format,
                      std::vector<uint32_t> &keep, size_t limit) {
    if (f->filter_on < 0) {  // lib.rs:71-76
        if (format == FB2_FORMAT_FASTA) f->filter_on = 0;
        else if (format == FB2_FORMAT_FASTQ) f->filter_on = 1;
        else return fb2_fail(FB2_EEMPTY, "Should have got a type");
    }
    keep.clear();
    const bool on = f->filter_on == 1;
    const bool strand = on && f->strand_filter > 0.0, err = on && f->err_filter > 0.0;
    static thread_local std::vector<uint8_t> ok_tls;       // strand survivors (reused across calls)
    static thread_local std::vector<uint64_t> hist_tls;
    std::vector<uint8_t> &ok = ok_tls;                     // (one TLS look-up, not one per loop iteration)
    std::vector<uint64_t> &hist = hist_tls;
    if (strand) ok.resize(n);
    if (err) hist.assign(4096, 0);
    if (strand || err) {
        const double cut = f->strand_filter;
        uint8_t *okp = ok.data();
        uint64_t *hp = hist.data();
        size_t hcap = hist.size();
        uint64_t small[4][64] = {};
        for (size_t i = 0; i < n; ++i) {
            const uint32_t c = counts[i];
            bool pass = true;
            if (strand && c >= 16) {  // filter_strands (filtering.rs:413-432): fewer observations are too noisy to call an adapter
                const uint32_t lowest = std::min(extras[i], c - extras[i]);
                // lowest / c >= cut, decided by a multiply when it is not within 1e-12 of the boundary (rounding is
                // monotonic, so away from the boundary the rounded quotient compares like the real one); the
                // division itself (the reference's expression, ~20 cycles) only for the borderline entries
                const double prod = cut * (double)c, lw = (double)lowest;
                if (lw >= prod * (1.0 + 1e-12)) pass = true;
                else if (lw <= prod * (1.0 - 1e-12)) pass = false;
                else pass = (lw / (double)c) >= cut;
            }
            if (strand) okp[i] = pass;
            if (err && pass && c) {   // hist of what is left (statistics.rs:30-47; filtering.rs:68-79)
                if (c <= 64) small[i & 3][c - 1]++;   // most counts are 1 or 2: four copies break the store-to-load chain
                else {
                    if (c > hcap) { hist.resize(std::max<size_t>(c, hcap * 2), 0); hp = hist.data(); hcap = hist.size(); }
                    hp[c - 1]++;
                }
            }
        }
        if (err) for (int q = 0; q < 4; ++q) for (int c = 0; c < 64; ++c) hp[c] += small[q][c];
    }
    if (err) {
        size_t max_count = hist.size();
        while (max_count && hist[max_count - 1] == 0) --max_count;
        hist.resize(max_count);
        const uint32_t cutoff = fb2_threshold_from_hist(hist, f->err_filter);
        if (f->has_abun_low) { if (cutoff > f->abun_low) f->abun_low = cutoff; }
        else { f->has_abun_low = 1; f->abun_low = cutoff; }
    }
    const bool abun = on && (f->has_abun_low || f->has_abun_high);   // filter_abundance (filtering.rs:329-343)
    const uint32_t lo = f->has_abun_low ? f->abun_low : 0u, hi = f->has_abun_high ? f->abun_high : UINT32_MAX;
    const uint8_t *okp = strand ? ok.data() : nullptr;
    for (size_t i = 0; i < n && keep.size() < limit; ++i) {
        if (okp && !okp[i]) continue;
        if (abun && !(lo <= counts[i] && counts[i] <= hi)) continue;
        keep.push_back((uint32_t)i);
    }
    return FB2_OK;
}

extern "C" int fb2_filter_counts(fb2_result *r, fb2_filter *f) {
    if (!r || !f) return fb2_fail(FB2_EINVAL, "null argument");
    std::vector<uint32_t> keep;
    const int rc = fb2_filter_select(r->counts, r->extras, (size_t)r->n, f, r->format, keep, SIZE_MAX);
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
// Pinned file-read buffers of the sketch_files workers.  Pinning is the most expensive set-up call a worker makes and
// the driver serialises it (tools/alloc_cost.cu on the 16-core B200 box: 14 ms per 32 MiB alone, 92 / 190 ms each when
// 8 / 16 threads ask at once), so a buffer is only as large as the files need (rbuf_need) and says how large it is.
static constexpr size_t RBUF_HDR = 64;
static uint8_t *rbuf_alloc(size_t cap) {
    uint8_t *base = nullptr;
    if (cudaHostAlloc((void **)&base, cap + RBUF_HDR, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    *reinterpret_cast<size_t *>(base) = cap;
    return base + RBUF_HDR;
}
static size_t rbuf_cap(const uint8_t *buf) { return buf ? *reinterpret_cast<const size_t *>(buf - RBUF_HDR) : 0; }
static void rbuf_free(uint8_t *buf) { if (buf) cudaFreeHost(buf - RBUF_HDR); }
static bool pool_acquire(const fb2_params *p, fb2_sketcher **s, uint8_t **buf);
static bool pool_release(const fb2_params *p, fb2_sketcher *s, uint8_t *buf);
static int finish_sketch(fb2_sketcher *s, const char *name, const fb2_params *p, const fb2_filter *f,
                         fb2_result *out) {
    return fb2_sketcher_sketch(s, name, p, f, out);
}

static int sketch_stream_two_ended(const uint8_t *bytes, size_t len, const char *name, const fb2_params *p,
                                   const fb2_filter *f, fb2_result *out);
static thread_local fb2_stats g_last_stream_stats;
static void stats_add_delta(fb2_stats &acc, const fb2_stats &a, const fb2_stats &b) {   // acc += b - a
    acc.kernel_launches += b.kernel_launches - a.kernel_launches; acc.h2d_bytes += b.h2d_bytes - a.h2d_bytes;
    acc.d2h_bytes += b.d2h_bytes - a.d2h_bytes; acc.chunks += b.chunks - a.chunks; acc.prunes += b.prunes - a.prunes;
    acc.hash_launches += b.hash_launches - a.hash_launches; acc.hash_symbols += b.hash_symbols - a.hash_symbols;
    acc.provisional_redos += b.provisional_redos - a.provisional_redos; acc.band_passes += b.band_passes - a.band_passes;
}
extern "C" int fb2_last_stream_stats(fb2_stats *out) {
    if (!out) return fb2_fail(FB2_EINVAL, "null argument");
    *out = g_last_stream_stats;
    return FB2_OK;
}
static size_t env_size_h(const char *name, size_t dflt) {
    const char *e = getenv(name);
    if (!e || !*e) return dflt;
    const long long v = atoll(e);
    return v > 0 ? (size_t)v : dflt;
}
// gzip / bzip2 / xz bytes in memory -> plain bytes (the decoders of the file path below)
static int decompress_all(const uint8_t *bytes, size_t len, const char *name, std::vector<uint8_t> &plain);
extern "C" int fb2_sketch_stream(const uint8_t *bytes, size_t len, const char *name, const fb2_params *p,
                                 const fb2_filter *f, fb2_result *out) {
    if (!p || !f || !out || (!bytes && len)) return fb2_fail(FB2_EINVAL, "null argument");
    memset(out, 0, sizeof(*out));
    // sketch_stream takes any reader and needletail sniffs compressed input (lib.rs:58-60): so does this call
    if (len >= 2 && ((bytes[0] == 0x1f && bytes[1] == 0x8b) || (bytes[0] == 'B' && bytes[1] == 'Z') || (bytes[0] == 0xFD && bytes[1] == '7'))) {
        std::vector<uint8_t> plain;
        const int rcd = decompress_all(bytes, len, name, plain);
        if (rcd != FB2_OK) return rcd;
        if (plain.size() >= 2 && ((plain[0] == 0x1f && plain[1] == 0x8b) || (plain[0] == 'B' && plain[1] == 'Z') || (plain[0] == 0xFD && plain[1] == '7')))
            return fb2_fail(FB2_EFORMAT, std::string(name ? name : "") + ": compressed data inside compressed data");
        return fb2_sketch_stream(plain.data(), plain.size(), name, p, f, out);
    }
    {   // A large FASTQ stream from both ends, host-framed and raw at once (sketch_stream_two_ended): the default on a host
        // with >= 8 cores (FB2_HOST_STRIP unset), forced by FB2_HOST_STRIP=2; 0 / 1 pick the single-mode paths
        const char *e = getenv("FB2_HOST_STRIP");
        const bool two = p->kind != FB2_KIND_ALLCOUNTS && (e ? *e == '2' : std::thread::hardware_concurrency() >= 8);
        if (two && len >= (env_size_h("FB2_TWO_ENDED_MIN_KB", 256u << 10) << 10) && bytes[0] == '@') {
            const int rc2 = sketch_stream_two_ended(bytes, len, name, p, f, out);
            if (rc2 == FB2_OK) return FB2_OK;
            fb2_result_free(out);
            memset(out, 0, sizeof(*out));          // anything else: the plain path below
        }
    }
    fb2_params pd = *p;
    if (pd.device < 0 && cudaGetDevice(&pd.device) != cudaSuccess) pd.device = -1;   // pool entries are keyed by device
    fb2_sketcher *s = nullptr;
    uint8_t *buf = nullptr;
    int rc;
    if (pool_acquire(&pd, &s, &buf)) rc = fb2_sketcher_reset(s);
    else rc = fb2_sketcher_create(&pd, &s);
    fb2_stats st0, st1;
    memset(&st0, 0, sizeof(st0)); memset(&st1, 0, sizeof(st1));
    if (rc == FB2_OK) fb2_sketcher_stats(s, &st0);
    if (rc == FB2_OK && p->kind == FB2_KIND_MASH) fb2_sketcher_hint_finish(s, p->final_size, f->filter_on);
    bool done = false;
    if (rc == FB2_OK) {   // a small stream in one go (engine.cu); 1 = not that kind of stream, or it wants the general path
        const int rs = fb2_sketcher_sketch_small(s, bytes, len, name, p, f, out);
        if (rs == 1) {
            fb2_result_free(out);
            memset(out, 0, sizeof(*out));
            rc = fb2_sketcher_reset(s);
            if (rc == FB2_OK && p->kind == FB2_KIND_MASH) fb2_sketcher_hint_finish(s, p->final_size, f->filter_on);
        } else { rc = rs; done = true; }
    }
    if (rc == FB2_OK && !done) rc = fb2_sketcher_feed_fastx(s, bytes, len, 1);
    if (rc == FB2_OK && !done) rc = finish_sketch(s, name, p, f, out);
    if (rc == FB2_OK) {
        fb2_sketcher_stats(s, &st1);
        memset(&g_last_stream_stats, 0, sizeof(g_last_stream_stats));
        stats_add_delta(g_last_stream_stats, st0, st1);
    }
    if (rc != FB2_OK) {   // keep the message across the clean-up calls
        const std::string msg = fb2_last_error();
        rbuf_free(buf);
        if (s) fb2_sketcher_destroy(s);
        return fb2_fail(rc, msg);
    }
    if (!pool_release(&pd, s, buf)) { rbuf_free(buf); fb2_sketcher_destroy(s); }
    return FB2_OK;
}

// ---- compressed input (needletail's parse_fastx_reader sniffs two magic bytes: gzip 1f 8b, bzip2 "BZ", xz fd 37) -------
// gzip through zlib (linked).  bzip2 and xz through the system's libbz2.so.1.0 / liblzma.so.5, loaded at first use: the
// image carries the shared objects (Python's bz2 / lzma modules use them) but not their headers, so the few
// declarations needed are restated here from the libraries' stable public ABI.
struct StreamDecoder {
    virtual ~StreamDecoder() {}
    virtual const char *name() const = 0;
    // consume from [in, in + avail_in), produce into [out, out + avail_out); 0 = go on, 1 = end of a stream, -1 = corrupt
    virtual int step(const uint8_t *&in, size_t &avail_in, uint8_t *&out, size_t &avail_out) = 0;
    virtual bool next_member() { return false; }     // more input after a stream end: decode it as another stream?
};
struct GzipDecoder : StreamDecoder {
    z_stream zs;
    bool open = false;
    ~GzipDecoder() override { if (open) inflateEnd(&zs); }
    const char *name() const override { return "gzip"; }
    bool init() { memset(&zs, 0, sizeof zs); open = inflateInit2(&zs, 15 + 32) == Z_OK; return open; }
    int step(const uint8_t *&in, size_t &avail_in, uint8_t *&out, size_t &avail_out) override {
        zs.next_in = const_cast<Bytef *>(in); zs.avail_in = (uInt)std::min<size_t>(avail_in, 1u << 30);
        zs.next_out = out; zs.avail_out = (uInt)std::min<size_t>(avail_out, 1u << 30);
        const uInt in0 = zs.avail_in, out0 = zs.avail_out;
        const int zr = inflate(&zs, Z_NO_FLUSH);
        in += in0 - zs.avail_in; avail_in -= in0 - zs.avail_in; out += out0 - zs.avail_out; avail_out -= out0 - zs.avail_out;
        if (zr == Z_STREAM_END) return 1;
        return (zr == Z_OK || zr == Z_BUF_ERROR) ? 0 : -1;
    }
    bool next_member() override { return inflateReset(&zs) == Z_OK; }   // MultiGzDecoder: concatenated members
};
static std::unique_ptr<StreamDecoder> make_gzip_decoder(std::string &why) {
    std::unique_ptr<GzipDecoder> d(new GzipDecoder());
    if (!d->init()) { why = "zlib initialisation failed"; return nullptr; }
    return d;
}
// bzlib.h: bz_stream and the three decompression entry points
struct Bz2Stream {
    char *next_in; unsigned int avail_in, total_in_lo32, total_in_hi32;
    char *next_out; unsigned int avail_out, total_out_lo32, total_out_hi32;
    void *state; void *(*bzalloc)(void *, int, int); void (*bzfree)(void *, void *); void *opaque;
};
struct Bz2Decoder : StreamDecoder {
    Bz2Stream bs;
    int (*fn_init)(Bz2Stream *, int, int) = nullptr;
    int (*fn_run)(Bz2Stream *) = nullptr;
    int (*fn_end)(Bz2Stream *) = nullptr;
    bool open = false;
    ~Bz2Decoder() override { if (open) fn_end(&bs); }
    const char *name() const override { return "bzip2"; }
    int step(const uint8_t *&in, size_t &avail_in, uint8_t *&out, size_t &avail_out) override {
        bs.next_in = reinterpret_cast<char *>(const_cast<uint8_t *>(in)); bs.avail_in = (unsigned)std::min<size_t>(avail_in, 1u << 30);
        bs.next_out = reinterpret_cast<char *>(out); bs.avail_out = (unsigned)std::min<size_t>(avail_out, 1u << 30);
        const unsigned in0 = bs.avail_in, out0 = bs.avail_out;
        const int r = fn_run(&bs);                   // BZ_OK 0, BZ_STREAM_END 4, errors negative
        in += in0 - bs.avail_in; avail_in -= in0 - bs.avail_in; out += out0 - bs.avail_out; avail_out -= out0 - bs.avail_out;
        return r == 4 ? 1 : (r == 0 ? 0 : -1);
    }
};
static std::unique_ptr<StreamDecoder> make_bz2_decoder(std::string &why) {
    static void *lib = dlopen("libbz2.so.1.0", RTLD_NOW | RTLD_LOCAL);
    if (!lib) lib = dlopen("libbz2.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!lib) { why = "bzip2 input needs libbz2.so.1.0, which this system does not have"; return nullptr; }
    std::unique_ptr<Bz2Decoder> d(new Bz2Decoder());
    d->fn_init = reinterpret_cast<int (*)(Bz2Stream *, int, int)>(dlsym(lib, "BZ2_bzDecompressInit"));
    d->fn_run = reinterpret_cast<int (*)(Bz2Stream *)>(dlsym(lib, "BZ2_bzDecompress"));
    d->fn_end = reinterpret_cast<int (*)(Bz2Stream *)>(dlsym(lib, "BZ2_bzDecompressEnd"));
    memset(&d->bs, 0, sizeof d->bs);
    if (!d->fn_init || !d->fn_run || !d->fn_end || d->fn_init(&d->bs, 0, 0) != 0) { why = "libbz2 initialisation failed"; return nullptr; }
    d->open = true;
    return d;
}
// lzma/base.h: lzma_stream and the stream decoder
struct XzStream {
    const uint8_t *next_in; size_t avail_in; uint64_t total_in;
    uint8_t *next_out; size_t avail_out; uint64_t total_out;
    const void *allocator; void *internal;
    void *reserved_ptr1, *reserved_ptr2, *reserved_ptr3, *reserved_ptr4;
    uint64_t reserved_int1, reserved_int2; size_t reserved_int3, reserved_int4;
    int reserved_enum1, reserved_enum2;
};
struct XzDecoder : StreamDecoder {
    XzStream xs;
    int (*fn_init)(XzStream *, uint64_t, uint32_t) = nullptr;
    int (*fn_code)(XzStream *, int) = nullptr;
    void (*fn_end)(XzStream *) = nullptr;
    bool open = false;
    ~XzDecoder() override { if (open) fn_end(&xs); }
    const char *name() const override { return "xz"; }
    int step(const uint8_t *&in, size_t &avail_in, uint8_t *&out, size_t &avail_out) override {
        xs.next_in = in; xs.avail_in = avail_in; xs.next_out = out; xs.avail_out = avail_out;
        const int r = fn_code(&xs, 0 /* LZMA_RUN */);   // LZMA_OK 0, LZMA_STREAM_END 1, LZMA_BUF_ERROR 10
        in = xs.next_in; avail_in = xs.avail_in; out = xs.next_out; avail_out = xs.avail_out;
        return r == 1 ? 1 : ((r == 0 || r == 10) ? 0 : -1);
    }
};
static std::unique_ptr<StreamDecoder> make_xz_decoder(std::string &why) {
    static void *lib = dlopen("liblzma.so.5", RTLD_NOW | RTLD_LOCAL);
    if (!lib) { why = "xz input needs liblzma.so.5, which this system does not have"; return nullptr; }
    std::unique_ptr<XzDecoder> d(new XzDecoder());
    d->fn_init = reinterpret_cast<int (*)(XzStream *, uint64_t, uint32_t)>(dlsym(lib, "lzma_stream_decoder"));
    d->fn_code = reinterpret_cast<int (*)(XzStream *, int)>(dlsym(lib, "lzma_code"));
    d->fn_end = reinterpret_cast<void (*)(XzStream *)>(dlsym(lib, "lzma_end"));
    memset(&d->xs, 0, sizeof d->xs);                 // LZMA_STREAM_INIT
    if (!d->fn_init || !d->fn_code || !d->fn_end || d->fn_init(&d->xs, UINT64_MAX, 0) != 0) { why = "liblzma initialisation failed"; return nullptr; }
    d->open = true;
    return d;
}

// One file through handle `s` (lib.rs:51-94 per file): read in pieces into the worker's pinned buffer,
// feed the raw bytes, finish the sketch.  gzip input (needletail sniffs the 1f 8b magic, lib.rs:60 via
// parse_fastx_reader), bzip2 and xz input are decompressed on the host by this worker thread (StreamDecoder above).
// FB2_TRACE_FILES=1: where the workers of sketch_files spend their time (summed over files, printed per call)
static std::atomic<uint64_t> g_ns_reset{0}, g_ns_read{0}, g_ns_feed{0}, g_ns_finish{0}, g_n_files{0}, g_ns_create{0}, g_n_created{0};
static inline uint64_t now_ns() {
    return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static int decompress_all(const uint8_t *bytes, size_t len, const char *name, std::vector<uint8_t> &plain) {
    std::string why;
    std::unique_ptr<StreamDecoder> dec;
    if (bytes[0] == 0x1f) dec = make_gzip_decoder(why);
    else if (bytes[0] == 'B') dec = make_bz2_decoder(why);
    else dec = make_xz_decoder(why);
    const std::string who = name ? name : "";
    if (!dec) return fb2_fail(FB2_EUNSUPPORTED, who + ": " + why);
    const uint8_t *in = bytes;
    size_t avail_in = len;
    bool member_done = false;
    plain.clear();
    plain.resize(std::max<size_t>(1u << 20, len * 4));
    size_t fill = 0;
    while (true) {
        if (avail_in == 0) {
            if (!member_done) return fb2_fail(FB2_EIO, who + ": truncated " + dec->name() + " stream");
            break;
        }
        if (member_done) {                      // more bytes after the end of a stream
            if (!dec->next_member()) break;     // bz2 / xz: one stream, the rest is ignored; gzip: another member follows
            member_done = false;
        }
        if (fill == plain.size()) plain.resize(plain.size() * 2);
        uint8_t *o = plain.data() + fill;
        size_t avail_out = std::min<size_t>(plain.size() - fill, 1u << 30);
        const int st = dec->step(in, avail_in, o, avail_out);
        fill = (size_t)(o - plain.data());
        if (st == 1) member_done = true;
        else if (st < 0) return fb2_fail(FB2_EIO, who + ": corrupt " + dec->name() + " stream");
    }
    plain.resize(fill);
    return FB2_OK;
}
static bool big_regular_file(FILE *fp) {
    struct stat sb;
    if (getenv("FB2_NO_PARALLEL_READ")) return false;
    const uint64_t least = (uint64_t)env_size_h("FB2_BIG_FILE_KB", 64u << 10) << 10;   // (test hook: the large-file paths on small files)
    return fstat(fileno(fp), &sb) == 0 && S_ISREG(sb.st_mode) && (uint64_t)sb.st_size >= least;
}
// FASTQ files framed on the host straight from the mapping: where the in-memory stream does the same (>= 8 cores),
// FB2_FILE_MMAP=0 / 1 forbids / forces it
static bool mapped_fastq_ok() {
    if (const char *e = getenv("FB2_FILE_MMAP")) return atoi(e) != 0;
    if (const char *e = getenv("FB2_HOST_STRIP")) if (*e == '0') return false;
    return std::thread::hardware_concurrency() >= 8;
}
// Is (nearly) all of the file in the page cache?  Only then is framing from the mapping a win: a cold file is read
// from its device either way, and the parallel pread path asks for it in large sequential requests instead of page
// faults.  (A sample of the pages, through a short-lived mapping.)
static bool file_is_cached(int fd) {
    struct stat sb;
    if (fstat(fd, &sb) != 0 || sb.st_size <= 0) return false;
    const size_t len = (size_t)sb.st_size;
    void *map = mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0);
    if (map == MAP_FAILED) return false;
    const size_t page = (size_t)sysconf(_SC_PAGESIZE), n_pages = (len + page - 1) / page;
    size_t seen = 0, resident = 0;
    unsigned char vec[64];
    const size_t stride = std::max<size_t>(64, n_pages / 256 / 64 * 64);      // ~256 windows of 64 pages
    for (size_t p0 = 0; p0 < n_pages; p0 += stride) {
        const size_t k = std::min<size_t>(64, n_pages - p0);
        if (mincore((char *)map + p0 * page, k * page, vec) != 0) { seen = 1; resident = 0; break; }
        for (size_t i = 0; i < k; ++i) resident += vec[i] & 1u;
        seen += k;
    }
    munmap(map, len);
    return seen > 0 && resident * 100 >= seen * 97;
}
static int sketch_one_file(fb2_sketcher *s, bool reuse, const char *path, uint8_t *buf, size_t piece,
                           const fb2_params *p, const fb2_filter *f, fb2_result *out) {
    const bool is_stdin = strcmp(path, "-") == 0;  // lib.rs:38-40
    FILE *fp = is_stdin ? stdin : fopen(path, "rb");
    if (!fp) return fb2_fail(FB2_EIO, std::string(path) + ": No such file or directory");
    uint64_t t0 = now_ns();
    int rc = reuse ? fb2_sketcher_reset(s) : FB2_OK;
    g_ns_reset += now_ns() - t0;
    if (p->kind == FB2_KIND_MASH) fb2_sketcher_hint_finish(s, p->final_size, f->filter_on);
    bool any = false, fed_final = false;
    // sniff the first two bytes
    unsigned char magic[2] = {0, 0};
    const size_t nmagic = rc == FB2_OK ? fread(magic, 1, 2, fp) : 0;
    std::unique_ptr<StreamDecoder> dec;
    if (rc == FB2_OK && nmagic == 2) {
        std::string why;
        if (magic[0] == 0x1f && magic[1] == 0x8b) dec = make_gzip_decoder(why);
        else if (magic[0] == 'B' && magic[1] == 'Z') dec = make_bz2_decoder(why);
        else if (magic[0] == 0xFD && magic[1] == '7') dec = make_xz_decoder(why);
        if (!why.empty()) rc = fb2_fail(FB2_EUNSUPPORTED, std::string(path) + ": " + why);
    }
    if (rc == FB2_OK && dec) {
        std::vector<unsigned char> in(1u << 20);
        memcpy(in.data(), magic, 2);
        const uint8_t *next_in = in.data();
        size_t avail_in = 2, fill = 0;
        bool eof = false, member_done = false;
        while (rc == FB2_OK) {
            if (avail_in == 0 && !eof) {
                const size_t got = fread(in.data(), 1, in.size(), fp);
                if (got == 0) eof = true;
                next_in = in.data(); avail_in = got;
            }
            if (avail_in == 0 && eof) {
                if (!member_done) rc = fb2_fail(FB2_EIO, std::string(path) + ": truncated " + dec->name() + " stream");
                break;
            }
            if (member_done) {                      // more bytes after the end of a stream
                if (!dec->next_member()) break;     // bz2 / xz: one stream, the rest is ignored (BzDecoder / XzDecoder)
                member_done = false;                // gzip: another member follows (MultiGzDecoder)
            }
            uint8_t *next_out = buf + fill;
            size_t avail_out = std::min<size_t>(piece - fill, 1u << 30);
            const int st = dec->step(next_in, avail_in, next_out, avail_out);
            fill = (size_t)(next_out - buf);
            if (st == 1) member_done = true;
            else if (st < 0) { rc = fb2_fail(FB2_EIO, std::string(path) + ": corrupt " + dec->name() + " stream"); break; }
            if (fill == piece) { any = true; rc = fb2_sketcher_feed_fastx(s, buf, fill, 0); fill = 0; }
        }
        if (rc == FB2_OK && fill) { any = true; rc = fb2_sketcher_feed_fastx(s, buf, fill, 0); }
    } else if (rc == FB2_OK && !is_stdin && big_regular_file(fp) && nmagic == 2 && magic[0] == '@' && mapped_fastq_ok() &&
               file_is_cached(fileno(fp))) {
        // A large plain FASTQ file: the host cores frame its records straight out of the page cache (the file is
        // mapped, nothing is copied first) and only the sequence lines go to the pinned staging and over PCIe -- the
        // in-memory host-framed mode (strip.cpp, engine.cu feed_fastq_stripped), which reports malformed records itself.
        // Copying the file into pinned memory first (the branch below) is bound by that copy: 20-25 GB/s on the 16-core box,
        // 125-160 ms for the 3.1 GB of C2; framed from the mapping: 100 ms (half of it minor page faults).
        struct stat sb;
        const int fd = fileno(fp);
        void *map = MAP_FAILED;
        if (fstat(fd, &sb) == 0 && sb.st_size > 0) map = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (map == MAP_FAILED) rc = fb2_fail(FB2_EIO, std::string(path) + ": mmap failed");
        else {
            const size_t total = (size_t)sb.st_size, step = (size_t)512 << 20;
            madvise(map, total, MADV_SEQUENTIAL);
            fb2_sketcher_set_force_strip(s, 1);
            const uint64_t t1 = now_ns();
            // (setting the page tables up ahead with MADV_POPULATE_READ from all cores was slower than letting the framing
            // threads fault their own pages in: 153 vs 100 ms for C2)
            for (size_t off = 0; off < total && rc == FB2_OK; off += step) {
                any = true;
                const size_t n = std::min(step, total - off);
                rc = fb2_sketcher_feed_fastx(s, (const uint8_t *)map + off, n, 0);
            }
            if (rc == FB2_OK) { rc = fb2_sketcher_feed_fastx(s, nullptr, 0, 1); fed_final = true; }   // (the framing reads the mapping until here)
            g_ns_feed += now_ns() - t1;
            fb2_sketcher_set_force_strip(s, 0);
            munmap(map, total);
        }
    } else if (rc == FB2_OK && !is_stdin && big_regular_file(fp)) {
        // A large plain file: one thread copying it out of the page cache (~5 GB/s) would be 10x slower than the PCIe
        // link it feeds.  Several threads pread() slices of the next 32 MiB piece into a second pinned buffer while the
        // current piece is being copied to the GPU.
        uint8_t *bufs[2] = {buf, nullptr};
        if (cudaHostAlloc((void **)&bufs[1], piece, cudaHostAllocDefault) != cudaSuccess) rc = fb2_fail(FB2_ECUDA, "cudaHostAlloc failed (second read buffer)");
        const int fd = fileno(fp);
        unsigned T = (unsigned)std::min<size_t>(8, std::max<size_t>(1, std::thread::hardware_concurrency() / 2));
        if (const char *e = getenv("FB2_READ_THREADS")) { const long v = atol(e); if (v >= 1 && v <= 64) T = (unsigned)v; }
        auto read_piece = [&](uint8_t *dst, uint64_t off) -> long {      // bytes read, or -1
            std::vector<long> got(T, 0);
            const size_t each = (piece / T + 4095) & ~(size_t)4095;
            auto part = [&](unsigned t) {
                size_t a = (size_t)t * each, b = std::min(piece, a + each), done = 0;
                while (a + done < b) {
                    const ssize_t r = pread(fd, dst + a + done, b - a - done, (off_t)(off + a + done));
                    if (r < 0) { got[t] = -1; return; }
                    if (r == 0) break;
                    done += (size_t)r;
                }
                got[t] = (long)done;
            };
            std::vector<std::thread> th;
            for (unsigned t = 1; t < T; ++t) th.emplace_back(part, t);
            part(0);
            for (auto &x : th) x.join();
            long total = 0;
            for (unsigned t = 0; t < T; ++t) { if (got[t] < 0) return -1; total += got[t]; if ((size_t)got[t] < std::min(piece, (size_t)(t + 1) * each) - std::min(piece, (size_t)t * each)) break; }
            return total;
        };
        uint64_t off = 0;
        long n_cur = -2, n_next = -2;
        std::thread reader;
        uint64_t t0r = now_ns();
        if (rc == FB2_OK) n_cur = read_piece(bufs[0], 0);
        g_ns_read += now_ns() - t0r;
        for (int which = 0; rc == FB2_OK; which ^= 1) {
            if (n_cur < 0) { rc = fb2_fail(FB2_EIO, std::string(path) + ": read error"); break; }
            if (n_cur == 0) break;
            off += (uint64_t)n_cur;
            const bool more = (size_t)n_cur == piece;
            if (more) reader = std::thread([&, which, off] { n_next = read_piece(bufs[which ^ 1], off); });
            any = true;
            const uint64_t t1 = now_ns();
            rc = fb2_sketcher_feed_fastx(s, bufs[which], (size_t)n_cur, 0);
            g_ns_feed += now_ns() - t1;
            if (more) { reader.join(); n_cur = n_next; } else break;
        }
        if (reader.joinable()) reader.join();
        if (bufs[1]) cudaFreeHost(bufs[1]);
    } else {
        size_t fill = nmagic;
        if (nmagic) memcpy(buf, magic, nmagic);
        bool first = true;
        while (rc == FB2_OK) {
            t0 = now_ns();
            const size_t got = fread(buf + fill, 1, piece - fill, fp) + fill;
            const uint64_t t1 = now_ns();
            g_ns_read += t1 - t0;
            fill = 0;
            if (first && got && got < piece) {   // the whole file is here: the small-stream path (engine.cu), if it applies
                const int rs = fb2_sketcher_sketch_small(s, buf, got, path, p, f, out);
                if (rs != 1) {
                    g_ns_feed += now_ns() - t1;
                    if (!is_stdin) fclose(fp);
                    g_n_files += 1;
                    return rs;
                }
                fb2_result_free(out);
                memset(out, 0, sizeof(*out));
                rc = fb2_sketcher_reset(s);
                if (rc == FB2_OK && p->kind == FB2_KIND_MASH) fb2_sketcher_hint_finish(s, p->final_size, f->filter_on);
                if (rc != FB2_OK) break;
            }
            first = false;
            if (got) { any = true; rc = fb2_sketcher_feed_fastx(s, buf, got, 0); }
            g_ns_feed += now_ns() - t1;
            if (got < piece) break;
        }
    }
    if (!is_stdin) fclose(fp);
    if (rc == FB2_OK && !any) rc = fb2_fail(FB2_EEMPTY, std::string(path) + ": empty input");
    t0 = now_ns();
    if (rc == FB2_OK && !fed_final) rc = fb2_sketcher_feed_fastx(s, nullptr, 0, 1);
    g_ns_feed += now_ns() - t0;
    t0 = now_ns();
    if (rc == FB2_OK) rc = finish_sketch(s, path, p, f, out);
    g_ns_finish += now_ns() - t0;
    g_n_files += 1;
    return rc;
}

// Idle worker handles (sketcher + pinned read buffer) kept between sketch_* calls: creating a handle allocates
// device buffers, pinned mirrors, streams and events (10+ ms, partly serialised by the driver), which would otherwise
// dominate small files.  The reference API has no release call, so retention is bounded PER DEVICE: at most
// FB2_POOL_MAX handles (default 16; 0 disables pooling) holding at most FB2_POOL_MB of device memory together
// (default 4096 MiB: a handle that sketched a 5 Mbp FASTA holds ~100 MB, one that streamed a large FASTQ ~1.3 GB -- the
// two-ended stream mode keeps two of those), and
// a handle is only re-used under the chunk / log settings it was created with.  fb2_sketch_files_release_pool()
// frees them (the Python mirror registers it with atexit).
struct PoolEntry { fb2_params p; fb2_sketcher *s; uint8_t *buf; std::string env; };
static std::mutex g_pool_mu;
static std::vector<PoolEntry> g_pool;
static size_t pool_max() {
    if (const char *e = getenv("FB2_POOL_MAX")) { const long v = atol(e); if (v >= 0 && v <= 64) return (size_t)v; }
    return 16;
}
static size_t pool_max_bytes() {
    if (const char *e = getenv("FB2_POOL_MB")) { const long v = atol(e); if (v >= 0) return (size_t)v << 20; }
    return (size_t)4096 << 20;
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
    size_t on_device = 0, bytes = fb2_sketcher_device_bytes(s);
    for (const auto &e : g_pool) if (e.p.device == p->device) { ++on_device; bytes += fb2_sketcher_device_bytes(e.s); }
    if (on_device >= pool_max() || bytes > pool_max_bytes() || p->stream) return false;
    g_pool.push_back(PoolEntry{*p, s, buf, pool_env_key()});
    return true;
}
extern "C" void fb2_sketch_files_release_pool(void) {
    std::vector<PoolEntry> old;
    { std::lock_guard<std::mutex> g(g_pool_mu); old.swap(g_pool); }
    for (auto &e : old) { rbuf_free(e.buf); fb2_sketcher_destroy(e.s); }
}

// sketch_files (lib.rs:29-49): the reference fans the files out over a rayon pool, one sketcher per task.
// Here every GPU gets a few host threads, each owning one sketcher handle (its own CUDA streams and buffers);
// files are assigned to GPUs longest-first (LPT by file size) and a GPU's threads pull from its own list, so the
// small kernels and host<->device round trips of different files overlap on each GPU; results land in input
// order.  The finished sketches need no device-side gather: this is ONE process driving all GPUs, every worker
// copies its 24 KB result straight into the caller's host array (a gather to one GPU followed by a copy to the
// host would move the same bytes twice; the NCCL gather lives where there is one process per GPU: bench.py).
// FB2_FILE_WORKERS overrides the thread count per GPU.
// Read-buffer size for one file: whole small files in one read (+ 1 byte, so that the read sees the end of the file and
// the small-stream path applies), in steps of 4 MiB; `piece` for large files, pipes and whatever cannot be stat'ed.
static size_t rbuf_need(const char *path, size_t piece) {
    struct stat sb;
    if (strcmp(path, "-") == 0 || stat(path, &sb) != 0 || !S_ISREG(sb.st_mode)) return piece;
    const uint64_t want = ((uint64_t)sb.st_size + (uint64_t)sb.st_size / 8 + 4096 + (4u << 20) - 1) / (4u << 20) * (4u << 20);
    return (size_t)std::min<uint64_t>(piece, want);
}
static int sketch_files_on(const char *const *paths, size_t n, const fb2_params *p, const fb2_filter *f,
                           fb2_result *outs, const std::vector<int> &devices) {
    for (size_t i = 0; i < n; ++i) memset(&outs[i], 0, sizeof(fb2_result));
    if (!n) return FB2_OK;
    const size_t G = devices.size();
    // Worker threads per GPU.  A worker's time per file is half file reading (page cache -> pinned buffer, ~5 GB/s per
    // thread) and half waiting on its GPU, so about one thread per core over all GPUs is where the throughput peaks once
    // the handles exist (1024 files, handles warm: 16 cores / 1 GPU: 8 workers 0.28 ms per file, 16 workers 0.23-0.24;
    // 24 cores / 2 GPUs: 8 per GPU 0.17, 16 per GPU 0.21; 32 cores / 8 GPUs: 4 per GPU 0.13, 8 per GPU 0.24).  Smaller
    // hosts keep half their cores (creating a handle pins memory, which the driver serialises: tools/alloc_cost.cu).
    const size_t cores = std::max(1u, std::thread::hardware_concurrency());
    size_t workers = std::min<size_t>(8, std::max<size_t>(std::max<size_t>(1, cores / (2 * G)), std::min<size_t>(4, std::max<size_t>(1, cores / G))));
    if (cores >= 16) workers = std::min<size_t>(16, std::max<size_t>(workers, cores / G));
    if (const char *e = getenv("FB2_FILE_WORKERS")) { const long v = atol(e); if (v >= 1 && v <= 64) workers = (size_t)v; }
    bool has_stdin = false;
    for (size_t i = 0; i < n; ++i) if (strcmp(paths[i], "-") == 0) has_stdin = true;   // stdin is consumed in order
    // LPT: files by decreasing size onto the least loaded GPU; each GPU's list stays in that order
    std::vector<std::vector<size_t>> lists(G);
    if (G == 1 || has_stdin) { for (size_t i = 0; i < n; ++i) lists[0].push_back(i); }
    else {
        std::vector<std::pair<uint64_t, size_t>> sized(n);
        for (size_t i = 0; i < n; ++i) {
            struct stat sb;
            sized[i] = {stat(paths[i], &sb) == 0 ? (uint64_t)sb.st_size : 0ull, i};
        }
        std::stable_sort(sized.begin(), sized.end(), [](const std::pair<uint64_t, size_t> &x, const std::pair<uint64_t, size_t> &y) { return x.first > y.first; });
        std::vector<uint64_t> load(G, 0);
        for (const auto &fz : sized) {
            size_t g = 0;
            for (size_t j = 1; j < G; ++j) if (load[j] < load[g]) g = j;
            lists[g].push_back(fz.second);
            load[g] += fz.first + 4096;
        }
    }
    const size_t piece = 32u << 20;
    std::vector<std::atomic<size_t>> next(G);
    for (auto &a : next) a.store(0);
    std::atomic<int> first_rc{FB2_OK};
    std::mutex mu;
    std::string first_msg;
    auto fail = [&](int rc) {
        std::lock_guard<std::mutex> g(mu);
        if (first_rc.load() == FB2_OK) { first_rc.store(rc); first_msg = fb2_last_error(); }
    };
    // more worker threads than cores (only by FB2_FILE_WORKERS, or many GPUs on a small host): waits must not spin
    size_t n_threads = 0;
    for (size_t g = 0; g < G; ++g) n_threads += has_stdin ? 1 : std::min(workers, lists[g].size());
    const bool polite = getenv("FB2_POLITE_SYNC") ? atoi(getenv("FB2_POLITE_SYNC")) != 0 : n_threads > cores;
    auto work = [&](size_t g) {
        fb2_params pd = *p;
        pd.device = devices[g];
        fb2_sketcher *s = nullptr;
        uint8_t *buf = nullptr;
        const uint64_t t_create = now_ns();
        bool reuse = pool_acquire(&pd, &s, &buf);     // an idle handle of an earlier call, same parameters and device
        int rc = FB2_OK;
        if (!reuse) g_n_created += 1;
        if (!reuse) rc = fb2_sketcher_create(&pd, &s);   // selects the device for this thread
        g_ns_create += now_ns() - t_create;
        if (rc == FB2_OK) fb2_sketcher_set_polite_sync(s, polite ? 1 : 0);
        while (rc == FB2_OK && first_rc.load() == FB2_OK) {
            const size_t q = next[g].fetch_add(1);
            if (q >= lists[g].size()) break;
            const size_t i = lists[g][q];
            const size_t need = rbuf_need(paths[i], piece);
            if (need > rbuf_cap(buf)) {                  // (handles pooled by fb2_sketch_stream_multi carry no read buffer)
                cudaSetDevice(pd.device);
                rbuf_free(buf);
                buf = rbuf_alloc(need);
                if (!buf) { rc = fb2_fail(FB2_ECUDA, "cudaHostAlloc failed (file read buffer)"); break; }
            }
            rc = sketch_one_file(s, reuse, paths[i], buf, rbuf_cap(buf), &pd, f, &outs[i]);
            reuse = true;
        }
        if (s) fb2_sketcher_set_polite_sync(s, 0);       // pooled handles go back to spinning
        if (rc != FB2_OK) fail(rc);
        if (rc == FB2_OK && s && buf && pool_release(&pd, s, buf)) return;   // kept for the next call
        rbuf_free(buf);
        if (s) fb2_sketcher_destroy(s);
    };
    std::vector<std::thread> th;
    for (size_t g = 0; g < G; ++g) {
        const size_t w = has_stdin ? 1 : std::min(workers, lists[g].size());
        for (size_t k = 0; k < w; ++k) th.emplace_back(work, g);
    }
    if (th.size() == 1) { th[0].join(); }
    else for (auto &t : th) t.join();
    if (getenv("FB2_TRACE_FILES")) {
        const double nf = (double)std::max<uint64_t>(1, g_n_files.load());
        fprintf(stderr, "sketch_files: %llu files, %zu threads%s; per file (worker time, us): reset %.0f  read %.0f  feed %.0f  finish %.0f; "
                        "%llu handles created, %.1f ms per thread getting one\n",
                (unsigned long long)g_n_files.load(), th.size(), polite ? " (polite waits)" : "", g_ns_reset.load() / nf / 1e3, g_ns_read.load() / nf / 1e3,
                g_ns_feed.load() / nf / 1e3, g_ns_finish.load() / nf / 1e3, (unsigned long long)g_n_created.load(),
                g_ns_create.load() / 1e6 / (double)std::max<size_t>(1, th.size()));
        g_ns_reset = 0; g_ns_read = 0; g_ns_feed = 0; g_ns_finish = 0; g_n_files = 0; g_ns_create = 0; g_n_created = 0;
    }
    const int rc = first_rc.load();
    if (rc != FB2_OK) {
        for (size_t i = 0; i < n; ++i) fb2_result_free(&outs[i]);
        return fb2_fail(rc, first_msg);
    }
    return FB2_OK;
}

static int device_list(const fb2_params *p, int ngpus, std::vector<int> &devices) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fb2_fail(FB2_ECUDA, "no CUDA device: finch_b200 has no CPU fallback");
    devices.clear();
    if (ngpus == 1 || (ngpus <= 0 && ndev == 1) || (ndev == 1 && !getenv("FB2_MULTI_OVERSUBSCRIBE"))) {   // one GPU: the one the parameters name (or the current one)
        int d = p->device;
        if (d < 0 && cudaGetDevice(&d) != cudaSuccess) d = 0;
        if (d >= ndev) return fb2_fail(FB2_EINVAL, "device ordinal out of range");
        devices.push_back(d);
        return FB2_OK;
    }
    // FB2_MULTI_OVERSUBSCRIBE=1 (tests on a one-GPU box): more "GPUs" than devices wrap around onto the visible ones
    const bool wrap = getenv("FB2_MULTI_OVERSUBSCRIBE") != nullptr;
    const int G = ngpus <= 0 ? ndev : (wrap ? ngpus : std::min(ngpus, ndev));
    for (int d = 0; d < G; ++d) devices.push_back(d % ndev);
    return FB2_OK;
}

extern "C" int fb2_sketch_files(const char *const *paths, size_t n, const fb2_params *p, const fb2_filter *f,
                                fb2_result *outs) {
    return fb2_sketch_files_multi(paths, n, p, f, outs, 1);
}
extern "C" int fb2_sketch_files_multi(const char *const *paths, size_t n, const fb2_params *p, const fb2_filter *f,
                                      fb2_result *outs, int ngpus) {
    if (!p || !f || (n && (!paths || !outs))) return fb2_fail(FB2_EINVAL, "null argument");
    std::vector<int> devices;
    const int rc = device_list(p, ngpus, devices);
    if (rc != FB2_OK) return rc;
    return sketch_files_on(paths, n, p, f, outs, devices);
}

// ---- one stream, several GPUs (second mode) ----------------------------------------------------
// sketch_stream (lib.rs:51-94) of ONE file with its bytes cut into one range per GPU: every range is parsed and
// sketched by its own GPU (the host->device copies run on every GPU's own PCIe link), then the tables are united
// exactly on the first GPU by kernels that read the peers' tables over NVLink (fb2_sketcher_merge_from) and the
// usual to_vec / filter / truncate tail runs there.
//   FASTQ: ranges start at record starts ('@' line whose third line starts with '+').  Whether a cut really is a
//          record start is VERIFIED, not assumed: every range but the last must end with its parser back in the
//          header phase right after a newline -- by induction from the first range all cuts are then true record
//          starts.  If not (pathological input), the stream is sketched on one GPU instead.
//   FASTA: ranges start at any line start; the host hands the next GPU the parser state there: whether the previous
//          line was a header, the two previous bytes, and the last `halo` symbols (k-mers span the cut).
static size_t find_fastq_record_start(const uint8_t *b, size_t len, size_t from) {
    size_t p = from;
    while (p < len) {
        const uint8_t *nl = (const uint8_t *)memchr(b + p, '\n', len - p);
        if (!nl) return len;
        const size_t q = (size_t)(nl - b) + 1;
        if (q >= len) return len;
        if (b[q] == '@') {
            const uint8_t *l1 = (const uint8_t *)memchr(b + q, '\n', len - q);
            if (!l1) return len;
            const size_t s1 = (size_t)(l1 - b) + 1;                       // would-be sequence line
            const uint8_t *l2 = s1 < len ? (const uint8_t *)memchr(b + s1, '\n', len - s1) : nullptr;
            if (!l2) return len;
            const size_t s2 = (size_t)(l2 - b) + 1;                       // would-be '+' line
            if (s2 < len && b[s2] == '+' && b[s1] != '@') return q;
        }
        p = q;
    }
    return len;
}
// The parser's carry at line start P (0 < P < len) of a FASTA stream.
static void fasta_carry_at(const uint8_t *b, size_t P, uint32_t halo, uint32_t *state, uint32_t *prev2,
                           std::vector<uint8_t> &syms) {
    syms.assign(halo, fb2::SYM_BREAK);
    *prev2 = P >= 2 ? b[P - 2] : (uint32_t)'\n';
    long idx = (long)halo - 1;
    size_t le = P - 1;                                   // position of the newline that ends the line before P
    bool first = true;
    while (true) {
        const void *r = le ? memrchr(b, '\n', le) : nullptr;
        const size_t ls = r ? (size_t)((const uint8_t *)r - b) + 1 : 0;
        const bool hdr = b[ls] == '>';
        if (first) { *state = hdr ? 1u : 0u; first = false; }
        if (hdr) break;                                  // a header start is a break symbol: nothing before it matters
        bool stop = false;
        for (size_t i = le; i > ls && idx >= 0; --i) {
            const uint8_t cls = fb2::classify_byte(b[i - 1]);
            if (cls <= 3) syms[(size_t)idx--] = cls;
            else if (cls == fb2::CLS_WS || cls == fb2::CLS_CR || cls == fb2::CLS_NL) continue;
            else { stop = true; break; }                 // a non-base byte is a break symbol
        }
        if (stop || idx < 0 || ls == 0) break;
        le = ls - 1;
    }
}

extern "C" int fb2_sketch_stream_multi(const uint8_t *bytes, size_t len, const char *name, const fb2_params *p,
                                       const fb2_filter *f, fb2_result *out, int ngpus) {
    if (!p || !f || !out || (!bytes && len)) return fb2_fail(FB2_EINVAL, "null argument");
    std::vector<int> devices;
    int rc = device_list(p, ngpus, devices);
    if (rc != FB2_OK) return rc;
    size_t min_range = 8u << 20;
    if (const char *e = getenv("FB2_MIN_RANGE_KB")) min_range = (size_t)std::max(1L, atol(e)) << 10;
    const bool fasta = len && bytes[0] == '>', fastq = len && bytes[0] == '@';
    size_t G = std::min(devices.size(), std::max<size_t>(1, len / min_range));
    if (G <= 1 || !(fasta || fastq) || p->kind == FB2_KIND_ALLCOUNTS) {   // (counter arrays are not merged across GPUs)                   // also: unknown / compressed formats get the usual errors
        fb2_params pd = *p;
        pd.device = devices[0];
        return fb2_sketch_stream(bytes, len, name, &pd, f, out);
    }
    memset(out, 0, sizeof(*out));
    // ---- cuts ----
    std::vector<size_t> cut(1, 0);
    for (size_t g = 1; g < G; ++g) {
        const size_t target = len / G * g;
        size_t q;
        if (fastq) q = find_fastq_record_start(bytes, len, std::max(target, cut.back()));
        else {
            const uint8_t *nl = (const uint8_t *)memchr(bytes + target, '\n', len - target);
            q = nl ? (size_t)(nl - bytes) + 1 : len;
        }
        if (q < len && q > cut.back()) cut.push_back(q);
    }
    G = cut.size();
    cut.push_back(len);
    if (G <= 1) { fb2_params pd = *p; pd.device = devices[0]; return fb2_sketch_stream(bytes, len, name, &pd, f, out); }
    // ---- one handle and one host thread per range ----
    std::vector<fb2_sketcher *> hs(G, nullptr);
    std::vector<uint8_t *> bufs(G, nullptr);
    std::vector<int> rcs(G, FB2_OK);
    std::vector<std::string> msgs(G);
    std::vector<uint32_t> end_state(G, 0), last_byte(G, '\n');
    std::vector<uint64_t> bad1(G, ~0ULL), bad2(G, ~0ULL);
    const int fmt = fasta ? FB2_FORMAT_FASTA : FB2_FORMAT_FASTQ;
    auto work = [&](size_t g) {
        fb2_params pd = *p;
        pd.device = devices[g]; pd.stream = nullptr;
        int r = FB2_OK;
        bool pooled = pool_acquire(&pd, &hs[g], &bufs[g]);
        if (pooled) r = fb2_sketcher_reset(hs[g]);
        else r = fb2_sketcher_create(&pd, &hs[g]);
        if (r == FB2_OK && p->kind == FB2_KIND_MASH) fb2_sketcher_hint_finish(hs[g], p->final_size, f->filter_on);
        if (r == FB2_OK) {
            uint32_t state = fasta ? 1u : 0u, prev2 = '\n';
            std::vector<uint8_t> syms;
            if (g > 0 && fasta) fasta_carry_at(bytes, cut[g], fb2_sketcher_halo(hs[g]), &state, &prev2, syms);
            r = fb2_sketcher_begin_range(hs[g], fmt, state, '\n', prev2, syms.empty() ? nullptr : syms.data(), cut[g],
                                         (uint64_t)g << 44, 0);
        }
        if (r == FB2_OK) r = fb2_sketcher_feed_fastx(hs[g], bytes + cut[g], cut[g + 1] - cut[g], g + 1 == G ? 1 : 0);
        if (r == FB2_OK && g + 1 < G) r = fb2_sketcher_end_range(hs[g], &end_state[g], &last_byte[g], &bad1[g], &bad2[g]);
        rcs[g] = r;
        if (r != FB2_OK) msgs[g] = fb2_last_error();
    };
    {
        std::vector<std::thread> th;
        for (size_t g = 0; g < G; ++g) th.emplace_back(work, g);
        for (auto &t : th) t.join();
    }
    auto release_all = [&](bool keep) {
        for (size_t g = 0; g < G; ++g) {
            if (!hs[g]) continue;
            fb2_params pd = *p;
            pd.device = devices[g]; pd.stream = nullptr;
            if (keep && pool_release(&pd, hs[g], bufs[g])) continue;
            rbuf_free(bufs[g]);
            fb2_sketcher_destroy(hs[g]);
        }
    };
    // a cut that is not a record start shows as a range that does not end in the header phase: redo on one GPU
    bool cuts_ok = true;
    if (fastq) for (size_t g = 0; g + 1 < G; ++g) if (rcs[g] == FB2_OK && (end_state[g] != 0u || last_byte[g] != '\n')) cuts_ok = false;
    if (!cuts_ok) {
        release_all(false);
        fb2_params pd = *p;
        pd.device = devices[0];
        return fb2_sketch_stream(bytes, len, name, &pd, f, out);
    }
    for (size_t g = 0; g < G && rc == FB2_OK; ++g) {     // first error in stream order
        if (rcs[g] != FB2_OK) {
            // a blank last range (trailing newlines longer than a range) cannot be judged alone: one GPU decides
            if (g + 1 == G && rcs[g] == FB2_EEMPTY) { release_all(false); fb2_params pd = *p; pd.device = devices[0]; return fb2_sketch_stream(bytes, len, name, &pd, f, out); }
            rc = fb2_fail(rcs[g], msgs[g]);
        } else if (g + 1 < G && fastq && (bad1[g] != ~0ULL || bad2[g] != ~0ULL)) {
            rc = fb2_fail(FB2_ERECORD, bad1[g] <= bad2[g]
                              ? "invalid FASTQ record: line at byte " + std::to_string(bad1[g]) + " does not start with the expected '@' / '+'"
                              : "invalid FASTQ record: sequence and quality lengths differ (record whose header ends at byte " + std::to_string(bad2[g]) + ")");
        }
    }
    for (size_t g = 1; g < G && rc == FB2_OK; ++g) rc = fb2_sketcher_merge_from(hs[0], hs[g]);
    if (rc == FB2_OK) rc = fb2_sketcher_sketch(hs[0], name, p, f, out);
    std::string msg = rc != FB2_OK ? fb2_last_error() : "";
    release_all(rc == FB2_OK);
    return rc == FB2_OK ? FB2_OK : fb2_fail(rc, msg);
}

// ---- one FASTQ stream, host cores AND the PCIe link (FB2_HOST_STRIP=2) ---------------------------------------------
// End to end from host memory a FASTQ stream is bound by the link (2.09 raw bytes per base); with the host pre-strip
// (strip.cpp) it is bound by the host cores instead and the link idles half of the time.  Here both work at once, from
// the two ends of the stream: handle A takes spans from the FRONT and frames their records on the host cores (one
// continuous stream: its spans follow each other), handle B takes spans from the BACK and copies them raw (each one a
// range of its own, begun at a guessed record start and VERIFIED to end on one, like fb2_sketch_stream_multi's cuts).
// They stop where they meet -- whichever resource is faster takes more -- and the two tables are united exactly
// (fb2_sketcher_merge_from; position ids ascend with the stream offset whatever the processing order).  Anything out of
// the ordinary (a cut that is no record start, any record error, blank stretches) sends the stream through the plain
// single-mode path, which produces the authoritative result or error.
static int sketch_stream_two_ended(const uint8_t *bytes, size_t len, const char *name, const fb2_params *p,
                                   const fb2_filter *f, fb2_result *out) {
    fb2_params pd = *p;
    if (pd.device < 0 && cudaGetDevice(&pd.device) != cudaSuccess) pd.device = 0;
    pd.stream = nullptr;
    // cuts at guessed record starts, about `unit` bytes apart
    // (four units stay under the 2 x chunk of raw bytes the host framing takes per block: one block per claim of A)
    const size_t unit = std::max<size_t>(64u << 10, env_size_h("FB2_TWO_ENDED_UNIT_KB", 60u << 10) << 10);
    std::vector<size_t> cut(1, 0);
    for (size_t target = unit; target + unit / 2 < len; target += unit) {
        const size_t q = find_fastq_record_start(bytes, len, std::max(target, cut.back()));
        if (q < len && q > cut.back()) cut.push_back(q);
    }
    const size_t N = cut.size();          // ranges [cut[g], cut[g + 1])
    cut.push_back(len);
    if (N < 4 || N > 1500) return FB2_EUNSUPPORTED;   // too small to be worth it (or position ids would not fit): the plain path
    fb2_sketcher *hs[2] = {nullptr, nullptr};
    uint8_t *bufs[2] = {nullptr, nullptr};
    int rcs[2] = {FB2_OK, FB2_OK};
    bool clean[2] = {true, true};
    std::mutex mu;
    // FB2_TWO_ENDED_PRIORITY=1 (A/B switch; measured no gain on the 16-core box): A's copies in flight, B's pieces wait
    // for them.  Process-wide and never destroyed: the driver lowers it from a stream callback that may run after this
    // call has returned.
    static std::atomic<int> link_flag{0};
    const size_t claim_div = std::max<size_t>(1, env_size_h("FB2_TWO_ENDED_DIV", 3)), back_max = std::max<size_t>(1, env_size_h("FB2_TWO_ENDED_BACK_MAX", 8));
    size_t lo = 0, hi = N;                // unclaimed ranges [lo, hi)
    auto claim = [&](bool front, size_t *a, size_t *b) -> bool {
        std::lock_guard<std::mutex> lk(mu);
        if (lo >= hi) return false;
        const size_t left = hi - lo, take = std::max<size_t>(1, std::min<size_t>(front ? 4 : back_max, left / claim_div));   // smaller claims towards the end
        if (front) { *a = lo; *b = lo + take; lo += take; }
        else { *a = hi - take; *b = hi; hi -= take; }
        return true;
    };
    fb2_stats st0[2];
    memset(st0, 0, sizeof(st0));
    auto open = [&](int w) -> int {
        int r;
        if (pool_acquire(&pd, &hs[w], &bufs[w])) r = fb2_sketcher_reset(hs[w]);
        else r = fb2_sketcher_create(&pd, &hs[w]);
        if (r == FB2_OK) fb2_sketcher_stats(hs[w], &st0[w]);
        if (r == FB2_OK && p->kind == FB2_KIND_MASH) fb2_sketcher_hint_finish(hs[w], p->final_size, f->filter_on);
        return r;
    };
    const char *only = getenv("FB2_TWO_ENDED_ONLY");   // diagnosis: "front" / "back" = the other side takes the minimum
    const bool only_front = only && only[0] == 'f', only_back = only && only[0] == 'b';
    size_t b_first = N;                   // B's first claim is [b_first, N): it holds the end of the stream
    {
        size_t a = N, b = N;
        claim(false, &a, &b);
        b_first = a;
    }
    auto front_work = [&]() {             // A: host-framed, one continuous stream from byte 0
        int r = open(0);
        if (r == FB2_OK && getenv("FB2_TWO_ENDED_PRIORITY")) fb2_sketcher_set_link_flag(hs[0], &link_flag, 1);
        if (r == FB2_OK) r = fb2_sketcher_begin_range(hs[0], FB2_FORMAT_FASTQ, 0u, '\n', '\n', nullptr, 0, 0, 1);
        size_t a, b;
        while (r == FB2_OK && !only_back && claim(true, &a, &b)) r = fb2_sketcher_feed_fastx(hs[0], bytes + cut[a], cut[b] - cut[a], 0);
        if (r == FB2_OK) {
            uint32_t st = 0, lb = '\n';
            uint64_t b1 = ~0ULL, b2 = ~0ULL;
            r = fb2_sketcher_end_range(hs[0], &st, &lb, &b1, &b2);
            if (r == FB2_OK && (st != 0u || b1 != ~0ULL || b2 != ~0ULL)) clean[0] = false;
        }
        rcs[0] = r;
    };
    auto back_work = [&]() {              // B: raw bytes over the link, spans from the back
        int r = open(1);
        if (r == FB2_OK) {
            fb2_sketcher_set_polite_copy(hs[1], (unsigned)env_size_h("FB2_TWO_ENDED_PIECE_MB", 64));
            if (getenv("FB2_TWO_ENDED_PRIORITY")) fb2_sketcher_set_link_flag(hs[1], &link_flag, 0);
        }
        size_t a = b_first, b = N;
        bool have = r == FB2_OK;
        while (have) {
            r = fb2_sketcher_begin_range(hs[1], FB2_FORMAT_FASTQ, 0u, '\n', '\n', nullptr, cut[a], (uint64_t)(a + 1) << 44, 0);
            if (r == FB2_OK) r = fb2_sketcher_feed_fastx(hs[1], bytes + cut[a], cut[b] - cut[a], b == N ? 1 : 0);
            if (r == FB2_OK && b != N) {
                uint32_t st = 0, lb = '\n';
                uint64_t b1 = ~0ULL, b2 = ~0ULL;
                r = fb2_sketcher_end_range(hs[1], &st, &lb, &b1, &b2);
                if (r == FB2_OK && (st != 0u || lb != '\n' || b1 != ~0ULL || b2 != ~0ULL)) clean[1] = false;
            }
            if (r != FB2_OK || !clean[1]) break;
            have = !only_front && claim(false, &a, &b);
        }
        rcs[1] = r;
    };
    const bool trace2 = getenv("FB2_TRACE_TWO_ENDED") != nullptr;
    auto now_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_start = now_ms();
    double t_front = 0, t_back = 0;
    {
        std::thread tb([&] { back_work(); t_back = now_ms(); });
        front_work();
        t_front = now_ms();
        tb.join();
    }
    const double t_joined = now_ms();
    auto release_all = [&](bool keep) {
        for (int w = 0; w < 2; ++w) {
            if (!hs[w]) continue;
            fb2_sketcher_set_polite_copy(hs[w], 0);
            fb2_sketcher_set_link_flag(hs[w], nullptr, 0);
            if (keep && pool_release(&pd, hs[w], bufs[w])) continue;
            rbuf_free(bufs[w]);
            fb2_sketcher_destroy(hs[w]);
        }
    };
    if (rcs[0] != FB2_OK || rcs[1] != FB2_OK || !clean[0] || !clean[1]) {
        release_all(false);
        return FB2_EUNSUPPORTED;          // the plain path decides (and words the error, if there is one)
    }
    int rc = fb2_sketcher_merge_from(hs[0], hs[1]);
    if (rc == FB2_OK) rc = fb2_sketcher_sketch(hs[0], name, p, f, out);
    if (rc == FB2_OK) {
        memset(&g_last_stream_stats, 0, sizeof(g_last_stream_stats));
        for (int w = 0; w < 2; ++w) {
            fb2_stats st1;
            memset(&st1, 0, sizeof(st1));
            fb2_sketcher_stats(hs[w], &st1);
            stats_add_delta(g_last_stream_stats, st0[w], st1);
        }
        if (trace2)
            fprintf(stderr, "two-ended: %zu ranges, host-framed [0, %zu), raw [%zu, %zu); front done +%.1f ms, back done +%.1f ms, "
                            "merge + sketch %.1f ms\n", N, lo, lo, N, t_front - t_start, t_back - t_start, now_ms() - t_joined);
    }
    const std::string msg = rc != FB2_OK ? fb2_last_error() : "";
    release_all(rc == FB2_OK);
    return rc == FB2_OK ? FB2_OK : fb2_fail(rc, msg);
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

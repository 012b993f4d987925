// engine.cu -- the sketcher object behind the C ABI (include/finch_b200.h).
//
// Host-side orchestration only: all arithmetic on sequence data runs in the kernels of
// parse.cu / hash.cu / table.cu.  There is no CPU fallback.
//
// Mirrors, per handle, the state of MashSketcher / ScaledSketcher (mash.rs:10-18, scaled.rs:10-19):
//   heap + counts map  ->  device hash table (TableView) + admission threshold
//   total_kmers        ->  host counter fed by the hash kernel's per-launch count
//   total_bases        ->  ParseCarry::total_bases (FASTX) + host sum (process())
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <sys/mman.h>

#include "common.cuh"
#include "kernels.h"
#include "strip.h"

using namespace fb2;

// ---- error plumbing --------------------------------------------------------------------------
static thread_local std::string g_err;
int fb2_fail(int code, const std::string &msg) { g_err = msg; return code; }
extern "C" const char *fb2_last_error(void) { return g_err.c_str(); }

#define CU(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess)                                                                  \
            return fb2_fail(FB2_ECUDA, std::string(#x) + ": " + cudaGetErrorString(e_));        \
    } while (0)
#define TRY(x) do { int r_ = (x); if (r_ != FB2_OK) return r_; } while (0)

// Every entry point works on its handle's device and leaves the calling thread's current device as it found it
// (a caller that mixes this library with other CUDA code, e.g. torch, must not find its device changed).
struct DeviceScope {
    int prev = -1;
    bool ok = true;
    explicit DeviceScope(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (dev >= 0 && dev != prev) ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceScope() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};
#define ON_DEVICE(dev) DeviceScope dev_scope_(dev); if (!dev_scope_.ok) return fb2_fail(FB2_ECUDA, "cudaSetDevice failed")

static size_t env_size(const char *name, size_t dflt) {
    const char *v = getenv(name);
    if (!v || !*v) return dflt;
    return (size_t)strtoull(v, nullptr, 10);
}
static uint32_t next_pow2(uint64_t v) {
    uint64_t p = 1;
    while (p < v) p <<= 1;
    return (uint32_t)p;
}
static inline uint32_t cdivu(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {  // contents are NOT preserved
        if (bytes <= cap) return FB2_OK;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        CU(cudaMalloc(&p, want));
        cap = want;
        return FB2_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct Table {
    DevBuf key, cnt, ext, posx, kmer;
    uint32_t cap = 0, kw = 1;            // kw: 64-bit words of 2-bit codes per k-mer (1 for k <= 32)
    int alloc(uint32_t c) {
        const size_t n = (size_t)c + 1;  // + side slot for hash == u64::MAX
        TRY(key.ensure(n * 8)); TRY(cnt.ensure(n * 8)); TRY(ext.ensure(n * 8));
        TRY(posx.ensure(n * 8)); TRY(kmer.ensure(n * 8 * kw));
        cap = c;
        return FB2_OK;
    }
    TableView view() const {
        TableView v;
        v.key = key.as<unsigned long long>(); v.cnt = cnt.as<unsigned long long>();
        v.ext = ext.as<unsigned long long>(); v.posx = posx.as<unsigned long long>();
        v.kmer = kmer.as<unsigned long long>();
        v.cap = cap; v.kw = kw;
        uint32_t lg = 0; while ((1u << lg) < cap) ++lg;
        v.shift = 64u - lg;
        return v;
    }
    void release() { key.release(); cnt.release(); ext.release(); posx.release(); kmer.release(); cap = 0; }
};

struct fb2_sketcher {
    fb2_params prm{};
    int device = 0;
    bool scaled = false;
    uint64_t size = 0, max_hash = 0;
    int k = 0;
    uint32_t kw = 1, halo = HALO_SMALL;   // words per k-mer; symbols carried across region / chunk seams
    cudaStream_t st = nullptr, copy_st = nullptr;
    cudaStream_t st2 = nullptr;      // absorb stream: log -> table of chunk c overlaps the parse kernels of chunk c+1
    cudaEvent_t ev_hash[2]{}, ev_st2 = nullptr;
    bool st2_dirty = false;          // work was queued on st2 since the last join
    bool skip_provisional = false;   // hash_range: the provisional threshold failed for this chunk, use the exact ramp
    bool own_stream = false;
    cudaEvent_t ev_h2d[2]{}, ev_rawfree[2]{};
    bool rawfree_pending[2] = {false, false};

    size_t chunk_bytes = 0;
    DevBuf d_raw[2], d_sym[2], d_stmap, d_ststate, d_rcount[2], d_tail, d_carry, d_state, d_seam;
    // fused single-pass parse (parse.cu, parse_fused_kernel): d_stmap holds the look-back status words
    DevBuf d_ptab[2], d_pcount[2];   // hash pieces planned by the parse kernel (PiecePlan), per chunk parity
    bool rec_pieces = true;          // FB2_PIECES=0: uniform 64-position pieces only (A/B)
    bool fresh = false;              // nothing fed since create / reset (fb2_sketcher_sketch_small needs an untouched handle)
    bool small_mode = false; int small_par = 0; uint64_t small_ord = 0;   // run_chunk stops after the parse kernels
    bool allcounts = false;          // FB2_KIND_ALLCOUNTS: d_ac holds counts[4^k] (counts.rs), no table / log / hashing
    DevBuf d_ac, d_ac_off, d_ac_meta;
    DevBuf d_fhist, d_fok;           // device-side sketch filters: histogram of counts (+ 4 meta words), strand flags
    DevBuf d_ticket;                 // supertile ticket counter (never reset: launches pass its value so far)
    uint32_t ticket_total = 0;
    uint32_t parse_epoch = 0;        // 1..65535, one per chunk; the status words are cleared when it wraps
    bool parse_fused = true;         // FB2_PARSE_V1=1: the three-kernel pipeline (phase / scan / pack)
    int tail_sel = 0;               // which half of d_tail holds the symbols carried into the next chunk
    uint64_t ordinal = 0;           // next position id (symbols of all regions so far + pushed k-mers)
    DevBuf log_hash[2], log_kmer[2], log_posx[2];   // one candidate log per in-flight chunk
    uint32_t log_cap = 0;
    int par = 0;                     // parity (slot / log / symbol buffer) of the next chunk
    struct Pending { bool valid = false; ChunkGeom g{}; uint64_t ord_base = 0; } pend[2];
    cudaEvent_t ev_chunk[2]{};
    SketchState *h_snap[2] = {nullptr, nullptr};   // pinned snapshots written after each async chunk
    bool steady = false;             // a whole chunk's candidates fit the log: chunks run asynchronously
    ChunkGeom last_geom{};
    std::vector<uint8_t> tail_host;   // last <= 4096 raw bytes of the open stream (end-of-stream checks)
    Table tab[2];
    int cur = 0;
    DevBuf sort_keys, sort_slots, sort_tkeys, sort_tslots, sort_hist, d_bins;
    DevBuf d_live_bins;          // live histogram of the current table's keys (table_upsert), 4096 bins
    uint32_t live_shift = 52;    // its shift (host copy of SketchState::hist_shift)
    DevBuf out_hash, out_cnt, out_ext, out_kmer, out_posx;
    DevBuf sel_hash, sel_cnt, sel_ext, sel_kmer, sel_posx, sel_bytes, sel_idx;
    std::vector<cudaEvent_t> ev_piece;   // one per piece of a large result on its way out (collect_rows)
    uint8_t *h_res = nullptr;        // pinned read-back staging
    size_t h_res_cap = 0;
    DevBuf d_push_bytes, d_push_offs, d_push_extra;

    uint8_t *h_pinned = nullptr;    // one pinned allocation holding the mirrors below and h_snap[]
    ParseCarry *h_carry = nullptr;  // pinned mirrors
    SketchState *h_state = nullptr;

    uint8_t *h_stage = nullptr;     // pinned staging for process() and small feed pieces
    size_t stage_cap = 0, stage_fill = 0;
    int stage_mode = -1;

    std::vector<uint8_t> push_bytes, push_extra;
    std::vector<uint32_t> push_offs;
    std::vector<std::string> arena;  // bytes of pushed k-mers (unit-test surface)
    size_t arena_flushed = 0;

    int format = FB2_FORMAT_UNKNOWN;
    bool stream_open = false;        // a FASTX stream has begun and not yet seen `final`
    // FB2_HOST_STRIP=1: FASTQ record framing on the host, only the sequence lines cross PCIe (strip.cpp)
    bool strip_on = false;               // active for the open stream
    bool polite = false;                 // waits yield / sleep instead of spinning (fb2_sketcher_set_polite_sync)
    bool force_strip = false;            // the next FASTQ stream is framed on the host whatever FB2_HOST_STRIP says (mapped files)
    unsigned polite_copy = 0;            // > 0: raw host-to-device copies go in pieces of that many MiB, one in flight (see feed_host_chunks)
    std::atomic<int> *link_flag = nullptr;   // two-ended streams: copies of the host-framed handle in flight (it raises the flag,
    bool link_owner = false;                 // the raw handle waits for zero before every piece: the framed lines go first)
    bool range_mode = false;             // the open stream is a range of a larger one (begin_range): several feed calls, no sync between them
    unsigned strip_threads = 1;
    std::vector<uint8_t> strip_carry;    // the incomplete record at the end of the previous piece
    uint64_t strip_off = 0;              // stream offset of the next unprocessed byte (of strip_carry[0] when it is not empty)
    uint8_t *strip_pin[2] = {nullptr, nullptr};   // two staging sets of strip_threads buffers each
    size_t strip_pin_each = 0;           // capacity of one thread's buffer
    cudaEvent_t ev_strip_free[2]{};
    bool strip_free_pending[2] = {false, false};
    int strip_set = 0;
    uint64_t strip_bad = ~0ULL, strip_len_bad = ~0ULL, strip_first_blank = ~0ULL, strip_last_nonblank = 0, strip_records = 0;
    std::vector<uint8_t> presniff;   // a first piece shorter than 2 bytes waits here until the format can be sniffed
    // fb2_sketcher_hint_finish: the caller will finish with fb2_sketcher_sketch(final_size, filter) and nothing else,
    // so a Mash heap larger than final_size is only needed when the filter ends up on
    bool hint_set = false; uint64_t hint_final = 0; int hint_filter = 0;
    uint64_t lines_bases = 0, total_kmers = 0;
    uint32_t next_launch = 0;
    fb2_stats stats{};
    bool timing = false;
    // kernel timing (fb2_sketcher_enable_timing): event pairs recorded around the hash / parse launches of
    // whichever path runs (the asynchronous path stays asynchronous); resolved lazily by fb2_sketcher_stats
    struct EvPair { cudaEvent_t a = nullptr, b = nullptr; };
    std::vector<EvPair> ev_free, ev_hash_pending, ev_parse_pending;
};

// Host waits.  A worker thread normally spins in cudaStreamSynchronize (lowest latency); when fb2_sketch_files runs more
// worker threads than the host has cores (FB2_FILE_WORKERS, many GPUs on few cores) a spinning wait takes the core a
// sibling needs for its file read, and throughput collapsed (8 GPUs x 16 workers on 32 cores: 10x slower than x 4).
// `polite` handles poll a few times and then sleep between polls.
static cudaError_t polite_wait(const std::function<cudaError_t()> &query) {
    for (int n = 0;; ++n) {
        const cudaError_t e = query();
        if (e != cudaErrorNotReady) return e;
        if (n < 16) std::this_thread::yield();
        else std::this_thread::sleep_for(std::chrono::microseconds(n < 64 ? 20 : 100));
    }
}
static cudaError_t wait_stream(const fb2_sketcher *s, cudaStream_t st) {
    if (!s->polite) return cudaStreamSynchronize(st);
    return polite_wait([st] { return cudaStreamQuery(st); });
}
static cudaError_t wait_event(const fb2_sketcher *s, cudaEvent_t ev) {
    if (!s->polite) return cudaEventSynchronize(ev);
    return polite_wait([ev] { return cudaEventQuery(ev); });
}

static bool timing_begin(fb2_sketcher *s, fb2_sketcher::EvPair &p) {
    if (!s->timing) return false;
    if (!s->ev_free.empty()) { p = s->ev_free.back(); s->ev_free.pop_back(); }
    else if (cudaEventCreate(&p.a) != cudaSuccess || cudaEventCreate(&p.b) != cudaSuccess) return false;
    cudaEventRecord(p.a, s->st);
    return true;
}
static void timing_end(fb2_sketcher *s, fb2_sketcher::EvPair &p, std::vector<fb2_sketcher::EvPair> &pending) {
    cudaEventRecord(p.b, s->st);
    pending.push_back(p);
}
static void timing_resolve(fb2_sketcher *s) {
    if (s->ev_hash_pending.empty() && s->ev_parse_pending.empty()) return;
    wait_stream(s, s->st);
    for (auto &p : s->ev_hash_pending) { float ms = 0; if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) s->stats.hash_kernel_ms += ms; s->ev_free.push_back(p); }
    for (auto &p : s->ev_parse_pending) { float ms = 0; if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) s->stats.parse_kernel_ms += ms; s->ev_free.push_back(p); }
    s->ev_hash_pending.clear(); s->ev_parse_pending.clear();
}

// ---- small helpers ---------------------------------------------------------------------------
static LogView log_view(fb2_sketcher *s, int par) {
    LogView l;
    l.hash = s->log_hash[par].as<unsigned long long>(); l.kmer = s->log_kmer[par].as<unsigned long long>();
    l.posx = s->log_posx[par].as<unsigned long long>(); l.cap = s->log_cap; l.kw = s->kw;
    return l;
}
static inline LaunchSlot *dev_slot(fb2_sketcher *s, int par) { return &((SketchState *)s->d_state.p)->slot[par]; }
// Everything queued on the absorb stream so far is ordered before whatever follows on the main stream.
static int join_absorb_stream(fb2_sketcher *s) {
    if (!s->st2_dirty) return FB2_OK;
    CU(cudaEventRecord(s->ev_st2, s->st2));
    CU(cudaStreamWaitEvent(s->st, s->ev_st2, 0));
    s->st2_dirty = false;
    return FB2_OK;
}
static int pull_state(fb2_sketcher *s) {  // device -> pinned mirrors, then wait
    TRY(join_absorb_stream(s));           // every host decision starts here: the table must be quiescent
    CU(cudaMemcpyAsync(s->h_state, s->d_state.p, sizeof(SketchState), cudaMemcpyDeviceToHost, s->st));
    CU(cudaMemcpyAsync(s->h_carry, s->d_carry.p, sizeof(ParseCarry), cudaMemcpyDeviceToHost, s->st));
    CU(wait_stream(s, s->st));
    s->stats.d2h_bytes += sizeof(SketchState) + sizeof(ParseCarry);
    return FB2_OK;
}
static int push_carry(fb2_sketcher *s) {
    CU(cudaMemcpyAsync(s->d_carry.p, s->h_carry, sizeof(ParseCarry), cudaMemcpyHostToDevice, s->st));
    CU(wait_stream(s, s->st));
    return FB2_OK;
}
static int push_state(fb2_sketcher *s) {
    CU(cudaMemcpyAsync(s->d_state.p, s->h_state, sizeof(SketchState), cudaMemcpyHostToDevice, s->st));
    CU(wait_stream(s, s->st));
    return FB2_OK;
}

static int ensure_table(fb2_sketcher *s, int which, uint32_t cap) {
    TRY(s->tab[which].alloc(cap));
    return FB2_OK;
}
static int ensure_sort(fb2_sketcher *s, uint32_t n) {
    const size_t m = (size_t)n + 64;
    TRY(s->sort_keys.ensure(m * 8)); TRY(s->sort_tkeys.ensure(m * 8));
    TRY(s->sort_slots.ensure(m * 4)); TRY(s->sort_tslots.ensure(m * 4));
    TRY(s->sort_hist.ensure((size_t)radix_hist_words(n) * 4));
    return FB2_OK;
}

static uint32_t shift_for_threshold(unsigned long long thr) {   // top 12 bits below the threshold
    uint32_t bits = 0;
    while (bits < 64 && (thr >> bits) != 0ULL) ++bits;
    return bits > 12 ? bits - 12 : 0;
}
// (Re)build the live histogram for the current table; called after every rebuild and at reset.
static int refresh_live_hist(fb2_sketcher *s) {
    s->live_shift = shift_for_threshold(s->h_state->threshold);
    s->h_state->hist_shift = s->live_shift;   // keep the host mirror in step (push_state copies it back)
    launch_live_hist_refresh(s->tab[s->cur].view(), (SketchState *)s->d_state.p, s->live_shift,
                             s->d_live_bins.as<uint32_t>(), s->st);
    s->stats.kernel_launches += 2;
    return FB2_OK;
}

// Candidate-log capacity for the current sketch size.  The provisional first launch lets ~16 * size candidates
// through and wants them in half the log; steady Scaled chunks log chunk_symbols * max_hash / 2^64 candidates and
// the asynchronous path wants a chunk's candidates in a quarter of the log.  FB2_LOG_M (Mi entries) overrides.
static void size_logs(fb2_sketcher *s) {
    uint64_t cap;
    if (const size_t m = env_size("FB2_LOG_M", 0)) cap = (uint64_t)m << 20;
    else {
        uint64_t want = 36ull * s->size;
        if (s->scaled) want = std::max<uint64_t>(want, (uint64_t)(8.0 * (double)s->chunk_bytes * ((double)s->max_hash / 18446744073709551616.0)));
        cap = next_pow2(std::min<uint64_t>(want, 8ull << 20));
        if (cap > (8u << 20)) cap = 8u << 20;
    }
    if (cap < 32ull * HASH_TILE) cap = 32ull * HASH_TILE;
    s->log_cap = (uint32_t)cap;
    // First launch of a stream: the threshold is still infinite, every k-mer is a candidate.  Take as
    // many positions as the log holds (a warp reserves 35 slots per 32 candidates) so that a small
    // file is hashed in ONE launch and the banded absorb picks the bottom band from the whole log.
    s->next_launch = std::max<uint32_t>(32u * HASH_TILE, (s->log_cap / 8u * 7u) / HASH_TILE * HASH_TILE);
}
static void apply_finish_hint(fb2_sketcher *s);
// The logs are allocated at first use (a handle that only ever sees small sketches stays small).
static int ensure_logs(fb2_sketcher *s) {
    for (int i = 0; i < 2; ++i) {
        TRY(s->log_hash[i].ensure((size_t)s->log_cap * 8));
        TRY(s->log_kmer[i].ensure((size_t)s->log_cap * 8 * s->kw));
        TRY(s->log_posx[i].ensure((size_t)s->log_cap * 8));
    }
    return FB2_OK;
}

static int reset_sketch_state(fb2_sketcher *s) {
    s->size = s->allcounts ? 0 : s->prm.kmers_to_sketch;   // a finish hint may have lowered it for the previous stream
    if (s->allcounts) CU(cudaMemsetAsync(s->d_ac.p, 0, ((size_t)1 << (2 * s->k)) * 4, s->st));
    memset(s->h_state, 0, sizeof(SketchState));
    s->h_state->threshold = (s->scaled && s->size == 0) ? s->max_hash : ~0ULL;
    TRY(s->d_live_bins.ensure(4096 * sizeof(uint32_t)));
    s->h_state->live_bins = s->d_live_bins.as<unsigned int>();
    s->live_shift = shift_for_threshold(s->h_state->threshold);
    s->h_state->hist_shift = s->live_shift;
    CU(cudaMemsetAsync(s->d_live_bins.p, 0, 4096 * sizeof(uint32_t), s->st));
    TRY(push_state(s));
    memset(s->h_carry, 0, sizeof(ParseCarry));
    s->h_carry->prev1 = s->h_carry->prev2 = '\n';
    s->h_carry->first_bad_pos = ~0ULL; s->h_carry->len_bad_pos = ~0ULL;
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 3; ++b) s->h_carry->last_nl[a][b] = NL_NONE;
    TRY(push_carry(s));
    launch_table_clear(s->tab[s->cur].view(), s->st);
    launch_fill_bytes(s->d_tail.as<uint8_t>(), 2 * HALO_BIG, SYM_BREAK, s->st);
    s->stats.kernel_launches += 2;
    CU(wait_stream(s, s->st));
    s->tail_sel = 0; s->ordinal = 0; s->par = 0; s->steady = false;
    s->pend[0].valid = s->pend[1].valid = false;
    s->format = FB2_FORMAT_UNKNOWN; s->stream_open = false;
    s->lines_bases = 0; s->total_kmers = 0; s->stage_fill = 0; s->stage_mode = -1;
    size_logs(s);
    s->presniff.clear();
    s->push_bytes.clear(); s->push_extra.clear(); s->push_offs.assign(1, 0u);
    s->arena.clear(); s->arena_flushed = 0;
    s->fresh = true; s->small_mode = false;
    return FB2_OK;
}

// ---- create / destroy --------------------------------------------------------------------------
extern "C" int fb2_sketcher_create(const fb2_params *p, fb2_sketcher **out) {
    if (!p || !out) return fb2_fail(FB2_EINVAL, "null argument");
    *out = nullptr;
    if (p->kind != FB2_KIND_MASH && p->kind != FB2_KIND_SCALED && p->kind != FB2_KIND_ALLCOUNTS) return fb2_fail(FB2_EINVAL, "unknown sketch kind");
    if (p->kmer_length < 1) return fb2_fail(FB2_EINVAL, "kmer_length must be in 1..=255");
    // AllCountsSketcher::new allocates 4^k counters (counts.rs:15-21): 16 GiB at k = 16 is where this build stops
    if (p->kind == FB2_KIND_ALLCOUNTS && p->kmer_length > 16) return fb2_fail(FB2_ENOMEM, "AllCounts: 4^k counters do not fit for k > 16");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fb2_fail(FB2_ECUDA, "no CUDA device: finch_b200 has no CPU fallback");
    int dev = p->device;
    if (dev < 0) { CU(cudaGetDevice(&dev)); }
    if (dev >= ndev) return fb2_fail(FB2_EINVAL, "device ordinal out of range");
    ON_DEVICE(dev);

    fb2_sketcher *s = new fb2_sketcher();
    s->prm = *p; s->device = dev; s->k = p->kmer_length;
    // k <= 32: the k-mer is one 64-bit word of 2-bit codes and the fast kernels apply; 33..=255: exact multi-word path
    s->kw = s->k <= 32 ? 1u : (uint32_t)(s->k + 31) / 32u;
    s->halo = s->k <= 32 ? HALO_SMALL : HALO_BIG;
    s->tab[0].kw = s->tab[1].kw = s->kw;
    s->scaled = p->kind == FB2_KIND_SCALED;
    s->allcounts = p->kind == FB2_KIND_ALLCOUNTS;
    s->size = s->allcounts ? 0 : p->kmers_to_sketch;
    if (s->scaled) {
        // scaled.rs:23,31: iscale = (1./scale) as u64 (saturating); max_hash = u64::MAX / iscale
        const double inv = 1.0 / p->scale;
        uint64_t iscale;
        if (!(inv == inv) || inv <= 0.0) iscale = 0;
        else if (inv >= 18446744073709551616.0) iscale = UINT64_MAX;
        else iscale = (uint64_t)inv;
        if (iscale == 0) { delete s; return fb2_fail(FB2_EINVAL, "scale must be in (0, 1]"); }
        s->max_hash = UINT64_MAX / iscale;
    }
    if (p->stream) { s->st = (cudaStream_t)p->stream; s->own_stream = false; }
    else { CU(cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking)); s->own_stream = true; }
    CU(cudaStreamCreateWithFlags(&s->copy_st, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&s->st2, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&s->ev_st2, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) CU(cudaEventCreateWithFlags(&s->ev_hash[i], cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
        CU(cudaEventCreateWithFlags(&s->ev_h2d[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&s->ev_rawfree[i], cudaEventDisableTiming));
    }

    s->chunk_bytes = env_size("FB2_CHUNK_MB", 128) << 20;
    s->parse_fused = !getenv("FB2_PARSE_V1");
    s->rec_pieces = !(getenv("FB2_PIECES") && atoi(getenv("FB2_PIECES")) == 0);
    if (s->chunk_bytes < (1u << 20)) s->chunk_bytes = 1u << 20;
    if (s->chunk_bytes > (1ull << 30)) s->chunk_bytes = 1ull << 30;

    int rc = FB2_OK;
    do {
        if ((rc = s->d_carry.ensure(sizeof(ParseCarry))) != FB2_OK) break;
        if ((rc = s->d_state.ensure(sizeof(SketchState))) != FB2_OK) break;
        if ((rc = s->d_tail.ensure(2 * HALO_BIG)) != FB2_OK) break;
        {   // the pinned mirrors in ONE allocation (pinning calls are serialised by the driver and queue behind each
            // other when sketch_files starts its workers: tools/alloc_cost.cu)
            const size_t a_carry = (sizeof(ParseCarry) + 255) & ~(size_t)255, a_state = (sizeof(SketchState) + 255) & ~(size_t)255;
            if (cudaHostAlloc((void **)&s->h_pinned, a_carry + 3 * a_state, cudaHostAllocDefault) != cudaSuccess) {
                rc = fb2_fail(FB2_ECUDA, "cudaHostAlloc failed"); break;
            }
            s->h_carry = reinterpret_cast<ParseCarry *>(s->h_pinned);
            s->h_state = reinterpret_cast<SketchState *>(s->h_pinned + a_carry);
            for (int i = 0; i < 2; ++i) s->h_snap[i] = reinterpret_cast<SketchState *>(s->h_pinned + a_carry + (size_t)(1 + i) * a_state);
        }
        for (int i = 0; i < 2 && rc == FB2_OK; ++i)
            if (cudaEventCreateWithFlags(&s->ev_chunk[i], cudaEventDisableTiming) != cudaSuccess) rc = fb2_fail(FB2_ECUDA, "cudaEventCreate failed");
        if (rc != FB2_OK) break;
        // 4 slots per kept key (FB2_TABLE_MULT: A/B switch; 8 measured no faster on the C2 workload)
        uint64_t want = s->size ? env_size("FB2_TABLE_MULT", 4) * s->size : 1;
        if (want < (1u << 16)) want = 1u << 16;
        if (want > (1u << 22)) want = 1u << 22;  // grows on demand
        if ((rc = ensure_table(s, 0, next_pow2(want))) != FB2_OK) break;
        s->cur = 0;
        if (s->allcounts && (rc = s->d_ac.ensure(((size_t)1 << (2 * s->k)) * 4)) != FB2_OK) break;
        if ((rc = reset_sketch_state(s)) != FB2_OK) break;
    } while (0);
    if (rc != FB2_OK) { fb2_sketcher_destroy(s); return rc; }
    *out = s;
    return FB2_OK;
}

extern "C" void fb2_sketcher_destroy(fb2_sketcher *s) {
    if (!s) return;
    DeviceScope dev_scope_(s->device);
    if (s->st) wait_stream(s, s->st);
    if (s->copy_st) wait_stream(s, s->copy_st);
    if (s->st2) wait_stream(s, s->st2);
    for (int i = 0; i < 2; ++i) {
        s->d_raw[i].release(); s->tab[i].release();
        if (s->ev_h2d[i]) cudaEventDestroy(s->ev_h2d[i]);
        if (s->ev_rawfree[i]) cudaEventDestroy(s->ev_rawfree[i]);
    }
    timing_resolve(s);
    for (auto &p : s->ev_free) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (int i = 0; i < 2; ++i) {
        s->d_sym[i].release(); s->d_rcount[i].release(); s->d_ptab[i].release(); s->d_pcount[i].release(); s->log_hash[i].release(); s->log_kmer[i].release(); s->log_posx[i].release();
        if (s->ev_chunk[i]) cudaEventDestroy(s->ev_chunk[i]);
    }
    s->d_stmap.release(); s->d_ststate.release(); s->d_tail.release(); s->d_seam.release(); s->d_ticket.release(); s->d_fhist.release(); s->d_fok.release(); s->d_ac.release(); s->d_ac_off.release(); s->d_ac_meta.release();
    s->d_carry.release(); s->d_state.release();
    s->sort_keys.release(); s->sort_slots.release(); s->sort_tkeys.release(); s->sort_tslots.release(); s->sort_hist.release(); s->d_bins.release(); s->d_live_bins.release();
    s->out_hash.release(); s->out_cnt.release(); s->out_ext.release(); s->out_kmer.release(); s->out_posx.release();
    s->sel_hash.release(); s->sel_cnt.release(); s->sel_ext.release(); s->sel_kmer.release(); s->sel_posx.release();
    s->sel_bytes.release(); s->sel_idx.release();
    if (s->h_res) cudaFreeHost(s->h_res);
    for (cudaEvent_t e : s->ev_piece) cudaEventDestroy(e);
    s->d_push_bytes.release(); s->d_push_offs.release(); s->d_push_extra.release();
    if (s->h_pinned) cudaFreeHost(s->h_pinned);   // h_carry, h_state, h_snap[]
    if (s->h_stage) cudaFreeHost(s->h_stage);
    for (int i = 0; i < 2; ++i) { if (s->strip_pin[i]) cudaFreeHost(s->strip_pin[i]); if (s->ev_strip_free[i]) cudaEventDestroy(s->ev_strip_free[i]); }
    if (s->own_stream && s->st) cudaStreamDestroy(s->st);
    if (s->copy_st) cudaStreamDestroy(s->copy_st);
    if (s->st2) cudaStreamDestroy(s->st2);
    if (s->ev_st2) cudaEventDestroy(s->ev_st2);
    for (int i = 0; i < 2; ++i) if (s->ev_hash[i]) cudaEventDestroy(s->ev_hash[i]);
    delete s;
}

extern "C" int fb2_sketcher_reset(fb2_sketcher *s) {
    if (!s) return fb2_fail(FB2_EINVAL, "null handle");
    ON_DEVICE(s->device);
    CU(wait_stream(s, s->st));
    CU(wait_stream(s, s->copy_st));
    CU(wait_stream(s, s->st2));
    s->st2_dirty = false;
    return reset_sketch_state(s);
}

// ---- pruning: bottom-s selection + table rebuild -------------------------------------------------
// Sort the occupied keys, decide how many the sketch keeps (select_keep_kernel), rebuild the other
// table from those and lower the threshold.  Grows the table when what must be kept needs it.
static int sort_table(fb2_sketcher *s, uint32_t *n_out) {
    TRY(pull_state(s));
    const uint32_t n_all = s->h_state->occupied + (s->h_state->has_max_key ? 1u : 0u);
    TRY(ensure_sort(s, std::max(n_all, 1u)));
    TRY(s->d_bins.ensure(3 * 4096 * sizeof(uint32_t)));
    SketchState *dst = (SketchState *)s->d_state.p;
    unsigned long long *keys = s->sort_keys.as<unsigned long long>(), *tkeys = s->sort_tkeys.as<unsigned long long>();
    uint32_t *slots = s->sort_slots.as<uint32_t>(), *tslots = s->sort_tslots.as<uint32_t>();
    launch_gather(s->tab[s->cur].view(), dst, tkeys, tslots, s->st);          // unsorted live keys -> (tkeys, tslots)
    s->stats.kernel_launches += 2;
    TRY(pull_state(s));
    const uint32_t n = s->h_state->gather_count;                               // keys at or below the threshold
    // bucket + rank sort on the top 12 bits below the threshold; keys above the threshold cannot
    // exist in the table except the u64::MAX side slot (threshold == MAX then)
    const unsigned long long thr = s->h_state->threshold;
    uint32_t bits = 0;
    while (bits < 64 && (thr >> bits) != 0ULL) ++bits;
    const uint32_t shift = bits > 12 ? bits - 12 : 0;
    if (n >= 2) {
        uint32_t *bins = s->d_bins.as<uint32_t>();
        // scatter target: (keys, slots) as scratch is not possible (rank needs a third buffer): use
        // out_hash/out_cnt-sized scratch from the sort pool: tkeys -> keys (scatter) -> tkeys (rank)
        launch_bucket_sort(tkeys, tslots, keys, slots, tkeys, tslots, n, shift, bins, bins + 4096, bins + 8192, dst, s->st);
        s->stats.kernel_launches += 4;
        TRY(pull_state(s));
        if (s->h_state->gather_count <= bucket_cap()) {
            // result is in (tkeys, tslots): copy to (keys, slots) where the consumers expect it
            CU(cudaMemcpyAsync(keys, tkeys, (size_t)n * 8, cudaMemcpyDeviceToDevice, s->st));
            CU(cudaMemcpyAsync(slots, tslots, (size_t)n * 4, cudaMemcpyDeviceToDevice, s->st));
        } else {
            // non-uniform keys: (keys, slots) hold the bucket-scattered (complete) set: radix sort them
            launch_radix_sort(keys, slots, tkeys, tslots, n, s->sort_hist.as<uint32_t>(), s->st);
            s->stats.kernel_launches += 24;
        }
    } else if (n == 1) {
        CU(cudaMemcpyAsync(keys, tkeys, 8, cudaMemcpyDeviceToDevice, s->st));
        CU(cudaMemcpyAsync(slots, tslots, 4, cudaMemcpyDeviceToDevice, s->st));
    }
    launch_select_keep(keys, n, s->scaled ? 1 : 0, s->size, s->max_hash, dst, s->st);
    s->stats.kernel_launches += 1;
    TRY(pull_state(s));
    *n_out = n;
    return FB2_OK;
}
// Histogram-select a bin-boundary threshold that keeps >= size keys, gather the survivors,
// rebuild them into the other table (grown when needed) and commit the lower threshold.
// With `limit`: only commit when the selected threshold is <= *limit (i.e. the table already holds
// >= size keys at or below it); otherwise leave everything untouched and report *ok = false.
static int prune(fb2_sketcher *s, uint32_t need_room, const unsigned long long *limit = nullptr, bool *ok = nullptr) {
    if (ok) *ok = true;
    const uint32_t occ_ub = s->tab[s->cur].cap + 1;
    TRY(ensure_sort(s, occ_ub));
    TRY(s->d_bins.ensure(4096 * sizeof(uint32_t)));
    SketchState *dst = (SketchState *)s->d_state.p;
    const unsigned long long thr = s->h_state->threshold;  // mirror may be stale-high: still a valid bound
    uint32_t bits = 0;
    while (bits < 64 && (thr >> bits) != 0ULL) ++bits;
    const uint32_t shift = bits > 12 ? bits - 12 : 0;
    launch_prune_select(s->tab[s->cur].view(), dst, shift, s->d_bins.as<uint32_t>(), s->scaled ? 1 : 0,
                        (!s->scaled && s->size == 0) ? 0ULL : s->size, s->max_hash,
                        s->sort_keys.as<unsigned long long>(), s->sort_slots.as<uint32_t>(), s->st);
    s->stats.kernel_launches += 4;
    TRY(pull_state(s));
    if (limit && s->h_state->new_threshold > *limit) { if (ok) *ok = false; return FB2_OK; }
    uint32_t keep = s->h_state->gather_count;
    if (!s->scaled && s->size == 0) keep = 0;  // MashSketcher::new(0, ..) keeps nothing
    uint32_t cap = s->tab[s->cur].cap;
    while ((uint64_t)keep + need_room > (uint64_t)cap / 4 * 3 || (uint64_t)keep * 2 > cap) {
        if (cap >= (1u << 30)) return fb2_fail(FB2_ENOMEM, "sketch table would exceed 2^30 slots");
        cap <<= 1;
    }
    const int other = s->cur ^ 1;
    TRY(ensure_table(s, other, cap));
    launch_rebuild(s->sort_keys.as<unsigned long long>(), s->sort_slots.as<uint32_t>(), keep,
                   s->tab[s->cur].view(), s->tab[other].view(), dst, s->st);
    launch_commit_threshold(dst, s->st);
    s->stats.kernel_launches += 4;
    s->cur = other;
    s->stats.prunes++;
    TRY(pull_state(s));
    TRY(refresh_live_hist(s));
    return FB2_OK;
}

// Absorb log[0, cnt) into the table, pruning / growing so the table never passes 3/4 load.
// Between pulls the host only knows an upper bound of the occupancy (every absorbed entry may be
// a new key); it fetches the exact value before deciding to prune.
// Large logs (threshold still falling): most entries would be pruned right after being inserted.
// Histogram the log's hashes, absorb only the lowest band that should already contain `size` distinct
// keys, and let the table histogram confirm it (the selected threshold must not exceed the band);
// widen the band only when it does not.  Entries above the final band never touch the table.
static int absorb_log_banded(fb2_sketcher *s, int par, uint32_t cnt, bool *done) {
    *done = false;
    SketchState *dst = (SketchState *)s->d_state.p;
    TRY(pull_state(s));
    const unsigned long long thr = s->h_state->threshold;
    uint32_t bits = 0;
    while (bits < 64 && (thr >> bits) != 0ULL) ++bits;
    const uint32_t shift = bits > 12 ? bits - 12 : 0;
    TRY(s->d_bins.ensure(4096 * sizeof(uint32_t)));
    launch_log_hist(log_view(s, par), cnt, dst, shift, s->d_bins.as<uint32_t>(), s->st);
    s->stats.kernel_launches++;
    std::vector<uint32_t> bins(4096);
    CU(cudaMemcpyAsync(bins.data(), s->d_bins.p, 4096 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->st));
    CU(wait_stream(s, s->st));
    s->stats.d2h_bytes += 4096 * sizeof(uint32_t);
    std::vector<uint64_t> cum(4096);
    uint64_t run = 0;
    for (int b = 0; b < 4096; ++b) { run += bins[b]; cum[b] = run; }
    uint64_t target = s->size + s->size / 4 + 1024;
    unsigned long long lo = 0; int use_lo = 0; uint64_t lo_cum = 0;
    while (true) {
        int b = 0;
        while (b < 4095 && cum[b] < target) ++b;
        unsigned long long hi = thr;
        if (b < 4095 && shift < 64) {
            const unsigned long long top = (((unsigned long long)b + 1ULL) << shift) - 1ULL;
            if (top < thr) hi = top;
        }
        if (s->scaled && hi < s->max_hash) {      // everything <= max_hash is kept: the band must reach it
            hi = std::min<unsigned long long>(thr, s->max_hash);
            b = (int)std::min<unsigned long long>(4095ULL, hi >> shift);
        }
        const uint64_t n_band = cum[b] - lo_cum;  // upper bound of the entries inside (lo, hi]
        // room for n_band new keys (prune / grow first when the table is too full)
        {
            const uint32_t cap = s->tab[s->cur].cap, limit = cap / 4 * 3;
            if ((uint64_t)s->h_state->occupied + n_band > limit) TRY(prune(s, (uint32_t)std::min<uint64_t>(n_band, 1u << 30)));
        }
        launch_absorb_band(log_view(s, par), cnt, s->tab[s->cur].view(), dst, lo, use_lo, hi, s->st);
        s->stats.kernel_launches += 2; s->stats.band_passes++;
        if (getenv("FB2_TRACE_BANDS")) fprintf(stderr, "band pass: cnt=%u target=%llu b=%d n_band=%llu occupied=%u lo_cum=%llu hi=%llu thr=%llu\n", cnt, (unsigned long long)target, b, (unsigned long long)n_band, s->h_state->occupied, (unsigned long long)lo_cum, hi, thr);
        if (hi >= thr) break;                     // the whole log has been absorbed
        bool ok = false;
        TRY(prune(s, 0, &hi, &ok));               // pulls the state; commits only if >= size keys are <= hi
        if (ok) { *done = true; return FB2_OK; }
        lo = hi; use_lo = 1; lo_cum = cum[b];
        // widen: extrapolate from the multiplicity seen so far (entries absorbed per distinct key in the
        // table), at least doubling -- high-coverage reads repeat every key several times per chunk
        const double mult = (double)cum[b] / std::max<double>(1.0, (double)s->h_state->occupied);
        const double est = ((double)s->size * 1.15 + 1024.0) * std::max(1.0, mult);
        target = std::max<uint64_t>(cum[b] * 2 + 1024, (uint64_t)std::min(est, 4.0e9));
    }
    TRY(pull_state(s));
    return FB2_OK;
}

static int absorb_log(fb2_sketcher *s, int par, uint32_t cnt) {
    if (s->size > 0 && cnt > 65536u && (uint64_t)cnt > 2 * s->size) {
        bool done = false;
        TRY(absorb_log_banded(s, par, cnt, &done));
        if (done) return FB2_OK;
        // the whole log was absorbed band by band: fall through to the usual post-absorb pruning
        cnt = 0;
    }
    uint32_t i = 0;
    bool exact = true;  // h_state->occupied is exact right after a pull
    while (i < cnt) {
        const uint32_t cap = s->tab[s->cur].cap;
        const uint32_t limit = cap / 4 * 3;
        const uint32_t occ = s->h_state->occupied;
        const uint32_t room = occ < limit ? limit - occ : 0;
        const uint32_t left = cnt - i;
        if (room < std::min(left, cap / 4)) {
            if (!exact) { TRY(pull_state(s)); exact = true; continue; }
            TRY(prune(s, std::min(left, cap / 4)));
            continue;
        }
        const uint32_t m = std::min(left, room);
        launch_absorb(log_view(s, par), i, i + m, s->tab[s->cur].view(), (SketchState *)s->d_state.p, s->st);
        s->stats.kernel_launches += 2;
        i += m;
        s->h_state->occupied = occ + m;  // upper bound until the next pull
        exact = false;
    }
    // keep the threshold moving: first finite threshold as soon as `size` keys exist, then at half load
    auto wants_prune = [&]() {
        const uint32_t occ = s->h_state->occupied;
        const bool infinite = s->h_state->threshold == ~0ULL;
        const bool can_lower = s->size > 0 && occ >= s->size;
        return (infinite && can_lower) || occ > s->tab[s->cur].cap / 2;
    };
    if (wants_prune()) {
        if (!exact) TRY(pull_state(s));
        if (wants_prune()) TRY(prune(s, 0));
    }
    return FB2_OK;
}

// ---- one chunk of raw bytes resident in HBM ---------------------------------------------------------
// The hash pieces of a chunk (device_types.cuh, PiecePlan): the parse kernel fills the table, the hash kernel walks it.
static PiecePlan piece_plan(const fb2_sketcher *s, int par) {
    PiecePlan pp;
    memset(&pp, 0, sizeof(pp));
    if (s->k > 32) return pp;                      // hash_big_kernel walks the regions itself
    pp.table = s->d_ptab[par].as<uint32_t>(); pp.count = s->d_pcount[par].as<uint32_t>();
    pp.stride = PIECE_STRIDE; pp.k = (uint32_t)s->k;
    pp.pmax = s->rec_pieces ? PIECE_SPAN / 32u - (uint32_t)s->k : 0u;   // 0: uniform pieces only
    pp.pmax_inv = pp.pmax ? ((1u << 22) + pp.pmax - 1u) / pp.pmax : 0u;
    return pp;
}
static ChunkGeom make_geom(uint32_t len) {
    ChunkGeom g;
    g.len = len;
    g.n_tiles = cdivu(len, TILE_BYTES);
    uint32_t stt = g.n_tiles / 4096u;  // ~4096 supertiles per full chunk: several waves of blocks, small tails
    const uint32_t forced = (uint32_t)env_size("FB2_ST_TILES", 0);   // tests: supertile size whatever the chunk size
    if (forced) stt = forced;
    if (stt < 1) stt = 1;
    if (stt > (uint32_t)parse_fused_max_tiles()) stt = (uint32_t)parse_fused_max_tiles();   // one supertile = one batch of the fused parse
    g.st_tiles = stt;
    g.n_st = cdivu(g.n_tiles, stt);
    g.st_bytes = stt * TILE_BYTES;
    g.region_stride = g.st_bytes + SYM_FRONT;
    g.hash_tiles = cdivu(g.st_bytes, HASH_TILE);
    return g;
}

// Synchronous hashing of a chunk's regions: launches sized so the candidate log cannot overflow
// unnoticed, the host reads the counters after each launch and absorbs the log (pruning as needed).
// Used while the threshold is still falling (early in a stream), for redo after a log overflow,
// and when kernel timing is on.
static int hash_range(fb2_sketcher *s, const ChunkGeom &g, uint64_t ord_base, int par) {
    SketchState *dst = (SketchState *)s->d_state.p;
    LaunchSlot *slot = dev_slot(s, par);
    const uint32_t total_blocks = g.n_st;      // launches cover whole symbol regions
    const uint32_t per_blk = g.st_bytes;       // positions a region can hold
    uint32_t b = 0;
    bool known = false;
    double fill = 1.0;  // symbols per launched position
    // Infinite threshold (start of a stream): every k-mer is a candidate.  If the whole chunk fits into
    // one launch of the log take it (small files: one launch, one banded absorb); otherwise start small,
    // so the threshold is finite before most of the chunk is hashed.
    // A chunk too large for that gets a PROVISIONAL finite threshold instead of the slow ramp: the
    // chunk holds N symbols, so T0 = 16 * size * 2^64 / N lets ~16 * size candidate occurrences through
    // (bounded, whatever the duplication) and the whole chunk is hashed in one launch.  T0 is a valid
    // admission threshold iff at least `size` DISTINCT keys turn out to lie at or below it, which is
    // checked after the absorb; if not (very repetitive input) the table is cleared and the chunk is
    // redone with the exact ramp.  Either way the result is that of an infinite initial threshold.
    bool provisional = false;
    unsigned long long prov_kmers = 0;
    if (s->h_state->threshold == ~0ULL) {
        const bool big = (uint64_t)total_blocks * per_blk > s->next_launch;
        const uint64_t N = s->h_carry->chunk_syms, want = 16ull * s->size;
        // (also when the chunk would fit one infinite-threshold launch: ~16 * size candidates instead of N)
        // only on an EMPTY table: the fallback below clears it, which must not lose keys of earlier chunks
        const bool empty = s->h_state->occupied == 0 && s->h_state->has_max_key == 0;
        if (empty && !s->skip_provisional && s->size > 0 && want + want / 8 <= s->log_cap / 2 && N > 4 * want &&
            !getenv("FB2_NO_PROVISIONAL")) {
            unsigned long long T0 = (~0ULL / N) * want;
            if (s->scaled && T0 < s->max_hash) T0 = s->max_hash;
            s->h_state->threshold = T0;
            TRY(push_state(s));
            TRY(refresh_live_hist(s));
            provisional = true;
            s->next_launch = 0x7FFFFFFFu / HASH_TILE * HASH_TILE;   // the whole chunk
        } else if (big) {
            s->next_launch = 32u * HASH_TILE;
        }
    }
    s->skip_provisional = false;
    while (b < total_blocks) {
        uint32_t nb = std::max<uint32_t>(s->next_launch / per_blk, 1u);
        nb = std::min(nb, total_blocks - b);
        CU(cudaMemsetAsync(slot, 0, sizeof(LaunchSlot), s->st));
        fb2_sketcher::EvPair evp;
        const bool timed = timing_begin(s, evp);
        // candidate-dense launches (infinite / provisional threshold, early ramp) reserve log slots in big batches
        launch_hash(s->k, s->d_sym[par].as<uint8_t>(), g, b, b + nb, s->d_rcount[par].as<uint32_t>(),
                    (const ParseCarry *)s->d_carry.p, ord_base, dst, slot, log_view(s, par), s->prm.hash_seed, 31u, piece_plan(s, par), s->st);
        if (timed) timing_end(s, evp, s->ev_hash_pending);
        s->stats.kernel_launches++; s->stats.hash_launches++;
        TRY(pull_state(s));
        if (!known) {
            known = true;
            s->stats.hash_symbols += s->h_carry->chunk_syms;
            fill = std::max(1e-6, (double)s->h_carry->chunk_syms / ((double)total_blocks * per_blk));
            if (s->h_carry->chunk_syms == 0) break;  // nothing to hash in this chunk
        }
        const uint32_t cnt = s->h_state->slot[par].log_count;
        if (cnt > s->log_cap) {  // log overflowed: nothing committed, redo this range in smaller launches
            s->next_launch = std::max<uint32_t>(per_blk, std::min((nb * per_blk) / 4, s->log_cap / HASH_TILE * HASH_TILE));
            s->steady = false;
            continue;
        }
        s->total_kmers += s->h_state->slot[par].launch_kmers;
        if (provisional) prov_kmers += s->h_state->slot[par].launch_kmers;
        TRY(absorb_log(s, par, cnt));
        b += nb;
        // next launch: aim the candidate count at a quarter of the log
        const double walked = std::max(1.0, (double)nb * per_blk * fill);
        const double frac = std::max((double)cnt, 1.0) / walked;           // candidates per symbol
        double next = (double)(s->log_cap / 4) / frac / fill;               // launched positions
        if (next > 2147483648.0) next = 2147483648.0;
        s->next_launch = std::max<uint32_t>(32u * HASH_TILE, (uint32_t)next);
    }
    if (provisional) {
        TRY(pull_state(s));
        const uint64_t have = (uint64_t)s->h_state->occupied + (s->h_state->has_max_key ? 1u : 0u);
        if (have < s->size) {
            // the provisional threshold was too low for this input: forget the chunk and redo it exactly
            s->total_kmers -= prov_kmers;
            launch_table_clear(s->tab[s->cur].view(), s->st);
            s->stats.kernel_launches++;
            s->h_state->threshold = ~0ULL; s->h_state->occupied = 0; s->h_state->has_max_key = 0;
            TRY(push_state(s));
            TRY(refresh_live_hist(s));
            s->skip_provisional = true;
            s->stats.provisional_redos++;
            return hash_range(s, g, ord_base, par);
        }
    }
    return FB2_OK;
}

// Verify an asynchronously hashed chunk one chunk later: commit its counters, or redo / prune in the
// rare cases the device-side guard declined to absorb its log.
static int settle(fb2_sketcher *s, int q) {
    if (!s->pend[q].valid) return FB2_OK;
    s->pend[q].valid = false;
    CU(wait_event(s, s->ev_chunk[q]));
    const SketchState snap = *s->h_snap[q];
    const LaunchSlot sl = snap.slot[q];
    const ChunkGeom g = s->pend[q].g;
    s->stats.hash_symbols += sl.chunk_syms;
    if (sl.decision == DECIDE_OVERFLOW) {
        // nothing of this chunk was absorbed: hash it again in bounded launches
        const double pos = (double)g.n_st * g.st_bytes;
        s->next_launch = std::max<uint32_t>(32u * HASH_TILE, (uint32_t)std::min(pos / 4.0, (double)(s->log_cap / 2)));
        s->steady = false;
        TRY(pull_state(s));
        return hash_range(s, g, s->pend[q].ord_base, q);
    }
    s->total_kmers += sl.launch_kmers;
    if (sl.decision == DECIDE_FULL) {
        TRY(pull_state(s));
        TRY(absorb_log(s, q, sl.log_count));
    } else {
        *s->h_state = snap;   // exact as of the end of this chunk's absorb
    }
    // adapt the launch size to the observed candidate rate
    const double pos = std::max(1.0, (double)g.n_st * g.st_bytes);
    const double per_pos = std::max((double)sl.log_count, 1.0) / pos;
    double next = (double)(s->log_cap / 4) / per_pos;
    if (next > 2147483648.0) next = 2147483648.0;
    s->next_launch = std::max<uint32_t>(32u * HASH_TILE, (uint32_t)next);
    if (s->h_state->occupied > s->tab[s->cur].cap / 2) {
        TRY(pull_state(s));   // drains the stream: later chunks may already have absorbed more
        if (s->h_state->occupied > s->tab[s->cur].cap / 2) TRY(prune(s, 0));
    }
    return FB2_OK;
}
static int settle_all(fb2_sketcher *s) {
    TRY(settle(s, s->par));       // older chunk first
    TRY(settle(s, s->par ^ 1));
    return FB2_OK;
}

static int run_chunk(fb2_sketcher *s, const uint8_t *d_raw, uint32_t len, int mode, int rawbuf /* -1: not ours */) {
    if (!len) return FB2_OK;
    const int par = s->par;
    TRY(ensure_logs(s));
    TRY(settle(s, par));           // its buffers are about to be reused (normally already settled)
    const ChunkGeom g = make_geom(len);
    s->last_geom = g;
    const void *stmap_was = s->d_stmap.p;
    TRY(s->d_stmap.ensure((size_t)g.n_st * 4)); TRY(s->d_ststate.ensure((size_t)g.n_st * 4));
    const bool status_fresh = s->d_stmap.p != stmap_was;
    TRY(s->d_rcount[par].ensure((size_t)g.n_st * 4));
    if (s->k <= 32) { TRY(s->d_ptab[par].ensure((size_t)g.n_st * PIECE_STRIDE * 4)); TRY(s->d_pcount[par].ensure((size_t)g.n_st * 4)); }
    const PiecePlan pp = piece_plan(s, par);
    if (mode == MODE_FASTQ) TRY(s->d_seam.ensure((size_t)g.n_st * sizeof(SeamNl)));
    TRY(s->d_sym[par].ensure((size_t)SYM_FRONT + (size_t)g.n_st * g.region_stride + 2 * HASH_TILE));
    ParseCarry *dc = (ParseCarry *)s->d_carry.p;
    SketchState *dst = (SketchState *)s->d_state.p;
    uint8_t *tail_in = s->d_tail.as<uint8_t>() + s->halo * s->tail_sel, *tail_out = s->d_tail.as<uint8_t>() + s->halo * (s->tail_sel ^ 1);
    fb2_sketcher::EvPair evparse;
    const bool parse_timed = timing_begin(s, evparse);
    if (s->parse_fused) {
        if (!s->d_ticket.p) {
            TRY(s->d_ticket.ensure(4));
            CU(cudaMemsetAsync(s->d_ticket.p, 0, 4, s->st));
            s->ticket_total = 0;
        }
        if (status_fresh || ++s->parse_epoch > 0xFFFFu) {   // new words, or the epoch wrapped: no stale word may match
            CU(cudaMemsetAsync(s->d_stmap.p, 0, s->d_stmap.cap, s->st));
            s->parse_epoch = 1;
        }
        launch_parse_fused(mode, d_raw, g, dc, s->d_ststate.as<uint32_t>(), s->d_sym[par].as<uint8_t>(), s->d_rcount[par].as<uint32_t>(),
                           tail_in, tail_out, s->d_seam.as<SeamNl>(), s->tail_sel, s->halo, s->d_stmap.as<uint32_t>(), s->parse_epoch,
                           s->d_ticket.as<uint32_t>(), s->ticket_total, pp, s->st);
        s->ticket_total += g.n_st;
    } else {
        launch_phase(mode, d_raw, g, dc, s->d_stmap.as<uint32_t>(), s->d_ststate.as<uint32_t>(), s->st);
        launch_pack(mode, d_raw, g, dc, s->d_ststate.as<uint32_t>(), s->d_sym[par].as<uint8_t>(), s->d_rcount[par].as<uint32_t>(),
                    tail_in, tail_out, s->d_seam.as<SeamNl>(), s->tail_sel, s->halo, s->st);
        launch_uniform_pieces(s->d_rcount[par].as<uint32_t>(), g.n_st, pp, dc, s->st);
    }
    if (parse_timed) timing_end(s, evparse, s->ev_parse_pending);
    if (rawbuf >= 0) { CU(cudaEventRecord(s->ev_rawfree[rawbuf], s->st)); s->rawfree_pending[rawbuf] = true; }
    s->tail_sel ^= 1;
    s->stats.kernel_launches += s->parse_fused ? 3 : (mode == MODE_LINES ? 3 : 4); s->stats.chunks++;
    const uint64_t ord_base = s->ordinal;
    s->ordinal += (uint64_t)g.n_st * g.st_bytes;
    s->par ^= 1;

    if (s->small_mode) { s->small_par = par; s->small_ord = ord_base; return FB2_OK; }   // fb2_sketcher_sketch_small drives the rest
    if (s->allcounts) {   // AllCountsSketcher::process (counts.rs:24-36): every valid window bumps its counter
        launch_count_kmers(s->d_sym[par].as<uint8_t>(), g, g.n_st, pp, (uint32_t)s->k, s->d_ac.as<uint32_t>(), s->st);
        s->stats.kernel_launches++;
        return FB2_OK;
    }
    const uint32_t total_blocks = g.n_st;      // regions
    const double positions = (double)g.n_st * g.st_bytes;
    if (positions <= (double)s->next_launch) s->steady = true;
    if (s->steady && positions <= (double)s->next_launch) {
        // asynchronous: hash the whole chunk, let the device decide about absorbing, snapshot the
        // state; the host looks at the outcome while the next chunk is already running
        // main stream: parse (above) and hash; absorb stream: log -> table, soft threshold, snapshot.
        // The absorb of this chunk then runs under the parse kernels of the next one (the persistent
        // hash kernel fills every SM, the latency-bound absorb kernels fit beside phase / pack).
        LaunchSlot *slot = dev_slot(s, par);
        cudaStream_t ab = getenv("FB2_NO_ABSORB_STREAM") ? s->st : s->st2;
        CU(cudaMemsetAsync(slot, 0, sizeof(LaunchSlot), s->st));
        launch_note_chunk_syms(slot, dc, s->st);
        fb2_sketcher::EvPair evp;
        const bool timed = timing_begin(s, evp);
        launch_hash(s->k, s->d_sym[par].as<uint8_t>(), g, 0, total_blocks, s->d_rcount[par].as<uint32_t>(), dc, ord_base, dst,
                    slot, log_view(s, par), s->prm.hash_seed, 3u, pp, s->st);
        if (timed) timing_end(s, evp, s->ev_hash_pending);
        if (ab != s->st) {
            CU(cudaEventRecord(s->ev_hash[par], s->st));
            CU(cudaStreamWaitEvent(ab, s->ev_hash[par], 0));
            s->st2_dirty = true;
        }
        launch_absorb_guarded(log_view(s, par), slot, s->tab[s->cur].view(), dst, dc, s->log_cap / 4, ab);
        if (s->size > 0 && !getenv("FB2_NO_SOFT_THRESHOLD")) {   // keep the admission threshold tight between rebuilds
            launch_soft_threshold(s->d_live_bins.as<uint32_t>(), s->live_shift, s->scaled ? 1 : 0, s->size, s->max_hash, dst, ab);
            s->stats.kernel_launches += 2;
        }
        CU(cudaMemcpyAsync(s->h_snap[par], dst, sizeof(SketchState), cudaMemcpyDeviceToHost, ab));
        CU(cudaEventRecord(s->ev_chunk[par], ab));
        s->stats.kernel_launches += 5; s->stats.hash_launches++;
        s->stats.d2h_bytes += sizeof(SketchState);
        s->pend[par].valid = true; s->pend[par].g = g; s->pend[par].ord_base = ord_base;
        TRY(settle(s, par ^ 1));   // the previous chunk, while this one runs
    } else {
        TRY(settle(s, par ^ 1));
        TRY(pull_state(s));
        TRY(hash_range(s, g, ord_base, par));
    }
    return FB2_OK;
}

// Host bytes -> chunks, H2D double-buffered on the copy stream against the compute stream.
static int feed_host_chunks(fb2_sketcher *s, const uint8_t *bytes, size_t len, int mode) {
    if (!len) return FB2_OK;
    const size_t chunk = s->chunk_bytes;
    const size_t nch = (len + chunk - 1) / chunk;
    const size_t bufsz = std::min(chunk, (len + 4095) / 4096 * 4096) + 64;
    auto issue = [&](size_t c) -> int {
        const int b = (int)(c & 1);
        const size_t off = c * chunk, n = std::min(chunk, len - off);
        TRY(s->d_raw[b].ensure(bufsz));
        if (s->rawfree_pending[b]) { CU(cudaStreamWaitEvent(s->copy_st, s->ev_rawfree[b], 0)); s->rawfree_pending[b] = false; }
        if (s->polite_copy) {
            // Another handle's (smaller, host-framed) copies share the link: hand it over after every piece instead of
            // queueing whole chunks ahead of them (hostlogic.cpp, sketch_stream_two_ended)
            const size_t piece = (size_t)s->polite_copy << 20;
            for (size_t q = 0; q < n; q += piece) {
                const size_t m = std::min(piece, n - q);
                if (s->link_flag && !s->link_owner)
                    while (s->link_flag->load(std::memory_order_acquire) > 0) std::this_thread::sleep_for(std::chrono::microseconds(20));
                CU(cudaMemcpyAsync(s->d_raw[b].as<uint8_t>() + q, bytes + off + q, m, cudaMemcpyHostToDevice, s->copy_st));
                CU(cudaEventRecord(s->ev_h2d[b], s->copy_st));
                if (q + m < n) CU(wait_event(s, s->ev_h2d[b]));
            }
        } else {
            CU(cudaMemcpyAsync(s->d_raw[b].p, bytes + off, n, cudaMemcpyHostToDevice, s->copy_st));
            CU(cudaEventRecord(s->ev_h2d[b], s->copy_st));
        }
        s->stats.h2d_bytes += n;
        return FB2_OK;
    };
    // make sure earlier work that used the raw buffers is ordered before we overwrite them
    for (int b = 0; b < 2; ++b) if (s->rawfree_pending[b]) { CU(cudaStreamWaitEvent(s->copy_st, s->ev_rawfree[b], 0)); s->rawfree_pending[b] = false; }
    TRY(issue(0));
    for (size_t c = 0; c < nch; ++c) {
        if (c + 1 < nch) TRY(issue(c + 1));
        const int b = (int)(c & 1);
        const size_t off = c * chunk, n = std::min(chunk, len - off);
        CU(cudaStreamWaitEvent(s->st, s->ev_h2d[b], 0));
        TRY(run_chunk(s, s->d_raw[b].as<uint8_t>(), (uint32_t)n, mode, b));
    }
    return FB2_OK;
}

static int flush_stage(fb2_sketcher *s) {
    if (!s->stage_fill) return FB2_OK;
    const size_t n = s->stage_fill;
    const int mode = s->stage_mode;
    s->stage_fill = 0;
    TRY(feed_host_chunks(s, s->h_stage, n, mode));
    // the staging buffer is reused right away: wait for its H2D copies
    CU(wait_stream(s, s->copy_st));
    return FB2_OK;
}
static int ensure_stage(fb2_sketcher *s) {
    if (s->h_stage) return FB2_OK;
    s->stage_cap = std::min<size_t>(s->chunk_bytes, 64u << 20);
    CU(cudaHostAlloc((void **)&s->h_stage, s->stage_cap, cudaHostAllocDefault));
    return FB2_OK;
}

static int flush_push(fb2_sketcher *s) {
    const uint32_t n = (uint32_t)s->push_extra.size();
    if (!n) return FB2_OK;
    TRY(settle_all(s));
    TRY(ensure_logs(s));
    const int par = s->par;
    TRY(s->d_push_bytes.ensure(s->push_bytes.size() + 16));
    TRY(s->d_push_offs.ensure((size_t)(n + 1) * 4));
    TRY(s->d_push_extra.ensure(n));
    CU(cudaMemcpyAsync(s->d_push_bytes.p, s->push_bytes.data(), s->push_bytes.size(), cudaMemcpyHostToDevice, s->st));
    CU(cudaMemcpyAsync(s->d_push_offs.p, s->push_offs.data(), (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, s->st));
    CU(cudaMemcpyAsync(s->d_push_extra.p, s->push_extra.data(), n, cudaMemcpyHostToDevice, s->st));
    s->stats.h2d_bytes += s->push_bytes.size() + (size_t)(n + 1) * 4 + n;
    SketchState *dst = (SketchState *)s->d_state.p;
    LaunchSlot *slot = dev_slot(s, par);
    CU(cudaMemsetAsync(slot, 0, sizeof(LaunchSlot), s->st));
    launch_push_hash(s->d_push_bytes.as<uint8_t>(), s->d_push_offs.as<uint32_t>(), s->d_push_extra.as<uint8_t>(), n,
                     s->arena_flushed, s->ordinal, dst, slot, log_view(s, par), s->prm.hash_seed, s->st);
    s->ordinal += n;
    s->stats.kernel_launches += 2;
    TRY(pull_state(s));
    s->total_kmers += s->h_state->slot[par].launch_kmers;
    s->arena_flushed += n;
    s->push_bytes.clear(); s->push_extra.clear(); s->push_offs.assign(1, 0u);
    TRY(absorb_log(s, par, s->h_state->slot[par].log_count));
    return FB2_OK;
}
static int flush_all(fb2_sketcher *s) {
    TRY(flush_stage(s));
    TRY(flush_push(s));
    TRY(settle_all(s));
    return FB2_OK;
}

// ---- SketchScheme::process ------------------------------------------------------------------------
extern "C" int fb2_sketcher_process(fb2_sketcher *s, const uint8_t *seq, size_t len) {
    if (!s || (!seq && len)) return fb2_fail(FB2_EINVAL, "null argument");
    ON_DEVICE(s->device);
    if (s->stream_open) return fb2_fail(FB2_EINVAL, "process() while a FASTX stream is open");
    s->fresh = false;
    TRY(flush_push(s));
    TRY(ensure_stage(s));
    if (s->stage_mode != MODE_LINES) { TRY(flush_stage(s)); s->stage_mode = MODE_LINES; }
    s->lines_bases += len;  // mash.rs:72: total_bases += seq.sequence().len()
    size_t done = 0;
    // the record goes into staging with interior '\n' rewritten to ' ' (both are dropped by
    // normalize) and one '\n' appended as the record separator; long records span flushes.
    while (done < len) {
        if (s->stage_fill == s->stage_cap) { TRY(flush_stage(s)); s->stage_mode = MODE_LINES; }
        const size_t n = std::min(len - done, s->stage_cap - s->stage_fill);
        uint8_t *dst = s->h_stage + s->stage_fill;
        for (size_t i = 0; i < n; ++i) { const uint8_t c = seq[done + i]; dst[i] = c == '\n' ? ' ' : c; }
        s->stage_fill += n; done += n;
    }
    if (s->stage_fill == s->stage_cap) { TRY(flush_stage(s)); s->stage_mode = MODE_LINES; }
    s->h_stage[s->stage_fill++] = '\n';
    return FB2_OK;
}

// ---- push -----------------------------------------------------------------------------------------
extern "C" int fb2_sketcher_push(fb2_sketcher *s, const uint8_t *kmer, size_t k, uint8_t extra_count) {
    if (s && s->allcounts) return fb2_fail(FB2_EUNSUPPORTED, "AllCountsSketcher has no push (counts.rs)");
    if (!s || (!kmer && k)) return fb2_fail(FB2_EINVAL, "null argument");
    if (k > 255) return fb2_fail(FB2_EINVAL, "k-mer longer than 255 bytes");
    s->fresh = false;
    ON_DEVICE(s->device);
    TRY(flush_stage(s));
    s->push_bytes.insert(s->push_bytes.end(), kmer, kmer + k);
    s->push_offs.push_back((uint32_t)s->push_bytes.size());
    s->push_extra.push_back(extra_count);
    s->arena.emplace_back((const char *)kmer, k);
    if (s->push_extra.size() >= (size_t)s->log_cap / 2) TRY(flush_push(s));
    return FB2_OK;
}

// ---- FASTX streams -------------------------------------------------------------------------------
// remember the last bytes of the stream (host copy) for the end-of-stream checks
static void note_tail(fb2_sketcher *s, const uint8_t *bytes, size_t len) {
    const size_t keep = 4096;
    if (len >= keep) { s->tail_host.assign(bytes + (len - keep), bytes + len); return; }
    s->tail_host.insert(s->tail_host.end(), bytes, bytes + len);
    if (s->tail_host.size() > keep) s->tail_host.erase(s->tail_host.begin(), s->tail_host.end() - keep);
}

static int begin_stream(fb2_sketcher *s, const uint8_t *first, size_t n) {
    const uint8_t c = first[0];
    if (c == '>') s->format = FB2_FORMAT_FASTA;
    else if (c == '@') s->format = FB2_FORMAT_FASTQ;
    else if ((c == 0x1f && n > 1 && first[1] == 0x8b) || (c == 'B' && n > 1 && first[1] == 'Z') ||
             (c == 0xFD && n > 1 && first[1] == '7'))
        return fb2_fail(FB2_EUNSUPPORTED, "compressed input is not supported; decompress first");
    else return fb2_fail(FB2_EFORMAT, "could not detect FASTA/FASTQ: first byte is neither '>' nor '@'");
    TRY(flush_all(s));
    TRY(pull_state(s));
    apply_finish_hint(s);
    ParseCarry *c_ = s->h_carry;
    c_->state = s->format == FB2_FORMAT_FASTA ? 1u : 0u;  // FASTA: pretend a header line precedes byte 0
    c_->prev1 = c_->prev2 = '\n';
    c_->raw_total = 0; c_->n_records = 0; c_->first_bad_pos = ~0ULL; c_->last_sig = 0; c_->error = 0;
    c_->len_bad_pos = ~0ULL;
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 3; ++b) c_->last_nl[a][b] = NL_NONE;
    TRY(push_carry(s));
    s->stream_open = true;
    s->range_mode = false;
    s->tail_host.clear();
    s->strip_on = false;
    if (s->format == FB2_FORMAT_FASTQ) {
        const char *e = getenv("FB2_HOST_STRIP");
        if ((e && *e == '1') || s->force_strip) {
            s->strip_on = true;
            s->strip_threads = (unsigned)std::min<size_t>(64, std::max<size_t>(1, env_size("FB2_STRIP_THREADS", std::min(16u, std::max(1u, std::thread::hardware_concurrency())))));
            s->strip_carry.clear(); s->strip_off = 0;
            s->strip_bad = s->strip_len_bad = s->strip_first_blank = ~0ULL; s->strip_last_nonblank = 0; s->strip_records = 0;
        }
    }
    return FB2_OK;
}

// ---- FASTQ with the host pre-strip (strip.cpp) ------------------------------------------------------------------
static void strip_note(fb2_sketcher *s, const StripOut &o) {
    s->strip_bad = std::min(s->strip_bad, o.bad_pos); s->strip_len_bad = std::min(s->strip_len_bad, o.len_bad_pos);
    s->strip_first_blank = std::min(s->strip_first_blank, o.first_blank);
    s->strip_last_nonblank = std::max(s->strip_last_nonblank, o.last_nonblank);
    s->strip_records += o.records;
    s->lines_bases += o.bases;          // mash.rs:72: total_bases += seq.sequence().len()
}
// append already stripped lines to the pinned staging of the MODE_LINES path (small pieces)
static int stage_lines(fb2_sketcher *s, const uint8_t *data, size_t n) {
    TRY(ensure_stage(s));
    if (s->stage_mode != MODE_LINES) { TRY(flush_stage(s)); s->stage_mode = MODE_LINES; }
    size_t done = 0;
    while (done < n) {
        if (s->stage_fill == s->stage_cap) { TRY(flush_stage(s)); s->stage_mode = MODE_LINES; }
        const size_t m = std::min(n - done, s->stage_cap - s->stage_fill);
        memcpy(s->h_stage + s->stage_fill, data + done, m);
        s->stage_fill += m; done += m;
    }
    return FB2_OK;
}
// strip [p, q) (whole records from a record start) on this thread and stage the lines
static int strip_small(fb2_sketcher *s, const uint8_t *p, const uint8_t *q, uint64_t off, const uint8_t **stopped) {
    StripOut o;
    o.origin = p;
    o.spill.resize((size_t)(q - p) / 2 + 16);
    o.spilled = true;
    *stopped = strip_records(p, q, nullptr, off, o);
    strip_note(s, o);
    return stage_lines(s, o.spill.data(), o.out_len);
}
static int feed_fastq_stripped(fb2_sketcher *s, const uint8_t *bytes, size_t len, int final) {
    const uint8_t *p = bytes, *end = bytes + len;
    // 1. the record left incomplete by the previous piece
    if (!s->strip_carry.empty() && len) {
        const uint8_t *q = complete_record(s->strip_carry, p, end);
        if (!q) { s->strip_carry.insert(s->strip_carry.end(), p, end); p = end; }
        else {
            std::vector<uint8_t> joined(s->strip_carry);
            joined.insert(joined.end(), p, q);
            const uint8_t *st = nullptr;
            TRY(strip_small(s, joined.data(), joined.data() + joined.size(), s->strip_off, &st));
            s->strip_off += joined.size();
            s->strip_carry.clear();
            p = q;
        }
    }
    // 2. whole records of this piece
    if (p < end && s->strip_carry.empty()) {
        if ((size_t)(end - p) < (1u << 20)) {
            const uint8_t *st = nullptr;
            TRY(strip_small(s, p, end, s->strip_off, &st));
            s->strip_off += (uint64_t)(st - p);
            p = st;
        } else {
            TRY(flush_stage(s));                         // staged (older) lines go first: stream order = position order
            const unsigned T = s->strip_threads;
            size_t block = 2 * s->chunk_bytes;           // stripped, a block is under half its size: one device chunk
            const size_t each = (block / T) / 2 + (1u << 20);
            if (!s->strip_pin[0] || s->strip_pin_each < each) {
                for (int i = 0; i < 2; ++i) {
                    if (s->strip_pin[i]) cudaFreeHost(s->strip_pin[i]);
                    s->strip_pin[i] = nullptr;
                    CU(cudaHostAlloc((void **)&s->strip_pin[i], each * T, cudaHostAllocDefault));
                    if (!s->ev_strip_free[i]) CU(cudaEventCreateWithFlags(&s->ev_strip_free[i], cudaEventDisableTiming));
                    s->strip_free_pending[i] = false;
                }
                s->strip_pin_each = each;
            }
            std::vector<StripOut> outs(T);
            while (p < end) {
                const uint8_t *blk_end = (size_t)(end - p) > block ? p + block : end;
                const int set = s->strip_set;
                if (s->strip_free_pending[set]) { CU(wait_event(s, s->ev_strip_free[set])); s->strip_free_pending[set] = false; }
                for (unsigned t = 0; t < T; ++t) {
                    outs[t] = StripOut();
                    outs[t].out = s->strip_pin[set] + (size_t)t * s->strip_pin_each;
                    outs[t].out_cap = s->strip_pin_each;
                }
                const uint8_t *e = strip_parallel(p, blk_end, s->strip_off, T, outs);
                size_t total = 0;
                for (auto &o : outs) { strip_note(s, o); total += o.out_len; }
                if (total) {
                    if (total > s->chunk_bytes) return fb2_fail(FB2_EINVAL, "internal: stripped block exceeds a chunk");
                    const int b = set;                   // raw buffer b pairs with staging set b
                    TRY(s->d_raw[b].ensure(std::min(s->chunk_bytes, (total + 4095) / 4096 * 4096) + 64));
                    if (s->rawfree_pending[b]) { CU(cudaStreamWaitEvent(s->copy_st, s->ev_rawfree[b], 0)); s->rawfree_pending[b] = false; }
                    size_t at = 0;
                    for (auto &o : outs) {
                        if (!o.out_len) continue;
                        CU(cudaMemcpyAsync(s->d_raw[b].as<uint8_t>() + at, o.data(), o.out_len, cudaMemcpyHostToDevice, s->copy_st));
                        at += o.out_len;
                        if (o.spilled) CU(wait_stream(s, s->copy_st));   // pageable source: gone after this iteration
                    }
                    CU(cudaEventRecord(s->ev_h2d[b], s->copy_st));
                    if (s->link_flag && s->link_owner) {   // lowered by the driver when these copies are through
                        s->link_flag->fetch_add(1, std::memory_order_acq_rel);
                        CU(cudaLaunchHostFunc(s->copy_st, [](void *f) { ((std::atomic<int> *)f)->fetch_sub(1, std::memory_order_acq_rel); }, s->link_flag));
                    }
                    CU(cudaEventRecord(s->ev_strip_free[set], s->copy_st));
                    s->strip_free_pending[set] = true;
                    s->stats.h2d_bytes += total;
                    CU(cudaStreamWaitEvent(s->st, s->ev_h2d[b], 0));
                    TRY(run_chunk(s, s->d_raw[b].as<uint8_t>(), (uint32_t)total, MODE_LINES, b));
                    s->strip_set ^= 1;
                }
                s->strip_off += (uint64_t)(e - p);
                if (e == p) {                            // not one complete record in the block
                    if (blk_end == end) break;
                    block *= 2;                          // a very long record: look further
                    continue;
                }
                p = e;
            }
        }
        if (p < end) s->strip_carry.assign(p, end);      // the incomplete record waits for the next piece
    }
    if (!final) return FB2_OK;
    // 3. end of the stream: the reader's end-of-input rules on what is left
    int bad_tail = 0;
    if (!s->strip_carry.empty()) {
        StripOut o;
        o.origin = s->strip_carry.data();
        o.spill.resize(s->strip_carry.size() / 2 + 16);
        o.spilled = true;
        bad_tail = strip_final(s->strip_carry.data(), s->strip_carry.data() + s->strip_carry.size(), s->strip_off, o);
        strip_note(s, o);
        TRY(stage_lines(s, o.spill.data(), o.out_len));
        s->strip_carry.clear();
    }
    TRY(flush_stage(s));
    TRY(settle_all(s));
    TRY(pull_state(s));
    s->h_carry->state = 0; s->h_carry->prev1 = s->h_carry->prev2 = '\n';
    TRY(push_carry(s));
    launch_fill_bytes(s->d_tail.as<uint8_t>(), 2 * HALO_BIG, SYM_BREAK, s->st);
    s->stats.kernel_launches++;
    s->stream_open = false;
    s->strip_on = false;
    if (s->strip_bad != ~0ULL)
        return fb2_fail(FB2_ERECORD, "invalid FASTQ record: line at byte " + std::to_string(s->strip_bad) + " does not start with the expected '@' / '+'");
    if (s->strip_len_bad != ~0ULL)
        return fb2_fail(FB2_ERECORD, "invalid FASTQ record: sequence and quality lengths differ (record whose header ends at byte " +
                                         std::to_string(s->strip_len_bad) + ")");
    if (s->strip_first_blank != ~0ULL && s->strip_first_blank < s->strip_last_nonblank)
        return fb2_fail(FB2_ERECORD, "invalid FASTQ record: blank lines at byte " + std::to_string(s->strip_first_blank) + " inside the stream");
    if (bad_tail) return fb2_fail(FB2_ERECORD, "truncated or invalid FASTQ record at end of input");
    if (s->strip_records == 0) return fb2_fail(FB2_EEMPTY, "no records in FASTQ stream");
    return FB2_OK;
}

// What the reader makes of the stream now that its last byte is in: FASTA -- the last record's length; FASTQ -- the
// end-of-input rules (truncated record, record errors seen by the kernels, the last record's length check).  Works on
// the host mirror of the carry (pulled by the caller) and the host's copy of the stream's last bytes.
static int stream_verdict(fb2_sketcher *s) {
    ParseCarry *c = s->h_carry;
    int rc = FB2_OK;
    if (s->format == FB2_FORMAT_FASTA) {
        // last record: raw sequence loses its final '\n' (+ CR before it) or a dangling CR
        if (c->state == 0u) {
            if (c->prev1 == '\n') { c->total_bases -= 1; if (c->prev2 == '\r') c->total_bases -= 1; }
            else if (c->prev1 == '\r') c->total_bases -= 1;
        }
    } else if (s->format == FB2_FORMAT_FASTQ) {
        // Trailing blank lines are tolerated; the last byte that is neither CR nor LF must lie in a
        // quality line (phase 3), and every line before it must have passed the '@' / '+' checks.
        // c->state is the phase of the line the next byte would belong to.
        const std::vector<uint8_t> &t = s->tail_host;
        size_t n_trail = 0, nl_trail = 0;
        while (n_trail < t.size() && (t[t.size() - 1 - n_trail] == '\n' || t[t.size() - 1 - n_trail] == '\r')) {
            nl_trail += t[t.size() - 1 - n_trail] == '\n';
            ++n_trail;
        }
        const uint64_t L = c->raw_total;
        if (n_trail == L) rc = fb2_fail(FB2_EEMPTY, "no records in FASTQ stream");
        else if (n_trail == t.size()) rc = fb2_fail(FB2_ERECORD, "too many trailing line terminators after the last FASTQ record");
        else {
            const uint32_t phase_last = (c->state + 4u - (uint32_t)(nl_trail % 4)) & 3u;
            const uint64_t last_sig_pos = L - n_trail - 1;
            // A last record with an EMPTY quality line: the last significant byte then lies in its '+' line.  The
            // reader accepts it when the '+' line is terminated and the sequence is as long as what follows (nothing,
            // or carriage returns: one is trimmed).  Decided here from the host's copy of the stream's last bytes.
            bool empty_qual_ok = false, empty_qual_lens_differ = false;
            if (phase_last == 2u && nl_trail >= 1) {
                const size_t sig = t.size() - 1 - n_trail;                 // index of the last significant byte in t
                size_t l3 = sig + 1;
                while (t[l3] != '\n') ++l3;                                 // newline of the '+' line (exists: nl_trail >= 1)
                size_t crs = 0;
                for (size_t j = l3 + 1; j < t.size() && t[j] != '\n'; ++j) ++crs;
                const uint64_t qual_len = crs ? crs - 1 : 0;
                long e1 = (long)sig;
                while (e1 >= 0 && t[(size_t)e1] != '\n') --e1;              // newline that ends the sequence line
                if (e1 >= 0) {
                    long e0 = e1 - 1;
                    while (e0 >= 0 && t[(size_t)e0] != '\n') --e0;          // newline that ends the header line
                    if (e0 >= 0) {
                        const uint64_t seq_len = (uint64_t)(e1 - e0 - 1) - ((e1 - 1 > e0 && t[(size_t)e1 - 1] == '\r') ? 1 : 0);
                        empty_qual_ok = seq_len == qual_len;
                        empty_qual_lens_differ = !empty_qual_ok;
                    } else if ((uint64_t)e1 > qual_len + 1) empty_qual_lens_differ = true;   // sequence longer than the host's tail copy
                }
            }
            if (empty_qual_lens_differ)
                rc = fb2_fail(FB2_ERECORD, "invalid FASTQ record: sequence and quality lengths differ (last record)");
            else if (phase_last != 3u && !empty_qual_ok)
                rc = fb2_fail(FB2_ERECORD, "truncated FASTQ record at end of input");
            else if (c->first_bad_pos != ~0ULL && c->first_bad_pos <= last_sig_pos)
                rc = fb2_fail(FB2_ERECORD, "invalid FASTQ record: line at byte " + std::to_string(c->first_bad_pos) +
                                               " does not start with the expected '@' / '+'");
            else {
                // needletail: sequence and quality line of a record have the same length (lib.rs:63 panics otherwise).
                // The kernels checked every quality line that ends in a newline; a last one that does not is checked
                // here from the stream's last three newlines (the header, sequence and '+' line ends of that record).
                uint64_t bad = c->len_bad_pos;
                if (nl_trail == 0 && phase_last == 3u) {
                    const uint64_t *e = c->last_nl[s->tail_sel];   // as left by the last chunk
                    const unsigned long long PM = ~(1ULL << 63);
                    if (e[0] != NL_NONE && e[1] != NL_NONE && e[2] != NL_NONE) {
                        const uint64_t ls = (e[1] & PM) - (e[0] & PM) - 1 - (e[1] >> 63);
                        const uint64_t lq = L - (e[2] & PM) - 1 - (t.back() == '\r' ? 1 : 0);
                        if (ls != lq) bad = std::min<uint64_t>(bad, e[0] & PM);
                    }
                }
                if (bad != ~0ULL && bad <= last_sig_pos)
                    rc = fb2_fail(FB2_ERECORD, "invalid FASTQ record: sequence and quality lengths differ (record whose header ends at byte " +
                                                   std::to_string(bad) + ")");
            }
        }
    }
    return rc;
}
static int end_stream(fb2_sketcher *s) {
    TRY(flush_stage(s));
    TRY(settle_all(s));
    TRY(pull_state(s));
    ParseCarry *c = s->h_carry;
    const int rc = stream_verdict(s);
    // a later stream or record must not join this one: break the carried symbols
    c->state = 0; c->prev1 = c->prev2 = '\n';
    TRY(push_carry(s));
    launch_fill_bytes(s->d_tail.as<uint8_t>(), 2 * HALO_BIG, SYM_BREAK, s->st);
    s->stats.kernel_launches++;
    s->stream_open = false;
    return rc;
}

extern "C" int fb2_sketcher_feed_fastx(fb2_sketcher *s, const uint8_t *bytes, size_t len, int final) {
    if (!s || (!bytes && len)) return fb2_fail(FB2_EINVAL, "null argument");
    s->fresh = false;
    ON_DEVICE(s->device);
    if (!s->stream_open) {
        // format sniffing looks at two bytes (compressed-input magic): a shorter first piece waits for the next one
        const size_t have = s->presniff.size();
        if (have + len < 2 && !final) {
            if (len) s->presniff.insert(s->presniff.end(), bytes, bytes + len);
            return FB2_OK;
        }
        if (have) {
            std::vector<uint8_t> joined(s->presniff);
            s->presniff.clear();
            joined.insert(joined.end(), bytes, bytes + len);
            return fb2_sketcher_feed_fastx(s, joined.data(), joined.size(), final);
        }
    }
    if (len) {
        if (!s->stream_open) TRY(begin_stream(s, bytes, len));
        const int mode = s->format == FB2_FORMAT_FASTA ? MODE_FASTA : MODE_FASTQ;
        if (s->strip_on) {
            TRY(feed_fastq_stripped(s, bytes, len, final));
            if (len >= (1u << 20) && !s->range_mode) CU(wait_stream(s, s->copy_st));   // (staging sets are ours; the caller's bytes were only read by the CPU)
            return FB2_OK;
        }
        note_tail(s, bytes, len);
        if (len < (1u << 20)) {  // small piece: gather in pinned staging
            TRY(ensure_stage(s));
            if (s->stage_mode != mode) { TRY(flush_stage(s)); s->stage_mode = mode; }
            size_t done = 0;
            while (done < len) {
                if (s->stage_fill == s->stage_cap) { TRY(flush_stage(s)); s->stage_mode = mode; }
                const size_t n = std::min(len - done, s->stage_cap - s->stage_fill);
                memcpy(s->h_stage + s->stage_fill, bytes + done, n);
                s->stage_fill += n; done += n;
            }
        } else {
            TRY(flush_stage(s));
            TRY(feed_host_chunks(s, bytes, len, mode));
            CU(wait_stream(s, s->copy_st));  // caller may reuse `bytes` after we return
        }
    }
    if (final) {
        if (!s->stream_open) return fb2_fail(FB2_EEMPTY, "empty input: no records");
        if (s->strip_on) return feed_fastq_stripped(s, nullptr, 0, 1);
        TRY(end_stream(s));
    }
    return FB2_OK;
}

extern "C" int fb2_sketcher_feed_device(fb2_sketcher *s, const uint8_t *dev, size_t len, int final) {
    if (!s || (!dev && len)) return fb2_fail(FB2_EINVAL, "null argument");
    ON_DEVICE(s->device);
    if (len) {
    s->fresh = false;
        if (!s->stream_open) {
            uint8_t first[2] = {0, 0};
            CU(cudaMemcpy(first, dev, std::min<size_t>(2, len), cudaMemcpyDeviceToHost));
            TRY(begin_stream(s, first, std::min<size_t>(2, len)));
        }
        TRY(flush_stage(s));
        const int mode = s->format == FB2_FORMAT_FASTA ? MODE_FASTA : MODE_FASTQ;
        {
            const size_t tl = std::min<size_t>(len, 4096);
            std::vector<uint8_t> tmp(tl);
            CU(cudaMemcpy(tmp.data(), dev + (len - tl), tl, cudaMemcpyDeviceToHost));
            note_tail(s, tmp.data(), tl);
        }
        const size_t chunk = s->chunk_bytes;
        const bool aligned = ((uintptr_t)dev & 15u) == 0;
        for (size_t off = 0; off < len; off += chunk) {
            const size_t n = std::min(chunk, len - off);
            if (aligned) {
                TRY(run_chunk(s, dev + off, (uint32_t)n, mode, -1));
            } else {
                TRY(s->d_raw[0].ensure(n + 64));
                CU(cudaMemcpyAsync(s->d_raw[0].p, dev + off, n, cudaMemcpyDeviceToDevice, s->st));
                TRY(run_chunk(s, s->d_raw[0].as<uint8_t>(), (uint32_t)n, mode, -1));
                CU(wait_stream(s, s->st));   // d_raw[0] is reused by the next piece
            }
        }
        if (aligned) {  // the caller may release `dev` after we return: the parse kernels must be done
            CU(cudaEventRecord(s->ev_rawfree[0], s->st));
            CU(wait_event(s, s->ev_rawfree[0]));
        }
    }
    if (final) {
        if (!s->stream_open) return fb2_fail(FB2_EEMPTY, "empty input: no records");
        TRY(end_stream(s));
    }
    return FB2_OK;
}

// ---- one FASTX stream split into byte ranges over several sketchers (hostlogic.cpp: fb2_sketch_stream_multi) ------
// Internal API.  A range starts at a line start chosen by the host, which supplies what the parser would have
// carried to that point: the line state, the two previous raw bytes, the `halo` symbols that precede the range in
// stream order (k-mers span the cut), the stream offset and a position-id base (later ranges get larger ids so the
// first occurrence of a hash in stream order survives the merge).
static void apply_finish_hint(fb2_sketcher *s) {
    if (s->hint_set && !s->scaled && s->hint_final < s->size && s->h_state->occupied == 0 && s->h_state->has_max_key == 0 &&
        s->total_kmers == 0 && (s->hint_filter == 0 || (s->hint_filter < 0 && s->format == FB2_FORMAT_FASTA))) {
        // the filter resolves to off (lib.rs:71-76) and the result is truncated to final_size: the bottom final_size
        // of the bottom kmers_to_sketch is the bottom final_size -- sketch with the small heap
        s->size = s->hint_final;
        size_logs(s);
    }
}
int fb2_sketcher_begin_range(fb2_sketcher *s, int format, uint32_t state, uint32_t prev1, uint32_t prev2,
                             const uint8_t *tail_syms /* halo symbols, may be null = all breaks */, uint64_t raw_base,
                             uint64_t ord_base, int strip /* FASTQ range at a record start: frame the records on the host */) {
    if (!s || s->stream_open) return fb2_fail(FB2_EINVAL, "begin_range: bad handle state");
    s->fresh = false;
    ON_DEVICE(s->device);
    s->format = format;
    TRY(flush_all(s));
    TRY(pull_state(s));
    apply_finish_hint(s);
    ParseCarry *c = s->h_carry;
    c->state = state; c->prev1 = prev1; c->prev2 = prev2;
    c->raw_total = raw_base; c->n_records = 0; c->first_bad_pos = ~0ULL; c->last_sig = 0; c->error = 0;
    c->len_bad_pos = ~0ULL;
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 3; ++b) c->last_nl[a][b] = NL_NONE;
    TRY(push_carry(s));
    if (tail_syms) {
        CU(cudaMemcpyAsync(s->d_tail.as<uint8_t>() + s->halo * s->tail_sel, tail_syms, s->halo, cudaMemcpyHostToDevice, s->st));
        CU(wait_stream(s, s->st));
    }
    s->ordinal = ord_base;
    s->stream_open = true;
    s->range_mode = true;
    s->tail_host.clear();
    s->strip_on = false;
    if (strip && format == FB2_FORMAT_FASTQ) {
        s->strip_on = true;
        s->strip_threads = (unsigned)std::min<size_t>(64, std::max<size_t>(1, env_size("FB2_STRIP_THREADS", std::min(16u, std::max(1u, std::thread::hardware_concurrency())))));
        s->strip_carry.clear(); s->strip_off = raw_base;
        s->strip_bad = s->strip_len_bad = s->strip_first_blank = ~0ULL; s->strip_last_nonblank = 0; s->strip_records = 0;
    }
    return FB2_OK;
}
// End of a range that is NOT the end of the stream: everything queued is done, no end-of-stream rules applied.
// Reports the parser state the next range must have started from, and the first record errors seen (~0: none).
int fb2_sketcher_end_range(fb2_sketcher *s, uint32_t *end_state, uint32_t *last_byte, uint64_t *first_bad_pos, uint64_t *len_bad_pos) {
    if (!s) return fb2_fail(FB2_EINVAL, "null handle");
    ON_DEVICE(s->device);
    TRY(flush_stage(s));
    TRY(settle_all(s));
    TRY(pull_state(s));
    if (s->strip_on) {
        // host-framed range: it ends at a record start iff no bytes of an incomplete record are left over; blank
        // "records" are reported as a framing error (only the end of the whole stream may hold them)
        const bool clean = s->strip_carry.empty() && s->strip_first_blank == ~0ULL;
        if (end_state) *end_state = clean ? 0u : 1u;
        if (last_byte) *last_byte = '\n';
        if (first_bad_pos) *first_bad_pos = s->strip_bad;
        if (len_bad_pos) *len_bad_pos = s->strip_len_bad;
        s->strip_carry.clear();
        s->strip_on = false;
        s->h_carry->state = 0; s->h_carry->prev1 = s->h_carry->prev2 = '\n';
        TRY(push_carry(s));
        launch_fill_bytes(s->d_tail.as<uint8_t>(), 2 * HALO_BIG, SYM_BREAK, s->st);
        s->stats.kernel_launches++;
        s->stream_open = false;
        return FB2_OK;
    }
    if (end_state) *end_state = s->h_carry->state;
    if (last_byte) *last_byte = s->h_carry->prev1;
    if (first_bad_pos) *first_bad_pos = s->h_carry->first_bad_pos;
    if (len_bad_pos) *len_bad_pos = s->h_carry->len_bad_pos;
    s->stream_open = false;
    return FB2_OK;
}
void fb2_sketcher_set_polite_copy(fb2_sketcher *s, unsigned piece_mb) { if (s) s->polite_copy = piece_mb; }
void fb2_sketcher_set_polite_sync(fb2_sketcher *s, int on) { if (s) s->polite = on != 0; }
void fb2_sketcher_set_force_strip(fb2_sketcher *s, int on) { if (s) s->force_strip = on != 0; }
void fb2_sketcher_set_link_flag(fb2_sketcher *s, std::atomic<int> *flag, int owner) { if (s) { s->link_flag = flag; s->link_owner = owner != 0; } }
uint32_t fb2_sketcher_halo(const fb2_sketcher *s) { return s->halo; }
int fb2_sketcher_device(const fb2_sketcher *s) { return s->device; }

// Exact union of `src`'s sketch state into `dst` (both idle, same parameters): the live entries of src's table are
// merged into dst's table by kernels running on dst's GPU that read src's table through peer access (NVLink); when
// the devices cannot address each other the table is first copied with cudaMemcpyPeer.  Totals add up.
int fb2_sketcher_merge_from(fb2_sketcher *dst, fb2_sketcher *src) {
    if (!dst || !src || dst == src) return fb2_fail(FB2_EINVAL, "merge: bad handles");
    if (dst->k != src->k || dst->scaled != src->scaled || dst->size != src->size || dst->prm.hash_seed != src->prm.hash_seed ||
        (dst->scaled && dst->max_hash != src->max_hash))
        return fb2_fail(FB2_EINVAL, "merge: sketchers differ in parameters");
    DeviceScope dev_scope_(dst->device);      // restores the caller's device on every return path
    CU(cudaSetDevice(src->device));
    TRY(flush_all(src));
    TRY(pull_state(src));
    const SketchState ss = *src->h_state;
    const TableView sv = src->tab[src->cur].view();
    CU(cudaSetDevice(dst->device));
    TRY(flush_all(dst));
    TRY(pull_state(dst));
    // room for every live source entry (upper bound: its occupancy)
    {
        const uint64_t need = (uint64_t)ss.occupied + 1;
        const uint32_t cap = dst->tab[dst->cur].cap;
        if ((uint64_t)dst->h_state->occupied + need > (uint64_t)cap / 4 * 3) TRY(prune(dst, (uint32_t)std::min<uint64_t>(need, 1u << 30)));
    }
    if (ss.threshold < dst->h_state->threshold) {      // both are valid admission thresholds of the whole stream: keep the lower
        dst->h_state->threshold = ss.threshold;
        TRY(push_state(dst));
    }
    TableView from = sv;
    DevBuf cp[5];
    if (src->device != dst->device) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, dst->device, src->device);
        bool direct = can != 0 && !getenv("FB2_NO_PEER_ACCESS");
        if (direct) {
            const cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) direct = false;
            cudaGetLastError();
        }
        if (!direct) {   // staged: peer copies of the five table arrays into scratch on dst's device
            const size_t n = (size_t)sv.cap + 1;
            const void *srcp[5] = {sv.key, sv.cnt, sv.ext, sv.posx, sv.kmer};
            for (int a = 0; a < 5; ++a) {
                const size_t bytes = n * 8 * (a == 4 ? sv.kw : 1);
                TRY(cp[a].ensure(bytes));
                CU(cudaMemcpyPeerAsync(cp[a].p, dst->device, srcp[a], src->device, bytes, dst->st));
            }
            from.key = cp[0].as<unsigned long long>(); from.cnt = cp[1].as<unsigned long long>(); from.ext = cp[2].as<unsigned long long>();
            from.posx = cp[3].as<unsigned long long>(); from.kmer = cp[4].as<unsigned long long>();
        }
    }
    launch_merge_tables(from, ss.threshold, ss.has_max_key, dst->tab[dst->cur].view(), (SketchState *)dst->d_state.p, dst->st);
    dst->stats.kernel_launches += 2;
    TRY(pull_state(dst));
    for (auto &b : cp) b.release();
    dst->total_kmers += src->total_kmers;
    dst->lines_bases += src->h_carry->total_bases + src->lines_bases;   // host-side sum: seq_length = carry + lines_bases
    return FB2_OK;
}

// Internal (hostlogic.cpp): the caller promises to finish this handle's streams with
// fb2_sketcher_sketch(final_size, filter) only; see begin_stream.  Survives reset.
void fb2_sketcher_hint_finish(fb2_sketcher *s, uint64_t final_size, int filter_on) {
    s->hint_set = true; s->hint_final = final_size; s->hint_filter = filter_on;
}
size_t fb2_sketcher_chunk_bytes(const fb2_sketcher *s) { return s->chunk_bytes; }
// Device memory this handle holds right now (its pooled retention cost, hostlogic.cpp).
size_t fb2_sketcher_device_bytes(const fb2_sketcher *s) {
    size_t n = 0;
    const DevBuf *bufs[] = {&s->d_raw[0], &s->d_raw[1], &s->d_sym[0], &s->d_sym[1], &s->d_stmap, &s->d_ststate, &s->d_rcount[0],
                            &s->d_rcount[1], &s->d_ptab[0], &s->d_ptab[1], &s->d_pcount[0], &s->d_pcount[1], &s->d_tail, &s->d_carry, &s->d_state, &s->d_seam, &s->log_hash[0], &s->log_hash[1],
                            &s->log_kmer[0], &s->log_kmer[1], &s->log_posx[0], &s->log_posx[1], &s->sort_keys, &s->sort_slots,
                            &s->sort_tkeys, &s->sort_tslots, &s->sort_hist, &s->d_bins, &s->d_live_bins, &s->out_hash, &s->out_cnt,
                            &s->out_ext, &s->out_kmer, &s->out_posx, &s->sel_hash, &s->sel_cnt, &s->sel_ext, &s->sel_kmer, &s->sel_posx,
                            &s->sel_bytes, &s->sel_idx, &s->d_push_bytes, &s->d_push_offs, &s->d_push_extra};
    for (const DevBuf *b : bufs) n += b->cap;
    for (int t = 0; t < 2; ++t) n += s->tab[t].key.cap + s->tab[t].cnt.cap + s->tab[t].ext.cap + s->tab[t].posx.cap + s->tab[t].kmer.cap;
    return n;
}

extern "C" int fb2_sketcher_format(fb2_sketcher *s, int32_t *format) {
    if (!s || !format) return fb2_fail(FB2_EINVAL, "null argument");
    *format = s->format;
    return FB2_OK;
}

// ---- totals / result -------------------------------------------------------------------------------
static int ensure_hres(fb2_sketcher *s, size_t bytes);
// AllCountsSketcher::to_vec (counts.rs:38-63) on the device: which indices make an entry, where it goes (two-level
// scan, index order kept), then hash = index, count, extra_count and the k-mer's bytes.  out == nullptr: only the sum
// of the counters (total_bases_and_kmers).
static int allcounts_result(fb2_sketcher *s, fb2_result *out, uint64_t *sum_out) {
    const uint64_t n = (uint64_t)1 << (2 * s->k);
    const uint32_t nb = allcounts_blocks(n);
    TRY(s->d_ac_off.ensure((size_t)nb * 4)); TRY(s->d_ac_meta.ensure(16));
    CU(cudaMemsetAsync(s->d_ac_meta.p, 0, 16, s->st));
    launch_allcounts_plan(s->d_ac.as<uint32_t>(), n, (uint32_t)s->k, s->d_ac_off.as<uint32_t>(), s->d_ac_meta.as<unsigned long long>(), s->st);
    s->stats.kernel_launches += 2;
    TRY(ensure_hres(s, 16));
    CU(cudaMemcpyAsync(s->h_res, s->d_ac_meta.p, 16, cudaMemcpyDeviceToHost, s->st));
    CU(wait_stream(s, s->st));
    const uint64_t m = ((uint64_t *)s->h_res)[0], sum = ((uint64_t *)s->h_res)[1];
    if (sum_out) *sum_out = sum;
    if (!out) return FB2_OK;
    if (m > 0xFFFFFFFFull) return fb2_fail(FB2_ENOMEM, "AllCounts: too many entries to return");
    const size_t k = (size_t)s->k;
    out->n = m;
    out->kmer_stride = (uint32_t)k;
    out->hashes = (uint64_t *)malloc(std::max<size_t>(1, (size_t)m * 8));
    out->counts = (uint32_t *)malloc(std::max<size_t>(1, (size_t)m * 4));
    out->extras = (uint32_t *)malloc(std::max<size_t>(1, (size_t)m * 4));
    out->kmers = (uint8_t *)malloc(std::max<size_t>(1, (size_t)m * k));
    if (!out->hashes || !out->counts || !out->extras || !out->kmers) { fb2_result_free(out); return fb2_fail(FB2_ENOMEM, "malloc"); }
    if (m) {
        TRY(s->out_hash.ensure((size_t)m * 8)); TRY(s->out_cnt.ensure((size_t)m * 4)); TRY(s->out_ext.ensure((size_t)m * 4));
        TRY(s->sel_bytes.ensure((size_t)m * k));
        launch_allcounts_emit(s->d_ac.as<uint32_t>(), n, (uint32_t)s->k, s->d_ac_off.as<uint32_t>(), s->out_hash.as<unsigned long long>(),
                              s->out_cnt.as<uint32_t>(), s->out_ext.as<uint32_t>(), s->sel_bytes.as<uint8_t>(), s->st);
        s->stats.kernel_launches++;
        CU(cudaMemcpyAsync(out->hashes, s->out_hash.p, (size_t)m * 8, cudaMemcpyDeviceToHost, s->st));
        CU(cudaMemcpyAsync(out->counts, s->out_cnt.p, (size_t)m * 4, cudaMemcpyDeviceToHost, s->st));
        CU(cudaMemcpyAsync(out->extras, s->out_ext.p, (size_t)m * 4, cudaMemcpyDeviceToHost, s->st));
        CU(cudaMemcpyAsync(out->kmers, s->sel_bytes.p, (size_t)m * k, cudaMemcpyDeviceToHost, s->st));
        CU(wait_stream(s, s->st));
        s->stats.d2h_bytes += (size_t)m * (16 + k);
    }
    out->seq_length = 0;                 // counts.rs never adds to total_bases
    out->num_valid_kmers = sum;
    out->format = s->format;
    out->filters.filter_on = 0;
    return FB2_OK;
}
extern "C" int fb2_sketcher_totals(fb2_sketcher *s, uint64_t *total_bases, uint64_t *total_kmers) {
    if (!s) return fb2_fail(FB2_EINVAL, "null handle");
    ON_DEVICE(s->device);
    TRY(flush_all(s));
    TRY(pull_state(s));
    if (s->allcounts) {   // counts.rs:38-43: total_bases is never updated there; the k-mer total is the sum of the counters
        uint64_t sum = 0;
        TRY(allcounts_result(s, nullptr, &sum));
        if (total_bases) *total_bases = 0;
        if (total_kmers) *total_kmers = sum;
        return FB2_OK;
    }
    if (total_bases) *total_bases = s->h_carry->total_bases + s->lines_bases;
    if (total_kmers) *total_kmers = s->total_kmers;
    return FB2_OK;
}

extern "C" void fb2_result_free(fb2_result *r) {
    if (!r) return;
    free(r->hashes); free(r->counts); free(r->extras); free(r->kmers); free(r->kmer_lens);
    r->hashes = nullptr; r->counts = nullptr; r->extras = nullptr; r->kmers = nullptr; r->kmer_lens = nullptr; r->n = 0;
}

// hostlogic.cpp
uint32_t fb2_threshold_from_hist(const std::vector<uint64_t> &hist, double filter_level);
int fb2_filter_select(const uint32_t *counts, const uint32_t *extras, size_t n, fb2_filter *f, int format,
                      std::vector<uint32_t> &keep, size_t limit);

// Pinned host staging for result read-back (grown on demand).
// Copy `n` bytes with a few host threads (the result slabs are large; one memcpy thread cannot keep
// up with the kernel + D2H pipeline).
static void parallel_memcpy(void *dst, const void *src, size_t n) {
    const size_t min_part = 8u << 20;
    unsigned parts = (unsigned)std::min<size_t>(8, n / min_part);
    if (parts < 2) { memcpy(dst, src, n); return; }
    std::vector<std::thread> th;
    const size_t per = (n / parts + 63) & ~(size_t)63;
    for (unsigned i = 1; i < parts; ++i) {
        const size_t off = per * i, len = i + 1 == parts ? n - off : per;
        th.emplace_back([=] { memcpy((char *)dst + off, (const char *)src + off, len); });
    }
    memcpy(dst, src, per);
    for (auto &t : th) t.join();
}

static int ensure_hres(fb2_sketcher *s, size_t bytes) {
    if (bytes <= s->h_res_cap) return FB2_OK;
    if (s->h_res) cudaFreeHost(s->h_res);
    s->h_res = nullptr; s->h_res_cap = 0;
    const size_t want = std::max<size_t>(bytes + bytes / 4 + 4096, 256u << 10);   // (one allocation serves a usual sketch: pinning calls are slow)
    CU(cudaHostAlloc((void **)&s->h_res, want, cudaHostAllocDefault));
    s->h_res_cap = want;
    return FB2_OK;
}

// Sort the table and export the kept entries as device SoA (out_hash/out_cnt/out_ext/out_kmer/out_posx).
static int export_sorted(fb2_sketcher *s, uint32_t *keep_out) {
    uint32_t n = 0;
    TRY(sort_table(s, &n));
    const uint32_t keep = s->h_state->keep_count;
    if (keep) {
        TRY(s->out_hash.ensure((size_t)keep * 8)); TRY(s->out_kmer.ensure((size_t)keep * 8 * s->kw));
        TRY(s->out_posx.ensure((size_t)keep * 8));
        TRY(s->out_cnt.ensure((size_t)keep * 4)); TRY(s->out_ext.ensure((size_t)keep * 4));
        launch_export(s->sort_keys.as<unsigned long long>(), s->sort_slots.as<uint32_t>(), keep, s->tab[s->cur].view(),
                      s->out_hash.as<unsigned long long>(), s->out_cnt.as<uint32_t>(), s->out_ext.as<uint32_t>(),
                      s->out_kmer.as<unsigned long long>(), s->out_posx.as<unsigned long long>(), s->st);
        s->stats.kernel_launches++;
    }
    *keep_out = keep;
    return FB2_OK;
}

// Result arrays.  A Scaled sketch of a whole genome has millions of entries (C4: 2.96 M, 139 MB): filling freshly
// mapped 4 KiB pages was most of its end-of-stream time (26 of 30 ms, mostly page faults, and another 10 ms when the
// caller freed them).  Large arrays are 2 MiB-aligned and marked for transparent huge pages; free() takes both kinds.
static void *result_alloc(size_t n) {
    if (n >= ((size_t)8 << 20)) {
        const size_t a = (size_t)2 << 20, r = (n + a - 1) & ~(a - 1);
        void *p = aligned_alloc(a, r);
        if (p) { madvise(p, r, MADV_HUGEPAGE); return p; }
    }
    return malloc(std::max<size_t>(1, n));
}
// Large results leave the pinned staging block in pieces: every piece has its own device-to-host copy and event, and a
// few host threads copy their stripe of a piece as soon as it has arrived, so the link and the host copies overlap.
struct ResultPiece { void *dst; size_t off, len; };

// Bring m exported rows to the host as an fb2_result: rows idx[0..m) (device indices) or the first m.
static int collect_rows(fb2_sketcher *s, const uint32_t *h_idx, uint32_t m, fb2_result *out, bool idx_on_device = false) {
    const size_t stride = (size_t)s->k;   // pushed k-mers may be longer: fixed up below
    out->n = m;
    out->hashes = (uint64_t *)result_alloc((size_t)m * 8);
    out->counts = (uint32_t *)result_alloc((size_t)m * 4);
    out->extras = (uint32_t *)result_alloc((size_t)m * 4);
    if (!out->hashes || !out->counts || !out->extras) { fb2_result_free(out); return fb2_fail(FB2_ENOMEM, "malloc"); }
    // Entries that came through push() carry the caller's bytes (arena) and need their codes / position
    // words on the host; everything else is expanded to ASCII on the device and copied as one block.
    const bool has_arena = !s->arena.empty();
    std::vector<unsigned long long> h_kmer(has_arena ? m : 0), h_posx(has_arena ? m : 0);
    const uint8_t *h_bytes = nullptr;
    if (m) {
        TRY(s->sel_hash.ensure((size_t)m * 8)); TRY(s->sel_kmer.ensure((size_t)m * 8)); TRY(s->sel_posx.ensure((size_t)m * 8));
        TRY(s->sel_cnt.ensure((size_t)m * 4)); TRY(s->sel_ext.ensure((size_t)m * 4)); TRY(s->sel_bytes.ensure((size_t)m * stride));
        const uint32_t *d_idx = idx_on_device ? s->sel_idx.as<uint32_t>() : nullptr;   // left there by filter_select_kernel
        if (h_idx && !idx_on_device) {
            TRY(s->sel_idx.ensure((size_t)m * 4));
            CU(cudaMemcpyAsync(s->sel_idx.p, h_idx, (size_t)m * 4, cudaMemcpyHostToDevice, s->st));
            s->stats.h2d_bytes += (size_t)m * 4;
            d_idx = s->sel_idx.as<uint32_t>();
        }
        launch_select_rows(d_idx, m, s->k, (uint32_t)stride, s->kw, s->out_hash.as<unsigned long long>(), s->out_cnt.as<uint32_t>(),
                           s->out_ext.as<uint32_t>(), s->out_kmer.as<unsigned long long>(), s->out_posx.as<unsigned long long>(),
                           s->sel_hash.as<unsigned long long>(), s->sel_cnt.as<uint32_t>(), s->sel_ext.as<uint32_t>(),
                           s->sel_kmer.as<unsigned long long>(), s->sel_posx.as<unsigned long long>(),
                           s->sel_bytes.as<uint8_t>(), s->st);
        s->stats.kernel_launches++;
        // one pinned staging block: hash | cnt | ext | bytes [| kmer | posx]
        const size_t o_hash = 0, o_cnt = (size_t)m * 8, o_ext = (size_t)m * 12, o_bytes = (size_t)m * 16;
        const size_t o_kmer = (o_bytes + (size_t)m * stride + 7) & ~(size_t)7, o_posx = o_kmer + (size_t)m * 8;
        const size_t total = has_arena ? o_posx + (size_t)m * 8 : o_bytes + (size_t)m * stride;
        TRY(ensure_hres(s, total));
        if (!has_arena && total >= ((size_t)32 << 20) && !getenv("FB2_NO_PIPED_RESULT")) {
            // the pipelined way out (see ResultPiece)
            out->kmers = (uint8_t *)result_alloc((size_t)m * stride);
            if (!out->kmers) { fb2_result_free(out); return fb2_fail(FB2_ENOMEM, "malloc"); }
            out->kmer_stride = (uint32_t)stride;
            std::vector<ResultPiece> pieces;
            const size_t piece = (size_t)16 << 20;
            auto add = [&](void *dst, const void *dev, size_t off, size_t len) -> int {
                for (size_t a = 0; a < len; a += piece) {
                    const size_t n = std::min(piece, len - a);
                    CU(cudaMemcpyAsync(s->h_res + off + a, (const char *)dev + a, n, cudaMemcpyDeviceToHost, s->st));
                    if (pieces.size() >= s->ev_piece.size()) {
                        cudaEvent_t e;
                        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                        s->ev_piece.push_back(e);
                    }
                    CU(cudaEventRecord(s->ev_piece[pieces.size()], s->st));
                    pieces.push_back(ResultPiece{(char *)dst + a, off + a, n});
                }
                return FB2_OK;
            };
            TRY(add(out->hashes, s->sel_hash.p, o_hash, (size_t)m * 8));
            TRY(add(out->counts, s->sel_cnt.p, o_cnt, (size_t)m * 4));
            TRY(add(out->extras, s->sel_ext.p, o_ext, (size_t)m * 4));
            TRY(add(out->kmers, s->sel_bytes.p, o_bytes, (size_t)m * stride));
            const unsigned T = std::max(1u, std::min(12u, std::thread::hardware_concurrency()));
            std::atomic<int> bad{0};
            auto work = [&](unsigned t) {
                if (t) cudaSetDevice(s->device);
                for (size_t q = 0; q < pieces.size(); ++q) {
                    if (cudaEventSynchronize(s->ev_piece[q]) != cudaSuccess) { bad.store(1); return; }
                    const ResultPiece &pc = pieces[q];
                    const size_t per = (pc.len / T + 4095) & ~(size_t)4095, a = std::min(pc.len, per * t), b = std::min(pc.len, a + per);
                    if (b > a) memcpy((char *)pc.dst + a, s->h_res + pc.off + a, b - a);
                }
            };
            std::vector<std::thread> th;
            for (unsigned t = 1; t < T; ++t) th.emplace_back(work, t);
            work(0);
            for (auto &x : th) x.join();
            if (bad.load()) { fb2_result_free(out); return fb2_fail(FB2_ECUDA, "cudaEventSynchronize failed (result pieces)"); }
            s->stats.d2h_bytes += total;
            out->seq_length = s->h_carry->total_bases + s->lines_bases;
            out->num_valid_kmers = s->total_kmers;
            out->format = s->format;
            out->filters.filter_on = 0;
            return FB2_OK;
        }
        CU(cudaMemcpyAsync(s->h_res + o_hash, s->sel_hash.p, (size_t)m * 8, cudaMemcpyDeviceToHost, s->st));
        CU(cudaMemcpyAsync(s->h_res + o_cnt, s->sel_cnt.p, (size_t)m * 4, cudaMemcpyDeviceToHost, s->st));
        CU(cudaMemcpyAsync(s->h_res + o_ext, s->sel_ext.p, (size_t)m * 4, cudaMemcpyDeviceToHost, s->st));
        CU(cudaMemcpyAsync(s->h_res + o_bytes, s->sel_bytes.p, (size_t)m * stride, cudaMemcpyDeviceToHost, s->st));
        if (has_arena) {
            CU(cudaMemcpyAsync(s->h_res + o_kmer, s->sel_kmer.p, (size_t)m * 8, cudaMemcpyDeviceToHost, s->st));
            CU(cudaMemcpyAsync(s->h_res + o_posx, s->sel_posx.p, (size_t)m * 8, cudaMemcpyDeviceToHost, s->st));
        }
        CU(wait_stream(s, s->st));
        s->stats.d2h_bytes += total;
        parallel_memcpy(out->hashes, s->h_res + o_hash, (size_t)m * 8);
        parallel_memcpy(out->counts, s->h_res + o_cnt, (size_t)m * 4);
        parallel_memcpy(out->extras, s->h_res + o_ext, (size_t)m * 4);
        h_bytes = s->h_res + o_bytes;
        if (has_arena) {
            memcpy(h_kmer.data(), s->h_res + o_kmer, (size_t)m * 8);
            memcpy(h_posx.data(), s->h_res + o_posx, (size_t)m * 8);
        }
    }
    size_t ostride = stride;
    if (has_arena)
        for (uint32_t i = 0; i < m; ++i)
            if (h_posx[i] & (1ULL << 8)) ostride = std::max(ostride, s->arena[(size_t)h_kmer[i]].size());
    out->kmer_stride = (uint32_t)ostride;
    if (!has_arena) {
        out->kmers = (uint8_t *)result_alloc((size_t)m * ostride);
        if (!out->kmers) { fb2_result_free(out); return fb2_fail(FB2_ENOMEM, "malloc"); }
        if (m) parallel_memcpy(out->kmers, h_bytes, (size_t)m * stride);
    } else {
        out->kmers = (uint8_t *)calloc(std::max<size_t>(1, (size_t)m * ostride), 1);
        out->kmer_lens = (uint32_t *)malloc(std::max<size_t>(1, (size_t)m * 4));
        if (!out->kmers || !out->kmer_lens) { fb2_result_free(out); return fb2_fail(FB2_ENOMEM, "malloc"); }
        for (uint32_t i = 0; i < m; ++i) {
            uint8_t *dst = out->kmers + (size_t)i * ostride;
            if (h_posx[i] & (1ULL << 8)) {
                const std::string &a = s->arena[(size_t)h_kmer[i]];
                memcpy(dst, a.data(), a.size());
                out->kmer_lens[i] = (uint32_t)a.size();
            } else { memcpy(dst, h_bytes + (size_t)i * stride, stride); out->kmer_lens[i] = (uint32_t)stride; }
        }
    }
    out->seq_length = s->h_carry->total_bases + s->lines_bases;
    out->num_valid_kmers = s->total_kmers;
    out->format = s->format;
    out->filters.filter_on = 0;
    return FB2_OK;
}

extern "C" int fb2_sketcher_result(fb2_sketcher *s, fb2_result *out) {
    if (!s || !out) return fb2_fail(FB2_EINVAL, "null argument");
    ON_DEVICE(s->device);
    memset(out, 0, sizeof(*out));
    TRY(flush_all(s));
    if (s->allcounts) return allcounts_result(s, out, nullptr);
    uint32_t keep = 0;
    TRY(export_sorted(s, &keep));
    return collect_rows(s, nullptr, keep, out);
}

// to_vec -> filter_counts -> process_post_filter (lib.rs:78-82), moving only what is needed:
// counts/extras of every entry for the filter decisions, then the surviving rows.
extern "C" int fb2_sketcher_sketch(fb2_sketcher *s, const char *name, const fb2_params *p, const fb2_filter *f,
                                   fb2_result *out) {
    if (!s || !p || !f || !out) return fb2_fail(FB2_EINVAL, "null argument");
    ON_DEVICE(s->device);
    memset(out, 0, sizeof(*out));
    const bool trace = getenv("FB2_TRACE_SKETCH") != nullptr;   // where the end-of-stream time goes (stderr)
    auto now = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = trace ? now() : 0.0;
    TRY(flush_all(s));
    if (s->allcounts) {   // to_vec, then the generic tail of sketch_stream (lib.rs:78-82): filters if they are on
        TRY(allcounts_result(s, out, nullptr));
        fb2_filter ff = *f;
        if (ff.filter_on < 0) ff.filter_on = s->format == FB2_FORMAT_FASTQ ? 1 : 0;
        if (s->format == FB2_FORMAT_UNKNOWN && f->filter_on < 0) { fb2_result_free(out); return fb2_fail(FB2_EEMPTY, "Should have got a type"); }
        if (ff.filter_on == 1) {
            out->format = s->format;
            const int rcf = fb2_filter_counts(out, &ff);
            if (rcf != FB2_OK) { fb2_result_free(out); return rcf; }
        }
        out->filters = ff;
        return FB2_OK;
    }
    const double t1 = trace ? now() : 0.0;
    uint32_t keep = 0;
    TRY(export_sorted(s, &keep));
    const double t2 = trace ? now() : 0.0;
    fb2_filter ff = *f;
    if (ff.filter_on < 0) {  // lib.rs:71-76
        if (s->format == FB2_FORMAT_FASTA) ff.filter_on = 0;
        else if (s->format == FB2_FORMAT_FASTQ) ff.filter_on = 1;
        else return fb2_fail(FB2_EEMPTY, "Should have got a type");
    }
    std::vector<uint32_t> sel;
    const uint32_t *idx = nullptr;
    uint32_t m = keep;
    bool idx_on_device = false;
    if (ff.filter_on == 1 && keep && !getenv("FB2_HOST_FILTER")) {
        // the filters on the device (table.cu, filter_pass / filter_select): per-entry decisions in parallel, the
        // host only walks the histogram of counts (guess_filter_threshold); counts past FILTER_HCAP: host path below
        constexpr uint32_t FILTER_HCAP = 1u << 16;
        const bool strand = ff.strand_filter > 0.0, err = ff.err_filter > 0.0;
        TRY(s->d_fhist.ensure((size_t)FILTER_HCAP * 4 + 16)); TRY(s->d_fok.ensure(keep));
        uint32_t *d_hist = s->d_fhist.as<uint32_t>(), *d_meta = d_hist + FILTER_HCAP;
        bool device_ok = true;
        if (strand || err) {
            CU(cudaMemsetAsync(d_hist, 0, (size_t)FILTER_HCAP * 4 + 16, s->st));
            launch_filter_pass(s->out_cnt.as<uint32_t>(), s->out_ext.as<uint32_t>(), keep, strand ? 1 : 0, ff.strand_filter,
                               err ? 1 : 0, d_hist, FILTER_HCAP, s->d_fok.as<uint8_t>(), d_meta, s->st);
            s->stats.kernel_launches++;
        }
        if (err) {
            TRY(ensure_hres(s, (size_t)FILTER_HCAP * 4 + 16));
            uint32_t *h_meta = (uint32_t *)s->h_res, *h_hist = h_meta + 4;
            CU(cudaMemcpyAsync(h_meta, d_meta, 8, cudaMemcpyDeviceToHost, s->st));
            CU(cudaMemcpyAsync(h_hist, d_hist, 4096 * 4, cudaMemcpyDeviceToHost, s->st));    // nearly always enough
            CU(wait_stream(s, s->st));
            const uint32_t over = h_meta[0], maxc = h_meta[1];
            if (over) device_ok = false;
            else {
                if (maxc > 4096u) {
                    CU(cudaMemcpyAsync(h_hist, d_hist, (size_t)maxc * 4, cudaMemcpyDeviceToHost, s->st));
                    CU(wait_stream(s, s->st));
                }
                s->stats.d2h_bytes += 8 + (size_t)std::max(maxc, 4096u) * 4;
                std::vector<uint64_t> hist(maxc);
                for (uint32_t c = 0; c < maxc; ++c) hist[c] = h_hist[c];
                const uint32_t cutoff = fb2_threshold_from_hist(hist, ff.err_filter);    // filtering.rs:68-79
                if (ff.has_abun_low) { if (cutoff > ff.abun_low) ff.abun_low = cutoff; }
                else { ff.has_abun_low = 1; ff.abun_low = cutoff; }
            }
        }
        if (device_ok) {
            const uint32_t limit = p->kind == FB2_KIND_MASH ? (uint32_t)std::min<uint64_t>(p->final_size, keep) : keep;
            TRY(s->sel_idx.ensure((size_t)std::max(limit, 1u) * 4));
            launch_filter_select(s->out_cnt.as<uint32_t>(), s->d_fok.as<uint8_t>(), keep, strand ? 1 : 0,
                                 (ff.has_abun_low || ff.has_abun_high) ? 1 : 0, ff.has_abun_low ? ff.abun_low : 0u,
                                 ff.has_abun_high ? ff.abun_high : UINT32_MAX, limit, s->sel_idx.as<uint32_t>(), d_meta, s->st);
            s->stats.kernel_launches++;
            TRY(ensure_hres(s, 16));
            CU(cudaMemcpyAsync(s->h_res, d_meta + 2, 4, cudaMemcpyDeviceToHost, s->st));
            CU(wait_stream(s, s->st));
            m = *(uint32_t *)s->h_res;
            idx_on_device = true;
            if (trace) fprintf(stderr, "sketch(): device filter %.0f us (%u entries -> %u)\n", now() - t2, keep, m);
        }
    }
    if (ff.filter_on == 1 && keep && !idx_on_device) {
        TRY(ensure_hres(s, (size_t)keep * 8));
        uint32_t *hc = (uint32_t *)s->h_res, *hx = hc + keep;
        CU(cudaMemcpyAsync(hc, s->out_cnt.p, (size_t)keep * 4, cudaMemcpyDeviceToHost, s->st));
        CU(cudaMemcpyAsync(hx, s->out_ext.p, (size_t)keep * 4, cudaMemcpyDeviceToHost, s->st));
        CU(wait_stream(s, s->st));
        s->stats.d2h_bytes += (size_t)keep * 8;
        const double t3 = trace ? now() : 0.0;
        TRY(fb2_filter_select(hc, hx, keep, &ff, s->format, sel, p->kind == FB2_KIND_MASH ? (size_t)p->final_size : SIZE_MAX));
        if (trace) fprintf(stderr, "sketch(): counts D2H %.0f us, host filter %.0f us (%u entries)\n", t3 - t2, now() - t3, keep);
        m = (uint32_t)sel.size();
        idx = sel.data();
    }
    const double t4 = trace ? now() : 0.0;
    if (p->kind == FB2_KIND_MASH) {  // process_post_filter (mod.rs:115-128)
        if (m > p->final_size) m = (uint32_t)p->final_size;
        if (!p->no_strict && m < p->final_size)
            return fb2_fail(FB2_ETOOFEW, std::string(name ? name : "") + " had too few kmers (" + std::to_string(m) +
                                             ") to sketch");
    }
    const int rc = collect_rows(s, idx, m, out, idx_on_device);
    if (trace) fprintf(stderr, "sketch(): flush/settle %.0f us, sort+export %.0f us, filter stage %.0f us, collect %.0f us\n",
                       t1 - t0, t2 - t1, t4 - t2, now() - t4);
    if (rc == FB2_OK) out->filters = ff;
    return rc;
}

// Debug/inspection: symbol regions of the most recent chunk (for tests of the parse kernels).
extern "C" int fb2_sketcher_debug_symbols(fb2_sketcher *s, uint32_t *geom7, uint32_t *counts, size_t counts_cap,
                                          uint8_t *sym, size_t sym_cap) {
    if (!s || !geom7) return fb2_fail(FB2_EINVAL, "null argument");
    ON_DEVICE(s->device);
    CU(wait_stream(s, s->st));
    const int par = s->par ^ 1;   // the chunk that ran last
    const ChunkGeom g = s->last_geom;
    const uint32_t gv[7] = {g.len, g.n_tiles, g.st_tiles, g.n_st, g.st_bytes, g.region_stride, g.hash_tiles};
    memcpy(geom7, gv, sizeof gv);
    if (counts && counts_cap >= g.n_st) CU(cudaMemcpy(counts, s->d_rcount[par].p, (size_t)g.n_st * 4, cudaMemcpyDeviceToHost));
    const size_t need = (size_t)SYM_FRONT + (size_t)g.n_st * g.region_stride + HASH_W;
    if (sym && sym_cap >= need) CU(cudaMemcpy(sym, s->d_sym[par].p, need, cudaMemcpyDeviceToHost));
    return FB2_OK;
}

// Test hook: add (add_count, add_extra) to the 64-bit totals the table keeps for `hash` (which must be present).
extern "C" int fb2_sketcher_debug_bump(fb2_sketcher *s, uint64_t hash, uint64_t add_count, uint64_t add_extra) {
    if (!s) return fb2_fail(FB2_EINVAL, "null handle");
    ON_DEVICE(s->device);
    TRY(flush_all(s));
    TRY(pull_state(s));
    unsigned int *found = &((SketchState *)s->d_state.p)->gather_count;   // scratch word of the state block
    launch_debug_bump(s->tab[s->cur].view(), hash, add_count, add_extra, found, s->st);
    TRY(pull_state(s));
    if (!s->h_state->gather_count) return fb2_fail(FB2_EINVAL, "hash not in the table");
    return FB2_OK;
}

// ---- a whole small stream in one go ---------------------------------------------------------------------------------
// sketch_stream (lib.rs:51-94) of a stream that fits ONE chunk, on an untouched handle, when the filters resolve to off
// (a FASTA file with the command line's defaults: the heap is final_size then): the general path would walk
// through ~24 host round trips (state pulls between the steps of the first-chunk ramp, the prune, the sort) for 0.2 ms
// of kernels.  Here everything up to the gathered keys is queued without looking:
//     H2D | parse | hash under the provisional threshold | guarded absorb | gather         -> pull A
//     bucket sort | select                                                                 -> pull B
//     export | rows | D2H                                                                  -> done
// The provisional threshold T0 = 16 s 2^64 / len is hash_range's (there with the chunk's symbol count N <= len, so
// this one is a little lower) and valid under the same condition, checked at pull A: at least s distinct keys lie at
// or below it.  Returns 1 when the stream is not of that kind or anything unusual shows up (candidates past the log,
// non-uniform keys, a threshold that turned out too low): the caller then resets the handle and takes the general path,
// which decides.
int fb2_sketcher_sketch_small(fb2_sketcher *s, const uint8_t *bytes, size_t len, const char *name, const fb2_params *p,
                              const fb2_filter *f, fb2_result *out) {
    if (!s || !bytes || !p || !f || !out) return 1;
    if (!s->fresh || s->allcounts || s->scaled || s->k > 32 || s->stream_open || getenv("FB2_NO_SMALL_PATH")) return 1;
    if (len < (64u << 10) || len > s->chunk_bytes || len >= (1ull << 31)) return 1;
    const uint8_t c0 = bytes[0];
    if (c0 != '>' && c0 != '@') return 1;
    const int format = c0 == '>' ? FB2_FORMAT_FASTA : FB2_FORMAT_FASTQ;
    if (format == FB2_FORMAT_FASTQ && getenv("FB2_HOST_STRIP") && *getenv("FB2_HOST_STRIP") == '1') return 1;
    const int filter_on = f->filter_on >= 0 ? f->filter_on : (format == FB2_FORMAT_FASTQ ? 1 : 0);   // lib.rs:71-76
    if (filter_on != 0 || p->kind != FB2_KIND_MASH) return 1;
    ON_DEVICE(s->device);
    memset(out, 0, sizeof(*out));
    s->fresh = false;
    s->format = format;
    apply_finish_hint(s);                       // heap = final_size when the caller said so (fb2_sketcher_hint_finish)
    const uint64_t size = s->size, want = 16ull * size;
    if (size == 0 || want + want / 8 > s->log_cap / 2 || len < 16 * want) return 1;
    if ((uint64_t)s->tab[s->cur].cap / 4 * 3 < 2 * want) return 1;   // room for every candidate as a new key
    // begin_stream's carry, the provisional threshold
    ParseCarry *c = s->h_carry;
    c->state = format == FB2_FORMAT_FASTA ? 1u : 0u;
    c->prev1 = c->prev2 = '\n';
    c->raw_total = 0; c->n_records = 0; c->first_bad_pos = ~0ULL; c->last_sig = 0; c->error = 0; c->len_bad_pos = ~0ULL;
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 3; ++b) c->last_nl[a][b] = NL_NONE;
    s->h_state->threshold = (~0ULL / (unsigned long long)len) * want;
    CU(cudaMemcpyAsync(s->d_carry.p, c, sizeof(ParseCarry), cudaMemcpyHostToDevice, s->st));
    CU(cudaMemcpyAsync(s->d_state.p, s->h_state, sizeof(SketchState), cudaMemcpyHostToDevice, s->st));
    s->stream_open = true; s->range_mode = false; s->strip_on = false;
    s->tail_host.clear();
    note_tail(s, bytes, len);
    // the bytes
    const size_t bufsz = (len + 4095) / 4096 * 4096 + 64;
    TRY(s->d_raw[0].ensure(bufsz));
    CU(cudaMemcpyAsync(s->d_raw[0].p, bytes, len, cudaMemcpyHostToDevice, s->st));
    s->stats.h2d_bytes += len + sizeof(ParseCarry) + sizeof(SketchState);
    // parse (run_chunk stops after the parse kernels), hash, absorb, gather
    s->small_mode = true;
    const int rcp = run_chunk(s, s->d_raw[0].as<uint8_t>(), (uint32_t)len, format == FB2_FORMAT_FASTA ? MODE_FASTA : MODE_FASTQ, -1);
    s->small_mode = false;
    if (rcp != FB2_OK) return rcp;
    const int par = s->small_par;
    const ChunkGeom g = s->last_geom;
    SketchState *dst = (SketchState *)s->d_state.p;
    ParseCarry *dc = (ParseCarry *)s->d_carry.p;
    LaunchSlot *slot = dev_slot(s, par);
    CU(cudaMemsetAsync(slot, 0, sizeof(LaunchSlot), s->st));
    launch_hash(s->k, s->d_sym[par].as<uint8_t>(), g, 0, g.n_st, s->d_rcount[par].as<uint32_t>(), dc, s->small_ord, dst, slot,
                log_view(s, par), s->prm.hash_seed, 31u, piece_plan(s, par), s->st);
    launch_absorb_guarded(log_view(s, par), slot, s->tab[s->cur].view(), dst, dc, (uint32_t)want, s->st);
    const uint32_t occ_ub = s->tab[s->cur].cap + 1;
    TRY(ensure_sort(s, occ_ub));
    TRY(s->d_bins.ensure(3 * 4096 * sizeof(uint32_t)));
    unsigned long long *keys = s->sort_keys.as<unsigned long long>(), *tkeys = s->sort_tkeys.as<unsigned long long>();
    uint32_t *slots = s->sort_slots.as<uint32_t>(), *tslots = s->sort_tslots.as<uint32_t>();
    launch_gather(s->tab[s->cur].view(), dst, tkeys, tslots, s->st);
    s->stats.kernel_launches += 6; s->stats.hash_launches++;
    TRY(pull_state(s));                                                                           // ---- pull A
    const LaunchSlot &sl = s->h_state->slot[par];
    s->stats.hash_symbols += s->h_carry->chunk_syms;
    const uint64_t have = (uint64_t)s->h_state->occupied + (s->h_state->has_max_key ? 1u : 0u);
    if (sl.decision != DECIDE_GO || sl.log_count > s->log_cap || have < size) return 1;
    s->total_kmers += sl.launch_kmers;
    s->stream_open = false;
    {
        const int rcv = stream_verdict(s);       // the reader's end-of-input rules (and the last record's length)
        if (rcv != FB2_OK) return rcv;
        // (the verdict may have corrected the carry's totals: the device copy follows, like end_stream's push)
        s->h_carry->state = 0; s->h_carry->prev1 = s->h_carry->prev2 = '\n';
        CU(cudaMemcpyAsync(s->d_carry.p, s->h_carry, sizeof(ParseCarry), cudaMemcpyHostToDevice, s->st));
    }
    const uint32_t n = s->h_state->gather_count;
    const unsigned long long thr = s->h_state->threshold;
    if (n < 2) return 1;
    const uint32_t shift = shift_for_threshold(thr);
    uint32_t *bins = s->d_bins.as<uint32_t>();
    launch_bucket_sort(tkeys, tslots, keys, slots, tkeys, tslots, n, shift, bins, bins + 4096, bins + 8192, dst, s->st);
    CU(cudaMemcpyAsync(keys, tkeys, (size_t)n * 8, cudaMemcpyDeviceToDevice, s->st));
    CU(cudaMemcpyAsync(slots, tslots, (size_t)n * 4, cudaMemcpyDeviceToDevice, s->st));
    launch_select_keep(keys, n, 0, s->size, s->max_hash, dst, s->st);
    s->stats.kernel_launches += 5;
    TRY(pull_state(s));                                                                           // ---- pull B
    if (s->h_state->gather_count > bucket_cap()) return 1;      // keys too uneven for the bucket sort: the general path sorts by radix
    const uint32_t keep = s->h_state->keep_count;
    if (keep) {
        TRY(s->out_hash.ensure((size_t)keep * 8)); TRY(s->out_kmer.ensure((size_t)keep * 8 * s->kw));
        TRY(s->out_posx.ensure((size_t)keep * 8));
        TRY(s->out_cnt.ensure((size_t)keep * 4)); TRY(s->out_ext.ensure((size_t)keep * 4));
        launch_export(keys, slots, keep, s->tab[s->cur].view(), s->out_hash.as<unsigned long long>(), s->out_cnt.as<uint32_t>(),
                      s->out_ext.as<uint32_t>(), s->out_kmer.as<unsigned long long>(), s->out_posx.as<unsigned long long>(), s->st);
        s->stats.kernel_launches++;
    }
    uint32_t m = keep;
    if (m > p->final_size) m = (uint32_t)p->final_size;      // process_post_filter (mod.rs:115-128)
    if (!p->no_strict && m < p->final_size)
        return fb2_fail(FB2_ETOOFEW, std::string(name ? name : "") + " had too few kmers (" + std::to_string(m) + ") to sketch");
    const int rc = collect_rows(s, nullptr, m, out);
    if (rc == FB2_OK) { fb2_filter ff = *f; ff.filter_on = 0; out->filters = ff; }
    return rc;
}

extern "C" int fb2_sketcher_stats(fb2_sketcher *s, fb2_stats *out) {
    if (!s || !out) return fb2_fail(FB2_EINVAL, "null argument");
    ON_DEVICE(s->device);
    timing_resolve(s);
    *out = s->stats;
    return FB2_OK;
}
extern "C" int fb2_sketcher_enable_timing(fb2_sketcher *s, int on) {
    if (!s) return fb2_fail(FB2_EINVAL, "null handle");
    s->timing = on != 0;
    return FB2_OK;
}

// ---- dist --------------------------------------------------------------------------------------------
static cudaError_t upload_large(void *dst, const void *src, size_t bytes);   // below
static unsigned long long dist_max_hash(double scale) {
    // distance.rs:100: u64::MAX / scale.recip() as u64
    const double rec = 1.0 / scale;
    const uint64_t d = rec >= 18446744073709551616.0 ? UINT64_MAX : (uint64_t)rec;
    return d ? UINT64_MAX / d : UINT64_MAX;
}
static int dist_common(const uint64_t *hashes, const uint32_t *lens, size_t n_sk, size_t stride, int32_t device,
                       DevBuf &d_h, DevBuf &d_l) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fb2_fail(FB2_ECUDA, "no CUDA device: finch_b200 has no CPU fallback");
    if (device >= 0) CU(cudaSetDevice(device));   // (the callers hold a DeviceScope)
    TRY(d_h.ensure(std::max<size_t>(8, n_sk * stride * 8)));
    TRY(d_l.ensure(std::max<size_t>(4, n_sk * 4)));
    CU(upload_large(d_h.p, hashes, n_sk * stride * 8));
    CU(cudaMemcpy(d_l.p, lens, n_sk * 4, cudaMemcpyHostToDevice));
    return FB2_OK;
}
// minmer_matrix (distance.rs:344-364): n_sk x n_ref, row-major; sketch i's hashes / counts are
// sk_hashes[sk_off[i] .. sk_off[i + 1]) (ascending), ref_hashes ascending and distinct.
extern "C" int fb2_minmer_matrix(const uint64_t *ref_hashes, size_t n_ref, const uint64_t *sk_hashes, const uint32_t *sk_counts,
                                 const uint64_t *sk_off, size_t n_sk, int32_t *result, int32_t device) {
    if ((n_ref && !ref_hashes) || (n_sk && (!sk_off || !result)) || (n_sk && sk_off[n_sk] && (!sk_hashes || !sk_counts)))
        return fb2_fail(FB2_EINVAL, "null argument");
    if (n_ref >= (1ull << 31) || n_sk >= (1ull << 31)) return fb2_fail(FB2_EINVAL, "matrix too large");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fb2_fail(FB2_ECUDA, "no CUDA device: finch_b200 has no CPU fallback");
    if (!n_sk || !n_ref) return FB2_OK;
    DeviceScope dev_scope_(device);
    const uint64_t total = sk_off[n_sk];
    uint32_t max_len = 0;
    for (size_t i = 0; i < n_sk; ++i) {
        if (sk_off[i + 1] < sk_off[i]) return fb2_fail(FB2_EINVAL, "sketch offsets must ascend");
        max_len = (uint32_t)std::max<uint64_t>(max_len, std::min<uint64_t>(sk_off[i + 1] - sk_off[i], 0xFFFFFFFFull));
    }
    DevBuf d_ref, d_h, d_c, d_o, d_r;
    int rc = FB2_OK;
    do {
        if ((rc = d_ref.ensure(n_ref * 8)) != FB2_OK) break;
        if ((rc = d_h.ensure(std::max<uint64_t>(1, total) * 8)) != FB2_OK) break;
        if ((rc = d_c.ensure(std::max<uint64_t>(1, total) * 4)) != FB2_OK) break;
        if ((rc = d_o.ensure((n_sk + 1) * 8)) != FB2_OK) break;
        if ((rc = d_r.ensure(n_sk * n_ref * 4)) != FB2_OK) break;
        cudaMemcpy(d_ref.p, ref_hashes, n_ref * 8, cudaMemcpyHostToDevice);
        if (total) { cudaMemcpy(d_h.p, sk_hashes, total * 8, cudaMemcpyHostToDevice); cudaMemcpy(d_c.p, sk_counts, total * 4, cudaMemcpyHostToDevice); }
        cudaMemcpy(d_o.p, sk_off, (n_sk + 1) * 8, cudaMemcpyHostToDevice);
        cudaMemset(d_r.p, 0, n_sk * n_ref * 4);
        launch_minmer_matrix(d_ref.as<unsigned long long>(), (uint32_t)n_ref, d_h.as<unsigned long long>(), d_c.as<uint32_t>(),
                             d_o.as<unsigned long long>(), (uint32_t)n_sk, max_len, d_r.as<int32_t>(), 0);
        const cudaError_t e = cudaMemcpy(result, d_r.p, n_sk * n_ref * 4, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fb2_fail(FB2_ECUDA, cudaGetErrorString(e));
    } while (0);
    d_ref.release(); d_h.release(); d_c.release(); d_o.release(); d_r.release();
    return rc;
}
extern "C" int fb2_dist_batch(const uint64_t *hashes, const uint32_t *lens, size_t n_sk, size_t stride, double scale,
                              const uint32_t *q_idx, const uint32_t *r_idx, size_t n_pairs, fb2_pair_out *out,
                              int32_t device) {
    if ((!hashes && n_sk && stride) || !lens || (n_pairs && (!q_idx || !r_idx || !out)))
        return fb2_fail(FB2_EINVAL, "null argument");
    for (size_t i = 0; i < n_sk; ++i) if (lens[i] > stride) return fb2_fail(FB2_EINVAL, "sketch length exceeds stride");
    for (size_t i = 0; i < n_pairs; ++i)
        if (q_idx[i] >= n_sk || r_idx[i] >= n_sk) return fb2_fail(FB2_EINVAL, "pair index out of range");
    DeviceScope dev_scope_(-1);   // dist_common may select `device`: the caller's device comes back on return
    DevBuf d_h, d_l, d_q, d_r, d_o;
    int rc = dist_common(hashes, lens, n_sk, stride, device, d_h, d_l);
    if (rc == FB2_OK && n_pairs) {
        do {
            if ((rc = d_q.ensure(n_pairs * 4)) != FB2_OK) break;
            if ((rc = d_r.ensure(n_pairs * 4)) != FB2_OK) break;
            if ((rc = d_o.ensure(n_pairs * sizeof(fb2_pair_out))) != FB2_OK) break;
            cudaMemcpy(d_q.p, q_idx, n_pairs * 4, cudaMemcpyHostToDevice);
            cudaMemcpy(d_r.p, r_idx, n_pairs * 4, cudaMemcpyHostToDevice);
            launch_dist_pairs(d_h.as<unsigned long long>(), d_l.as<uint32_t>(), (uint32_t)stride, d_q.as<uint32_t>(),
                              d_r.as<uint32_t>(), n_pairs, scale > 0.0, scale > 0.0 ? dist_max_hash(scale) : 0,
                              d_o.as<fb2_pair_out>(), 0);
            cudaError_t e = cudaMemcpy(out, d_o.p, n_pairs * sizeof(fb2_pair_out), cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) rc = fb2_fail(FB2_ECUDA, cudaGetErrorString(e));
        } while (0);
    }
    d_h.release(); d_l.release(); d_q.release(); d_r.release(); d_o.release();
    return rc;
}
static thread_local double g_dist_kernel_ms = 0.0;

// Host (pageable) -> device for the large inputs of the dist calls.  A plain cudaMemcpy stages pageable memory through
// the driver's own bounce buffer on ONE thread (~10 GB/s: 80 ms for the 800 MB matrix of C5); here four threads each
// own a pinned 8 MiB buffer and a stream and copy their share of the pieces, so the link and the host copies overlap.
static cudaError_t upload_large(void *dst, const void *src, size_t bytes) {
    const size_t piece = (size_t)8 << 20;
    if (bytes < 8 * piece || getenv("FB2_NO_STAGED_UPLOAD")) return cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
    const unsigned T = std::max(1u, std::min(4u, std::thread::hardware_concurrency()));
    int dev = 0;
    cudaGetDevice(&dev);
    std::atomic<size_t> next{0};
    std::atomic<int> err{(int)cudaSuccess};
    const size_t n_pieces = (bytes + piece - 1) / piece;
    auto work = [&](unsigned t) {
        if (t) cudaSetDevice(dev);
        uint8_t *pin = nullptr;
        cudaStream_t st = nullptr;
        cudaError_t e = cudaHostAlloc((void **)&pin, piece, cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        while (e == cudaSuccess) {
            const size_t q = next.fetch_add(1);
            if (q >= n_pieces) break;
            const size_t off = q * piece, n = std::min(piece, bytes - off);
            memcpy(pin, (const char *)src + off, n);
            e = cudaMemcpyAsync((char *)dst + off, pin, n, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);          // the buffer is reused right away
        }
        if (e != cudaSuccess) err.store((int)e);
        if (st) cudaStreamDestroy(st);
        if (pin) cudaFreeHost(pin);
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < T; ++t) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    if (err.load() != (int)cudaSuccess) {   // (e.g. no pinned memory to be had): the plain copy decides
        cudaGetLastError();
        return cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
    }
    return cudaSuccess;
}
extern "C" double fb2_dist_last_kernel_ms(void) { return g_dist_kernel_ms; }

extern "C" int fb2_dist_all_pairs(const uint64_t *hashes, const uint32_t *lens, size_t n_sk, size_t stride,
                                  double scale, size_t q0, size_t q1, fb2_pair_out *out, int32_t device) {
    if ((!hashes && n_sk && stride) || !lens || q0 > q1 || q1 > n_sk) return fb2_fail(FB2_EINVAL, "bad argument");
    for (size_t i = 0; i < n_sk; ++i) if (lens[i] > stride) return fb2_fail(FB2_EINVAL, "sketch length exceeds stride");
    const uint64_t n_pairs = (uint64_t)(q1 - q0) * n_sk;
    if (n_pairs && !out) return fb2_fail(FB2_EINVAL, "null output");
    DeviceScope dev_scope_(-1);   // dist_common may select `device`: the caller's device comes back on return
    DevBuf d_h, d_l, d_o[2];
    fb2_pair_out *h_pin[2] = {nullptr, nullptr};
    cudaStream_t st_k = nullptr, st_c = nullptr;
    cudaEvent_t ev_k[2] = {nullptr, nullptr}, ev_c[2] = {nullptr, nullptr};
    int rc = dist_common(hashes, lens, n_sk, stride, device, d_h, d_l);
    if (rc == FB2_OK && n_pairs) {
        // Slabs of whole query rows through a 3-stage pipeline: kernel (slab k) | D2H into pinned staging
        // (slab k-1) | host copy into the caller's (pageable) array (slab k-2).
        const uint64_t rows = std::max<uint64_t>(1, (1ull << 22) / std::max<size_t>(1, n_sk));
        const uint64_t slab_pairs = std::min<uint64_t>(n_pairs, rows * n_sk);
        // short query sketches (the usual n <= 1000): shared-memory tiled kernel; otherwise warp-per-pair
        uint32_t max_qlen = 0;
        for (size_t q = q0; q < q1; ++q) max_qlen = std::max(max_qlen, lens[q]);
        const bool tiled = max_qlen <= dist_tile_max_len() && !getenv("FB2_DIST_NO_TILE");
        const int scaled = scale > 0.0;
        const unsigned long long max_hash = scaled ? dist_max_hash(scale) : 0;
        auto cu = [&](cudaError_t e) { if (e != cudaSuccess && rc == FB2_OK) rc = fb2_fail(FB2_ECUDA, cudaGetErrorString(e)); };
        cu(cudaStreamCreateWithFlags(&st_k, cudaStreamNonBlocking));
        cu(cudaStreamCreateWithFlags(&st_c, cudaStreamNonBlocking));
        // a caller-provided PINNED result array (cudaHostAlloc / cudaHostRegister) takes the D2H copies directly
        bool out_pinned = false;
        {
            cudaPointerAttributes attr;
            if (cudaPointerGetAttributes(&attr, out) == cudaSuccess && attr.type == cudaMemoryTypeHost) out_pinned = true;
            else cudaGetLastError();   // clear the error an unregistered pointer may leave on older runtimes
        }
        for (int i = 0; i < 2 && rc == FB2_OK; ++i) {
            rc = d_o[i].ensure(slab_pairs * sizeof(fb2_pair_out));
            if (rc == FB2_OK && !out_pinned) cu(cudaHostAlloc((void **)&h_pin[i], slab_pairs * sizeof(fb2_pair_out), cudaHostAllocDefault));
            cu(cudaEventCreateWithFlags(&ev_k[i], cudaEventDisableTiming));
            cu(cudaEventCreateWithFlags(&ev_c[i], cudaEventDisableTiming));
        }
        cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
        cu(cudaEventCreate(&ev_t0)); cu(cudaEventCreate(&ev_t1));
        cu(cudaEventRecord(ev_t0, st_k));
        struct Slab { uint64_t q, m; };
        std::vector<Slab> slabs;
        for (uint64_t q = q0; q < q1; q += rows) slabs.push_back({q, std::min<uint64_t>(rows, q1 - q) * n_sk});
        auto drain = [&](size_t k) {   // slab k: wait for its D2H, copy to the caller
            cu(cudaEventSynchronize(ev_c[k & 1]));
            if (rc == FB2_OK && !out_pinned) parallel_memcpy(out + (slabs[k].q - q0) * n_sk, h_pin[k & 1], slabs[k].m * sizeof(fb2_pair_out));
        };
        std::vector<cudaEvent_t> ev_s0(slabs.size(), nullptr), ev_s1(slabs.size(), nullptr);   // per-slab kernel time
        for (size_t k = 0; k < slabs.size() && rc == FB2_OK; ++k) {
            const int b = (int)(k & 1);
            const uint64_t q = slabs[k].q, m = slabs[k].m;
            cu(cudaEventCreate(&ev_s0[k])); cu(cudaEventCreate(&ev_s1[k]));
            cu(cudaEventRecord(ev_s0[k], st_k));
            if (tiled) {
                if (launch_dist_tile(d_h.as<unsigned long long>(), d_l.as<uint32_t>(), (uint32_t)stride, (uint32_t)n_sk,
                                     (uint32_t)q, (uint32_t)(q + m / n_sk), scaled, max_hash, d_o[b].as<fb2_pair_out>(), st_k) != 0)
                    rc = fb2_fail(FB2_ECUDA, "dist_tile_kernel: could not reserve shared memory");
            } else {
                launch_dist_all(d_h.as<unsigned long long>(), d_l.as<uint32_t>(), (uint32_t)stride, (uint32_t)n_sk,
                                (uint32_t)q, m, scaled, max_hash, d_o[b].as<fb2_pair_out>(), st_k);
            }
            cu(cudaEventRecord(ev_s1[k], st_k));
            cu(cudaEventRecord(ev_k[b], st_k));
            if (k >= 2) drain(k - 2);               // host copy of slab k-2 while the kernel of slab k runs; frees h_pin[b]
            cu(cudaStreamWaitEvent(st_c, ev_k[b], 0));
            cu(cudaMemcpyAsync(out_pinned ? (void *)(out + (q - q0) * n_sk) : (void *)h_pin[b], d_o[b].p, m * sizeof(fb2_pair_out),
                               cudaMemcpyDeviceToHost, st_c));
            cu(cudaEventRecord(ev_c[b], st_c));
            cu(cudaStreamWaitEvent(st_k, ev_c[b], 0));                  // the kernel of slab k+2 reuses d_o[b]
        }
        if (ev_t1) cu(cudaEventRecord(ev_t1, st_k));
        if (rc == FB2_OK && slabs.size() >= 2) drain(slabs.size() - 2);
        if (rc == FB2_OK && slabs.size() >= 1) drain(slabs.size() - 1);
        if (st_k) cudaStreamSynchronize(st_k);
        if (st_c) cudaStreamSynchronize(st_c);
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess && rc == FB2_OK) rc = fb2_fail(FB2_ECUDA, cudaGetErrorString(e));
        // device time from the first kernel's start to the last kernel's end (includes waits for free slabs)
        float ms = 0.f;
        double sum = 0.0;
        for (size_t k = 0; k < slabs.size(); ++k) {
            if (rc == FB2_OK && ev_s0[k] && ev_s1[k] && cudaEventElapsedTime(&ms, ev_s0[k], ev_s1[k]) == cudaSuccess) sum += ms;
            if (ev_s0[k]) cudaEventDestroy(ev_s0[k]);
            if (ev_s1[k]) cudaEventDestroy(ev_s1[k]);
        }
        if (rc == FB2_OK) g_dist_kernel_ms = sum;   // kernel time only (the span also holds waits for free slabs)
        if (ev_t0) cudaEventDestroy(ev_t0);
        if (ev_t1) cudaEventDestroy(ev_t1);
    }
    for (int i = 0; i < 2; ++i) {
        if (h_pin[i]) cudaFreeHost(h_pin[i]);
        if (ev_k[i]) cudaEventDestroy(ev_k[i]);
        if (ev_c[i]) cudaEventDestroy(ev_c[i]);
        d_o[i].release();
    }
    if (st_k) cudaStreamDestroy(st_k);
    if (st_c) cudaStreamDestroy(st_c);
    d_h.release(); d_l.release();
    return rc;
}

// ---- dist with the max_distance cut ---------------------------------------------------------------------------
// jaccard bound for mash_distance <= d (distance.rs:36-41): -ln(2j / (1 + j)) / k <= d  <=>  j >= x / (2 - x) with
// x = exp(-d k); placed a hair BELOW the exact value so that rounding can only let extra pairs through (the caller
// applies the exact test).  d >= 1 keeps everything (the distance is clamped to 1).
static double dist_cut_jlow(double max_distance, uint8_t k) {
    if (!(max_distance < 1.0)) return -1.0;
    if (max_distance < 0.0) return 2.0;                       // nothing can pass (distances are >= 0)
    const double x = std::exp(-max_distance * (double)k);
    const double j = x / (2.0 - x);
    return j * (1.0 - 1e-9) - 1e-12;
}
static bool host_passes_cut(const fb2_pair_out &o, double jlow) {
    const uint32_t total = o.i - o.common + o.j;
    const double jac = total == 0u ? 1.0 : (double)o.common / (double)total;
    return jlow < 0.0 || jac >= jlow;
}
// Where the surviving pairs of a row block go: straight into the caller's array when one GPU does all rows (they
// arrive in final order), else into a vector of the block that is placed once the blocks before it are known.
struct HitSink {
    fb2_pair_hit *dst = nullptr;
    size_t cap = 0;
    uint64_t n = 0;                       // hits seen (may exceed cap: the caller then reports the count)
    std::vector<fb2_pair_hit> vec;
    void append(const fb2_pair_hit *p, size_t k) {
        if (dst) {
            if (n < cap) memcpy(dst + n, p, std::min<size_t>(k, cap - (size_t)n) * sizeof(fb2_pair_hit));
        } else vec.insert(vec.end(), p, p + k);
        n += k;
    }
    void push(const fb2_pair_hit &h) { append(&h, 1); }
};
// Rows [qa, qb) against all n_sk sketches on the current device; d_h / d_l hold the matrix.  Hits ascending by (q, r).
static int dist_cut_rows(const unsigned long long *d_h, const uint32_t *d_l, const uint32_t *lens, size_t n_sk, size_t stride,
                         int scaled, unsigned long long max_hash, size_t qa, size_t qb, int skip_self, double jlow,
                         HitSink &out, double *kernel_ms, int *index_only = nullptr) {
    // index_only (in: non-null = do the rows only if the inverted index is chosen; out: 1 = done, 0 = nothing done)
    if (qb <= qa || !n_sk) { if (index_only) *index_only = 0; return FB2_OK; }
    uint32_t max_qlen = 0;
    for (size_t q = qa; q < qb; ++q) max_qlen = std::max(max_qlen, lens[q]);
    const bool tiled = max_qlen <= dist_tile_max_len() && !getenv("FB2_DIST_NO_TILE");
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    size_t rows = env_size("FB2_DIST_ROWS", (size_t)sms * 9 * 2);    // query rows per launch: two 9-row tiles per SM
    uint32_t cap = (uint32_t)env_size("FB2_DIST_HITS_M", 8) << 20;      // hit capacity per launch
    if (!tiled) rows = std::max<size_t>(1, (1u << 22) / n_sk);          // dense fallback: small slabs
    cudaStream_t st = nullptr;
    CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    DevBuf hits, sorted, keys, tkeys, vals, tvals, hist, counter, dense;
    fb2_pair_hit *h_pin = nullptr;
    unsigned int *h_cnt = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = FB2_OK;
    auto cu = [&](cudaError_t e) { if (e != cudaSuccess && rc == FB2_OK) rc = fb2_fail(FB2_ECUDA, cudaGetErrorString(e)); };
    cu(cudaEventCreate(&e0)); cu(cudaEventCreate(&e1));
    cu(cudaHostAlloc((void **)&h_cnt, sizeof(unsigned int), cudaHostAllocDefault));
    if (rc == FB2_OK) rc = counter.ensure(sizeof(unsigned int));
    size_t pin_cap = 0;
    double kms = 0.0;
    std::vector<fb2_pair_out> h_dense;
    // The cut through an inverted index (dist.cu): when a positive jaccard bound rules out every pair without a common
    // hash, large jobs count shared hashes per query from sorted (hash, sketch) postings instead of probing every pair.
    // Chosen when it is applicable (no empty sketch: an empty one has jaccard 1 with everything, distance.rs:119-123;
    // lengths fit the 16-bit counters) and its work -- the sum of squared posting-run lengths -- is far below the
    // pairs x hashes of the tile kernel.  FB2_DIST_INVERTED=0 / 1 forbids / forces it (where applicable).
    bool inverted = false;
    DevBuf p_keys, p_vals, p_tkeys, p_tvals, p_hist, p_off, p_sk, p_run, p_sum;
    {
        const char *inv_env = getenv("FB2_DIST_INVERTED");
        const bool forbid = inv_env && atoi(inv_env) == 0, force = inv_env && atoi(inv_env) != 0;
        uint64_t n_post = 0;
        uint32_t min_len = ~0u, max_len = 0;
        for (size_t i = 0; i < n_sk; ++i) { n_post += lens[i]; min_len = std::min(min_len, lens[i]); max_len = std::max(max_len, lens[i]); }
        const double pairs = (double)(qb - qa) * (double)n_sk;
        size_t mem_free = 0, mem_total = 0;
        if (cudaMemGetInfo(&mem_free, &mem_total) != cudaSuccess) mem_free = 0;
        const bool fits = (double)n_post * 44.0 < (double)mem_free * 0.7;     // postings, their sort buffers, run table
        if (!forbid && jlow > 0.0 && min_len > 0 && max_len <= 65535u && n_post < 0xFFFFFF00ull && n_sk < 0x7FFFFFFFull && fits &&
            (force || pairs >= 2e8)) {
            std::vector<uint32_t> off(n_sk + 1);
            uint64_t run = 0;
            for (size_t i = 0; i < n_sk; ++i) { off[i] = (uint32_t)run; run += lens[i]; }
            off[n_sk] = (uint32_t)run;
            const uint32_t np = (uint32_t)n_post;
            const double t_ix0 = getenv("FB2_TRACE_DIST") ? std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count() : 0.0;
            if ((rc = p_keys.ensure((size_t)np * 8 + 512)) == FB2_OK && (rc = p_tkeys.ensure((size_t)np * 8 + 512)) == FB2_OK &&
                (rc = p_vals.ensure((size_t)np * 4 + 256)) == FB2_OK && (rc = p_tvals.ensure((size_t)np * 4 + 256)) == FB2_OK &&
                (rc = p_hist.ensure((size_t)radix_hist_words(np) * 4)) == FB2_OK && (rc = p_off.ensure((n_sk + 1) * 4)) == FB2_OK &&
                (rc = p_sk.ensure((size_t)np * 4)) == FB2_OK && (rc = p_run.ensure((size_t)np * 8)) == FB2_OK &&
                (rc = p_sum.ensure(8)) == FB2_OK) {
                unsigned long long sum_sq = 0;
                cu(cudaMemcpyAsync(p_off.p, off.data(), (n_sk + 1) * 4, cudaMemcpyHostToDevice, st));
                cu(cudaMemsetAsync(p_sum.p, 0, 8, st));
                cu(cudaEventRecord(e0, st));
                launch_postings_fill(d_h, d_l, (uint32_t)stride, (uint32_t)n_sk, p_off.as<uint32_t>(), p_keys.as<unsigned long long>(),
                                     p_vals.as<uint32_t>(), st);
                launch_radix_sort(p_keys.as<unsigned long long>(), p_vals.as<uint32_t>(), p_tkeys.as<unsigned long long>(),
                                  p_tvals.as<uint32_t>(), np, p_hist.as<uint32_t>(), st);
                launch_postings_runs(p_keys.as<unsigned long long>(), p_vals.as<uint32_t>(), np, p_off.as<uint32_t>(), (uint32_t)n_sk,
                                     p_sk.as<uint32_t>(), p_run.as<unsigned long long>(), p_sum.as<unsigned long long>(), st);
                cu(cudaEventRecord(e1, st));
                cu(cudaMemcpyAsync(&sum_sq, p_sum.p, 8, cudaMemcpyDeviceToHost, st));
                cu(cudaStreamSynchronize(st));
                if (rc == FB2_OK) {
                    float ms = 0.f;
                    if (cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) kms += ms;
                    // bumps of the counting pass (all rows) against probes of the tile kernel (all rows), with a margin
                    inverted = force || (double)sum_sq * 8.0 < (double)n_sk * (double)n_post;
                    if (getenv("FB2_TRACE_DIST"))
                        fprintf(stderr, "dist: %u postings, sum of squared runs %.3e vs %.3e probes: %s (index kernels %.1f ms, with its buffers %.1f ms)\n", np,
                                (double)sum_sq, (double)n_sk * (double)n_post, inverted ? "inverted index" : "tile kernel", ms,
                                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count() - t_ix0);
                }
                // what the counting kernel does not read
                p_keys.release(); p_vals.release(); p_tkeys.release(); p_tvals.release(); p_hist.release();
                if (!inverted) { p_sk.release(); p_run.release(); }
            }
            if (rc != FB2_OK) {   // (e.g. out of device memory after all): the tile kernel does the job
                inverted = false;
                p_keys.release(); p_vals.release(); p_tkeys.release(); p_tvals.release(); p_hist.release(); p_off.release();
                p_sk.release(); p_run.release(); p_sum.release();
                cudaGetLastError();
                rc = FB2_OK;
            }
        }
        if (inverted) rows = env_size("FB2_DIST_ROWS", 16384);
    }
    if (index_only) {
        *index_only = inverted ? 1 : 0;
        if (!inverted) {
            if (kernel_ms) *kernel_ms = kms;
            if (h_cnt) cudaFreeHost(h_cnt);
            if (e0) cudaEventDestroy(e0);
            if (e1) cudaEventDestroy(e1);
            if (st) cudaStreamDestroy(st);
            counter.release(); p_off.release(); p_sum.release();
            return rc;
        }
    }
    const bool trace_rows = getenv("FB2_TRACE_DIST") != nullptr;
    auto now_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_wait = 0.0, t_sortcopy = 0.0, t_insert = 0.0, t_alloc = 0.0;
    for (size_t q = qa; q < qb && rc == FB2_OK;) {
        const size_t m = std::min(rows, qb - q);
        if (tiled || inverted) {
            double t0 = now_ms();
            if ((rc = hits.ensure((size_t)cap * sizeof(fb2_pair_hit))) != FB2_OK) break;
            if ((rc = sorted.ensure((size_t)cap * sizeof(fb2_pair_hit))) != FB2_OK) break;
            if ((rc = keys.ensure((size_t)cap * 8 + 512)) != FB2_OK || (rc = tkeys.ensure((size_t)cap * 8 + 512)) != FB2_OK) break;
            if ((rc = vals.ensure((size_t)cap * 4 + 256)) != FB2_OK || (rc = tvals.ensure((size_t)cap * 4 + 256)) != FB2_OK) break;
            if ((rc = hist.ensure((size_t)radix_hist_words(cap) * 4)) != FB2_OK) break;
            t_alloc += now_ms() - t0; t0 = now_ms();
            cu(cudaMemsetAsync(counter.p, 0, sizeof(unsigned int), st));
            cu(cudaEventRecord(e0, st));
            if (inverted) {
                if (launch_dist_inverted_cut(d_h, d_l, (uint32_t)stride, (uint32_t)n_sk, (uint32_t)q, (uint32_t)(q + m), p_off.as<uint32_t>(),
                                             p_sk.as<uint32_t>(), p_run.as<unsigned long long>(), scaled, max_hash, hits.as<fb2_pair_hit>(),
                                             keys.as<unsigned long long>(), counter.as<unsigned int>(), cap, skip_self, jlow, st) != 0) {
                    rc = fb2_fail(FB2_ECUDA, "dist_inverted_kernel: could not reserve shared memory"); break;
                }
            } else
            if (launch_dist_tile_cut(d_h, d_l, (uint32_t)stride, (uint32_t)n_sk, (uint32_t)q, (uint32_t)(q + m), scaled, max_hash,
                                     hits.as<fb2_pair_hit>(), keys.as<unsigned long long>(), counter.as<unsigned int>(), cap, skip_self,
                                     jlow, st) != 0) { rc = fb2_fail(FB2_ECUDA, "dist_tile_kernel: could not reserve shared memory"); break; }
            cu(cudaEventRecord(e1, st));
            cu(cudaMemcpyAsync(h_cnt, counter.p, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
            cu(cudaStreamSynchronize(st));
            if (rc != FB2_OK) break;
            const unsigned int n = *h_cnt;
            t_wait += now_ms() - t0; t0 = now_ms();
            if (n > cap) {                       // more survivors than the buffer holds: fewer rows, then a larger buffer
                if (m > 9) { rows = inverted ? std::max<size_t>(1, m / 2) : std::max<size_t>(9, (m / 2 + 8) / 9 * 9); continue; }
                if (inverted && m > 1) { rows = std::max<size_t>(1, m / 2); continue; }
                if (cap >= (1u << 30)) { rc = fb2_fail(FB2_ENOMEM, "dist: too many surviving pairs for one query tile"); break; }
                cap <<= 1;
                continue;
            }
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) kms += ms;
            if (n) {
                launch_iota(vals.as<uint32_t>(), n, st);
                launch_radix_sort(keys.as<unsigned long long>(), vals.as<uint32_t>(), tkeys.as<unsigned long long>(), tvals.as<uint32_t>(),
                                  n, hist.as<uint32_t>(), st);
                launch_gather_hits(hits.as<fb2_pair_hit>(), vals.as<uint32_t>(), n, sorted.as<fb2_pair_hit>(), st);
                if ((size_t)n > pin_cap) {
                    if (h_pin) cudaFreeHost(h_pin);
                    h_pin = nullptr;
                    pin_cap = (size_t)n + n / 4 + 4096;
                    cu(cudaHostAlloc((void **)&h_pin, pin_cap * sizeof(fb2_pair_hit), cudaHostAllocDefault));
                }
                if (rc != FB2_OK) break;
                cu(cudaMemcpyAsync(h_pin, sorted.p, (size_t)n * sizeof(fb2_pair_hit), cudaMemcpyDeviceToHost, st));
                cu(cudaStreamSynchronize(st));
                if (rc != FB2_OK) break;
                t_sortcopy += now_ms() - t0; t0 = now_ms();
                out.append(h_pin, n);
                t_insert += now_ms() - t0;
            }
        } else {                                 // sketches longer than the tile kernel takes: dense slab, cut on the host
            const uint64_t np = (uint64_t)m * n_sk;
            if ((rc = dense.ensure(np * sizeof(fb2_pair_out))) != FB2_OK) break;
            h_dense.resize(np);
            cu(cudaEventRecord(e0, st));
            launch_dist_all(d_h, d_l, (uint32_t)stride, (uint32_t)n_sk, (uint32_t)q, np, scaled, max_hash, dense.as<fb2_pair_out>(), st);
            cu(cudaEventRecord(e1, st));
            cu(cudaMemcpyAsync(h_dense.data(), dense.p, np * sizeof(fb2_pair_out), cudaMemcpyDeviceToHost, st));
            cu(cudaStreamSynchronize(st));
            if (rc != FB2_OK) break;
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) kms += ms;
            for (size_t a = 0; a < m; ++a)
                for (size_t r = 0; r < n_sk; ++r) {
                    const fb2_pair_out &o = h_dense[a * n_sk + r];
                    if ((skip_self && q + a == r) || !host_passes_cut(o, jlow)) continue;
                    out.push(fb2_pair_hit{(uint32_t)(q + a), (uint32_t)r, o.common, o.i, o.j});
                }
        }
        q += m;
    }
    if (kernel_ms) *kernel_ms = kms;
    if (trace_rows)
        fprintf(stderr, "dist rows [%zu, %zu): buffers %.1f ms, kernel + wait %.1f ms, hit sort + D2H %.1f ms, append %.1f ms\n", qa, qb,
                t_alloc, t_wait, t_sortcopy, t_insert);
    if (h_pin) cudaFreeHost(h_pin);
    if (h_cnt) cudaFreeHost(h_cnt);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
    hits.release(); sorted.release(); keys.release(); tkeys.release(); vals.release(); tvals.release(); hist.release();
    counter.release(); dense.release();
    p_keys.release(); p_vals.release(); p_tkeys.release(); p_tvals.release(); p_hist.release(); p_off.release(); p_sk.release();
    p_run.release(); p_sum.release();
    return rc;
}

extern "C" int fb2_dist_all_pairs_cut(const uint64_t *hashes, const uint32_t *lens, size_t n_sk, size_t stride, double scale,
                                      size_t q0, size_t q1, uint8_t kmer_length, double max_distance, int skip_self,
                                      fb2_pair_hit *hits, size_t cap, uint64_t *n_hits, int32_t device, int ngpus) {
    if ((!hashes && n_sk && stride) || !lens || q0 > q1 || q1 > n_sk || !n_hits || (cap && !hits)) return fb2_fail(FB2_EINVAL, "bad argument");
    if (n_sk > 0xFFFFFFFFull || kmer_length == 0) return fb2_fail(FB2_EINVAL, "bad argument");
    for (size_t i = 0; i < n_sk; ++i) if (lens[i] > stride) return fb2_fail(FB2_EINVAL, "sketch length exceeds stride");
    *n_hits = 0;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fb2_fail(FB2_ECUDA, "no CUDA device: finch_b200 has no CPU fallback");
    DeviceScope dev_scope_(-1);   // the single-GPU path runs on this thread and selects its device
    std::vector<int> devs;
    if (ngpus == 1 || ndev == 1) {
        int d = device;
        if (d < 0 && cudaGetDevice(&d) != cudaSuccess) d = 0;
        if (d >= ndev) return fb2_fail(FB2_EINVAL, "device ordinal out of range");
        devs.push_back(d);
    } else for (int d = 0; d < (ngpus <= 0 ? ndev : std::min(ngpus, ndev)); ++d) devs.push_back(d);
    const size_t G = std::min<size_t>(devs.size(), std::max<size_t>(1, (q1 - q0 + 8) / 9));
    const bool trace = getenv("FB2_TRACE_DIST") != nullptr;
    auto now_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now_ms();
    std::vector<double> t_up(G, 0.0), t_rows(G, 0.0);
    const int scaled = scale > 0.0;
    const unsigned long long max_hash = scaled ? dist_max_hash(scale) : 0;
    const double jlow = dist_cut_jlow(max_distance, kmer_length);
    const size_t mat_bytes = std::max<size_t>(8, n_sk * stride * 8), len_bytes = std::max<size_t>(4, n_sk * 4);
    std::vector<HitSink> found(G);
    if (G == 1) { found[0].dst = hits; found[0].cap = cap; }
    std::vector<int> rcs(G, FB2_OK);
    std::vector<std::string> msgs(G);
    std::vector<double> kms(G, 0.0);
    std::vector<DevBuf> d_h(G), d_l(G);
    // the matrix goes host -> first GPU once; the others take it from there over peer copies (NVLink)
    std::mutex mu;
    std::condition_variable cv;
    int src_ready = 0;   // 0 not yet, 1 ok, -1 failed
    // row blocks in multiples of the 9-row query tile
    const size_t rows_total = q1 - q0;
    size_t per = ((rows_total + G - 1) / G + 8) / 9 * 9;
    auto work = [&](size_t g) {
        int r = FB2_OK;
        do {
            if (cudaSetDevice(devs[g]) != cudaSuccess) { r = fb2_fail(FB2_ECUDA, "cudaSetDevice failed"); break; }
            if ((r = d_h[g].ensure(mat_bytes)) != FB2_OK || (r = d_l[g].ensure(len_bytes)) != FB2_OK) break;
            if (g == 0) {
                if (src_ready != 1) {   // (already there when the index was tried first)
                    cudaError_t e = upload_large(d_h[0].p, hashes, n_sk * stride * 8);
                    if (e == cudaSuccess) e = cudaMemcpy(d_l[0].p, lens, n_sk * 4, cudaMemcpyHostToDevice);
                    if (e != cudaSuccess) r = fb2_fail(FB2_ECUDA, cudaGetErrorString(e));
                }
                { std::lock_guard<std::mutex> lk(mu); src_ready = r == FB2_OK ? 1 : -1; }
                cv.notify_all();
                if (r != FB2_OK) break;
            } else {
                { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return src_ready != 0; }); }
                if (src_ready < 0) { r = fb2_fail(FB2_ECUDA, "matrix upload failed on the first GPU"); break; }
                cudaError_t e = cudaMemcpyPeer(d_h[g].p, devs[g], d_h[0].p, devs[0], n_sk * stride * 8);
                if (e == cudaSuccess) e = cudaMemcpyPeer(d_l[g].p, devs[g], d_l[0].p, devs[0], n_sk * 4);
                if (e != cudaSuccess) r = fb2_fail(FB2_ECUDA, cudaGetErrorString(e));
                if (r != FB2_OK) break;
            }
            const size_t qa = std::min(q1, q0 + g * per), qb = std::min(q1, qa + per);
            t_up[g] = now_ms();
            r = dist_cut_rows(d_h[g].as<unsigned long long>(), d_l[g].as<uint32_t>(), lens, n_sk, stride, scaled, max_hash, qa, qb,
                              skip_self, jlow, found[g], &kms[g]);
            t_rows[g] = now_ms();
        } while (0);
        if (g == 0 && src_ready == 0) { { std::lock_guard<std::mutex> lk(mu); src_ready = -1; } cv.notify_all(); }
        rcs[g] = r;
        if (r != FB2_OK) msgs[g] = fb2_last_error();
    };
    // Several GPUs: the inverted index (when it applies) does ALL rows on the first GPU in less time than it takes to
    // hand the matrix to the others (C5: 0.2 s on one GPU; every GPU would have to sort all postings anyway).  Only when
    // the index declines (dense collections, no positive bound, empty sketches) are the rows cut over the GPUs.
    bool done_on_first = false;
    if (G > 1 && jlow > 0.0 && !(getenv("FB2_DIST_INVERTED") && atoi(getenv("FB2_DIST_INVERTED")) == 0)) {
        int r = FB2_OK;
        do {
            if (cudaSetDevice(devs[0]) != cudaSuccess) { r = fb2_fail(FB2_ECUDA, "cudaSetDevice failed"); break; }
            if ((r = d_h[0].ensure(mat_bytes)) != FB2_OK || (r = d_l[0].ensure(len_bytes)) != FB2_OK) break;
            cudaError_t e = upload_large(d_h[0].p, hashes, n_sk * stride * 8);
            if (e == cudaSuccess) e = cudaMemcpy(d_l[0].p, lens, n_sk * 4, cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { r = fb2_fail(FB2_ECUDA, cudaGetErrorString(e)); break; }
            src_ready = 1;
            int only = 1;
            HitSink direct;
            direct.dst = hits; direct.cap = cap;
            t_up[0] = now_ms();
            r = dist_cut_rows(d_h[0].as<unsigned long long>(), d_l[0].as<uint32_t>(), lens, n_sk, stride, scaled, max_hash, q0, q1,
                              skip_self, jlow, direct, &kms[0], &only);
            t_rows[0] = now_ms();
            if (r == FB2_OK && only) { found[0] = std::move(direct); done_on_first = true; }
        } while (0);
        if (r != FB2_OK) {
            const std::string msg = fb2_last_error();
            DeviceScope restore(-1);
            cudaSetDevice(devs[0]); d_h[0].release(); d_l[0].release();
            return fb2_fail(r, msg);
        }
    }
    if (done_on_first) { /* nothing left */ }
    else if (G == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (size_t g = 0; g < G; ++g) th.emplace_back(work, g);
        for (auto &t : th) t.join();
    }
    {
        DeviceScope restore(-1);
        for (size_t g = 0; g < G; ++g)   // (only devices that hold something: selecting an untouched GPU creates its context, ~1 s)
            if (d_h[g].p || d_l[g].p) { cudaSetDevice(devs[g]); d_h[g].release(); d_l[g].release(); }
    }
    for (size_t g = 0; g < G; ++g) if (rcs[g] != FB2_OK) return fb2_fail(rcs[g], msgs[g]);
    uint64_t total = 0;
    double kmax = 0.0;
    for (size_t g = 0; g < G; ++g) {
        if (!found[g].dst) {
            const size_t room = total < cap ? cap - (size_t)total : 0, n = std::min(room, found[g].vec.size());
            if (n) memcpy(hits + total, found[g].vec.data(), n * sizeof(fb2_pair_hit));
        }
        total += found[g].n;
        kmax = std::max(kmax, kms[g]);
    }
    g_dist_kernel_ms = kmax;     // the slowest GPU's summed kernel time
    *n_hits = total;
    if (trace)
        fprintf(stderr, "dist: matrix on GPU 0 after %.1f ms, its rows done after %.1f ms (kernels %.1f ms), everything after %.1f ms\n",
                t_up[0] - t_begin, t_rows[0] - t_begin, kms[0], now_ms() - t_begin);
    if (total > cap) return fb2_fail(FB2_ENOMEM, "dist: " + std::to_string(total) + " surviving pairs, room for " + std::to_string(cap));
    return FB2_OK;
}

extern "C" int fb2_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}
extern "C" const char *fb2_version(void) { return "finch_b200 0.1.0 (sm_100a)"; }

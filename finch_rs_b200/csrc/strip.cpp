// strip.cpp -- optional host pre-strip of FASTQ input (FB2_HOST_STRIP): record FRAMING on the host CPU, so that
// only what the GPU needs crosses PCIe.
//
// The end-to-end rate of fb2_sketcher_feed_fastx from host memory is the PCIe link's: 2.09 raw bytes
// travel per base of a 150 bp FASTQ, of which 52 % are header, '+' and quality lines that the parse kernels throw
// away on arrival.  With the strip on, worker threads frame the records here -- the part of needletail's FASTQ reader
// that is a newline scan (call site lib/src/lib.rs:60-68) -- and ship the sequence lines alone, one '\n' after each,
// which the GPU then takes through the same MODE_LINES path SketchScheme::process uses.  Everything downstream
// (normalisation, canonical k-mers, hashing, bottom-s) stays on the GPU; what this file decides is exactly what the
// parse kernels decide for a FASTQ stream: where the sequence lines are, sequence().len() of every record (line
// length minus one trailing CR), and whether the stream is well formed ('@' / '+' line starts, equal sequence and
// quality lengths, no truncated last record; trailing blank lines tolerated).
//
// Threads own whole records.  A range boundary is a GUESS (first '@' line after the cut whose third line starts with
// '+'), verified for free: the previous thread walks record by record from its own verified start and must arrive
// exactly at the guessed position; where it does not, the caller re-strips from the position it did arrive at.
#include <cstdint>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "strip.h"

namespace fb2 {

namespace {

// position of the next '\n' in [p, end), or end
#if defined(__x86_64__)
__attribute__((target("avx2"))) const uint8_t *find_nl_avx2(const uint8_t *p, const uint8_t *end) {
    const __m256i nl = _mm256_set1_epi8('\n');
    while (p + 32 <= end) {
        const uint32_t m = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_loadu_si256((const __m256i *)p), nl));
        if (m) return p + __builtin_ctz(m);
        p += 32;
    }
    const void *r = p < end ? memchr(p, '\n', (size_t)(end - p)) : nullptr;
    return r ? (const uint8_t *)r : end;
}
#endif
const uint8_t *find_nl_plain(const uint8_t *p, const uint8_t *end) {
    const void *r = p < end ? memchr(p, '\n', (size_t)(end - p)) : nullptr;
    return r ? (const uint8_t *)r : end;
}
typedef const uint8_t *(*find_nl_fn)(const uint8_t *, const uint8_t *);
find_nl_fn pick_find_nl() {
#if defined(__x86_64__)
    if (__builtin_cpu_supports("avx2")) return find_nl_avx2;
#endif
    return find_nl_plain;
}

}  // namespace

// Strip whole records from [p, end), p at a record start.  Stops at the first record that is not complete inside the
// range (its four newlines are not all there) or, with `stop_at` set, at the first record start >= stop_at.
// Appends "sequence line\n" per record to out (capacity: (end - p) / 2 + 2 is always enough: a quality line is as
// long as its sequence line).  Returns the position where it stopped (a record start, or where the incomplete
// record begins).
static inline bool all_blank(const uint8_t *a, const uint8_t *b) {
    for (; a < b; ++a) if (*a != '\n' && *a != '\r') return false;
    return true;
}
// One record whose four newlines are known: the reader's checks, then ship "sequence line\n".  Returns false when the
// output buffer is full (nothing consumed).
static inline bool take_record(const uint8_t *p, const uint8_t *l1, const uint8_t *l2, const uint8_t *l3, const uint8_t *l4,
                               uint64_t base_off, StripOut &o, uint8_t *obase, size_t ocap, size_t &olen) {
    const uint8_t *seq = l1 + 1, *sep = l2 + 1, *qual = l3 + 1;
    size_t sl = (size_t)(l2 - seq), ql = (size_t)(l4 - qual);
    if (sl && seq[sl - 1] == '\r') --sl;
    if (ql && qual[ql - 1] == '\r') --ql;
    if (olen + sl + 1 > ocap) return false;
    const uint64_t off = base_off + (uint64_t)(p - o.origin);
    if (*p != '@' && all_blank(p, l4)) {
        // four lines of terminators only: legal as trailing blank lines, an error when content follows
        if (off < o.first_blank) o.first_blank = off;
        return true;
    }
    if (off + 1 > o.last_nonblank) o.last_nonblank = off + 1;
    if (*p != '@') { if (off < o.bad_pos) o.bad_pos = off; }
    if (*sep != '+') { const uint64_t so = base_off + (uint64_t)(sep - o.origin); if (so < o.bad_pos) o.bad_pos = so; }
    if (sl != ql) { const uint64_t ho = base_off + (uint64_t)(l1 - o.origin); if (ho < o.len_bad_pos) o.len_bad_pos = ho; }
    memcpy(obase + olen, seq, sl);
    obase[olen + sl] = '\n';
    olen += sl + 1;
    o.bases += sl;
    o.records += 1;
    return true;
}

#if defined(__x86_64__)
// Output of the AVX2 path goes through a small cache-resident buffer and leaves it with non-temporal stores: the
// stripped lines are written once and next read by the DMA engine, so they should neither be read for ownership
// nor displace the input stream from the caches.
struct WcOut {
    static constexpr size_t CAP = 8192;
    alignas(64) uint8_t buf[CAP + 64];
    size_t fill = 0;
    uint8_t *dst;                 // next 32-byte aligned position of the real output
    size_t head = 0;              // bytes in front of dst that were written directly (unaligned start)
};
__attribute__((target("avx2")))
static inline void wc_flush(WcOut &w, bool all) {
    size_t n = all ? w.fill : (w.fill & ~(size_t)31);
    size_t i = 0;
    if (((uintptr_t)w.dst & 31u) == 0) {
        for (; i + 32 <= n; i += 32) _mm256_stream_si256((__m256i *)(w.dst + i), _mm256_load_si256((const __m256i *)(w.buf + i)));
    }
    if (i < n) { memcpy(w.dst + i, w.buf + i, n - i); i = n; }
    w.dst += n;
    const size_t rest = w.fill - n;
    if (rest) memmove(w.buf, w.buf + n, rest);
    w.fill = rest;
}
// Single pass: 64 bytes per step -> a 64-bit newline mask; every set bit is a line end, every fourth one a record.
__attribute__((target("avx2,bmi,bmi2,lzcnt,popcnt")))
static const uint8_t *strip_records_avx2(const uint8_t *p, const uint8_t *end, const uint8_t *stop_at, uint64_t base_off, StripOut &o) {
    uint8_t *const obase = o.spilled ? o.spill.data() : o.out;
    const size_t ocap = o.spilled ? o.spill.size() : o.out_cap;
    size_t olen = o.out_len;
    WcOut w;
    w.dst = obase + olen;
    const __m256i nlv = _mm256_set1_epi8('\n');
    const uint8_t *rec = p;                 // start of the record being framed
    const uint8_t *nl[3];                   // its newlines so far
    int have = 0;
    const uint8_t *q = p;                   // scan position
    bool done = stop_at && rec >= stop_at;
    while (!done && q + 64 <= end) {
        const uint32_t m0 = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_loadu_si256((const __m256i *)q), nlv));
        const uint32_t m1 = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_loadu_si256((const __m256i *)(q + 32)), nlv));
        uint64_t m = (uint64_t)m0 | ((uint64_t)m1 << 32);
        while (m) {
            const uint8_t *e = q + __builtin_ctzll(m);
            m &= m - 1;
            if (have < 3) { nl[have++] = e; continue; }
            // ---- a complete record: rec .. e ----
            const uint8_t *seq = nl[0] + 1, *sep = nl[1] + 1, *qual = nl[2] + 1;
            size_t sl = (size_t)(nl[1] - seq), ql = (size_t)(e - qual);
            if (sl && seq[sl - 1] == '\r') --sl;
            if (ql && qual[ql - 1] == '\r') --ql;
            if (olen + sl + 1 > ocap) { done = true; break; }
            const uint64_t off = base_off + (uint64_t)(rec - o.origin);
            if (*rec != '@' && all_blank(rec, e)) {
                if (off < o.first_blank) o.first_blank = off;
            } else {
                if (off + 1 > o.last_nonblank) o.last_nonblank = off + 1;
                if (*rec != '@') { if (off < o.bad_pos) o.bad_pos = off; }
                if (*sep != '+') { const uint64_t so = base_off + (uint64_t)(sep - o.origin); if (so < o.bad_pos) o.bad_pos = so; }
                if (sl != ql) { const uint64_t ho = base_off + (uint64_t)(nl[0] - o.origin); if (ho < o.len_bad_pos) o.len_bad_pos = ho; }
                if (sl + 1 > WcOut::CAP) {              // a very long line: around the staging buffer
                    wc_flush(w, true);
                    memcpy(w.dst, seq, sl);
                    w.dst[sl] = '\n';
                    w.dst += sl + 1;
                } else {
                    if (w.fill + sl + 1 > WcOut::CAP) wc_flush(w, false);
                    memcpy(w.buf + w.fill, seq, sl);
                    w.buf[w.fill + sl] = '\n';
                    w.fill += sl + 1;
                }
                olen += sl + 1;
                o.bases += sl;
                o.records += 1;
            }
            rec = e + 1;
            have = 0;
            if (stop_at && rec >= stop_at) { done = true; break; }
        }
        q += 64;
    }
    wc_flush(w, true);
    _mm_sfence();
    o.out_len = olen;
    return rec;     // the caller's generic loop takes the last < 64 bytes (from the start of the open record)
}
#endif

static const uint8_t *strip_records_generic(const uint8_t *p, const uint8_t *end, const uint8_t *stop_at, uint64_t base_off, StripOut &o);
const uint8_t *strip_records(const uint8_t *p, const uint8_t *end, const uint8_t *stop_at, uint64_t base_off, StripOut &o) {
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2");
    if (avx2 && end - p >= 4096) p = strip_records_avx2(p, end, stop_at, base_off, o);
#endif
    return strip_records_generic(p, end, stop_at, base_off, o);
}
static const uint8_t *strip_records_generic(const uint8_t *p, const uint8_t *end, const uint8_t *stop_at, uint64_t base_off, StripOut &o) {
    static const find_nl_fn find_nl = pick_find_nl();
    uint8_t *const obase = o.spilled ? o.spill.data() : o.out;
    const size_t ocap = o.spilled ? o.spill.size() : o.out_cap;
    size_t olen = o.out_len;
    while (p < end && (!stop_at || p < stop_at)) {
        const uint8_t *l1 = find_nl(p, end);
        if (l1 == end) break;
        const uint8_t *seq = l1 + 1;
        const uint8_t *l2 = find_nl(seq, end);
        if (l2 == end) break;
        const uint8_t *sep = l2 + 1;
        const uint8_t *l3 = find_nl(sep, end);
        if (l3 == end) break;
        const uint8_t *qual = l3 + 1;
        const uint8_t *l4 = find_nl(qual, end);
        if (l4 == end) break;
        size_t sl = (size_t)(l2 - seq), ql = (size_t)(l4 - qual);
        if (sl && seq[sl - 1] == '\r') --sl;
        if (ql && qual[ql - 1] == '\r') --ql;
        if (olen + sl + 1 > ocap) break;                         // out of room: the caller redoes from here
        const uint64_t off = base_off + (uint64_t)(p - o.origin);
        if (*p != '@' && all_blank(p, l4)) {
            // four lines of terminators only: legal as trailing blank lines, an error when content follows
            if (off < o.first_blank) o.first_blank = off;
            p = l4 + 1;
            continue;
        } else {
            // a record: the reader's checks
            if (off + 1 > o.last_nonblank) o.last_nonblank = off + 1;
            if (*p != '@') { if (off < o.bad_pos) o.bad_pos = off; }
            if (*sep != '+') { const uint64_t so = base_off + (uint64_t)(sep - o.origin); if (so < o.bad_pos) o.bad_pos = so; }
            if (sl != ql) { const uint64_t ho = base_off + (uint64_t)(l1 - o.origin); if (ho < o.len_bad_pos) o.len_bad_pos = ho; }
        }
        // ship the sequence line as the reader's sequence() sees it (CR trimmed) + the record separator
        memcpy(obase + olen, seq, sl);
        obase[olen + sl] = '\n';
        olen += sl + 1;
        o.bases += sl;
        o.records += 1;
        p = l4 + 1;
    }
    o.out_len = olen;
    return p;
}

int strip_final(const uint8_t *p, const uint8_t *end, uint64_t base_off, StripOut &o) {
    static const find_nl_fn find_nl = pick_find_nl();
    uint8_t *const obase = o.spilled ? o.spill.data() : o.out;
    const size_t ocap = o.spilled ? o.spill.size() : o.out_cap;
    while (p < end) {
        if (all_blank(p, end)) break;                            // trailing blank lines
        const uint64_t off = base_off + (uint64_t)(p - o.origin);
        if (off + 1 > o.last_nonblank) o.last_nonblank = off + 1;
        if (*p != '@') return 1;
        const uint8_t *l1 = find_nl(p, end);
        if (l1 == end) return 1;
        const uint8_t *seq = l1 + 1;
        const uint8_t *l2 = find_nl(seq, end);
        if (l2 == end) return 1;
        const uint8_t *sep = l2 + 1;
        if (sep >= end || *sep != '+') return 1;
        const uint8_t *l3 = find_nl(sep, end);
        if (l3 == end) return 1;
        const uint8_t *qual = l3 + 1;
        const uint8_t *l4 = find_nl(qual, end);                  // may be `end`: the last line needs no newline
        size_t sl = (size_t)(l2 - seq), ql = (size_t)(l4 - qual);
        if (sl && seq[sl - 1] == '\r') --sl;
        if (ql && qual[ql - 1] == '\r') --ql;
        if (sl != ql) return 1;
        if (o.out_len + sl + 1 > ocap) return 1;                 // cannot happen: the caller sizes for (end - p) / 2 + 2
        memcpy(obase + o.out_len, seq, sl);
        obase[o.out_len + sl] = '\n';
        o.out_len += sl + 1;
        o.bases += sl;
        o.records += 1;
        p = l4 < end ? l4 + 1 : end;
    }
    return 0;
}

const uint8_t *complete_record(const std::vector<uint8_t> &carry, const uint8_t *p, const uint8_t *end) {
    static const find_nl_fn find_nl = pick_find_nl();
    size_t have = 0;
    for (uint8_t c : carry) have += c == '\n';
    // the carry never holds a complete record: have <= 3
    const uint8_t *q = p;
    for (size_t need = 4 - (have & 3); need > 0; --need) {
        const uint8_t *nl = find_nl(q, end);
        if (nl == end) return nullptr;
        q = nl + 1;
    }
    return q;
}

// First position q >= from that looks like a record start: follows a '\n', holds '@', its next line does not start
// with '@' and the line after that starts with '+'.  Returns end when there is none.
const uint8_t *guess_record_start(const uint8_t *begin, const uint8_t *from, const uint8_t *end) {
    static const find_nl_fn find_nl = pick_find_nl();
    const uint8_t *p = from;
    if (p == begin) return p;
    --p;                                     // so that a record starting exactly at `from` is found
    while (p < end) {
        const uint8_t *nl = find_nl(p, end);
        if (nl == end) return end;
        const uint8_t *q = nl + 1;
        if (q >= end) return end;
        if (*q == '@') {
            const uint8_t *l1 = find_nl(q, end);
            if (l1 == end) return end;
            const uint8_t *s1 = l1 + 1;
            const uint8_t *l2 = s1 < end ? find_nl(s1, end) : end;
            if (l2 == end) return end;
            const uint8_t *s2 = l2 + 1;
            if (s2 < end && *s2 == '+' && *s1 != '@') return q;
        }
        p = q;
    }
    return end;
}

// Worker threads of strip_parallel, kept between calls: a block of 256 MiB is framed in 4 ms, and starting sixteen
// threads for each of them was a tenth of that.  run(n, f) executes f(0) .. f(n - 1), f(0) on the caller; one call at
// a time (concurrent callers queue on the mutex).
namespace {
class StripPool {
public:
    void run(unsigned n, const std::function<void(unsigned)> &f) {
        std::lock_guard<std::mutex> call(call_mu_);
        {
            std::unique_lock<std::mutex> lk(mu_);
            while (workers_.size() + 1 < n) {
                const unsigned id = (unsigned)workers_.size() + 1;
                workers_.emplace_back([this, id] { loop(id); });
                workers_.back().detach();
            }
            fn_ = &f; n_ = n; pending_ = n - 1; ++epoch_;
        }
        cv_.notify_all();
        f(0);
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }
private:
    void loop(unsigned id) {
        unsigned long long seen = 0;
        while (true) {
            const std::function<void(unsigned)> *f = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return epoch_ != seen; });
                seen = epoch_;
                if (id < n_) f = fn_;
            }
            if (f) {
                (*f)(id);
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    std::mutex call_mu_, mu_;
    std::condition_variable cv_, done_;
    std::vector<std::thread> workers_;
    const std::function<void(unsigned)> *fn_ = nullptr;
    unsigned n_ = 0, pending_ = 0;
    unsigned long long epoch_ = 0;
};
StripPool &strip_pool() { static StripPool *p = new StripPool(); return *p; }   // (leaked on purpose: detached workers outlive statics)
}  // namespace

// Strip [p, end) (p at a record start) with `threads` threads into outs[0..threads): ranges in stream order.
// Returns where the complete records end (start of the incomplete tail, or end).
const uint8_t *strip_parallel(const uint8_t *p, const uint8_t *end, uint64_t base_off, unsigned threads,
                              std::vector<StripOut> &outs) {
    const size_t n = (size_t)(end - p);
    if (threads < 1) threads = 1;
    if (n < (size_t)threads * (256u << 10)) threads = 1;
    std::vector<const uint8_t *> start(threads + 1, end), stopped(threads, end);
    start[0] = p;
    for (unsigned t = 1; t < threads; ++t) {
        const uint8_t *g = guess_record_start(p, p + n / threads * t, end);
        start[t] = g < start[t - 1] ? start[t - 1] : g;
    }
    start[threads] = end;
    auto run = [&](unsigned t) {
        StripOut &o = outs[t];
        o.origin = p;
        if (start[t] >= end) { stopped[t] = end; return; }
        // run to the next thread's guessed start (or, for the last thread, as far as complete records go)
        stopped[t] = strip_records(start[t], end, t + 1 < threads ? start[t + 1] : nullptr, base_off, o);
    };
    if (threads == 1) run(0);
    else strip_pool().run(threads, run);
    // verification: every thread must have arrived exactly at the next thread's start
    for (unsigned t = 0; t + 1 < threads; ++t) {
        if (start[t + 1] >= end && stopped[t] >= start[t + 1]) continue;
        if (stopped[t] != start[t + 1]) {
            // the guess was not a record start (or a record ran past it): everything after thread t is redone
            // sequentially from where thread t really stopped, into thread t's successor slots
            for (unsigned u = t + 1; u < threads; ++u) outs[u].reset_counts();
            StripOut &o = outs[t + 1];
            o.origin = p;
            // the successor buffers are sized for their own ranges: a growable pageable buffer takes the rest
            o.spill.resize((size_t)(end - stopped[t]) / 2 + 16);
            o.spilled = true;
            return strip_records(stopped[t], end, nullptr, base_off, o);
        }
    }
    return stopped[threads - 1];
}

}  // namespace fb2

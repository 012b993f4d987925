// parse.cu -- K1/K2: raw FASTA/FASTQ bytes in HBM -> symbol regions in HBM.
//
// Replaces, in bulk, needletail's record loop (lib/src/lib.rs:60-68) and the per-record
// `seq.normalize(false)` of SketchScheme::process (lib/src/sketch_schemes/mash.rs:72-73):
// every base that `process` would see becomes one byte 0..3 (A,C,G,T); every other kept
// position (N, IUPAC, gaps) and every record boundary becomes SYM_BREAK (4); whitespace and
// non-sequence lines vanish.  k-mers of the symbol stream == canonical_kmers of the records.
//
// A chunk is cut into supertiles of st_tiles x 4 KiB.  Per chunk:
//   phase_kernel      : per supertile, how the line state changes across it (FASTQ: newline
//                       count mod 4; FASTA: is the last line a header) -- newline masks only
//   phase_scan_kernel : scan of those state maps -> start state of every supertile
//   pack_kernel       : one block per supertile walks its tiles with the known state and writes
//                       the symbols into the supertile's own region (count in region_count[]);
//                       no global compaction is needed because the hash kernel walks regions
//   front_fix_kernel  : copies the last 32 symbols before each region into its front pad (and
//                       the chunk's tail into the carry buffer) so k-mers span region / chunk seams
// Byte classification is SIMD-in-register on the common case (ACGTacgt + '\n'), with an exact
// per-byte LUT fallback whenever a thread meets anything else among its sequence bytes.
#include "common.cuh"
#include "device_types.cuh"

namespace fb2 {

// 0x80 in every byte of x that is zero (exact, no cross-byte borrows)
__device__ __forceinline__ uint32_t zero_bytes80(uint32_t x) {
    return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);
}
// flags at bits 7,15,23,31 -> 4 contiguous bits
__device__ __forceinline__ uint32_t gather4(uint32_t f80) { return ((f80 >> 7) * 0x01020408u) >> 24; }
// 4 bits -> 0xFF in each flagged byte
__device__ __forceinline__ uint32_t spread4(uint32_t nib) { return ((nib * 0x00204081u) & 0x01010101u) * 0xFFu; }

struct Raw16 {
    uint32_t w[4];
    uint32_t valid;   // bytes that exist (off + i < len)
    uint32_t nl;      // '\n'
};
__device__ __forceinline__ uint32_t raw_byte(const Raw16 &r, int i) {
    const uint64_t lo = (uint64_t)r.w[0] | ((uint64_t)r.w[1] << 32), hi = (uint64_t)r.w[2] | ((uint64_t)r.w[3] << 32);
    return (uint32_t)((i < 8 ? (lo >> (8 * i)) : (hi >> (8 * (i - 8)))) & 0xFFu);
}
// issue the 16-byte load of a full piece early (software pipelining of the tile loop)
__device__ __forceinline__ uint4 fetch_raw(const uint8_t *__restrict__ raw, uint32_t off, uint32_t len) {
    if (off + 16u <= len) return __ldg(reinterpret_cast<const uint4 *>(raw + off));
    return make_uint4(0, 0, 0, 0);
}
__device__ __forceinline__ void load_raw(const uint8_t *__restrict__ raw, uint32_t off, uint32_t len, Raw16 &r,
                                         const uint4 *pre = nullptr) {
    r.w[0] = r.w[1] = r.w[2] = r.w[3] = 0;
    int nvalid;
    if (off + 16u <= len) {
        const uint4 v = pre ? *pre : __ldg(reinterpret_cast<const uint4 *>(raw + off));
        r.w[0] = v.x; r.w[1] = v.y; r.w[2] = v.z; r.w[3] = v.w;
        nvalid = 16;
    } else {
        nvalid = off < len ? (int)(len - off) : 0;
#pragma unroll
        for (int i = 0; i < 16; ++i)   // static indices keep w[] in registers
            if (i < nvalid) r.w[i >> 2] |= (uint32_t)raw[off + i] << (8 * (i & 3));
    }
    r.valid = nvalid >= 16 ? 0xFFFFu : ((1u << nvalid) - 1u);
    uint32_t nl = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) nl |= gather4(zero_bytes80(r.w[j] ^ 0x0A0A0A0Au)) << (4 * j);
    r.nl = nl & r.valid;
}

// Classification of the sequence bytes `sm` of a thread (bytes outside sm are ignored).
struct SeqCls {
    uint32_t wsm;    // ' ', '\t', '\r' among sm          (removed by normalize)
    uint32_t cr;     // '\r' among the 16 bytes            (0 on the fast path: none among sm)
    uint32_t bad;    // kept but not a base, among sm      (N, IUPAC, '>', digits, ...)
    uint32_t cw[4];  // per byte: 2-bit base code in the low bits
};
__device__ __forceinline__ void classify_seq(const Raw16 &r, uint32_t sm, const uint8_t *lut, SeqCls &c) {
    // fast path: code from bits 1,2 of the upper-cased letter; re-expand and compare, which is
    // equal exactly for ACGTacgt.  Any mismatch inside sm sends the thread to the exact LUT.
    uint32_t diff = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t u = r.w[j] & 0xDFDFDFDFu;
        const uint32_t c2 = ((u >> 1) ^ (u >> 2)) & 0x03030303u;
        c.cw[j] = c2;
        uint32_t z = (c2 | (c2 >> 4)) & 0x00FF00FFu;
        z = (z | (z >> 8)) & 0xFFFFu;
        const uint32_t e = __byte_perm(0x54474341u, 0u, z);
        diff |= (e ^ u) & spread4((sm >> (4 * j)) & 0xFu);
    }
    c.wsm = 0; c.cr = 0; c.bad = 0;
    if (diff != 0u) {  // exact per-byte classes (needletail normalize(false))
        uint32_t wsm = 0, cr = 0, bad = 0;
        c.cw[0] = c.cw[1] = c.cw[2] = c.cw[3] = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint32_t b = (r.w[i >> 2] >> (8 * (i & 3))) & 0xFFu;
            const uint32_t k = lut[b];
            cr |= (k == CLS_CR ? 1u : 0u) << i;
            wsm |= ((k == CLS_WS || k == CLS_CR) ? 1u : 0u) << i;
            bad |= ((k >= 4u && k != CLS_WS && k != CLS_CR && k != CLS_NL) ? 1u : 0u) << i;
            c.cw[i >> 2] |= (k & 3u) << (8 * (i & 3));
        }
        c.wsm = wsm & sm; c.cr = cr & r.valid; c.bad = bad & sm;
    }
}
// Final symbol code words: base code, or SYM_BREAK (4) where `four` is set.
__device__ __forceinline__ void code_words(const SeqCls &c, uint32_t four, uint32_t out[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t m1 = (((four >> (4 * j)) & 0xFu) * 0x00204081u) & 0x01010101u;
        out[j] = (c.cw[j] & ~(m1 * 3u)) | (m1 << 2);
    }
}

// ---- block-wide helpers (256 threads) ---------------------------------------------------------
__device__ __forceinline__ uint32_t block_exscan_add(uint32_t v, uint32_t *sh8, uint32_t &total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += n;
    }
    __syncthreads();  // protect sh8 from a previous use
    if (lane == 31) sh8[wid] = x;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < TILE_THREADS / 32; ++i) {
        const uint32_t s = sh8[i];
        if (i < wid) base += s;
        tot += s;
    }
    total = tot;
    return base + x - v;
}
// Same for four 16-bit counters packed into 64 bits (fields must not overflow: <= 4096 each here).
__device__ __forceinline__ unsigned long long block_exscan_add64(unsigned long long v, unsigned long long *sh8l,
                                                                 unsigned long long &total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned long long x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long n = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += n;
    }
    __syncthreads();  // protect sh8l from a previous use
    if (lane == 31) sh8l[wid] = x;
    __syncthreads();
    unsigned long long base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < TILE_THREADS / 32; ++i) {
        const unsigned long long s = sh8l[i];
        if (i < wid) base += s;
        tot += s;
    }
    total = tot;
    return base + x - v;
}
__device__ __forceinline__ uint32_t block_exscan_max(uint32_t v, uint32_t *sh8, uint32_t &total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x = max(x, n);
    }
    uint32_t ex = __shfl_up_sync(0xffffffffu, x, 1);
    if (lane == 0) ex = 0;
    __syncthreads();
    if (lane == 31) sh8[wid] = x;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < TILE_THREADS / 32; ++i) {
        const uint32_t s = sh8[i];
        if (i < wid) base = max(base, s);
        tot = max(tot, s);
    }
    total = tot;
    return max(base, ex);
}
struct OpAdd { __device__ unsigned long long operator()(unsigned long long a, unsigned long long b) const { return a + b; } };
struct OpMin { __device__ unsigned long long operator()(unsigned long long a, unsigned long long b) const { return a < b ? a : b; } };
struct OpMax { __device__ unsigned long long operator()(unsigned long long a, unsigned long long b) const { return a > b ? a : b; } };
template <class Op>
__device__ __forceinline__ unsigned long long block_reduce64(unsigned long long v, unsigned long long *sh8, Op op,
                                                             unsigned long long ident) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = op(v, __shfl_down_sync(0xffffffffu, v, d));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh8[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long t = ident;
#pragma unroll
    for (int i = 0; i < TILE_THREADS / 32; ++i) t = op(t, sh8[i]);
    return t;
}

// previous raw byte of each thread's first byte: neighbour lane's last byte, or memory for lane 0
__device__ __forceinline__ uint32_t prev_byte(const Raw16 &r, const uint8_t *__restrict__ raw, uint32_t off, uint32_t carried) {
    uint32_t p = __shfl_up_sync(0xffffffffu, r.w[3] >> 24, 1);
    if ((threadIdx.x & 31) == 0) p = off ? raw[off - 1] : carried;
    return p;
}

// FASTQ: mask of this thread's bytes that lie in sequence lines (their '\n' included), given
// the phase of the line containing byte 0.
__device__ __forceinline__ uint32_t fastq_seq_mask(uint32_t nl, uint32_t phase0) {
    uint32_t qm = 0, mm = nl, lo = 0, ph = phase0 & 3u;
    while (true) {
        const int nb = mm ? (__ffs(mm) - 1) : 16;
        const uint32_t upto = nb >= 15 ? 0xFFFFu : ((2u << nb) - 1u);
        const uint32_t seg = upto & ~((1u << lo) - 1u);
        if (ph == 1u) qm |= seg;
        if (nb >= 15) break;
        mm &= mm - 1u;
        lo = (uint32_t)nb + 1u;
        ph = (ph + 1u) & 3u;
    }
    return qm;
}
// FASTA: which line starts begin with '>' (line starts are few per thread: point look-ups)
__device__ __forceinline__ uint32_t header_starts(const Raw16 &r, uint32_t ls) {
    uint32_t hs = 0, l = ls;
    while (l) {
        const int i = __ffs(l) - 1;
        l &= l - 1u;
        if (raw_byte(r, i) == '>') hs |= 1u << i;
    }
    return hs;
}
// FASTA: header-byte mask (header lines including their '\n').  Carry-propagation flood: a seed at
// the bottom of a run of non-newline bytes floods the run and the newline that ends it.
__device__ __forceinline__ uint32_t fasta_header_mask(uint32_t hs, uint32_t nl, bool seed0) {
    const uint32_t g = hs | (seed0 ? 1u : 0u);
    const uint32_t p = ~nl & 0xFFFFu;
    return ((g + p) ^ p) & 0x1FFFFu;
}

// ------------------------------------------------------------------------------------------------
// Per supertile: state map as 2 bits per start state.  FASTQ: phase += newlines.  FASTA: state of
// the line containing the supertile's last byte (0 sequence, 1 header), or unchanged if the
// supertile holds no line start.
template <int MODE>
__global__ void __launch_bounds__(TILE_THREADS)
phase_kernel(const uint8_t *__restrict__ raw, ChunkGeom g, const ParseCarry *__restrict__ carry,
             uint32_t *__restrict__ st_map) {
    __shared__ unsigned long long sh8l[8];
    const int tid = threadIdx.x;
    const uint32_t st = blockIdx.x;
    const uint32_t t0 = st * g.st_tiles, t1 = min(t0 + g.st_tiles, g.n_tiles);
    unsigned long long acc = 0;  // FASTQ: newline count; FASTA: max over line starts of (pos+1) << 1 | is_header
    uint4 nextv = fetch_raw(raw, t0 * (uint32_t)TILE_BYTES + (uint32_t)tid * 16u, g.len);
    for (uint32_t t = t0; t < t1; ++t) {
        const uint32_t off = t * (uint32_t)TILE_BYTES + (uint32_t)tid * 16u;
        Raw16 r;
        const uint4 curv = nextv;
        if (t + 1 < t1) nextv = fetch_raw(raw, off + (uint32_t)TILE_BYTES, g.len);
        uint32_t nl_count = 0;
        if (MODE == MODE_FASTQ && off + 16u <= g.len) {
            // a full piece only needs its newline COUNT here: one flag per byte, no gather into a position mask
            nl_count = __popc(zero_bytes80(curv.x ^ 0x0A0A0A0Au)) + __popc(zero_bytes80(curv.y ^ 0x0A0A0A0Au)) +
                       __popc(zero_bytes80(curv.z ^ 0x0A0A0A0Au)) + __popc(zero_bytes80(curv.w ^ 0x0A0A0A0Au));
        } else {
            load_raw(raw, off, g.len, r, &curv);
            nl_count = __popc(r.nl);
        }
        if (MODE == MODE_FASTQ) {
            acc += nl_count;
        } else {
            const uint32_t p1 = prev_byte(r, raw, off, carry->prev1);
            const uint32_t ls = ((r.nl << 1) | (p1 == '\n' ? 1u : 0u)) & r.valid;
            if (ls) {
                const int hi = 31 - __clz(ls);
                acc = ((unsigned long long)(off + hi + 1u) << 1) | (raw_byte(r, hi) == '>' ? 1ull : 0ull);
            }
        }
    }
    uint32_t map;
    if (MODE == MODE_FASTQ) {
        const uint32_t n = (uint32_t)block_reduce64(acc, sh8l, OpAdd(), 0ull);
        map = 0;
        for (uint32_t h = 0; h < 4; ++h) map |= ((h + n) & 3u) << (2 * h);
    } else {
        const unsigned long long last = block_reduce64(acc, sh8l, OpMax(), 0ull);
        const uint32_t o = (uint32_t)(last & 1ull);
        map = last ? (o | (o << 2) | (o << 4) | (o << 6)) : 0xE4u;
    }
    if (tid == 0) st_map[st] = map;
}

__device__ __forceinline__ uint32_t fcompose(uint32_t first, uint32_t then) {  // 2 bits per state
    uint32_t r = 0;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        const uint32_t mid = (first >> (2 * h)) & 3u;
        r |= ((then >> (2 * mid)) & 3u) << (2 * h);
    }
    return r;
}
// One block of 1024 threads: exclusive scan (function composition) of the supertile state maps.
__global__ void __launch_bounds__(1024)
phase_scan_kernel(const uint32_t *__restrict__ st_map, ChunkGeom g, ParseCarry *carry, uint32_t *__restrict__ st_state,
                  const uint8_t *__restrict__ raw, int has_maps) {
    __shared__ uint32_t w_fun[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint32_t IDENT = 0xE4u;
    const uint32_t G = (g.n_st + 1023u) / 1024u;
    const uint32_t a = min(tid * G, g.n_st), b = min(a + G, g.n_st);
    uint32_t mine = IDENT;
    if (has_maps) for (uint32_t i = a; i < b; ++i) mine = fcompose(mine, st_map[i]);
    uint32_t f = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t h = __shfl_up_sync(0xffffffffu, f, d);
        if (lane >= (uint32_t)d) f = fcompose(h, f);
    }
    if (lane == 31u) w_fun[wid] = f;
    uint32_t f_ex = __shfl_up_sync(0xffffffffu, f, 1);
    if (lane == 0u) f_ex = IDENT;
    __syncthreads();
    if (wid == 0u) {
        uint32_t q = w_fun[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t h = __shfl_up_sync(0xffffffffu, q, d);
            if (lane >= (uint32_t)d) q = fcompose(h, q);
        }
        uint32_t q_ex = __shfl_up_sync(0xffffffffu, q, 1);
        if (lane == 0u) q_ex = IDENT;
        w_fun[lane] = q_ex;
    }
    __syncthreads();
    uint32_t before = fcompose(w_fun[wid], f_ex);
    const uint32_t st0 = carry->state & 3u;
    for (uint32_t i = a; i < b; ++i) {
        st_state[i] = (before >> (2 * st0)) & 3u;
        if (has_maps) before = fcompose(before, st_map[i]);
    }
    __syncthreads();  // every thread has read carry->state
    if (tid == 1023u) {
        carry->state = (before >> (2 * st0)) & 3u;
        carry->cprev1 = carry->prev1; carry->cprev2 = carry->prev2;
        if (g.len >= 2) { carry->prev2 = raw[g.len - 2]; carry->prev1 = raw[g.len - 1]; }
        else if (g.len == 1) { carry->prev2 = carry->prev1; carry->prev1 = raw[0]; }
        carry->chunk_syms = 0;
        carry->max_region_syms = 0;
        carry->max_region_pieces = 0;
        carry->chunk_raw_base = carry->raw_total;
        carry->raw_total += g.len;
    }
}

// ---- FASTQ record check: sequence line and quality line must have the same length (needletail's reader;
// lib.rs:63 panics on the error).  With E0..E3 the newlines that end the header, sequence, '+' and quality line
// (chunk-relative position, bit 31 = preceded by a CR, which the reader trims):
//   len(seq) = E1 - E0 - 1 - cr(E1),  len(qual) = E3 - E2 - 1 - cr(E3).
// Returns the chunk-relative position of E0 when the record is invalid, else ~0.
__device__ __forceinline__ uint32_t fastq_len_check(uint32_t e0, uint32_t e1, uint32_t e2, uint32_t e3) {
    const uint32_t M = 0x7FFFFFFFu;
    const uint32_t ls = (e1 & M) - (e0 & M) - 1u - (e1 >> 31), lq = (e3 & M) - (e2 & M) - 1u - (e3 >> 31);
    return ls == lq ? 0xFFFFFFFFu : (e0 & M);
}
// The block's window of the last three newlines seen so far in its supertile (c_nl[2] the most recent).
__device__ __forceinline__ void nl_window_push(uint32_t *c_nl, uint32_t v) { c_nl[0] = c_nl[1]; c_nl[1] = c_nl[2]; c_nl[2] = v; }

// ------------------------------------------------------------------------------------------------
// Exact tile walk of one supertile: 16 raw bytes per thread, every byte classified, line state by
// block scans, symbols written byte by byte.  Handles anything (blanks / CRs inside sequence lines,
// thousands of lines per tile); pack_kernel below falls back to it when its fast path declines.
template <int MODE>
__device__ __forceinline__ void pack_exact(const uint8_t *__restrict__ raw, const ChunkGeom &g, uint32_t t0, uint32_t t1,
                                           uint64_t raw_base, uint32_t cprev1, uint32_t cprev2, uint32_t state,
                                           uint8_t *__restrict__ region, const uint8_t *lut, uint32_t *sh8,
                                           uint32_t &out_off, long long &bases_delta, uint32_t &recs,
                                           unsigned long long &bad_pos, uint16_t *t_nl, uint32_t *c_nl,
                                           uint32_t &nl_before, uint32_t &lbad, SeamNl *seam_st) {
    const int tid = threadIdx.x;
    uint4 nextv = fetch_raw(raw, t0 * (uint32_t)TILE_BYTES + (uint32_t)tid * 16u, g.len);
    for (uint32_t t = t0; t < t1; ++t) {
        const uint32_t off = t * (uint32_t)TILE_BYTES + (uint32_t)tid * 16u;
        Raw16 r;
        const uint4 curv = nextv;
        if (t + 1 < t1) nextv = fetch_raw(raw, off + (uint32_t)TILE_BYTES, g.len);   // next tile's bytes in flight
        load_raw(raw, off, g.len, r, &curv);
        uint32_t em = 0, four = 0;
        SeqCls c;

        if (MODE == MODE_LINES) {
            const uint32_t sm = r.valid & ~r.nl;
            classify_seq(r, sm, lut, c);
            em = (sm & ~c.wsm) | r.nl;     // '\n' separates records
            four = c.bad | r.nl;
        } else if (MODE == MODE_FASTQ) {
            uint32_t total_nl;
            const uint32_t rel = block_exscan_add(__popc(r.nl), sh8, total_nl);
            const uint32_t ph0 = (state + rel) & 3u;
            const uint32_t qm = fastq_seq_mask(r.nl, ph0) & r.valid;
            const uint32_t sm = qm & ~r.nl;
            if (qm) classify_seq(r, sm, lut, c);
            else { c.wsm = c.cr = c.bad = 0; c.cw[0] = c.cw[1] = c.cw[2] = c.cw[3] = 0; }
            em = (sm & ~c.wsm) | (qm & r.nl);
            four = c.bad | (qm & r.nl);
            const uint32_t p1 = prev_byte(r, raw, off, cprev1);
            // sequence().len(): the line without its '\n' and without one CR right before it
            const uint32_t crprev = ((c.cr << 1) | (p1 == '\r' ? 1u : 0u)) & 0xFFFFu;
            bases_delta += (long long)__popc(sm) - (long long)__popc(qm & r.nl & crprev);
            // line-start checks: phase 0 must start with '@', phase 2 with '+'
            uint32_t ls = ((r.nl << 1) | (p1 == '\n' ? 1u : 0u)) & r.valid;
            while (ls) {
                const int i = __ffs(ls) - 1;
                ls &= ls - 1u;
                const uint32_t ph = (ph0 + __popc(r.nl & ((1u << i) - 1u))) & 3u;
                if (ph == 0u) recs++;
                bool ok = true;
                if (ph == 0u || ph == 2u) ok = raw[off + (uint32_t)i] == (ph == 0u ? '@' : '+');
                if (!ok) bad_pos = min(bad_pos, (unsigned long long)(raw_base + off + (uint32_t)i));
            }
            // sequence / quality length check (see fastq_len_check): this tile's newlines as a sorted list in
            // shared memory (t_nl: tile-relative position, bit 15 = preceded by a CR)
            {
                uint32_t m = r.nl, at = rel;
                while (m) {
                    const int i = __ffs(m) - 1;
                    m &= m - 1u;
                    const uint32_t pb = i >= 1 ? raw_byte(r, i - 1) : p1;
                    t_nl[at++] = (uint16_t)(((uint32_t)tid * 16u + (uint32_t)i) | (pb == '\r' ? 0x8000u : 0u));
                }
                __syncthreads();
                const uint32_t tile_off = t * (uint32_t)TILE_BYTES;
                auto nlq = [&](int j) -> uint32_t {      // j-th newline of the tile, or (j < 0) of the carried window
                    if (j >= 0) { const uint32_t v = t_nl[j]; return (tile_off + (v & 0x7FFFu)) | ((v >> 15) << 31); }
                    return c_nl[3 + j];
                };
                m = r.nl;
                int j = (int)rel;
                while (m) {
                    m &= m - 1u;
                    const uint32_t gidx = nl_before + (uint32_t)j;            // index of this newline in the supertile
                    if (gidx < 3u) seam_st->first[gidx] = nlq(j);
                    else if (((state + (uint32_t)j) & 3u) == 3u) lbad = min(lbad, fastq_len_check(nlq(j - 3), nlq(j - 2), nlq(j - 1), nlq(j)));
                    ++j;
                }
                __syncthreads();
                if (tid == 0) {
                    const uint32_t n0 = total_nl > 3u ? total_nl - 3u : 0u;
                    for (uint32_t q = n0; q < total_nl; ++q) nl_window_push(c_nl, nlq((int)q));
                }
                nl_before += total_nl;
                __syncthreads();
            }
            state = (state + total_nl) & 3u;
        } else {  // MODE_FASTA
            const uint32_t p1 = prev_byte(r, raw, off, cprev1);
            const uint32_t ls = ((r.nl << 1) | (p1 == '\n' ? 1u : 0u)) & r.valid;
            const uint32_t hs = header_starts(r, ls);
            uint32_t mine = 0;
            if (ls) {
                const int hi = 31 - __clz(ls);
                mine = ((uint32_t)(tid + 1) << 1) | ((hs >> hi) & 1u);
            }
            uint32_t last;
            const uint32_t prev = block_exscan_max(mine, sh8, last);
            const uint32_t st_in = prev ? (prev & 1u) : (state & 1u);
            const uint32_t hm = fasta_header_mask(hs, r.nl, st_in && !(ls & 1u));
            const uint32_t sm = ~hm & ~r.nl & r.valid;
            if (sm) classify_seq(r, sm, lut, c);
            else { c.wsm = c.cr = c.bad = 0; c.cw[0] = c.cw[1] = c.cw[2] = c.cw[3] = 0; }
            em = (sm & ~c.wsm) | hs;
            four = c.bad | hs;
            recs += __popc(hs);
            bases_delta += __popc(~hm & r.valid);
            // a new header ends the previous record: its raw sequence loses the final '\n'
            // (and one CR before it) when that newline closed a sequence line
            uint32_t h = hs;
            while (h) {
                const int i = __ffs(h) - 1;
                h &= h - 1u;
                const bool prev_in_hdr = i >= 1 ? ((hm >> (i - 1)) & 1u) : (st_in != 0u);
                if (!prev_in_hdr) {
                    bases_delta -= 1;
                    uint32_t b2;
                    if (i >= 2) b2 = raw_byte(r, i - 2);
                    else if (i == 1) b2 = p1;
                    else b2 = off >= 2 ? raw[off - 2] : (off == 1 ? cprev1 : cprev2);
                    if (b2 == '\r') bases_delta -= 1;
                }
            }
            if (last) state = last & 1u;
        }

        // ---- emit ---------------------------------------------------------------------------
        uint32_t tile_total;
        const uint32_t local = block_exscan_add(__popc(em), sh8, tile_total);
        if (em) {
            four &= em;
            uint32_t cw[4];
            if (four) code_words(c, four, cw);
            else { cw[0] = c.cw[0]; cw[1] = c.cw[1]; cw[2] = c.cw[2]; cw[3] = c.cw[3]; }
            // a contiguous run of flags (whole piece, prefix or suffix: nearly always) is written with
            // statically indexed predicated stores
            uint8_t *o = region + out_off + local;
            const int a0 = __ffs(em) - 1;
            const uint32_t run = em >> a0;
            if ((run & (run + 1u)) == 0u) {
                uint8_t *ob = o - a0;
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if ((em >> i) & 1u) ob[i] = (uint8_t)((cw[i >> 2] >> (8 * (i & 3))) & 0xFFu);
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if ((em >> i) & 1u) *o++ = (uint8_t)((cw[i >> 2] >> (8 * (i & 3))) & 0xFFu);
            }
        }
        out_off += tile_total;
    }
}

// ------------------------------------------------------------------------------------------------
// pack_kernel: one block per supertile.  Fast path, per batch of 8 KiB staged in shared memory:
//   1. 2 x 16 bytes per thread -> shared memory; newline masks from registers
//   2. block scan of newline counts -> sorted newline positions s_nl[]: the batch is now a list of
//      line pieces (N newlines -> N + 1 pieces; piece 0 / piece N may be cut by the batch edges)
//   3. one thread per piece: line state of the piece (FASTQ: phase = state + index; FASTA: header
//      iff the line starts with '>'), the record / base counters and FASTQ '@' / '+' checks of the
//      reference parser, and the piece's output entry (source offset, kept length, trailing BREAK);
//      block scan of the output lengths; every 16-byte output word learns its first entry
//   4. one thread per ALIGNED 16-byte output word: unaligned 16-byte read from shared memory
//      (5 LDS + 4 PRMT), SIMD-in-register base codes, one 16-byte store
// Only sequence bytes are classified and every store is a full aligned word.  The fast path
// assumes what normalize(false) would remove inside a sequence line is at most one CR at its end;
// a blank or CR anywhere else, or more than PB_NLCAP newlines in a batch, makes the block redo its
// supertile with pack_exact.
constexpr int PB_TILES = 4;                        // tiles per batch (16 bytes per thread each)
constexpr int PB_BYTES = PB_TILES * TILE_BYTES;    // <= 16 KiB: piece lengths must fit 15 bits below the break flag
constexpr int PB_NLCAP = 1535;
static_assert(PB_TILES == 2 || PB_TILES == 4, "newline counts are scanned as 16-bit fields of one 64-bit word");
static_assert(PB_BYTES < 0x8000, "e_len keeps a flag in bit 15");

__device__ __forceinline__ void ones128(uint32_t nbytes, uint64_t &lo, uint64_t &hi) {   // nbytes in 0..16
    lo = nbytes >= 8u ? ~0ULL : ((1ULL << (8u * nbytes)) - 1ULL);
    hi = nbytes <= 8u ? 0ULL : (nbytes >= 16u ? ~0ULL : ((1ULL << (8u * (nbytes - 8u))) - 1ULL));
}

template <int MODE>
__global__ void __launch_bounds__(TILE_THREADS, 4)   // 64 registers: four blocks per SM
pack_kernel(const uint8_t *__restrict__ raw, ChunkGeom g, ParseCarry *carry, const uint32_t *__restrict__ st_state,
            uint8_t *__restrict__ sym, uint32_t *__restrict__ region_count, SeamNl *__restrict__ seam) {
    __shared__ uint8_t lut[256];
    __shared__ uint32_t sh8[8];
    __shared__ unsigned long long sh8l[8];
    __shared__ __align__(16) uint8_t s_rawbuf[16 + PB_BYTES + 32];   // [16 front pad | batch | back pad]
    __shared__ uint16_t s_nl[PB_NLCAP + 1];
    __shared__ uint16_t e_src[PB_NLCAP + 1], e_len[PB_NLCAP + 1], e_out[PB_NLCAP + 1];
    __shared__ uint16_t s_first[PB_BYTES / 16 + 2];
    __shared__ uint32_t s_flag[2];   // [0] fast path declined, [1] FASTA: state after the batch
    __shared__ uint32_t c_nl[3];     // FASTQ: the last three newlines of this supertile before the current batch
    __shared__ uint16_t s_bw[PB_BYTES / 16 + 8];   // output words that need the general (multi-entry) path
    __shared__ uint32_t s_nbw;
    const int tid = threadIdx.x;
    lut[tid] = classify_byte((uint8_t)tid);
    if (tid < 2) s_flag[tid] = 0;
    __syncthreads();
    const uint32_t st = blockIdx.x;
    const uint32_t t0 = st * g.st_tiles, t1 = min(t0 + g.st_tiles, g.n_tiles);
    const uint32_t B0 = t0 * (uint32_t)TILE_BYTES, B1 = min(t1 * (uint32_t)TILE_BYTES, g.len);
    const uint64_t raw_base = carry->chunk_raw_base;
    const uint32_t cprev1 = carry->cprev1, cprev2 = carry->cprev2;
    uint8_t *region = sym + (size_t)SYM_FRONT + (size_t)st * g.region_stride;
    const uint32_t state0 = MODE == MODE_LINES ? 0u : st_state[st];  // state of the line holding the previous byte
    constexpr uint32_t STEP = MODE == MODE_FASTQ ? 4u : 1u;          // FASTQ: only every 4th piece emits

    uint32_t state = state0;
    uint32_t out_off = 0;                                     // symbols written so far in this region
    long long bases_delta = 0;
    uint32_t recs = 0;
    unsigned long long bad_pos = ~0ULL;
    uint32_t nl_before = 0, lbad = 0xFFFFFFFFu;               // FASTQ: newlines of this supertile so far; min bad record
    SeamNl *seam_st = MODE == MODE_FASTQ ? seam + st : nullptr;
    uint8_t *rawb = s_rawbuf + 16;
    const uint32_t *raw32 = reinterpret_cast<const uint32_t *>(s_rawbuf);
    bool declined = false;

    uint4 nextv[PB_TILES];
#pragma unroll
    for (int p = 0; p < PB_TILES; ++p) nextv[p] = fetch_raw(raw, B0 + (uint32_t)p * TILE_BYTES + (uint32_t)tid * 16u, B1);
    for (uint32_t b = B0; b < B1; b += PB_BYTES) {
        const uint32_t blen = min((uint32_t)PB_BYTES, B1 - b);
        // ---- 1. stage the batch, newline masks --------------------------------------------------
        uint32_t nlm[PB_TILES];
#pragma unroll
        for (int p = 0; p < PB_TILES; ++p) {
            const uint32_t po = (uint32_t)p * TILE_BYTES + (uint32_t)tid * 16u;
            const uint4 curv = nextv[p];
            if (b + PB_BYTES < B1) nextv[p] = fetch_raw(raw, b + PB_BYTES + po, B1);   // next batch in flight
            Raw16 r;
            if (po < blen) load_raw(raw, b + po, B1, r, &curv);
            else { r.w[0] = r.w[1] = r.w[2] = r.w[3] = 0; r.nl = 0; }
            *reinterpret_cast<uint4 *>(rawb + po) = make_uint4(r.w[0], r.w[1], r.w[2], r.w[3]);
            nlm[p] = r.nl;
        }
        if (tid == 0) {   // the two bytes in front of the batch
            s_nbw = 0;
            rawb[-1] = (uint8_t)(b >= 1u ? raw[b - 1] : cprev1);
            rawb[-2] = (uint8_t)(b >= 2u ? raw[b - 2] : (b == 1u ? cprev1 : cprev2));
        }
        // ---- 2. newline positions ---------------------------------------------------------------
        unsigned long long cnt4 = 0, tot4;
#pragma unroll
        for (int p = 0; p < PB_TILES; ++p) cnt4 |= (unsigned long long)__popc(nlm[p]) << (16 * p);
        const unsigned long long ex4 = block_exscan_add64(cnt4, sh8l, tot4);
        uint32_t N = 0;
#pragma unroll
        for (int p = 0; p < PB_TILES; ++p) N += (uint32_t)(tot4 >> (16 * p)) & 0xFFFFu;
        if (N > (uint32_t)PB_NLCAP) { declined = true; break; }   // uniform
        uint32_t before = 0;                                      // newlines of the pieces in front
#pragma unroll
        for (int p = 0; p < PB_TILES; ++p) {
            uint32_t m = nlm[p], at = before + ((uint32_t)(ex4 >> (16 * p)) & 0xFFFFu);
            before += (uint32_t)(tot4 >> (16 * p)) & 0xFFFFu;
            const uint32_t po = (uint32_t)p * TILE_BYTES + (uint32_t)tid * 16u;
            while (m) {
                s_nl[at++] = (uint16_t)(po + (uint32_t)(__ffs(m) - 1));
                m &= m - 1u;
            }
        }
        __syncthreads();
        // ---- 3. pieces -> entries ---------------------------------------------------------------
        const uint32_t NP = N + 1u;
        const uint32_t G = (NP + TILE_THREADS - 1u) / TILE_THREADS;
        const uint32_t i_lo = min((uint32_t)tid * G, NP), i_hi = min(i_lo + G, NP);
        const bool ls0 = rawb[-1] == '\n';
        // FASTA: the piece holding the batch's last byte (the trailing piece unless it is empty)
        const uint32_t fa_tgt = ((N ? (uint32_t)s_nl[N - 1u] + 1u : 0u) < blen) ? N : N - 1u;
        uint32_t sum = 0;
        for (uint32_t i = i_lo; i < i_hi; ++i) {
            const uint32_t start = i ? (uint32_t)s_nl[i - 1] + 1u : 0u;
            const bool has_nl = i < N;
            const uint32_t end = has_nl ? (uint32_t)s_nl[i] : blen;
            const uint32_t len = end - start;
            const bool line_start = i > 0u || ls0;
            const uint32_t pb = rawb[(int)end - 1];            // byte in front of the piece's end
            uint32_t klen = len - ((len > 0u && pb == '\r') ? 1u : 0u);   // one trailing CR goes
            uint32_t L = 0;
            if (MODE == MODE_LINES) {
                L = klen | (has_nl ? 0x8000u : 0u);
            } else if (MODE == MODE_FASTQ) {
                const uint32_t ph = (state + i) & 3u;
                if (line_start && start < blen) {
                    if (ph == 0u) recs++;
                    if ((ph == 0u || ph == 2u) && rawb[start] != (ph == 0u ? '@' : '+'))
                        bad_pos = min(bad_pos, (unsigned long long)(raw_base + b + start));
                }
                if (ph == 1u) {
                    // sequence().len(): the line without its '\n' and without one CR right before it
                    bases_delta += (long long)len - ((has_nl && pb == '\r') ? 1 : 0);
                    L = klen | (has_nl ? 0x8000u : 0u);
                }
                if (has_nl) {   // sequence / quality length check (see fastq_len_check)
                    auto nlq = [&](int j) -> uint32_t {    // j-th newline of the batch, or (j < 0) of the carried window
                        if (j >= 0) { const uint32_t q = s_nl[j]; return (b + q) | (rawb[(int)q - 1] == '\r' ? 0x80000000u : 0u); }
                        return c_nl[3 + j];
                    };
                    if (i >= 3u) {   // the usual case: the record's four newlines are all in this batch (klen is the quality's length)
                        if (ph == 3u) {
                            const uint32_t e1 = s_nl[i - 2], e0 = s_nl[i - 3];
                            const uint32_t ls = e1 - e0 - 1u - (rawb[(int)e1 - 1] == '\r' ? 1u : 0u);
                            if (ls != klen) lbad = min(lbad, b + e0);
                        }
                    } else {         // the first three pieces: newlines carried from earlier batches, or left to the seam check
                        const uint32_t gidx = nl_before + i;
                        if (gidx < 3u) seam_st->first[gidx] = nlq((int)i);
                        else if (ph == 3u) lbad = min(lbad, fastq_len_check(nlq((int)i - 3), nlq((int)i - 2), nlq((int)i - 1), nlq((int)i)));
                    }
                }
            } else {  // MODE_FASTA
                const bool hdr = line_start ? (start < blen && rawb[start] == '>') : ((state & 1u) != 0u);
                if (line_start && hdr) {
                    recs++;
                    L = 0x8000u;                               // the header start is a record break
                    bool prev_in_hdr;
                    if (i == 0u) prev_in_hdr = state != 0u;
                    else {
                        const uint32_t ps = i > 1u ? (uint32_t)s_nl[i - 2] + 1u : 0u;
                        prev_in_hdr = (i > 1u || ls0) ? rawb[ps] == '>' : ((state & 1u) != 0u);
                    }
                    // a new header ends the previous record: its raw sequence loses the final '\n'
                    // (and one CR before it) when that newline closed a sequence line
                    if (!prev_in_hdr) {
                        bases_delta -= 1;
                        if (rawb[(int)start - 2] == '\r') bases_delta -= 1;
                    }
                } else if (!hdr) {
                    bases_delta += (long long)len + (has_nl ? 1 : 0);
                    L = klen;
                }
                // state for the next batch: the line holding this batch's last byte
                if (i == fa_tgt) s_flag[1] = hdr ? 1u : 0u;
            }
            e_src[i] = (uint16_t)start;
            e_len[i] = (uint16_t)L;
            e_out[i] = (uint16_t)sum;
            sum += (L & 0x7FFFu) + (L >> 15);
        }
        uint32_t T;
        const uint32_t exo = block_exscan_add(sum, sh8, T);
        const uint32_t a = out_off & 15u;
        for (uint32_t i = i_lo; i < i_hi; ++i) {
            const uint32_t eo = (uint32_t)e_out[i] + exo, L = e_len[i];
            const uint32_t outlen = (L & 0x7FFFu) + (L >> 15);
            e_out[i] = (uint16_t)eo;
            if (outlen) {   // words whose first byte this entry provides
                if (eo == 0u) s_first[0] = (uint16_t)i;
                const uint32_t w_hi = (eo + outlen - 1u + a) >> 4;
                for (uint32_t w = (eo + a + 15u) >> 4; w <= w_hi; ++w) s_first[w] = (uint16_t)i;
            }
        }
        __syncthreads();
        if (MODE == MODE_FASTQ) {
            if (tid == 0) {
                const uint32_t n0 = N > 3u ? N - 3u : 0u;
                for (uint32_t q = n0; q < N; ++q) {
                    const uint32_t pos = s_nl[q];
                    nl_window_push(c_nl, (b + pos) | (rawb[(int)pos - 1] == '\r' ? 0x80000000u : 0u));
                }
            }
            nl_before += N;
        }
        // ---- 4. aligned output words ------------------------------------------------------------
        const uint32_t W = (a + T + 15u) >> 4;
        uint8_t *obase = region + (out_off - a);
        // Pass A: every word that comes out of ONE line whole (9 of 10 words of 150 bp reads): no masks, no merging --
        // unaligned 16-byte read, SIMD-in-register codes, one aligned store.  The others (words holding a line end, the
        // ragged first / last word of the batch, anything but ACGTacgt) are only LISTED here and done in pass B by as
        // few warps as it takes: inside a warp they would drag all 32 lanes through the general path.
        for (uint32_t w = tid; w < W; w += TILE_THREADS) {
            const uint32_t f0 = w == 0u ? a : 0u;
            const uint32_t pos = 16u * w + f0 - a;             // output position within the batch
            const uint32_t i = s_first[w];
            const uint32_t d = pos - (uint32_t)e_out[i];
            if (f0 == 0u && d + 16u <= ((uint32_t)e_len[i] & 0x7FFFu)) {
                const uint32_t A = 16u + (uint32_t)e_src[i] + d;                  // s_rawbuf offset of output byte 0
                const uint32_t wi = A >> 2, sel = 0x3210u + 0x1111u * (A & 3u);
                uint32_t x[5], cw[4], diff = 0;
#pragma unroll
                for (int j = 0; j < 5; ++j) x[j] = raw32[wi + j];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t u = __byte_perm(x[j], x[j + 1], sel) & 0xDFDFDFDFu;
                    const uint32_t c2 = ((u >> 1) ^ (u >> 2)) & 0x03030303u;
                    cw[j] = c2;
                    uint32_t z = (c2 | (c2 >> 4)) & 0x00FF00FFu;
                    z = (z | (z >> 8)) & 0xFFFFu;
                    diff |= __byte_perm(0x54474341u, 0u, z) ^ u;
                }
                if (diff == 0u) {
                    *reinterpret_cast<uint4 *>(obase + 16u * w) = make_uint4(cw[0], cw[1], cw[2], cw[3]);
                    continue;
                }
            }
            s_bw[atomicAdd(&s_nbw, 1u)] = (uint16_t)w;
        }
        __syncthreads();
        // Pass B: the listed words, densely over the threads
        const uint32_t nbw = s_nbw;
        for (uint32_t idx = tid; idx < nbw; idx += TILE_THREADS) {
            const uint32_t w = s_bw[idx];
            const uint32_t f0 = w == 0u ? a : 0u;
            uint32_t filled = f0;
            uint32_t pos = 16u * w + filled - a;               // output position within the batch
            uint32_t i = s_first[w];
            uint32_t d = pos - (uint32_t)e_out[i];
            uint64_t alo = 0, ahi = 0;
            while (filled < 16u && pos < T) {
                const uint32_t L = e_len[i], len = L & 0x7FFFu, brk = L >> 15;
                if (d < len) {
                    const uint32_t n = min(16u - filled, len - d);
                    const uint32_t A = 16u + (uint32_t)e_src[i] + d - filled;   // s_rawbuf offset of output byte 0
                    const uint32_t wi = A >> 2, sel = 0x3210u + 0x1111u * (A & 3u);
                    uint32_t x[5], v[4];
#pragma unroll
                    for (int j = 0; j < 5; ++j) x[j] = raw32[wi + j];
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = __byte_perm(x[j], x[j + 1], sel);
                    uint64_t l0, h0, l1, h1;
                    ones128(filled, l0, h0);
                    ones128(filled + n, l1, h1);
                    const uint64_t mlo = l1 & ~l0, mhi = h1 & ~h0;
                    const uint32_t m[4] = {(uint32_t)mlo, (uint32_t)(mlo >> 32), (uint32_t)mhi, (uint32_t)(mhi >> 32)};
                    uint32_t cw[4], diff = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t u = v[j] & 0xDFDFDFDFu;
                        const uint32_t c2 = ((u >> 1) ^ (u >> 2)) & 0x03030303u;
                        cw[j] = c2;
                        uint32_t z = (c2 | (c2 >> 4)) & 0x00FF00FFu;
                        z = (z | (z >> 8)) & 0xFFFFu;
                        diff |= (__byte_perm(0x54474341u, 0u, z) ^ u) & m[j];
                    }
                    if (diff) {   // exact per-byte classes (needletail normalize(false))
                        bool blank = false;
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            const uint32_t k = lut[(v[q >> 2] >> (8 * (q & 3))) & 0xFFu];
                            const bool in = (uint32_t)q >= filled && (uint32_t)q < filled + n;
                            blank |= in && (k == CLS_WS || k == CLS_CR);
                            cw[q >> 2] = (cw[q >> 2] & ~(0xFFu << (8 * (q & 3)))) | ((k < 4u ? k : (uint32_t)SYM_BREAK) << (8 * (q & 3)));
                        }
                        if (blank) s_flag[0] = 1u;
                    }
                    alo |= ((uint64_t)cw[0] | ((uint64_t)cw[1] << 32)) & mlo;
                    ahi |= ((uint64_t)cw[2] | ((uint64_t)cw[3] << 32)) & mhi;
                    filled += n; pos += n; d += n;
                }
                if (filled < 16u && d == len && brk) {
                    if (filled < 8u) alo |= (uint64_t)SYM_BREAK << (8u * filled);
                    else ahi |= (uint64_t)SYM_BREAK << (8u * (filled - 8u));
                    filled++; pos++; d++;
                }
                if (d >= len + brk) { i += STEP; d = 0; }
            }
            uint8_t *o = obase + 16u * w;
            if (f0 == 0u && filled == 16u) {
                *reinterpret_cast<uint4 *>(o) = make_uint4((uint32_t)alo, (uint32_t)(alo >> 32), (uint32_t)ahi, (uint32_t)(ahi >> 32));
            } else {   // first / last word of the batch
                for (uint32_t q = f0; q < filled; ++q)
                    o[q] = (uint8_t)((q < 8u ? (alo >> (8u * q)) : (ahi >> (8u * (q - 8u)))) & 0xFFu);
            }
        }
        out_off += T;
        if (MODE == MODE_FASTQ) state = (state + N) & 3u;
        __syncthreads();
        if (MODE == MODE_FASTA) state = s_flag[1];
        if (s_flag[0]) { declined = true; break; }   // uniform: read after the barrier
    }

    if (declined) {
        __syncthreads();
        out_off = 0; bases_delta = 0; recs = 0; bad_pos = ~0ULL; nl_before = 0; lbad = 0xFFFFFFFFu;
        pack_exact<MODE>(raw, g, t0, t1, raw_base, cprev1, cprev2, state0, region, lut, sh8, out_off, bases_delta, recs, bad_pos,
                         reinterpret_cast<uint16_t *>(s_rawbuf), c_nl, nl_before, lbad, seam_st);
    }
    if (MODE == MODE_FASTQ) {
        __syncthreads();
        if (tid == 0) { seam_st->last[0] = c_nl[0]; seam_st->last[1] = c_nl[1]; seam_st->last[2] = c_nl[2]; seam_st->n = nl_before; }
    }

    // the hash kernel walks up to HASH_W positions past the region's end: make them breaks
    if (tid < HASH_W) region[out_off + tid] = SYM_BREAK;
    if (tid == 0) { region_count[st] = out_off; atomicAdd(&carry->chunk_syms, out_off); atomicMax(&carry->max_region_syms, out_off); }
    if (MODE != MODE_LINES) {
        const unsigned long long bsum = block_reduce64((unsigned long long)bases_delta, sh8l, OpAdd(), 0ull);
        const unsigned long long rsum = block_reduce64(recs, sh8l, OpAdd(), 0ull);
        if (tid == 0) {
            if (bsum) atomicAdd((unsigned long long *)&carry->total_bases, bsum);
            if (rsum) atomicAdd((unsigned long long *)&carry->n_records, rsum);
        }
        if (MODE == MODE_FASTQ) {
            const unsigned long long bmin = block_reduce64(bad_pos, sh8l, OpMin(), ~0ULL);
            if (tid == 0 && bmin != ~0ULL) atomicMin((unsigned long long *)&carry->first_bad_pos, bmin);
            const unsigned long long lmin = block_reduce64(lbad == 0xFFFFFFFFu ? ~0ULL : raw_base + lbad, sh8l, OpMin(), ~0ULL);
            if (tid == 0 && lmin != ~0ULL) atomicMin((unsigned long long *)&carry->len_bad_pos, lmin);
        }
    }
}

// ================================================================================================
// parse_fused_kernel: phase + scan + pack in ONE pass over the raw bytes.
//
// The three-kernel pipeline above reads every raw byte twice (phase_kernel counts newlines so that pack_kernel can
// start each supertile in the right line state) and pays two launches for a 2-bit piece of information.  Here a
// block owns one supertile of <= 32 KiB, stages it in shared memory once, and learns its start state from its
// predecessors while it works -- a decoupled look-back on a per-supertile status word:
//     status[st] = epoch << 16 | kind << 8 | value
//     kind 1 (aggregate): what the supertile does to the line state -- FASTQ: its newline count mod 4;
//                         FASTA: 0 = holds no line start (identity), 2 | h = its last line start is (h = 1) or is
//                         not (h = 0) a header line
//     kind 2 (inclusive): the line state AFTER the supertile (the state of the line holding its last byte)
// Supertile indices are handed out by a ticket counter, so every predecessor of a running block is running or done
// (no deadlock whatever the block scheduling order); `epoch` (the chunk number) makes stale words of earlier chunks
// invisible without clearing the array.  The whole supertile is ONE batch: one newline list, one piece pass, one
// output pass, symbols start at the region's first byte (no ragged first word).
//
// Everything else is pack_kernel's scheme (line pieces -> entries -> aligned 16-byte output words; words cut by a
// line end or holding anything but ACGTacgt are listed and done densely in a second pass; blanks / CRs inside
// sequence lines or more than PF_NLCAP newlines make the block redo its supertile with pack_exact).
constexpr int PF_MAX_TILES = 8;
constexpr int PF_BYTES = PF_MAX_TILES * TILE_BYTES;      // 32 KiB
constexpr int PF_NLCAP = 2047;
constexpr int PF_LIST = 3900;                            // newline list (<= PF_NLCAP + 1 entries), then the listed output words (<= PF_BYTES / 16 + 1):
                                                         // behind the newlines while the state is a guess (a redo needs them again), else over them
constexpr int PF_ENT = 2050;                             // entries (PF_NLCAP + 1 pieces)
constexpr int PF_FIRST = PF_BYTES / 16 + 2;
constexpr uint32_t PF_SMEM = 16u + PF_BYTES + 32u + 2u * (PF_LIST + 3 * PF_ENT + PF_FIRST);
static_assert(PF_LIST >= PF_NLCAP + 1 && PF_LIST >= PF_BYTES / 16 + 1 && PF_ENT >= PF_NLCAP + 2, "list sizes");
static_assert(4 * (PF_SMEM + 1024 + 256) <= 228 * 1024 && PF_SMEM + 256 <= 227 * 1024 / 4 + 1024, "four blocks per SM");

struct ClsTable {
    uint8_t v[256];
    constexpr ClsTable() : v{} { for (int i = 0; i < 256; ++i) v[i] = classify_byte((uint8_t)i); }
};
__device__ const ClsTable g_cls_table{};

__device__ __forceinline__ uint32_t nl_mask16(const uint4 v) {
    return gather4(zero_bytes80(v.x ^ 0x0A0A0A0Au)) | (gather4(zero_bytes80(v.y ^ 0x0A0A0A0Au)) << 4) |
           (gather4(zero_bytes80(v.z ^ 0x0A0A0A0Au)) << 8) | (gather4(zero_bytes80(v.w ^ 0x0A0A0A0Au)) << 12);
}
// two packed (4 x 16-bit) counters scanned together: one pair of barriers
__device__ __forceinline__ void block_exscan_add64x2(unsigned long long &a, unsigned long long &b, unsigned long long *sh16l,
                                                     unsigned long long &ta, unsigned long long &tb) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned long long xa = a, xb = b;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long na = __shfl_up_sync(0xffffffffu, xa, d), nb = __shfl_up_sync(0xffffffffu, xb, d);
        if (lane >= d) { xa += na; xb += nb; }
    }
    __syncthreads();  // protect sh16l from a previous use
    if (lane == 31) { sh16l[wid] = xa; sh16l[8 + wid] = xb; }
    __syncthreads();
    unsigned long long basea = 0, baseb = 0, tota = 0, totb = 0;
#pragma unroll
    for (int i = 0; i < TILE_THREADS / 32; ++i) {
        const unsigned long long sa = sh16l[i], sb = sh16l[8 + i];
        if (i < wid) { basea += sa; baseb += sb; }
        tota += sa; totb += sb;
    }
    ta = tota; tb = totb;
    a = basea + xa - a; b = baseb + xb - b;
}
__device__ __forceinline__ uint4 ld_status4(const uint32_t *p) {   // 16-byte aligned, straight from L2
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(uint32_t *p, uint32_t v) { *reinterpret_cast<volatile uint32_t *>(p) = v; }
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// Line state at the start of supertile st (whole warp; see the status word above).  A lane takes one aligned group
// of four status words per round (128 predecessors per round); the entry in front of the chunk's first supertile is
// the carried state of the stream.  Within a lane and across lanes the walk goes from the nearest predecessor back
// to the first inclusive word.
template <int MODE>
__device__ __forceinline__ uint32_t lookback_state(const uint32_t *status, uint32_t st, uint32_t epoch, uint32_t carry_state,
                                                   uint32_t lane) {
    if (st == 0u) return carry_state;
    uint32_t add = 0;                       // FASTQ: newlines of the nearer predecessors folded so far
    int q0 = (int)((st - 1u) >> 2);         // group of the nearest predecessor not folded yet
    while (true) {
        const int q = q0 - (int)lane;
        // lane summary: res = 1 when the lane's group settles the state (an inclusive word; FASTA: also a supertile
        // holding a line start), `val` = that state with the lane's nearer aggregates applied; else val = the
        // lane's aggregates alone (FASTQ: their newline sum)
        uint32_t res, val;
        if (q < 0) { res = 1u; val = carry_state; }
        else {
            uint32_t w[4];
            const uint32_t hi = min(3u, st - 1u - 4u * (uint32_t)q);     // entries above st - 1 are not predecessors
            while (true) {
                const uint4 v = ld_status4(status + 4 * q);
                w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
                bool ready = true;
#pragma unroll
                for (int e = 0; e < 4; ++e) ready &= (uint32_t)e > hi || (w[e] >> 16) == epoch;
                if (ready) break;
            }
            res = 0u; val = 0u;
#pragma unroll
            for (int e = 3; e >= 0; --e) {
                if ((uint32_t)e <= hi && !res) {
                    const uint32_t kind = (w[e] >> 8) & 0xFFu, x = w[e] & 0xFFu;
                    if (MODE == MODE_FASTQ) {
                        if (kind == 2u) { res = 1u; val = (x + val) & 3u; }
                        else val += x;
                    } else {
                        if (kind == 2u || (x & 2u)) { res = 1u; val = x & 1u; }
                    }
                }
            }
        }
        const uint32_t done = __ballot_sync(0xffffffffu, res != 0u);
        const uint32_t upto = done ? (uint32_t)(__ffs(done) - 1) : 32u;   // lanes [0, upto) only add; lane upto settles
        if (MODE == MODE_FASTQ) {
            add += __reduce_add_sync(0xffffffffu, lane < upto ? val : 0u);
            if (done) return (__shfl_sync(0xffffffffu, val, (int)upto) + add) & 3u;
        } else {
            if (done) return __shfl_sync(0xffffffffu, val, (int)upto) & 1u;
        }
        q0 -= 32;
    }
}

template <int MODE>
__global__ void __launch_bounds__(TILE_THREADS, 4)
parse_fused_kernel(const uint8_t *__restrict__ raw, ChunkGeom g, ParseCarry *carry, uint32_t *__restrict__ st_state,
                   uint8_t *__restrict__ sym, uint32_t *__restrict__ region_count, SeamNl *__restrict__ seam,
                   uint32_t *status, uint32_t epoch, uint32_t *ticket, uint32_t ticket_base, PiecePlan pp) {
    extern __shared__ __align__(16) uint8_t pf_smem[];
    __shared__ unsigned long long sh16l[16];
    __shared__ uint32_t sh8[8];
    __shared__ uint32_t c_nl[3];
    __shared__ uint32_t s_misc[12];
    enum { M_ST = 0, M_DECL = 1, M_NBW = 2, M_STATE = 3, M_LASTNL = 4, M_BASES = 5, M_RECS = 6, M_BAD = 7, M_LBAD = 8, M_T = 9, M_GUESS = 10, M_NP = 11 };
    uint8_t *rawb = pf_smem + 16;                                          // [16 front pad | supertile | back pad]
    const uint32_t *raw32 = reinterpret_cast<const uint32_t *>(pf_smem);
    uint16_t *s_nl = reinterpret_cast<uint16_t *>(pf_smem + 16 + PF_BYTES + 32);
    uint16_t *e_src = s_nl + PF_LIST, *e_len = e_src + PF_ENT, *e_out = e_len + PF_ENT, *s_first = e_out + PF_ENT;
    const uint8_t *lut = g_cls_table.v;
    const int tid = threadIdx.x;
    const uint32_t lane = tid & 31u, wid = tid >> 5;
    if (tid == 0) {
        s_misc[M_ST] = atomicAdd(ticket, 1u) - ticket_base;
        s_misc[M_DECL] = 0; s_misc[M_NBW] = 0; s_misc[M_STATE] = 0; s_misc[M_LASTNL] = 0;
        s_misc[M_BASES] = 0; s_misc[M_RECS] = 0; s_misc[M_BAD] = 0xFFFFFFFFu; s_misc[M_LBAD] = 0xFFFFFFFFu;
    }
    __syncthreads();
    const uint32_t st = s_misc[M_ST];
    const uint32_t B0 = st * g.st_bytes, B1 = min(B0 + g.st_bytes, g.len), blen = B1 - B0;
    const uint64_t raw_base = carry->chunk_raw_base;
    const uint32_t cprev1 = carry->cprev1, cprev2 = carry->cprev2;
    const uint32_t carry_state = MODE == MODE_LINES ? 0u : (carry->state & 3u);
    uint8_t *region = sym + (size_t)SYM_FRONT + (size_t)st * g.region_stride;
    SeamNl *seam_st = MODE == MODE_FASTQ ? seam + st : nullptr;
    constexpr uint32_t STEP = MODE == MODE_FASTQ ? 4u : 1u;              // FASTQ: only every 4th piece emits

    // ---- 1. stage the supertile, newline masks --------------------------------------------------------
    uint32_t nlm[PF_MAX_TILES];
    if (blen == (uint32_t)PF_BYTES) {                                    // the usual supertile of a large chunk
        uint4 v[PF_MAX_TILES];
#pragma unroll
        for (int p = 0; p < PF_MAX_TILES; ++p)
            v[p] = __ldg(reinterpret_cast<const uint4 *>(raw + B0 + (uint32_t)p * TILE_BYTES + (uint32_t)tid * 16u));
#pragma unroll
        for (int p = 0; p < PF_MAX_TILES; ++p) {
            *reinterpret_cast<uint4 *>(rawb + (uint32_t)p * TILE_BYTES + (uint32_t)tid * 16u) = v[p];
            nlm[p] = nl_mask16(v[p]);
        }
    } else {
#pragma unroll
        for (int p = 0; p < PF_MAX_TILES; ++p) {
            nlm[p] = 0;
            const uint32_t po = (uint32_t)p * TILE_BYTES + (uint32_t)tid * 16u;
            if ((uint32_t)p * TILE_BYTES < blen) {
                Raw16 r;
                if (po < blen) load_raw(raw, B0 + po, B1, r);
                else { r.w[0] = r.w[1] = r.w[2] = r.w[3] = 0; r.nl = 0; }
                *reinterpret_cast<uint4 *>(rawb + po) = make_uint4(r.w[0], r.w[1], r.w[2], r.w[3]);
                nlm[p] = r.nl;
            }
        }
    }
    if (tid == 0) {   // the two bytes in front of the supertile
        rawb[-1] = (uint8_t)(B0 >= 1u ? raw[B0 - 1] : cprev1);
        rawb[-2] = (uint8_t)(B0 >= 2u ? raw[B0 - 2] : (B0 == 1u ? cprev1 : cprev2));
    }
    if (MODE == MODE_FASTA) {   // position after the last newline that is followed by a byte of this supertile
        uint32_t mx = 0;
#pragma unroll
        for (int p = 0; p < PF_MAX_TILES; ++p) {
            const uint32_t po = (uint32_t)p * TILE_BYTES + (uint32_t)tid * 16u;
            uint32_t m = nlm[p];
            if (m) {
                const uint32_t room = blen - 1u - po;                    // bits below `room` are followed by a byte
                if (room < 16u) m &= (1u << room) - 1u;
                if (m) mx = po + (uint32_t)(31 - __clz(m)) + 1u;
            }
        }
        mx = __reduce_max_sync(0xffffffffu, mx);
        if (lane == 0 && mx) atomicMax(&s_misc[M_LASTNL], mx);
    }
    // ---- 2. newline positions ----------------------------------------------------------------------------
    unsigned long long exa = 0, exb = 0, tota, totb;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        exa |= (unsigned long long)__popc(nlm[p]) << (16 * p);
        exb |= (unsigned long long)__popc(nlm[4 + p]) << (16 * p);
    }
    block_exscan_add64x2(exa, exb, sh16l, tota, totb);                   // (its barriers also publish the staged bytes)
    uint32_t N = 0;
#pragma unroll
    for (int p = 0; p < 4; ++p) N += ((uint32_t)(tota >> (16 * p)) & 0xFFFFu) + ((uint32_t)(totb >> (16 * p)) & 0xFFFFu);
    bool declined = N > (uint32_t)PF_NLCAP;                              // uniform
    // what this supertile does to the line state
    uint32_t agg = 0, after_const = 0;
    bool agg_const = false;
    if (MODE == MODE_FASTQ) agg = N & 3u;
    if (MODE == MODE_FASTA) {
        const uint32_t ln = s_misc[M_LASTNL];
        const bool has_ls = ln != 0u || rawb[-1] == '\n';
        if (has_ls) { after_const = rawb[ln] == '>' ? 1u : 0u; agg = 2u | after_const; agg_const = true; }
    }
    // Warp PF_LB (the last one) is the look-back warp: it publishes the aggregate, and while the other seven warps
    // frame the pieces (3a, which does not need the state) it walks back over the predecessors' status words.
    constexpr uint32_t PF_LB = TILE_THREADS / 32 - 1, PF_WT = PF_LB * 32;   // worker threads
    if (MODE != MODE_LINES && tid == (int)PF_WT) st_status(status + st, (epoch << 16) | (1u << 8) | agg);
    if (!declined) {
        uint32_t before = 0;                                             // newlines of the tiles in front
#pragma unroll
        for (int p = 0; p < PF_MAX_TILES; ++p) {
            const unsigned long long ex = p < 4 ? exa : exb, tot = p < 4 ? tota : totb;
            uint32_t m = nlm[p], at = before + ((uint32_t)(ex >> (16 * (p & 3))) & 0xFFFFu);
            before += (uint32_t)(tot >> (16 * (p & 3))) & 0xFFFFu;
            const uint32_t po = (uint32_t)p * TILE_BYTES + (uint32_t)tid * 16u;
            while (m) {
                s_nl[at++] = (uint16_t)(po + (uint32_t)(__ffs(m) - 1));
                m &= m - 1u;
            }
        }
    }
    // Pieces: N newlines -> N + 1 pieces, G consecutive ones per worker thread.  Geometry of piece i (both passes).
    const uint32_t NP = N + 1u;
    const uint32_t G = (NP + PF_WT - 1u) / PF_WT;
    const uint32_t i_lo = min((uint32_t)tid * G, NP), i_hi = wid == PF_LB ? i_lo : min(i_lo + G, NP);
    unsigned long long vsum = 0, vex = 0, vtot = 0;   // output lengths per variant of the unknown state (16-bit fields)
    unsigned long long psum = 0, pex = 0, ptot = 0;   // hash pieces, likewise
    uint32_t state0 = 0;                             // state of the line holding the byte in front of the supertile
    bool have_state = MODE == MODE_LINES;            // false: state0 is a guess, verified when the look-back is in
    if (wid == PF_LB) {
        named_arrive(1, TILE_THREADS);               // this warp's newlines are in the list
        if (MODE != MODE_LINES) {
            state0 = lookback_state<MODE>(status, st, epoch, carry_state, lane);
            have_state = true;
            if (lane == 0) {
                const uint32_t after = MODE == MODE_FASTQ ? ((state0 + N) & 3u) : (agg_const ? after_const : state0);
                st_status(status + st, (epoch << 16) | (2u << 8) | after);
                st_state[st] = state0;
                s_misc[M_STATE] = state0;
                if (st == g.n_st - 1u) carry->state_next = after;
            }
            __syncwarp();
            named_arrive(3, TILE_THREADS);           // the workers pick the state up when they need it
        }
    } else {
        named_sync(1, TILE_THREADS);                 // the newline list is complete
        if (!declined) {
            // ---- 3a. output length of every piece, for each state the supertile may start in -----------
            //   FASTQ: piece i emits iff (state0 + i) % 4 == 1: field i % 4;  FASTA: only piece 0 (when it continues
            //   a line of the previous supertile) depends on the state: field 0 = sequence, field 1 = header
            const bool ls0 = rawb[-1] == '\n';
            for (uint32_t i = i_lo; i < i_hi; ++i) {
                const uint32_t start = i ? (uint32_t)s_nl[i - 1] + 1u : 0u;
                const bool has_nl = i < N;
                const uint32_t end = has_nl ? (uint32_t)s_nl[i] : blen;
                const uint32_t len = end - start;
                const uint32_t pb = rawb[(int)end - 1];
                const uint32_t klen = len - ((len > 0u && pb == '\r') ? 1u : 0u);
                // hash pieces of the line if it is a record's sequence (see PiecePlan): its valid k-mer end positions in
                // runs of <= pp.pmax; a line continued from the previous supertile has no dead prefix here
                if (MODE != MODE_FASTA && pp.pmax) {
                    const uint32_t nv = (i == 0u && !ls0) ? klen : (klen >= pp.k ? klen - (pp.k - 1u) : 0u);
                    const uint32_t np = ((nv + pp.pmax - 1u) * pp.pmax_inv) >> 22;
                    psum += (unsigned long long)np << (MODE == MODE_FASTQ ? 16u * (i & 3u) : 0u);
                }
                if (MODE == MODE_LINES) vsum += klen + (has_nl ? 1u : 0u);
                else if (MODE == MODE_FASTQ) {
                    vsum += (unsigned long long)(klen + (has_nl ? 1u : 0u)) << (16u * (i & 3u));
                    if (has_nl && i < 3u) seam_st->first[i] = (B0 + end) | (pb == '\r' ? 0x80000000u : 0u);   // for front_fix_kernel's seam check
                } else {
                    const unsigned long long both = 0x0000000000010001ULL;
                    if (i > 0u || ls0) vsum += ((start < blen && rawb[start] == '>') ? 1u : klen) * both;
                    else vsum += klen;
                }
            }
            if (MODE == MODE_FASTQ && tid == 0) {   // the supertile's last three newlines, for the seam checks
                const uint32_t n0 = N > 3u ? N - 3u : 0u;
                for (uint32_t q = n0; q < N; ++q) {
                    const uint32_t pos = s_nl[q];
                    seam_st->last[3u - (N - q)] = (B0 + pos) | (rawb[(int)pos - 1] == '\r' ? 0x80000000u : 0u);
                }
                seam_st->n = N;
            }
            // The workers do not wait for the look-back: they go on with a GUESS of the state and check it at the end
            // (wrong: 3b and 4 are redone).  FASTQ: a line that starts with '@' and whose second next line starts with
            // '+' is a header line in any well-formed file; FASTA: the supertile rarely starts inside a header line.
            if (MODE != MODE_LINES && tid == (int)PF_WT - 1) {
                uint32_t guess = MODE == MODE_FASTA ? 0u : 0xFFu;
                if (MODE == MODE_FASTQ) {
                    for (uint32_t i = ls0 ? 0u : 1u; i + 2u <= N && i < 12u; ++i) {
                        const uint32_t s_i = i ? (uint32_t)s_nl[i - 1] + 1u : 0u, s_i2 = (uint32_t)s_nl[i + 1] + 1u;
                        if (s_i2 < blen && rawb[s_i] == '@' && rawb[s_i2] == '+') { guess = (4u - (i & 3u)) & 3u; break; }
                    }
                }
                s_misc[M_GUESS] = guess;
            }
            // exclusive scan over the worker threads (output lengths and piece counts together)
            unsigned long long x = vsum, y = psum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long n = __shfl_up_sync(0xffffffffu, x, d), m = __shfl_up_sync(0xffffffffu, y, d);
                if (lane >= (uint32_t)d) { x += n; y += m; }
            }
            named_sync(2, PF_WT);                    // sh16l: the first scan's reads are over
            if (lane == 31u) { sh16l[wid] = x; sh16l[8 + wid] = y; }
            named_sync(2, PF_WT);
            unsigned long long base = 0, pbase = 0;
#pragma unroll
            for (uint32_t w = 0; w < PF_LB; ++w) {
                const unsigned long long sv = sh16l[w], pv = sh16l[8 + w];
                if (w < wid) { base += sv; pbase += pv; }
                vtot += sv; ptot += pv;
            }
            vex = base + x - vsum;
            pex = pbase + y - psum;
        }
    }
    uint32_t out_off = 0;
    int bases_delta = 0;
    uint32_t recs = 0, bad_rel = 0xFFFFFFFFu, lbad = 0xFFFFFFFFu;        // chunk-relative positions
    // the listed words of pass A go behind the newline list when a redo may need the newlines again
    const bool spec_room = N + 2u + (uint32_t)PF_BYTES / 16u <= (uint32_t)PF_LIST;
    uint16_t *s_bw = spec_room ? s_nl + N + 1u : s_nl;
    if (wid != PF_LB && MODE != MODE_LINES) {
        const uint32_t guess = (declined || !spec_room) ? 0xFFu : s_misc[M_GUESS];
        if (guess == 0xFFu) { named_sync(3, TILE_THREADS); state0 = s_misc[M_STATE]; have_state = true; }
        else state0 = guess;
    }
    while (wid != PF_LB && !declined) {
        // ---- 3b. pieces -> entries, with the state known (or guessed) ---------------------------------
        const uint32_t vsel = MODE == MODE_FASTQ ? 16u * ((1u - state0) & 3u) : (MODE == MODE_FASTA ? 16u * (state0 & 1u) : 0u);
        if (tid == 0) s_misc[M_T] = (uint32_t)(vtot >> vsel) & 0xFFFFu;
        // hash pieces: one per record run when they fit the region's table, else (and for FASTA) uniform ones below
        const uint32_t n_rec_pieces = (uint32_t)(ptot >> vsel) & 0xFFFFu;
        const bool rec_pieces = MODE != MODE_FASTA && pp.table != nullptr && pp.pmax != 0u && n_rec_pieces <= pp.stride;
        uint32_t pslot = (uint32_t)(pex >> vsel) & 0xFFFFu;
        uint32_t *ptab = pp.table ? pp.table + (size_t)st * pp.stride : nullptr;
        const bool ls0 = rawb[-1] == '\n';
        uint32_t sum = (uint32_t)(vex >> vsel) & 0xFFFFu;                // output offset of this thread's first piece
        for (uint32_t i = i_lo; i < i_hi; ++i) {
            const uint32_t start = i ? (uint32_t)s_nl[i - 1] + 1u : 0u;
            const bool has_nl = i < N;
            const uint32_t end = has_nl ? (uint32_t)s_nl[i] : blen;
            const uint32_t len = end - start;
            const bool line_start = i > 0u || ls0;
            const uint32_t pb = rawb[(int)end - 1];            // byte in front of the piece's end
            const uint32_t klen = len - ((len > 0u && pb == '\r') ? 1u : 0u);   // one trailing CR goes
            uint32_t L = 0, brk = 0;
            if (MODE == MODE_LINES) {
                L = klen; brk = has_nl ? 1u : 0u;
            } else if (MODE == MODE_FASTQ) {
                const uint32_t ph = (state0 + i) & 3u;
                if (line_start && start < blen) {
                    if (ph == 0u) recs++;
                    if ((ph == 0u || ph == 2u) && rawb[start] != (ph == 0u ? '@' : '+')) bad_rel = min(bad_rel, B0 + start);
                }
                if (ph == 1u) {
                    // sequence().len(): the line without its '\n' and without one CR right before it
                    bases_delta += (int)len - ((has_nl && pb == '\r') ? 1 : 0);
                    L = klen; brk = has_nl ? 1u : 0u;
                }
                // sequence / quality length check (see fastq_len_check) of a record whose four newlines are all in
                // this supertile (klen is the quality's length); the first three newlines: front_fix_kernel
                if (has_nl && i >= 3u && ph == 3u) {
                    const uint32_t e1 = s_nl[i - 2], e0 = s_nl[i - 3];
                    const uint32_t ls = e1 - e0 - 1u - (rawb[(int)e1 - 1] == '\r' ? 1u : 0u);
                    if (ls != klen) lbad = min(lbad, B0 + e0);
                }
            } else {  // MODE_FASTA
                const bool hdr = line_start ? (start < blen && rawb[start] == '>') : ((state0 & 1u) != 0u);
                if (line_start && hdr) {
                    recs++;
                    brk = 1u;                                  // the header start is a record break
                    bool prev_in_hdr;
                    if (i == 0u) prev_in_hdr = state0 != 0u;
                    else {
                        const uint32_t ps = i > 1u ? (uint32_t)s_nl[i - 2] + 1u : 0u;
                        prev_in_hdr = (i > 1u || ls0) ? rawb[ps] == '>' : ((state0 & 1u) != 0u);
                    }
                    // a new header ends the previous record: its raw sequence loses the final '\n'
                    // (and one CR before it) when that newline closed a sequence line
                    if (!prev_in_hdr) {
                        bases_delta -= 1;
                        if (rawb[(int)start - 2] == '\r') bases_delta -= 1;
                    }
                } else if (!hdr) {
                    bases_delta += (int)len + (has_nl ? 1 : 0);
                    L = klen;
                }
            }
            e_src[i] = (uint16_t)((start & 0x7FFFu) | (brk << 15));
            e_len[i] = (uint16_t)L;
            e_out[i] = (uint16_t)sum;
            if (MODE != MODE_FASTA && rec_pieces && (MODE == MODE_LINES || ((state0 + i) & 3u) == 1u)) {
                const uint32_t nv = (i == 0u && !ls0) ? L : (L >= pp.k ? L - (pp.k - 1u) : 0u);
                const uint32_t np = ((nv + pp.pmax - 1u) * pp.pmax_inv) >> 22;
                if (np) {   // nv positions in np runs whose lengths differ by at most one
                    const uint32_t q0 = nv / np, rem = nv - q0 * np;
                    uint32_t ppos = sum + L - nv;
                    for (uint32_t q = 0; q < np; ++q) {
                        const uint32_t pl = q0 + (q < rem ? 1u : 0u);
                        ptab[pslot + q] = ppos | (pl << 16);
                        ppos += pl;
                    }
                    pslot += np;
                }
            }
            const uint32_t outlen = L + brk;
            if (outlen) {   // words whose first byte this entry provides
                const uint32_t w_hi = (sum + outlen - 1u) >> 4;
                for (uint32_t w = (sum + 15u) >> 4; w <= w_hi; ++w) s_first[w] = (uint16_t)i;
            }
            sum += outlen;
        }
        named_sync(2, PF_WT);
        const uint32_t T = s_misc[M_T];
        // ---- 4. aligned output words ------------------------------------------------------------------
        // Pass A: every word that comes out of ONE line whole: unaligned 16-byte read, SIMD-in-register codes, one
        // aligned store.  The others are listed and done in pass B by as few warps as it takes.
        const uint32_t W = (T + 15u) >> 4;
        for (uint32_t w = tid; w < W; w += PF_WT) {
            const uint32_t i = s_first[w];
            const uint32_t d = 16u * w - (uint32_t)e_out[i];
            if (d + 16u <= (uint32_t)e_len[i]) {
                const uint32_t A = 16u + ((uint32_t)e_src[i] & 0x7FFFu) + d;      // pf_smem offset of output byte 0
                const uint32_t wi = A >> 2, sel = 0x3210u + 0x1111u * (A & 3u);
                uint32_t x[5], cw[4], diff = 0;
#pragma unroll
                for (int j = 0; j < 5; ++j) x[j] = raw32[wi + j];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t u = __byte_perm(x[j], x[j + 1], sel) & 0xDFDFDFDFu;
                    const uint32_t c2 = ((u >> 1) ^ (u >> 2)) & 0x03030303u;
                    cw[j] = c2;
                    uint32_t z = (c2 | (c2 >> 4)) & 0x00FF00FFu;
                    z = (z | (z >> 8)) & 0xFFFFu;
                    diff |= __byte_perm(0x54474341u, 0u, z) ^ u;
                }
                if (diff == 0u) {
                    *reinterpret_cast<uint4 *>(region + 16u * w) = make_uint4(cw[0], cw[1], cw[2], cw[3]);
                    continue;
                }
            }
            s_bw[atomicAdd(&s_misc[M_NBW], 1u)] = (uint16_t)w;
        }
        named_sync(2, PF_WT);
        // Pass B: the listed words, densely over the threads
        const uint32_t nbw = s_misc[M_NBW];
        for (uint32_t idx = tid; idx < nbw; idx += PF_WT) {
            const uint32_t w = s_bw[idx];
            uint32_t filled = 0;
            uint32_t pos = 16u * w;                            // output position within the region
            uint32_t i = s_first[w];
            uint32_t d = pos - (uint32_t)e_out[i];
            uint64_t alo = 0, ahi = 0;
            while (filled < 16u && pos < T) {
                const uint32_t len = e_len[i], brk = (uint32_t)e_src[i] >> 15;
                if (d < len) {
                    const uint32_t n = min(16u - filled, len - d);
                    const uint32_t A = 16u + ((uint32_t)e_src[i] & 0x7FFFu) + d - filled;   // pf_smem offset of output byte 0
                    const uint32_t wi = A >> 2, sel = 0x3210u + 0x1111u * (A & 3u);
                    uint32_t x[5], v[4];
#pragma unroll
                    for (int j = 0; j < 5; ++j) x[j] = raw32[wi + j];
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = __byte_perm(x[j], x[j + 1], sel);
                    uint64_t l0, h0, l1, h1;
                    ones128(filled, l0, h0);
                    ones128(filled + n, l1, h1);
                    const uint64_t mlo = l1 & ~l0, mhi = h1 & ~h0;
                    const uint32_t m[4] = {(uint32_t)mlo, (uint32_t)(mlo >> 32), (uint32_t)mhi, (uint32_t)(mhi >> 32)};
                    uint32_t cw[4], diff = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t u = v[j] & 0xDFDFDFDFu;
                        const uint32_t c2 = ((u >> 1) ^ (u >> 2)) & 0x03030303u;
                        cw[j] = c2;
                        uint32_t z = (c2 | (c2 >> 4)) & 0x00FF00FFu;
                        z = (z | (z >> 8)) & 0xFFFFu;
                        diff |= (__byte_perm(0x54474341u, 0u, z) ^ u) & m[j];
                    }
                    if (diff) {   // exact per-byte classes (needletail normalize(false))
                        bool blank = false;
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            const uint32_t k = lut[(v[q >> 2] >> (8 * (q & 3))) & 0xFFu];
                            const bool in = (uint32_t)q >= filled && (uint32_t)q < filled + n;
                            blank |= in && (k == CLS_WS || k == CLS_CR);
                            cw[q >> 2] = (cw[q >> 2] & ~(0xFFu << (8 * (q & 3)))) | ((k < 4u ? k : (uint32_t)SYM_BREAK) << (8 * (q & 3)));
                        }
                        if (blank) s_misc[M_DECL] = 1u;
                    }
                    alo |= ((uint64_t)cw[0] | ((uint64_t)cw[1] << 32)) & mlo;
                    ahi |= ((uint64_t)cw[2] | ((uint64_t)cw[3] << 32)) & mhi;
                    filled += n; pos += n; d += n;
                }
                if (filled < 16u && d == len && brk) {
                    if (filled < 8u) alo |= (uint64_t)SYM_BREAK << (8u * filled);
                    else ahi |= (uint64_t)SYM_BREAK << (8u * (filled - 8u));
                    filled++; pos++; d++;
                }
                if (d >= len + brk) { i += STEP; d = 0; }
            }
            if (filled < 16u) {   // the region's last word: the symbols past its end are breaks anyway (see below)
                uint64_t l0, h0;
                ones128(filled, l0, h0);
                alo |= 0x0404040404040404ULL & ~l0;
                ahi |= 0x0404040404040404ULL & ~h0;
            }
            *reinterpret_cast<uint4 *>(region + 16u * w) = make_uint4((uint32_t)alo, (uint32_t)(alo >> 32), (uint32_t)ahi, (uint32_t)(ahi >> 32));
        }
        out_off = T;
        if (pp.table) {
            uint32_t np_region = n_rec_pieces;
            if (!rec_pieces) {
                np_region = (T + 63u) >> 6;
                for (uint32_t j = tid; j < np_region; j += PF_WT) ptab[j] = (64u * j) | (min(64u, T - 64u * j) << 16);
            }
            if (tid == 0) s_misc[M_NP] = np_region;
        }
        if (have_state) break;
        // the guess against the look-back's answer
        named_sync(3, TILE_THREADS);
        have_state = true;
        if (s_misc[M_STATE] == state0) break;
        state0 = s_misc[M_STATE];
        bases_delta = 0; recs = 0; bad_rel = 0xFFFFFFFFu; lbad = 0xFFFFFFFFu;
        named_sync(2, PF_WT);                                  // everyone has read the pass-B list and the flags
        if (tid == 0) { s_misc[M_NBW] = 0; s_misc[M_DECL] = 0; }
        named_sync(2, PF_WT);
    }
    __syncthreads();
    if (s_misc[M_DECL]) declined = true;                       // uniform: read after the barrier

    if (declined) {   // anything the fast path does not do: the exact tile walk, from the supertile's start state
        __syncthreads();
        long long bd = 0;
        unsigned long long bp = ~0ULL;
        uint32_t nl_before = 0;
        out_off = 0; recs = 0; lbad = 0xFFFFFFFFu;
        const uint32_t t0 = st * g.st_tiles, t1 = min(t0 + g.st_tiles, g.n_tiles);
        pack_exact<MODE>(raw, g, t0, t1, raw_base, cprev1, cprev2, state0, region, lut, sh8, out_off, bd, recs, bp,
                         reinterpret_cast<uint16_t *>(pf_smem), c_nl, nl_before, lbad, seam_st);
        bases_delta = (int)bd;
        bad_rel = bp == ~0ULL ? 0xFFFFFFFFu : (uint32_t)(bp - raw_base);
        if (pp.table) {
            const uint32_t np_region = (out_off + 63u) >> 6;
            uint32_t *ptab = pp.table + (size_t)st * pp.stride;
            for (uint32_t j = tid; j < np_region; j += TILE_THREADS) ptab[j] = (64u * j) | (min(64u, out_off - 64u * j) << 16);
            if (tid == 0) s_misc[M_NP] = np_region;
        }
        if (MODE == MODE_FASTQ) {
            __syncthreads();
            if (tid == 0) { seam_st->last[0] = c_nl[0]; seam_st->last[1] = c_nl[1]; seam_st->last[2] = c_nl[2]; seam_st->n = nl_before; }
        }
    }

    // the hash kernel walks up to HASH_W positions past the region's end: make them breaks
    if (tid < HASH_W) region[out_off + tid] = SYM_BREAK;
    if (MODE != MODE_LINES) {
        const int bsum = __reduce_add_sync(0xffffffffu, bases_delta);
        const uint32_t rsum = __reduce_add_sync(0xffffffffu, recs);
        if (lane == 0) {
            if (bsum) atomicAdd(reinterpret_cast<int *>(&s_misc[M_BASES]), bsum);
            if (rsum) atomicAdd(&s_misc[M_RECS], rsum);
        }
        if (MODE == MODE_FASTQ) {
            const uint32_t bmin = __reduce_min_sync(0xffffffffu, bad_rel), lmin = __reduce_min_sync(0xffffffffu, lbad);
            if (lane == 0) {
                if (bmin != 0xFFFFFFFFu) atomicMin(&s_misc[M_BAD], bmin);
                if (lmin != 0xFFFFFFFFu) atomicMin(&s_misc[M_LBAD], lmin);
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        region_count[st] = out_off;
        atomicAdd(&carry->chunk_syms, out_off);
        atomicMax(&carry->max_region_syms, out_off);
        if (pp.table) { pp.count[st] = s_misc[M_NP]; atomicMax(&carry->max_region_pieces, s_misc[M_NP]); }
        if (MODE != MODE_LINES) {
            const long long bsum = (long long)(int)s_misc[M_BASES];
            if (bsum) atomicAdd((unsigned long long *)&carry->total_bases, (unsigned long long)bsum);
            if (s_misc[M_RECS]) atomicAdd((unsigned long long *)&carry->n_records, (unsigned long long)s_misc[M_RECS]);
            if (MODE == MODE_FASTQ) {
                if (s_misc[M_BAD] != 0xFFFFFFFFu) atomicMin((unsigned long long *)&carry->first_bad_pos, raw_base + s_misc[M_BAD]);
                if (s_misc[M_LBAD] != 0xFFFFFFFFu) atomicMin((unsigned long long *)&carry->len_bad_pos, raw_base + s_misc[M_LBAD]);
            }
        }
    }
}

// What phase_scan_kernel does to the carry besides the state: the stream's last two bytes and the chunk counters.
__global__ void chunk_begin_kernel(const uint8_t *__restrict__ raw, ChunkGeom g, ParseCarry *carry) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    carry->cprev1 = carry->prev1; carry->cprev2 = carry->prev2;
    if (g.len >= 2) { carry->prev2 = raw[g.len - 2]; carry->prev1 = raw[g.len - 1]; }
    else if (g.len == 1) { carry->prev2 = carry->prev1; carry->prev1 = raw[0]; }
    carry->chunk_syms = 0;
    carry->max_region_syms = 0;
    carry->max_region_pieces = 0;
    carry->chunk_raw_base = carry->raw_total;
    carry->raw_total += g.len;
}

// Uniform hash pieces (64 positions each) for the regions of a chunk parsed by the three-kernel pipeline.
__global__ void uniform_pieces_kernel(const uint32_t *__restrict__ region_count, uint32_t n_st, PiecePlan pp, ParseCarry *carry) {
    const uint32_t st = blockIdx.x;
    if (st >= n_st) return;
    const uint32_t T = region_count[st], np = (T + 63u) >> 6;
    uint32_t *ptab = pp.table + (size_t)st * pp.stride;
    for (uint32_t j = threadIdx.x; j < np; j += blockDim.x) ptab[j] = (64u * j) | (min(64u, T - 64u * j) << 16);
    if (threadIdx.x == 0) { pp.count[st] = np; atomicMax(&carry->max_region_pieces, np); }
}

// Front pads: warp r fills the `halo` bytes before region r (r < n_st) or the outgoing chunk tail
// (r == n_st) with the last `halo` symbols that precede it in stream order -- every lane fetches halo / 32 of
// them, walking back over short or empty regions and finally into the incoming chunk tail.  halo = 32 covers
// k <= 32; k up to 255 carries 256 symbols.
// FASTQ: warp r also runs the sequence / quality length check of the (up to three) newlines at the front of
// supertile r whose preceding newlines lie in earlier supertiles or chunks, and warp n_st leaves the stream's last
// three newlines in the carry for the next chunk.
__device__ __forceinline__ unsigned long long seam_stream_nl(uint32_t v, unsigned long long raw_base) {
    return (raw_base + (v & 0x7FFFFFFFu)) | ((unsigned long long)(v >> 31) << 63);
}
// the `want` (<= 3) newlines that precede supertile r's newline number j, most recent first
__device__ __forceinline__ int seam_collect(const SeamNl *seam, uint32_t r, int j, const ParseCarry *carry, int nl_in,
                                            unsigned long long raw_base, unsigned long long out[3]) {
    int got = 0;
    for (int q = j - 1; q >= 0 && got < 3; --q) out[got++] = seam_stream_nl(seam[r].first[q], raw_base);
    for (int t = (int)r - 1; t >= 0 && got < 3; --t) {
        const uint32_t n = min(seam[t].n, 3u);
        for (uint32_t q = 0; q < n && got < 3; ++q) out[got++] = seam_stream_nl(seam[t].last[2u - q], raw_base);
    }
    for (int q = 2; q >= 0 && got < 3; --q) {
        const unsigned long long v = carry->last_nl[nl_in][q];
        if (v == NL_NONE) return got;
        out[got++] = v;
    }
    return got;
}
__global__ void front_fix_kernel(uint8_t *__restrict__ sym, ChunkGeom g, const uint32_t *__restrict__ region_count,
                                 const uint8_t *__restrict__ tail_in, uint8_t *__restrict__ tail_out,
                                 const SeamNl *__restrict__ seam, const uint32_t *__restrict__ st_state, ParseCarry *carry,
                                 int nl_in, uint32_t halo, int commit_state) {
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (r > g.n_st) return;
    if (commit_state && r == g.n_st && lane == 0u) carry->state = carry->state_next;   // fused parse: see parse_fused_kernel
    if (seam != nullptr && lane < 3u) {
        const unsigned long long raw_base = carry->chunk_raw_base;
        const unsigned long long PM = ~(1ULL << 63);
        if (r < g.n_st) {
            const int j = (int)lane;
            if ((uint32_t)j < min(seam[r].n, 3u) && ((st_state[r] + (uint32_t)j) & 3u) == 3u) {
                unsigned long long e[3];
                if (seam_collect(seam, r, j, carry, nl_in, raw_base, e) == 3) {
                    const unsigned long long e3 = seam_stream_nl(seam[r].first[j], raw_base);
                    const unsigned long long ls = (e[1] & PM) - (e[2] & PM) - 1ULL - (e[1] >> 63);
                    const unsigned long long lq = (e3 & PM) - (e[0] & PM) - 1ULL - (e3 >> 63);
                    if (ls != lq) atomicMin((unsigned long long *)&carry->len_bad_pos, e[2] & PM);
                }
            }
        } else if (lane == 0u) {
            unsigned long long e[3];
            const int got = seam_collect(seam, g.n_st, 0, carry, nl_in, raw_base, e);
            for (int q = 0; q < 3; ++q) carry->last_nl[nl_in ^ 1][2 - q] = q < got ? e[q] : NL_NONE;
        }
    }
    uint8_t *dst = r < g.n_st ? (sym + (size_t)SYM_FRONT + (size_t)r * g.region_stride - halo) : tail_out;
    for (uint32_t i = lane; i < halo; i += 32u) {
        uint32_t back = halo - 1u - i;               // symbols between the wanted one and the end of the stream so far
        uint8_t v = 0;
        bool found = false;
        for (int j = (int)r - 1; j >= 0; --j) {
            const uint32_t n = region_count[j];
            if (back < n) { v = sym[(size_t)SYM_FRONT + (size_t)j * g.region_stride + (n - 1u - back)]; found = true; break; }
            back -= n;
        }
        if (!found) v = tail_in[halo - 1u - back];   // back < halo here: the incoming tail holds the halo symbols before the chunk
        dst[i] = v;
    }
}

__global__ void fill_bytes_kernel(uint8_t *p, uint32_t n, uint8_t v) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---- launchers ---------------------------------------------------------------------------------
void launch_phase(int mode, const uint8_t *raw, ChunkGeom g, ParseCarry *carry, uint32_t *st_map, uint32_t *st_state,
                  cudaStream_t s) {
    if (mode == MODE_FASTQ) phase_kernel<MODE_FASTQ><<<g.n_st, TILE_THREADS, 0, s>>>(raw, g, carry, st_map);
    else if (mode == MODE_FASTA) phase_kernel<MODE_FASTA><<<g.n_st, TILE_THREADS, 0, s>>>(raw, g, carry, st_map);
    phase_scan_kernel<<<1, 1024, 0, s>>>(st_map, g, carry, st_state, raw, mode == MODE_LINES ? 0 : 1);
}
void launch_pack(int mode, const uint8_t *raw, ChunkGeom g, ParseCarry *carry, const uint32_t *st_state, uint8_t *sym,
                 uint32_t *region_count, const uint8_t *tail_in, uint8_t *tail_out, SeamNl *seam, int nl_in, uint32_t halo,
                 cudaStream_t s) {
    if (mode == MODE_LINES) pack_kernel<MODE_LINES><<<g.n_st, TILE_THREADS, 0, s>>>(raw, g, carry, st_state, sym, region_count, nullptr);
    else if (mode == MODE_FASTA) pack_kernel<MODE_FASTA><<<g.n_st, TILE_THREADS, 0, s>>>(raw, g, carry, st_state, sym, region_count, nullptr);
    else pack_kernel<MODE_FASTQ><<<g.n_st, TILE_THREADS, 0, s>>>(raw, g, carry, st_state, sym, region_count, seam);
    front_fix_kernel<<<(g.n_st + 1 + 7) / 8, 256, 0, s>>>(sym, g, region_count, tail_in, tail_out,
                                                        mode == MODE_FASTQ ? seam : nullptr, st_state, carry, nl_in, halo, 0);
}
// The fused single-pass parse of a chunk (g.st_tiles <= PF_MAX_TILES): chunk_begin + parse_fused + front_fix.
// `status`: g.n_st words, zeroed when allocated; `epoch`: 1..65535, different from the previous 65534 launches over
// the same words; `ticket`: a device counter whose value before this launch is `ticket_base`.
int parse_fused_max_tiles() { return PF_MAX_TILES; }
template <int MODE>
static void launch_parse_fused_m(const uint8_t *raw, ChunkGeom g, ParseCarry *carry, uint32_t *st_state, uint8_t *sym,
                                 uint32_t *region_count, SeamNl *seam, uint32_t *status, uint32_t epoch, uint32_t *ticket,
                                 uint32_t ticket_base, PiecePlan pp, cudaStream_t s) {
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!attr_set[dev]) {
        cudaFuncSetAttribute(parse_fused_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PF_SMEM);
        attr_set[dev] = true;
    }
    parse_fused_kernel<MODE><<<g.n_st, TILE_THREADS, PF_SMEM, s>>>(raw, g, carry, st_state, sym, region_count, seam, status, epoch,
                                                                   ticket, ticket_base, pp);
}
void launch_parse_fused(int mode, const uint8_t *raw, ChunkGeom g, ParseCarry *carry, uint32_t *st_state, uint8_t *sym,
                        uint32_t *region_count, const uint8_t *tail_in, uint8_t *tail_out, SeamNl *seam, int nl_in, uint32_t halo,
                        uint32_t *status, uint32_t epoch, uint32_t *ticket, uint32_t ticket_base, PiecePlan pp, cudaStream_t s) {
    chunk_begin_kernel<<<1, 32, 0, s>>>(raw, g, carry);
    if (mode == MODE_LINES) launch_parse_fused_m<MODE_LINES>(raw, g, carry, st_state, sym, region_count, nullptr, status, epoch, ticket, ticket_base, pp, s);
    else if (mode == MODE_FASTA) launch_parse_fused_m<MODE_FASTA>(raw, g, carry, st_state, sym, region_count, nullptr, status, epoch, ticket, ticket_base, pp, s);
    else launch_parse_fused_m<MODE_FASTQ>(raw, g, carry, st_state, sym, region_count, seam, status, epoch, ticket, ticket_base, pp, s);
    front_fix_kernel<<<(g.n_st + 1 + 7) / 8, 256, 0, s>>>(sym, g, region_count, tail_in, tail_out,
                                                        mode == MODE_FASTQ ? seam : nullptr, st_state, carry, nl_in, halo,
                                                        mode == MODE_LINES ? 0 : 1);
}
void launch_uniform_pieces(const uint32_t *region_count, uint32_t n_st, PiecePlan pp, ParseCarry *carry, cudaStream_t s) {
    if (pp.table && n_st) uniform_pieces_kernel<<<n_st, 128, 0, s>>>(region_count, n_st, pp, carry);
}
void launch_fill_bytes(uint8_t *p, uint32_t n, uint8_t v, cudaStream_t s) {
    if (n) fill_bytes_kernel<<<(n + 255) / 256, 256, 0, s>>>(p, n, v);
}

}  // namespace fb2

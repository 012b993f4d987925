// parse.cu -- K1/K2: raw FASTA/FASTQ bytes in HBM -> dense symbol stream in HBM.
//
// Replaces, in bulk, needletail's record loop (lib/src/lib.rs:60-68) and the per-record
// `seq.normalize(false)` of SketchScheme::process (lib/src/sketch_schemes/mash.rs:72-73):
// every base that `process` would see becomes one byte 0..3 (A,C,G,T); every other kept
// position (N, IUPAC, gaps) and every record boundary becomes SYM_BREAK (4); whitespace and
// non-sequence lines vanish.  k-mers of the symbol stream == canonical_kmers of the records.
//
// Three kernels per chunk:
//   tile_summary_kernel : per 4 KiB tile, a transducer summary (next state + symbols emitted for
//                         every possible start state)
//   tile_scan_kernel    : exclusive scan of the summaries -> per-tile start state and output offset
//   pack_kernel         : re-reads the tile, writes symbols at their final offsets, accumulates
//                         total_bases / record count / FASTQ validity.
#include "common.cuh"
#include "device_types.cuh"

namespace fb2 {

struct Masks16 {
    uint32_t valid, nl, cr, ws, gt, at, plus;
    uint64_t codes;  // 4 bits per byte: min(class, 4)
};

__device__ __forceinline__ void load_classify(const uint8_t *__restrict__ raw, uint32_t off, uint32_t len,
                                              const uint8_t *lut, Masks16 &m) {
    uint32_t w[4] = {0, 0, 0, 0};
    int nvalid;
    if (off + 16u <= len) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(raw + off));
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        nvalid = 16;
    } else {
        nvalid = off < len ? (int)(len - off) : 0;
        for (int i = 0; i < nvalid; ++i) w[i >> 2] |= (uint32_t)raw[off + i] << (8 * (i & 3));
    }
    m.valid = nvalid >= 16 ? 0xFFFFu : ((1u << nvalid) - 1u);
    m.nl = m.cr = m.ws = m.gt = m.at = m.plus = 0;
    m.codes = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint32_t b = (w[i >> 2] >> (8 * (i & 3))) & 0xFFu;
        const uint32_t c = lut[b];
        m.nl |= (c == CLS_NL ? 1u : 0u) << i;
        m.cr |= (c == CLS_CR ? 1u : 0u) << i;
        m.ws |= (c == CLS_WS ? 1u : 0u) << i;
        m.gt |= (c == CLS_GT ? 1u : 0u) << i;
        m.at |= (c == CLS_AT ? 1u : 0u) << i;
        m.plus |= (c == CLS_PLUS ? 1u : 0u) << i;
        m.codes |= (uint64_t)(c < 4u ? c : 4u) << (4 * i);
    }
    m.nl &= m.valid; m.cr &= m.valid; m.ws &= m.valid; m.gt &= m.valid; m.at &= m.valid; m.plus &= m.valid;
}

// byte at chunk offset off-d (d = 1, 2); before the chunk: the carried stream bytes.
__device__ __forceinline__ uint32_t byte_before(const uint8_t *__restrict__ raw, uint32_t off, uint32_t d,
                                                uint32_t prev1, uint32_t prev2) {
    if (off >= d) return raw[off - d];
    const uint32_t back = d - off;  // 1 or 2 bytes before the chunk start
    return back == 1 ? prev1 : prev2;
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
__device__ __forceinline__ unsigned long long block_reduce_add64(unsigned long long v, unsigned long long *sh8) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh8[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long t = 0;
#pragma unroll
    for (int i = 0; i < TILE_THREADS / 32; ++i) t += sh8[i];
    return t;
}
__device__ __forceinline__ unsigned long long block_reduce_min64(unsigned long long v, unsigned long long *sh8) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = min(v, __shfl_down_sync(0xffffffffu, v, d));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh8[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long t = ~0ULL;
#pragma unroll
    for (int i = 0; i < TILE_THREADS / 32; ++i) t = min(t, sh8[i]);
    return t;
}
__device__ __forceinline__ unsigned long long block_reduce_max64(unsigned long long v, unsigned long long *sh8) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = max(v, __shfl_down_sync(0xffffffffu, v, d));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh8[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long t = 0;
#pragma unroll
    for (int i = 0; i < TILE_THREADS / 32; ++i) t = max(t, sh8[i]);
    return t;
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

// FASTA: header-byte mask (header lines including their '\n') for this thread's 16 bytes.
//   hs      : header starts ('>' at a line start) among the bytes
//   seed0   : byte 0 continues a header line begun earlier
// Carry-propagation flood: a seed at the bottom of a run of non-newline bytes floods the run and
// the newline that ends it.
__device__ __forceinline__ uint32_t fasta_header_mask(uint32_t hs, uint32_t nl, bool seed0) {
    const uint32_t g = hs | (seed0 ? 1u : 0u);
    const uint32_t p = ~nl & 0xFFFFu;
    return ((g + p) ^ p) & 0x1FFFFu;  // bit 16 = still in a header after byte 15
}

// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(TILE_THREADS)
tile_summary_kernel(const uint8_t *__restrict__ raw, uint32_t len, const ParseCarry *__restrict__ carry,
                    TileSummary *__restrict__ out) {
    __shared__ uint8_t lut[256];
    __shared__ uint32_t sh8[8];
    __shared__ unsigned long long sh8l[8];
    const int tid = threadIdx.x;
    lut[tid] = classify_byte((uint8_t)tid);
    __syncthreads();
    const uint32_t off = blockIdx.x * (uint32_t)TILE_BYTES + (uint32_t)tid * 16u;
    Masks16 m;
    load_classify(raw, off, len, lut, m);
    const uint32_t wsm = m.ws | m.cr;
    TileSummary s;
    s.next = 0; s.cnt[0] = s.cnt[1] = s.cnt[2] = s.cnt[3] = 0;

    if (MODE == MODE_LINES) {
        const unsigned long long t = block_reduce_add64(__popc(m.valid & ~wsm), sh8l);
        s.next = 0; s.cnt[0] = (uint32_t)t;
    } else if (MODE == MODE_FASTQ) {
        uint32_t total_nl;
        const uint32_t rel = block_exscan_add(__popc(m.nl), sh8, total_nl);
        // symbols per relative line class, 16 bits each
        unsigned long long acc = 0;
        uint32_t mm = m.nl, lo = 0, r = rel & 3u;
        while (true) {
            const int nb = mm ? (__ffs(mm) - 1) : 16;
            const uint32_t upto = nb >= 15 ? 0xFFFFu : ((2u << nb) - 1u);
            const uint32_t seg = upto & ~((1u << lo) - 1u);
            acc += (unsigned long long)__popc(seg & m.valid & ~wsm) << (16 * r);
            if (nb >= 15) break;
            mm &= mm - 1u;
            lo = (uint32_t)nb + 1u;
            r = (r + 1u) & 3u;
        }
        acc = block_reduce_add64(acc, sh8l);
        for (uint32_t h = 0; h < 4; ++h) {
            s.next |= ((h + total_nl) & 3u) << (2 * h);
            s.cnt[h] = (uint32_t)((acc >> (16 * ((1u - h) & 3u))) & 0xFFFFu);
        }
    } else {  // MODE_FASTA
        const uint32_t p1 = byte_before(raw, off, 1, carry->prev1, carry->prev2);
        const uint32_t ls = ((m.nl << 1) | (p1 == '\n' ? 1u : 0u)) & m.valid;
        const uint32_t hs = ls & m.gt;
        // "last line start so far" scan: value = (tid+1) << 1 | is_header
        uint32_t mine = 0;
        if (ls) {
            const int hi = 31 - __clz(ls);
            mine = ((uint32_t)(tid + 1) << 1) | ((m.gt >> hi) & 1u);
        }
        uint32_t last;
        const uint32_t prev = block_exscan_max(mine, sh8, last);
        const bool defined_in = prev != 0;
        const bool in_hdr = defined_in && (prev & 1u);
        const uint32_t hm = fasta_header_mask(hs, m.nl, in_hdr && !(ls & 1u));
        const uint32_t pre = (ls ? ((ls & (0u - ls)) - 1u) : 0xFFFFu) & m.valid;  // before first line start
        const uint32_t seqsym = ~hm & ~wsm & ~m.nl & m.valid;
        uint32_t post_cnt, pre_cnt;
        if (defined_in) { post_cnt = __popc(seqsym) + __popc(hs); pre_cnt = 0; }
        else { post_cnt = __popc(seqsym & ~pre) + __popc(hs); pre_cnt = __popc(seqsym & pre); }
        const unsigned long long t =
            block_reduce_add64(((unsigned long long)pre_cnt << 32) | post_cnt, sh8l);
        const uint32_t post = (uint32_t)t, pres = (uint32_t)(t >> 32);
        const bool has = last != 0;
        const uint32_t outst = last & 1u;
        // states: 0 = sequence line, 1 = header line
        s.next = (has ? outst : 0u) | ((has ? outst : 1u) << 2);
        s.cnt[0] = post + pres;
        s.cnt[1] = post;
    }
    if (tid == 0) out[blockIdx.x] = s;
}

// ------------------------------------------------------------------------------------------------
// Single block.  Thread t composes tiles [t*G, (t+1)*G), thread 0 chains the 1024 group summaries,
// then every thread replays its tiles with a known start state.
__global__ void __launch_bounds__(1024)
tile_scan_kernel(const TileSummary *__restrict__ sums, uint32_t n_tiles, ParseCarry *carry,
                 TilePrefix *__restrict__ pre, const uint8_t *__restrict__ raw, uint32_t len) {
    __shared__ uint32_t g_next[1024];
    __shared__ uint32_t g_cnt[4][1024];
    __shared__ uint32_t g_state[1024];
    __shared__ uint32_t g_off[1024];
    const uint32_t tid = threadIdx.x;
    const uint32_t G = (n_tiles + 1023u) / 1024u;
    const uint32_t t0 = min(tid * G, n_tiles), t1 = min(t0 + G, n_tiles);
    uint32_t nx[4] = {0, 1, 2, 3}, ct[4] = {0, 0, 0, 0};
    for (uint32_t t = t0; t < t1; ++t) {
        const TileSummary s = sums[t];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const uint32_t mid = nx[h];
            ct[h] += s.cnt[mid];
            nx[h] = (s.next >> (2 * mid)) & 3u;
        }
    }
    g_next[tid] = nx[0] | (nx[1] << 2) | (nx[2] << 4) | (nx[3] << 6);
#pragma unroll
    for (int h = 0; h < 4; ++h) g_cnt[h][tid] = ct[h];
    __syncthreads();
    if (tid == 0) {
        uint32_t st = carry->state & 3u, off = 0;
        for (uint32_t g = 0; g < 1024; ++g) {
            g_state[g] = st; g_off[g] = off;
            off += g_cnt[st][g];
            st = (g_next[g] >> (2 * st)) & 3u;
        }
        carry->cprev1 = carry->prev1; carry->cprev2 = carry->prev2;
        if (len >= 2) { carry->prev2 = raw[len - 2]; carry->prev1 = raw[len - 1]; }
        else if (len == 1) { carry->prev2 = carry->prev1; carry->prev1 = raw[0]; }
        carry->state = st;
        carry->chunk_syms = off;
        carry->chunk_raw_base = carry->raw_total;
        carry->raw_total += len;
        carry->chunk_ord_base = carry->ordinal;
        carry->ordinal += off;
    }
    __syncthreads();
    uint32_t st = g_state[tid], off = g_off[tid];
    for (uint32_t t = t0; t < t1; ++t) {
        const TileSummary s = sums[t];
        TilePrefix p; p.state_in = st; p.sym_off = off;
        pre[t] = p;
        off += s.cnt[st];
        st = (s.next >> (2 * st)) & 3u;
    }
}

// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(TILE_THREADS)
pack_kernel(const uint8_t *__restrict__ raw, uint32_t len, ParseCarry *carry,
            const TilePrefix *__restrict__ pre, uint8_t *__restrict__ sym) {
    __shared__ uint8_t lut[256];
    __shared__ uint32_t sh8[8];
    __shared__ unsigned long long sh8l[8];
    const int tid = threadIdx.x;
    lut[tid] = classify_byte((uint8_t)tid);
    __syncthreads();
    const uint32_t off = blockIdx.x * (uint32_t)TILE_BYTES + (uint32_t)tid * 16u;
    Masks16 m;
    load_classify(raw, off, len, lut, m);
    const uint32_t wsm = m.ws | m.cr;
    const TilePrefix tp = pre[blockIdx.x];
    const uint64_t raw_base = carry->chunk_raw_base;

    uint32_t em = 0;               // bytes that emit a symbol
    long long bases_delta = 0;     // contribution to total_bases
    uint32_t recs = 0;

    if (MODE == MODE_LINES) {
        em = m.valid & ~wsm;       // '\n' (class NL -> code 4) separates records
    } else if (MODE == MODE_FASTQ) {
        uint32_t total_nl;
        const uint32_t rel = block_exscan_add(__popc(m.nl), sh8, total_nl);
        const uint32_t ph0 = (tp.state_in + rel) & 3u;
        const uint32_t qm = fastq_seq_mask(m.nl, ph0) & m.valid;
        em = qm & ~wsm;
        const uint32_t p1 = byte_before(raw, off, 1, carry->cprev1, carry->cprev2);
        const uint32_t crprev = ((m.cr << 1) | (p1 == '\r' ? 1u : 0u)) & 0xFFFFu;
        // sequence().len(): the line without its '\n' and without one CR right before it
        bases_delta = (long long)__popc(qm & ~m.nl) - (long long)__popc(qm & m.nl & crprev);
        // line-start checks: phase 0 must start with '@', phase 2 with '+'
        uint32_t ls = ((m.nl << 1) | (p1 == '\n' ? 1u : 0u)) & m.valid;
        unsigned long long bad = ~0ULL;
        while (ls) {
            const int i = __ffs(ls) - 1;
            ls &= ls - 1u;
            const uint32_t ph = (ph0 + __popc(m.nl & ((1u << i) - 1u))) & 3u;
            if (ph == 0u) recs++;
            const bool ok = ph == 0u ? ((m.at >> i) & 1u) : (ph == 2u ? ((m.plus >> i) & 1u) : 1u);
            if (!ok) bad = min(bad, (unsigned long long)(raw_base + off + (uint32_t)i));
        }
        bad = block_reduce_min64(bad, sh8l);
        unsigned long long sig = 0;
        const uint32_t sg = m.valid & ~(m.nl | m.cr);
        if (sg) {
            const int i = 31 - __clz(sg);
            const uint32_t ph = (ph0 + __popc(m.nl & ((1u << i) - 1u))) & 3u;
            sig = ((unsigned long long)(raw_base + off + (uint32_t)i + 1u) << 2) | ph;
        }
        sig = block_reduce_max64(sig, sh8l);
        if (tid == 0) {
            if (bad != ~0ULL) atomicMin((unsigned long long *)&carry->first_bad_pos, bad);
            if (sig) atomicMax((unsigned long long *)&carry->last_sig, sig);
        }
    } else {  // MODE_FASTA
        const uint32_t p1 = byte_before(raw, off, 1, carry->cprev1, carry->cprev2);
        const uint32_t p2 = byte_before(raw, off, 2, carry->cprev1, carry->cprev2);
        const uint32_t ls = ((m.nl << 1) | (p1 == '\n' ? 1u : 0u)) & m.valid;
        const uint32_t hs = ls & m.gt;
        uint32_t mine = 0;
        if (ls) {
            const int hi = 31 - __clz(ls);
            mine = ((uint32_t)(tid + 1) << 1) | ((m.gt >> hi) & 1u);
        }
        uint32_t last;
        const uint32_t prev = block_exscan_max(mine, sh8, last);
        // state of the line containing the byte before this thread's first byte
        const uint32_t st_in = prev ? (prev & 1u) : (tp.state_in & 1u);
        const uint32_t hm = fasta_header_mask(hs, m.nl, st_in && !(ls & 1u));
        em = ((~hm & ~wsm & ~m.nl) | hs) & m.valid;
        recs = __popc(hs);
        bases_delta = __popc(~hm & m.valid);
        // a new header ends the previous record: its raw sequence loses the final '\n'
        // (and one CR before it) when that newline closed a sequence line
        uint32_t h = hs;
        const uint32_t crm = m.cr;
        while (h) {
            const int i = __ffs(h) - 1;
            h &= h - 1u;
            const bool prev_in_hdr = i >= 1 ? ((hm >> (i - 1)) & 1u) : (st_in != 0u);
            if (!prev_in_hdr) {
                bases_delta -= 1;
                const bool cr2 = i >= 2 ? ((crm >> (i - 2)) & 1u) : ((i == 1 ? p1 : p2) == '\r');
                if (cr2) bases_delta -= 1;
            }
        }
    }

    // ---- emit -----------------------------------------------------------------------------
    uint32_t tile_total;
    const uint32_t local = block_exscan_add(__popc(em), sh8, tile_total);
    uint8_t *o = sym + tp.sym_off + local;
    uint32_t e = em;
    while (e) {
        const int i = __ffs(e) - 1;
        e &= e - 1u;
        *o++ = (uint8_t)((m.codes >> (4 * i)) & 0xFu);
    }
    if (MODE != MODE_LINES) {
        const unsigned long long packed =
            block_reduce_add64((unsigned long long)bases_delta, sh8l);  // wrapping add is fine
        const unsigned long long r = block_reduce_add64(recs, sh8l);
        if (tid == 0) {
            if (packed) atomicAdd((unsigned long long *)&carry->total_bases, packed);
            if (r) atomicAdd((unsigned long long *)&carry->n_records, r);
        }
    }
}

// After the hash kernel has consumed a chunk: move the last SYM_FRONT symbols of
// (front pad + chunk) to the front pad for the next chunk.  One block of SYM_FRONT threads.
__global__ void carry_front_kernel(uint8_t *symbuf /* start of the front pad */, const ParseCarry *carry) {
    const uint32_t n = carry->chunk_syms;
    const uint8_t v = symbuf[n + threadIdx.x];
    __syncthreads();
    symbuf[threadIdx.x] = v;
}

__global__ void fill_bytes_kernel(uint8_t *p, uint32_t n, uint8_t v) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---- launchers ---------------------------------------------------------------------------------
void launch_tile_summary(int mode, const uint8_t *raw, uint32_t len, const ParseCarry *carry,
                         TileSummary *out, uint32_t n_tiles, cudaStream_t st) {
    if (mode == MODE_LINES) tile_summary_kernel<MODE_LINES><<<n_tiles, TILE_THREADS, 0, st>>>(raw, len, carry, out);
    else if (mode == MODE_FASTA) tile_summary_kernel<MODE_FASTA><<<n_tiles, TILE_THREADS, 0, st>>>(raw, len, carry, out);
    else tile_summary_kernel<MODE_FASTQ><<<n_tiles, TILE_THREADS, 0, st>>>(raw, len, carry, out);
}
void launch_tile_scan(const TileSummary *sums, uint32_t n_tiles, ParseCarry *carry, TilePrefix *pre,
                      const uint8_t *raw, uint32_t len, cudaStream_t st) {
    tile_scan_kernel<<<1, 1024, 0, st>>>(sums, n_tiles, carry, pre, raw, len);
}
void launch_pack(int mode, const uint8_t *raw, uint32_t len, ParseCarry *carry, const TilePrefix *pre,
                 uint8_t *sym, uint32_t n_tiles, cudaStream_t st) {
    if (mode == MODE_LINES) pack_kernel<MODE_LINES><<<n_tiles, TILE_THREADS, 0, st>>>(raw, len, carry, pre, sym);
    else if (mode == MODE_FASTA) pack_kernel<MODE_FASTA><<<n_tiles, TILE_THREADS, 0, st>>>(raw, len, carry, pre, sym);
    else pack_kernel<MODE_FASTQ><<<n_tiles, TILE_THREADS, 0, st>>>(raw, len, carry, pre, sym);
}
void launch_carry_front(uint8_t *symbuf, const ParseCarry *carry, cudaStream_t st) {
    carry_front_kernel<<<1, SYM_FRONT, 0, st>>>(symbuf, carry);
}
void launch_fill_bytes(uint8_t *p, uint32_t n, uint8_t v, cudaStream_t st) {
    if (n) fill_bytes_kernel<<<(n + 255) / 256, 256, 0, st>>>(p, n, v);
}

}  // namespace fb2

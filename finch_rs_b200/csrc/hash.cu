// hash.cu -- K3: symbol stream -> canonical k-mer -> murmur3_x64_128 h1 -> threshold -> candidate log.
//
// Replaces the per-k-mer hot loop of the reference:
//   for (_, kmer, is_rc) in norm_seq.canonical_kmers(k, &rc) { self.push(kmer, is_rc as u8) }
//   (lib/src/sketch_schemes/mash.rs:76-79, scaled.rs:74-77) up to and including the admission
//   test of `push` (mash.rs:36-42 / scaled.rs:41), for all positions of a chunk at once.
// Each thread owns HASH_W consecutive k-mer END positions, warms its rolling 2-bit forward and
// reverse-complement words on the k-1 symbols before them, and for every valid window picks the
// canonical strand by integer compare, expands it to the ASCII bytes the reference hashes,
// computes h1 and appends (hash, k-mer codes, position|strand) to the log when h1 <= threshold.
// Order-independence of the sketch makes the unordered log exact.
#include "common.cuh"
#include "device_types.cuh"

namespace fb2 {

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ---- 32-bit-half arithmetic, placed on the pipes by hand ----------------------------------------
// The kernel is bound by the ALU pipe (LOP3/SHF/IADD3/ISETP/SEL/PRMT, 16 lanes per SM sub-partition);
// IMAD runs on the FMA pipe, which has room.  Everything that can be phrased as a multiply-add is:
// 64-bit adds (IMAD.WIDE + IMAD), *5+c, shared-memory addressing (idx * stride + base with the stride
// in a register so ptxas cannot turn it back into LEA), ">> 1" of a high word (IMAD.HI by 2^31).
// 64-bit rotates are two funnel shifts.
struct U2 { uint32_t lo, hi; };

__device__ __forceinline__ uint32_t mad_lo(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ U2 mad_wide(uint32_t a, uint32_t b, U2 c) {   // a * b + c (64-bit)
    U2 d;
    asm("{\n.reg .b64 t, u;\nmov.b64 u, {%4, %5};\nmad.wide.u32 t, %2, %3, u;\nmov.b64 {%0, %1}, t;\n}"
        : "=r"(d.lo), "=r"(d.hi) : "r"(a), "r"(b), "r"(c.lo), "r"(c.hi));
    return d;
}
__device__ __forceinline__ U2 mul_wide(uint32_t a, uint32_t b) {
    U2 d;
    asm("{\n.reg .b64 t;\nmul.wide.u32 t, %2, %3;\nmov.b64 {%0, %1}, t;\n}" : "=r"(d.lo), "=r"(d.hi) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t shr1_fma(uint32_t x) {               // x >> 1 on the FMA pipe
    uint32_t d;
    asm("mul.hi.u32 %0, %1, 0x80000000;" : "=r"(d) : "r"(x));
    return d;
}
__device__ __forceinline__ U2 add_u2(U2 a, U2 b) {                        // IMAD.WIDE + IMAD
    U2 t = mad_wide(a.lo, 1u, b);
    t.hi = mad_lo(a.hi, 1u, t.hi);
    return t;
}
template <uint64_t C>
__device__ __forceinline__ U2 mul_u2(U2 x) {                              // x * C mod 2^64: 3 IMAD
    U2 w = mul_wide(x.lo, (uint32_t)C);
    w.hi = mad_lo(x.lo, (uint32_t)(C >> 32), w.hi);
    w.hi = mad_lo(x.hi, (uint32_t)C, w.hi);
    return w;
}
template <uint32_t ADD>
__device__ __forceinline__ U2 mul5add_u2(U2 x) {                          // x * 5 + ADD
    U2 c; c.lo = ADD; c.hi = 0;
    U2 w = mad_wide(x.lo, 5u, c);
    w.hi = mad_lo(x.hi, 5u, w.hi);
    return w;
}
template <int R>
__device__ __forceinline__ U2 rotl_u2(U2 x) {                             // 2 SHF
    U2 d;
    if (R == 32) { d.lo = x.hi; d.hi = x.lo; }
    else if (R < 32) { d.hi = __funnelshift_l(x.lo, x.hi, R); d.lo = __funnelshift_l(x.hi, x.lo, R); }
    else { d.hi = __funnelshift_l(x.hi, x.lo, R - 32); d.lo = __funnelshift_l(x.lo, x.hi, R - 32); }
    return d;
}
__device__ __forceinline__ U2 fmix_u2(U2 k) {
    k.lo ^= shr1_fma(k.hi);  k = mul_u2<0xff51afd7ed558ccdULL>(k);       // k ^= k >> 33 touches the low word only
    k.lo ^= shr1_fma(k.hi);  k = mul_u2<0xc4ceb9fe1a85ec53ULL>(k);
    k.lo ^= shr1_fma(k.hi);
    return k;
}
__device__ __forceinline__ uint32_t byte_of(uint32_t x, int n) { return __byte_perm(x, 0u, 0x4440u + (uint32_t)n); }

// ---- murmur3 with the first multiply taken from shared-memory tables ---------------------------
// Every 8-base word w of the k-mer enters murmur3 as  w * c  (c = c1 for k1-type words, c2 for
// k2-type words), w being the 8 ASCII bytes of the bases.  Multiplication distributes over the
// byte groups:  w * c = A4(lo) * c + (A4(hi) * c << 32)  with A4(b) the 4 ASCII bytes of the 4
// bases b.  lut_c[b] = A4(b) * c (64 bit) turns "expand 2-bit codes to ASCII, then multiply" into
// two shared-memory loads and one add.  For k = 21 the 5-byte tail (10 bits of codes) has its own
// 1024-entry table, so the whole tail word is one load.
struct MulLut { uint32_t c1, c2, t5; uint32_t stride; };   // shared-window byte addresses; stride == 8 (in a register)

__device__ __forceinline__ uint2 lut_load(uint32_t table, uint32_t idx, uint32_t stride) {
    const uint32_t addr = mad_lo(idx, stride, table);
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lut_load_lo(uint32_t table, uint32_t idx, uint32_t stride) {
    const uint32_t addr = mad_lo(idx, stride, table);
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// (8 bases given as two table indices) * c
__device__ __forceinline__ U2 mul_word8(uint32_t table, uint32_t i0, uint32_t i1, uint32_t stride) {
    const uint2 a = lut_load(table, i0, stride);
    const uint32_t b = lut_load_lo(table, i1, stride);
    U2 r; r.lo = a.x; r.hi = mad_lo(b, 1u, a.y);
    return r;
}
// NBYTES (1..7) bases in g16 (base 0 lowest, the rest zero bytes) times C
template <int NBYTES, bool IS_C1>
__device__ __forceinline__ U2 mul_word_part(uint32_t g16, const MulLut &L) {
    const uint32_t T = IS_C1 ? L.c1 : L.c2;
    const uint64_t C = IS_C1 ? MM_C1 : MM_C2;
    U2 r;
    if (NBYTES > 4) {
        const uint2 a = lut_load(T, g16 & 0xFFu, L.stride);
        const uint32_t hi4 = expand4(g16 >> 8) & (uint32_t)low_bytes_mask(NBYTES - 4);
        r.lo = a.x; r.hi = mad_lo(hi4, (uint32_t)C, a.y);
    } else if (NBYTES == 4) {
        const uint2 a = lut_load(T, g16 & 0xFFu, L.stride);
        r.lo = a.x; r.hi = a.y;
    } else {
        const uint32_t lo4 = expand4(g16 & 0xFFu) & (uint32_t)low_bytes_mask(NBYTES);
        r = mul_wide(lo4, (uint32_t)C);
        r.hi = mad_lo(lo4, (uint32_t)(C >> 32), r.hi);
    }
    return r;
}

template <int K, bool SEED0>
__device__ __forceinline__ U2 murmur_kmer_h1_lut(U2 codes, U2 seed, const MulLut &L) {
    static_assert(K >= 1 && K <= 32, "k out of range");
    constexpr int NB = K / 16, T = K & 15;
    U2 h1 = seed, h2 = seed;
    if (NB >= 1) {
        U2 k1 = mul_word8(L.c1, byte_of(codes.lo, 0), byte_of(codes.lo, 1), L.stride);
        U2 k2 = mul_word8(L.c2, byte_of(codes.lo, 2), byte_of(codes.lo, 3), L.stride);
        k1 = rotl_u2<31>(k1); k1 = mul_u2<MM_C2>(k1);
        if (SEED0) { h1 = k1; } else { h1.lo ^= k1.lo; h1.hi ^= k1.hi; }
        h1 = rotl_u2<27>(h1); if (!SEED0) h1 = add_u2(h1, h2); h1 = mul5add_u2<0x52dce729u>(h1);
        k2 = rotl_u2<33>(k2); k2 = mul_u2<MM_C1>(k2);
        if (SEED0) { h2 = k2; } else { h2.lo ^= k2.lo; h2.hi ^= k2.hi; }
        h2 = rotl_u2<31>(h2); h2 = add_u2(h2, h1); h2 = mul5add_u2<0x38495ab5u>(h2);
    }
    if (NB >= 2) {
        U2 k1 = mul_word8(L.c1, byte_of(codes.hi, 0), byte_of(codes.hi, 1), L.stride);
        U2 k2 = mul_word8(L.c2, byte_of(codes.hi, 2), byte_of(codes.hi, 3), L.stride);
        k1 = rotl_u2<31>(k1); k1 = mul_u2<MM_C2>(k1); h1.lo ^= k1.lo; h1.hi ^= k1.hi;
        h1 = rotl_u2<27>(h1); h1 = add_u2(h1, h2); h1 = mul5add_u2<0x52dce729u>(h1);
        k2 = rotl_u2<33>(k2); k2 = mul_u2<MM_C1>(k2); h2.lo ^= k2.lo; h2.hi ^= k2.hi;
        h2 = rotl_u2<31>(h2); h2 = add_u2(h2, h1); h2 = mul5add_u2<0x38495ab5u>(h2);
    }
    // tail words: 16-bit groups TW (k1-type) and TW + 1 (k2-type) of the codes
    const uint32_t tword = (NB == 0) ? codes.lo : codes.hi;     // NB == 2 has no tail
    if (T > 8) {
        U2 k2;
        if (T == 16) k2 = mul_word8(L.c2, byte_of(tword, 2), byte_of(tword, 3), L.stride);
        else k2 = mul_word_part<(T > 8 ? T - 8 : 1), false>(tword >> 16, L);
        k2 = rotl_u2<33>(k2); k2 = mul_u2<MM_C1>(k2); h2.lo ^= k2.lo; h2.hi ^= k2.hi;
    }
    if (T > 0) {
        U2 k1;
        if (T >= 8) k1 = mul_word8(L.c1, byte_of(tword, 0), byte_of(tword, 1), L.stride);
        else if (K == 21) {                    // 5 bases = the whole (masked) high word of the codes
            const uint2 a = lut_load(L.t5, tword, L.stride);
            k1.lo = a.x; k1.hi = a.y;
        } else k1 = mul_word_part<(T > 0 && T < 8 ? T : 1), true>(tword & 0xFFFFu, L);
        k1 = rotl_u2<31>(k1); k1 = mul_u2<MM_C2>(k1); h1.lo ^= k1.lo; h1.hi ^= k1.hi;
    }
    h1.lo ^= (uint32_t)K; h2.lo ^= (uint32_t)K;
    h1 = add_u2(h1, h2); h2 = add_u2(h2, h1);
    h1 = fmix_u2(h1); h2 = fmix_u2(h2);
    return add_u2(h1, h2);
}

constexpr uint32_t LOG_RESERVE = 3;   // extra log slots a warp reserves per atomic (<= 31)

// ---- TMA (1-D bulk copy) staging of the block's symbol tile ------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy by the TMA unit; completion is signalled on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// ---- rolling canonical k-mer state, LSB-first on both strands ------------------------------------
//   A = LSB-first codes of the forward window (base i of the window at bits [2i, 2i+1])
//   B = LSB-first codes of its reverse complement
// With M(x) the MSB-first integer whose order is the byte order of the ASCII strings (common.cuh):
// M(fwd) = ~B & mask and M(rc) = ~A & mask, so  fwd < rc  <=>  B > A, and needletail's choice
// (fwd < rc ? fwd : rc, a palindrome reporting rc) is  is_rc = (A >= B), codes = is_rc ? B : A.
template <int K>
struct Roll2 {
    U2 A, B;
    __device__ __forceinline__ void push(uint32_t s /* symbol in the low byte, other bits arbitrary */, int k_rt,
                                         uint64_t mask_rt) {
        if (K > 0) {
            constexpr int TOP = 2 * (K - 1);                 // bit position where a new base enters A
            constexpr uint64_t MASK = K >= 32 ? ~0ULL : ((1ULL << (2 * (K > 0 ? K : 1))) - 1ULL);
            // A = (A >> 2) | (c << TOP)
            const uint32_t alo = __funnelshift_r(A.lo, A.hi, 2);
            if (TOP >= 32) {
                A.hi = (A.hi >> 2) | ((s & 3u) << (TOP - 32));
                A.lo = alo;
            } else {
                A.hi = 0;
                A.lo = alo | ((s & 3u) << TOP);
            }
            // B = ((B << 2) | (3 - c)) & mask
            const uint32_t bhi = __funnelshift_l(B.lo, B.hi, 2);
            B.lo = mad_lo(B.lo, 4u, (~s) & 3u) & (uint32_t)MASK;
            B.hi = bhi & (uint32_t)(MASK >> 32);
        } else {
            const uint64_t c = s & 3u;
            uint64_t a = ((uint64_t)A.hi << 32) | A.lo, b = ((uint64_t)B.hi << 32) | B.lo;
            a = (a >> 2) | (c << (2 * (k_rt - 1)));
            b = ((b << 2) | (c ^ 3ULL)) & mask_rt;
            A.lo = (uint32_t)a; A.hi = (uint32_t)(a >> 32); B.lo = (uint32_t)b; B.hi = (uint32_t)(b >> 32);
        }
    }
};

template <int K, bool SEED0>
__global__ void __launch_bounds__(HASH_THREADS)
hash_kernel(const uint8_t *__restrict__ symbuf,   // symbol buffer: region r starts at SYM_FRONT + r * region_stride
            ChunkGeom g, uint32_t b0,             // first hash block (region-major) of this launch
            const uint32_t *__restrict__ region_count, uint64_t ord_base, const SketchState *st,
            LaunchSlot *slot, LogView log, int k_rt, uint64_t seed, uint32_t lut_stride /* == 8 */) {
    __shared__ uint2 lut_c1[256], lut_c2[256];
    __shared__ uint2 lut_t5[K == 21 ? 1024 : 1];
    if (K > 0) {
        if (threadIdx.x < 256) {
            const uint32_t a4 = expand4(threadIdx.x & 0xFFu);
            const uint64_t p1 = (uint64_t)a4 * MM_C1, p2 = (uint64_t)a4 * MM_C2;
            lut_c1[threadIdx.x] = make_uint2((uint32_t)p1, (uint32_t)(p1 >> 32));
            lut_c2[threadIdx.x] = make_uint2((uint32_t)p2, (uint32_t)(p2 >> 32));
        }
        if (K == 21) {
            for (uint32_t i = threadIdx.x; i < 1024u; i += HASH_THREADS) {
                const uint64_t w = (uint64_t)expand4(i & 0xFFu) | ((uint64_t)(expand4(i >> 8) & 0xFFu) << 32);
                const uint64_t p = w * MM_C1;
                lut_t5[i] = make_uint2((uint32_t)p, (uint32_t)(p >> 32));
            }
        }
        __syncthreads();
    }
    MulLut L; L.c1 = smem_u32(lut_c1); L.c2 = smem_u32(lut_c2); L.t5 = smem_u32(lut_t5); L.stride = lut_stride;
    __shared__ __align__(128) uint8_t tile[32 + HASH_TILE];   // 32 symbols of halo, then the block's positions
    __shared__ __align__(8) uint64_t tile_bar;
    const int k = K > 0 ? K : k_rt;
    const uint64_t mask = kmer_mask(k);
    const uint32_t blk = b0 + blockIdx.x;
    const uint32_t region = blk / g.hash_tiles, lt = blk - region * g.hash_tiles;
    const uint32_t end = region_count[region];
    const uint32_t pb = lt * HASH_TILE;          // first position of this block in its region
    if (pb >= end) return;                        // block-uniform: nothing to do
    const uint8_t *sym = symbuf + (size_t)SYM_FRONT + (size_t)region * g.region_stride;
    const uint64_t ord_region = ord_base + (uint64_t)region * g.st_bytes;
    const unsigned long long T = st->threshold;
    const uint32_t T_hi = (uint32_t)(T >> 32);
    const uint32_t lane = threadIdx.x & 31u;
    U2 seed2; seed2.lo = (uint32_t)seed; seed2.hi = (uint32_t)(seed >> 32);
    // ---- stage [pb - 32, min(end + HASH_W, pb + HASH_TILE)) with one TMA bulk copy --------------
    // (positions in [end, end + HASH_W) hold SYM_BREAK, written by pack_kernel)
    {
        const uint32_t npos = min(end + (uint32_t)HASH_W - pb, (uint32_t)HASH_TILE);
        const uint32_t bytes = (32u + npos + 15u) & ~15u;
        if (threadIdx.x == 0) { mbar_init(&tile_bar, 1); fence_mbar_init(); }
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_arrive_expect_tx(&tile_bar, bytes);
            tma_load_1d(tile, sym + pb - 32, bytes, &tile_bar);
        }
        mbar_wait(&tile_bar, 0);
    }
    const uint32_t t0 = threadIdx.x * (uint32_t)HASH_W;   // this thread's first position within the tile
    const uint32_t p0 = pb + t0;
    // Warp-uniform early exit: a warp's positions are contiguous and ascending.
    if (__all_sync(0xffffffffu, p0 >= end)) return;
    // lanes past the end of the region (in a warp that is not entirely past it) never read: they
    // walk SYM_BREAK words.  Live lanes stay inside [p0 - 32, p0 + HASH_W), which was staged.
    const bool live = p0 < end;

    Roll2<K> r; r.A.lo = r.A.hi = r.B.lo = r.B.hi = 0;
    // brk = position (relative to the current group of 4) of the last non-base symbol; the window
    // ending at relative position b is valid  <=>  brk <= b - k.
    int brk = -k;
    const uint32_t *wp = reinterpret_cast<const uint32_t *>(tile + 32 + t0);
    // ---- warm-up on the 32 symbols before p0 (only the last k-1 matter) -------------------
    if (live) {
        const uint4 a = *reinterpret_cast<const uint4 *>(tile + t0);
        const uint4 b = *reinterpret_cast<const uint4 *>(tile + t0 + 16);
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            if (K > 0 && i < 32 - (K - 1)) continue;  // compile-time skip
            const uint32_t s = w[i >> 2] >> (8 * (i & 3));
            if (K == 0 && i < 32 - (k - 1)) continue;
            r.push(s, k, mask);
            if ((s & 0xFFu) >= 4u) brk = i - 32;
        }
    } else {
        brk = -1;
    }
    uint32_t nvalid = 0;
    uint32_t res_base = 0, res_left = 0;   // warp-uniform: this warp's reserved slice of the log
    uint32_t word = live ? wp[0] : 0x04040404u;
#pragma unroll 1
    for (int j = 0; j < HASH_W / 4; ++j) {
        const uint32_t cur = word;
        if (j + 1 < HASH_W / 4) word = live ? wp[j + 1] : 0x04040404u;  // next 4 symbols
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            r.push(cur >> (8 * b), k, mask);
            if (cur & (0xFCu << (8 * b))) brk = b;
            const bool ok = brk <= b - k;
            const bool is_rc = (((uint64_t)r.A.hi << 32) | r.A.lo) >= (((uint64_t)r.B.hi << 32) | r.B.lo);
            U2 codes; codes.lo = is_rc ? r.B.lo : r.A.lo; codes.hi = is_rc ? r.B.hi : r.A.hi;
            U2 h;
            if (K > 0) h = murmur_kmer_h1_lut<(K > 0 ? K : 1), SEED0>(codes, seed2, L);
            else {
                const uint64_t hv = murmur_kmer_h1<0>(((uint64_t)codes.hi << 32) | codes.lo, k, seed);
                h.lo = (uint32_t)hv; h.hi = (uint32_t)(hv >> 32);
            }
            nvalid += ok ? 1u : 0u;
            // hot path: compare only the high words (conservative); the exact test is in the branch
            const bool maybe = ok && (h.hi <= T_hi);
            if (__any_sync(0xffffffffu, maybe)) {
              const unsigned long long hv = ((unsigned long long)h.hi << 32) | h.lo;
              const bool emit = maybe && (hv <= T);
              const uint32_t em = __ballot_sync(0xffffffffu, emit);
              if (em) {
                // Warp-private bump reservation in the log: the global atomic (and the wait for its
                // result) happens once per LOG_RESERVE candidates, not once per candidate.
                const uint32_t n = __popc(em);
                if (n > res_left) {                                   // warp-uniform
                    if (lane < res_left && res_base + lane < log.cap) log.posx[res_base + lane] = ~0ULL;  // unused slots
                    const uint32_t want = n + LOG_RESERVE;
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(&slot->log_count, want);
                    res_base = __shfl_sync(0xffffffffu, base, 0);
                    res_left = want;
                }
                if (emit) {
                    const uint32_t idx = res_base + __popc(em & lanemask_lt());
                    if (idx < log.cap) {
                        const uint32_t p = p0 + 4u * (uint32_t)j + (uint32_t)b;
                        log.hash[idx] = hv;
                        log.kmer[idx] = ((unsigned long long)codes.hi << 32) | codes.lo;
                        log.posx[idx] = ((ord_region + p) << 9) | (is_rc ? 1ull : 0ull);
                    }
                }
                res_base += n; res_left -= n;
              }
            }
        }
        brk -= 4;
    }
    if (lane < res_left && res_base + lane < log.cap) log.posx[res_base + lane] = ~0ULL;  // unused tail of the reservation
    // valid-window count of this launch (committed to total_kmers by the host on success)
    nvalid = __reduce_add_sync(0xffffffffu, nvalid);
    if (lane == 0 && nvalid) atomicAdd(&slot->launch_kmers, (unsigned long long)nvalid);
}

// `push` unit-test surface (mash.rs:34 / scaled.rs:37): hash arbitrary byte strings.
__global__ void push_hash_kernel(const uint8_t *__restrict__ bytes, const uint32_t *__restrict__ offs,
                                 const uint8_t *__restrict__ extra, uint32_t n, uint64_t arena_base,
                                 uint64_t ord_base, const SketchState *st, LaunchSlot *slot, LogView log, uint64_t seed) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t a = offs[i], b = offs[i + 1];
    const uint64_t h = murmur_bytes_h1(bytes + a, b - a, seed);
    const unsigned long long T = st->threshold;
    if (h <= T) {
        const uint32_t idx = atomicAdd(&slot->log_count, 1u);
        if (idx < log.cap) {
            log.hash[idx] = h;
            log.kmer[idx] = arena_base + i;
            log.posx[idx] = ((ord_base + i) << 9) | (1ull << 8) | (unsigned long long)extra[i];
        }
    }
}
__global__ void push_commit_kernel(LaunchSlot *slot, uint32_t n) { slot->launch_kmers += n; }

template <int K>
static void launch_hash_k(uint32_t blocks, const uint8_t *symbuf, ChunkGeom g, uint32_t b0, const uint32_t *region_count,
                          uint64_t ord_base, const SketchState *st, LaunchSlot *slot, LogView log, int k, uint64_t seed,
                          cudaStream_t stream) {
    if (seed == 0 && K > 0)
        hash_kernel<K, true><<<blocks, HASH_THREADS, 0, stream>>>(symbuf, g, b0, region_count, ord_base, st, slot, log, k, seed, 8u);
    else
        hash_kernel<K, false><<<blocks, HASH_THREADS, 0, stream>>>(symbuf, g, b0, region_count, ord_base, st, slot, log, k, seed, 8u);
}
void launch_hash(int k, const uint8_t *symbuf, ChunkGeom g, uint32_t b0, uint32_t b1, const uint32_t *region_count,
                 uint64_t ord_base, const SketchState *st, LaunchSlot *slot, LogView log, uint64_t seed,
                 cudaStream_t stream) {
    if (b1 <= b0) return;
    const uint32_t blocks = b1 - b0;
    if (k == 21) launch_hash_k<21>(blocks, symbuf, g, b0, region_count, ord_base, st, slot, log, k, seed, stream);
    else if (k == 31) launch_hash_k<31>(blocks, symbuf, g, b0, region_count, ord_base, st, slot, log, k, seed, stream);
    else launch_hash_k<0>(blocks, symbuf, g, b0, region_count, ord_base, st, slot, log, k, seed, stream);
}
void launch_push_hash(const uint8_t *bytes, const uint32_t *offs, const uint8_t *extra, uint32_t n,
                      uint64_t arena_base, uint64_t ord_base, const SketchState *st, LaunchSlot *slot, LogView log,
                      uint64_t seed, cudaStream_t stream) {
    if (!n) return;
    push_hash_kernel<<<(n + 255) / 256, 256, 0, stream>>>(bytes, offs, extra, n, arena_base, ord_base, st, slot, log, seed);
    push_commit_kernel<<<1, 1, 0, stream>>>(slot, n);
}

}  // namespace fb2

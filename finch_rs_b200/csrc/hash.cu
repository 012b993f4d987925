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
#include <algorithm>
#include <cstdlib>
#include "common.cuh"
#include "device_types.cuh"

#ifndef FMIX_HI_A
#define FMIX_HI_A 1
#endif
#ifndef FMIX_HI_B
#define FMIX_HI_B 1
#endif
#ifndef HASH_VAR_DEFAULT
#define HASH_VAR_DEFAULT 0      // see murmur_kmer_h1_lut: which arithmetic form each step takes
#endif

namespace fb2 {

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ---- 32-bit-half arithmetic, placed on the pipes by hand ----------------------------------------
// Measured issue rates on the B200 (tools/pipe_rates.cu), warp instructions per clock per SM
// sub-partition: LOP3 / SHF / PRMT / IADD3 (ALU pipe) 0.5, IMAD 0.5, IMAD.WIDE and IMAD.HI 0.25, and
// ALU + IMAD interleave to 0.83.  The kernel is bound by integer issue, so the murmur arithmetic is
// written in 32-bit halves with the instruction chosen per operation: 64-bit rotates are two funnel
// shifts; x * C is one IMAD.WIDE + two IMAD; x * 5 + c is LEA + LEA.HI.X + add; shared-memory table
// addresses are idx * stride + base with the stride in a register (so ptxas keeps an IMAD instead of
// a shift + add); one of the three ">> 1" of each fmix goes through IMAD.HI to level the two pipes.
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
// k ^= k >> 33 touches the low word only.  HI_FIRST picks which of the three ">> 1" go through IMAD.HI
// (FMA pipe, 4 issue cycles) instead of SHF (ALU pipe, 2): measured pipe rates on B200
// (tools/pipe_rates.cu): LOP3/SHF/PRMT/IADD3/IMAD 0.5 per clock per sub-partition, IMAD.WIDE and
// IMAD.HI 0.25.  The mix below keeps the two pipes level.
template <int NHI>
__device__ __forceinline__ U2 fmix_u2(U2 k) {
    k.lo ^= (NHI >= 1 ? shr1_fma(k.hi) : (k.hi >> 1));  k = mul_u2<0xff51afd7ed558ccdULL>(k);
    k.lo ^= (NHI >= 2 ? shr1_fma(k.hi) : (k.hi >> 1));  k = mul_u2<0xc4ceb9fe1a85ec53ULL>(k);
    k.lo ^= (NHI >= 3 ? shr1_fma(k.hi) : (k.hi >> 1));
    return k;
}
__device__ __forceinline__ uint32_t byte_of(uint32_t x, int n) { return __byte_perm(x, 0u, 0x4440u + (uint32_t)n); }

// ---- murmur3 with the first multiply taken from shared-memory tables ---------------------------
// Every 8-base word w of the k-mer enters murmur3 as  w * c  (c = c1 for k1-type words, c2 for
// k2-type words), w being the 8 ASCII bytes of the bases.  Multiplication distributes over the
// byte groups:  w * c = A4(lo) * c + (A4(hi) * c << 32)  with A4(b) the 4 ASCII bytes of the 4
// bases b.  A 256-entry table of A4(b) * c turns "expand 2-bit codes to ASCII, then multiply" into
// two shared-memory loads and one add.
//
// Random table indices from 32 lanes conflict on the 32 shared-memory banks (measured: 6.2
// wavefronts per load with one copy of each table, which made the LSU data pipe the limiter).
// The tables are therefore REPLICATED: entry i of copy c sits at (i * R + c), and lane l reads
// copy l % R, so with R = 16 the 16 lanes of a 64-bit half-warp wavefront always hit 16 different
// bank pairs (no conflicts), and the 32-bit "high group" tables see at most 32/R lanes per copy.
constexpr int HP_R64 = 16;              // copies of the 64-bit tables (A4 * c, both words)
constexpr int HP_R32 = 8;               // copies of the 32-bit tables (low word of A4 * c)
struct HashConsts {                     // kernel parameters: values ptxas must not fold away
    uint32_t stride64, stride32, stride1;   // HP_R64 * 8, HP_R32 * 4, 8 (bytes between consecutive entries)
    uint32_t one;                           // 1: multiplier that keeps the table-word adds on the FMA pipe
    uint32_t five;                          // 5: the multiplier of h * 5 + c in its IMAD form
    uint32_t log_reserve;                   // extra log slots a warp reserves per atomic, <= 31 (candidate-dense launches: 31)
};
struct MulLut {                         // per-lane shared-window byte addresses (copy l % R already applied)
    uint32_t c1_64, c2_64, c1_32, c2_32, t1;
    uint32_t stride64, stride32, stride1, one, five;
    U2 add1, add2;                      // the +c of h * 5 + c
};

__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ U2 add_one(U2 a, U2 b, uint32_t /*one*/) {   // a + b (ptxas: IADD3 + IMAD.X; it splits IMAD.WIDE adds anyway)
    const uint64_t s = (((uint64_t)a.hi << 32) | a.lo) + (((uint64_t)b.hi << 32) | b.lo);
    U2 t; t.lo = (uint32_t)s; t.hi = (uint32_t)(s >> 32);
    return t;
}
// x * 5 + c.  LEA form: LEA + LEA.HI.X + the add (3 ALU + 1 FMA-pipe carry); IMAD form: IMAD.WIDE + IMAD (FMA pipe only,
// 6 issue cycles there, two instructions fewer).  Which one wins depends on which integer pipe is the longer pole of the
// whole body: chosen per use (hash_kernel's VAR).
template <bool LEA>
__device__ __forceinline__ U2 mul5add_r(U2 x, U2 c, uint32_t five) {
    if (LEA) {
        const uint64_t v = ((uint64_t)x.hi << 32) | x.lo;
        const uint64_t s = (v << 2) + v + (((uint64_t)c.hi << 32) | c.lo);
        U2 t; t.lo = (uint32_t)s; t.hi = (uint32_t)(s >> 32);
        return t;
    }
    U2 w = mad_wide(x.lo, five, c);          // `five` is a kernel parameter: ptxas keeps the multiply instead of rewriting it as shifts
    w.hi = mad_lo(x.hi, five, w.hi);
    return w;
}

// (NB bases, 1..8, of the 32-bit codes word cw starting at byte BYTE0) * c.  Bits of cw above the
// k-mer are zero (the rolling state is masked).
template <bool IS_C1, int BYTE0, int NB>
__device__ __forceinline__ U2 word_mul(uint32_t cw, const MulLut &L) {
    static_assert(NB >= 1 && NB <= 8 && (BYTE0 == 0 || BYTE0 == 2), "bad word");
    const uint32_t T64 = IS_C1 ? L.c1_64 : L.c2_64, T32 = IS_C1 ? L.c1_32 : L.c2_32;
    const uint64_t C = IS_C1 ? MM_C1 : MM_C2;
    U2 r;
    if (NB >= 4) {
        const uint2 a = lds64(mad_lo(byte_of(cw, BYTE0), L.stride64, T64));
        r.lo = a.x; r.hi = a.y;
        if (NB == 8) {
            r.hi = mad_lo(lds32(mad_lo(byte_of(cw, BYTE0 + 1), L.stride32, T32)), L.one, r.hi);
        } else if (NB > 4) {
            // the remaining 1..3 bases are the top of the (masked) word: a plain shift isolates them
            const uint32_t rem = BYTE0 == 0 ? byte_of(cw, 1) : (cw >> 24);
            r.hi = mad_lo(expand4(rem) & (uint32_t)low_bytes_mask(NB - 4), (uint32_t)C, r.hi);
        }
    } else {
        const uint32_t lo4 = expand4(byte_of(cw, BYTE0)) & (uint32_t)low_bytes_mask(NB);
        r = mul_wide(lo4, (uint32_t)C);
        r.hi = mad_lo(lo4, (uint32_t)(C >> 32), r.hi);
    }
    return r;
}

// VAR: bit 0 / bit 1 = h1 / h2 take the IMAD form of x * 5 + c; bits 2-3 / 4-5 = how many of the three ">> 1" of the
// first / second fmix go through IMAD.HI (default 1 each).
template <int K, bool SEED0, int VAR>
__device__ __forceinline__ U2 murmur_kmer_h1_lut(U2 codes, U2 seed, const MulLut &L) {
    static_assert(K >= 1 && K <= 32, "k out of range");
    constexpr bool LEA1 = !(VAR & 1), LEA2 = !(VAR & 2);
    constexpr int NHI_A = ((VAR >> 2) & 3) == 0 ? FMIX_HI_A : ((VAR >> 2) & 3) - 1, NHI_B = ((VAR >> 4) & 3) == 0 ? FMIX_HI_B : ((VAR >> 4) & 3) - 1;
    constexpr int NB = K / 16, T = K & 15;
    U2 h1 = seed, h2 = seed;
    if (NB >= 1) {
        U2 k1 = word_mul<true, 0, 8>(codes.lo, L);
        U2 k2 = word_mul<false, 2, 8>(codes.lo, L);
        k1 = rotl_u2<31>(k1); k1 = mul_u2<MM_C2>(k1);
        if (SEED0) { h1 = k1; } else { h1.lo ^= k1.lo; h1.hi ^= k1.hi; }
        h1 = rotl_u2<27>(h1); if (!SEED0) h1 = add_one(h1, h2, L.one); h1 = mul5add_r<LEA1>(h1, L.add1, L.five);
        k2 = rotl_u2<33>(k2); k2 = mul_u2<MM_C1>(k2);
        if (SEED0) { h2 = k2; } else { h2.lo ^= k2.lo; h2.hi ^= k2.hi; }
        h2 = rotl_u2<31>(h2); h2 = add_one(h2, h1, L.one); h2 = mul5add_r<LEA2>(h2, L.add2, L.five);
    }
    if (NB >= 2) {
        U2 k1 = word_mul<true, 0, 8>(codes.hi, L);
        U2 k2 = word_mul<false, 2, 8>(codes.hi, L);
        k1 = rotl_u2<31>(k1); k1 = mul_u2<MM_C2>(k1); h1.lo ^= k1.lo; h1.hi ^= k1.hi;
        h1 = rotl_u2<27>(h1); h1 = add_one(h1, h2, L.one); h1 = mul5add_r<LEA1>(h1, L.add1, L.five);
        k2 = rotl_u2<33>(k2); k2 = mul_u2<MM_C1>(k2); h2.lo ^= k2.lo; h2.hi ^= k2.hi;
        h2 = rotl_u2<31>(h2); h2 = add_one(h2, h1, L.one); h2 = mul5add_r<LEA2>(h2, L.add2, L.five);
    }
    // tail: 16-bit groups (k1-type first, then k2-type) of the next codes word; NB == 2 has none
    const uint32_t tword = (NB == 0) ? codes.lo : codes.hi;
    if (T > 8) {
        U2 k2 = word_mul<false, 2, (T > 8 ? T - 8 : 1)>(tword, L);
        k2 = rotl_u2<33>(k2); k2 = mul_u2<MM_C1>(k2); h2.lo ^= k2.lo; h2.hi ^= k2.hi;
    }
    if (T > 0) {
        U2 k1;
        if (K == 21) {                       // 5 bases = the whole (masked) high word of the codes: one load
            const uint2 a = lds64(mad_lo(tword, L.stride1, L.t1));
            k1.lo = a.x; k1.hi = a.y;
        } else k1 = word_mul<true, 0, (T >= 8 ? 8 : (T > 0 ? T : 1))>(tword, L);
        k1 = rotl_u2<31>(k1); k1 = mul_u2<MM_C2>(k1); h1.lo ^= k1.lo; h1.hi ^= k1.hi;
    }
    h1.lo ^= (uint32_t)K; h2.lo ^= (uint32_t)K;
    h1 = add_one(h1, h2, L.one); h2 = add_one(h2, h1, L.one);
    h1 = fmix_u2<NHI_A>(h1); h2 = fmix_u2<NHI_B>(h2);
    return add_one(h1, h2, L.one);
}

// A warp reserves n + log_reserve log slots per atomic on the launch's counter (same-address L2 atomics
// serialise at ~1 per ns: a candidate-dense launch must not issue one per position).

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

// ---- the persistent kernel ---------------------------------------------------------------------------
// One CTA of HP_WARPS warps per SM builds the replicated tables once, then every warp streams
// through work items on its own: an item is HP_SLICE = 32 lanes x HASH_W consecutive positions of
// one symbol region, staged (with a 32-symbol halo in front) into the warp's own shared-memory
// buffer by one 1-D TMA bulk copy signalled on the warp's own mbarrier.  With two buffers per warp
// the copy of the next item is in flight while the current one is hashed; there is no block-wide
// synchronisation after the table build.
constexpr int HP_WARPS = 32;
constexpr int HP_NBUF = 2;               // (the main loop is written for two)
constexpr uint32_t HP_SLICE = 32u * HASH_W;                  // positions per item of uniform 64-position pieces
constexpr uint32_t HP_BUF = 2208u;                           // bytes per warp buffer (multiple of 16): 32-symbol halo + up to 15
                                                             // symbols of alignment + PIECE_SPAN symbols + rounding
constexpr uint32_t HP_SPAN_MAX = HP_BUF - 32u - 16u;         // symbols from the aligned start a staged group may cover
static_assert(HP_SPAN_MAX >= PIECE_SPAN + 15u && HP_SPAN_MAX >= HP_SLICE + 15u, "a full group of planned pieces must fit the buffer");
constexpr uint32_t HP_OFF_C1_64 = 0;
constexpr uint32_t HP_OFF_C2_64 = HP_OFF_C1_64 + 256u * HP_R64 * 8u;
constexpr uint32_t HP_OFF_C1_32 = HP_OFF_C2_64 + 256u * HP_R64 * 8u;
constexpr uint32_t HP_OFF_C2_32 = HP_OFF_C1_32 + 256u * HP_R32 * 4u;
constexpr uint32_t HP_OFF_T1 = HP_OFF_C2_32 + 256u * HP_R32 * 4u;
constexpr uint32_t HP_OFF_BAR = HP_OFF_T1 + 1024u * 8u;   // T1: (5 ASCII bytes) * c1 for k = 21, one copy
constexpr uint32_t HP_OFF_BUF = (HP_OFF_BAR + HP_WARPS * HP_NBUF * 8u + 127u) & ~127u;
constexpr uint32_t HP_SMEM = HP_OFF_BUF + HP_WARPS * HP_NBUF * HP_BUF;
static_assert(HP_SMEM <= 232448u, "hash kernel shared memory exceeds 227 KB");
static_assert(HP_BUF % 16u == 0, "TMA size granularity");

template <int K, bool SEED0, int VAR>
__global__ void __launch_bounds__(HP_WARPS * 32, 1)
hash_kernel(const uint8_t *__restrict__ symbuf,   // symbol buffer: region r starts at SYM_FRONT + r * region_stride
            ChunkGeom g, uint32_t r0, uint32_t n_regions, // regions [r0, r0 + n_regions) of this launch
            const uint32_t *__restrict__ region_count, const ParseCarry *__restrict__ carry, uint64_t ord_base,
            const SketchState *st,
            LaunchSlot *slot, LogView log, int k_rt, uint64_t seed, HashConsts hc, PiecePlan pp) {
    extern __shared__ __align__(128) uint8_t hp_smem[];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    // ---- tables (once per CTA) and the warps' mbarriers ----
    if (K > 0) {
        uint2 *c1_64 = reinterpret_cast<uint2 *>(hp_smem + HP_OFF_C1_64), *c2_64 = reinterpret_cast<uint2 *>(hp_smem + HP_OFF_C2_64);
        uint32_t *c1_32 = reinterpret_cast<uint32_t *>(hp_smem + HP_OFF_C1_32), *c2_32 = reinterpret_cast<uint32_t *>(hp_smem + HP_OFF_C2_32);
        uint2 *t1 = reinterpret_cast<uint2 *>(hp_smem + HP_OFF_T1);
        for (uint32_t i = tid; i < 256u * HP_R64; i += blockDim.x) {
            const uint32_t a4 = expand4(i / HP_R64);
            const uint64_t p1 = (uint64_t)a4 * MM_C1, p2 = (uint64_t)a4 * MM_C2;
            c1_64[i] = make_uint2((uint32_t)p1, (uint32_t)(p1 >> 32));
            c2_64[i] = make_uint2((uint32_t)p2, (uint32_t)(p2 >> 32));
        }
        for (uint32_t i = tid; i < 256u * HP_R32; i += blockDim.x) {
            const uint32_t a4 = expand4(i / HP_R32);
            c1_32[i] = a4 * (uint32_t)MM_C1;
            c2_32[i] = a4 * (uint32_t)MM_C2;
        }
        if (K == 21 && tid < 1024u) {        // 5-base tail word of k = 21
            const uint64_t w = (uint64_t)expand4(tid & 0xFFu) | ((uint64_t)(expand4(tid >> 8) & 0xFFu) << 32);
            const uint64_t p = w * MM_C1;
            t1[tid] = make_uint2((uint32_t)p, (uint32_t)(p >> 32));
        }
    }
    uint64_t *bars = reinterpret_cast<uint64_t *>(hp_smem + HP_OFF_BAR) + warp * HP_NBUF;
    uint8_t *bufs = hp_smem + HP_OFF_BUF + warp * (HP_NBUF * HP_BUF);
    if (lane == 0) {
#pragma unroll
        for (int n = 0; n < HP_NBUF; ++n) mbar_init(&bars[n], 1);
        fence_mbar_init();
    }
    __syncthreads();

    MulLut L;
    L.c1_64 = smem_u32(hp_smem + HP_OFF_C1_64) + (lane % HP_R64) * 8u;
    L.c2_64 = smem_u32(hp_smem + HP_OFF_C2_64) + (lane % HP_R64) * 8u;
    L.c1_32 = smem_u32(hp_smem + HP_OFF_C1_32) + (lane % HP_R32) * 4u;
    L.c2_32 = smem_u32(hp_smem + HP_OFF_C2_32) + (lane % HP_R32) * 4u;
    L.t1 = smem_u32(hp_smem + HP_OFF_T1);
    L.stride64 = hc.stride64; L.stride32 = hc.stride32; L.stride1 = hc.stride1; L.one = hc.one; L.five = hc.five;
    L.add1.lo = 0x52dce729u; L.add1.hi = 0u; L.add2.lo = 0x38495ab5u; L.add2.hi = 0u;

    const int k = K > 0 ? K : k_rt;
    const uint64_t mask = kmer_mask(k);
    const unsigned long long T = st->threshold;
    const uint32_t T_hi = (uint32_t)(T >> 32);
    U2 seed2; seed2.lo = (uint32_t)seed; seed2.hi = (uint32_t)(seed >> 32);

    // Work: the hash pieces the parse kernel planned (PiecePlan, device_types.cuh): runs of k-mer end positions, a
    // lane walks one.  An item is a group of 32 consecutive pieces of one region, in GROUP-major order (item it = group
    // it / n_regions of region r0 + it % n_regions) so that concurrently running warps spread over the regions; the item
    // space is cut at the chunk's largest piece count, and items are claimed from a per-launch counter.  The symbols
    // under a group (from 32 in front of its first position, 16-byte aligned, to its last position) are staged with
    // one TMA copy; a group whose pieces lie further apart than the buffer (records without any k-mer in between) is
    // taken in several rounds.
    const uint32_t groups = min((carry->max_region_pieces + 31u) / 32u, (pp.stride + 31u) / 32u);
    const uint32_t w1 = groups * n_regions;      // one past the last item
    uint32_t it_region = 0, it_i0 = 0, it_i1 = 0;                 // pieces [it_i0, it_i1) of the claimed item not staged yet
    uint32_t nx_region = 0, nx_pbal = 0, nx_bytes = 0, nx_piece = 0;   // the group staged next (nx_piece: this lane's piece, 0 = none)
    auto next_group = [&]() -> bool {
        while (it_i0 >= it_i1) {
            uint32_t it = 0;
            if (lane == 0) it = atomicAdd(&slot->next_item, 1u);
            it = __shfl_sync(0xffffffffu, it, 0);
            if (it >= w1) return false;
            const uint32_t grp = it / n_regions, region = r0 + (it - grp * n_regions);
            const uint32_t np = min(pp.count[region], pp.stride);
            if (32u * grp < np) { it_region = region; it_i0 = 32u * grp; it_i1 = min(np, 32u * grp + 32u); }
        }
        const uint32_t idx = it_i0 + lane;
        uint32_t piece = idx < it_i1 ? pp.table[(size_t)it_region * pp.stride + idx] : 0u;
        const uint32_t pbal = (__shfl_sync(0xffffffffu, piece, 0) & 0xFFFFu) & ~15u;
        const uint32_t endl = (piece & 0xFFFFu) + (piece >> 16);
        const uint32_t fits = __ballot_sync(0xffffffffu, idx < it_i1 && endl - pbal <= HP_SPAN_MAX);
        const uint32_t cnt = fits == 0xffffffffu ? 32u : (uint32_t)(__ffs(~fits) - 1);   // >= 1: one piece always fits
        if (lane >= cnt) piece = 0u;
        const uint32_t last_end = __shfl_sync(0xffffffffu, endl, (int)cnt - 1);
        nx_region = it_region; nx_pbal = pbal; nx_piece = piece;
        nx_bytes = (32u + (last_end - pbal) + 15u) & ~15u;
        it_i0 += cnt;
        return true;
    };
    // stage [pbal - 32, last position of the group] of the region into buffer n
    auto issue = [&](int n) {
        if (lane == 0) {
            const uint8_t *src = symbuf + (size_t)SYM_FRONT + (size_t)nx_region * g.region_stride + nx_pbal - 32;
            mbar_arrive_expect_tx(&bars[n], nx_bytes);
            tma_load_1d(bufs + n * HP_BUF, src, nx_bytes, &bars[n]);
        }
    };

    uint32_t nvalid = 0;
    uint32_t res_base = 0, res_left = 0;   // warp-uniform: this warp's reserved slice of the log
    uint32_t phases = 0;                   // bit n: parity to wait for on buffer n
    int cb = 0;
    bool have = next_group();
    if (have) issue(0);
    while (have) {
        const uint32_t region = nx_region, pbal = nx_pbal, piece = nx_piece;
        const bool have_next = next_group();           // the next group's copy is in flight while this one is hashed
        if (have_next) issue(cb ^ 1);
        mbar_wait(&bars[cb], (phases >> cb) & 1u);
        phases ^= 1u << cb;
        const uint8_t *tile = bufs + cb * HP_BUF;       // tile[32 + x] = symbol pbal + x of the region
        const uint64_t ord_region = ord_base + (uint64_t)region * g.st_bytes;
        // this lane's run [p0, p0 + n): walked from p0 exactly, four symbols per step -- the staged bytes are read as
        // aligned words and shifted into place (one PRMT per step), so a run costs ceil(n / 4) steps wherever it starts
        const uint32_t n = piece >> 16, off = (piece & 0xFFFFu) - pbal;
        const bool live = n != 0u;
        const uint32_t t0 = live ? off : 0u;
        const int nend = (int)n;                        // walk positions [0, nend) are the lane's
        const uint32_t nw4 = ((uint32_t)__reduce_max_sync(0xffffffffu, nend) + 3u) >> 2;
        const uint32_t pw = pbal + t0;                  // region position of the walk's first symbol
        const uint32_t sel = 0x3210u + 0x1111u * (t0 & 3u);   // PRMT selector: bytes (t0 & 3) .. (t0 & 3) + 3 of a word pair

        Roll2<K> r; r.A.lo = r.A.hi = r.B.lo = r.B.hi = 0;
        // brk = position (relative to the current group of 4) of the last non-base symbol; the window
        // ending at relative position b is valid  <=>  brk <= b - k.
        int brk = -k;
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(tile + 32 + (t0 & ~3u));   // aligned word holding the walk's first symbol
        // ---- warm-up on the 32 symbols before the walk (only the last k-1 matter) -------------------
        if (live) {
            const uint32_t *hw = reinterpret_cast<const uint32_t *>(tile + (t0 & ~3u));
            uint32_t w[8];
            {
                uint32_t a = hw[0];
#pragma unroll
                for (int q = 0; q < 8; ++q) { const uint32_t b2 = hw[q + 1]; w[q] = __byte_perm(a, b2, sel); a = b2; }
            }
            if (K == 21) {
                // The 20 symbols in front are the words w[3..7]: build both rolling words at once instead of
                // 20 pushes.  pack4: the 2-bit codes of 4 symbol bytes -> one byte (multiply gathers them).
                uint32_t y[5];
#pragma unroll
                for (int q = 0; q < 5; ++q) y[q] = ((w[3 + q] & 0x03030303u) * 0x01041040u) >> 24;
                const uint32_t plo = y[0] | (y[1] << 8) | (y[2] << 16) | (y[3] << 24), phi = y[4];   // code j at bits [2j, 2j+1]
                // A = sum c_j << (2 + 2j): the 20 symbols sit one place above the slot the next push fills
                r.A.lo = plo << 2; r.A.hi = __funnelshift_l(plo, phi, 2);
                // B = sum (3 - c_j) << 2(19 - j): reverse the pair order (bit reverse + swap inside pairs), complement
                uint32_t rhi = __brev(plo), rlo = __brev(phi);
                rhi = ((rhi >> 1) & 0x55555555u) | ((rhi & 0x55555555u) << 1);
                rlo = ((rlo >> 1) & 0x55555555u) | ((rlo & 0x55555555u) << 1);
                r.B.lo = ~__funnelshift_r(rlo, rhi, 24);
                r.B.hi = ~(rhi >> 24) & 0xFFu;
#pragma unroll
                for (int q = 0; q < 5; ++q) {              // last non-base symbol (symbols are 0..4: bit 2 marks a break)
                    const uint32_t t = w[3 + q] & 0x04040404u;
                    if (t) brk = 4 * (3 + q) + ((31 - __clz(t)) >> 3) - 32;
                }
            } else if (K == 31) {
                // same for k = 31: the 30 symbols in front are bytes 2..31 of the halo
                uint32_t y[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) y[q] = ((w[q] & 0x03030303u) * 0x01041040u) >> 24;
                const uint32_t plo = y[0] | (y[1] << 8) | (y[2] << 16) | (y[3] << 24);
                const uint32_t phi = y[4] | (y[5] << 8) | (y[6] << 16) | (y[7] << 24);     // code of byte i at bits [2i, 2i+1]
                r.A.lo = __funnelshift_r(plo, phi, 2) & ~3u; r.A.hi = phi >> 2;            // codes 2..31 at pairs 1..30
                uint32_t rhi = __brev(plo), rlo = __brev(phi);
                rhi = ((rhi >> 1) & 0x55555555u) | ((rhi & 0x55555555u) << 1);
                rlo = ((rlo >> 1) & 0x55555555u) | ((rlo & 0x55555555u) << 1);
                r.B.lo = ~rlo; r.B.hi = ~rhi & 0x0FFFFFFFu;                                 // complement of code i at pair 31 - i, i >= 2
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint32_t t = w[q] & (q == 0 ? 0x04040000u : 0x04040404u);
                    if (t) brk = 4 * q + ((31 - __clz(t)) >> 3) - 32;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (K > 0 && i < 32 - (K - 1)) continue;  // compile-time skip
                    const uint32_t s = w[i >> 2] >> (8 * (i & 3));
                    if (K == 0 && i < 32 - (k - 1)) continue;
                    r.push(s, k, mask);
                    if ((s & 0xFFu) >= 4u) brk = i - 32;
                }
            }
        }
        uint32_t wlo = nend > 0 ? wp[0] : 0u, whi = nend > 0 ? wp[1] : 0u;
        uint32_t word = nend > 0 ? __byte_perm(wlo, whi, sel) : 0x04040404u;
        int lim = nend;                                 // positions of the current group of 4 below lim are the lane's
#pragma unroll 1
        for (uint32_t j = 0; j < nw4; ++j) {
            const uint32_t cw = word;
            if (j + 1u < nw4) {                         // next 4 symbols (lanes past their run read nothing)
                wlo = whi;
                if (lim > 4) whi = wp[j + 2];
                word = lim > 4 ? __byte_perm(wlo, whi, sel) : 0x04040404u;
            }
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                r.push(cw >> (8 * b), k, mask);
                if (cw & (0xFCu << (8 * b))) brk = b;
                const bool ok = brk <= b - k && b < lim;
                const bool is_rc = (((uint64_t)r.A.hi << 32) | r.A.lo) >= (((uint64_t)r.B.hi << 32) | r.B.lo);
                U2 codes; codes.lo = is_rc ? r.B.lo : r.A.lo; codes.hi = is_rc ? r.B.hi : r.A.hi;
                U2 h;
                if (K > 0) h = murmur_kmer_h1_lut<(K > 0 ? K : 1), SEED0, VAR>(codes, seed2, L);
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
                    // result) happens once per reservation, not once per candidate.
                    const uint32_t n = __popc(em);
                    if (n > res_left) {                                   // warp-uniform
                        if (lane < res_left && res_base + lane < log.cap) log.posx[res_base + lane] = ~0ULL;  // unused slots (res_left <= 31)
                        const uint32_t want = n + hc.log_reserve;
                        uint32_t base = 0;
                        if (lane == 0) base = atomicAdd(&slot->log_count, want);
                        res_base = __shfl_sync(0xffffffffu, base, 0);
                        res_left = want;
                    }
                    if (emit) {
                        const uint32_t idx = res_base + __popc(em & lanemask_lt());
                        if (idx < log.cap) {
                            const uint32_t p = pw + 4u * j + (uint32_t)b;
                            log.hash[idx] = hv;
                            log.kmer[idx] = ((unsigned long long)codes.hi << 32) | codes.lo;   // k <= 32: log.kw == 1
                            log.posx[idx] = ((ord_region + p) << 9) | (is_rc ? 1ull : 0ull);
                        }
                    }
                    res_base += n; res_left -= n;
                  }
                }
            }
            brk -= 4; lim -= 4;
        }
        __syncwarp();                      // every lane is done with this buffer before it is refilled
        have = have_next; cb ^= 1;
    }
    if (lane < res_left && res_base + lane < log.cap) log.posx[res_base + lane] = ~0ULL;  // unused tail of the reservation
    // valid-window count of this launch (committed to total_kmers by the host on success)
    nvalid = __reduce_add_sync(0xffffffffu, nvalid);
    if (lane == 0 && nvalid) atomicAdd(&slot->launch_kmers, (unsigned long long)nvalid);
}

// ---- 33 <= k <= 255: the exact, slower kernel -------------------------------------------------------
// The reference takes any u8 k-mer length (sketch_schemes/mod.rs:59, cli.rs:161-166).  Beyond 32 bases the
// strands no longer fit one 64-bit word, so this kernel does what needletail's canonical_kmers and the murmur3
// crate do, literally, per position: byte-wise lexicographic compare of the forward window against its reverse
// complement (first difference decides; a palindrome reports rc), then murmur3_x64_128 over the k ASCII bytes of
// the chosen strand, 16-byte block by block.  A thread owns HB_W consecutive k-mer end positions and carries the
// position of the last non-base symbol; symbols are read straight from the region (the 256-symbol front pad holds
// what precedes it in stream order).  Candidates carry their k-mer as ceil(k / 32) words of 2-bit codes.
constexpr int HB_THREADS = 256, HB_W = 16;
__device__ __forceinline__ uint32_t ascii_of(uint32_t code) { return (0x54474341u >> (8u * code)) & 0xFFu; }
__global__ void __launch_bounds__(HB_THREADS)
hash_big_kernel(const uint8_t *__restrict__ symbuf, ChunkGeom g, uint32_t r0, const uint32_t *__restrict__ region_count,
                uint64_t ord_base, const SketchState *st, LaunchSlot *slot, LogView log, int k, uint64_t seed) {
    const uint32_t region = r0 + blockIdx.y, lane = threadIdx.x & 31u;
    const int end = (int)region_count[region];
    if ((int)(blockIdx.x * HB_THREADS * HB_W) >= end) return;          // block-uniform
    const uint8_t *sym = symbuf + (size_t)SYM_FRONT + (size_t)region * g.region_stride;
    const uint64_t ord_region = ord_base + (uint64_t)region * g.st_bytes;
    const unsigned long long T = st->threshold;
    const int p0 = (int)((blockIdx.x * HB_THREADS + threadIdx.x) * HB_W);
    int lastb = p0 - k;                                               // last non-base symbol before the current position
    if (p0 < end)
        for (int i = p0 - 1; i > p0 - k; --i) if (sym[i] >= 4u) { lastb = i; break; }
    uint32_t nvalid = 0;
    for (int j = 0; j < HB_W; ++j) {
        const int p = p0 + j;
        const bool live = p < end;
        bool ok = false;
        if (live) {
            if (sym[p] >= 4u) lastb = p;
            ok = p - lastb >= k;
        }
        unsigned long long h1 = 0;
        bool is_rc = false;
        if (ok) {
            const uint8_t *f = sym + (p - k + 1);                     // forward window: f[i]; reverse complement: 3 - f[k-1-i]
            is_rc = true;                                             // tie (palindrome) => rc, as needletail's `fwd < rc` test
            for (int i = 0; i < k; ++i) {
                const uint32_t a = f[i], b = 3u - f[k - 1 - i];
                if (a != b) { is_rc = !(a < b); break; }
            }
            auto byte_at = [&](int i) -> unsigned long long {
                return (unsigned long long)ascii_of(is_rc ? 3u - f[k - 1 - i] : (uint32_t)f[i]);
            };
            unsigned long long h2 = seed;
            h1 = seed;
            const int nblocks = k / 16;
            for (int blk = 0; blk < nblocks; ++blk) {
                unsigned long long k1 = 0, k2 = 0;
                for (int q = 7; q >= 0; --q) { k1 = (k1 << 8) | byte_at(16 * blk + q); k2 = (k2 << 8) | byte_at(16 * blk + 8 + q); }
                k1 *= MM_C1; k1 = rotl64(k1, 31); k1 *= MM_C2; h1 ^= k1;
                h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729ULL;
                k2 *= MM_C2; k2 = rotl64(k2, 33); k2 *= MM_C1; h2 ^= k2;
                h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5ULL;
            }
            const int t = k & 15, tb = 16 * nblocks;
            if (t > 8) {
                unsigned long long k2 = 0;
                for (int q = t - 1; q >= 8; --q) k2 = (k2 << 8) | byte_at(tb + q);
                k2 *= MM_C2; k2 = rotl64(k2, 33); k2 *= MM_C1; h2 ^= k2;
            }
            if (t > 0) {
                unsigned long long k1 = 0;
                for (int q = (t > 8 ? 8 : t) - 1; q >= 0; --q) k1 = (k1 << 8) | byte_at(tb + q);
                k1 *= MM_C1; k1 = rotl64(k1, 31); k1 *= MM_C2; h1 ^= k1;
            }
            h1 ^= (unsigned long long)k; h2 ^= (unsigned long long)k;
            h1 += h2; h2 += h1;
            h1 = fmix64(h1); h2 = fmix64(h2);
            h1 += h2;
            ++nvalid;
        }
        const bool emit = ok && h1 <= T;
        const uint32_t em = __ballot_sync(0xffffffffu, emit);
        if (em) {
            uint32_t base = 0;
            if (lane == (uint32_t)(__ffs(em) - 1)) base = atomicAdd(&slot->log_count, (unsigned int)__popc(em));
            base = __shfl_sync(0xffffffffu, base, __ffs(em) - 1);
            if (emit) {
                const uint32_t idx = base + __popc(em & lanemask_lt());
                if (idx < log.cap) {
                    const uint8_t *f = sym + (p - k + 1);
                    log.hash[idx] = h1;
                    log.posx[idx] = ((ord_region + (uint64_t)p) << 9) | (is_rc ? 1ull : 0ull);
                    unsigned long long *kw = log.kmer + (size_t)idx * log.kw;
                    for (uint32_t w = 0; w < log.kw; ++w) {
                        unsigned long long acc = 0;
                        for (int i = 32 * (int)w; i < k && i < 32 * (int)w + 32; ++i)
                            acc |= (unsigned long long)(is_rc ? 3u - f[k - 1 - i] : (uint32_t)f[i]) << (2 * (i & 31));
                        kw[w] = acc;
                    }
                }
            }
        }
    }
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
            log.kmer[(size_t)idx * log.kw] = arena_base + i;
            log.posx[idx] = ((ord_base + i) << 9) | (1ull << 8) | (unsigned long long)extra[i];
        }
    }
}
__global__ void push_commit_kernel(LaunchSlot *slot, uint32_t n) { slot->launch_kmers += n; }

template <int K, bool SEED0, int VAR>
static void launch_hash_ks(uint32_t r0, uint32_t n_regions, const uint8_t *symbuf, ChunkGeom g, const uint32_t *region_count,
                           const ParseCarry *carry, uint64_t ord_base, const SketchState *st, LaunchSlot *slot, LogView log,
                           int k, uint64_t seed, uint32_t log_reserve, PiecePlan pp, cudaStream_t stream) {
    static int sms[64] = {};   // per device: SM count, set once the shared-memory attribute is in place
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!sms[dev]) {
        cudaFuncSetAttribute(hash_kernel<K, SEED0, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HP_SMEM);
        int n = 148;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        sms[dev] = n > 0 ? n : 148;
    }
    HashConsts hc;
    hc.stride64 = HP_R64 * 8u; hc.stride32 = HP_R32 * 4u; hc.stride1 = 8u; hc.one = 1u; hc.five = 5u;
    hc.log_reserve = log_reserve;
    const uint64_t items_ub = (uint64_t)n_regions * ((pp.stride + 31u) / 32u);   // the device cuts it at the chunk's largest piece count
    const uint32_t ctas = (uint32_t)std::min<uint64_t>((uint64_t)sms[dev], (items_ub + HP_WARPS - 1) / HP_WARPS);
    hash_kernel<K, SEED0, VAR><<<ctas, HP_WARPS * 32, HP_SMEM, stream>>>(symbuf, g, r0, n_regions, region_count, carry, ord_base, st,
                                                                   slot, log, k, seed, hc, pp);
}
template <int K>
static void launch_hash_k(uint32_t r0, uint32_t n_regions, const uint8_t *symbuf, ChunkGeom g, const uint32_t *region_count,
                          const ParseCarry *carry, uint64_t ord_base, const SketchState *st, LaunchSlot *slot, LogView log,
                          int k, uint64_t seed, uint32_t log_reserve, PiecePlan pp, cudaStream_t stream) {
#ifdef FB2_HASH_VARIANTS   // A/B builds: FB2_HASH_VAR picks the arithmetic variant of the k = 21 / seed 0 kernel at run time
    if (seed == 0 && K == 21) {
        static const int var = getenv("FB2_HASH_VAR") ? atoi(getenv("FB2_HASH_VAR")) : HASH_VAR_DEFAULT;
#define FB2_V(V) case V: launch_hash_ks<K, true, V>(r0, n_regions, symbuf, g, region_count, carry, ord_base, st, slot, log, k, seed, log_reserve, pp, stream); return;
        switch (var) { FB2_V(0) FB2_V(1) FB2_V(2) FB2_V(3) FB2_V(12) FB2_V(15) FB2_V(60) FB2_V(63) default: break; }
#undef FB2_V
    }
#endif
    if (seed == 0 && K > 0) launch_hash_ks<K, true, HASH_VAR_DEFAULT>(r0, n_regions, symbuf, g, region_count, carry, ord_base, st, slot, log, k, seed, log_reserve, pp, stream);
    else launch_hash_ks<K, false, HASH_VAR_DEFAULT>(r0, n_regions, symbuf, g, region_count, carry, ord_base, st, slot, log, k, seed, log_reserve, pp, stream);
}
// Hash the symbol regions [r0, r1) of a chunk.
void launch_hash(int k, const uint8_t *symbuf, ChunkGeom g, uint32_t r0, uint32_t r1, const uint32_t *region_count,
                 const ParseCarry *carry, uint64_t ord_base, const SketchState *st, LaunchSlot *slot, LogView log,
                 uint64_t seed, uint32_t log_reserve, PiecePlan pp, cudaStream_t stream) {
    if (r1 <= r0) return;
    if (k > 32) {
        const dim3 grid((g.st_bytes + HB_THREADS * HB_W - 1) / (HB_THREADS * HB_W), r1 - r0);
        hash_big_kernel<<<grid, HB_THREADS, 0, stream>>>(symbuf, g, r0, region_count, ord_base, st, slot, log, k, seed);
        return;
    }
    if (k == 21) launch_hash_k<21>(r0, r1 - r0, symbuf, g, region_count, carry, ord_base, st, slot, log, k, seed, log_reserve, pp, stream);
    else if (k == 31) launch_hash_k<31>(r0, r1 - r0, symbuf, g, region_count, carry, ord_base, st, slot, log, k, seed, log_reserve, pp, stream);
    else launch_hash_k<0>(r0, r1 - r0, symbuf, g, region_count, carry, ord_base, st, slot, log, k, seed, log_reserve, pp, stream);
}
void launch_push_hash(const uint8_t *bytes, const uint32_t *offs, const uint8_t *extra, uint32_t n,
                      uint64_t arena_base, uint64_t ord_base, const SketchState *st, LaunchSlot *slot, LogView log,
                      uint64_t seed, cudaStream_t stream) {
    if (!n) return;
    push_hash_kernel<<<(n + 255) / 256, 256, 0, stream>>>(bytes, offs, extra, n, arena_base, ord_base, st, slot, log, seed);
    push_commit_kernel<<<1, 1, 0, stream>>>(slot, n);
}

}  // namespace fb2

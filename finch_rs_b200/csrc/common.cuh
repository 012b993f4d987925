// common.cuh -- arithmetic shared by the sm_100a kernels (and, for unit tests only, by a host
// build of the same inline functions: tests/kernel_math_host.cpp).
//
// Restates, for the GPU, the arithmetic of
//   murmurhash3 0.0.5 :: murmurhash3_x64_128        (call site lib/src/sketch_schemes/hashing.rs:11)
//   needletail 0.5.0  :: normalize / canonical_kmers (call sites lib/src/sketch_schemes/mash.rs:73-76)
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define FB2_HD __host__ __device__ __forceinline__
#else
#define FB2_HD inline
#endif

namespace fb2 {

// ---- byte classes of the raw stream (needletail normalize(false)) -------------
// 0..3 = A,C,G,T (a,c,g,t,u,U fold in); everything below is not a base.
enum : uint8_t {
    CLS_A = 0, CLS_C = 1, CLS_G = 2, CLS_T = 3,
    CLS_BAD = 4,   // maps to 'N' / '-' : occupies a position, never part of a k-mer
    CLS_WS = 5,    // ' ' and '\t'      : removed by normalize
    CLS_NL = 6,    // '\n'              : removed by normalize; line structure
    CLS_GT = 7,    // '>'  (BAD as a symbol; FASTA header start at line start)
    CLS_CR = 8,    // '\r'              : removed by normalize; trimmed at line ends
    CLS_AT = 9,    // '@'  (BAD as a symbol; FASTQ header check)
    CLS_PLUS = 10  // '+'  (BAD as a symbol; FASTQ separator check)
};
constexpr uint8_t SYM_BREAK = 4;  // symbol-stream code for "not a base" (N, record break)

constexpr FB2_HD uint8_t classify_byte(uint8_t c) {
    switch (c) {
    case 'A': case 'a': return CLS_A;
    case 'C': case 'c': return CLS_C;
    case 'G': case 'g': return CLS_G;
    case 'T': case 't': case 'U': case 'u': return CLS_T;
    case ' ': case '\t': return CLS_WS;
    case '\n': return CLS_NL;
    case '\r': return CLS_CR;
    case '>': return CLS_GT;
    case '@': return CLS_AT;
    case '+': return CLS_PLUS;
    default: return CLS_BAD;
    }
}

// ---- murmur3 x64_128 ------------------------------------------------------------------------
constexpr uint64_t MM_C1 = 0x87c37b91114253d5ULL;
constexpr uint64_t MM_C2 = 0x4cf5ad432745937fULL;

FB2_HD uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
FB2_HD uint64_t fmix64(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

// 4 bases (2 bits each, base 0 in bits 0-1) -> 4 ASCII bytes, base 0 in byte 0.
FB2_HD uint32_t expand4(uint32_t b8) {
#if defined(__CUDA_ARCH__)
    uint32_t x = (b8 | (b8 << 4)) & 0x0F0Fu;
    x = (x | (x << 2)) & 0x3333u;               // one base per nibble
    return __byte_perm(0x54474341u, 0u, x);     // nibble -> 'A','C','G','T'
#else
    const uint32_t lut[4] = {0x41u, 0x43u, 0x47u, 0x54u};
    return lut[b8 & 3] | (lut[(b8 >> 2) & 3] << 8) | (lut[(b8 >> 4) & 3] << 16) | (lut[(b8 >> 6) & 3] << 24);
#endif
}
// 8 bases (16 bits) -> 8 ASCII bytes as a little-endian u64 (byte 0 = base 0).
FB2_HD uint64_t expand8(uint32_t b16) {
    return (uint64_t)expand4(b16 & 0xFFu) | ((uint64_t)expand4((b16 >> 8) & 0xFFu) << 32);
}
FB2_HD uint64_t low_bytes_mask(int nbytes) {  // nbytes in 0..8
    return nbytes >= 8 ? ~0ULL : ((1ULL << (8 * nbytes)) - 1ULL);
}

// h1 of murmurhash3_x64_128 over the k ASCII bytes of a k-mer given as 2-bit codes,
// base i in bits [2i, 2i+1] ("LSB-first"), 1 <= k <= 32.  K > 0 fixes k at compile time.
template <int K>
FB2_HD uint64_t murmur_kmer_h1(uint64_t codes, int k_rt, uint64_t seed) {
    const int k = K > 0 ? K : k_rt;
    uint64_t h1 = seed, h2 = seed;
    uint64_t w0 = expand8((uint32_t)(codes & 0xFFFFu));
    uint64_t w1 = expand8((uint32_t)((codes >> 16) & 0xFFFFu));
    uint64_t t1, t2;   // tail words (bytes 16*nblocks ..)
    if (k >= 16) {
        uint64_t k1 = w0, k2 = w1;
        k1 *= MM_C1; k1 = rotl64(k1, 31); k1 *= MM_C2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729ULL;
        k2 *= MM_C2; k2 = rotl64(k2, 33); k2 *= MM_C1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5ULL;
        t1 = expand8((uint32_t)((codes >> 32) & 0xFFFFu));
        t2 = expand8((uint32_t)((codes >> 48) & 0xFFFFu));
        if (k == 32) {
            k1 = t1; k2 = t2;
            k1 *= MM_C1; k1 = rotl64(k1, 31); k1 *= MM_C2; h1 ^= k1;
            h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729ULL;
            k2 *= MM_C2; k2 = rotl64(k2, 33); k2 *= MM_C1; h2 ^= k2;
            h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5ULL;
        }
    } else {
        t1 = w0; t2 = w1;
    }
    const int t = k & 15;
    if (t > 8) {
        uint64_t k2 = t2 & low_bytes_mask(t - 8);
        k2 *= MM_C2; k2 = rotl64(k2, 33); k2 *= MM_C1; h2 ^= k2;
    }
    if (t > 0) {
        uint64_t k1 = t1 & low_bytes_mask(t > 8 ? 8 : t);
        k1 *= MM_C1; k1 = rotl64(k1, 31); k1 *= MM_C2; h1 ^= k1;
    }
    h1 ^= (uint64_t)k; h2 ^= (uint64_t)k;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2;
    return h1;
}

// Generic murmur3_x64_128 h1 over raw bytes (used by the `push` unit-test surface).
FB2_HD uint64_t murmur_bytes_h1(const uint8_t *data, uint32_t len, uint64_t seed) {
    uint64_t h1 = seed, h2 = seed;
    const uint32_t nblocks = len / 16;
    for (uint32_t b = 0; b < nblocks; ++b) {
        uint64_t k1 = 0, k2 = 0;
        for (int i = 7; i >= 0; --i) { k1 = (k1 << 8) | data[16 * b + i]; k2 = (k2 << 8) | data[16 * b + 8 + i]; }
        k1 *= MM_C1; k1 = rotl64(k1, 31); k1 *= MM_C2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729ULL;
        k2 *= MM_C2; k2 = rotl64(k2, 33); k2 *= MM_C1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5ULL;
    }
    const uint8_t *tail = data + 16 * nblocks;
    const uint32_t t = len & 15;
    uint64_t k1 = 0, k2 = 0;
    if (t > 8) {
        for (uint32_t i = t; i > 8; --i) k2 = (k2 << 8) | tail[i - 1];
        k2 *= MM_C2; k2 = rotl64(k2, 33); k2 *= MM_C1; h2 ^= k2;
    }
    if (t > 0) {
        const uint32_t m = t > 8 ? 8 : t;
        for (uint32_t i = m; i > 0; --i) k1 = (k1 << 8) | tail[i - 1];
        k1 *= MM_C1; k1 = rotl64(k1, 31); k1 *= MM_C2; h1 ^= k1;
    }
    h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2;
    return h1;
}

// ---- rolling canonical k-mer state ------------------------------------------------------------
// fwd: base i of the window at bits [2(k-1-i), ..] ("MSB-first") so integer order == byte-
// lexicographic order of the ASCII strings (A<C<G<T  <=>  0<1<2<3).
// rc : reverse complement of the window, also MSB-first.
// Identity used below:  LSB-first(rc) == ~fwd & mask  and  LSB-first(fwd) == ~rc & mask.
struct Roll {
    uint64_t fwd, rc;
    uint32_t run;  // number of consecutive bases ending here (callers walk far fewer than 2^32 symbols)
};
FB2_HD uint64_t kmer_mask(int k) { return k >= 32 ? ~0ULL : ((1ULL << (2 * k)) - 1ULL); }
FB2_HD void roll_push(Roll &r, uint32_t sym, int k, uint64_t mask) {
    const uint64_t c = sym & 3u;
    r.fwd = ((r.fwd << 2) | c) & mask;
    r.rc = (r.rc >> 2) | ((c ^ 3ULL) << (2 * (k - 1)));
    r.run = sym < 4u ? r.run + 1u : 0u;
}
// Canonical choice of needletail's canonical_kmers: fwd < rc ? (fwd,false) : (rc,true);
// a palindrome therefore reports is_rc = true.  Returns LSB-first codes of the chosen k-mer.
FB2_HD uint64_t roll_canonical_lsb(const Roll &r, uint64_t mask, bool &is_rc) {
    is_rc = !(r.fwd < r.rc);
    return (is_rc ? ~r.fwd : ~r.rc) & mask;
}
// LSB-first codes -> k ASCII bytes.
FB2_HD void codes_to_ascii(uint64_t codes, int k, uint8_t *out) {
    for (int i = 0; i < k; ++i) out[i] = (uint8_t)"ACGT"[(codes >> (2 * i)) & 3u];
}

}  // namespace fb2

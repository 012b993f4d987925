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

// ---- murmur3 with the first multiply taken from shared-memory tables ---------------------------
// Every 8-base word w of the k-mer enters murmur3 as  w * c  (c = c1 for k1-type words, c2 for
// k2-type words), w being the 8 ASCII bytes of the bases.  Multiplication distributes over the
// byte groups:  w * c = A4(lo) * c + (A4(hi) * c << 32)  with A4(b) the 4 ASCII bytes of the 4
// bases b.  lut_c[b] = A4(b) * c (64 bit) turns "expand 2-bit codes to ASCII, then multiply" into
// two shared-memory loads and one add, which moves ~20 instructions per k-mer off the ALU pipe.
struct MulLut { uint32_t c1; uint32_t c2; };   // shared-window byte addresses of the two tables

// The kernel is bound by the ALU pipe (LOP3/SHF/IADD3/ISETP); IMAD runs on the FMA pipe.  These
// helpers keep 64-bit additions and table address arithmetic on the FMA pipe.
__device__ __forceinline__ uint64_t add64_fma(uint64_t a, uint64_t b) {
    uint64_t t;
    asm("mad.wide.u32 %0, %1, 1, %2;" : "=l"(t) : "r"((uint32_t)a), "l"(b));       // b + a.lo (64-bit)
    uint32_t hi;
    asm("mad.lo.u32 %0, %1, 1, %2;" : "=r"(hi) : "r"((uint32_t)(a >> 32)), "r"((uint32_t)(t >> 32)));
    return ((uint64_t)hi << 32) | (uint32_t)t;
}
__device__ __forceinline__ uint2 lut_load(uint32_t table, uint32_t idx) {
    uint32_t addr;
    asm("mad.lo.u32 %0, %1, 8, %2;" : "=r"(addr) : "r"(idx), "r"(table));
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}

template <int NBYTES, bool IS_C1>
__device__ __forceinline__ uint64_t mul_word(uint32_t g16, const MulLut &L) {
    // g16: 8 bases (2 bits each, base 0 lowest); NBYTES of them exist, the rest are zero bytes.
    const uint32_t T = IS_C1 ? L.c1 : L.c2;
    const uint64_t C = IS_C1 ? MM_C1 : MM_C2;
    if (NBYTES >= 8) {
        const uint2 a = lut_load(T, g16 & 0xFFu), b = lut_load(T, g16 >> 8);
        uint32_t hi;
        asm("mad.lo.u32 %0, %1, 1, %2;" : "=r"(hi) : "r"(b.x), "r"(a.y));
        return ((uint64_t)hi << 32) | a.x;
    } else if (NBYTES > 4) {
        const uint2 a = lut_load(T, g16 & 0xFFu);
        const uint32_t hi4 = expand4(g16 >> 8) & (uint32_t)low_bytes_mask(NBYTES - 4);
        return ((uint64_t)(a.y + hi4 * (uint32_t)C) << 32) | a.x;
    } else if (NBYTES == 4) {
        const uint2 a = lut_load(T, g16 & 0xFFu);
        return ((uint64_t)a.y << 32) | a.x;
    } else {
        const uint32_t lo4 = expand4(g16 & 0xFFu) & (uint32_t)low_bytes_mask(NBYTES);
        return (uint64_t)lo4 * C;
    }
}

template <int K>
__device__ __forceinline__ uint64_t murmur_kmer_h1_lut(uint64_t codes, uint64_t seed, const MulLut &L) {
    static_assert(K >= 1 && K <= 32, "k out of range");
    uint64_t h1 = seed, h2 = seed;
    constexpr int NB = K / 16, T = K & 15;
    if (NB >= 1) {
        uint64_t k1 = mul_word<8, true>((uint32_t)(codes & 0xFFFFu), L);
        uint64_t k2 = mul_word<8, false>((uint32_t)((codes >> 16) & 0xFFFFu), L);
        k1 = rotl64(k1, 31); k1 *= MM_C2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 = add64_fma(h1, h2); h1 = h1 * 5 + 0x52dce729ULL;
        k2 = rotl64(k2, 33); k2 *= MM_C1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 = add64_fma(h2, h1); h2 = h2 * 5 + 0x38495ab5ULL;
    }
    if (NB >= 2) {
        uint64_t k1 = mul_word<8, true>((uint32_t)((codes >> 32) & 0xFFFFu), L);
        uint64_t k2 = mul_word<8, false>((uint32_t)((codes >> 48) & 0xFFFFu), L);
        k1 = rotl64(k1, 31); k1 *= MM_C2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 = add64_fma(h1, h2); h1 = h1 * 5 + 0x52dce729ULL;
        k2 = rotl64(k2, 33); k2 *= MM_C1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 = add64_fma(h2, h1); h2 = h2 * 5 + 0x38495ab5ULL;
    }
    constexpr int TW = 2 * NB;  // first tail word
    if (T > 8) {
        uint64_t k2 = mul_word<(T > 8 ? T - 8 : 1), false>((uint32_t)((codes >> (16 * ((TW + 1) & 3))) & 0xFFFFu), L);
        k2 = rotl64(k2, 33); k2 *= MM_C1; h2 ^= k2;
    }
    if (T > 0) {
        uint64_t k1 = mul_word<(T > 8 ? 8 : (T > 0 ? T : 1)), true>((uint32_t)((codes >> (16 * (TW & 3))) & 0xFFFFu), L);
        k1 = rotl64(k1, 31); k1 *= MM_C2; h1 ^= k1;
    }
    h1 ^= (uint64_t)K; h2 ^= (uint64_t)K;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2;
    return h1;
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

template <int K>
__global__ void __launch_bounds__(HASH_THREADS)
hash_kernel(const uint8_t *__restrict__ symbuf,   // symbol buffer: region r starts at SYM_FRONT + r * region_stride
            ChunkGeom g, uint32_t b0,             // first hash block (region-major) of this launch
            const uint32_t *__restrict__ region_count, uint64_t ord_base, const SketchState *st,
            LaunchSlot *slot, LogView log, int k_rt, uint64_t seed) {
    __shared__ uint2 lut_c1[256], lut_c2[256];
    if (K > 0) {
        const uint32_t a4 = expand4(threadIdx.x & 0xFFu);
        const uint64_t p1 = (uint64_t)a4 * MM_C1, p2 = (uint64_t)a4 * MM_C2;
        if (threadIdx.x < 256) {
            lut_c1[threadIdx.x] = make_uint2((uint32_t)p1, (uint32_t)(p1 >> 32));
            lut_c2[threadIdx.x] = make_uint2((uint32_t)p2, (uint32_t)(p2 >> 32));
        }
        __syncthreads();
    }
    MulLut L; L.c1 = smem_u32(lut_c1); L.c2 = smem_u32(lut_c2);
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

    Roll r; r.fwd = 0; r.rc = 0; r.run = 0;
    const uint32_t *wp = reinterpret_cast<const uint32_t *>(tile + 32 + t0);
    // ---- warm-up on the 32 symbols before p0 (only the last k-1 matter) -------------------
    if (live) {
        const uint4 a = *reinterpret_cast<const uint4 *>(tile + t0);
        const uint4 b = *reinterpret_cast<const uint4 *>(tile + t0 + 16);
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            if (K > 0 && i < 32 - (K - 1)) continue;  // compile-time skip
            roll_push(r, (w[i >> 2] >> (8 * (i & 3))) & 0xFFu, k, mask);
        }
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
            const uint32_t p = p0 + 4u * (uint32_t)j + (uint32_t)b;
            roll_push(r, (cur >> (8 * b)) & 0xFFu, k, mask);
            const bool ok = r.run >= (uint32_t)k;
            bool is_rc;
            const uint64_t codes = roll_canonical_lsb(r, mask, is_rc);
            uint64_t h;
            if (K > 0) h = murmur_kmer_h1_lut<(K > 0 ? K : 1)>(codes, seed, L);
            else h = murmur_kmer_h1<0>(codes, k, seed);
            nvalid += ok ? 1u : 0u;
            // hot path: compare only the high words (conservative); the exact test is in the branch
            const bool maybe = ok && ((uint32_t)(h >> 32) <= T_hi);
            if (__any_sync(0xffffffffu, maybe)) {
              const bool emit = maybe && (h <= T);
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
                        log.hash[idx] = h;
                        log.kmer[idx] = codes;
                        log.posx[idx] = ((ord_region + p) << 9) | (is_rc ? 1ull : 0ull);
                    }
                }
                res_base += n; res_left -= n;
              }
            }
        }
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

void launch_hash(int k, const uint8_t *symbuf, ChunkGeom g, uint32_t b0, uint32_t b1, const uint32_t *region_count,
                 uint64_t ord_base, const SketchState *st, LaunchSlot *slot, LogView log, uint64_t seed,
                 cudaStream_t stream) {
    if (b1 <= b0) return;
    const uint32_t blocks = b1 - b0;
    if (k == 21) hash_kernel<21><<<blocks, HASH_THREADS, 0, stream>>>(symbuf, g, b0, region_count, ord_base, st, slot, log, k, seed);
    else if (k == 31) hash_kernel<31><<<blocks, HASH_THREADS, 0, stream>>>(symbuf, g, b0, region_count, ord_base, st, slot, log, k, seed);
    else hash_kernel<0><<<blocks, HASH_THREADS, 0, stream>>>(symbuf, g, b0, region_count, ord_base, st, slot, log, k, seed);
}
void launch_push_hash(const uint8_t *bytes, const uint32_t *offs, const uint8_t *extra, uint32_t n,
                      uint64_t arena_base, uint64_t ord_base, const SketchState *st, LaunchSlot *slot, LogView log,
                      uint64_t seed, cudaStream_t stream) {
    if (!n) return;
    push_hash_kernel<<<(n + 255) / 256, 256, 0, stream>>>(bytes, offs, extra, n, arena_base, ord_base, st, slot, log, seed);
    push_commit_kernel<<<1, 1, 0, stream>>>(slot, n);
}

}  // namespace fb2

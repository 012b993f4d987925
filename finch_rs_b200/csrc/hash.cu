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

template <int K>
__global__ void __launch_bounds__(HASH_THREADS)
hash_kernel(const uint8_t *__restrict__ symbuf,   // symbol buffer: region r starts at SYM_FRONT + r * region_stride
            ChunkGeom g, uint32_t b0,             // first hash block (region-major) of this launch
            const uint32_t *__restrict__ region_count, uint64_t ord_base, const SketchState *st,
            LaunchSlot *slot, LogView log, int k_rt, uint64_t seed) {
    const int k = K > 0 ? K : k_rt;
    const uint64_t mask = kmer_mask(k);
    const uint32_t blk = b0 + blockIdx.x;
    const uint32_t region = blk / g.hash_tiles, lt = blk - region * g.hash_tiles;
    const uint32_t end = region_count[region];
    const uint8_t *sym = symbuf + (size_t)SYM_FRONT + (size_t)region * g.region_stride;
    const uint64_t ord_region = ord_base + (uint64_t)region * g.st_bytes;
    const unsigned long long T = st->threshold;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t p0 = lt * HASH_TILE + threadIdx.x * (uint32_t)HASH_W;
    // Warp-uniform early exit: a warp's positions are contiguous and ascending.
    if (__all_sync(0xffffffffu, p0 >= end)) return;

    Roll r; r.fwd = 0; r.rc = 0; r.run = 0;
    const uint32_t *wp = reinterpret_cast<const uint32_t *>(sym + p0);
    // ---- warm-up on the 32 symbols before p0 (only the last k-1 matter) -------------------
    {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(sym + p0) - 2);
        const uint4 b = __ldg(reinterpret_cast<const uint4 *>(sym + p0) - 1);
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            if (K > 0 && i < 32 - (K - 1)) continue;  // compile-time skip
            roll_push(r, (w[i >> 2] >> (8 * (i & 3))) & 0xFFu, k, mask);
        }
    }
    uint32_t nvalid = 0;
    uint32_t word = __ldg(wp);
#pragma unroll 1
    for (int j = 0; j < HASH_W / 4; ++j) {
        const uint32_t cur = word;
        if (j + 1 < HASH_W / 4) word = __ldg(wp + j + 1);  // prefetch next 4 symbols
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const uint32_t p = p0 + 4u * (uint32_t)j + (uint32_t)b;
            roll_push(r, (cur >> (8 * b)) & 0xFFu, k, mask);
            const bool ok = (r.run >= (uint32_t)k) && (p < end);
            bool is_rc;
            const uint64_t codes = roll_canonical_lsb(r, mask, is_rc);
            const uint64_t h = murmur_kmer_h1<K>(codes, k, seed);
            nvalid += ok ? 1u : 0u;
            const bool emit = ok && (h <= T);
            const uint32_t em = __ballot_sync(0xffffffffu, emit);
            if (em) {  // warp-aggregated append
                const int leader = __ffs(em) - 1;
                uint32_t base = 0;
                if ((int)lane == leader) base = atomicAdd(&slot->log_count, (unsigned int)__popc(em));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (emit) {
                    const uint32_t idx = base + __popc(em & lanemask_lt());
                    if (idx < log.cap) {
                        log.hash[idx] = h;
                        log.kmer[idx] = codes;
                        log.posx[idx] = ((ord_region + p) << 9) | (is_rc ? 1ull : 0ull);
                    }
                }
            }
        }
    }
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

// dist.cu -- K6: batched sorted-hash intersection, the integer part of
//   raw_distance (lib/src/distance.rs:66-126): two-pointer merge (:82-95) + scaled tail (:99-115).
// Output per ordered pair: (common, i, j); containment / jaccard / mash distance are f64 and
// are finished on the host (fb2_distance_finish) exactly as distance.rs:117-125 and :35-41 do.
#include <algorithm>
#include "common.cuh"
#include "device_types.cuh"
#include "../../include/finch_b200.h"

namespace fb2 {

// #{x in a[0..n) : x < v}
__device__ __forceinline__ uint32_t lower_bound_u64(const unsigned long long *__restrict__ a, uint32_t n,
                                                    unsigned long long v) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}

// One warp per pair.  Lanes stride over the query hashes and binary-search the reference; the
// consumed counts i, j follow from the closed form of the merge loop:
//   t = min(max A, max B), i = #{a <= t}, j = #{b <= t}; empty list => i = j = 0.
__device__ __forceinline__ void pair_warp(const unsigned long long *__restrict__ A, uint32_t na,
                                          const unsigned long long *__restrict__ B, uint32_t nb,
                                          int scaled, unsigned long long max_hash, fb2_pair_out *out) {
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t common = 0, i = 0, j = 0;
    if (na && nb) {
        for (uint32_t x = lane; x < na; x += 32) {
            const unsigned long long a = A[x];
            const uint32_t p = lower_bound_u64(B, nb, a);
            common += (p < nb && B[p] == a) ? 1u : 0u;
        }
        common = __reduce_add_sync(0xffffffffu, common);
        const unsigned long long ma = A[na - 1], mb = B[nb - 1];
        if (ma <= mb) { i = na; j = (ma == mb) ? nb : lower_bound_u64(B, nb, ma + 1ULL); }
        else { j = nb; i = lower_bound_u64(A, na, mb + 1ULL); }
    }
    if (scaled) {  // distance.rs:99-115: advance while hash < max_hash (strict)
        const uint32_t ia = lower_bound_u64(A, na, max_hash), jb = lower_bound_u64(B, nb, max_hash);
        i = max(i, ia); j = max(j, jb);
    }
    if (lane == 0) { out->common = common; out->i = i; out->j = j; }
}

__global__ void __launch_bounds__(256)
dist_pairs_kernel(const unsigned long long *__restrict__ hashes, const uint32_t *__restrict__ lens, uint32_t stride,
                  const uint32_t *__restrict__ q_idx, const uint32_t *__restrict__ r_idx, uint64_t n_pairs,
                  int scaled, unsigned long long max_hash, fb2_pair_out *__restrict__ out) {
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= n_pairs) return;
    const uint32_t q = q_idx[warp], r = r_idx[warp];
    pair_warp(hashes + (uint64_t)q * stride, lens[q], hashes + (uint64_t)r * stride, lens[r], scaled, max_hash,
              out + warp);
}
__global__ void __launch_bounds__(256)
dist_all_kernel(const unsigned long long *__restrict__ hashes, const uint32_t *__restrict__ lens, uint32_t stride,
                uint32_t n_sk, uint32_t q0, uint64_t n_pairs, int scaled, unsigned long long max_hash,
                fb2_pair_out *__restrict__ out) {
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= n_pairs) return;
    const uint32_t q = q0 + (uint32_t)(warp / n_sk), r = (uint32_t)(warp % n_sk);
    pair_warp(hashes + (uint64_t)q * stride, lens[q], hashes + (uint64_t)r * stride, lens[r], scaled, max_hash,
              out + warp);
}


// ---- tiled all-pairs kernel -------------------------------------------------------------------------
// A block keeps DT_TQ query sketches in shared memory, each as (a) its sorted hash list and (b) an
// open-addressing table of 16-bit entries (6-bit fingerprint | 10-bit index into the sorted list,
// load factor <= 1/8), and streams reference sketches past them: a warp takes one reference, every
// lane probes its 1/32 of the reference's hashes against all DT_TQ tables.  A reference row is read
// once per DT_TQ pairs; a membership test is ~1 shared-memory load instead of a 10-step binary
// search.  Murmur outputs are uniform in their low bits (a bottom-s sketch constrains the top bits
// only), so slot and fingerprint come straight from bits [0,13) and [13,19) of the hash.
// i and j follow from the closed form of the merge loop as in pair_warp.
constexpr int DT_TQ = 9;                 // query sketches per block
constexpr int DT_SLOTS = 8192;           // table slots per query (uint16_t each)
constexpr int DT_MAXLEN = 1023;          // longest query sketch the tiled kernel takes
constexpr int DT_WARPS = 16;
constexpr uint32_t DT_EMPTY = 0xFFFFu;
struct DistTileSmem {
    unsigned long long keys[DT_TQ][DT_MAXLEN + 1];   // sorted hashes; [1023] is never a valid index
    uint16_t tab[DT_TQ][DT_SLOTS];
    unsigned long long maxa[DT_TQ];
    uint32_t na[DT_TQ];
    uint32_t lb[DT_TQ];                               // scaled: #{a < max_hash}
};

__device__ __forceinline__ uint32_t lower_bound_sm(const unsigned long long *a, uint32_t n, unsigned long long v) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}

__global__ void __launch_bounds__(DT_WARPS * 32, 1)
dist_tile_kernel(const unsigned long long *__restrict__ hashes, const uint32_t *__restrict__ lens, uint32_t stride,
                 uint32_t n_sk, uint32_t q0, uint32_t q1, uint32_t r_chunk, int scaled, unsigned long long max_hash,
                 fb2_pair_out *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char dt_raw[];
    DistTileSmem &S = *reinterpret_cast<DistTileSmem *>(dt_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t qb = q0 + blockIdx.x * DT_TQ;
    const uint32_t nq = min((uint32_t)DT_TQ, q1 - qb);
    const uint32_t r_begin = blockIdx.y * r_chunk, r_end = min(n_sk, r_begin + r_chunk);
    // ---- build the query tables ----
    {
        uint32_t *t32 = reinterpret_cast<uint32_t *>(&S.tab[0][0]);
        for (uint32_t i = tid; i < DT_TQ * DT_SLOTS / 2; i += blockDim.x) t32[i] = 0xFFFFFFFFu;
        for (uint32_t q = 0; q < DT_TQ; ++q) {
            const uint32_t len = q < nq ? lens[qb + q] : 0u;
            const unsigned long long *A = hashes + (uint64_t)(qb + (q < nq ? q : 0u)) * stride;
            for (uint32_t x = tid; x <= DT_MAXLEN; x += blockDim.x) S.keys[q][x] = x < len ? A[x] : 0ULL;
            if (tid == 0) {
                S.na[q] = len;
                S.maxa[q] = len ? A[len - 1] : 0ULL;
                S.lb[q] = scaled ? lower_bound_u64(A, len, max_hash) : 0u;
            }
        }
        __syncthreads();
        for (uint32_t q = 0; q < nq; ++q) {
            const uint32_t len = S.na[q];
            for (uint32_t x = tid; x < len; x += blockDim.x) {
                const unsigned long long h = S.keys[q][x];
                uint32_t s = (uint32_t)h & (DT_SLOTS - 1);
                const unsigned short val = (unsigned short)(((((uint32_t)h >> 13) & 63u) << 10) | x);
                while (atomicCAS(&S.tab[q][s], (unsigned short)DT_EMPTY, val) != (unsigned short)DT_EMPTY)
                    s = (s + 1) & (DT_SLOTS - 1);
            }
        }
        __syncthreads();
    }
    // ---- stream the references ----
    for (uint32_t r = r_begin + warp; r < r_end; r += DT_WARPS) {
        const uint32_t nb = lens[r];
        const unsigned long long *__restrict__ B = hashes + (uint64_t)r * stride;
        uint32_t cnt[DT_TQ];
#pragma unroll
        for (int q = 0; q < DT_TQ; ++q) cnt[q] = 0;
        for (uint32_t x = lane; x < nb; x += 32) {
            const unsigned long long b = B[x];
            const uint32_t s0 = (uint32_t)b & (DT_SLOTS - 1), fp = ((uint32_t)b >> 13) & 63u;
#pragma unroll
            for (int q = 0; q < DT_TQ; ++q) {
                uint32_t s = s0;
                while (true) {
                    const uint32_t e = S.tab[q][s];
                    if (e == DT_EMPTY) break;
                    if ((e >> 10) == fp && S.keys[q][e & 1023u] == b) { ++cnt[q]; break; }
                    s = (s + 1) & (DT_SLOTS - 1);
                }
            }
        }
        uint32_t mine = 0;
#pragma unroll
        for (int q = 0; q < DT_TQ; ++q) {
            const uint32_t c = __reduce_add_sync(0xffffffffu, cnt[q]);
            if (lane == (uint32_t)q) mine = c;
        }
        if (lane < nq) {
            const uint32_t q = lane, na = S.na[q];
            uint32_t i = 0, j = 0;
            if (na && nb) {
                const unsigned long long ma = S.maxa[q], mb = B[nb - 1];
                if (ma <= mb) { i = na; j = (ma == mb) ? nb : lower_bound_u64(B, nb, ma + 1ULL); }
                else { j = nb; i = lower_bound_sm(S.keys[q], na, mb + 1ULL); }
            } else mine = 0;
            if (scaled) { i = max(i, S.lb[q]); j = max(j, lower_bound_u64(B, nb, max_hash)); }
            fb2_pair_out o; o.common = mine; o.i = i; o.j = j;
            out[(uint64_t)(qb + q - q0) * n_sk + r] = o;
        }
    }
}

void launch_dist_pairs(const unsigned long long *hashes, const uint32_t *lens, uint32_t stride, const uint32_t *q_idx,
                       const uint32_t *r_idx, uint64_t n_pairs, int scaled, unsigned long long max_hash,
                       fb2_pair_out *out, cudaStream_t s) {
    if (!n_pairs) return;
    const uint64_t blocks = (n_pairs * 32 + 255) / 256;
    dist_pairs_kernel<<<(unsigned)blocks, 256, 0, s>>>(hashes, lens, stride, q_idx, r_idx, n_pairs, scaled, max_hash, out);
}
void launch_dist_all(const unsigned long long *hashes, const uint32_t *lens, uint32_t stride, uint32_t n_sk,
                     uint32_t q0, uint64_t n_pairs, int scaled, unsigned long long max_hash, fb2_pair_out *out,
                     cudaStream_t s) {
    if (!n_pairs) return;
    const uint64_t blocks = (n_pairs * 32 + 255) / 256;
    dist_all_kernel<<<(unsigned)blocks, 256, 0, s>>>(hashes, lens, stride, n_sk, q0, n_pairs, scaled, max_hash, out);
}

uint32_t dist_tile_max_len() { return DT_MAXLEN; }
// All ordered pairs (q, r), q in [q0, q1), r in [0, n_sk); every query sketch must be <= dist_tile_max_len() long.
int launch_dist_tile(const unsigned long long *hashes, const uint32_t *lens, uint32_t stride, uint32_t n_sk, uint32_t q0,
                     uint32_t q1, int scaled, unsigned long long max_hash, fb2_pair_out *out, cudaStream_t s) {
    if (q1 <= q0 || !n_sk) return 0;
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        if (cudaFuncSetAttribute(dist_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DistTileSmem)) != cudaSuccess)
            return -1;
        attr_set[dev] = true;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t q_tiles = (q1 - q0 + DT_TQ - 1) / DT_TQ;
    // split the references so that the grid fills the chip a few times over, >= 64 references per block
    uint32_t splits = std::max<uint32_t>(1u, (uint32_t)(4 * sms + q_tiles - 1) / q_tiles);
    splits = std::min<uint32_t>(splits, std::max<uint32_t>(1u, n_sk / 64u));
    splits = std::min<uint32_t>(splits, 65535u);
    const uint32_t r_chunk = (n_sk + splits - 1) / splits;
    const dim3 grid(q_tiles, (n_sk + r_chunk - 1) / r_chunk);
    dist_tile_kernel<<<grid, DT_WARPS * 32, sizeof(DistTileSmem), s>>>(hashes, lens, stride, n_sk, q0, q1, r_chunk, scaled,
                                                                    max_hash, out);
    return 0;
}

}  // namespace fb2

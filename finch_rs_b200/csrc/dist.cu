// dist.cu -- K6: batched sorted-hash intersection, the integer part of
//   raw_distance (lib/src/distance.rs:66-126): two-pointer merge (:82-95) + scaled tail (:99-115).
// Output per ordered pair: (common, i, j); containment / jaccard / mash distance are f64 and
// are finished on the host (fb2_distance_finish) exactly as distance.rs:117-125 and :35-41 do.
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

}  // namespace fb2

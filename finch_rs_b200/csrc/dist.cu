// dist.cu -- K6: batched sorted-hash intersection, the integer part of
//   raw_distance (lib/src/distance.rs:66-126): two-pointer merge (:82-95) + scaled tail (:99-115).
// Output per ordered pair: (common, i, j); containment / jaccard / mash distance are f64 and
// are finished on the host (fb2_distance_finish) exactly as distance.rs:117-125 and :35-41 do.
#include <algorithm>
#include <cstdlib>
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
// A block keeps DT_TQ query sketches in shared memory, each as (a) its sorted hash list and (b) a
// CUCKOO table of 32-bit entries (22-bit fingerprint | 10-bit index into the sorted list; two hash
// functions, one slot each, load <= 1/4), and streams reference sketches past them: a warp takes one
// reference, every lane probes its 1/32 of the reference's hashes against all DT_TQ tables.  A
// membership test is exactly two independent shared-memory loads (no probe chains, so the lanes of a
// warp never wait for the longest chain among them); the key itself is only read when a fingerprint
// matches.  A reference row is read once per DT_TQ pairs.
// Murmur outputs are uniform in every bit slice (a bottom-s sketch constrains the top bits only):
// slot 1 = bits [0,12), slot 2 = bits [12,24), fingerprint = bits [24,46).  A query whose keys cannot be
// placed (adversarial, non-random hashes) is flagged and answered by binary search instead.
// i and j follow from the closed form of the merge loop as in pair_warp.
constexpr int DT_TQ = 9;                 // query sketches per block
constexpr int DT_SLOTS = 4096;           // cuckoo slots per query (uint32_t each)
constexpr int DT_MAXLEN = 1023;          // longest query sketch the tiled kernel takes
constexpr int DT_WARPS = 16;
constexpr int DT_MAXKICK = 96;
constexpr uint32_t DT_EMPTY = 0xFFFFFFFFu;   // fingerprint all ones, index 1023 (never a valid index)
struct DistTileSmem {
    unsigned long long keys[DT_TQ][DT_MAXLEN + 1];   // sorted hashes; [1023] = 0: what an EMPTY entry points at
    uint32_t tab[DT_TQ][DT_SLOTS];
    unsigned long long maxa[DT_TQ];
    uint32_t na[DT_TQ];
    uint32_t lb[DT_TQ];                               // scaled: #{a < max_hash}
    uint32_t slow_mask;                               // queries whose cuckoo build failed
};
__device__ __forceinline__ uint32_t dt_h1(unsigned long long k) { return (uint32_t)k & (DT_SLOTS - 1); }
__device__ __forceinline__ uint32_t dt_h2(unsigned long long k) { return (uint32_t)(k >> 12) & (DT_SLOTS - 1); }
__device__ __forceinline__ uint32_t dt_fp(unsigned long long k) { return (uint32_t)(k >> 24) & 0x3FFFFFu; }

__device__ __forceinline__ uint32_t lower_bound_sm(const unsigned long long *a, uint32_t n, unsigned long long v) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}

// CUT mode (`finch dist --max-dist`, cli/src/main.rs:326-330 keeps a pair iff mash_distance <= max_dist): instead of
// the dense (q, r) matrix the kernel appends only the pairs whose jaccard reaches `jlow` -- a bound placed slightly
// BELOW the exact one, so no pair the host's exact f64 test would keep is lost; the host re-checks every hit -- to a
// compact list with one atomic per warp.  10^10 pairs cannot be returned densely; their hits can.
struct DistCut {
    fb2_pair_hit *hits;            // compacted survivors (unordered)
    unsigned long long *keys;      // (q << 32) | r per hit, for the sort that follows
    unsigned int *counter;         // hits appended so far (may run past cap: the host then retries with fewer rows)
    uint32_t cap;
    int skip_self;                 // drop q == r
    double jlow;                   // keep iff jaccard >= jlow (negative: keep all)
};
template <bool CUT>
__global__ void __launch_bounds__(DT_WARPS * 32, 1)
dist_tile_kernel(const unsigned long long *__restrict__ hashes, const uint32_t *__restrict__ lens, uint32_t stride,
                 uint32_t n_sk, uint32_t q0, uint32_t q1, uint32_t r_chunk, int scaled, unsigned long long max_hash,
                 fb2_pair_out *__restrict__ out, DistCut cut) {
    extern __shared__ __align__(16) unsigned char dt_raw[];
    DistTileSmem &S = *reinterpret_cast<DistTileSmem *>(dt_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t qb = q0 + blockIdx.x * DT_TQ;
    const uint32_t nq = min((uint32_t)DT_TQ, q1 - qb);
    const uint32_t r_begin = blockIdx.y * r_chunk, r_end = min(n_sk, r_begin + r_chunk);
    // ---- build the query tables ----
    {
        uint32_t *t32 = &S.tab[0][0];
        for (uint32_t i = tid; i < DT_TQ * DT_SLOTS; i += blockDim.x) t32[i] = DT_EMPTY;
        if (tid == 0) S.slow_mask = 0;
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
                // cuckoo insertion: swap into the slot, carry the evicted entry to its other slot
                unsigned long long k = S.keys[q][x];
                uint32_t cur = (dt_fp(k) << 10) | x, slot = dt_h1(k);
                int kick = 0;
                for (; kick < DT_MAXKICK; ++kick) {
                    const uint32_t old = atomicExch(&S.tab[q][slot], cur);
                    if (old == DT_EMPTY) break;
                    cur = old;
                    k = S.keys[q][old & 1023u];
                    slot = (slot == dt_h1(k)) ? dt_h2(k) : dt_h1(k);
                }
                if (kick == DT_MAXKICK) atomicOr(&S.slow_mask, 1u << q);   // an entry was dropped: table unusable
            }
        }
        __syncthreads();
    }
    const uint32_t slow_mask = S.slow_mask;
    // ---- stream the references ----
    for (uint32_t r = r_begin + warp; r < r_end; r += DT_WARPS) {
        const uint32_t nb = lens[r];
        const unsigned long long *__restrict__ B = hashes + (uint64_t)r * stride;
        uint32_t cnt[DT_TQ];
#pragma unroll
        for (int q = 0; q < DT_TQ; ++q) cnt[q] = 0;
        for (uint32_t x = lane; x < nb; x += 32) {
            const unsigned long long b = B[x];
            const uint32_t s1 = dt_h1(b), s2 = dt_h2(b), fp = dt_fp(b);
            uint32_t e1[DT_TQ], e2[DT_TQ];
#pragma unroll
            for (int q = 0; q < DT_TQ; ++q) { e1[q] = S.tab[q][s1]; e2[q] = S.tab[q][s2]; }   // 2 * DT_TQ loads in flight
#pragma unroll
            for (int q = 0; q < DT_TQ; ++q) {
                const bool m1 = (e1[q] >> 10) == fp, m2 = (e2[q] >> 10) == fp;
                if ((m1 || m2) && !((slow_mask >> q) & 1u)) {   // rare unless b is a member: verify against the key itself
                    const bool hit = (m1 && S.keys[q][e1[q] & 1023u] == b) || (m2 && S.keys[q][e2[q] & 1023u] == b);
                    cnt[q] += hit ? 1u : 0u;
                }
            }
            if (slow_mask) {      // block-uniform; adversarial hash sets only
#pragma unroll
                for (int q = 0; q < DT_TQ; ++q)
                    if ((slow_mask >> q) & 1u) {
                        const uint32_t na = S.na[q], p = lower_bound_sm(S.keys[q], na, b);
                        cnt[q] = cnt[q] + ((p < na && S.keys[q][p] == b) ? 1u : 0u);
                    }
            }
        }
        uint32_t mine = 0;
#pragma unroll
        for (int q = 0; q < DT_TQ; ++q) {
            const uint32_t c = __reduce_add_sync(0xffffffffu, cnt[q]);
            if (lane == (uint32_t)q) mine = c;
        }
        bool pass = false;
        fb2_pair_out o; o.common = 0; o.i = 0; o.j = 0;
        if (lane < nq) {
            const uint32_t q = lane, na = S.na[q];
            uint32_t i = 0, j = 0;
            if (na && nb) {
                const unsigned long long ma = S.maxa[q], mb = B[nb - 1];
                if (ma <= mb) { i = na; j = (ma == mb) ? nb : lower_bound_u64(B, nb, ma + 1ULL); }
                else { j = nb; i = lower_bound_sm(S.keys[q], na, mb + 1ULL); }
            } else mine = 0;
            if (scaled) { i = max(i, S.lb[q]); j = max(j, lower_bound_u64(B, nb, max_hash)); }
            o.common = mine; o.i = i; o.j = j;
            if (!CUT) out[(uint64_t)(qb + q - q0) * n_sk + r] = o;
            else {
                const uint32_t total = i - mine + j;
                const double jac = total == 0u ? 1.0 : (double)mine / (double)total;      // distance.rs:119-125
                pass = !(cut.skip_self && qb + q == r) && (cut.jlow < 0.0 || jac >= cut.jlow);
            }
        }
        if (CUT) {
            const uint32_t m = __ballot_sync(0xffffffffu, pass);
            if (m) {
                uint32_t base = 0;
                if (lane == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(cut.counter, (unsigned int)__popc(m));
                base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
                if (pass) {
                    const uint32_t idx = base + __popc(m & ((1u << lane) - 1u));
                    if (idx < cut.cap) {
                        fb2_pair_hit h; h.q = qb + lane; h.r = r; h.common = o.common; h.i = o.i; h.j = o.j;
                        cut.hits[idx] = h;
                        cut.keys[idx] = ((unsigned long long)(qb + lane) << 32) | r;
                    }
                }
            }
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
static int dist_tile_launch(const unsigned long long *hashes, const uint32_t *lens, uint32_t stride, uint32_t n_sk, uint32_t q0,
                            uint32_t q1, int scaled, unsigned long long max_hash, fb2_pair_out *out, const DistCut *cut,
                            cudaStream_t s) {
    if (q1 <= q0 || !n_sk) return 0;
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        if (cudaFuncSetAttribute(dist_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DistTileSmem)) != cudaSuccess ||
            cudaFuncSetAttribute(dist_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DistTileSmem)) != cudaSuccess)
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
    if (cut) dist_tile_kernel<true><<<grid, DT_WARPS * 32, sizeof(DistTileSmem), s>>>(hashes, lens, stride, n_sk, q0, q1, r_chunk, scaled,
                                                                                   max_hash, nullptr, *cut);
    else dist_tile_kernel<false><<<grid, DT_WARPS * 32, sizeof(DistTileSmem), s>>>(hashes, lens, stride, n_sk, q0, q1, r_chunk, scaled,
                                                                                max_hash, out, DistCut{});
    return 0;
}
int launch_dist_tile(const unsigned long long *hashes, const uint32_t *lens, uint32_t stride, uint32_t n_sk, uint32_t q0,
                     uint32_t q1, int scaled, unsigned long long max_hash, fb2_pair_out *out, cudaStream_t s) {
    return dist_tile_launch(hashes, lens, stride, n_sk, q0, q1, scaled, max_hash, out, nullptr, s);
}
// CUT mode: survivors of rows [q0, q1) appended to hits / keys (see DistCut).
int launch_dist_tile_cut(const unsigned long long *hashes, const uint32_t *lens, uint32_t stride, uint32_t n_sk, uint32_t q0,
                         uint32_t q1, int scaled, unsigned long long max_hash, fb2_pair_hit *hits, unsigned long long *keys,
                         unsigned int *counter, uint32_t cap, int skip_self, double jlow, cudaStream_t s) {
    DistCut c; c.hits = hits; c.keys = keys; c.counter = counter; c.cap = cap; c.skip_self = skip_self; c.jlow = jlow;
    return dist_tile_launch(hashes, lens, stride, n_sk, q0, q1, scaled, max_hash, nullptr, &c, s);
}
// ---- the cut through an inverted index ------------------------------------------------------------------------
// calc_sketch_distances keeps a pair iff mash_distance <= max_dist (cli/src/main.rs:326-330).  With a positive jaccard
// bound a pair that shares NO hash cannot pass (jaccard 0; empty sketches aside), and in a large collection almost no
// pair shares one: C5 holds 10^10 ordered pairs, 10^7 of them inside its 1000 clusters.  The tile kernel above still
// probes every pair (1000 membership tests each).  Here the collection is turned into postings (hash, sketch), sorted
// by hash; a block then owns ONE query sketch, walks the posting run of each of its hashes and counts, per reference
// sketch, the hashes they share in 16-bit counters in shared memory (2 bytes x sketches of a column block).  Only
// references with a non-zero count are candidates; their (common, i, j) are exact (i, j from the closed form of the
// merge loop, like pair_warp) and go through the same jaccard test and hit list as the tile kernel's.
// Work: sum over distinct hashes of (run length)^2 counter bumps instead of pairs x hashes probes.

// one warp per sketch: keys[off[sk] + p] = its p-th hash, vals[...] = off[sk] + p (the posting's own index)
__global__ void __launch_bounds__(256)
postings_fill_kernel(const unsigned long long *__restrict__ hashes, const uint32_t *__restrict__ lens, uint32_t stride,
                     uint32_t n_sk, const uint32_t *__restrict__ off, unsigned long long *__restrict__ keys,
                     uint32_t *__restrict__ vals) {
    const uint32_t sk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (sk >= n_sk) return;
    const uint32_t n = lens[sk], base = off[sk];
    const unsigned long long *A = hashes + (uint64_t)sk * stride;
    for (uint32_t p = lane; p < n; p += 32) { keys[base + p] = A[p]; vals[base + p] = base + p; }
}
// After the (stable) sort by hash: sorted_sk[e] = sketch of posting e (ascending inside a run); the first posting of
// every run tells each member where the run lies: runinfo[posting index] = start | length << 32.
__global__ void __launch_bounds__(256)
postings_runs_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ vals, uint32_t n,
                     const uint32_t *__restrict__ off, uint32_t n_sk, uint32_t *__restrict__ sorted_sk,
                     unsigned long long *__restrict__ runinfo, unsigned long long *__restrict__ sum_sq) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long sq = 0;
    if (e < n) {
        const uint32_t v = vals[e];
        uint32_t lo = 0, hi = n_sk;                       // sk = last index with off[sk] <= v
        while (hi - lo > 1u) { const uint32_t mid = (lo + hi) >> 1; if (off[mid] <= v) lo = mid; else hi = mid; }
        sorted_sk[e] = lo;
        const unsigned long long k = keys[e];
        if (e == 0u || keys[e - 1] != k) {
            uint32_t t = e + 1u;
            while (t < n && keys[t] == k) ++t;
            const unsigned long long info = (unsigned long long)e | ((unsigned long long)(t - e) << 32);
            for (uint32_t x = e; x < t; ++x) runinfo[vals[x]] = info;
            sq = (unsigned long long)(t - e) * (unsigned long long)(t - e);
        }
    }
    // block sum of the squared run lengths (the cost of the counting pass)
    __shared__ unsigned long long part[8];
    for (int d = 16; d > 0; d >>= 1) sq += __shfl_down_sync(0xffffffffu, sq, d);
    if ((threadIdx.x & 31u) == 0u) part[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long s = 0;
        for (int w = 0; w < 8; ++w) s += part[w];
        if (s) atomicAdd(sum_sq, s);
    }
}
constexpr int DI_THREADS = 1024;
__global__ void __launch_bounds__(DI_THREADS, 1)
dist_inverted_kernel(const unsigned long long *__restrict__ hashes, const uint32_t *__restrict__ lens, uint32_t stride,
                     uint32_t n_sk, uint32_t q0, const uint32_t *__restrict__ off, const uint32_t *__restrict__ sorted_sk,
                     const unsigned long long *__restrict__ runinfo, uint32_t cb, int scaled, unsigned long long max_hash,
                     DistCut cut) {
    extern __shared__ uint32_t di_cnt[];                  // cb / 2 words: two 16-bit counters each
    const uint32_t q = q0 + blockIdx.x, tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const unsigned long long *A = hashes + (uint64_t)q * stride;
    const uint32_t na = lens[q], base = off[q], words = cb >> 1;
    for (uint32_t c0 = 0; c0 < n_sk; c0 += cb) {
        for (uint32_t x = tid; x < words; x += DI_THREADS) di_cnt[x] = 0u;
        __syncthreads();
        for (uint32_t p = warp; p < na; p += DI_THREADS / 32) {
            const unsigned long long info = runinfo[base + p];
            const uint32_t start = (uint32_t)info, len = (uint32_t)(info >> 32);
            for (uint32_t x = lane; x < len; x += 32) {
                const uint32_t r = sorted_sk[start + x] - c0;        // (unsigned: sketches below c0 wrap past cb)
                if (r < cb) atomicAdd(&di_cnt[r >> 1], 1u << (16u * (r & 1u)));
            }
        }
        __syncthreads();
        for (uint32_t x = tid; x < words; x += DI_THREADS) {
            const uint32_t w = di_cnt[x];
            if (!w) continue;
#pragma unroll
            for (uint32_t half = 0; half < 2; ++half) {
                const uint32_t c = (w >> (16u * half)) & 0xFFFFu, r = c0 + 2u * x + half;
                if (!c || r >= n_sk) continue;
                const unsigned long long *B = hashes + (uint64_t)r * stride;
                const uint32_t nb = lens[r];
                uint32_t i, j;                                       // na, nb > 0: the sketches share a hash
                const unsigned long long ma = A[na - 1], mb = B[nb - 1];
                if (ma <= mb) { i = na; j = (ma == mb) ? nb : lower_bound_u64(B, nb, ma + 1ULL); }
                else { j = nb; i = lower_bound_u64(A, na, mb + 1ULL); }
                if (scaled) { i = max(i, lower_bound_u64(A, na, max_hash)); j = max(j, lower_bound_u64(B, nb, max_hash)); }
                const uint32_t total = i - c + j;
                const double jac = total == 0u ? 1.0 : (double)c / (double)total;      // distance.rs:119-125
                if ((cut.skip_self && q == r) || !(cut.jlow < 0.0 || jac >= cut.jlow)) continue;
                const uint32_t idx = atomicAdd(cut.counter, 1u);
                if (idx < cut.cap) {
                    fb2_pair_hit h; h.q = q; h.r = r; h.common = c; h.i = i; h.j = j;
                    cut.hits[idx] = h;
                    cut.keys[idx] = ((unsigned long long)q << 32) | r;
                }
            }
        }
        __syncthreads();
    }
}
void launch_postings_fill(const unsigned long long *hashes, const uint32_t *lens, uint32_t stride, uint32_t n_sk,
                          const uint32_t *off, unsigned long long *keys, uint32_t *vals, cudaStream_t s) {
    if (!n_sk) return;
    postings_fill_kernel<<<(unsigned)(((uint64_t)n_sk * 32 + 255) / 256), 256, 0, s>>>(hashes, lens, stride, n_sk, off, keys, vals);
}
void launch_postings_runs(const unsigned long long *keys, const uint32_t *vals, uint32_t n, const uint32_t *off, uint32_t n_sk,
                          uint32_t *sorted_sk, unsigned long long *runinfo, unsigned long long *sum_sq, cudaStream_t s) {
    if (!n) return;
    postings_runs_kernel<<<(n + 255u) / 256u, 256, 0, s>>>(keys, vals, n, off, n_sk, sorted_sk, runinfo, sum_sq);
}
// Column block of the counting kernel: as many sketches as 16-bit counters fit into the shared memory of an SM.
uint32_t dist_inverted_block(uint32_t n_sk) {
    const uint32_t most = 112u * 1024u;                   // 224 KiB of counters
    if (const char *e = getenv("FB2_DIST_CB")) { const long v = atol(e); if (v >= 2 && v <= (long)most) return (uint32_t)v & ~1u; }
    return std::min(most, (n_sk + 1u) & ~1u);
}
// Rows [q0, q1): survivors appended to hits / keys like launch_dist_tile_cut.  No sketch may be empty or longer than
// 65535 hashes, jlow must be positive (the caller checks).  Returns -1 when the shared memory cannot be reserved.
int launch_dist_inverted_cut(const unsigned long long *hashes, const uint32_t *lens, uint32_t stride, uint32_t n_sk, uint32_t q0,
                             uint32_t q1, const uint32_t *off, const uint32_t *sorted_sk, const unsigned long long *runinfo,
                             int scaled, unsigned long long max_hash, fb2_pair_hit *hits, unsigned long long *keys,
                             unsigned int *counter, uint32_t cap, int skip_self, double jlow, cudaStream_t s) {
    if (q1 <= q0 || !n_sk) return 0;
    const uint32_t cb = dist_inverted_block(n_sk);
    const size_t smem = (size_t)cb * 2u;
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        if (cudaFuncSetAttribute(dist_inverted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024) != cudaSuccess) return -1;
        attr_set[dev] = true;
    }
    DistCut c; c.hits = hits; c.keys = keys; c.counter = counter; c.cap = cap; c.skip_self = skip_self; c.jlow = jlow;
    dist_inverted_kernel<<<q1 - q0, DI_THREADS, smem, s>>>(hashes, lens, stride, n_sk, q0, off, sorted_sk, runinfo, cb, scaled, max_hash, c);
    return 0;
}
// hits[order[i]] -> sorted[i]
__global__ void gather_hits_kernel(const fb2_pair_hit *__restrict__ hits, const uint32_t *__restrict__ order, uint32_t n,
                                   fb2_pair_hit *__restrict__ sorted) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sorted[i] = hits[order[i]];
}
__global__ void iota_kernel(uint32_t *v, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}
void launch_iota(uint32_t *v, uint32_t n, cudaStream_t s) { if (n) iota_kernel<<<(n + 255) / 256, 256, 0, s>>>(v, n); }
void launch_gather_hits(const fb2_pair_hit *hits, const uint32_t *order, uint32_t n, fb2_pair_hit *sorted, cudaStream_t s) {
    if (n) gather_hits_kernel<<<(n + 255) / 256, 256, 0, s>>>(hits, order, n, sorted);
}

// ---- minmer_matrix (lib/src/distance.rs:344-364) ---------------------------------------------------------------------
// result[i][c] = the count sketch i holds for reference hash c (0 when it does not hold it).  The reference walks a
// pointer along the reference for every sketch; for ascending, distinct hashes that is a membership test, done here
// per (sketch, hash) with a binary search of the reference list.
__global__ void minmer_matrix_kernel(const unsigned long long *__restrict__ ref, uint32_t n_ref,
                                     const unsigned long long *__restrict__ sk_hash, const uint32_t *__restrict__ sk_cnt,
                                     const unsigned long long *__restrict__ sk_off, uint32_t n_sk, int32_t *__restrict__ result) {
    const uint32_t i = blockIdx.y;
    if (i >= n_sk) return;
    const unsigned long long a = sk_off[i], b = sk_off[i + 1];
    for (unsigned long long j = a + blockIdx.x * blockDim.x + threadIdx.x; j < b; j += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long h = sk_hash[j];
        uint32_t lo = 0, hi = n_ref;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (ref[mid] < h) lo = mid + 1; else hi = mid; }
        if (lo < n_ref && ref[lo] == h) result[(size_t)i * n_ref + lo] = (int32_t)sk_cnt[j];
    }
}
void launch_minmer_matrix(const unsigned long long *ref, uint32_t n_ref, const unsigned long long *sk_hash, const uint32_t *sk_cnt,
                          const unsigned long long *sk_off, uint32_t n_sk, uint32_t max_len, int32_t *result, cudaStream_t s) {
    if (!n_sk || !n_ref || !max_len) return;
    const dim3 grid(std::min<uint32_t>((max_len + 255u) / 256u, 64u), n_sk);
    minmer_matrix_kernel<<<grid, 256, 0, s>>>(ref, n_ref, sk_hash, sk_cnt, sk_off, n_sk, result);
}

}  // namespace fb2

// table.cu -- K4/K5: candidate log -> distinct-hash table with count / extra_count / first k-mer,
// pruning to the bottom-s, and the final ascending export.
//
// Replaces the state behind MashSketcher::push / ScaledSketcher::push
//   BinaryHeap<HashedItem<Vec<u8>>> + HashMap<ItemHash,(u32,u32)>   (mash.rs:10-18,43-61; scaled.rs:41-60)
// and to_vec (mash.rs:86-102 / scaled.rs:84-100) by their closed forms:
//   Mash(s):          the s smallest distinct hashes of the whole input with their full totals
//   Scaled(s, m):     the max(|{h <= m}|, s) smallest (just {h <= m} when s == 0)
// The admission threshold only ever decreases, so a key that survives to the result was never
// rejected or purged and its totals are complete.
#include <algorithm>
#include <cstdlib>
#include "common.cuh"
#include "device_types.cuh"

namespace fb2 {

__device__ __forceinline__ uint32_t home_slot(unsigned long long key, uint32_t shift) {
    return (uint32_t)((key * 0x9E3779B97F4A7C15ULL) >> shift);
}

// Find-or-insert `key`; returns its slot.  The table always has free slots (host guarantees
// occupied + batch <= 3/4 cap), so the probe terminates.
// `inserted` is set when the key is new: the caller adds the warp's new keys to st->occupied with ONE
// atomic (commit_inserted) -- a per-key atomic on that single address serialises large absorbs.
__device__ __forceinline__ uint32_t table_upsert(const TableView &t, unsigned long long key, SketchState *st, bool &inserted) {
    inserted = false;
    if (key == EMPTY_KEY) {                       // u64::MAX cannot be stored in a slot: side slot
        if (atomicExch(&st->has_max_key, 1u) == 0u) { /* first use */ }
        return t.cap;
    }
    uint32_t slot = home_slot(key, t.shift);
    const uint32_t maskc = t.cap - 1u;
    while (true) {
        unsigned long long cur = t.key[slot];
        if (cur == key) return slot;
        if (cur == EMPTY_KEY) {
            const unsigned long long prev = atomicCAS(&t.key[slot], EMPTY_KEY, key);
            if (prev == EMPTY_KEY) {
                inserted = true;
                // live histogram of the table's keys (soft threshold updates between rebuilds)
                unsigned int *lb = st->live_bins;
                if (lb) atomicAdd(&lb[min(4095ULL, key >> st->hist_shift)], 1u);
                return slot;
            }
            if (prev == key) return slot;
        }
        slot = (slot + 1u) & maskc;
    }
}
// Called by every lane of a (converged) warp: one atomic for all the warp's new keys.
__device__ __forceinline__ void commit_inserted(SketchState *st, bool inserted) {
    const uint32_t m = __ballot_sync(0xffffffffu, inserted);
    if (m && (threadIdx.x & 31u) == (uint32_t)(__ffs(m) - 1)) atomicAdd(&st->occupied, (unsigned int)__popc(m));
}
__device__ __forceinline__ uint32_t table_find(const TableView &t, unsigned long long key) {
    if (key == EMPTY_KEY) return t.cap;
    uint32_t slot = home_slot(key, t.shift);
    const uint32_t maskc = t.cap - 1u;
    while (true) {
        const unsigned long long cur = t.key[slot];
        if (cur == key) return slot;
        if (cur == EMPTY_KEY) return 0xFFFFFFFFu;
        slot = (slot + 1u) & maskc;
    }
}

// k-mer codes are kw 64-bit words per entry (kw = 1 for k <= 32: the hot configuration keeps its single-word copy)
__device__ __forceinline__ void copy_kmer(unsigned long long *dst, const unsigned long long *src, uint32_t kw) {
    dst[0] = src[0];
    for (uint32_t w = 1; w < kw; ++w) dst[w] = src[w];
}

// Pass 1: counts, strand counts and the minimum position per key.  Entries above the current
// threshold (logged under an older, larger threshold) are dropped.
// Optional hash band (lo, hi]: only entries inside it are absorbed (banded absorb of large logs).
struct Band { unsigned long long lo, hi; int use_lo; };
__device__ __forceinline__ bool in_band(const Band &b, unsigned long long key) {
    return key <= b.hi && (!b.use_lo || key > b.lo);
}
__global__ void absorb_count_kernel(LogView log, uint32_t i0, uint32_t i1, TableView t, SketchState *st, Band band) {
    const uint32_t i = i0 + blockIdx.x * blockDim.x + threadIdx.x;   // no early return: the warp votes below
    unsigned long long px = ~0ULL, key = 0;
    bool valid = i < i1;
    if (valid) { px = log.posx[i]; valid = px != ~0ULL; }            // ~0: unused slot of a warp's reservation
    if (valid) { key = log.hash[i]; valid = key <= st->threshold && in_band(band, key); }
    bool ins = false;
    uint32_t slot = 0;
    if (valid) slot = table_upsert(t, key, st, ins);
    commit_inserted(st, ins);
    if (!valid) return;
    atomicAdd(&t.cnt[slot], 1ULL);
    const unsigned long long extra = px & 0xFFULL;
    if (extra) atomicAdd(&t.ext[slot], extra);
    if (px < __ldcg(&t.posx[slot])) atomicMin(&t.posx[slot], px);   // stored value only decreases: a stale read is safe
}
// Pass 2: the occurrence that owns the minimum position donates the k-mer (mash.rs:52-55 keeps
// the k-mer of the first push of a hash).
__global__ void absorb_kmer_kernel(LogView log, uint32_t i0, uint32_t i1, TableView t, const SketchState *st, Band band) {
    const uint32_t i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= i1) return;
    const unsigned long long px = log.posx[i];
    if (px == ~0ULL) return;
    const unsigned long long key = log.hash[i];
    if (key > st->threshold || !in_band(band, key)) return;
    const uint32_t slot = table_find(t, key);
    if (slot == 0xFFFFFFFFu) return;
    if (t.posx[slot] == px) copy_kmer(t.kmer + (size_t)slot * t.kw, log.kmer + (size_t)i * log.kw, t.kw);
}

// ---- device-decided absorb (asynchronous chunks) ---------------------------------------------------
// One thread decides whether the slot's log may be absorbed: the log must not have overflowed and
// the table must stay under 3/4 load even if every entry is a new key.  The host reads the decision
// one chunk later and redoes / prunes in the (rare) other cases.
__global__ void absorb_decide_kernel(LaunchSlot *slot, const SketchState *st, const ParseCarry *carry,
                                     uint32_t table_cap, uint32_t log_cap) {
    const uint32_t cnt = slot->log_count;
    unsigned int d = DECIDE_GO;
    if (cnt > log_cap) d = DECIDE_OVERFLOW;
    else if ((unsigned long long)st->occupied + cnt > (unsigned long long)(table_cap / 4) * 3) d = DECIDE_FULL;
    slot->decision = d;
    (void)carry;   // chunk_syms is recorded by note_chunk_syms_kernel on the parse stream (the carry may already belong to the next chunk)
}
__global__ void note_chunk_syms_kernel(LaunchSlot *slot, const ParseCarry *carry) { slot->chunk_syms = carry->chunk_syms; }
__global__ void absorb_count_guarded_kernel(LogView log, const LaunchSlot *slot, TableView t, SketchState *st) {
    if (slot->decision != DECIDE_GO) return;
    const uint32_t n = slot->log_count, stride = gridDim.x * blockDim.x;
    const unsigned long long thr = st->threshold;
    const uint32_t lane = threadIdx.x & 31u;
    // warp-uniform trip count: every lane reaches the vote in commit_inserted
    for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x - lane); base < n; base += stride) {
        const uint32_t i = base + lane;
        unsigned long long px = ~0ULL, key = 0;
        bool valid = i < n;
        if (valid) { px = log.posx[i]; valid = px != ~0ULL; }        // ~0: unused slot of a warp's reservation
        if (valid) { key = log.hash[i]; valid = key <= thr; }
        bool ins = false;
        uint32_t s = 0;
        if (valid) s = table_upsert(t, key, st, ins);
        commit_inserted(st, ins);
        if (!valid) continue;
        atomicAdd(&t.cnt[s], 1ULL);
        const unsigned long long extra = px & 0xFFULL;
        if (extra) atomicAdd(&t.ext[s], extra);
        if (px < __ldcg(&t.posx[s])) atomicMin(&t.posx[s], px);
    }
}
__global__ void absorb_kmer_guarded_kernel(LogView log, const LaunchSlot *slot, TableView t, const SketchState *st) {
    if (slot->decision != DECIDE_GO) return;
    const uint32_t n = slot->log_count, stride = gridDim.x * blockDim.x;
    const unsigned long long thr = st->threshold;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long px = log.posx[i];
        if (px == ~0ULL) continue;
        const unsigned long long key = log.hash[i];
        if (key > thr) continue;
        const uint32_t s = table_find(t, key);
        if (s == 0xFFFFFFFFu) continue;
        if (t.posx[s] == px) copy_kmer(t.kmer + (size_t)s * t.kw, log.kmer + (size_t)i * log.kw, t.kw);
    }
}

__global__ void table_clear_kernel(TableView t) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= t.cap) {  // includes the side slot
        t.key[i] = EMPTY_KEY; t.cnt[i] = 0; t.ext[i] = 0; t.posx[i] = ~0ULL;
        for (uint32_t w = 0; w < t.kw; ++w) t.kmer[(size_t)i * t.kw + w] = 0;
    }
}

// Occupied slots -> (key, slot) pairs, arbitrary order.
// One global atomic per block (warp ballots -> shared counter -> block base), as gather_le_kernel below.
__global__ void __launch_bounds__(256)
gather_kernel(TableView t, SketchState *st, unsigned long long *keys, uint32_t *slots) {
    __shared__ uint32_t blk_count, blk_base;
    if (threadIdx.x == 0) blk_count = 0;
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long thr = st->threshold;
    unsigned long long key = EMPTY_KEY;
    bool occ = false;
    if (i < t.cap) { key = t.key[i]; occ = key != EMPTY_KEY && key <= thr; }   // dead keys (above a soft threshold) stay behind
    else if (i == t.cap) occ = st->has_max_key != 0u && thr == EMPTY_KEY;
    const uint32_t m = __ballot_sync(0xffffffffu, occ);
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t wbase = 0;
    if (lane == 0 && m) wbase = atomicAdd(&blk_count, (unsigned int)__popc(m));
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
    __syncthreads();
    if (threadIdx.x == 0 && blk_count) blk_base = atomicAdd(&st->gather_count, blk_count);
    __syncthreads();
    if (occ) {
        const uint32_t idx = blk_base + wbase + __popc(m & ((1u << lane) - 1u));
        keys[idx] = key;
        slots[idx] = i;
    }
}

__global__ void reset_gather_kernel(SketchState *st);

// ---- pruning without a sort: 4096-bin histogram of the occupied keys -> threshold ----------------
// Any threshold T' with #{keys <= T'} >= size is valid (a key of the final
// bottom-s is never above an intermediate threshold), so the cut is placed on a bin boundary.
constexpr int PRUNE_BINS = 4096;
__global__ void __launch_bounds__(256)
table_hist_kernel(TableView t, const SketchState *st, uint32_t shift, uint32_t *__restrict__ bins) {
    __shared__ uint32_t h[PRUNE_BINS];
    for (int i = threadIdx.x; i < PRUNE_BINS; i += blockDim.x) h[i] = 0;
    __syncthreads();
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= t.cap; i += stride) {
        unsigned long long key;
        bool occ;
        if (i == t.cap) { occ = st->has_max_key != 0u; key = EMPTY_KEY; }
        else { key = t.key[i]; occ = key != EMPTY_KEY; }
        if (occ) atomicAdd(&h[min((unsigned long long)(PRUNE_BINS - 1), key >> shift)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PRUNE_BINS; i += blockDim.x)
        if (h[i]) atomicAdd(&bins[i], h[i]);
}
// Histogram of the hashes of a candidate log (valid entries at or below the threshold).
__global__ void __launch_bounds__(256)
log_hist_kernel(LogView log, uint32_t n, const SketchState *st, uint32_t shift, uint32_t *__restrict__ bins) {
    __shared__ uint32_t h[PRUNE_BINS];
    for (int i = threadIdx.x; i < PRUNE_BINS; i += blockDim.x) h[i] = 0;
    __syncthreads();
    const unsigned long long thr = st->threshold;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (log.posx[i] == ~0ULL) continue;
        const unsigned long long key = log.hash[i];
        if (key > thr) continue;
        atomicAdd(&h[min((unsigned long long)(PRUNE_BINS - 1), key >> shift)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PRUNE_BINS; i += blockDim.x)
        if (h[i]) atomicAdd(&bins[i], h[i]);
}
// One block of 1024 threads, 4 bins each: smallest bin boundary with cumulative count >= size.
__global__ void __launch_bounds__(1024)
table_select_kernel(const uint32_t *__restrict__ bins, uint32_t shift, int scaled, unsigned long long size,
                    unsigned long long max_hash, SketchState *st) {
    __shared__ uint32_t wsum[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint32_t b0 = bins[4 * tid], b1 = bins[4 * tid + 1], b2 = bins[4 * tid + 2], b3 = bins[4 * tid + 3];
    const uint32_t c = b0 + b1 + b2 + b3;
    uint32_t x = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, x, d); if (lane >= (uint32_t)d) x += n; }
    if (lane == 31u) wsum[wid] = x;
    __syncthreads();
    if (wid == 0u) {
        uint32_t v = wsum[lane], y = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, y, d); if (lane >= (uint32_t)d) y += n; }
        wsum[lane] = y - v;
    }
    __syncthreads();
    const unsigned long long before = (unsigned long long)wsum[wid] + x - c;  // keys in bins < 4*tid
    if (tid == 0u) st->new_threshold = st->threshold;                          // default: no change
    __syncthreads();
    if (size > 0 && before < size && before + c >= size) {                     // the crossing is in my 4 bins
        unsigned long long cum = before;
        uint32_t b = 4 * tid;
        const uint32_t bb[4] = {b0, b1, b2, b3};
#pragma unroll
        for (int j = 0; j < 4; ++j) { cum += bb[j]; if (cum >= size) { b = 4 * tid + j; break; } }
        unsigned long long thr = ~0ULL;
        if (b + 1u < (uint32_t)PRUNE_BINS && shift < 64u) {
            const unsigned long long top = ((unsigned long long)(b + 1u)) << shift;
            // (b+1) << shift may exceed 64 bits only for the last bin, excluded above
            thr = top - 1ULL;
        }
        if (scaled && thr < max_hash) thr = max_hash;
        if (thr < st->threshold) st->new_threshold = thr;
    }
}
// Occupied slots with key <= new_threshold -> (key, slot), arbitrary order; count in gather_count.
// One global atomic per block (warp ballots -> shared counter -> block base).
__global__ void __launch_bounds__(256)
gather_le_kernel(TableView t, SketchState *st, unsigned long long *keys, uint32_t *slots) {
    __shared__ uint32_t blk_count, blk_base;
    if (threadIdx.x == 0) blk_count = 0;
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long thr = st->new_threshold;
    unsigned long long key = EMPTY_KEY;
    bool occ = false;
    if (i < t.cap) { key = t.key[i]; occ = key != EMPTY_KEY; }
    else if (i == t.cap) occ = st->has_max_key != 0u;
    occ = occ && key <= thr;
    const uint32_t m = __ballot_sync(0xffffffffu, occ);
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t wbase = 0;
    if (lane == 0 && m) wbase = atomicAdd(&blk_count, (unsigned int)__popc(m));
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
    __syncthreads();
    if (threadIdx.x == 0 && blk_count) blk_base = atomicAdd(&st->gather_count, blk_count);
    __syncthreads();
    if (occ) {
        const uint32_t idx = blk_base + wbase + __popc(m & ((1u << lane) - 1u));
        keys[idx] = key;
        slots[idx] = i;
    }
}

// ---- LSD radix sort, 8 bits per pass, (u64 key, u32 value), stable --------------------------
// Work unit = one warp over a contiguous segment of SORT_SEG keys.
constexpr int SORT_WARPS = 8;
constexpr uint32_t SORT_SEG = 2048;

__global__ void __launch_bounds__(SORT_WARPS * 32)
radix_hist_kernel(const unsigned long long *__restrict__ keys, uint32_t n, int shift, uint32_t n_segs,
                  uint32_t *__restrict__ hist /* [256][n_segs] */) {
    __shared__ uint32_t h[SORT_WARPS][256];
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    for (int i = lane; i < 256; i += 32) h[w][i] = 0;
    __syncwarp();
    const uint32_t seg = blockIdx.x * SORT_WARPS + w;
    if (seg < n_segs) {
        const uint32_t a = seg * SORT_SEG, b = min(a + SORT_SEG, n);
        for (uint32_t i = a + lane; i < b; i += 32) atomicAdd(&h[w][(uint32_t)(keys[i] >> shift) & 255u], 1u);
        __syncwarp();
        for (int d = lane; d < 256; d += 32) hist[(uint32_t)d * n_segs + seg] = h[w][d];
    }
}
// Exclusive scan of hist in (digit-major, segment-minor) order.  Single block of 1024 threads.
__global__ void __launch_bounds__(1024) radix_scan_kernel(uint32_t *hist, uint32_t total) {
    __shared__ uint32_t part[1024];
    const uint32_t tid = threadIdx.x;
    const uint32_t G = (total + 1023u) / 1024u;
    const uint32_t a = min(tid * G, total), b = min(a + G, total);
    uint32_t s = 0;
    for (uint32_t i = a; i < b; ++i) s += hist[i];
    part[tid] = s;
    __syncthreads();
    // inclusive Hillis-Steele over 1024 partials
    for (uint32_t d = 1; d < 1024; d <<= 1) {
        const uint32_t v = tid >= d ? part[tid - d] : 0u;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    uint32_t run = part[tid] - s;
    for (uint32_t i = a; i < b; ++i) { const uint32_t v = hist[i]; hist[i] = run; run += v; }
}
__global__ void __launch_bounds__(SORT_WARPS * 32)
radix_scatter_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ vals, uint32_t n,
                     int shift, uint32_t n_segs, const uint32_t *__restrict__ hist,
                     unsigned long long *__restrict__ okeys, uint32_t *__restrict__ ovals) {
    __shared__ uint32_t off[SORT_WARPS][256];
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const uint32_t seg = blockIdx.x * SORT_WARPS + w;
    if (seg >= n_segs) return;
    for (int d = lane; d < 256; d += 32) off[w][d] = hist[(uint32_t)d * n_segs + seg];
    __syncwarp();
    const uint32_t a = seg * SORT_SEG, b = min(a + SORT_SEG, n);
    for (uint32_t base = a; base < b; base += 32) {
        const uint32_t i = base + lane;
        const bool in = i < b;
        unsigned long long key = 0; uint32_t val = 0, d = 0;
        if (in) { key = keys[i]; val = vals[i]; d = (uint32_t)(key >> shift) & 255u; }
        const uint32_t act = __ballot_sync(0xffffffffu, in);
        const uint32_t peers = __match_any_sync(0xffffffffu, in ? d : (256u + lane)) & act;
        uint32_t dst = 0;
        if (in) {
            const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
            dst = off[w][d] + rank;
        }
        __syncwarp();
        if (in && (int)lane == (31 - __clz(peers))) off[w][d] += __popc(peers);  // one writer per digit
        __syncwarp();
        if (in) { okeys[dst] = key; ovals[dst] = val; }
    }
}

// ---- the same sort for LARGE inputs: one block per tile of 4096 keys -------------------------------
// The warp-per-segment scatter above writes every key to its own place (32 lanes, up to 32 digits: 8-byte writes all
// over the output); fine for a table's worth of keys, 19 ms per pass for the 10^8 postings of `dist` (dist.cu).  Here a
// block first orders its tile by digit in shared memory (stable: warps own consecutive sub-segments, lanes rank by
// match_any as above) and then writes it out digit by digit, so that consecutive threads write consecutive addresses
// (runs of ~16 keys per digit); the offsets come from a per-digit scan over the tiles (one block per digit).
constexpr uint32_t RT_TILE = 4096, RT_THREADS = 256, RT_WARPS = RT_THREADS / 32, RT_SUB = RT_TILE / RT_WARPS, RT_IT = RT_SUB / 32;
constexpr uint32_t RT_SMEM = RT_TILE * 8u + RT_TILE * 4u + RT_WARPS * 256u * 4u + 256u * 4u * 2u + 64u;

__global__ void __launch_bounds__(RT_THREADS)
radix_tile_hist_kernel(const unsigned long long *__restrict__ keys, uint32_t n, int shift, uint32_t n_tiles,
                       uint32_t *__restrict__ hist /* [256][n_tiles] */) {
    __shared__ uint32_t h[RT_WARPS][256];
    const uint32_t tid = threadIdx.x, w = tid >> 5;
    for (uint32_t i = tid; i < RT_WARPS * 256u; i += RT_THREADS) (&h[0][0])[i] = 0u;
    __syncthreads();
    const uint32_t tile = blockIdx.x, a = tile * RT_TILE, b = min(a + RT_TILE, n);
    for (uint32_t i = a + tid; i < b; i += RT_THREADS) atomicAdd(&h[w][(uint32_t)(keys[i] >> shift) & 255u], 1u);
    __syncthreads();
    uint32_t c = 0;
#pragma unroll
    for (uint32_t x = 0; x < RT_WARPS; ++x) c += h[x][tid];
    hist[tid * n_tiles + tile] = c;
}
// block d: exclusive scan of digit d's counts over the tiles (in place), its total to totals[d]
__global__ void __launch_bounds__(1024) radix_tile_scan_kernel(uint32_t *hist, uint32_t n_tiles, uint32_t *__restrict__ totals) {
    __shared__ uint32_t part[1024];
    const uint32_t tid = threadIdx.x;
    uint32_t *row = hist + (size_t)blockIdx.x * n_tiles;
    const uint32_t G = (n_tiles + 1023u) / 1024u;
    const uint32_t a = min(tid * G, n_tiles), b = min(a + G, n_tiles);
    uint32_t s = 0;
    for (uint32_t i = a; i < b; ++i) s += row[i];
    part[tid] = s;
    __syncthreads();
    for (uint32_t d = 1; d < 1024; d <<= 1) {
        const uint32_t v = tid >= d ? part[tid - d] : 0u;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    uint32_t run = part[tid] - s;
    for (uint32_t i = a; i < b; ++i) { const uint32_t v = row[i]; row[i] = run; run += v; }
    if (tid == 1023u) totals[blockIdx.x] = part[1023];
}
// exclusive scan of one value per thread over the 256 threads of a block
__device__ __forceinline__ uint32_t block_exscan_256(uint32_t v, uint32_t *tmp /* RT_WARPS + 1 words */) {
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= (uint32_t)d) x += y; }
    __syncthreads();                       // tmp may still be read from a previous call
    if (lane == 31u) tmp[w] = x;
    __syncthreads();
    uint32_t base = 0;
#pragma unroll
    for (uint32_t i = 0; i < RT_WARPS; ++i) if (i < w) base += tmp[i];
    return base + x - v;
}
__global__ void __launch_bounds__(RT_THREADS)
radix_tile_scatter_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ vals, uint32_t n, int shift,
                          uint32_t n_tiles, const uint32_t *__restrict__ hist, const uint32_t *__restrict__ totals,
                          unsigned long long *__restrict__ okeys, uint32_t *__restrict__ ovals) {
    extern __shared__ __align__(16) uint8_t rt_smem[];
    unsigned long long *sk = reinterpret_cast<unsigned long long *>(rt_smem);
    uint32_t *sv = reinterpret_cast<uint32_t *>(rt_smem + RT_TILE * 8u);
    uint32_t *cnt = sv + RT_TILE;                       // [RT_WARPS][256]: counts, then the running offset of (warp, digit)
    uint32_t *dbase = cnt + RT_WARPS * 256u;            // [256] first tile-local position of a digit
    uint32_t *gbase = dbase + 256u;                     // [256] global position of the tile's first key of a digit
    uint32_t *tmp = gbase + 256u;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    const uint32_t tile = blockIdx.x, a = tile * RT_TILE, nt = min(RT_TILE, n - a);
    for (uint32_t i = tid; i < RT_WARPS * 256u; i += RT_THREADS) cnt[i] = 0u;
    __syncthreads();
    unsigned long long k[RT_IT];
    uint32_t v[RT_IT];
#pragma unroll
    for (uint32_t it = 0; it < RT_IT; ++it) {
        const uint32_t idx = w * RT_SUB + it * 32u + lane;
        const bool in = idx < nt;
        k[it] = in ? keys[a + idx] : 0ULL;
        v[it] = in ? vals[a + idx] : 0u;
        if (in) atomicAdd(&cnt[w * 256u + ((uint32_t)(k[it] >> shift) & 255u)], 1u);
    }
    __syncthreads();
    {   // thread = digit
        uint32_t run = 0;
#pragma unroll
        for (uint32_t x = 0; x < RT_WARPS; ++x) { const uint32_t c = cnt[x * 256u + tid]; cnt[x * 256u + tid] = run; run += c; }
        const uint32_t local = block_exscan_256(run, tmp);
        const uint32_t gstart = block_exscan_256(totals[tid], tmp);
        dbase[tid] = local;
        gbase[tid] = gstart + hist[tid * n_tiles + tile];
    }
    __syncthreads();
#pragma unroll
    for (uint32_t it = 0; it < RT_IT; ++it) {
        const uint32_t idx = w * RT_SUB + it * 32u + lane;
        const bool in = idx < nt;
        const uint32_t d = (uint32_t)(k[it] >> shift) & 255u;
        const uint32_t act = __ballot_sync(0xffffffffu, in);
        const uint32_t peers = __match_any_sync(0xffffffffu, in ? d : (256u + lane)) & act;
        uint32_t p = 0;
        if (in) p = dbase[d] + cnt[w * 256u + d] + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
        if (in && (int)lane == (31 - __clz(peers))) cnt[w * 256u + d] += __popc(peers);   // one writer per digit
        __syncwarp();
        if (in) { sk[p] = k[it]; sv[p] = v[it]; }
    }
    __syncthreads();
    for (uint32_t p = tid; p < nt; p += RT_THREADS) {
        const unsigned long long key = sk[p];
        const uint32_t d = (uint32_t)(key >> shift) & 255u, dst = gbase[d] + (p - dbase[d]);
        okeys[dst] = key;
        ovals[dst] = sv[p];
    }
}

// ---- bucket + rank sort for (nearly) uniform keys ------------------------------------------------
// Sketch keys are murmur hashes below a known threshold, i.e. uniform: 4096 buckets on the top bits
// hold ~n/4096 keys each, and a block ranks one bucket in shared memory by counting.  4 launches
// instead of the radix sort's 24.  The host falls back to the radix sort when a bucket exceeds
// BUCKET_CAP (non-uniform keys).
constexpr uint32_t BUCKET_CAP = 2048;
__global__ void __launch_bounds__(256)
bucket_hist_kernel(const unsigned long long *__restrict__ keys, uint32_t n, uint32_t shift, uint32_t *__restrict__ bins) {
    __shared__ uint32_t h[PRUNE_BINS];
    for (int i = threadIdx.x; i < PRUNE_BINS; i += blockDim.x) h[i] = 0;
    __syncthreads();
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        atomicAdd(&h[min((unsigned long long)(PRUNE_BINS - 1), keys[i] >> shift)], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < PRUNE_BINS; i += blockDim.x)
        if (h[i]) atomicAdd(&bins[i], h[i]);
}
// bins[4096] -> exclusive offsets in offs[4096] (+ cursors), largest bucket in st->gather_count
__global__ void __launch_bounds__(1024) bucket_scan_kernel(const uint32_t *__restrict__ bins, uint32_t *__restrict__ offs,
                                                           uint32_t *__restrict__ cursor, SketchState *st) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t wmax[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint32_t b0 = bins[4 * tid], b1 = bins[4 * tid + 1], b2 = bins[4 * tid + 2], b3 = bins[4 * tid + 3];
    const uint32_t c = b0 + b1 + b2 + b3;
    uint32_t mx = max(max(b0, b1), max(b2, b3));
    uint32_t x = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, x, d); if (lane >= (uint32_t)d) x += v; }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    if (lane == 31u) wsum[wid] = x;
    if (lane == 0u) wmax[wid] = mx;
    __syncthreads();
    if (wid == 0u) {
        uint32_t v = wsum[lane], y = v, m2 = wmax[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, y, d); if (lane >= (uint32_t)d) y += u; }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) m2 = max(m2, __shfl_xor_sync(0xffffffffu, m2, d));
        wsum[lane] = y - v;
        if (lane == 0u) st->gather_count = m2;   // reused as "largest bucket"
    }
    __syncthreads();
    uint32_t o = wsum[wid] + x - c;
    offs[4 * tid] = o; cursor[4 * tid] = o; o += b0;
    offs[4 * tid + 1] = o; cursor[4 * tid + 1] = o; o += b1;
    offs[4 * tid + 2] = o; cursor[4 * tid + 2] = o; o += b2;
    offs[4 * tid + 3] = o; cursor[4 * tid + 3] = o;
}
__global__ void bucket_scatter_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ vals,
                                      uint32_t n, uint32_t shift, uint32_t *__restrict__ cursor,
                                      unsigned long long *__restrict__ okeys, uint32_t *__restrict__ ovals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = keys[i];
    const uint32_t b = (uint32_t)min((unsigned long long)(PRUNE_BINS - 1), k >> shift);
    const uint32_t pos = atomicAdd(&cursor[b], 1u);
    okeys[pos] = k; ovals[pos] = vals[i];
}
// one block per bucket: rank every key among the bucket's keys (all distinct) and write it in order
__global__ void __launch_bounds__(128)
bucket_rank_kernel(const unsigned long long *__restrict__ ikeys, const uint32_t *__restrict__ ivals,
                   const uint32_t *__restrict__ offs, const uint32_t *__restrict__ bins,
                   unsigned long long *__restrict__ okeys, uint32_t *__restrict__ ovals) {
    __shared__ unsigned long long sk[BUCKET_CAP];
    const uint32_t b = blockIdx.x, o = offs[b], m = bins[b];
    if (m == 0 || m > BUCKET_CAP) return;
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) sk[i] = ikeys[o + i];
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
        const unsigned long long k = sk[i];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < m; ++j) rank += sk[j] < k ? 1u : 0u;
        okeys[o + rank] = k; ovals[o + rank] = ivals[o + i];
    }
}

// After the sort: how many entries the sketch keeps, and the threshold that follows.
//   Mash:   keep = min(n, s)                      new_threshold = key[s-1] when n >= s
//   Scaled: keep = max(#{key <= max_hash}, min(n, s)); threshold = max(max_hash, key[s-1]) when n >= s
__global__ void select_keep_kernel(const unsigned long long *keys, uint32_t n, int scaled,
                                   unsigned long long size, unsigned long long max_hash, SketchState *st) {
    if (threadIdx.x || blockIdx.x) return;
    uint32_t keep = (unsigned long long)n < size ? n : (uint32_t)size;
    unsigned long long thr = st->threshold;
    if (!scaled) {
        if ((unsigned long long)n >= size && size > 0) thr = min(thr, keys[size - 1]);
    } else {
        uint32_t lo = 0, hi = n;  // first index with key > max_hash
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (keys[mid] <= max_hash) lo = mid + 1; else hi = mid; }
        if (lo > keep) keep = lo;
        if (size == 0) { keep = lo; thr = min(thr, max_hash); }
        else if ((unsigned long long)n >= size) thr = min(thr, max(max_hash, keys[size - 1]));
    }
    st->keep_count = keep;
    st->new_threshold = thr;
}
__global__ void commit_threshold_kernel(SketchState *st) { st->threshold = st->new_threshold; }

// Re-insert the first `keep` sorted entries of the old table into a (cleared) new table.
__global__ void rebuild_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ slots,
                               uint32_t keep, TableView from, TableView to, SketchState *st) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool ins = false;
    uint32_t dst = 0;
    if (i < keep) dst = table_upsert(to, keys[i], st, ins);
    commit_inserted(st, ins);
    if (i >= keep) return;
    const uint32_t src = slots[i];
    to.cnt[dst] = from.cnt[src]; to.ext[dst] = from.ext[src];
    to.posx[dst] = from.posx[src];
    copy_kmer(to.kmer + (size_t)dst * to.kw, from.kmer + (size_t)src * from.kw, to.kw);
}
__global__ void reset_occupancy_kernel(SketchState *st) { st->occupied = 0; st->has_max_key = 0; }
__global__ void reset_gather_kernel(SketchState *st) { st->gather_count = 0; }

// to_vec export: first `keep` sorted entries -> SoA with saturating u32 counts (mash.rs:48-49).
__global__ void export_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ slots,
                              uint32_t keep, TableView t, unsigned long long *o_hash, uint32_t *o_cnt,
                              uint32_t *o_ext, unsigned long long *o_kmer, unsigned long long *o_posx) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= keep) return;
    const uint32_t s = slots[i];
    const unsigned long long c = t.cnt[s], e = t.ext[s];
    o_hash[i] = keys[i];
    o_cnt[i] = c > 0xFFFFFFFFULL ? 0xFFFFFFFFu : (uint32_t)c;
    o_ext[i] = e > 0xFFFFFFFFULL ? 0xFFFFFFFFu : (uint32_t)e;
    copy_kmer(o_kmer + (size_t)i * t.kw, t.kmer + (size_t)s * t.kw, t.kw);
    o_posx[i] = t.posx[s];
}

// Select rows of the exported SoA by index (idx == nullptr: the first m rows) and expand the 2-bit
// k-mer codes to the ASCII bytes KmerCount::kmer holds.  Pushed k-mers (arena flag) keep their index.
__global__ void select_rows_kernel(const uint32_t *__restrict__ idx, uint32_t m, int k, uint32_t stride, uint32_t kw,
                                   const unsigned long long *__restrict__ i_hash, const uint32_t *__restrict__ i_cnt,
                                   const uint32_t *__restrict__ i_ext, const unsigned long long *__restrict__ i_kmer,
                                   const unsigned long long *__restrict__ i_posx, unsigned long long *o_hash,
                                   uint32_t *o_cnt, uint32_t *o_ext, unsigned long long *o_kmer,
                                   unsigned long long *o_posx, uint8_t *o_bytes) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const uint32_t i = idx ? idx[j] : j;
    const unsigned long long *codes = i_kmer + (size_t)i * kw;
    const unsigned long long px = i_posx[i];
    o_hash[j] = i_hash[i]; o_cnt[j] = i_cnt[i]; o_ext[j] = i_ext[i]; o_kmer[j] = codes[0]; o_posx[j] = px;
    uint8_t *dst = o_bytes + (size_t)j * stride;
    if (px & (1ULL << 8)) { for (uint32_t t = 0; t < stride; ++t) dst[t] = 0; }
    else {
        for (int t = 0; t < k; ++t) dst[t] = (uint8_t)"ACGT"[(codes[t >> 5] >> (2 * (t & 31))) & 3ULL];
        for (uint32_t t = (uint32_t)k; t < stride; ++t) dst[t] = 0;
    }
}

// ---- exact union of two sketch tables (one file split across GPUs) ------------------------------
// `src` is the table of another sketcher -- on a PEER GPU: the kernels below run on the destination GPU and read the
// source table straight out of the peer's HBM over NVLink (peer access), so the gather of the peer's entries and
// their merge into the local table are one pass.  Every live entry of src (key <= both thresholds) is upserted:
// totals add (64-bit), the smaller position id wins and donates the k-mer (position ids of later byte ranges are
// larger, so "first occurrence in stream order" survives the merge).  Exact because a hash of the global bottom-s is
// live in every local table that saw it, with complete local totals.
__global__ void merge_count_kernel(TableView src, unsigned long long src_thr, unsigned int src_has_max, TableView dst,
                                   SketchState *st) {
    const uint32_t stride = gridDim.x * blockDim.x, lane = threadIdx.x & 31u;
    const unsigned long long thr = st->threshold < src_thr ? st->threshold : src_thr;
    for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x - lane); base <= src.cap; base += stride) {
        const uint32_t i = base + lane;
        unsigned long long key = EMPTY_KEY;
        bool valid = false;
        if (i < src.cap) { key = src.key[i]; valid = key != EMPTY_KEY && key <= thr; }
        else if (i == src.cap) valid = src_has_max != 0u && thr == EMPTY_KEY;
        bool ins = false;
        uint32_t d = 0;
        if (valid) d = table_upsert(dst, key, st, ins);
        commit_inserted(st, ins);
        if (!valid) continue;
        atomicAdd(&dst.cnt[d], src.cnt[i]);
        const unsigned long long e = src.ext[i];
        if (e) atomicAdd(&dst.ext[d], e);
        atomicMin(&dst.posx[d], src.posx[i]);
    }
}
__global__ void merge_kmer_kernel(TableView src, unsigned long long src_thr, unsigned int src_has_max, TableView dst,
                                  const SketchState *st) {
    const uint32_t stride = gridDim.x * blockDim.x;
    const unsigned long long thr = st->threshold < src_thr ? st->threshold : src_thr;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= src.cap; i += stride) {
        unsigned long long key = EMPTY_KEY;
        bool valid = false;
        if (i < src.cap) { key = src.key[i]; valid = key != EMPTY_KEY && key <= thr; }
        else valid = src_has_max != 0u && thr == EMPTY_KEY;
        if (!valid) continue;
        const uint32_t d = table_find(dst, key);
        if (d == 0xFFFFFFFFu) continue;
        if (dst.posx[d] == src.posx[i]) copy_kmer(dst.kmer + (size_t)d * dst.kw, src.kmer + (size_t)i * src.kw, dst.kw);
    }
}
void launch_merge_tables(TableView src, unsigned long long src_thr, unsigned int src_has_max, TableView dst, SketchState *st,
                         cudaStream_t s) {
    const uint32_t blocks = min((src.cap + 256u) / 256u, 2368u);
    merge_count_kernel<<<blocks, 256, 0, s>>>(src, src_thr, src_has_max, dst, st);
    merge_kmer_kernel<<<blocks, 256, 0, s>>>(src, src_thr, src_has_max, dst, st);
}

// Test hook (fb2_sketcher_debug_bump): add to the 64-bit totals of an existing key, so a test can bring a count to
// the edge of u32 without pushing 2^32 k-mers (mash.rs:48-49 saturate; export_kernel clamps).
__global__ void debug_bump_kernel(TableView t, unsigned long long key, unsigned long long add_cnt, unsigned long long add_ext,
                                  unsigned int *found) {
    const uint32_t slot = table_find(t, key);
    if (slot == 0xFFFFFFFFu) { *found = 0u; return; }
    t.cnt[slot] += add_cnt; t.ext[slot] += add_ext;
    *found = 1u;
}
void launch_debug_bump(TableView t, unsigned long long key, unsigned long long add_cnt, unsigned long long add_ext,
                       unsigned int *found, cudaStream_t s) {
    debug_bump_kernel<<<1, 1, 0, s>>>(t, key, add_cnt, add_ext, found);
}

// ---- launchers ---------------------------------------------------------------------------------
static inline uint32_t cdiv(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

void launch_absorb(LogView log, uint32_t i0, uint32_t i1, TableView t, SketchState *st, cudaStream_t s) {
    if (i1 <= i0) return;
    Band all; all.lo = 0; all.hi = ~0ULL; all.use_lo = 0;
    absorb_count_kernel<<<cdiv(i1 - i0, 256), 256, 0, s>>>(log, i0, i1, t, st, all);
    absorb_kmer_kernel<<<cdiv(i1 - i0, 256), 256, 0, s>>>(log, i0, i1, t, st, all);
}
void launch_absorb_band(LogView log, uint32_t n, TableView t, SketchState *st, unsigned long long lo, int use_lo,
                        unsigned long long hi, cudaStream_t s) {
    if (!n) return;
    Band b; b.lo = lo; b.hi = hi; b.use_lo = use_lo;
    absorb_count_kernel<<<cdiv(n, 256), 256, 0, s>>>(log, 0, n, t, st, b);
    absorb_kmer_kernel<<<cdiv(n, 256), 256, 0, s>>>(log, 0, n, t, st, b);
}
void launch_log_hist(LogView log, uint32_t n, const SketchState *st, uint32_t shift, uint32_t *bins, cudaStream_t s) {
    cudaMemsetAsync(bins, 0, PRUNE_BINS * sizeof(uint32_t), s);
    if (n) log_hist_kernel<<<min(cdiv(n, 256), 1184u), 256, 0, s>>>(log, n, st, shift, bins);
}
void launch_note_chunk_syms(LaunchSlot *slot, const ParseCarry *carry, cudaStream_t s) { note_chunk_syms_kernel<<<1, 1, 0, s>>>(slot, carry); }
void launch_absorb_guarded(LogView log, LaunchSlot *slot, TableView t, SketchState *st, const ParseCarry *carry,
                           uint32_t expect, cudaStream_t s) {
    absorb_decide_kernel<<<1, 1, 0, s>>>(slot, st, carry, t.cap, log.cap);
    const uint32_t blocks = max(64u, min(cdiv(expect, 256), 2048u));
    absorb_count_guarded_kernel<<<blocks, 256, 0, s>>>(log, slot, t, st);
    absorb_kmer_guarded_kernel<<<blocks, 256, 0, s>>>(log, slot, t, st);
}
void launch_table_clear(TableView t, cudaStream_t s) {
    table_clear_kernel<<<cdiv(t.cap + 1, 256), 256, 0, s>>>(t);
}
void launch_gather(TableView t, SketchState *st, unsigned long long *keys, uint32_t *slots, cudaStream_t s) {
    reset_gather_kernel<<<1, 1, 0, s>>>(st);
    gather_kernel<<<cdiv(t.cap + 1, 256), 256, 0, s>>>(t, st, keys, slots);
}
// Sorts (keys, vals) of length n; result ends in (keys, vals) (8 passes ping-pong back).
void launch_radix_sort(unsigned long long *keys, uint32_t *vals, unsigned long long *tkeys, uint32_t *tvals,
                       uint32_t n, uint32_t *hist, cudaStream_t s) {
    if (n < 2) return;
    unsigned long long *ka = keys, *kb = tkeys;
    uint32_t *va = vals, *vb = tvals;
    if (n >= (1u << 18) && !getenv("FB2_RADIX_V1")) {   // large: tiles ordered in shared memory, coalesced output
        static bool attr_set[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        bool ok = true;
        if (dev >= 0 && dev < 64 && !attr_set[dev]) {
            ok = cudaFuncSetAttribute(radix_tile_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RT_SMEM) == cudaSuccess;
            attr_set[dev] = ok;
        }
        if (ok) {
            const uint32_t n_tiles = cdiv(n, RT_TILE);
            uint32_t *totals = hist + (size_t)256u * n_tiles;          // (radix_hist_words leaves room)
            for (int pass = 0; pass < 8; ++pass) {
                radix_tile_hist_kernel<<<n_tiles, RT_THREADS, 0, s>>>(ka, n, pass * 8, n_tiles, hist);
                radix_tile_scan_kernel<<<256, 1024, 0, s>>>(hist, n_tiles, totals);
                radix_tile_scatter_kernel<<<n_tiles, RT_THREADS, RT_SMEM, s>>>(ka, va, n, pass * 8, n_tiles, hist, totals, kb, vb);
                unsigned long long *tk = ka; ka = kb; kb = tk;
                uint32_t *tv = va; va = vb; vb = tv;
            }
            return;
        }
    }
    const uint32_t n_segs = cdiv(n, SORT_SEG);
    const uint32_t blocks = cdiv(n_segs, SORT_WARPS);
    for (int pass = 0; pass < 8; ++pass) {
        radix_hist_kernel<<<blocks, SORT_WARPS * 32, 0, s>>>(ka, n, pass * 8, n_segs, hist);
        radix_scan_kernel<<<1, 1024, 0, s>>>(hist, 256u * n_segs);
        radix_scatter_kernel<<<blocks, SORT_WARPS * 32, 0, s>>>(ka, va, n, pass * 8, n_segs, hist, kb, vb);
        unsigned long long *tk = ka; ka = kb; kb = tk;
        uint32_t *tv = va; va = vb; vb = tv;
    }
}
// hist -> select -> filtered gather; the host then reads gather_count (= keys kept).
void launch_prune_select(TableView t, SketchState *st, uint32_t shift, uint32_t *bins, int scaled,
                         unsigned long long size, unsigned long long max_hash, unsigned long long *keys,
                         uint32_t *slots, cudaStream_t s) {
    cudaMemsetAsync(bins, 0, PRUNE_BINS * sizeof(uint32_t), s);
    const uint32_t blocks = min(cdiv(t.cap + 1, 256), 592u);
    table_hist_kernel<<<blocks, 256, 0, s>>>(t, st, shift, bins);
    table_select_kernel<<<1, 1024, 0, s>>>(bins, shift, scaled, size, max_hash, st);
    reset_gather_kernel<<<1, 1, 0, s>>>(st);
    gather_le_kernel<<<cdiv(t.cap + 1, 256), 256, 0, s>>>(t, st, keys, slots);
}
// Bucket sort (keys <= 2^bits - 1 assumed roughly uniform).  Result lands in (tkeys, tvals); the
// largest bucket is left in st->gather_count for the host to check against bucket_cap().
void launch_bucket_sort(const unsigned long long *keys, const uint32_t *vals, unsigned long long *tkeys, uint32_t *tvals,
                        unsigned long long *okeys, uint32_t *ovals, uint32_t n, uint32_t shift, uint32_t *bins,
                        uint32_t *offs, uint32_t *cursor, SketchState *st, cudaStream_t s) {
    cudaMemsetAsync(bins, 0, PRUNE_BINS * sizeof(uint32_t), s);
    if (!n) return;
    bucket_hist_kernel<<<min(cdiv(n, 256), 1184u), 256, 0, s>>>(keys, n, shift, bins);
    bucket_scan_kernel<<<1, 1024, 0, s>>>(bins, offs, cursor, st);
    bucket_scatter_kernel<<<cdiv(n, 256), 256, 0, s>>>(keys, vals, n, shift, cursor, tkeys, tvals);
    bucket_rank_kernel<<<PRUNE_BINS, 128, 0, s>>>(tkeys, tvals, offs, bins, okeys, ovals);
}
uint32_t bucket_cap() { return BUCKET_CAP; }
uint32_t radix_hist_words(uint32_t n) { return 256u * cdiv(n ? n : 1, SORT_SEG) + 256u; }   // (+ the tile path's digit totals)

void launch_select_rows(const uint32_t *idx, uint32_t m, int k, uint32_t stride, uint32_t kw, const unsigned long long *i_hash,
                        const uint32_t *i_cnt, const uint32_t *i_ext, const unsigned long long *i_kmer,
                        const unsigned long long *i_posx, unsigned long long *o_hash, uint32_t *o_cnt, uint32_t *o_ext,
                        unsigned long long *o_kmer, unsigned long long *o_posx, uint8_t *o_bytes, cudaStream_t s) {
    if (m) select_rows_kernel<<<cdiv(m, 128), 128, 0, s>>>(idx, m, k, stride, kw, i_hash, i_cnt, i_ext, i_kmer, i_posx, o_hash,
                                                          o_cnt, o_ext, o_kmer, o_posx, o_bytes);
}
void launch_select_keep(const unsigned long long *keys, uint32_t n, int scaled, unsigned long long size,
                        unsigned long long max_hash, SketchState *st, cudaStream_t s) {
    select_keep_kernel<<<1, 1, 0, s>>>(keys, n, scaled, size, max_hash, st);
}
void launch_commit_threshold(SketchState *st, cudaStream_t s) { commit_threshold_kernel<<<1, 1, 0, s>>>(st); }
// Rebuild the live histogram of table `t` with a new shift (after a rebuild / reset).
__global__ void set_hist_shift_kernel(SketchState *st, uint32_t shift) { st->hist_shift = shift; }
void launch_live_hist_refresh(TableView t, SketchState *st, uint32_t shift, uint32_t *live_bins, cudaStream_t s) {
    cudaMemsetAsync(live_bins, 0, PRUNE_BINS * sizeof(uint32_t), s);
    set_hist_shift_kernel<<<1, 1, 0, s>>>(st, shift);
    const uint32_t blocks = min(cdiv(t.cap + 1, 256), 592u);
    table_hist_kernel<<<blocks, 256, 0, s>>>(t, st, shift, live_bins);
}
// Soft threshold update: lower the admission threshold to the smallest bin boundary that keeps >= size
// keys, from the live histogram alone (no rebuild; entries above it simply stop being updated).
void launch_soft_threshold(const uint32_t *live_bins, uint32_t shift, int scaled, unsigned long long size,
                           unsigned long long max_hash, SketchState *st, cudaStream_t s) {
    table_select_kernel<<<1, 1024, 0, s>>>(live_bins, shift, scaled, size, max_hash, st);
    commit_threshold_kernel<<<1, 1, 0, s>>>(st);
}
void launch_rebuild(const unsigned long long *keys, const uint32_t *slots, uint32_t keep, TableView from,
                    TableView to, SketchState *st, cudaStream_t s) {
    launch_table_clear(to, s);
    reset_occupancy_kernel<<<1, 1, 0, s>>>(st);
    if (keep) rebuild_kernel<<<cdiv(keep, 256), 256, 0, s>>>(keys, slots, keep, from, to, st);
}
void launch_export(const unsigned long long *keys, const uint32_t *slots, uint32_t keep, TableView t,
                   unsigned long long *o_hash, uint32_t *o_cnt, uint32_t *o_ext, unsigned long long *o_kmer,
                   unsigned long long *o_posx, cudaStream_t s) {
    if (keep) export_kernel<<<cdiv(keep, 256), 256, 0, s>>>(keys, slots, keep, t, o_hash, o_cnt, o_ext, o_kmer, o_posx);
}

// ---- sketch filters on the exported columns (FilterParams::filter_counts, filtering.rs:60-87) -------------------
// The host version (hostlogic.cpp, fb2_filter_select) walks 200 000 counts per C2 sketch; at the end of a stream
// that is half a millisecond of an idle device.  Same decisions here, per entry in parallel:
//   filter_pass_kernel  : strand filter flag (filter_strands, filtering.rs:413-432: the same f64 division, correctly
//                         rounded on both sides) and the histogram of the counts that pass (statistics.rs:30-47),
//                         up to FILTER_HCAP bins; larger counts are only counted -- the caller then takes the host path
//   (host)              : guess_filter_threshold over that histogram (a walk over max-count bins)
//   filter_select_kernel: abundance cut (filter_abundance, filtering.rs:329-343) and the first `limit` survivors,
//                         in order, as row indices for select_rows
constexpr uint32_t FILTER_SMEM_BINS = 2048;
__global__ void filter_pass_kernel(const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ ext, uint32_t n, int strand,
                                   double cut, int err, uint32_t *__restrict__ hist, uint32_t hcap, uint8_t *__restrict__ ok,
                                   uint32_t *meta /* [0] counts above hcap, [1] largest count that passed */) {
    __shared__ uint32_t sh[FILTER_SMEM_BINS];
    for (uint32_t i = threadIdx.x; i < FILTER_SMEM_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    uint32_t mx = 0, over = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t c = cnt[i];
        bool pass = true;
        if (strand && c >= 16u) {   // fewer observations are too noisy to call an adapter
            const uint32_t e = ext[i], lowest = min(e, c - e);
            pass = __ddiv_rn((double)lowest, (double)c) >= cut;
        }
        if (strand) ok[i] = pass ? 1 : 0;
        if (err && pass && c) {
            if (c <= FILTER_SMEM_BINS) atomicAdd(&sh[c - 1u], 1u);
            else if (c <= hcap) atomicAdd(&hist[c - 1u], 1u);
            else ++over;
            mx = max(mx, c);
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < FILTER_SMEM_BINS; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
    mx = __reduce_max_sync(0xffffffffu, mx);
    over = __reduce_add_sync(0xffffffffu, over);
    if ((threadIdx.x & 31u) == 0u) {
        if (mx) atomicMax(&meta[1], mx);
        if (over) atomicAdd(&meta[0], over);
    }
}
// One block: the first `limit` rows that pass, ascending.  meta[2] = how many were written.
__global__ void __launch_bounds__(1024)
filter_select_kernel(const uint32_t *__restrict__ cnt, const uint8_t *__restrict__ ok, uint32_t n, int use_ok, int abun,
                     uint32_t lo, uint32_t hi, uint32_t limit, uint32_t *__restrict__ idx_out, uint32_t *meta) {
    __shared__ uint32_t wsum[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    uint32_t base = 0;
    for (uint32_t tile = 0; tile < n && base < limit; tile += 1024u) {
        const uint32_t i = tile + tid;
        bool f = i < n;
        if (f && use_ok) f = ok[i] != 0;
        if (f && abun) { const uint32_t c = cnt[i]; f = lo <= c && c <= hi; }
        const uint32_t b = __ballot_sync(0xffffffffu, f);
        __syncthreads();                       // wsum from the previous tile
        if (lane == 0u) wsum[wid] = __popc(b);
        __syncthreads();
        uint32_t before = 0, total = 0;
        for (uint32_t w = 0; w < 32u; ++w) { const uint32_t v = wsum[w]; if (w < wid) before += v; total += v; }
        const uint32_t pos = base + before + __popc(b & ((1u << lane) - 1u));
        if (f && pos < limit) idx_out[pos] = i;
        base += total;
    }
    if (tid == 0u) meta[2] = min(base, limit);
}
void launch_filter_pass(const uint32_t *cnt, const uint32_t *ext, uint32_t n, int strand, double cut, int err, uint32_t *hist,
                        uint32_t hcap, uint8_t *ok, uint32_t *meta, cudaStream_t s) {
    if (!n) return;
    filter_pass_kernel<<<min(cdiv(n, 256), 296u), 256, 0, s>>>(cnt, ext, n, strand, cut, err, hist, hcap, ok, meta);
}
void launch_filter_select(const uint32_t *cnt, const uint8_t *ok, uint32_t n, int use_ok, int abun, uint32_t lo, uint32_t hi,
                          uint32_t limit, uint32_t *idx_out, uint32_t *meta, cudaStream_t s) {
    filter_select_kernel<<<1, 1024, 0, s>>>(cnt, ok, n, use_ok, abun, lo, hi, limit, idx_out, meta);
}

// ---- AllCountsSketcher (lib/src/sketch_schemes/counts.rs:7-70): `--sketch-type none` ---------------------------------
// counts[ix] over ALL 4^k forward k-mers, ix = the k-mer's 2-bit codes MSB-first (needletail's BitKmer: A, C, G, T =
// 0..3, first base in the highest pair), one saturating increment per valid window (bit_kmers(k, false): windows
// holding anything but ACGT are skipped, like canonical_kmers).  A thread walks one hash piece of the parse kernel's
// plan straight out of the symbol region (this is not the headline path: no staging, no persistent blocks).
__global__ void count_kmers_kernel(const uint8_t *__restrict__ symbuf, ChunkGeom g, uint32_t n_regions, PiecePlan pp, uint32_t k,
                                   uint32_t *__restrict__ counts) {
    const uint64_t mask = k >= 32u ? ~0ULL : ((1ULL << (2u * k)) - 1ULL);
    const uint64_t items = (uint64_t)n_regions * pp.stride;
    for (uint64_t item = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; item < items; item += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t region = (uint32_t)(item / pp.stride), j = (uint32_t)(item % pp.stride);
        if (j >= pp.count[region]) continue;
        const uint32_t piece = pp.table[(size_t)region * pp.stride + j];
        const int p0 = (int)(piece & 0xFFFFu), n = (int)(piece >> 16);
        const uint8_t *sym = symbuf + (size_t)SYM_FRONT + (size_t)region * g.region_stride;   // sym[-1 .. -halo]: what precedes the region
        uint64_t fwd = 0;
        uint32_t run = 0;
        for (int i = p0 - (int)k + 1; i < p0 + n; ++i) {
            const uint32_t c = sym[i];
            if (c < 4u) { fwd = ((fwd << 2) | c) & mask; ++run; } else run = 0;
            if (i >= p0 && run >= k) {
                const uint32_t old = atomicAdd(&counts[fwd], 1u);
                if (old == 0xFFFFFFFFu) atomicSub(&counts[fwd], 1u);      // saturating_add: a wrapping add undoes itself
            }
        }
    }
}
__device__ __forceinline__ uint64_t revcomp_msb(uint64_t ix, uint32_t k) {       // needletail bitkmer::reverse_complement
    uint64_t x = ~ix;
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    x = ((x >> 8) & 0x00FF00FF00FF00FFULL) | ((x & 0x00FF00FF00FF00FFULL) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFULL) | ((x & 0x0000FFFF0000FFFFULL) << 16);
    x = (x >> 32) | (x << 32);
    return x >> (64u - 2u * k);
}
// to_vec (counts.rs:38-63): walking ix upwards, an entry is made for ix when its count is not zero and its reverse
// complement has not been folded into an earlier entry -- i.e. ix <= rc(ix), or counts[rc(ix)] == 0.
__device__ __forceinline__ bool allcounts_keeps(const uint32_t *counts, uint64_t ix, uint32_t k) {
    if (counts[ix] == 0u) return false;
    const uint64_t rc = revcomp_msb(ix, k);
    return ix <= rc || counts[rc] == 0u;
}
constexpr uint32_t AC_BLOCK = 1024;
__global__ void __launch_bounds__(AC_BLOCK)
allcounts_flag_kernel(const uint32_t *__restrict__ counts, uint64_t n, uint32_t k, uint32_t *__restrict__ block_kept,
                      unsigned long long *total) {
    const uint64_t ix = (uint64_t)blockIdx.x * AC_BLOCK + threadIdx.x;
    const bool keep = ix < n && allcounts_keeps(counts, ix, k);
    const uint32_t kept = __syncthreads_count(keep);
    unsigned long long c = ix < n ? counts[ix] : 0u;                      // total_bases_and_kmers: the sum of the counts
    for (int d = 16; d > 0; d >>= 1) c += __shfl_down_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31u) == 0u && c) atomicAdd(total, c);
    if (threadIdx.x == 0) block_kept[blockIdx.x] = kept;
}
__global__ void __launch_bounds__(1024)
allcounts_scan_kernel(uint32_t *block_kept, uint32_t n_blocks, unsigned long long *n_out) {   // exclusive scan in place, one block
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_blocks; base += 1024u) {
        const uint32_t i = base + threadIdx.x;
        const unsigned long long v = i < n_blocks ? block_kept[i] : 0u;
        unsigned long long x = v;
        for (int d = 1; d < 32; d <<= 1) { const unsigned long long y = __shfl_up_sync(0xffffffffu, x, d); if ((threadIdx.x & 31u) >= (uint32_t)d) x += y; }
        if ((threadIdx.x & 31u) == 31u) wsum[threadIdx.x >> 5] = x;
        __syncthreads();
        unsigned long long before = carry;
        for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) before += wsum[w];
        // offsets are written as 32-bit values: more than 2^32 - 1 entries cannot be returned anyway (checked by the host)
        if (i < n_blocks) block_kept[i] = (uint32_t)min(before + x - v, 0xFFFFFFFFull);
        __syncthreads();
        if (threadIdx.x == 1023u) carry = before + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = carry;
}
__global__ void __launch_bounds__(AC_BLOCK)
allcounts_emit_kernel(const uint32_t *__restrict__ counts, uint64_t n, uint32_t k, const uint32_t *__restrict__ block_off,
                      unsigned long long *__restrict__ o_hash, uint32_t *__restrict__ o_cnt, uint32_t *__restrict__ o_ext,
                      uint8_t *__restrict__ o_kmer) {
    __shared__ uint32_t wsum[32];
    const uint64_t ix = (uint64_t)blockIdx.x * AC_BLOCK + threadIdx.x;
    const bool keep = ix < n && allcounts_keeps(counts, ix, k);
    const uint32_t b = __ballot_sync(0xffffffffu, keep);
    if ((threadIdx.x & 31u) == 0u) wsum[threadIdx.x >> 5] = __popc(b);
    __syncthreads();
    uint32_t before = block_off[blockIdx.x];
    for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) before += wsum[w];
    if (!keep) return;
    const size_t at = (size_t)before + __popc(b & ((1u << (threadIdx.x & 31u)) - 1u));
    const uint32_t extra = counts[revcomp_msb(ix, k)];
    o_hash[at] = ix;
    o_ext[at] = extra;
    o_cnt[at] = counts[ix] + extra;                                       // `count += extra_count`
    for (uint32_t i = 0; i < k; ++i) o_kmer[at * k + i] = (uint8_t)"ACGT"[(ix >> (2u * (k - 1u - i))) & 3u];   // bitmer_to_bytes
}
void launch_count_kmers(const uint8_t *symbuf, ChunkGeom g, uint32_t n_regions, PiecePlan pp, uint32_t k, uint32_t *counts, cudaStream_t s) {
    if (!n_regions) return;
    const uint64_t items = (uint64_t)n_regions * pp.stride;
    count_kmers_kernel<<<(uint32_t)std::min<uint64_t>((items + 255) / 256, 148u * 32u), 256, 0, s>>>(symbuf, g, n_regions, pp, k, counts);
}
uint32_t allcounts_blocks(uint64_t n) { return (uint32_t)((n + AC_BLOCK - 1) / AC_BLOCK); }
void launch_allcounts_plan(const uint32_t *counts, uint64_t n, uint32_t k, uint32_t *block_off, unsigned long long *meta /* [0] entries, [1] sum */,
                           cudaStream_t s) {
    const uint32_t nb = allcounts_blocks(n);
    allcounts_flag_kernel<<<nb, AC_BLOCK, 0, s>>>(counts, n, k, block_off, meta + 1);
    allcounts_scan_kernel<<<1, 1024, 0, s>>>(block_off, nb, meta);
}
void launch_allcounts_emit(const uint32_t *counts, uint64_t n, uint32_t k, const uint32_t *block_off, unsigned long long *o_hash,
                           uint32_t *o_cnt, uint32_t *o_ext, uint8_t *o_kmer, cudaStream_t s) {
    allcounts_emit_kernel<<<allcounts_blocks(n), AC_BLOCK, 0, s>>>(counts, n, k, block_off, o_hash, o_cnt, o_ext, o_kmer);
}

}  // namespace fb2

// device_types.cuh -- structures shared between the kernels and the host engine.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fb2 {

constexpr int TILE_THREADS = 256;   // parse: threads per tile
constexpr int TILE_BYTES = 4096;    // parse: raw bytes per tile (16 per thread)
constexpr int SYM_FRONT = 256;      // pad in front of every symbol region: holds the `halo` symbols that precede it in stream order
constexpr int HALO_SMALL = 32;      // halo (symbols carried across region / chunk seams) for k <= 32
constexpr int HALO_BIG = 256;       // ... for 33 <= k <= 255 (k - 1 <= 254 symbols)
constexpr int KMER_WORDS_MAX = 8;   // 64-bit words of 2-bit codes per k-mer: 1 for k <= 32, ceil(k / 32) otherwise
constexpr int HASH_THREADS = 256;
constexpr int HASH_W = 64;          // k-mer end positions per thread
constexpr uint32_t HASH_TILE = HASH_THREADS * HASH_W;  // symbols per block

enum ParseMode : int { MODE_LINES = 0, MODE_FASTA = 1, MODE_FASTQ = 2 };

// Geometry of one chunk: tiles of 4 KiB grouped into supertiles; every supertile owns a symbol
// region of st_bytes (+ SYM_FRONT pad in front of it) in the symbol buffer.
struct ChunkGeom {
    uint32_t len;            // raw bytes in the chunk
    uint32_t n_tiles;        // ceil(len / TILE_BYTES)
    uint32_t st_tiles;       // tiles per supertile
    uint32_t n_st;           // supertiles
    uint32_t st_bytes;       // st_tiles * TILE_BYTES == region capacity in symbols
    uint32_t region_stride;  // st_bytes + SYM_FRONT
    uint32_t hash_tiles;     // hash blocks per region = ceil(st_bytes / HASH_TILE)
};

// Parser + stream state that persists across chunks (device resident; mirrored to the host
// after every chunk).  All positions are global (stream) offsets.
struct ParseCarry {
    uint32_t state;          // state of the line containing the next byte
    uint32_t prev1, prev2;   // last / second-to-last raw byte of the stream so far
    uint32_t error;          // sticky device-side error bits
    uint64_t raw_total;      // raw bytes consumed
    uint64_t total_bases;    // Sum of record sequence().len() (wrapping arithmetic)
    uint64_t n_records;
    uint64_t first_bad_pos;  // FASTQ: min raw pos of a line start failing the '@' / '+' check
    uint64_t last_sig;       // FASTQ: max ((raw pos + 1) << 2 | phase) over bytes that are not CR/LF
    uint64_t chunk_raw_base; // raw_total before the current chunk
    uint32_t chunk_syms;     // symbols produced by the current chunk (all regions)
    uint32_t cprev1, cprev2; // prev1 / prev2 as they were at the start of the current chunk
    uint32_t max_region_syms;// largest region of the current chunk (symbols): bounds the hash kernel's item space
    // FASTQ: needletail rejects a record whose sequence and quality lines differ in length (lib.rs:63 panics).
    uint64_t len_bad_pos;    // min stream position of the header-line newline of such a record (~0: none)
    uint64_t last_nl[2][3];  // the last three newlines of the stream before / after the current chunk (ascending;
                             // stream position, bit 63 = preceded by a CR; ~0 = none); halves alternate per chunk
    uint32_t state_next;     // fused parse: line state after the current chunk (its last supertile writes it while other
                             // blocks may still read `state`; front_fix_kernel commits it to `state`)
    uint32_t max_region_pieces;  // largest hash-piece count of a region of the current chunk: bounds the hash kernel's item space
};
// Hash pieces: what a lane of the hash kernel walks.  The parse kernel lists, per region, runs of k-mer END positions
// (p0 | n << 16: positions [p0, p0 + n) of the region) that together hold every position a valid k-mer can end at,
// each exactly once.  For records with short sequences (reads) a run never starts inside the k - 1 positions after a
// record break -- those windows all hold the break -- and a record's positions are cut into runs of nearly equal
// length <= pmax, so that 32 consecutive runs span at most PIECE_SPAN symbols (= what a warp stages at once);
// otherwise runs are the uniform 64-position slices of the region.
constexpr uint32_t PIECE_STRIDE = 1024;     // table entries per region (uniform slices need 512)
constexpr uint32_t PIECE_SPAN = 2144;       // 32 * (pmax + k): bound on the symbols under 32 consecutive pieces of records
struct PiecePlan {
    uint32_t *table;         // n_st * stride entries (nullptr: no plan, k > 32)
    uint32_t *count;         // pieces per region
    uint32_t stride;         // PIECE_STRIDE
    uint32_t k;              // k-mer length
    uint32_t pmax;           // longest run of a record: PIECE_SPAN / 32 - k
    uint32_t pmax_inv;       // ceil(2^22 / pmax): x / pmax == (x * pmax_inv) >> 22 for x < 2^15 + pmax
};
constexpr unsigned long long NL_NONE = ~0ULL;

// FASTQ length check across supertile seams: the first and the last (up to) three newlines of a supertile,
// chunk-relative position, bit 31 = the byte before the newline is a CR.  n = newlines in the supertile.
struct SeamNl { uint32_t first[3], last[3]; uint32_t n, pad; };

// Counters of one hash launch (two slots: chunk c+1 may be hashed while chunk c is verified).
struct LaunchSlot {
    unsigned long long launch_kmers;  // valid k-mers seen by the launch
    unsigned int log_count;           // entries appended to the slot's log (may exceed its capacity)
    unsigned int decision;            // absorb_decide_kernel: 0 none, 1 absorbed, 2 log overflowed, 3 table too full
    unsigned int chunk_syms;          // symbols of the chunk (copied from ParseCarry for the host)
    unsigned int next_item;           // hash_kernel: work-item counter of the launch (warps claim items dynamically)
};
constexpr unsigned int DECIDE_GO = 1, DECIDE_OVERFLOW = 2, DECIDE_FULL = 3;

// Sketch state (device resident; mirrored to the host when the host needs to decide).
struct SketchState {
    unsigned long long threshold;     // admit hash <= threshold (only ever decreases)
    unsigned long long total_kmers;   // unused on device (the host commits per-launch counts)
    LaunchSlot slot[2];               // per-launch counters, one per in-flight chunk
    unsigned int occupied;            // table slots in use (excl. the u64::MAX side slot)
    unsigned int has_max_key;         // side slot for hash == u64::MAX in use
    unsigned int gather_count;
    unsigned int keep_count;          // result of select_keep
    unsigned int hist_shift;          // live_bins[key >> hist_shift] counts the keys inserted into the current table
    unsigned long long new_threshold;
    unsigned int *live_bins;          // 4096 bins (device), maintained by table_upsert; nullptr = off
};

struct LogView {
    unsigned long long *hash;   // murmur h1
    unsigned long long *kmer;   // kw words per entry: LSB-first 2-bit codes (base i in bits [2i, 2i+1] of the multi-word),
                                // or the arena index of a pushed k-mer in word 0
    unsigned long long *posx;   // position id << 9 | is_arena << 8 | extra_count (u8)
    unsigned int cap;
    unsigned int kw;            // words per k-mer (1 for k <= 32)
};
struct TableView {
    unsigned long long *key;    // EMPTY_KEY when free; slot `cap` is the side slot for u64::MAX
    unsigned long long *cnt;
    unsigned long long *ext;
    unsigned long long *posx;   // min posx over occurrences (first occurrence wins the kmer)
    unsigned long long *kmer;   // kw words per slot
    unsigned int cap;           // power of two
    unsigned int shift;         // 64 - log2(cap)
    unsigned int kw;            // words per k-mer (1 for k <= 32)
};
constexpr unsigned long long EMPTY_KEY = ~0ULL;

}  // namespace fb2

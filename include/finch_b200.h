/*
 * finch_b200.h -- C ABI of the B200-native MinHash sketching engine (libfinch_b200.so).
 *
 * Drop-in boundary for ONE path of onecodex/finch-rs (reference @ bb481c0):
 *   FASTA/FASTQ bytes -> canonical k-mers -> murmurhash3_x64_128(kmer, seed).0
 *   -> bottom-s distinct hashes with count / extra_count (MashSketcher / ScaledSketcher),
 *   the host filter that follows it, and `dist`'s sorted-hash intersection.
 *
 * The reference has no FFI; its operator boundary is the Rust trait `SketchScheme`
 * (lib/src/sketch_schemes/mod.rs:24-51), the factory `SketchParams::create_sketcher`
 * (mod.rs:86-113), `sketch_files` / `sketch_stream` (lib/src/lib.rs:29-94) and
 * `distance` / `raw_distance` (lib/src/distance.rs:9-126).  Each entry point below names the
 * reference item it replaces; INTEGRATION.md shows the Rust `extern "C"` binding and the
 * `impl SketchScheme` shim a maintainer would add.
 *
 * Conventions: plain pointers and sizes only; every call returns FB2_OK (0) or a negative
 * FB2_E* code and never aborts/unwinds across the ABI; fb2_last_error() returns a thread-local
 * message.  All compute runs in hand-written sm_100a CUDA kernels; there is NO CPU fallback:
 * without a CUDA device every compute entry point returns FB2_ECUDA.
 * Thread-safety: distinct handles may be used from distinct threads (as rayon does with one
 * sketcher per file, lib.rs:36-46); a single handle is not re-entrant (`&mut self`).
 */
#ifndef FINCH_B200_H
#define FINCH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FB2_OK 0
#define FB2_EINVAL (-1)       /* bad argument */
#define FB2_ECUDA (-2)        /* CUDA runtime failure / no device */
#define FB2_EFORMAT (-3)      /* first byte neither '>' nor '@'      (lib.rs:60 panics) */
#define FB2_ERECORD (-4)      /* invalid / truncated FASTQ record    (lib.rs:63 panics) */
#define FB2_EEMPTY (-5)       /* no records in the stream            (lib.rs:72 panics) */
#define FB2_ETOOFEW (-6)      /* "<name> had too few kmers (<n>) to sketch" (mod.rs:115-128) */
#define FB2_EUNSUPPORTED (-7) /* bz2 / xz input, compressed bytes handed to the raw feed calls */
#define FB2_EIO (-8)          /* file could not be opened / read     (lib.rs:43) */
#define FB2_ENOMEM (-9)

#define FB2_KIND_MASH 0   /* SketchParams::Mash   (mod.rs:55-61) */
#define FB2_KIND_SCALED 1 /* SketchParams::Scaled (mod.rs:62-67) */
#define FB2_KIND_ALLCOUNTS 2 /* SketchParams::AllCounts (mod.rs:68-70): counts of all 4^k k-mers (sketch_schemes/counts.rs), k <= 16;
                              * process / feed / totals / result / sketch only (the reference type has no push) */

#define FB2_FORMAT_UNKNOWN 0
#define FB2_FORMAT_FASTA 1 /* needletail::parser::Format::Fasta */
#define FB2_FORMAT_FASTQ 2 /* needletail::parser::Format::Fastq */

/* SketchParams (mod.rs:53-71) + placement. */
typedef struct fb2_params {
    int32_t kind;             /* FB2_KIND_* */
    uint64_t kmers_to_sketch; /* Mash: heap size; Scaled: fill-up size (0 = pure scaled) */
    uint64_t final_size;      /* Mash only: process_post_filter truncation */
    int32_t no_strict;        /* Mash only */
    uint8_t kmer_length;      /* 1..=255 like the reference (mod.rs:59); 33..=255 run the exact multi-word kernel */
    uint64_t hash_seed;
    double scale;             /* Scaled only */
    int32_t device;           /* CUDA ordinal; -1 = current device */
    void *stream;             /* cudaStream_t to launch on; NULL = library-owned stream */
} fb2_params;

/* FilterParams (lib/src/filtering.rs:11-16). */
typedef struct fb2_filter {
    int32_t filter_on;     /* -1 = None (auto by format, lib.rs:71-76), 0 = Some(false), 1 = Some(true) */
    int32_t has_abun_low;  uint32_t abun_low;   /* abun_filter.0 */
    int32_t has_abun_high; uint32_t abun_high;  /* abun_filter.1 */
    double err_filter;     /* already multiplied by k/100 as cli.rs:264-265 does */
    double strand_filter;
} fb2_filter;

/* Vec<KmerCount> (mod.rs:15-22) as library-owned SoA, ascending by hash; plus the Sketch
 * metadata sketch_stream fills (lib.rs:84-93).  Free with fb2_result_free. */
typedef struct fb2_result {
    uint64_t n;
    uint64_t *hashes;      /* n */
    uint32_t *counts;      /* n, saturating u32 (mash.rs:48) */
    uint32_t *extras;      /* n, saturating u32 (mash.rs:49) */
    uint8_t *kmers;        /* n * kmer_stride bytes: canonical k-mer of the first occurrence */
    uint32_t kmer_stride;
    uint64_t seq_length;       /* total_bases  (mash.rs:72,82-84) */
    uint64_t num_valid_kmers;  /* total_kmers  (mash.rs:35) */
    int32_t format;            /* FB2_FORMAT_* seen by the parser */
    fb2_filter filters;        /* FilterParams as updated by filter_counts (sketch_* calls only) */
    uint32_t *kmer_lens;       /* NULL: every k-mer is kmer_length bytes.  Else n byte lengths: entries that came through
                                  fb2_sketcher_push keep the caller's bytes, of any length up to kmer_stride (mash.rs:52-55) */
} fb2_result;

typedef struct fb2_sketcher fb2_sketcher;

/* ---- SketchParams::create_sketcher / MashSketcher::new / ScaledSketcher::new ------------- */
/* (mod.rs:86-113, mash.rs:21, scaled.rs:22) */
int fb2_sketcher_create(const fb2_params *p, fb2_sketcher **out);
/* Drop */
void fb2_sketcher_destroy(fb2_sketcher *s);
/* Forget all state but keep device buffers (one handle re-used across files). */
int fb2_sketcher_reset(fb2_sketcher *s);

/* ---- SketchScheme::process (mod.rs:25-28; mash.rs:67-80, scaled.rs:65-78) ---------------- */
/* One record's RAW, un-normalised sequence bytes.  Batches internally; the GPU work happens
 * when a staging chunk fills or on totals/result. */
int fb2_sketcher_process(fb2_sketcher *s, const uint8_t *seq, size_t len);

/* ---- MashSketcher::push / ScaledSketcher::push (mash.rs:34, scaled.rs:37) ---------------- */
/* Unit-test surface: hashes the given bytes as they are (any bytes, any length <= 255). */
int fb2_sketcher_push(fb2_sketcher *s, const uint8_t *kmer, size_t k, uint8_t extra_count);

/* ---- bulk replacement of the record loop lib.rs:60-68 ------------------------------------ */
/* Raw FASTA/FASTQ file bytes in arbitrary pieces (seams anywhere); `final` != 0 on the last
 * piece.  `bytes` may be pageable or pinned host memory. */
int fb2_sketcher_feed_fastx(fb2_sketcher *s, const uint8_t *bytes, size_t len, int final);
/* Same, bytes already resident in device memory (HBM) of the sketcher's device. */
int fb2_sketcher_feed_device(fb2_sketcher *s, const uint8_t *dev_bytes, size_t len, int final);
/* needletail SequenceRecord::format() of the stream fed so far (lib.rs:64-66). */
int fb2_sketcher_format(fb2_sketcher *s, int32_t *format);

/* ---- SketchScheme::total_bases_and_kmers (mash.rs:82-84) --------------------------------- */
int fb2_sketcher_totals(fb2_sketcher *s, uint64_t *total_bases, uint64_t *total_kmers);
/* ---- SketchScheme::to_vec (mash.rs:86-102, scaled.rs:84-100) ----------------------------- */
/* Non-destructive (the trait takes &self): more input may follow. */
int fb2_sketcher_result(fb2_sketcher *s, fb2_result *out);
void fb2_result_free(fb2_result *r);
/* ---- the tail of sketch_stream (lib.rs:78-93) in one call ----------------------------------- */
/* to_vec + FilterParams::filter_counts + SketchParams::process_post_filter with identical results,
 * but only the surviving entries cross PCIe (counts/extras of all entries do, for the filter).
 * `p` supplies final_size / no_strict; `name` is used in the "too few kmers" message. */
int fb2_sketcher_sketch(fb2_sketcher *s, const char *name, const fb2_params *p, const fb2_filter *f,
                        fb2_result *out);

/* Counters for benchmarking: kernels launched and bytes moved by this handle so far. */
typedef struct fb2_stats {
    uint64_t kernel_launches, h2d_bytes, d2h_bytes, chunks, prunes, hash_launches;
    double hash_kernel_ms;   /* CUDA-event time of the k-mer hash kernel (only if timing enabled) */
    double parse_kernel_ms;  /* CUDA-event time of the parse/pack kernels */
    uint64_t hash_symbols;   /* symbols (bases + record breaks) the hash kernel walked */
    uint64_t provisional_redos; /* chunks redone because the provisional first threshold was too low */
    uint64_t band_passes;       /* passes of the banded absorb over a candidate log */
} fb2_stats;
int fb2_sketcher_stats(fb2_sketcher *s, fb2_stats *out);
/* The same counters for the most recent fb2_sketch_stream call of the calling thread (summed over the handles it
 * used: the two-ended mode FB2_HOST_STRIP=2 drives two). */
int fb2_last_stream_stats(fb2_stats *out);
/* Inspection hook for tests: geometry (7 x u32), per-region symbol counts and the raw symbol buffer of the
 * most recent chunk (see finch_rs_b200/csrc/parse.cu for the layout). */
int fb2_sketcher_debug_symbols(fb2_sketcher *s, uint32_t *geom7, uint32_t *counts, size_t counts_cap,
                               uint8_t *sym, size_t sym_cap);
int fb2_sketcher_enable_timing(fb2_sketcher *s, int on);
/* Test hook: add to the 64-bit totals kept for `hash` (must be present), to reach the u32 saturation of
 * mash.rs:48-49 without 2^32 pushes. */
int fb2_sketcher_debug_bump(fb2_sketcher *s, uint64_t hash, uint64_t add_count, uint64_t add_extra);

/* ---- FilterParams::filter_counts + SketchParams::process_post_filter ---------------------- */
/* (filtering.rs:60-87, mod.rs:115-128).  In place on `r`; `f` is updated like the reference
 * (filter_on resolved from `format` when -1; abun_low raised to the guessed cutoff). */
int fb2_filter_counts(fb2_result *r, fb2_filter *f);
int fb2_process_post_filter(fb2_result *r, const fb2_params *p, const char *name);
/* statistics.rs:30-47 / filtering.rs:154-195 (exposed for parity tests) */
uint32_t fb2_guess_filter_threshold(const uint32_t *counts, size_t n, double filter_level);

/* ---- sketch_stream (lib.rs:51-94) over a host buffer -------------------------------------- */
/* gzip / bzip2 / xz bytes are decompressed on the host first, as the reader sketch_stream goes through does
 * (needletail sniffs the first bytes); the handle-level feed calls above take plain FASTA / FASTQ bytes only. */
int fb2_sketch_stream(const uint8_t *bytes, size_t len, const char *name, const fb2_params *p,
                      const fb2_filter *f, fb2_result *out);
/* ---- sketch_files (lib.rs:29-49): outs[i] <- paths[i], input order; "-" is stdin ---------- */
int fb2_sketch_files(const char *const *paths, size_t n, const fb2_params *p, const fb2_filter *f,
                     fb2_result *outs);

/* ---- the same over several GPUs of this process ------------------------------------------------ */
/* sketch_files with the files sharded over `ngpus` devices (0 = all visible; devices 0..ngpus-1): longest-first
 * assignment by file size, a few worker handles per GPU, results in input order (the rayon fan-out of lib.rs:34-36
 * across GPUs).  fb2_sketch_files == ngpus 1 on p->device. */
int fb2_sketch_files_multi(const char *const *paths, size_t n, const fb2_params *p, const fb2_filter *f,
                           fb2_result *outs, int ngpus);
/* sketch_stream of ONE file cut into byte ranges over `ngpus` devices: every range is parsed and sketched on its own
 * GPU, the tables are united exactly on the first GPU over peer memory, then the usual filter / truncate tail.
 * Bit-identical to fb2_sketch_stream (hashes, counts, extra counts, first-occurrence k-mers, totals). */
int fb2_sketch_stream_multi(const uint8_t *bytes, size_t len, const char *name, const fb2_params *p,
                            const fb2_filter *f, fb2_result *out, int ngpus);

/* The sketch_* calls keep idle worker handles (device buffers, a pinned read buffer sized by the files read so far) for
 * the next call with the same parameters, bounded per device: at most FB2_POOL_MAX handles (default 16; 0 = none) holding
 * at most FB2_POOL_MB MiB of device memory together (default 4096).  This frees them. */
void fb2_sketch_files_release_pool(void);

/* ---- raw_distance (distance.rs:66-126), integer part, batched ----------------------------- */
typedef struct fb2_pair_out {
    uint32_t common; /* |A n B| up to the stopping point */
    uint32_t i;      /* query hashes consumed */
    uint32_t j;      /* reference hashes consumed */
} fb2_pair_out;
/* hashes: n_sk sketches, sketch s occupies hashes[s*stride .. s*stride+lens[s]), STRICTLY ascending (what to_vec
 * returns: the keys of a hash map, sorted; the merge loop's behaviour on repeated hashes is not reproduced).
 * For pair p: query q_idx[p], reference r_idx[p].  scale as raw_distance's (0 = none).
 * The f64 epilogue (containment, jaccard, mash distance) stays on the host: fb2_distance_finish. */
int fb2_dist_batch(const uint64_t *hashes, const uint32_t *lens, size_t n_sk, size_t stride,
                   double scale, const uint32_t *q_idx, const uint32_t *r_idx, size_t n_pairs,
                   fb2_pair_out *out, int32_t device);
/* All ordered pairs (q, r), q in [q0,q1), r in [0,n_sk): out[(q-q0)*n_sk + r].  `out` is host memory; a pinned
 * array (cudaHostAlloc / cudaHostRegister) receives the device copies directly, pageable memory goes through staging. */
int fb2_dist_all_pairs(const uint64_t *hashes, const uint32_t *lens, size_t n_sk, size_t stride,
                       double scale, size_t q0, size_t q1, fb2_pair_out *out, int32_t device);
/* ---- calc_sketch_distances with the max_distance cut (cli/src/main.rs:315-334) --------------------------- */
/* One surviving pair: query row q, reference row r and raw_distance's integers. */
typedef struct fb2_pair_hit {
    uint32_t q, r, common, i, j;
} fb2_pair_hit;
/* All ordered pairs (q in [q0,q1), r in [0,n_sk)) whose mash distance (distance.rs:36-41, k = kmer_length) can be
 * <= max_distance, ascending by (q, r).  The device applies the cut conservatively (jaccard bound slightly below the
 * exact one) and compacts the survivors; the caller finishes each with fb2_distance_finish and applies the exact
 * `mash_distance <= max_distance` test of main.rs:328 -- nothing that test keeps is missing.  skip_self drops q == r.
 * `hits` has room for `cap` entries; *n_hits receives the number found -- when it exceeds cap the call returns
 * FB2_ENOMEM with *n_hits = the required capacity and the first `cap` hits written.
 * ngpus > 1 (0 = all): the query rows are cut into one block per GPU, the hash matrix is loaded once and handed to
 * the other GPUs over peer copies (replicas + row blocks). */
int fb2_dist_all_pairs_cut(const uint64_t *hashes, const uint32_t *lens, size_t n_sk, size_t stride, double scale,
                           size_t q0, size_t q1, uint8_t kmer_length, double max_distance, int skip_self,
                           fb2_pair_hit *hits, size_t cap, uint64_t *n_hits, int32_t device, int ngpus);
/* Measurement aid: summed device time (ms) of the kernels of this thread's last fb2_dist_all_pairs. */
double fb2_dist_last_kernel_ms(void);
/* distance.rs:117-125 and :35-41 from the integers of one pair. */
void fb2_distance_finish(const fb2_pair_out *p, uint8_t kmer_length, double *containment,
                         double *jaccard, double *mash_distance, uint64_t *common_hashes,
                         uint64_t *total_hashes);

/* minmer_matrix (distance.rs:344-364): result[i * n_ref + c] = the count sketch i holds for reference hash c, 0 when it
 * does not hold it.  Sketch i = sk_hashes / sk_counts[sk_off[i] .. sk_off[i + 1]); all hash lists ascending. */
int fb2_minmer_matrix(const uint64_t *ref_hashes, size_t n_ref, const uint64_t *sk_hashes, const uint32_t *sk_counts,
                      const uint64_t *sk_off, size_t n_sk, int32_t *result, int32_t device);

/* old_distance (distance.rs:136-157, `--old-dist`) from `common` of a scale-0 fb2_dist_batch pair and the two
 * sketch lengths; FB2_EINVAL where the reference panics (empty query, non-empty reference). */
int fb2_old_distance_finish(uint64_t common, uint64_t query_len, uint64_t ref_len, uint8_t kmer_length,
                            double *containment, double *jaccard, double *mash_distance,
                            uint64_t *common_hashes, uint64_t *total_hashes);

/* ---- sketch files ---------------------------------------------------------------------------
 * Collections of finished sketches in the reference's three file formats; host only (no device is needed).
 *   open  <- open_sketch_file (lib/src/lib.rs:96-117): the file SUFFIX decides -- *.msh Mash Cap'n Proto
 *            (serialization/mash.rs:73-132), *.bsk finch Cap'n Proto (serialization/mod.rs:178-224), *.sk / *.json
 *            Mash-compatible JSON (serialization/json.rs); anything else: "File suffix is not *.bsk, *.msh, or *.sk"
 *   save  <- write_finch_file (serialization/mod.rs:123-176; what Multisketch.save of the Python module writes,
 *            python.rs:180-186), write_mash_file (mash.rs:12-71), MultiSketch JSON (json.rs:64-89)
 * This is what the Python mirror of the reference's `finch` module (finch_rs_b200/pyfinch.py) is built on. */
#define FB2_FILE_SK 0  /* .sk  */
#define FB2_FILE_BSK 1 /* .bsk */
#define FB2_FILE_MSH 2 /* .msh */
typedef struct fb2_sketch_set fb2_sketch_set;
/* `Sketch` (serialization/mod.rs:45-55).  From fb2_sketch_set_get the pointers belong to the set and stay valid
 * until its next _get / _add / _remove / _close. */
typedef struct fb2_sketch_view {
    const char *name;
    const char *comment;
    uint64_t seq_length;
    uint64_t num_valid_kmers;
    fb2_params params;         /* sketch_params (device / stream unused) */
    fb2_filter filter;         /* filter_params */
    uint64_t n;                /* hashes.len() */
    const uint64_t *hashes;    /* ascending */
    const uint32_t *counts;
    const uint32_t *extras;
    const uint8_t *kmers;      /* k-mer bytes back to back (files may hold none: all offsets equal) */
    const uint64_t *kmer_offs; /* n + 1 offsets into kmers */
} fb2_sketch_view;
int fb2_sketch_set_new(fb2_sketch_set **out);
int fb2_sketch_set_open(const char *path, fb2_sketch_set **out);
uint64_t fb2_sketch_set_len(const fb2_sketch_set *set);
int fb2_sketch_set_get(const fb2_sketch_set *set, uint64_t i, fb2_sketch_view *out);
int fb2_sketch_set_add(fb2_sketch_set *set, const fb2_sketch_view *v); /* copies */
int fb2_sketch_set_remove(fb2_sketch_set *set, uint64_t i);
int fb2_sketch_set_save(const fb2_sketch_set *set, const char *path, int file_format);
void fb2_sketch_set_close(fb2_sketch_set *set);

/* ---- misc --------------------------------------------------------------------------------- */
const char *fb2_last_error(void);
int fb2_device_count(void);
const char *fb2_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FINCH_B200_H */

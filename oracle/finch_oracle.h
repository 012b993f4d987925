/*
 * finch_oracle.h -- CPU ORACLE for the finch-rs MinHash sketching hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.  The
 * product (finch_rs_b200/libfinch_b200.so) never links, loads or calls anything here.
 *
 * It restates, in plain C, the algorithm of onecodex/finch-rs @ bb481c0 for the path
 *   FASTA/FASTQ bytes -> canonical k-mers -> murmurhash3_x64_128(kmer, seed).0
 *   -> bottom-s distinct hashes with count / extra_count  (MashSketcher / ScaledSketcher)
 *   -> filter_counts -> process_post_filter ;  plus raw_distance / distance.
 *
 * Two third-party crates carry arithmetic on this path and are NOT under /root/reference:
 *   murmurhash3 0.0.5 (Cargo.lock:471-474)  -- restated from the published MurmurHash3_x64_128
 *   needletail  0.5.0 (Cargo.lock:490-502)  -- normalize / reverse_complement / canonical_kmers /
 *                                              FASTX record rules restated from the crate's
 *                                              documented behaviour.
 * Parity is PINNED on the reference's own golden vectors (tests/test_oracle_kats.py):
 *   mash.rs:115-154, scaled.rs:118-213, distance.rs:176-242, filtering.rs:197-505,
 *   statistics.rs:53-129, cli/tests/test_cli.rs:80-149 (+ cli/tests/data/query.fa).
 * Unpinned by any reference test (this oracle is then the only authority):
 *   seq_length semantics, FASTQ record rules, IUPAC/whitespace normalisation, palindromes.
 */
#ifndef FINCH_ORACLE_H
#define FINCH_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- hashing.rs:9-12 -> murmurhash3 0.0.5 -------------------------------------------- */
void fo_murmur3_x64_128(const uint8_t *data, size_t len, uint64_t seed, uint64_t out[2]);
uint64_t fo_hash_f(const uint8_t *item, size_t len, uint64_t seed); /* = .0 (h1) */

/* ---- needletail Sequence::{normalize(false), reverse_complement} --------------------- */
/* out must hold n bytes; returns normalized length. */
size_t fo_normalize(const uint8_t *in, size_t n, uint8_t *out);
void fo_reverse_complement(const uint8_t *in, size_t n, uint8_t *out);

/* Stream of (hash, is_rc) for every canonical k-mer of ONE raw record sequence, in the
 * order needletail's canonical_kmers yields them (mash.rs:72-79).  Returns the count;
 * writes at most cap entries (pass cap=0 / NULL to count only). */
size_t fo_kmer_stream(const uint8_t *seq, size_t len, uint8_t k, uint64_t seed,
                      uint64_t *hashes, uint8_t *is_rc, uint8_t *kmers /* cap*k or NULL */,
                      size_t cap);

/* ---- sketchers: mash.rs:10-113, scaled.rs:10-110 -------------------------------------- */
typedef struct fo_sketcher fo_sketcher;
fo_sketcher *fo_mash_new(size_t size, uint8_t kmer_length, uint64_t seed);
fo_sketcher *fo_scaled_new(size_t size, double scale, uint8_t kmer_length, uint64_t seed);
void fo_sketcher_free(fo_sketcher *s);
void fo_push(fo_sketcher *s, const uint8_t *kmer, size_t klen, uint8_t extra_count);
void fo_process(fo_sketcher *s, const uint8_t *raw_seq, size_t len); /* SketchScheme::process */
void fo_totals(const fo_sketcher *s, uint64_t *total_bases, uint64_t *total_kmers);
uint64_t fo_scaled_max_hash(const fo_sketcher *s);
size_t fo_result_len(const fo_sketcher *s);
/* to_vec(): ascending by hash.  kmers: n * kmer_stride bytes (each entry's kmer, zero padded). */
size_t fo_result(const fo_sketcher *s, uint64_t *hashes, uint32_t *counts, uint32_t *extras,
                 uint8_t *kmers, size_t kmer_stride);

/* ---- FASTX record iteration (needletail parse_fastx_reader semantics) --- */
#define FO_FMT_FASTA 1
#define FO_FMT_FASTQ 2
#define FO_OK 0
#define FO_E_EMPTY -1      /* no records (lib.rs:72 panics "Should have got a type") */
#define FO_E_FORMAT -2     /* first byte neither '>' nor '@' (lib.rs:60 panic) */
#define FO_E_RECORD -3     /* invalid record (lib.rs:63 panic) */
#define FO_E_TOO_FEW -4    /* process_post_filter: "had too few kmers" (mod.rs:115-128) */
typedef void (*fo_record_cb)(void *ctx, const uint8_t *raw_seq, size_t len);
int fo_parse_fastx(const uint8_t *data, size_t len, fo_record_cb cb, void *ctx, int *format,
                   uint64_t *n_records);

/* ---- filtering.rs / statistics.rs ----------------------------------------------------- */
/* hist (statistics.rs:30-47): returns max_count; out[i] = #kmers with count == i+1, i < cap. */
uint64_t fo_hist(const uint32_t *counts, size_t n, uint64_t *out, size_t cap);
uint32_t fo_guess_filter_threshold(const uint32_t *counts, size_t n, double filter_level);
/* The three filters return the number kept and write kept INDICES into keep[]. */
size_t fo_filter_strands(const uint32_t *counts, const uint32_t *extras, size_t n,
                         double ratio_cutoff, uint32_t *keep);
size_t fo_filter_abundance(const uint32_t *counts, size_t n, int has_low, uint32_t low,
                           int has_high, uint32_t high, uint32_t *keep);

typedef struct {
    int filter_on;      /* -1 = None (auto by format), 0 = Some(false), 1 = Some(true) */
    int has_abun_low;   uint32_t abun_low;
    int has_abun_high;  uint32_t abun_high;
    double err_filter;
    double strand_filter;
} fo_filter_params;
/* FilterParams::filter_counts (filtering.rs:60-87); updates fp->abun_low like the reference. */
size_t fo_filter_counts(fo_filter_params *fp, const uint32_t *counts, const uint32_t *extras,
                        size_t n, uint32_t *keep);

/* ---- sketch_stream (lib.rs:51-94) ------------------------------------------------------ */
typedef struct {
    int kind;               /* 0 = Mash, 1 = Scaled */
    uint64_t kmers_to_sketch;
    uint64_t final_size;    /* Mash only */
    int no_strict;          /* Mash only */
    uint8_t kmer_length;
    uint64_t hash_seed;
    double scale;           /* Scaled only */
} fo_sketch_params;
typedef struct {
    uint64_t seq_length, num_valid_kmers;
    size_t n;
    uint64_t *hashes;
    uint32_t *counts, *extras;
    uint8_t *kmers;         /* n * kmer_length */
    fo_filter_params filters; /* as updated (filter_on resolved, minCopies guessed) */
    int format;
} fo_sketch;
int fo_sketch_stream(const uint8_t *data, size_t len, const fo_sketch_params *sp,
                     const fo_filter_params *fp, fo_sketch *out);
void fo_sketch_free(fo_sketch *sk);

/* ---- distance.rs:9-126 ---------------------------------------------------------------- */
int fo_old_distance(const uint64_t *q, size_t nq, const uint64_t *r, size_t nr,
                    double *containment, double *jaccard, uint64_t *common_out, uint64_t *total_out);
void fo_raw_distance(const uint64_t *q, size_t nq, const uint64_t *r, size_t nr, double scale,
                     double *containment, double *jaccard, uint64_t *common, uint64_t *total);
double fo_mash_distance(double jaccard, uint8_t k); /* distance.rs:35-41 */

#ifdef __cplusplus
}
#endif
/* AllCountsSketcher (counts.rs) and minmer_matrix (distance.rs:344-364) */
void fo_allcounts_process(uint32_t *counts, uint8_t k, const uint8_t *raw_seq, size_t len);
size_t fo_allcounts_to_vec(const uint32_t *self_counts, uint8_t k, uint64_t *hashes, uint32_t *cnt, uint32_t *ext,
                           uint8_t *kmers, size_t cap);
uint64_t fo_allcounts_total(const uint32_t *counts, uint8_t k);
void fo_minmer_matrix(const uint64_t *ref_hashes, size_t n_ref, const uint64_t *const *sk_hashes, const uint32_t *const *sk_counts,
                      const size_t *sk_len, size_t n_sk, int32_t *result);

#endif

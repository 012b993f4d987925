// kernels.h -- host-callable launchers of the sm_100a kernels (parse.cu, hash.cu, table.cu, dist.cu).
#pragma once
#include "device_types.cuh"
#include "../../include/finch_b200.h"

namespace fb2 {

// parse.cu
void launch_phase(int mode, const uint8_t *raw, ChunkGeom g, ParseCarry *carry, uint32_t *st_map, uint32_t *st_state,
                  cudaStream_t s);
void launch_pack(int mode, const uint8_t *raw, ChunkGeom g, ParseCarry *carry, const uint32_t *st_state, uint8_t *sym,
                 uint32_t *region_count, const uint8_t *tail_in, uint8_t *tail_out, SeamNl *seam, int nl_in, uint32_t halo,
                 cudaStream_t s);
int parse_fused_max_tiles();
void launch_parse_fused(int mode, const uint8_t *raw, ChunkGeom g, ParseCarry *carry, uint32_t *st_state, uint8_t *sym,
                        uint32_t *region_count, const uint8_t *tail_in, uint8_t *tail_out, SeamNl *seam, int nl_in, uint32_t halo,
                        uint32_t *status, uint32_t epoch, uint32_t *ticket, uint32_t ticket_base, PiecePlan pp, cudaStream_t s);
void launch_uniform_pieces(const uint32_t *region_count, uint32_t n_st, PiecePlan pp, ParseCarry *carry, cudaStream_t s);
void launch_fill_bytes(uint8_t *p, uint32_t n, uint8_t v, cudaStream_t st);

// hash.cu
void launch_hash(int k, const uint8_t *symbuf, ChunkGeom g, uint32_t r0, uint32_t r1, const uint32_t *region_count,
                 const ParseCarry *carry, uint64_t ord_base, const SketchState *st, LaunchSlot *slot, LogView log,
                 uint64_t seed, uint32_t log_reserve, PiecePlan pp, cudaStream_t stream);
void launch_push_hash(const uint8_t *bytes, const uint32_t *offs, const uint8_t *extra, uint32_t n,
                      uint64_t arena_base, uint64_t ord_base, const SketchState *st, LaunchSlot *slot, LogView log,
                      uint64_t seed, cudaStream_t stream);

// table.cu
void launch_absorb(LogView log, uint32_t i0, uint32_t i1, TableView t, SketchState *st, cudaStream_t s);
void launch_absorb_band(LogView log, uint32_t n, TableView t, SketchState *st, unsigned long long lo, int use_lo,
                        unsigned long long hi, cudaStream_t s);
void launch_log_hist(LogView log, uint32_t n, const SketchState *st, uint32_t shift, uint32_t *bins, cudaStream_t s);
void launch_note_chunk_syms(LaunchSlot *slot, const ParseCarry *carry, cudaStream_t s);
void launch_absorb_guarded(LogView log, LaunchSlot *slot, TableView t, SketchState *st, const ParseCarry *carry,
                           uint32_t expect, cudaStream_t s);
void launch_table_clear(TableView t, cudaStream_t s);
void launch_gather(TableView t, SketchState *st, unsigned long long *keys, uint32_t *slots, cudaStream_t s);
void launch_radix_sort(unsigned long long *keys, uint32_t *vals, unsigned long long *tkeys, uint32_t *tvals,
                       uint32_t n, uint32_t *hist, cudaStream_t s);
uint32_t radix_hist_words(uint32_t n);
void launch_bucket_sort(const unsigned long long *keys, const uint32_t *vals, unsigned long long *tkeys, uint32_t *tvals,
                        unsigned long long *okeys, uint32_t *ovals, uint32_t n, uint32_t shift, uint32_t *bins,
                        uint32_t *offs, uint32_t *cursor, SketchState *st, cudaStream_t s);
uint32_t bucket_cap();
void launch_prune_select(TableView t, SketchState *st, uint32_t shift, uint32_t *bins, int scaled,
                         unsigned long long size, unsigned long long max_hash, unsigned long long *keys,
                         uint32_t *slots, cudaStream_t s);
void launch_select_rows(const uint32_t *idx, uint32_t m, int k, uint32_t stride, uint32_t kw, const unsigned long long *i_hash,
                        const uint32_t *i_cnt, const uint32_t *i_ext, const unsigned long long *i_kmer,
                        const unsigned long long *i_posx, unsigned long long *o_hash, uint32_t *o_cnt, uint32_t *o_ext,
                        unsigned long long *o_kmer, unsigned long long *o_posx, uint8_t *o_bytes, cudaStream_t s);
void launch_select_keep(const unsigned long long *keys, uint32_t n, int scaled, unsigned long long size,
                        unsigned long long max_hash, SketchState *st, cudaStream_t s);
void launch_commit_threshold(SketchState *st, cudaStream_t s);
void launch_live_hist_refresh(TableView t, SketchState *st, uint32_t shift, uint32_t *live_bins, cudaStream_t s);
void launch_soft_threshold(const uint32_t *live_bins, uint32_t shift, int scaled, unsigned long long size,
                           unsigned long long max_hash, SketchState *st, cudaStream_t s);
void launch_rebuild(const unsigned long long *keys, const uint32_t *slots, uint32_t keep, TableView from,
                    TableView to, SketchState *st, cudaStream_t s);
void launch_export(const unsigned long long *keys, const uint32_t *slots, uint32_t keep, TableView t,
                   unsigned long long *o_hash, uint32_t *o_cnt, uint32_t *o_ext, unsigned long long *o_kmer,
                   unsigned long long *o_posx, cudaStream_t s);

void launch_filter_pass(const uint32_t *cnt, const uint32_t *ext, uint32_t n, int strand, double cut, int err, uint32_t *hist,
                        uint32_t hcap, uint8_t *ok, uint32_t *meta, cudaStream_t s);
void launch_filter_select(const uint32_t *cnt, const uint8_t *ok, uint32_t n, int use_ok, int abun, uint32_t lo, uint32_t hi,
                          uint32_t limit, uint32_t *idx_out, uint32_t *meta, cudaStream_t s);

void launch_count_kmers(const uint8_t *symbuf, ChunkGeom g, uint32_t n_regions, PiecePlan pp, uint32_t k, uint32_t *counts, cudaStream_t s);
uint32_t allcounts_blocks(uint64_t n);
void launch_allcounts_plan(const uint32_t *counts, uint64_t n, uint32_t k, uint32_t *block_off, unsigned long long *meta, cudaStream_t s);
void launch_allcounts_emit(const uint32_t *counts, uint64_t n, uint32_t k, const uint32_t *block_off, unsigned long long *o_hash,
                           uint32_t *o_cnt, uint32_t *o_ext, uint8_t *o_kmer, cudaStream_t s);

void launch_merge_tables(TableView src, unsigned long long src_thr, unsigned int src_has_max, TableView dst, SketchState *st,
                         cudaStream_t s);
void launch_debug_bump(TableView t, unsigned long long key, unsigned long long add_cnt, unsigned long long add_ext,
                       unsigned int *found, cudaStream_t s);

// dist.cu
void launch_dist_pairs(const unsigned long long *hashes, const uint32_t *lens, uint32_t stride, const uint32_t *q_idx,
                       const uint32_t *r_idx, uint64_t n_pairs, int scaled, unsigned long long max_hash,
                       fb2_pair_out *out, cudaStream_t s);
uint32_t dist_tile_max_len();
int launch_dist_tile(const unsigned long long *hashes, const uint32_t *lens, uint32_t stride, uint32_t n_sk, uint32_t q0,
                     uint32_t q1, int scaled, unsigned long long max_hash, fb2_pair_out *out, cudaStream_t s);
int launch_dist_tile_cut(const unsigned long long *hashes, const uint32_t *lens, uint32_t stride, uint32_t n_sk, uint32_t q0,
                         uint32_t q1, int scaled, unsigned long long max_hash, fb2_pair_hit *hits, unsigned long long *keys,
                         unsigned int *counter, uint32_t cap, int skip_self, double jlow, cudaStream_t s);
void launch_minmer_matrix(const unsigned long long *ref, uint32_t n_ref, const unsigned long long *sk_hash, const uint32_t *sk_cnt,
                          const unsigned long long *sk_off, uint32_t n_sk, uint32_t max_len, int32_t *result, cudaStream_t s);
void launch_iota(uint32_t *v, uint32_t n, cudaStream_t s);
void launch_gather_hits(const fb2_pair_hit *hits, const uint32_t *order, uint32_t n, fb2_pair_hit *sorted, cudaStream_t s);
// the cut through an inverted index over all (hash, sketch) postings (dist.cu)
void launch_postings_fill(const unsigned long long *hashes, const uint32_t *lens, uint32_t stride, uint32_t n_sk,
                          const uint32_t *off, unsigned long long *keys, uint32_t *vals, cudaStream_t s);
void launch_postings_runs(const unsigned long long *keys, const uint32_t *vals, uint32_t n, const uint32_t *off, uint32_t n_sk,
                          uint32_t *sorted_sk, unsigned long long *runinfo, unsigned long long *sum_sq, cudaStream_t s);
uint32_t dist_inverted_block(uint32_t n_sk);
int launch_dist_inverted_cut(const unsigned long long *hashes, const uint32_t *lens, uint32_t stride, uint32_t n_sk, uint32_t q0,
                             uint32_t q1, const uint32_t *off, const uint32_t *sorted_sk, const unsigned long long *runinfo,
                             int scaled, unsigned long long max_hash, fb2_pair_hit *hits, unsigned long long *keys,
                             unsigned int *counter, uint32_t cap, int skip_self, double jlow, cudaStream_t s);
void launch_dist_all(const unsigned long long *hashes, const uint32_t *lens, uint32_t stride, uint32_t n_sk,
                     uint32_t q0, uint64_t n_pairs, int scaled, unsigned long long max_hash, fb2_pair_out *out,
                     cudaStream_t s);

}  // namespace fb2

//! `#[repr(C)]` mirrors and the `extern "C"` block of include/finch_b200.h -- nothing else.
//! NOT COMPILED in this repository's image (no Rust toolchain); the same ABI is exercised by the Python ctypes
//! mirror (finch_rs_b200/__init__.py) and the C++ mirror (finch_rs_b200/host/finch_b200.hpp), which are tested.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const FB2_OK: c_int = 0;
pub const FB2_KIND_MASH: i32 = 0;
pub const FB2_KIND_SCALED: i32 = 1;
pub const FB2_KIND_ALLCOUNTS: i32 = 2;
pub const FB2_FILE_SK: c_int = 0;
pub const FB2_FILE_BSK: c_int = 1;
pub const FB2_FILE_MSH: c_int = 2;
pub const FB2_FORMAT_FASTA: i32 = 1;
pub const FB2_FORMAT_FASTQ: i32 = 2;

#[repr(C)]
pub struct fb2_params {
    pub kind: i32,
    pub kmers_to_sketch: u64,
    pub final_size: u64,
    pub no_strict: i32,
    pub kmer_length: u8,
    pub hash_seed: u64,
    pub scale: f64,
    pub device: i32,
    pub stream: *mut c_void,
}
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct fb2_filter {
    pub filter_on: i32,
    pub has_abun_low: i32,
    pub abun_low: u32,
    pub has_abun_high: i32,
    pub abun_high: u32,
    pub err_filter: f64,
    pub strand_filter: f64,
}
#[repr(C)]
pub struct fb2_result {
    pub n: u64,
    pub hashes: *mut u64,
    pub counts: *mut u32,
    pub extras: *mut u32,
    pub kmers: *mut u8,
    pub kmer_stride: u32,
    pub seq_length: u64,
    pub num_valid_kmers: u64,
    pub format: i32,
    pub filters: fb2_filter,
    pub kmer_lens: *mut u32,
}
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct fb2_pair_out {
    pub common: u32,
    pub i: u32,
    pub j: u32,
}
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct fb2_pair_hit {
    pub q: u32,
    pub r: u32,
    pub common: u32,
    pub i: u32,
    pub j: u32,
}
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct fb2_stats {
    pub kernel_launches: u64,
    pub h2d_bytes: u64,
    pub d2h_bytes: u64,
    pub chunks: u64,
    pub prunes: u64,
    pub hash_launches: u64,
    pub hash_kernel_ms: f64,
    pub parse_kernel_ms: f64,
    pub hash_symbols: u64,
    pub provisional_redos: u64,
    pub band_passes: u64,
}
/// `Sketch` (serialization/mod.rs:45-55) of a sketch file as plain arrays
#[repr(C)]
pub struct fb2_sketch_view {
    pub name: *const c_char,
    pub comment: *const c_char,
    pub seq_length: u64,
    pub num_valid_kmers: u64,
    pub params: fb2_params,
    pub filter: fb2_filter,
    pub n: u64,
    pub hashes: *const u64,
    pub counts: *const u32,
    pub extras: *const u32,
    pub kmers: *const u8,
    pub kmer_offs: *const u64,
}
pub enum fb2_sketcher {}
pub enum fb2_sketch_set {}

extern "C" {
    pub fn fb2_sketcher_create(p: *const fb2_params, out: *mut *mut fb2_sketcher) -> c_int;
    pub fn fb2_sketcher_destroy(s: *mut fb2_sketcher);
    pub fn fb2_sketcher_reset(s: *mut fb2_sketcher) -> c_int;
    pub fn fb2_sketcher_process(s: *mut fb2_sketcher, seq: *const u8, len: usize) -> c_int;
    pub fn fb2_sketcher_push(s: *mut fb2_sketcher, kmer: *const u8, k: usize, extra: u8) -> c_int;
    pub fn fb2_sketcher_feed_fastx(s: *mut fb2_sketcher, bytes: *const u8, len: usize, fin: c_int) -> c_int;
    pub fn fb2_sketcher_format(s: *mut fb2_sketcher, format: *mut i32) -> c_int;
    pub fn fb2_sketcher_totals(s: *mut fb2_sketcher, bases: *mut u64, kmers: *mut u64) -> c_int;
    pub fn fb2_sketcher_result(s: *mut fb2_sketcher, out: *mut fb2_result) -> c_int;
    pub fn fb2_sketcher_sketch(s: *mut fb2_sketcher, name: *const c_char, p: *const fb2_params, f: *const fb2_filter,
                               out: *mut fb2_result) -> c_int;
    pub fn fb2_result_free(r: *mut fb2_result);
    pub fn fb2_sketch_stream(bytes: *const u8, len: usize, name: *const c_char, p: *const fb2_params, f: *const fb2_filter,
                             out: *mut fb2_result) -> c_int;
    pub fn fb2_sketch_stream_multi(bytes: *const u8, len: usize, name: *const c_char, p: *const fb2_params,
                                   f: *const fb2_filter, out: *mut fb2_result, ngpus: c_int) -> c_int;
    pub fn fb2_sketch_files(paths: *const *const c_char, n: usize, p: *const fb2_params, f: *const fb2_filter,
                            outs: *mut fb2_result) -> c_int;
    pub fn fb2_sketch_files_multi(paths: *const *const c_char, n: usize, p: *const fb2_params, f: *const fb2_filter,
                                  outs: *mut fb2_result, ngpus: c_int) -> c_int;
    pub fn fb2_sketch_files_release_pool();
    pub fn fb2_dist_batch(hashes: *const u64, lens: *const u32, n_sk: usize, stride: usize, scale: f64, q_idx: *const u32,
                          r_idx: *const u32, n_pairs: usize, out: *mut fb2_pair_out, device: i32) -> c_int;
    pub fn fb2_dist_all_pairs_cut(hashes: *const u64, lens: *const u32, n_sk: usize, stride: usize, scale: f64, q0: usize,
                                  q1: usize, kmer_length: u8, max_distance: f64, skip_self: c_int, hits: *mut fb2_pair_hit,
                                  cap: usize, n_hits: *mut u64, device: i32, ngpus: c_int) -> c_int;
    pub fn fb2_distance_finish(p: *const fb2_pair_out, kmer_length: u8, containment: *mut f64, jaccard: *mut f64,
                               mash_distance: *mut f64, common_hashes: *mut u64, total_hashes: *mut u64);
    pub fn fb2_sketcher_feed_device(s: *mut fb2_sketcher, dev_bytes: *const u8, len: usize, fin: c_int) -> c_int;
    pub fn fb2_sketcher_stats(s: *mut fb2_sketcher, out: *mut fb2_stats) -> c_int;
    pub fn fb2_last_stream_stats(out: *mut fb2_stats) -> c_int;
    pub fn fb2_sketcher_enable_timing(s: *mut fb2_sketcher, on: c_int) -> c_int;
    pub fn fb2_sketcher_debug_symbols(s: *mut fb2_sketcher, geom7: *mut u32, counts: *mut u32, counts_cap: usize, sym: *mut u8,
                                      sym_cap: usize) -> c_int;
    pub fn fb2_sketcher_debug_bump(s: *mut fb2_sketcher, hash: u64, add_count: u64, add_extra: u64) -> c_int;
    pub fn fb2_filter_counts(r: *mut fb2_result, f: *mut fb2_filter) -> c_int;
    pub fn fb2_process_post_filter(r: *mut fb2_result, p: *const fb2_params, name: *const c_char) -> c_int;
    pub fn fb2_guess_filter_threshold(counts: *const u32, n: usize, filter_level: f64) -> u32;
    pub fn fb2_dist_all_pairs(hashes: *const u64, lens: *const u32, n_sk: usize, stride: usize, scale: f64, q0: usize, q1: usize,
                              out: *mut fb2_pair_out, device: i32) -> c_int;
    pub fn fb2_dist_last_kernel_ms() -> f64;
    pub fn fb2_minmer_matrix(ref_hashes: *const u64, n_ref: usize, sk_hashes: *const u64, sk_counts: *const u32,
                             sk_off: *const u64, n_sk: usize, result: *mut i32, device: i32) -> c_int;
    pub fn fb2_old_distance_finish(common: u64, query_len: u64, ref_len: u64, kmer_length: u8, containment: *mut f64,
                                   jaccard: *mut f64, mash_distance: *mut f64, common_hashes: *mut u64,
                                   total_hashes: *mut u64) -> c_int;
    // sketch files (open_sketch_file lib.rs:96-117, write_finch_file / write_mash_file / MultiSketch JSON); host only
    pub fn fb2_sketch_set_new(out: *mut *mut fb2_sketch_set) -> c_int;
    pub fn fb2_sketch_set_open(path: *const c_char, out: *mut *mut fb2_sketch_set) -> c_int;
    pub fn fb2_sketch_set_len(set: *const fb2_sketch_set) -> u64;
    pub fn fb2_sketch_set_get(set: *const fb2_sketch_set, i: u64, out: *mut fb2_sketch_view) -> c_int;
    pub fn fb2_sketch_set_add(set: *mut fb2_sketch_set, v: *const fb2_sketch_view) -> c_int;
    pub fn fb2_sketch_set_remove(set: *mut fb2_sketch_set, i: u64) -> c_int;
    pub fn fb2_sketch_set_save(set: *const fb2_sketch_set, path: *const c_char, file_format: c_int) -> c_int;
    pub fn fb2_sketch_set_close(set: *mut fb2_sketch_set);
    pub fn fb2_last_error() -> *const c_char;
    pub fn fb2_device_count() -> c_int;
    pub fn fb2_version() -> *const c_char;
}

//! Rust shim over `libfinch_b200.so` (include/finch_b200.h): `impl SketchScheme` for GPU-backed
//! Mash / Scaled sketchers so they drop in behind `SketchParams::create_sketcher`
//! (finch-rs lib/src/sketch_schemes/mod.rs:86-113).  NOT COMPILED in this repository's image
//! (no Rust toolchain); kept as the binding a finch-rs maintainer would add.
use std::cell::RefCell;
use std::ffi::{c_void, CStr};
use std::os::raw::c_char;

use finch::sketch_schemes::{KmerCount, SketchParams, SketchScheme};
use needletail::Sequence;

#[repr(C)]
pub struct Fb2Params {
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
pub struct Fb2Filter {
    pub filter_on: i32,
    pub has_abun_low: i32,
    pub abun_low: u32,
    pub has_abun_high: i32,
    pub abun_high: u32,
    pub err_filter: f64,
    pub strand_filter: f64,
}
#[repr(C)]
pub struct Fb2Result {
    pub n: u64,
    pub hashes: *mut u64,
    pub counts: *mut u32,
    pub extras: *mut u32,
    pub kmers: *mut u8,
    pub kmer_stride: u32,
    pub seq_length: u64,
    pub num_valid_kmers: u64,
    pub format: i32,
    pub filters: Fb2Filter,
}
pub enum Fb2Sketcher {}

#[link(name = "finch_b200")]
extern "C" {
    fn fb2_sketcher_create(p: *const Fb2Params, out: *mut *mut Fb2Sketcher) -> i32;
    fn fb2_sketcher_destroy(s: *mut Fb2Sketcher);
    fn fb2_sketcher_process(s: *mut Fb2Sketcher, seq: *const u8, len: usize) -> i32;
    fn fb2_sketcher_push(s: *mut Fb2Sketcher, kmer: *const u8, k: usize, extra: u8) -> i32;
    fn fb2_sketcher_feed_fastx(s: *mut Fb2Sketcher, bytes: *const u8, len: usize, fin: i32) -> i32;
    fn fb2_sketcher_totals(s: *mut Fb2Sketcher, bases: *mut u64, kmers: *mut u64) -> i32;
    fn fb2_sketcher_result(s: *mut Fb2Sketcher, out: *mut Fb2Result) -> i32;
    fn fb2_result_free(r: *mut Fb2Result);
    fn fb2_last_error() -> *const c_char;
}

fn check(rc: i32) {
    if rc != 0 {
        let msg = unsafe { CStr::from_ptr(fb2_last_error()) }.to_string_lossy().into_owned();
        // SketchScheme::process / to_vec are infallible in the reference; it panics on bad input
        // (lib/src/lib.rs:60,63,72).  Same convention on this side of the ABI only.
        panic!("finch_b200: [{}] {}", rc, msg);
    }
}

/// GPU-resident MashSketcher / ScaledSketcher.  `to_vec` / `total_bases_and_kmers` take `&self`
/// in the trait but flush device work, hence the interior mutability.
pub struct GpuSketcher {
    handle: RefCell<*mut Fb2Sketcher>,
    params: SketchParams,
}

impl GpuSketcher {
    pub fn new(params: &SketchParams) -> Self {
        let p = match params {
            SketchParams::Mash { kmers_to_sketch, final_size, no_strict, kmer_length, hash_seed } => Fb2Params {
                kind: 0, kmers_to_sketch: *kmers_to_sketch as u64, final_size: *final_size as u64,
                no_strict: *no_strict as i32, kmer_length: *kmer_length, hash_seed: *hash_seed, scale: 0.0,
                device: -1, stream: std::ptr::null_mut(),
            },
            SketchParams::Scaled { kmers_to_sketch, kmer_length, scale, hash_seed } => Fb2Params {
                kind: 1, kmers_to_sketch: *kmers_to_sketch as u64, final_size: 0, no_strict: 0,
                kmer_length: *kmer_length, hash_seed: *hash_seed, scale: *scale, device: -1,
                stream: std::ptr::null_mut(),
            },
            SketchParams::AllCounts { .. } => panic!("AllCountsSketcher stays on the CPU path"),
        };
        let mut h: *mut Fb2Sketcher = std::ptr::null_mut();
        check(unsafe { fb2_sketcher_create(&p, &mut h) });
        GpuSketcher { handle: RefCell::new(h), params: params.clone() }
    }
    /// MashSketcher::push / ScaledSketcher::push (mash.rs:34, scaled.rs:37)
    pub fn push(&mut self, kmer: &[u8], extra_count: u8) {
        check(unsafe { fb2_sketcher_push(*self.handle.borrow(), kmer.as_ptr(), kmer.len(), extra_count) });
    }
    /// Bulk replacement of the record loop in sketch_stream (lib/src/lib.rs:60-68).
    pub fn feed_fastx(&mut self, bytes: &[u8], last: bool) {
        check(unsafe { fb2_sketcher_feed_fastx(*self.handle.borrow(), bytes.as_ptr(), bytes.len(), last as i32) });
    }
}

impl SketchScheme for GpuSketcher {
    fn process<'seq, 'a, 'inner>(&'a mut self, seq: &'seq dyn Sequence<'inner>)
    where
        'a: 'seq,
        'seq: 'inner,
    {
        let raw = seq.sequence(); // un-normalised record bytes, as mash.rs:72-73 receives them
        check(unsafe { fb2_sketcher_process(*self.handle.borrow(), raw.as_ptr(), raw.len()) });
    }

    fn total_bases_and_kmers(&self) -> (u64, u64) {
        let (mut b, mut k) = (0u64, 0u64);
        check(unsafe { fb2_sketcher_totals(*self.handle.borrow(), &mut b, &mut k) });
        (b, k)
    }

    fn to_vec(&self) -> Vec<KmerCount> {
        let mut r: Fb2Result = unsafe { std::mem::zeroed() };
        check(unsafe { fb2_sketcher_result(*self.handle.borrow(), &mut r) });
        let n = r.n as usize;
        let st = r.kmer_stride as usize;
        let k = self.params.k() as usize;
        let mut out = Vec::with_capacity(n);
        for i in 0..n {
            unsafe {
                out.push(KmerCount {
                    hash: *r.hashes.add(i),
                    kmer: std::slice::from_raw_parts(r.kmers.add(i * st), k).to_vec(),
                    count: *r.counts.add(i),
                    extra_count: *r.extras.add(i),
                    label: None,
                });
            }
        }
        unsafe { fb2_result_free(&mut r) };
        out
    }

    fn parameters(&self) -> SketchParams {
        self.params.clone()
    }
}

impl Drop for GpuSketcher {
    fn drop(&mut self) {
        unsafe { fb2_sketcher_destroy(*self.handle.borrow()) };
    }
}

//! lib/src/sketch_schemes/b200.rs -- copy this file into finch-rs (it is a module OF finch, not a crate that depends
//! on finch: no dependency cycle) and apply finch-b200.patch next to it.  `impl SketchScheme` for a GPU-resident
//! Mash / Scaled sketcher over libfinch_b200.so (crate finch_b200-sys).  NOT COMPILED in this repository's image.
use std::cell::RefCell;
use std::ffi::CStr;

use finch_b200_sys as sys;
use needletail::Sequence;

use crate::sketch_schemes::{KmerCount, SketchParams, SketchScheme};

fn check(rc: i32) {
    if rc != sys::FB2_OK {
        let msg = unsafe { CStr::from_ptr(sys::fb2_last_error()) }.to_string_lossy().into_owned();
        // SketchScheme::process / to_vec are infallible in the reference, which panics on bad input
        // (lib/src/lib.rs:60,63,72).  Same convention on this side of the ABI only.
        panic!("finch_b200: [{}] {}", rc, msg);
    }
}

/// `to_vec` / `total_bases_and_kmers` take `&self` in the trait but flush device work: interior mutability.
pub struct B200Sketcher {
    handle: RefCell<*mut sys::fb2_sketcher>,
    created_from: SketchParams,
}

impl B200Sketcher {
    pub fn new(params: &SketchParams) -> Self {
        let p = match params {
            SketchParams::Mash { kmers_to_sketch, final_size, no_strict, kmer_length, hash_seed } => sys::fb2_params {
                kind: sys::FB2_KIND_MASH, kmers_to_sketch: *kmers_to_sketch as u64, final_size: *final_size as u64,
                no_strict: *no_strict as i32, kmer_length: *kmer_length, hash_seed: *hash_seed, scale: 0.0,
                device: -1, stream: std::ptr::null_mut(),
            },
            SketchParams::Scaled { kmers_to_sketch, kmer_length, scale, hash_seed } => sys::fb2_params {
                kind: sys::FB2_KIND_SCALED, kmers_to_sketch: *kmers_to_sketch as u64, final_size: 0, no_strict: 0,
                kmer_length: *kmer_length, hash_seed: *hash_seed, scale: *scale, device: -1,
                stream: std::ptr::null_mut(),
            },
            SketchParams::AllCounts { .. } => unreachable!("create_sketcher keeps AllCounts on the CPU path"),
        };
        let mut h: *mut sys::fb2_sketcher = std::ptr::null_mut();
        check(unsafe { sys::fb2_sketcher_create(&p, &mut h) });
        B200Sketcher { handle: RefCell::new(h), created_from: params.clone() }
    }
    /// MashSketcher::push / ScaledSketcher::push (mash.rs:34, scaled.rs:37)
    pub fn push(&mut self, kmer: &[u8], extra_count: u8) {
        check(unsafe { sys::fb2_sketcher_push(*self.handle.borrow(), kmer.as_ptr(), kmer.len(), extra_count) });
    }
    /// Bulk replacement of the record loop in sketch_stream (lib/src/lib.rs:60-68): raw file bytes.
    pub fn feed_fastx(&mut self, bytes: &[u8], last: bool) {
        check(unsafe { sys::fb2_sketcher_feed_fastx(*self.handle.borrow(), bytes.as_ptr(), bytes.len(), last as i32) });
    }
    /// needletail's `seqrec.format()` of the stream fed so far (lib.rs:64-66).
    pub fn format(&self) -> i32 {
        let mut f = 0i32;
        check(unsafe { sys::fb2_sketcher_format(*self.handle.borrow(), &mut f) });
        f
    }
}

impl SketchScheme for B200Sketcher {
    fn process<'seq, 'a, 'inner>(&'a mut self, seq: &'seq dyn Sequence<'inner>)
    where
        'a: 'seq,
        'seq: 'inner,
    {
        let raw = seq.sequence(); // un-normalised record bytes, as mash.rs:72-73 receives them
        check(unsafe { sys::fb2_sketcher_process(*self.handle.borrow(), raw.as_ptr(), raw.len()) });
    }

    fn total_bases_and_kmers(&self) -> (u64, u64) {
        let (mut b, mut k) = (0u64, 0u64);
        check(unsafe { sys::fb2_sketcher_totals(*self.handle.borrow(), &mut b, &mut k) });
        (b, k)
    }

    fn to_vec(&self) -> Vec<KmerCount> {
        let mut r: sys::fb2_result = unsafe { std::mem::zeroed() };
        check(unsafe { sys::fb2_sketcher_result(*self.handle.borrow(), &mut r) });
        let (n, st, k) = (r.n as usize, r.kmer_stride as usize, self.created_from.k() as usize);
        let mut out = Vec::with_capacity(n);
        for i in 0..n {
            unsafe {
                // entries that came through push() keep the caller's bytes, whatever their length (mash.rs:52-55)
                let len = if r.kmer_lens.is_null() { k } else { *r.kmer_lens.add(i) as usize };
                out.push(KmerCount {
                    hash: *r.hashes.add(i),
                    kmer: std::slice::from_raw_parts(r.kmers.add(i * st), len).to_vec(),
                    count: *r.counts.add(i),
                    extra_count: *r.extras.add(i),
                    label: None,
                });
            }
        }
        unsafe { sys::fb2_result_free(&mut r) };
        out
    }

    /// What the reference's sketchers answer, quirks included: MashSketcher reports
    /// `final_size = size, no_strict = false` however it was created (mash.rs:104-112); ScaledSketcher recomputes
    /// `scale` from its integer `max_hash` (scaled.rs:23,31,102-109).
    fn parameters(&self) -> SketchParams {
        match &self.created_from {
            SketchParams::Mash { kmers_to_sketch, kmer_length, hash_seed, .. } => SketchParams::Mash {
                kmers_to_sketch: *kmers_to_sketch,
                final_size: *kmers_to_sketch,
                no_strict: false,
                kmer_length: *kmer_length,
                hash_seed: *hash_seed,
            },
            SketchParams::Scaled { kmers_to_sketch, kmer_length, scale, hash_seed } => {
                let iscale = (1. / *scale) as u64;
                let max_hash = u64::max_value() / iscale;
                SketchParams::Scaled {
                    kmers_to_sketch: *kmers_to_sketch,
                    kmer_length: *kmer_length,
                    scale: 1. / (u64::max_value() as f64 / max_hash as f64),
                    hash_seed: *hash_seed,
                }
            }
            other => other.clone(),
        }
    }
}

impl Drop for B200Sketcher {
    fn drop(&mut self) {
        unsafe { sys::fb2_sketcher_destroy(*self.handle.borrow()) };
    }
}

"""finch_rs_b200 -- host-side mirror (Python) of the finch-rs sketching surface over the C ABI
of ``libfinch_b200.so`` (hand-written sm_100a CUDA; see include/finch_b200.h).

The names and argument meanings follow the reference (paths relative to the reference root):

  ``SketchParams.mash / .scaled`` + ``create_sketcher``   lib/src/sketch_schemes/mod.rs:53-113
  ``MashSketcher`` / ``ScaledSketcher``                    lib/src/sketch_schemes/mash.rs, scaled.rs
     ``.push(kmer, extra_count)``, ``.process(seq)``, ``.total_bases_and_kmers()``, ``.to_vec()``
  ``FilterParams`` (+ ``filter_counts``)                   lib/src/filtering.rs:11-87
  ``sketch_stream`` / ``sketch_files``                     lib/src/lib.rs:29-94
  ``raw_distance`` / ``distance``                          lib/src/distance.rs:9-126

Every compute call goes to the GPU through the C ABI.  There is no CPU fallback: importing works
anywhere, but creating a sketcher without the built library or without a CUDA device raises.
"""
import atexit
import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfinch_b200.so")

OK = 0
EINVAL, ECUDA, EFORMAT, ERECORD, EEMPTY, ETOOFEW, EUNSUPPORTED, EIO, ENOMEM = range(-1, -10, -1)
KIND_MASH, KIND_SCALED, KIND_ALLCOUNTS = 0, 1, 2
FORMAT_UNKNOWN, FORMAT_FASTA, FORMAT_FASTQ = 0, 1, 2


class FinchError(RuntimeError):
    """FinchError::Message (lib/src/errors.rs:5-23) carrying the C-ABI status code."""

    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code
        self.message = msg


class _Params(C.Structure):
    _fields_ = [("kind", C.c_int32), ("kmers_to_sketch", C.c_uint64), ("final_size", C.c_uint64),
                ("no_strict", C.c_int32), ("kmer_length", C.c_uint8), ("hash_seed", C.c_uint64),
                ("scale", C.c_double), ("device", C.c_int32), ("stream", C.c_void_p)]


class _Filter(C.Structure):
    _fields_ = [("filter_on", C.c_int32), ("has_abun_low", C.c_int32), ("abun_low", C.c_uint32),
                ("has_abun_high", C.c_int32), ("abun_high", C.c_uint32),
                ("err_filter", C.c_double), ("strand_filter", C.c_double)]


class _Result(C.Structure):
    _fields_ = [("n", C.c_uint64), ("hashes", C.POINTER(C.c_uint64)), ("counts", C.POINTER(C.c_uint32)),
                ("extras", C.POINTER(C.c_uint32)), ("kmers", C.POINTER(C.c_uint8)),
                ("kmer_stride", C.c_uint32), ("seq_length", C.c_uint64),
                ("num_valid_kmers", C.c_uint64), ("format", C.c_int32), ("filters", _Filter),
                ("kmer_lens", C.POINTER(C.c_uint32))]


class _Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("chunks", C.c_uint64), ("prunes", C.c_uint64), ("hash_launches", C.c_uint64),
                ("hash_kernel_ms", C.c_double), ("parse_kernel_ms", C.c_double),
                ("hash_symbols", C.c_uint64), ("provisional_redos", C.c_uint64), ("band_passes", C.c_uint64)]


class _SketchView(C.Structure):
    """fb2_sketch_view: one `Sketch` (serialization/mod.rs:45-55) of a sketch file as plain arrays"""
    _fields_ = [("name", C.c_char_p), ("comment", C.c_char_p), ("seq_length", C.c_uint64), ("num_valid_kmers", C.c_uint64),
                ("params", _Params), ("filter", _Filter), ("n", C.c_uint64), ("hashes", C.POINTER(C.c_uint64)),
                ("counts", C.POINTER(C.c_uint32)), ("extras", C.POINTER(C.c_uint32)), ("kmers", C.POINTER(C.c_uint8)),
                ("kmer_offs", C.POINTER(C.c_uint64))]


class _PairOut(C.Structure):
    _fields_ = [("common", C.c_uint32), ("i", C.c_uint32), ("j", C.c_uint32)]


EXPORTS = [
    "fb2_sketcher_create", "fb2_sketcher_destroy", "fb2_sketcher_reset", "fb2_sketcher_process",
    "fb2_sketcher_push", "fb2_sketcher_feed_fastx", "fb2_sketcher_feed_device", "fb2_sketcher_format",
    "fb2_sketcher_totals", "fb2_sketcher_result", "fb2_sketcher_sketch", "fb2_result_free", "fb2_sketcher_stats", "fb2_last_stream_stats",
    "fb2_sketcher_enable_timing", "fb2_sketcher_debug_symbols", "fb2_sketcher_debug_bump", "fb2_filter_counts", "fb2_process_post_filter",
    "fb2_guess_filter_threshold", "fb2_sketch_stream", "fb2_sketch_files", "fb2_sketch_files_multi", "fb2_sketch_stream_multi",
    "fb2_sketch_files_release_pool", "fb2_dist_batch", "fb2_minmer_matrix",
    "fb2_dist_all_pairs", "fb2_dist_all_pairs_cut", "fb2_dist_last_kernel_ms", "fb2_distance_finish", "fb2_old_distance_finish", "fb2_last_error", "fb2_device_count", "fb2_version",
    "fb2_sketch_set_new", "fb2_sketch_set_open", "fb2_sketch_set_len", "fb2_sketch_set_get", "fb2_sketch_set_add", "fb2_sketch_set_remove",
    "fb2_sketch_set_save", "fb2_sketch_set_close",
]

_lib = None


def lib():
    """Load libfinch_b200.so (built by __graft_entry__.build() / make -C finch_rs_b200/csrc)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FinchError(ECUDA, f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                                f"g.build()'` -- finch_rs_b200 has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, sz = C.c_void_p, C.c_size_t
    L.fb2_sketcher_create.argtypes = [C.POINTER(_Params), C.POINTER(vp)]
    L.fb2_sketcher_destroy.argtypes = [vp]
    L.fb2_sketcher_destroy.restype = None
    L.fb2_sketcher_reset.argtypes = [vp]
    L.fb2_sketcher_process.argtypes = [vp, vp, sz]
    L.fb2_sketcher_push.argtypes = [vp, C.c_char_p, sz, C.c_uint8]
    L.fb2_sketcher_feed_fastx.argtypes = [vp, vp, sz, C.c_int]
    L.fb2_sketcher_feed_device.argtypes = [vp, vp, sz, C.c_int]
    L.fb2_sketcher_format.argtypes = [vp, C.POINTER(C.c_int32)]
    L.fb2_sketcher_totals.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.fb2_sketcher_result.argtypes = [vp, C.POINTER(_Result)]
    L.fb2_sketcher_sketch.argtypes = [vp, C.c_char_p, C.POINTER(_Params), C.POINTER(_Filter), C.POINTER(_Result)]
    L.fb2_result_free.argtypes = [C.POINTER(_Result)]
    L.fb2_result_free.restype = None
    L.fb2_sketcher_stats.argtypes = [vp, C.POINTER(_Stats)]
    L.fb2_last_stream_stats.argtypes = [C.POINTER(_Stats)]
    L.fb2_minmer_matrix.argtypes = [vp, sz, vp, vp, vp, sz, vp, C.c_int32]
    L.fb2_sketcher_enable_timing.argtypes = [vp, C.c_int]
    L.fb2_sketcher_debug_symbols.argtypes = [vp, vp, vp, sz, vp, sz]
    L.fb2_sketcher_debug_bump.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint64]
    L.fb2_filter_counts.argtypes = [C.POINTER(_Result), C.POINTER(_Filter)]
    L.fb2_process_post_filter.argtypes = [C.POINTER(_Result), C.POINTER(_Params), C.c_char_p]
    L.fb2_guess_filter_threshold.argtypes = [vp, sz, C.c_double]
    L.fb2_guess_filter_threshold.restype = C.c_uint32
    L.fb2_sketch_stream.argtypes = [vp, sz, C.c_char_p, C.POINTER(_Params), C.POINTER(_Filter), C.POINTER(_Result)]
    L.fb2_sketch_files.argtypes = [C.POINTER(C.c_char_p), sz, C.POINTER(_Params), C.POINTER(_Filter), C.POINTER(_Result)]
    L.fb2_sketch_files_multi.argtypes = [C.POINTER(C.c_char_p), sz, C.POINTER(_Params), C.POINTER(_Filter), C.POINTER(_Result), C.c_int]
    L.fb2_sketch_stream_multi.argtypes = [vp, sz, C.c_char_p, C.POINTER(_Params), C.POINTER(_Filter), C.POINTER(_Result), C.c_int]
    L.fb2_dist_batch.argtypes = [vp, vp, sz, sz, C.c_double, vp, vp, sz, vp, C.c_int32]
    L.fb2_dist_all_pairs.argtypes = [vp, vp, sz, sz, C.c_double, sz, sz, vp, C.c_int32]
    L.fb2_dist_all_pairs_cut.argtypes = [vp, vp, sz, sz, C.c_double, sz, sz, C.c_uint8, C.c_double, C.c_int, vp, sz,
                                         C.POINTER(C.c_uint64), C.c_int32, C.c_int]
    L.fb2_distance_finish.argtypes = [C.POINTER(_PairOut), C.c_uint8, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                      C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.fb2_distance_finish.restype = None
    L.fb2_old_distance_finish.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint8, C.POINTER(C.c_double),
                                          C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.fb2_dist_last_kernel_ms.argtypes = []
    L.fb2_dist_last_kernel_ms.restype = C.c_double
    L.fb2_last_error.restype = C.c_char_p
    L.fb2_version.restype = C.c_char_p
    L.fb2_sketch_set_new.argtypes = [C.POINTER(vp)]
    L.fb2_sketch_set_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.fb2_sketch_set_len.argtypes = [vp]
    L.fb2_sketch_set_len.restype = C.c_uint64
    L.fb2_sketch_set_get.argtypes = [vp, C.c_uint64, C.POINTER(_SketchView)]
    L.fb2_sketch_set_add.argtypes = [vp, C.POINTER(_SketchView)]
    L.fb2_sketch_set_remove.argtypes = [vp, C.c_uint64]
    L.fb2_sketch_set_save.argtypes = [vp, C.c_char_p, C.c_int]
    L.fb2_sketch_set_close.argtypes = [vp]
    L.fb2_sketch_set_close.restype = None
    _lib = L
    atexit.register(L.fb2_sketch_files_release_pool)   # idle worker handles of sketch_files (bounded, see finch_b200.h)
    return L


def _check(rc):
    if rc != OK:
        raise FinchError(rc, lib().fb2_last_error().decode("utf-8", "replace"))


def _addr(b):
    """bytes-like / numpy uint8 -> (address, nbytes, keepalive)"""
    if isinstance(b, np.ndarray):
        a = np.ascontiguousarray(b).view(np.uint8).reshape(-1)
        return a.ctypes.data, a.size, a
    if isinstance(b, (bytes, bytearray, memoryview)):
        a = np.frombuffer(b, dtype=np.uint8)
        return (a.ctypes.data if a.size else None), a.size, a
    raise TypeError(f"unsupported buffer type {type(b)}")


# ---------------------------------------------------------------------------------------------
@dataclass
class KmerCount:
    """lib/src/sketch_schemes/mod.rs:15-22"""
    hash: int
    kmer: bytes
    count: int
    extra_count: int
    label: Optional[bytes] = None


@dataclass
class FilterParams:
    """lib/src/filtering.rs:11-16.  err_filter is the internal value (CLI percent * k / 100)."""
    filter_on: Optional[bool] = False
    abun_filter: Tuple[Optional[int], Optional[int]] = (None, None)
    err_filter: float = 0.0
    strand_filter: float = 0.0

    def _c(self):
        lo, hi = self.abun_filter
        return _Filter(-1 if self.filter_on is None else int(bool(self.filter_on)),
                       lo is not None, lo or 0, hi is not None, hi or 0, self.err_filter, self.strand_filter)

    @staticmethod
    def _from_c(f):
        return FilterParams(None if f.filter_on < 0 else bool(f.filter_on),
                            (f.abun_low if f.has_abun_low else None, f.abun_high if f.has_abun_high else None),
                            f.err_filter, f.strand_filter)


@dataclass
class SketchParams:
    """enum SketchParams (mod.rs:53-71) restricted to the GPU path's variants."""
    kind: int = KIND_MASH
    kmers_to_sketch: int = 1000
    final_size: int = 1000
    no_strict: bool = False
    kmer_length: int = 21
    hash_seed: int = 0
    scale: float = 0.0
    device: int = -1

    @staticmethod
    def mash(kmers_to_sketch=1000, final_size=1000, no_strict=False, kmer_length=21, hash_seed=0, device=-1):
        return SketchParams(KIND_MASH, kmers_to_sketch, final_size, no_strict, kmer_length, hash_seed, 0.0, device)

    @staticmethod
    def scaled(kmers_to_sketch=1000, kmer_length=21, scale=0.001, hash_seed=0, device=-1):
        return SketchParams(KIND_SCALED, kmers_to_sketch, 0, False, kmer_length, hash_seed, scale, device)

    @staticmethod
    def allcounts(kmer_length=4, device=-1):
        """SketchParams::AllCounts (mod.rs:68-70): counts of all 4^k k-mers (sketch_schemes/counts.rs), k <= 16."""
        return SketchParams(KIND_ALLCOUNTS, 0, 0, False, kmer_length, 0, 0.0, device)

    @staticmethod
    def from_cli(sketch_type="mash", n_hashes=1000, kmer_length=21, seed=0, oversketch=200, scale=0.001,
                 no_strict=False, filters_enabled: Optional[bool] = None, device=-1):
        """parse_sketch_options (cli/src/cli.rs:277-340): Mash over-sketches n*200 unless --no-filter."""
        if sketch_type == "mash":
            size = n_hashes * oversketch if filters_enabled in (True, None) else n_hashes
            return SketchParams.mash(size, n_hashes, no_strict, kmer_length, seed, device)
        if sketch_type == "scaled":
            return SketchParams.scaled(n_hashes, kmer_length, scale, seed, device)
        if sketch_type == "none":                          # cli.rs:336
            return SketchParams.allcounts(kmer_length, device)
        raise FinchError(EINVAL, "A unknown sketch type was selected")

    def k(self):
        return self.kmer_length

    def expected_size(self):  # mod.rs:148-156
        if self.kind == KIND_ALLCOUNTS:
            return 4 ** self.kmer_length
        return self.final_size if self.kind == KIND_MASH else self.kmers_to_sketch

    def _c(self, stream=None):
        return _Params(self.kind, self.kmers_to_sketch, self.final_size, int(self.no_strict), self.kmer_length,
                       self.hash_seed, self.scale, self.device, stream)

    def create_sketcher(self, stream=None):
        """mod.rs:86-113"""
        return _Sketcher(self, stream)


def _result_to_py(r, k):
    n, st = int(r.n), int(r.kmer_stride)
    if n:
        h = np.ctypeslib.as_array(r.hashes, (n,)).copy()
        c = np.ctypeslib.as_array(r.counts, (n,)).copy()
        x = np.ctypeslib.as_array(r.extras, (n,)).copy()
        km = np.ctypeslib.as_array(r.kmers, (n * st,)).copy().reshape(n, st)
    else:
        h, c, x = np.zeros(0, np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.uint32)
        km = np.zeros((0, max(st, 1)), np.uint8)
    return h, c, x, km


@dataclass
class Sketch:
    """lib/src/serialization/mod.rs:45-55 (SoA arrays + a KmerCount view)."""
    name: str
    seq_length: int
    num_valid_kmers: int
    comment: str
    hashes_u64: np.ndarray
    counts: np.ndarray
    extra_counts: np.ndarray
    kmers: np.ndarray            # [n, stride] uint8
    filter_params: FilterParams
    sketch_params: SketchParams
    format: int = FORMAT_UNKNOWN

    def __len__(self):
        return len(self.hashes_u64)

    def kmer_bytes(self, i, k=None):
        return self.kmers[i, :(k or self.sketch_params.kmer_length)].tobytes()

    @property
    def hashes(self) -> List[KmerCount]:
        k = self.sketch_params.kmer_length
        return [KmerCount(int(self.hashes_u64[i]), self.kmers[i, :k].tobytes(), int(self.counts[i]),
                          int(self.extra_counts[i])) for i in range(len(self))]


class _Sketcher:
    """A GPU-resident MashSketcher / ScaledSketcher (trait SketchScheme, mod.rs:24-51)."""

    def __init__(self, params: SketchParams, stream=None):
        self.params = params
        self._h = C.c_void_p()
        cp = params._c(stream)
        _check(lib().fb2_sketcher_create(C.byref(cp), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().fb2_sketcher_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:      # interpreter shutdown: module globals may already be gone
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def reset(self):
        _check(lib().fb2_sketcher_reset(self._h))

    # -- SketchScheme ---------------------------------------------------------------------
    def process(self, seq):
        """process(&mut self, seq: &dyn Sequence): one record's raw sequence bytes."""
        addr, n, keep = _addr(seq)
        _check(lib().fb2_sketcher_process(self._h, addr, n))

    def push(self, kmer: bytes, extra_count: int):
        """MashSketcher::push / ScaledSketcher::push (mash.rs:34, scaled.rs:37)."""
        _check(lib().fb2_sketcher_push(self._h, bytes(kmer), len(kmer), extra_count))

    def total_bases_and_kmers(self):
        a, b = C.c_uint64(), C.c_uint64()
        _check(lib().fb2_sketcher_totals(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def parameters(self):
        """SketchScheme::parameters as the reference's sketchers report them (quirks Q7 / Q8): MashSketcher answers
        final_size = size, no_strict = false whatever it was created from (mash.rs:104-112); ScaledSketcher
        recomputes the scale from its integer max_hash (scaled.rs:102-109)."""
        p = self.params
        if p.kind == KIND_ALLCOUNTS:                      # counts.rs:65-69
            return SketchParams.allcounts(p.kmer_length, p.device)
        if p.kind == KIND_MASH:
            return SketchParams.mash(p.kmers_to_sketch, p.kmers_to_sketch, False, p.kmer_length, p.hash_seed, p.device)
        iscale = int(1.0 / p.scale)                       # scaled.rs:23: (1. / scale) as u64
        max_hash = (2**64 - 1) // iscale                  # scaled.rs:31
        return SketchParams.scaled(p.kmers_to_sketch, p.kmer_length, 1.0 / (float(2**64 - 1) / float(max_hash)), p.hash_seed, p.device)

    def to_arrays(self):
        """(hashes u64[n], counts u32[n], extras u32[n], kmers u8[n, stride], seq_length, n_kmers, format)"""
        r = _Result()
        _check(lib().fb2_sketcher_result(self._h, C.byref(r)))
        try:
            h, c, x, km = _result_to_py(r, self.params.kmer_length)
            return h, c, x, km, int(r.seq_length), int(r.num_valid_kmers), int(r.format)
        finally:
            lib().fb2_result_free(C.byref(r))

    def to_vec(self, kmer_len=None) -> List[KmerCount]:
        h, c, x, km, *_ = self.to_arrays()
        k = kmer_len if kmer_len is not None else self.params.kmer_length
        return [KmerCount(int(h[i]), km[i, :k].tobytes(), int(c[i]), int(x[i])) for i in range(len(h))]

    def to_sketch(self) -> Sketch:
        """mod.rs:33-50: name "", default filters."""
        h, c, x, km, sl, nk, fmt = self.to_arrays()
        return Sketch("", sl, nk, "", h, c, x, km, FilterParams(), self.parameters(), fmt)

    def sketch(self, name: str, filters: "FilterParams") -> "Sketch":
        """The tail of sketch_stream (lib.rs:78-93): to_vec + filter_counts + process_post_filter."""
        r = _Result()
        cp, cf = self.params._c(), filters._c()
        _check(lib().fb2_sketcher_sketch(self._h, name.encode(), C.byref(cp), C.byref(cf), C.byref(r)))
        return _finish(r, name, self.params)

    # -- bulk feeds (replace the record loop lib.rs:60-68) ---------------------------------------
    def feed_fastx(self, data, final=True):
        addr, n, keep = _addr(data)
        _check(lib().fb2_sketcher_feed_fastx(self._h, addr, n, int(final)))

    def feed_fastx_ptr(self, host_ptr: int, nbytes: int, final=True):
        _check(lib().fb2_sketcher_feed_fastx(self._h, host_ptr, nbytes, int(final)))

    def feed_device(self, dev_ptr: int, nbytes: int, final=True):
        _check(lib().fb2_sketcher_feed_device(self._h, dev_ptr, nbytes, int(final)))

    def format(self):
        f = C.c_int32()
        _check(lib().fb2_sketcher_format(self._h, C.byref(f)))
        return f.value

    def debug_symbols(self):
        """(geom dict, counts[n_st], list of per-region symbol arrays, raw buffer) of the last chunk."""
        g = np.zeros(7, np.uint32)
        _check(lib().fb2_sketcher_debug_symbols(self._h, g.ctypes.data, None, 0, None, 0))
        names = ("len", "n_tiles", "st_tiles", "n_st", "st_bytes", "region_stride", "hash_tiles")
        geom = dict(zip(names, (int(x) for x in g)))
        counts = np.zeros(geom["n_st"], np.uint32)
        front = geom["region_stride"] - geom["st_bytes"]      # SYM_FRONT: pad in front of every region
        buf = np.zeros(front + geom["n_st"] * geom["region_stride"] + 64, np.uint8)
        _check(lib().fb2_sketcher_debug_symbols(self._h, g.ctypes.data, counts.ctypes.data, counts.size,
                                                buf.ctypes.data, buf.size))
        regs = [buf[front + r * geom["region_stride"]: front + r * geom["region_stride"] + int(counts[r])] for r in range(geom["n_st"])]
        return geom, counts, regs, buf

    def debug_bump(self, hash_, add_count, add_extra):
        _check(lib().fb2_sketcher_debug_bump(self._h, hash_, add_count, add_extra))

    def enable_timing(self, on=True):
        _check(lib().fb2_sketcher_enable_timing(self._h, int(on)))

    def stats(self):
        s = _Stats()
        _check(lib().fb2_sketcher_stats(self._h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in _Stats._fields_}


def MashSketcher(size, kmer_length, seed, device=-1):
    """MashSketcher::new(size, kmer_length, seed)  (mash.rs:21)"""
    return SketchParams.mash(size, size, False, kmer_length, seed, device).create_sketcher()


def AllCountsSketcher(kmer_length, device=-1):
    """AllCountsSketcher::new(k)  (counts.rs:14-21): process / total_bases_and_kmers / to_vec / parameters, no push."""
    return SketchParams.allcounts(kmer_length, device).create_sketcher()


def ScaledSketcher(size, scale, kmer_length, seed, device=-1):
    """ScaledSketcher::new(size, scale, kmer_length, seed)  (scaled.rs:22)"""
    return SketchParams.scaled(size, kmer_length, scale, seed, device).create_sketcher()


# ---------------------------------------------------------------------------------------------
class _ResultOwner:
    """Keeps a library-owned fb2_result alive while numpy views of its arrays are in use."""

    def __init__(self, r):
        self.r = r

    def __del__(self):
        try:
            lib().fb2_result_free(C.byref(self.r))
        except Exception:
            pass


def _finish(r, name, sp):
    n, st = int(r.n), int(r.kmer_stride)
    if n < 65536:   # small: plain copies
        try:
            h, c, x, km = _result_to_py(r, sp.kmer_length)
            return Sketch(name, int(r.seq_length), int(r.num_valid_kmers), "", h, c, x, km,
                          FilterParams._from_c(r.filters), sp, int(r.format))
        finally:
            lib().fb2_result_free(C.byref(r))
    # large (e.g. Scaled sketches of whole genomes): zero-copy views, freed with the Sketch
    owner = _ResultOwner(r)
    h = np.ctypeslib.as_array(r.hashes, (n,))
    c = np.ctypeslib.as_array(r.counts, (n,))
    x = np.ctypeslib.as_array(r.extras, (n,))
    km = np.ctypeslib.as_array(r.kmers, (n * st,)).reshape(n, st)
    sk = Sketch(name, int(r.seq_length), int(r.num_valid_kmers), "", h, c, x, km,
                FilterParams._from_c(r.filters), sp, int(r.format))
    sk._owner = owner
    return sk


def sketch_stream(data, name: str, sketch_params: SketchParams, filters: FilterParams) -> Sketch:
    """lib/src/lib.rs:51-94 over an in-memory byte stream."""
    addr, n, keep = _addr(data)
    r = _Result()
    cp, cf = sketch_params._c(), filters._c()
    _check(lib().fb2_sketch_stream(addr, n, name.encode(), C.byref(cp), C.byref(cf), C.byref(r)))
    return _finish(r, name, sketch_params)


def sketch_stream_ptr(host_ptr: int, nbytes: int, name: str, sketch_params: SketchParams, filters: FilterParams) -> Sketch:
    """sketch_stream over host memory given by address (e.g. a pinned buffer)."""
    r = _Result()
    cp, cf = sketch_params._c(), filters._c()
    _check(lib().fb2_sketch_stream(host_ptr, nbytes, name.encode(), C.byref(cp), C.byref(cf), C.byref(r)))
    return _finish(r, name, sketch_params)


def last_stream_stats() -> dict:
    """Counters of the calling thread's most recent sketch_stream call (kernel launches, bytes over PCIe, ...)."""
    st = _Stats()
    _check(lib().fb2_last_stream_stats(C.byref(st)))
    return {k: getattr(st, k) for k, _ in _Stats._fields_}


def sketch_stream_multi(data, name: str, sketch_params: SketchParams, filters: FilterParams, ngpus=0) -> Sketch:
    """sketch_stream of one file cut into byte ranges over `ngpus` GPUs (0 = all), united exactly on the first."""
    if isinstance(data, int):
        raise TypeError("pass a buffer, or use sketch_stream_multi_ptr(ptr, nbytes, ...)")
    addr, n, keep = _addr(data)
    return sketch_stream_multi_ptr(addr, n, name, sketch_params, filters, ngpus)


def sketch_stream_multi_ptr(host_ptr, nbytes, name, sketch_params, filters, ngpus=0) -> Sketch:
    r = _Result()
    cp, cf = sketch_params._c(), filters._c()
    _check(lib().fb2_sketch_stream_multi(host_ptr, nbytes, name.encode(), C.byref(cp), C.byref(cf), C.byref(r), ngpus))
    return _finish(r, name, sketch_params)


def sketch_files(filenames, sketch_params: SketchParams, filters: FilterParams, ngpus=1) -> List[Sketch]:
    """lib/src/lib.rs:29-49; results in input order.  ngpus: shard the files over that many GPUs (0 = all)."""
    n = len(filenames)
    arr = (C.c_char_p * n)(*[f.encode() for f in filenames])
    outs = (_Result * n)()
    cp, cf = sketch_params._c(), filters._c()
    _check(lib().fb2_sketch_files_multi(arr, n, C.byref(cp), C.byref(cf), outs, ngpus))
    return [_finish(outs[i], filenames[i], sketch_params) for i in range(n)]


def filter_counts(filters: FilterParams, hashes, counts, extras, fmt=FORMAT_FASTQ):
    """FilterParams::filter_counts on plain arrays -> (kept hashes, counts, extras, updated FilterParams)."""
    h = np.ascontiguousarray(hashes, np.uint64).copy()
    c = np.ascontiguousarray(counts, np.uint32).copy()
    x = np.ascontiguousarray(extras, np.uint32).copy()
    km = np.zeros(max(1, len(h)), np.uint8)
    r = _Result(len(h), h.ctypes.data_as(C.POINTER(C.c_uint64)), c.ctypes.data_as(C.POINTER(C.c_uint32)),
                x.ctypes.data_as(C.POINTER(C.c_uint32)), km.ctypes.data_as(C.POINTER(C.c_uint8)), 1, 0, 0, fmt,
                _Filter(), None)
    cf = filters._c()
    _check(lib().fb2_filter_counts(C.byref(r), C.byref(cf)))
    n = int(r.n)
    return h[:n], c[:n], x[:n], FilterParams._from_c(cf)


def guess_filter_threshold(counts, level):
    c = np.ascontiguousarray(counts, np.uint32)
    return int(lib().fb2_guess_filter_threshold(c.ctypes.data, c.size, level))


# ---- distance ---------------------------------------------------------------------------------
@dataclass
class SketchDistance:
    """lib/src/serialization/mod.rs:31-43"""
    containment: float
    jaccard: float
    mash_distance: float
    common_hashes: int
    total_hashes: int
    query: str = ""
    reference: str = ""


def _pack(sketch_hashes):
    lens = np.array([len(s) for s in sketch_hashes], np.uint32)
    stride = max(1, int(lens.max()) if len(lens) else 1)
    mat = np.zeros((len(sketch_hashes), stride), np.uint64)
    for i, s in enumerate(sketch_hashes):
        mat[i, :len(s)] = np.asarray(s, np.uint64)
    return mat, lens, stride


def dist_batch(sketch_hashes, q_idx, r_idx, scale=0.0, device=-1):
    """Integer part of raw_distance for many pairs: -> array[(common, i, j)] (uint32 x 3)."""
    mat, lens, stride = _pack(sketch_hashes)
    q, r = np.ascontiguousarray(q_idx, np.uint32), np.ascontiguousarray(r_idx, np.uint32)
    out = np.zeros((len(q), 3), np.uint32)
    _check(lib().fb2_dist_batch(mat.ctypes.data, lens.ctypes.data, len(lens), stride, scale, q.ctypes.data,
                                r.ctypes.data, len(q), out.ctypes.data, device))
    return out


def dist_all_pairs(mat, lens, scale=0.0, q0=0, q1=None, device=-1, out=None):
    """All ordered pairs (q, r), q in [q0, q1): -> uint32 array [q1-q0, n, 3] of (common, i, j).
    `out`: optional preallocated C-contiguous uint32 array of that size (reused across calls)."""
    mat = np.ascontiguousarray(mat, np.uint64)
    lens = np.ascontiguousarray(lens, np.uint32)
    n, stride = mat.shape
    q1 = n if q1 is None else q1
    if out is None:
        out = np.zeros(((q1 - q0) * n, 3), np.uint32)
    elif out.dtype != np.uint32 or out.size != (q1 - q0) * n * 3 or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous uint32 array with (q1 - q0) * n * 3 elements")
    _check(lib().fb2_dist_all_pairs(mat.ctypes.data, lens.ctypes.data, n, stride, scale, q0, q1,
                                    out.ctypes.data, device))
    return out.reshape(q1 - q0, n, 3)


HIT_DTYPE = np.dtype([("q", np.uint32), ("r", np.uint32), ("common", np.uint32), ("i", np.uint32), ("j", np.uint32)])


def dist_all_pairs_cut(mat, lens, kmer_length, max_distance, scale=0.0, q0=0, q1=None, skip_self=True, device=-1, ngpus=1,
                       cap=1 << 20):
    """calc_sketch_distances with the max_distance cut (cli/src/main.rs:315-334): all ordered pairs (q in [q0, q1),
    r) whose mash distance can be <= max_distance, ascending by (q, r), as a structured array (q, r, common, i, j).
    The device cut is conservative; finish with distance_of_hits() and apply `mash_distance <= max_distance`."""
    mat = np.ascontiguousarray(mat, np.uint64)
    lens = np.ascontiguousarray(lens, np.uint32)
    n, stride = mat.shape
    q1 = n if q1 is None else q1
    while True:
        hits = np.zeros(max(1, cap), HIT_DTYPE)
        nh = C.c_uint64()
        rc = lib().fb2_dist_all_pairs_cut(mat.ctypes.data, lens.ctypes.data, n, stride, scale, q0, q1, kmer_length, max_distance,
                                          int(bool(skip_self)), hits.ctypes.data, hits.size, C.byref(nh), device, ngpus)
        if rc == ENOMEM and nh.value > hits.size:
            cap = int(nh.value)
            continue
        _check(rc)
        return hits[:nh.value]


def distance_of_hits(hits, kmer_length):
    """(containment, jaccard, mash_distance, common, total) arrays for a hit list, exactly as fb2_distance_finish
    (distance.rs:117-125, :35-41) computes them one pair at a time."""
    n = len(hits)
    cont, jac, md = np.zeros(n), np.zeros(n), np.zeros(n)
    com, tot = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
    for t in range(n):
        cont[t], jac[t], md[t], com[t], tot[t] = _finish_pair((hits["common"][t], hits["i"][t], hits["j"][t]), kmer_length)
    return cont, jac, md, com, tot


def _finish_pair(row, k):
    p = _PairOut(int(row[0]), int(row[1]), int(row[2]))
    cont, jac, md = C.c_double(), C.c_double(), C.c_double()
    com, tot = C.c_uint64(), C.c_uint64()
    lib().fb2_distance_finish(C.byref(p), k, C.byref(cont), C.byref(jac), C.byref(md), C.byref(com), C.byref(tot))
    return cont.value, jac.value, md.value, com.value, tot.value


def minmer_matrix(ref_sketch, sketches, device=-1):
    """distance.rs:344-364 (the reference's `numpy` feature): an int32 array [len(sketches), len(ref_sketch)] holding,
    for every sketch, its count of each of the reference sketch's hashes (0 where it does not hold the hash).
    `ref_sketch`: a Sketch or an ascending u64 array; `sketches`: Sketch objects or (hashes, counts) pairs."""
    ref = np.ascontiguousarray(getattr(ref_sketch, "hashes_u64", ref_sketch), np.uint64)
    hs, cs = [], []
    for sk in sketches:
        h, c = (sk.hashes_u64, sk.counts) if hasattr(sk, "hashes_u64") else sk
        hs.append(np.ascontiguousarray(h, np.uint64)); cs.append(np.ascontiguousarray(c, np.uint32))
        if len(hs[-1]) != len(cs[-1]):
            raise FinchError(EINVAL, "hashes and counts differ in length")
    off = np.zeros(len(hs) + 1, np.uint64)
    if hs:
        off[1:] = np.cumsum([len(h) for h in hs])
    allh = np.concatenate(hs) if hs else np.zeros(0, np.uint64)
    allc = np.concatenate(cs) if cs else np.zeros(0, np.uint32)
    out = np.zeros((len(hs), len(ref)), np.int32)
    _check(lib().fb2_minmer_matrix(ref.ctypes.data, len(ref), allh.ctypes.data, allc.ctypes.data, off.ctypes.data, len(hs),
                                   out.ctypes.data, device))
    return out


def raw_distance(query_hashes, ref_hashes, scale=0.0):
    """distance.rs:66-126 -> (containment, jaccard, common, total)"""
    out = dist_batch([query_hashes, ref_hashes], [0], [1], scale)
    cont, jac, _, com, tot = _finish_pair(out[0], 1)
    return cont, jac, com, tot


def distance(query: Sketch, ref: Sketch, old_mode=False) -> SketchDistance:
    """distance.rs:9-47, both modes."""
    if old_mode:   # old_distance (distance.rs:136-157): |Q n R| over all of R
        out = dist_batch([query.hashes_u64, ref.hashes_u64], [0], [1], 0.0)
        cont, jac, md = C.c_double(), C.c_double(), C.c_double()
        com, tot = C.c_uint64(), C.c_uint64()
        _check(lib().fb2_old_distance_finish(int(out[0][0]), len(query.hashes_u64), len(ref.hashes_u64),
                                             query.sketch_params.kmer_length, C.byref(cont), C.byref(jac), C.byref(md),
                                             C.byref(com), C.byref(tot)))
        return SketchDistance(cont.value, jac.value, md.value, com.value, tot.value, query.name, ref.name)
    s1 = query.sketch_params.scale if query.sketch_params.kind == KIND_SCALED else None
    s2 = ref.sketch_params.scale if ref.sketch_params.kind == KIND_SCALED else None
    min_scale = min(s1, s2) if (s1 is not None and s2 is not None) else 0.0
    out = dist_batch([query.hashes_u64, ref.hashes_u64], [0], [1], min_scale)
    cont, jac, md, com, tot = _finish_pair(out[0], query.sketch_params.kmer_length)
    return SketchDistance(cont, jac, md, com, tot, query.name, ref.name)

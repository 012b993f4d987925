"""ctypes binding of the CPU oracle (oracle/libfinch_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (finch_rs_b200) never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfinch_oracle.so")

FMT_FASTA, FMT_FASTQ = 1, 2
OK, E_EMPTY, E_FORMAT, E_RECORD, E_TOO_FEW = 0, -1, -2, -3, -4


def build(force=False):
    src = os.path.join(_HERE, "finch_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libfinch_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


class FilterParams(C.Structure):
    _fields_ = [("filter_on", C.c_int), ("has_abun_low", C.c_int), ("abun_low", C.c_uint32),
                ("has_abun_high", C.c_int), ("abun_high", C.c_uint32),
                ("err_filter", C.c_double), ("strand_filter", C.c_double)]


class SketchParams(C.Structure):
    _fields_ = [("kind", C.c_int), ("kmers_to_sketch", C.c_uint64), ("final_size", C.c_uint64),
                ("no_strict", C.c_int), ("kmer_length", C.c_uint8), ("hash_seed", C.c_uint64),
                ("scale", C.c_double)]


class Sketch(C.Structure):
    _fields_ = [("seq_length", C.c_uint64), ("num_valid_kmers", C.c_uint64), ("n", C.c_size_t),
                ("hashes", C.POINTER(C.c_uint64)), ("counts", C.POINTER(C.c_uint32)),
                ("extras", C.POINTER(C.c_uint32)), ("kmers", C.POINTER(C.c_uint8)),
                ("filters", FilterParams), ("format", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    u8p, u32p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    L.fo_murmur3_x64_128.argtypes = [C.c_char_p, C.c_size_t, C.c_uint64, u64p]
    L.fo_hash_f.argtypes = [C.c_char_p, C.c_size_t, C.c_uint64]
    L.fo_hash_f.restype = C.c_uint64
    L.fo_normalize.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
    L.fo_normalize.restype = C.c_size_t
    L.fo_reverse_complement.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
    L.fo_kmer_stream.argtypes = [C.c_void_p, C.c_size_t, C.c_uint8, C.c_uint64, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_size_t]
    L.fo_kmer_stream.restype = C.c_size_t
    L.fo_mash_new.argtypes = [C.c_size_t, C.c_uint8, C.c_uint64]
    L.fo_mash_new.restype = C.c_void_p
    L.fo_scaled_new.argtypes = [C.c_size_t, C.c_double, C.c_uint8, C.c_uint64]
    L.fo_scaled_new.restype = C.c_void_p
    L.fo_sketcher_free.argtypes = [C.c_void_p]
    L.fo_push.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_uint8]
    L.fo_process.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.fo_totals.argtypes = [C.c_void_p, u64p, u64p]
    L.fo_scaled_max_hash.argtypes = [C.c_void_p]
    L.fo_scaled_max_hash.restype = C.c_uint64
    L.fo_result_len.argtypes = [C.c_void_p]
    L.fo_result_len.restype = C.c_size_t
    L.fo_result.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    L.fo_result.restype = C.c_size_t
    L.fo_hist.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    L.fo_hist.restype = C.c_uint64
    L.fo_guess_filter_threshold.argtypes = [C.c_void_p, C.c_size_t, C.c_double]
    L.fo_guess_filter_threshold.restype = C.c_uint32
    L.fo_filter_strands.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_void_p]
    L.fo_filter_strands.restype = C.c_size_t
    L.fo_filter_abundance.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_int,
                                      C.c_uint32, C.c_void_p]
    L.fo_filter_abundance.restype = C.c_size_t
    L.fo_filter_counts.argtypes = [C.POINTER(FilterParams), C.c_void_p, C.c_void_p, C.c_size_t,
                                   C.c_void_p]
    L.fo_filter_counts.restype = C.c_size_t
    L.fo_sketch_stream.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(SketchParams),
                                   C.POINTER(FilterParams), C.POINTER(Sketch)]
    L.fo_sketch_stream.restype = C.c_int
    L.fo_sketch_free.argtypes = [C.POINTER(Sketch)]
    L.fo_raw_distance.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_double,
                                  C.POINTER(C.c_double), C.POINTER(C.c_double), u64p, u64p]
    L.fo_old_distance.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                  C.POINTER(C.c_double), C.POINTER(C.c_double), u64p, u64p]
    L.fo_allcounts_process.argtypes = [C.c_void_p, C.c_uint8, C.c_void_p, C.c_size_t]
    L.fo_allcounts_to_vec.argtypes = [C.c_void_p, C.c_uint8, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    L.fo_allcounts_to_vec.restype = C.c_size_t
    L.fo_allcounts_total.argtypes = [C.c_void_p, C.c_uint8]
    L.fo_allcounts_total.restype = C.c_uint64
    L.fo_minmer_matrix.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.fo_mash_distance.argtypes = [C.c_double, C.c_uint8]
    L.fo_mash_distance.restype = C.c_double
    _lib = L
    return L


def _buf(b):
    """bytes / bytearray / numpy uint8 -> (address, length, keepalive)"""
    if isinstance(b, np.ndarray):
        a = np.ascontiguousarray(b, dtype=np.uint8)
        return a.ctypes.data, a.size, a
    a = np.frombuffer(bytes(b), dtype=np.uint8) if len(b) else np.zeros(0, np.uint8)
    return a.ctypes.data, a.size, a


def murmur3_x64_128(data, seed=0):
    out = (C.c_uint64 * 2)()
    lib().fo_murmur3_x64_128(bytes(data), len(data), seed, out)
    return int(out[0]), int(out[1])


def hash_f(item, seed=0):
    return int(lib().fo_hash_f(bytes(item), len(item), seed))


def normalize(seq):
    out = C.create_string_buffer(max(1, len(seq)))
    n = lib().fo_normalize(bytes(seq), len(seq), out)
    return out.raw[:n]


def reverse_complement(seq):
    out = C.create_string_buffer(max(1, len(seq)))
    lib().fo_reverse_complement(bytes(seq), len(seq), out)
    return out.raw[:len(seq)]


def kmer_stream(seq, k, seed=0, want_kmers=False):
    """(hashes u64[n], is_rc u8[n][, kmers u8[n,k]]) for one raw record sequence."""
    addr, n, keep = _buf(seq)
    cnt = lib().fo_kmer_stream(addr, n, k, seed, None, None, None, 0)
    h = np.zeros(cnt, np.uint64)
    rc = np.zeros(cnt, np.uint8)
    km = np.zeros((cnt, k), np.uint8) if want_kmers else None
    lib().fo_kmer_stream(addr, n, k, seed, h.ctypes.data, rc.ctypes.data,
                         km.ctypes.data if want_kmers else None, cnt)
    return (h, rc, km) if want_kmers else (h, rc)


class Sketcher:
    """Mirror of MashSketcher / ScaledSketcher (mash.rs, scaled.rs)."""

    def __init__(self, handle, k):
        self._h, self.k = handle, k

    @classmethod
    def mash(cls, size, kmer_length, seed):
        return cls(lib().fo_mash_new(size, kmer_length, seed), kmer_length)

    @classmethod
    def scaled(cls, size, scale, kmer_length, seed):
        return cls(lib().fo_scaled_new(size, scale, kmer_length, seed), kmer_length)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().fo_sketcher_free(self._h)
            self._h = None

    def push(self, kmer, extra_count):
        lib().fo_push(self._h, bytes(kmer), len(kmer), extra_count)

    def process(self, raw_seq):
        addr, n, keep = _buf(raw_seq)
        lib().fo_process(self._h, addr, n)

    def total_bases_and_kmers(self):
        a, b = C.c_uint64(), C.c_uint64()
        lib().fo_totals(self._h, C.byref(a), C.byref(b))
        return a.value, b.value

    def max_hash(self):
        return int(lib().fo_scaled_max_hash(self._h))

    def to_vec(self, kmer_len=None):
        """dict(hashes, counts, extras, kmers[list of bytes]) ascending by hash."""
        n = lib().fo_result_len(self._h)
        kl = kmer_len if kmer_len is not None else self.k
        h, c, x = np.zeros(n, np.uint64), np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        km = np.zeros((n, max(1, kl)), np.uint8)
        lib().fo_result(self._h, h.ctypes.data, c.ctypes.data, x.ctypes.data, km.ctypes.data,
                        max(1, kl))
        return {"hashes": h, "counts": c, "extras": x,
                "kmers": [km[i, :kl].tobytes() for i in range(n)]}


def parse_fastx(data):
    """-> (rc, format, [raw record sequences])"""
    recs = []
    CB = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_size_t)

    def cb(ctx, p, n):
        recs.append(C.string_at(p, n) if n else b"")

    L = lib()
    L.fo_parse_fastx.argtypes = [C.c_void_p, C.c_size_t, CB, C.c_void_p, C.POINTER(C.c_int),
                                 C.POINTER(C.c_uint64)]
    L.fo_parse_fastx.restype = C.c_int
    addr, n, keep = _buf(data)
    fmt, nrec = C.c_int(), C.c_uint64()
    rc = L.fo_parse_fastx(addr, n, CB(cb), None, C.byref(fmt), C.byref(nrec))
    return rc, fmt.value, recs


def hist(counts):
    c = np.ascontiguousarray(counts, np.uint32)
    mx = lib().fo_hist(c.ctypes.data, c.size, None, 0)
    out = np.zeros(mx, np.uint64)
    lib().fo_hist(c.ctypes.data, c.size, out.ctypes.data, mx)
    return out


def guess_filter_threshold(counts, level):
    c = np.ascontiguousarray(counts, np.uint32)
    return int(lib().fo_guess_filter_threshold(c.ctypes.data, c.size, level))


def filter_strands(counts, extras, cutoff):
    c, x = np.ascontiguousarray(counts, np.uint32), np.ascontiguousarray(extras, np.uint32)
    keep = np.zeros(c.size, np.uint32)
    m = lib().fo_filter_strands(c.ctypes.data, x.ctypes.data, c.size, cutoff, keep.ctypes.data)
    return keep[:m]


def filter_abundance(counts, low=None, high=None):
    c = np.ascontiguousarray(counts, np.uint32)
    keep = np.zeros(c.size, np.uint32)
    m = lib().fo_filter_abundance(c.ctypes.data, c.size, low is not None, low or 0,
                                  high is not None, high or 0, keep.ctypes.data)
    return keep[:m]


def make_filter(filter_on=None, abun=(None, None), err_filter=0.0, strand_filter=0.0):
    return FilterParams(-1 if filter_on is None else int(bool(filter_on)),
                        abun[0] is not None, abun[0] or 0, abun[1] is not None, abun[1] or 0,
                        err_filter, strand_filter)


def filter_counts(fp, counts, extras):
    c, x = np.ascontiguousarray(counts, np.uint32), np.ascontiguousarray(extras, np.uint32)
    keep = np.zeros(c.size, np.uint32)
    m = lib().fo_filter_counts(C.byref(fp), c.ctypes.data, x.ctypes.data, c.size, keep.ctypes.data)
    return keep[:m]


def mash_params(kmers_to_sketch=1000, final_size=1000, no_strict=False, kmer_length=21, hash_seed=0):
    return SketchParams(0, kmers_to_sketch, final_size, int(no_strict), kmer_length, hash_seed, 0.0)


def scaled_params(kmers_to_sketch=1000, kmer_length=21, scale=0.001, hash_seed=0):
    return SketchParams(1, kmers_to_sketch, 0, 0, kmer_length, hash_seed, scale)


def sketch_stream(data, sp, fp, kmers_array=False):
    """lib.rs:51-94.  -> (rc, dict or None).  kmers_array: k-mers as one [n, k] uint8 array (large sketches)."""
    addr, n, keep = _buf(data)
    sk = Sketch()
    rc = lib().fo_sketch_stream(addr, n, C.byref(sp), C.byref(fp), C.byref(sk))
    if rc != OK:
        return rc, None
    k = sp.kmer_length
    m = sk.n
    out = {
        "seq_length": int(sk.seq_length), "num_valid_kmers": int(sk.num_valid_kmers),
        "hashes": np.ctypeslib.as_array(sk.hashes, (m,)).copy() if m else np.zeros(0, np.uint64),
        "counts": np.ctypeslib.as_array(sk.counts, (m,)).copy() if m else np.zeros(0, np.uint32),
        "extras": np.ctypeslib.as_array(sk.extras, (m,)).copy() if m else np.zeros(0, np.uint32),
        "kmers": ((np.ctypeslib.as_array(sk.kmers, (m * k,)).copy().reshape(m, k) if m else np.zeros((0, k), np.uint8))
                  if kmers_array else
                  [bytes(np.ctypeslib.as_array(sk.kmers, (m * k,))[i * k:(i + 1) * k]) for i in range(m)] if m else []),
        "format": sk.format,
        "filter_on": bool(sk.filters.filter_on == 1),
        "min_copies": int(sk.filters.abun_low) if sk.filters.has_abun_low else None,
        "max_copies": int(sk.filters.abun_high) if sk.filters.has_abun_high else None,
    }
    lib().fo_sketch_free(C.byref(sk))
    return rc, out


def raw_distance(q, r, scale=0.0):
    q, r = np.ascontiguousarray(q, np.uint64), np.ascontiguousarray(r, np.uint64)
    cont, jac, com, tot = C.c_double(), C.c_double(), C.c_uint64(), C.c_uint64()
    lib().fo_raw_distance(q.ctypes.data, q.size, r.ctypes.data, r.size, scale, C.byref(cont),
                          C.byref(jac), C.byref(com), C.byref(tot))
    return cont.value, jac.value, com.value, tot.value


def old_distance(q, r):
    """distance.rs:136-157; None where the reference panics (empty query with a non-empty reference)."""
    q, r = np.ascontiguousarray(q, np.uint64), np.ascontiguousarray(r, np.uint64)
    cont, jac, com, tot = C.c_double(), C.c_double(), C.c_uint64(), C.c_uint64()
    if lib().fo_old_distance(q.ctypes.data, q.size, r.ctypes.data, r.size, C.byref(cont), C.byref(jac),
                             C.byref(com), C.byref(tot)) != 0:
        return None
    return cont.value, jac.value, com.value, tot.value


def mash_distance(jaccard, k):
    return float(lib().fo_mash_distance(jaccard, k))


class AllCountsSketcher:
    """Mirror of AllCountsSketcher (lib/src/sketch_schemes/counts.rs)."""

    def __init__(self, k):
        self.k = k
        self.counts = np.zeros(4 ** k, np.uint32)

    def process(self, raw_seq):
        addr, n, keep = _buf(raw_seq)
        lib().fo_allcounts_process(self.counts.ctypes.data, self.k, addr, n)

    def total_bases_and_kmers(self):
        return 0, int(lib().fo_allcounts_total(self.counts.ctypes.data, self.k))

    def to_vec(self):
        cap = int(np.count_nonzero(self.counts))
        h, c, x = np.zeros(cap, np.uint64), np.zeros(cap, np.uint32), np.zeros(cap, np.uint32)
        km = np.zeros((cap, self.k), np.uint8)
        n = lib().fo_allcounts_to_vec(self.counts.ctypes.data, self.k, h.ctypes.data, c.ctypes.data, x.ctypes.data,
                                      km.ctypes.data, cap)
        return {"hashes": h[:n], "counts": c[:n], "extras": x[:n], "kmers": [km[i].tobytes() for i in range(n)]}


def minmer_matrix(ref_hashes, sketches):
    """distance.rs:344-364.  sketches: list of (hashes u64 ascending, counts u32)."""
    ref = np.ascontiguousarray(ref_hashes, np.uint64)
    hs = [np.ascontiguousarray(h, np.uint64) for h, _ in sketches]
    cs = [np.ascontiguousarray(c, np.uint32) for _, c in sketches]
    n = len(sketches)
    hp = (C.c_void_p * max(1, n))(*[a.ctypes.data for a in hs])
    cp = (C.c_void_p * max(1, n))(*[a.ctypes.data for a in cs])
    ln = (C.c_size_t * max(1, n))(*[len(a) for a in hs])
    out = np.zeros((n, len(ref)), np.int32)
    lib().fo_minmer_matrix(ref.ctypes.data, len(ref), hp, cp, ln, n, out.ctypes.data)
    return out


# ---- the reference's Python module (lib/src/python.rs): literal restatements of its host loops -------------------
# Entries are (hash, kmer, count, extra_count) tuples, ascending by hash, like `Sketch.hashes` returns them.
def py_merge_sketches(hashes1, hashes2, size=None, scale=None):
    """merge_sketches (python.rs:44-101): the two-pointer walk (it stops when EITHER list ends, the rest of the longer
    list is dropped), then the clip by `size` / `scale`.  u32 additions wrap (release build)."""
    new = []
    i = j = 0
    while i < len(hashes1) and j < len(hashes2):
        if hashes1[i][0] < hashes2[j][0]:
            new.append(hashes1[i]); i += 1
        elif hashes2[j][0] < hashes1[i][0]:
            new.append(hashes2[j]); j += 1
        else:
            new.append((hashes1[i][0], hashes1[i][1], (hashes1[i][2] + hashes2[j][2]) & 0xFFFFFFFF,
                        (hashes1[i][3] + hashes2[j][3]) & 0xFFFFFFFF))
            i += 1; j += 1
    if scale is not None:
        max_hash = ((1 << 64) - 1) // int(1.0 / scale)
        out = []
        for ix, h in enumerate(new):
            if not (h[0] <= max_hash or (size is not None and ix < size)):
                break
            out.append(h)
        new = out
    elif size is not None:
        new = new[:size]
    return new


def py_compare_counts(reference, query):
    """Sketch.compare_counts (python.rs:496-561), statement by statement."""
    common = ref_pos = ref_count = query_pos = query_count = 0
    mean = m2 = m3 = m4 = 0.0
    while ref_pos < len(reference) and query_pos < len(query):
        if reference[ref_pos][0] < query[query_pos][0]:
            ref_pos += 1
        elif query[query_pos][0] < reference[ref_pos][0]:
            query_pos += 1
        else:
            ref_count += reference[ref_pos][2]
            query_count += query[query_pos][2]
            n = float(common) + 1.0
            float_count = float(query[query_pos][2])
            delta = float_count - mean
            delta_n = delta / n
            delta_n2 = delta_n * delta_n
            term1 = delta * delta_n * (n - 1.0)
            mean += delta_n
            m4 += term1 * delta_n2 * (n * n - 3.0 * n + 3.0) + 6.0 * delta_n2 * m2 - 4.0 * delta_n * m3
            m3 += term1 * delta_n * (n - 2.0) - 3.0 * delta_n * m2
            m2 += term1
            ref_pos += 1; query_pos += 1; common += 1
    f = np.float64
    with np.errstate(all="ignore"):
        var = f(m2) / f(common)
        skew = np.sqrt(f(common)) * f(m3) / np.power(f(m2), f(1.5))
        kurt = f(common) * f(m4) / (f(m2) * f(m2)) - f(3.0)
    return (common, ref_pos, query_pos, ref_count, query_count, float(var), float(skew), float(kurt))


def py_set_counts(hashes, values):
    """Sketch.set_counts (python.rs:585-608) -> the new entries, or an error message"""
    if len(values) != len(hashes):
        return "counts must be same length as sketch"
    new = []
    for s, v in zip(hashes, values):
        if v < 0:
            return f"Negative count {v} not supported"
        if v > 0:
            new.append((s[0], s[1], int(v), s[3]))
    return new

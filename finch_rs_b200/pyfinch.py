"""pyfinch -- the reference's Python module (``import finch``, built from lib/src/python.rs with pyo3) over the
B200 engine: ``sketch_file``, ``Sketch``, ``Multisketch``, ``FinchError`` with the reference's names, argument
meanings, defaults and error behaviour.  ``python/finch/__init__.py`` re-exports this module under the reference's
name, so ``from finch import sketch_file, Multisketch`` keeps working with ``<repo>/python`` on the path.

Where the work happens:

  ``sketch_file``                    -> ``fb2_sketch_files`` (GPU: parse, canonical k-mers, murmur3, bottom-s, filters)
  ``Sketch.compare``, ``Multisketch.best_match`` / ``filter_to_matches``
                                     -> ``fb2_dist_batch`` (GPU: sorted-hash intersection) + ``fb2_distance_finish``
  ``Sketch.compare_matrix``          -> ``fb2_minmer_matrix`` (GPU)
  ``Multisketch.open`` / ``save``    -> ``fb2_sketch_set_open`` / ``_save`` (host: .sk / .bsk / .msh codecs)
  ``merge``, ``compare_counts``, ``counts``: O(sketch) host arithmetic, as in the reference (python.rs:24-103, 496-568)

Quirks of the reference kept on purpose (parity first; each is covered by a test):
  * ``Sketch.merge`` stops at the end of the SHORTER hash list: the tail of the longer one is dropped
    (python.rs:49-72), and it adds ``seq_length`` / ``num_valid_kmers`` before it checks compatibility (:26-27).
  * A negative index into a ``Multisketch`` is turned into ``len - key`` (python.rs:284-286), which is past the end:
    the reference panics there (pyo3 raises ``PanicException``, a ``BaseException``); so does this module.
  * ``Multisketch.best_match`` only replaces its candidate on a strictly larger containment, so all-zero
    containments return sketch 0; on an empty collection the reference panics (index out of bounds).
  * ``Sketch.compare(other)`` treats ``other`` as the QUERY and ``self`` as the reference (python.rs:484).
  * ``sketch_file`` passes ``err_filter = 1.0`` as the internal fraction (python.rs:668-673), not the command line's
    percent-times-k value.
"""
import ctypes as C
import math
from typing import Iterable, List, Optional, Tuple

import numpy as np

import finch_rs_b200 as _fb

__all__ = ["sketch_file", "Sketch", "Multisketch", "SketchIter", "FinchError", "PanicException"]

_U64_MAX = (1 << 64) - 1
FILE_SK, FILE_BSK, FILE_MSH = 0, 1, 2


class FinchError(Exception):
    """``create_exception!(finch, FinchError, PyException)`` (python.rs:17)"""


class PanicException(BaseException):
    """What pyo3 raises when the Rust side panics (``pyo3_runtime.PanicException`` derives from BaseException)."""


def _try(fn, *a, **kw):
    """py_try! (python.rs:18-22): library errors surface as finch.FinchError carrying the message alone."""
    try:
        return fn(*a, **kw)
    except _fb.FinchError as e:
        raise FinchError(e.message) from None


def _check(rc):
    if rc != _fb.OK:
        raise FinchError(_fb.lib().fb2_last_error().decode("utf-8", "replace"))


# ---- the Rust-side Sketch (serialization/mod.rs:45-55) ---------------------------------------------------------
class _SketchRs:
    __slots__ = ("name", "seq_length", "num_valid_kmers", "comment", "hashes", "kmers", "counts", "extra_counts",
                 "sketch_params", "filter_params")

    def __init__(self, name, seq_length, num_valid_kmers, comment, hashes, kmers, counts, extra_counts, sketch_params,
                 filter_params):
        self.name, self.seq_length, self.num_valid_kmers, self.comment = name, int(seq_length), int(num_valid_kmers), comment
        self.hashes = np.ascontiguousarray(hashes, np.uint64)
        self.kmers = list(kmers)                       # bytes per entry (files may hold none: b"")
        self.counts = np.ascontiguousarray(counts, np.uint32)
        self.extra_counts = np.ascontiguousarray(extra_counts, np.uint32)
        self.sketch_params, self.filter_params = sketch_params, filter_params

    def clone(self):
        return _SketchRs(self.name, self.seq_length, self.num_valid_kmers, self.comment, self.hashes.copy(), list(self.kmers),
                         self.counts.copy(), self.extra_counts.copy(), self.sketch_params, self.filter_params)

    @staticmethod
    def from_engine(sk: "_fb.Sketch"):
        k = sk.sketch_params.kmer_length
        kb = np.ascontiguousarray(sk.kmers[:, :k])
        return _SketchRs(sk.name, sk.seq_length, sk.num_valid_kmers, sk.comment, np.array(sk.hashes_u64, np.uint64),
                         [kb[i].tobytes() for i in range(len(sk))], np.array(sk.counts, np.uint32),
                         np.array(sk.extra_counts, np.uint32), sk.sketch_params, sk.filter_params)

    def scale(self) -> Optional[float]:
        """hash_info().3 (mod.rs:138-146)"""
        return self.sketch_params.scale if self.sketch_params.kind == _fb.KIND_SCALED else None


def _params_from_c(c) -> "_fb.SketchParams":
    if c.kind == _fb.KIND_MASH:
        return _fb.SketchParams.mash(int(c.kmers_to_sketch), int(c.final_size), bool(c.no_strict), int(c.kmer_length), int(c.hash_seed))
    if c.kind == _fb.KIND_SCALED:
        return _fb.SketchParams.scaled(int(c.kmers_to_sketch), int(c.kmer_length), float(c.scale), int(c.hash_seed))
    return _fb.SketchParams.allcounts(int(c.kmer_length))


def _hash_type(p) -> str:
    return "None" if p.kind == _fb.KIND_ALLCOUNTS else "MurmurHash3_x64_128"


def _check_compatibility(a, b) -> Optional[Tuple[str, str, str]]:
    """SketchParams::check_compatibility (mod.rs:186-214)"""
    if a.kmer_length != b.kmer_length:
        return "k", str(a.kmer_length), str(b.kmer_length)
    if _hash_type(a) != _hash_type(b):
        return "hash type", _hash_type(a), _hash_type(b)
    bits_a, bits_b = (0 if a.kind == _fb.KIND_ALLCOUNTS else 64), (0 if b.kind == _fb.KIND_ALLCOUNTS else 64)
    if bits_a != bits_b:
        return "hash bits", str(bits_a), str(bits_b)
    seed_a, seed_b = (0 if a.kind == _fb.KIND_ALLCOUNTS else a.hash_seed), (0 if b.kind == _fb.KIND_ALLCOUNTS else b.hash_seed)
    if seed_a != seed_b:
        return "hash seed", str(seed_a), str(seed_b)
    return None


def _max_hash_for_scale(sc: float) -> int:
    """``u64::max_value() / (1. / sc) as u64`` (python.rs:78,90): the cast saturates, a zero divisor panics."""
    inv = 1.0 / sc if sc != 0.0 else math.inf
    if inv != inv or inv <= 0.0:
        iscale = 0
    elif inv >= 18446744073709551616.0:
        iscale = _U64_MAX
    else:
        iscale = int(inv)
    if iscale == 0:
        raise PanicException("attempt to divide by zero")
    return _U64_MAX // iscale


def _merge_sketches(sketch: _SketchRs, other: _SketchRs, size: Optional[int]):
    """merge_sketches (python.rs:24-103), vectorised: the two-pointer walk stops when EITHER list ends, i.e. it emits
    the sorted union of the entries that are <= min(last hash of each list) -- the tail of the longer list is lost."""
    sketch.seq_length += other.seq_length
    sketch.num_valid_kmers += other.num_valid_kmers
    bad = _check_compatibility(sketch.sketch_params, other.sketch_params)
    if bad:
        raise FinchError(f"First sketch has {bad[0]} {bad[1]}, but second sketch has {bad[0]} {bad[2]}")
    h1, h2 = sketch.hashes, other.hashes
    if len(h1) == 0 or len(h2) == 0:
        nh, nk = np.zeros(0, np.uint64), []
        nc, nx = np.zeros(0, np.uint32), np.zeros(0, np.uint32)
    else:
        # the walk ends right after consuming the last element of the list whose maximum is smaller (ties: both)
        t = min(int(h1[-1]), int(h2[-1]))
        n1 = int(np.searchsorted(h1, np.uint64(t), side="right"))
        n2 = int(np.searchsorted(h2, np.uint64(t), side="right"))
        a, b = h1[:n1], h2[:n2]
        nh = np.union1d(a, b)
        ia = np.searchsorted(nh, a)
        ib = np.searchsorted(nh, b)
        cnt = np.zeros(len(nh), np.uint64)
        ext = np.zeros(len(nh), np.uint64)
        cnt[ia] += sketch.counts[:n1]; cnt[ib] += other.counts[:n2]
        ext[ia] += sketch.extra_counts[:n1]; ext[ib] += other.extra_counts[:n2]
        if (cnt > 0xFFFFFFFF).any() or (ext > 0xFFFFFFFF).any():
            # `count + count` on u32 (python.rs:62-63): a debug build panics; the wheels are release builds and wrap
            cnt &= 0xFFFFFFFF; ext &= 0xFFFFFFFF
        nc, nx = cnt.astype(np.uint32), ext.astype(np.uint32)
        # the k-mer (and label) of an entry present in both comes from the first sketch (python.rs:60-66)
        src_kmers: List[bytes] = [b""] * len(nh)
        for j, q in enumerate(ib):
            src_kmers[q] = other.kmers[j] if j < len(other.kmers) else b""
        for j, q in enumerate(ia):
            src_kmers[q] = sketch.kmers[j] if j < len(sketch.kmers) else b""
        nk = src_kmers
    sc = sketch.scale()
    keep = len(nh)
    if sc is not None:
        max_hash = _max_hash_for_scale(sc)
        below = int(np.searchsorted(nh, np.uint64(max_hash), side="right"))     # take_while(hash <= max_hash ...)
        keep = max(below, min(size, len(nh))) if size is not None else below    # ... || ix < size
    elif size is not None:
        keep = min(size, len(nh))
    sketch.hashes, sketch.kmers = nh[:keep].copy(), nk[:keep]
    sketch.counts, sketch.extra_counts = nc[:keep].copy(), nx[:keep].copy()


def _distances(query: _SketchRs, refs: List[_SketchRs], old_mode=False):
    """distance(query, ref, old_mode) (distance.rs:9-47) for every ref: -> list of (containment, jaccard).  One GPU call
    per distinct scale (usually one)."""
    if not refs:
        return []
    out = [None] * len(refs)
    k = query.sketch_params.kmer_length
    groups = {}
    for i, r in enumerate(refs):
        s1, s2 = query.scale(), r.scale()
        ms = 0.0 if (old_mode or s1 is None or s2 is None) else min(s1, s2)
        groups.setdefault(ms, []).append(i)
    for ms, idx in groups.items():
        lists = [query.hashes] + [refs[i].hashes for i in idx]
        rows = _try(_fb.dist_batch, lists, [0] * len(idx), list(range(1, len(idx) + 1)), ms)
        for row, i in zip(rows, idx):
            if old_mode:
                cont, jac, md = C.c_double(), C.c_double(), C.c_double()
                com, tot = C.c_uint64(), C.c_uint64()
                rc = _fb.lib().fb2_old_distance_finish(int(row[0]), len(query.hashes), len(refs[i].hashes), k, C.byref(cont),
                                                       C.byref(jac), C.byref(md), C.byref(com), C.byref(tot))
                if rc != _fb.OK:   # the reference indexes an empty query: panic (distance.rs:141)
                    raise PanicException("index out of bounds: the len is 0 but the index is 0")
                out[i] = (cont.value, jac.value)
            else:
                cont, jac, _, _, _ = _fb._finish_pair(row, k)
                out[i] = (cont, jac)
    return out


# ---- #[pyclass] Sketch (python.rs:310-616) -------------------------------------------------------------------
class Sketch:
    """A Sketch is a collection of deterministically-selected hashes from a single sequencing file."""

    def __init__(self, name: str):
        if not isinstance(name, str):
            raise TypeError("argument 'name': 'str' expected")
        # python.rs:318-338
        self.s = _SketchRs(name, 0, 0, "", np.zeros(0, np.uint64), [], np.zeros(0, np.uint32), np.zeros(0, np.uint32),
                           _fb.SketchParams.mash(1000, 1000, True, 21, 0), _fb.FilterParams(False, (None, None), 0.0, 0.0))

    @classmethod
    def _wrap(cls, s: _SketchRs) -> "Sketch":
        o = cls.__new__(cls)
        o.s = s
        return o

    def __repr__(self):
        return f'<Sketch "{self.s.name}">'

    def __len__(self):
        return len(self.s.hashes)

    @property
    def name(self) -> str:
        return self.s.name

    @name.setter
    def name(self, value: str):
        if not isinstance(value, str):
            raise TypeError("'str' expected")
        self.s.name = value

    @property
    def seq_length(self) -> int:
        return self.s.seq_length

    @property
    def num_valid_kmers(self) -> int:
        return self.s.num_valid_kmers

    @property
    def comment(self) -> str:
        return self.s.comment

    @comment.setter
    def comment(self, value: str):
        if not isinstance(value, str):
            raise TypeError("'str' expected")
        self.s.comment = value

    @property
    def hashes(self) -> List[Tuple[int, bytes, int, int]]:
        """(hash, kmer, count, extra_count) per entry, ascending by hash (python.rs:384-402)"""
        s = self.s
        return [(int(s.hashes[i]), s.kmers[i] if i < len(s.kmers) else b"", int(s.counts[i]), int(s.extra_counts[i]))
                for i in range(len(s.hashes))]

    @property
    def sketch_params(self) -> dict:
        """python.rs:423-463"""
        p = self.s.sketch_params
        if p.kind == _fb.KIND_MASH:
            return {"sketch_type": "mash", "kmers_to_sketch": p.kmers_to_sketch, "final_size": p.final_size,
                    "no_strict": bool(p.no_strict), "kmer_length": p.kmer_length, "hash_seed": p.hash_seed}
        if p.kind == _fb.KIND_SCALED:
            return {"sketch_type": "scaled", "kmers_to_sketch": p.kmers_to_sketch, "kmer_length": p.kmer_length,
                    "scale": p.scale, "hash_seed": p.hash_seed}
        return {"sketch_type": "none", "kmer_length": p.kmer_length}

    def merge(self, sketch: "Sketch", size: Optional[int] = None):
        """merge(self, sketch: Sketch, size: int) -- python.rs:472-474"""
        _require_sketch(sketch)
        if size is not None and (not isinstance(size, int) or size < 0):
            raise OverflowError("can't convert negative int to unsigned")
        _merge_sketches(self.s, sketch.s, size)

    def compare(self, sketch: "Sketch", old_mode: bool = False) -> Tuple[float, float]:
        """compare(self, sketch, old_mode=False) -> (containment, jaccard): `sketch` is the query, `self` the reference
        (python.rs:482-487)"""
        _require_sketch(sketch)
        return _distances(sketch.s, [self.s], bool(old_mode))[0]

    def compare_counts(self, sketch: "Sketch"):
        """compare_counts(self, sketch) -> (common, ref_pos, query_pos, ref_count, query_count, var, skew, kurt)
        (python.rs:496-561; `self` is the reference).  The running moments are order-dependent f64 sums, so they are
        accumulated in the reference's order, one common hash at a time."""
        _require_sketch(sketch)
        ref, qry = self.s, sketch.s
        rh, qh = ref.hashes, qry.hashes
        if len(rh) == 0 or len(qh) == 0:
            common_r = common_q = np.zeros(0, np.int64)
            ref_pos, query_pos = 0, 0
        else:
            # the walk stops when either list is exhausted: positions consumed are those <= min(last, last), except
            # that the list that ends first is consumed completely and the other up to (and including) equal hashes
            _, common_r, common_q = np.intersect1d(rh, qh, assume_unique=True, return_indices=True)
            t = min(int(rh[-1]), int(qh[-1]))
            if int(rh[-1]) < int(qh[-1]):
                ref_pos, query_pos = len(rh), int(np.searchsorted(qh, np.uint64(t), side="left"))
            elif int(qh[-1]) < int(rh[-1]):
                ref_pos, query_pos = int(np.searchsorted(rh, np.uint64(t), side="left")), len(qh)
            else:
                ref_pos, query_pos = len(rh), len(qh)
            # a common hash equal to t is consumed on both sides
            if len(common_r) and int(rh[common_r[-1]]) == t:
                ref_pos = max(ref_pos, int(common_r[-1]) + 1)
                query_pos = max(query_pos, int(common_q[-1]) + 1)
        common = 0
        ref_count = query_count = 0
        mean = m2 = m3 = m4 = 0.0
        for ir, iq in zip(common_r, common_q):
            ref_count += int(ref.counts[ir])
            query_count += int(qry.counts[iq])
            n = float(common) + 1.0
            fc = float(qry.counts[iq])
            delta = fc - mean
            delta_n = delta / n
            delta_n2 = delta_n * delta_n
            term1 = delta * delta_n * (n - 1.0)
            mean += delta_n
            m4 += term1 * delta_n2 * (n * n - 3.0 * n + 3.0) + 6.0 * delta_n2 * m2 - 4.0 * delta_n * m3
            m3 += term1 * delta_n * (n - 2.0) - 3.0 * delta_n * m2
            m2 += term1
            common += 1
        fcommon = float(common)
        var = _fdiv(m2, fcommon)
        skew = _fdiv(math.sqrt(fcommon) * m3, _powf(m2, 1.5))
        kurt = _fdiv(fcommon * m4, m2 * m2) - 3.0
        return (common, ref_pos, query_pos, ref_count, query_count, var, skew, kurt)

    def compare_matrix(self, *sketches: "Sketch") -> np.ndarray:
        """compare_matrix(self, *sketches) -> int32 array [len(sketches), len(self)] (python.rs:570-576, minmer_matrix)"""
        for sk in sketches:
            _require_sketch(sk)
        if len(self.s.hashes) == 0 and any(len(sk.s.hashes) for sk in sketches):
            raise PanicException("index out of bounds: the len is 0 but the index is 0")   # distance.rs:355 indexes ref_sketch[0]
        return _try(_fb.minmer_matrix, self.s.hashes, [(sk.s.hashes, sk.s.counts) for sk in sketches])

    @property
    def counts(self) -> np.ndarray:
        """python.rs:578-583"""
        return self.s.counts.astype(np.int32)

    @counts.setter
    def counts(self, value):
        """python.rs:585-608: entries whose new count is 0 are dropped"""
        val = np.asarray(value)
        if val.ndim != 1 or val.dtype.kind not in "iu":
            raise TypeError("counts must be a one-dimensional int32 array")
        if len(val) != len(self.s.hashes):
            raise FinchError("counts must be same length as sketch")
        val = val.astype(np.int64)
        neg = np.flatnonzero(val < 0)
        if len(neg):
            raise FinchError(f"Negative count {int(val[neg[0]])} not supported")
        keep = np.flatnonzero(val > 0)
        s = self.s
        s.hashes = s.hashes[keep].copy()
        s.kmers = [s.kmers[i] if i < len(s.kmers) else b"" for i in keep]
        s.counts = val[keep].astype(np.uint32)
        s.extra_counts = s.extra_counts[keep].copy()

    def copy(self) -> "Sketch":
        return Sketch._wrap(self.s.clone())


def _fdiv(a: float, b: float) -> float:
    """f64 division with IEEE results where Python raises"""
    if b == 0.0:
        if a != a or a == 0.0:
            return math.nan
        return math.copysign(math.inf, a) * math.copysign(1.0, b)
    return a / b


def _powf(x: float, y: float) -> float:
    try:
        return math.pow(x, y)
    except (ValueError, OverflowError):
        return math.nan if x < 0 else math.inf


def _require_sketch(x):
    if not isinstance(x, Sketch):
        raise TypeError(f"argument 'sketch': '{type(x).__name__}' object cannot be converted to 'Sketch'")


# ---- #[pyclass] SketchIter / Multisketch (python.rs:105-308) ----------------------------------------------------
class SketchIter:
    def __init__(self, sketches: List[Sketch]):
        self._sketches = list(sketches)
        self._at = 0

    def __iter__(self):
        return self

    def __next__(self) -> Sketch:
        if self._at >= len(self._sketches):
            raise StopIteration
        self._at += 1
        return self._sketches[self._at - 1]


def _get_sketch_index(sketches: List[_SketchRs], key) -> int:
    """python.rs:281-308"""
    if isinstance(key, (int, np.integer)):      # key.extract::<isize>() (a bool is an int there as well)
        k, l = int(key), len(sketches)
        if -l <= k < 0:
            return l - k                # sic: past the end; the caller indexes with it and panics
        if 0 <= k < l:
            return k
        raise IndexError("index out of range")
    if isinstance(key, str):
        for i, s in enumerate(sketches):
            if s.name == key:
                return i
        raise KeyError(key)
    raise FinchError("key is not a string or integer")


class Multisketch:
    """A Multisketch is a collection of Sketchs with information about their generation parameters (to make sure
    they're consistant for distance calculation)."""

    def __init__(self):
        raise TypeError("No constructor defined")     # a #[pyclass] without #[new]

    @classmethod
    def _make(cls, sketches: List[_SketchRs]) -> "Multisketch":
        o = cls.__new__(cls)
        o.sketches = sketches
        return o

    @classmethod
    def open(cls, filename: str) -> "Multisketch":
        """open(filename: str): a `.sk`, `.bsk` or `.msh` file -> the Multisketch it holds (python.rs:116-121)"""
        L = _fb.lib()
        h = C.c_void_p()
        _check(L.fb2_sketch_set_open(filename.encode(), C.byref(h)))
        try:
            out = []
            v = _fb._SketchView()
            for i in range(int(L.fb2_sketch_set_len(h))):
                _check(L.fb2_sketch_set_get(h, i, C.byref(v)))
                n = int(v.n)
                hashes = np.ctypeslib.as_array(v.hashes, (n,)).copy() if n else np.zeros(0, np.uint64)
                counts = np.ctypeslib.as_array(v.counts, (n,)).copy() if n else np.zeros(0, np.uint32)
                extras = np.ctypeslib.as_array(v.extras, (n,)).copy() if n else np.zeros(0, np.uint32)
                offs = np.ctypeslib.as_array(v.kmer_offs, (n + 1,)).copy()
                blob = C.string_at(v.kmers, int(offs[n])) if int(offs[n]) else b""
                kmers = [blob[int(offs[q]):int(offs[q + 1])] for q in range(n)]
                out.append(_SketchRs((v.name or b"").decode("utf-8", "replace"), v.seq_length, v.num_valid_kmers,
                                     (v.comment or b"").decode("utf-8", "replace"), hashes, kmers, counts, extras,
                                     _params_from_c(v.params), _fb.FilterParams._from_c(v.filter)))
        finally:
            L.fb2_sketch_set_close(h)
        return cls._make(out)

    @classmethod
    def from_sketches(cls, sketches: Iterable[Sketch]) -> "Multisketch":
        """from_sketches(sketches: List[Sketch]) (python.rs:128-132)"""
        sketches = list(sketches)
        for s in sketches:
            _require_sketch(s)
        return cls._make([s.s.clone() for s in sketches])

    def __repr__(self):
        n = len(self.sketches)
        return f"<Multisketch ({n} {'sketch' if n == 1 else 'sketches'})>"

    def __len__(self):
        return len(self.sketches)

    def __iter__(self):
        return SketchIter([Sketch._wrap(s.clone()) for s in self.sketches])

    def __getitem__(self, key) -> Sketch:
        idx = _get_sketch_index(self.sketches, key)
        if idx >= len(self.sketches):
            raise PanicException(f"index out of bounds: the len is {len(self.sketches)} but the index is {idx}")
        return Sketch._wrap(self.sketches[idx].clone())

    def __delitem__(self, key):
        idx = _get_sketch_index(self.sketches, key)
        if idx >= len(self.sketches):
            raise PanicException(f"removal index (is {idx}) should be < len (is {len(self.sketches)})")
        del self.sketches[idx]

    def __contains__(self, key) -> bool:
        if not isinstance(key, str):
            raise TypeError("argument 'key': 'str' expected")
        return any(s.name == key for s in self.sketches)

    def save(self, filename: str):
        """save(self, filename: str): always the finch binary format (`.bsk`), whatever the name (python.rs:180-186)"""
        self._save_as(filename, FILE_BSK)

    def _save_as(self, filename: str, file_format: int):
        L = _fb.lib()
        h = C.c_void_p()
        _check(L.fb2_sketch_set_new(C.byref(h)))
        try:
            for s in self.sketches:
                n = len(s.hashes)
                kmers = [s.kmers[i] if i < len(s.kmers) else b"" for i in range(n)]
                offs = np.zeros(n + 1, np.uint64)
                if n:
                    offs[1:] = np.cumsum([len(x) for x in kmers])
                blob = np.frombuffer(b"".join(kmers) or b"\0", np.uint8)
                hh, cc, xx = (np.ascontiguousarray(s.hashes, np.uint64), np.ascontiguousarray(s.counts, np.uint32),
                              np.ascontiguousarray(s.extra_counts, np.uint32))
                v = _fb._SketchView(s.name.encode(), s.comment.encode(), s.seq_length, s.num_valid_kmers, s.sketch_params._c(),
                                    s.filter_params._c(), n, hh.ctypes.data_as(C.POINTER(C.c_uint64)),
                                    cc.ctypes.data_as(C.POINTER(C.c_uint32)), xx.ctypes.data_as(C.POINTER(C.c_uint32)),
                                    blob.ctypes.data_as(C.POINTER(C.c_uint8)), offs.ctypes.data_as(C.POINTER(C.c_uint64)))
                _check(L.fb2_sketch_set_add(h, C.byref(v)))
            rc = L.fb2_sketch_set_save(h, filename.encode(), file_format)
            if rc == _fb.EIO:
                raise FinchError(f"Could not create {filename}")
            _check(rc)
        finally:
            L.fb2_sketch_set_close(h)

    def add(self, sketch: Sketch):
        _require_sketch(sketch)
        self.sketches.append(sketch.s.clone())

    def best_match(self, query: Sketch) -> Tuple[int, Sketch]:
        """best_match(self, query) -> (index, Sketch) of the largest containment of the query (python.rs:202-216)"""
        _require_sketch(query)
        best, max_containment = 0, 0.0
        for ix, (cont, _) in enumerate(_distances(query.s, self.sketches)):
            if cont > max_containment:
                max_containment, best = cont, ix
        if best >= len(self.sketches):
            raise PanicException("index out of bounds: the len is 0 but the index is 0")
        return best, Sketch._wrap(self.sketches[best].clone())

    def filter_to_matches(self, query: Sketch, threshold: float):
        """keep the sketches whose containment of the query is >= threshold (python.rs:223-235)"""
        _require_sketch(query)
        d = _distances(query.s, self.sketches)
        self.sketches = [s for s, (cont, _) in zip(self.sketches, d) if cont >= threshold]

    def filter_to_names(self, names: list):
        """python.rs:242-248"""
        if not isinstance(names, list):
            raise TypeError("argument 'names': 'list' expected")
        for n in names:
            if not isinstance(n, str):
                raise TypeError("'str' expected")
        keep = set(names)
        self.sketches = [s for s in self.sketches if s.name in keep]


# ---- #[pyfunction] sketch_file (python.rs:645-679) -------------------------------------------------------------
def sketch_file(filename: str, n_hashes: int = 1000, final_size: Optional[int] = None, kmer_length: int = 21,
                filter: bool = True, seed: int = 0, no_strict: bool = False) -> Sketch:
    """From the FASTA or FASTQ file path, create a Sketch."""
    sp = _fb.SketchParams.mash(n_hashes, n_hashes if final_size is None else final_size, bool(no_strict), kmer_length, seed)
    fp = _fb.FilterParams(bool(filter), (None, None), 1.0, 0.1)
    sks = _try(_fb.sketch_files, [filename], sp, fp)
    s = _SketchRs.from_engine(sks[0])
    s.filter_params = sks[0].filter_params
    return Sketch._wrap(s)

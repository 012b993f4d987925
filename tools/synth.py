"""Deterministic synthetic inputs over tools/libfb2_synth.so -- benchmark / test infrastructure,
deliberately outside the product library (the reference arm of bench.py must not map libfinch_b200.so)."""
import ctypes as C
import os
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfb2_synth.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    L = C.CDLL(LIB_PATH)
    vp, sz = C.c_void_p, C.c_size_t
    L.fb2_synth_genome.argtypes = [vp, sz, C.c_uint64]
    L.fb2_synth_genome.restype = sz
    L.fb2_synth_fasta.argtypes = [vp, sz, sz, C.c_uint32, C.c_uint32, C.c_double, C.c_double, C.c_uint64]
    L.fb2_synth_fasta.restype = sz
    L.fb2_synth_fastq.argtypes = [vp, sz, vp, sz, C.c_uint64, C.c_uint32, C.c_double, C.c_uint64, C.c_uint64,
                                  C.POINTER(C.c_uint64)]
    L.fb2_synth_fastq.restype = sz
    L.fb2_synth_sketches.argtypes = [vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64]
    L.fb2_synth_sketches.restype = None
    _lib = L
    return L


def synth_genome(n_bases, seed):
    out = np.empty(n_bases, np.uint8)
    lib().fb2_synth_genome(out.ctypes.data, n_bases, seed)
    return out


def synth_fasta(n_bases, n_records=1, line_width=80, lower_frac=0.0, n_frac=0.0, seed=1):
    need = lib().fb2_synth_fasta(None, 0, n_bases, n_records, line_width, lower_frac, n_frac, seed)
    out = np.empty(need, np.uint8)
    lib().fb2_synth_fasta(out.ctypes.data, need, n_bases, n_records, line_width, lower_frac, n_frac, seed)
    return out


def fastq_nbytes(n_reads, read_len, first_read_id=0):
    """Exact size of synth_fastq's output: '@r<id>\\n' + seq + '\\n+\\n' + qual + '\\n' per read."""
    total, lo = 0, first_read_id
    hi = first_read_id + n_reads
    while lo < hi:
        digits = len(str(lo))
        nxt = min(hi, 10 ** digits)
        total += (nxt - lo) * (2 + digits + 1 + read_len + 1 + 2 + read_len + 1)
        lo = nxt
    return total


def synth_fastq(genome, n_reads, read_len=150, err_rate=0.005, seed=3, first_read_id=0, out=None):
    genome = np.ascontiguousarray(genome, np.uint8)
    need = fastq_nbytes(n_reads, read_len, first_read_id)
    if out is None:
        out = np.empty(need, np.uint8)
    addr = out.ctypes.data if isinstance(out, np.ndarray) else int(out)
    nb = C.c_uint64()
    got = lib().fb2_synth_fastq(addr, need, genome.ctypes.data, genome.size, n_reads, read_len, err_rate, seed,
                                first_read_id, C.byref(nb))
    assert got == need, (got, need)
    return out, int(nb.value)


def synth_fastq_parallel(genome, n_reads, read_len, err_rate, seed, first_id=0, out_ptr=None, threads=None):
    """The same bytes as synth_fastq, generated in parallel slices (reads have their own RNG streams).
    -> (numpy buffer or None when out_ptr is given, nbytes, nbases)"""
    threads = threads or min(32, os.cpu_count() or 1)
    genome = np.ascontiguousarray(genome, np.uint8)
    need = fastq_nbytes(n_reads, read_len, first_id)
    if out_ptr is None:
        buf = np.empty(need, np.uint8)
        base = buf.ctypes.data
    else:
        buf, base = None, int(out_ptr)
    per = (n_reads + threads - 1) // threads
    jobs, off = [], 0
    for t in range(threads):
        a, b = t * per, min(n_reads, (t + 1) * per)
        if a >= b:
            break
        nb = fastq_nbytes(b - a, read_len, first_id + a)
        jobs.append((a, b - a, off, nb))
        off += nb
    assert off == need

    def work(a, n, o, nb):
        got = lib().fb2_synth_fastq(base + o, nb, genome.ctypes.data, genome.size, n, read_len, err_rate, seed,
                                    first_id + a, None)
        assert got == nb

    ths = [threading.Thread(target=work, args=j) for j in jobs]
    [t.start() for t in ths]
    [t.join() for t in ths]
    return buf, need, n_reads * read_len


def synth_sketches(n_sk, n_hashes=1000, n_clusters=1000, seed=5, first=0, count=None):
    """C5-shaped input: rows [first, first+count) of `n_sk` sketches of `n_hashes` sorted distinct
    u64; sketch i belongs to cluster i % n_clusters and shares 50-95 % of its hashes with the cluster's base
    set, the rest is its own.  Every row is a pure function of (seed, i): any slice can be generated anywhere."""
    count = n_sk - first if count is None else count
    out = np.empty((count, n_hashes), np.uint64)
    lib().fb2_synth_sketches(out.ctypes.data, count, n_hashes, n_clusters, seed, first, n_sk)
    return out

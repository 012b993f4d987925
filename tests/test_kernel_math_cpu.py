"""CPU checks of the exact inline arithmetic the CUDA kernels call (common.cuh, host build)
against the oracle: murmur over 2-bit k-mers, rolling canonical choice, byte classes."""
import ctypes as C

import numpy as np
import pytest


def _symbols(seq: bytes):
    """normalize + code: A,C,G,T -> 0..3, anything else kept -> 4 (as pack_kernel emits)."""
    lut = np.full(256, 4, np.uint8)
    for ch, v in ((b"A", 0), (b"C", 1), (b"G", 2), (b"T", 3)):
        lut[ch[0]] = v
    return lut[np.frombuffer(seq, np.uint8)]


@pytest.mark.parametrize("k", [1, 2, 4, 7, 8, 9, 15, 16, 17, 21, 24, 25, 31, 32])
@pytest.mark.parametrize("seed", [0, 42])
def test_stream_matches_oracle(oracle, kmath, k, seed):
    rng = np.random.default_rng(k * 1000 + seed)
    alphabet = np.frombuffer(b"ACGTACGTACGTACGTN", np.uint8)
    seq = bytes(rng.choice(alphabet, size=600))
    want_h, want_rc, want_km = oracle.kmer_stream(seq, k, seed, want_kmers=True)
    sym = _symbols(oracle.normalize(seq))
    n = len(sym)
    h, rc, codes = np.zeros(n, np.uint64), np.zeros(n, np.uint8), np.zeros(n, np.uint64)
    m = kmath.km_stream(sym.ctypes.data, n, k, seed, h.ctypes.data, rc.ctypes.data, codes.ctypes.data)
    assert m == len(want_h)
    assert np.array_equal(h[:m], want_h)
    assert np.array_equal(rc[:m], want_rc)
    buf = (C.c_uint8 * k)()
    for i in range(0, m, 37):
        kmath.km_codes_to_ascii(int(codes[i]), k, buf)
        assert bytes(buf) == want_km[i].tobytes()


def test_palindrome_reports_rc(oracle, kmath):
    seq = b"ACGT" * 8  # every even-k window of ACGT repeats is its own reverse complement
    h_o, rc_o = oracle.kmer_stream(seq, 4, 0)
    sym = _symbols(seq)
    n = len(sym)
    h, rc, codes = np.zeros(n, np.uint64), np.zeros(n, np.uint8), np.zeros(n, np.uint64)
    m = kmath.km_stream(sym.ctypes.data, n, 4, 0, h.ctypes.data, rc.ctypes.data, codes.ctypes.data)
    assert m == len(h_o) and np.array_equal(rc[:m], rc_o) and np.array_equal(h[:m], h_o)
    assert rc_o[0] == 1  # ACGT == revcomp(ACGT): tie goes to the rc slice


def test_classify_matches_normalize(oracle, kmath):
    for c in range(256):
        cls = kmath.km_classify(c)
        norm = oracle.normalize(bytes([c]))
        if cls <= 3:
            assert norm == b"ACGT"[cls:cls + 1]
        elif cls in (5, 6, 8):          # space/tab, newline, CR: removed
            assert norm == b""
        else:                           # kept, not a base
            assert len(norm) == 1 and norm not in (b"A", b"C", b"G", b"T")


def test_murmur_bytes(oracle, kmath):
    rng = np.random.default_rng(5)
    for n in list(range(0, 40)) + [63, 64, 65, 255]:
        b = bytes(rng.integers(0, 256, size=n, dtype=np.uint8))
        for seed in (0, 42, 2**63 + 5):
            assert kmath.km_murmur_bytes(b, n, seed) == oracle.hash_f(b, seed)

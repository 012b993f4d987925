"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol declared in
include/finch_b200.h, the host-only logic (filters, distance epilogue, generators) matches the
oracle, and compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import sys
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rust_sys_crate_declares_every_entry_point():
    """rust/finch_b200-sys is shipped uncompiled (no Rust toolchain here): at least its extern block must name every
    function of the header, so that the two cannot drift apart unnoticed."""
    hdr = open(os.path.join(ROOT, "include", "finch_b200.h")).read()
    rs = open(os.path.join(ROOT, "rust", "finch_b200-sys", "src", "lib.rs")).read()
    declared = set(re.findall(r"\b(fb2_[a-z0-9_]+)\s*\(", hdr))
    in_rust = set(re.findall(r"pub fn (fb2_[a-z0-9_]+)\s*\(", rs))
    assert declared == in_rust, (sorted(declared - in_rust), sorted(in_rust - declared))


def test_library_exports_every_declared_symbol(fb):
    hdr = open(os.path.join(ROOT, "include", "finch_b200.h")).read()
    declared = set(re.findall(r"\b(fb2_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = C.CDLL(fb.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(fb.EXPORTS)
    assert b"sm_100a" in fb.lib().fb2_version()


def test_no_cpu_fallback(fb):
    if fb.lib().fb2_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(fb.FinchError) as e:
        fb.MashSketcher(10, 21, 0)
    assert e.value.code == fb.ECUDA and "no CPU fallback" in e.value.message
    with pytest.raises(fb.FinchError) as e:
        fb.raw_distance([1, 2], [2, 3])
    assert e.value.code == fb.ECUDA


def test_product_does_not_touch_oracle():
    """The product package never imports, includes or links anything under oracle/."""
    pkg = os.path.join(ROOT, "finch_rs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "finch_oracle" not in txt and "libfinch_oracle" not in txt, f
                assert not re.search(r"^\s*(import|from)\s+oracle", txt, re.M), f
                assert not re.search(r"#include\s+[\"<].*oracle", txt), f
                assert not re.search(r"\bfo_[a-z]+\s*\(", txt), f


def test_guess_filter_threshold_reference_cases(fb):  # filtering.rs:197-327
    g = fb.guess_filter_threshold
    assert [g([], 0.2), g([1], 0.2), g([1, 1], 0.2), g([1, 9], 0.2), g([1, 10, 10, 9], 0.1),
            g([1, 1, 2, 4], 0.1), g([2], 1.0)] == [1, 1, 1, 8, 8, 1, 2]


def test_filter_counts_matches_oracle(fb, oracle):
    rng = np.random.default_rng(2)
    for trial in range(60):
        n = int(rng.integers(0, 400))
        # error k-mers (count 1-2) + genomic k-mers (Poisson coverage) + a few strand-biased adapters
        counts = np.where(rng.random(n) < 0.6, rng.integers(1, 3, size=n), rng.poisson(40, size=n) + 1).astype(np.uint32)
        extras = rng.binomial(counts, 0.5).astype(np.uint32)
        biased = rng.random(n) < 0.05
        extras[biased] = 0
        hashes = np.sort(rng.integers(0, 2**63, size=n, dtype=np.uint64))
        on = [True, False, None][trial % 3]
        abun = [(None, None), (2, None), (None, 30), (3, 50)][trial % 4]
        err, strand = [0.21, 0.0, 0.5][trial % 3], [0.1, 0.0][trial % 2]
        fmt = [fb.FORMAT_FASTQ, fb.FORMAT_FASTA][trial % 2]
        ofp = oracle.make_filter(on if on is not None else (fmt == fb.FORMAT_FASTQ), abun, err, strand)
        keep = oracle.filter_counts(ofp, counts, extras)
        h, c, x, fp = fb.filter_counts(fb.FilterParams(on, abun, err, strand), hashes, counts, extras, fmt)
        assert np.array_equal(h, hashes[keep]) and np.array_equal(c, counts[keep]) and np.array_equal(x, extras[keep])
        assert fp.abun_filter[0] == (int(ofp.abun_low) if ofp.has_abun_low else None)
        assert fp.filter_on == bool(ofp.filter_on == 1)


def test_distance_finish_matches_oracle(fb, oracle):
    rng = np.random.default_rng(3)
    for _ in range(200):
        common = int(rng.integers(0, 50)); i = common + int(rng.integers(0, 50)); j = common + int(rng.integers(0, 50))
        cont, jac, md, com, tot = fb._finish_pair((common, i, j), 21)
        assert cont == (0.0 if j == 0 else common / j)
        assert tot == i - common + j and jac == (1.0 if tot == 0 else common / tot)
        assert md == oracle.mash_distance(jac, 21)


def test_generators_are_deterministic_and_parseable(fb, synth, oracle):
    g = synth.synth_genome(5000, 2)
    assert set(np.unique(g).tolist()) <= set(b"ACGT") and np.array_equal(g, synth.synth_genome(5000, 2))
    fq, nb = synth.synth_fastq(g, 1234, 150, 0.005, 3, first_read_id=95)
    assert nb == 1234 * 150 and len(fq) == synth.fastq_nbytes(1234, 150, 95)
    rc, fmt, recs = oracle.parse_fastx(fq.tobytes())
    assert rc == oracle.OK and fmt == oracle.FMT_FASTQ and len(recs) == 1234 and all(len(r) == 150 for r in recs)
    # slices by read id compose to the same bytes
    a, _ = synth.synth_fastq(g, 600, 150, 0.005, 3, first_read_id=95)
    b, _ = synth.synth_fastq(g, 634, 150, 0.005, 3, first_read_id=695)
    assert np.array_equal(np.concatenate([a, b]), fq)
    fa = synth.synth_fasta(100_000, n_records=3, line_width=60, lower_frac=0.02, n_frac=0.01, seed=4).tobytes()
    rc, fmt, recs = oracle.parse_fastx(fa)
    assert rc == oracle.OK and fmt == oracle.FMT_FASTA and len(recs) == 3
    assert sum(len(oracle.normalize(r)) for r in recs) == 100_000


def test_oracle_literal_equals_closed_form(oracle):
    """heap+map push == bottom-s of the whole multiset (both sketchers)."""
    rng = np.random.default_rng(4)
    for trial in range(30):
        seq = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), size=int(rng.integers(30, 3000))))
        k = int(rng.integers(3, 12)); s = int(rng.integers(0, 200))
        h, rc = oracle.kmer_stream(seq, k, 1)
        uniq, inv = np.unique(h, return_inverse=True)
        cnt = np.bincount(inv, minlength=len(uniq)); ext = np.bincount(inv, weights=rc, minlength=len(uniq))
        m = oracle.Sketcher.mash(s, k, 1); m.process(seq); v = m.to_vec()
        assert np.array_equal(v["hashes"], uniq[:s]) and np.array_equal(v["counts"], cnt[:s])
        assert np.array_equal(v["extras"], ext[:s].astype(np.uint32))
        scale = 0.2
        sc = oracle.Sketcher.scaled(s, scale, k, 1); sc.process(seq); v = sc.to_vec()
        small = int(np.searchsorted(uniq, sc.max_hash(), side="right"))
        keep = small if s == 0 else max(small, min(s, len(uniq)))
        assert np.array_equal(v["hashes"], uniq[:keep]) and np.array_equal(v["counts"], cnt[:keep])


def build_hpp_mirror_test():
    """tests/hpp_mirror_test.cpp: the header-only C++ mirror (finch_rs_b200/host/finch_b200.hpp) compiled for real."""
    import subprocess
    src = os.path.join(ROOT, "tests", "hpp_mirror_test.cpp")
    exe = os.path.join(ROOT, "tests", "_hpp_mirror_test")
    libdir = os.path.join(ROOT, "finch_rs_b200")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-Wextra", "-o", exe, src, "-L" + libdir, "-lfinch_b200",
                           "-Wl,-rpath," + libdir])
    return exe


def test_cpp_mirror_header_compiles_and_links(fb):
    import subprocess
    exe = build_hpp_mirror_test()
    out = subprocess.run([exe, "--link-only"], capture_output=True, text=True)
    assert out.returncode == 0 and "sm_100a" in out.stdout, out.stderr


def test_cpp_mirror_sketch_files(fb, tmp_path):
    """open_sketch_file / write_sketch_file of the C++ mirror: three formats written and read back (host only)."""
    import subprocess
    exe = build_hpp_mirror_test()
    out = subprocess.run([exe, "--files", str(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0 and "files ok" in out.stdout, out.stderr
    # and the Python module reads what the C++ mirror wrote
    sys.path.insert(0, os.path.join(ROOT, "python"))
    import finch
    ms = finch.Multisketch.open(str(tmp_path / "m.bsk"))
    assert [s.name for s in ms] == ["a.fa", "b.fa"] and ms["a.fa"].hashes == [(10, b"ACGT", 1, 0), (17, b"CCCC", 2, 1), (24, b"TTGA", 3, 2)]


def test_sketch_stream_rejects_broken_compressed_bytes_before_any_device_work(fb):
    """fb2_sketch_stream inflates gzip / bzip2 / xz bytes on the host first: a truncated or corrupt stream is an I/O error
    whether or not there is a device."""
    import gzip
    data = gzip.compress(b"@r\nACGT\n+\nIIII\n" * 500)
    sp, fp = fb.SketchParams.mash(10, 10, True, 3, 0), fb.FilterParams(False, (None, None), 0.0, 0.0)
    with pytest.raises(fb.FinchError) as ei:
        fb.sketch_stream(data[:-20], "cut.fq.gz", sp, fp)
    assert ei.value.code == fb.EIO and "truncated gzip stream" in str(ei.value)
    bad = bytearray(data); bad[40] ^= 0xFF; bad[41] ^= 0xFF
    with pytest.raises(fb.FinchError) as ei:
        fb.sketch_stream(bytes(bad), "bad.fq.gz", sp, fp)
    assert ei.value.code == fb.EIO


def test_parameters_reports_what_the_reference_sketchers_report():
    """Quirks Q7 / Q8 (mash.rs:104-112, scaled.rs:102-109) in the mirror's parameters(), without a device."""
    import finch_rs_b200 as m

    class Fake(m._Sketcher):
        def __init__(self, params):
            self.params = params
    p = Fake(m.SketchParams.mash(200000, 1000, True, 21, 7)).parameters()
    assert (p.kmers_to_sketch, p.final_size, p.no_strict, p.kmer_length, p.hash_seed) == (200000, 200000, False, 21, 7)
    p = Fake(m.SketchParams.scaled(10, 31, 0.003, 3)).parameters()
    max_hash = (2**64 - 1) // 333                     # (1. / 0.003) as u64 == 333
    assert p.scale == 1.0 / (float(2**64 - 1) / float(max_hash)) and p.scale != 0.003 and p.kmers_to_sketch == 10


def test_host_strip_framing_cpu(oracle):
    """strip.cpp (FB2_HOST_STRIP) on the CPU, no device: the lines it ships are exactly the records' sequence()
    slices the oracle's reader yields, its verdict on malformed streams is the reader's, for any thread count."""
    import subprocess
    import gen
    exe = os.path.join(ROOT, "tests", "_strip_host_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "strip_host_test.cpp"),
                           os.path.join(ROOT, "finch_rs_b200", "csrc", "strip.cpp"), "-lpthread"])
    rng = np.random.default_rng(2)
    body = gen.fastq(rng, n_records=3000, max_len=400, messy=0.02)
    cases = [body, body.replace(b"\n", b"\r\n"), body[:-1], gen.fastq(rng, n_records=40, max_len=60_000, messy=0.0),
             b"@r\n\n+\n", b"@r\nAC\n+\nII", body + b"\n\r\n\n", b"@r\nA\n+\n\r\r", body + b"@r\nA\n+\n", body + b"@x\nACGT\n+\nII\n" + body,
             body[:5000] + b"\n\n\n\n" + body[5000:], body + b"@q\nAC\n-\nII\n", b"@a\nAC\n+\nII\n\n\n\n\n\n\n\n\n"]
    for data in cases:
        rc, fmt, recs = oracle.parse_fastx(data)
        for threads in (1, 3, 8):
            out = subprocess.run([exe, str(threads)], input=data, capture_output=True).stdout
            head, _, lines = out.partition(b"\n")
            n, bases, invalid = (int(x) for x in head.split())
            assert bool(invalid) == (rc != oracle.OK), (data[-30:], threads)
            if rc == oracle.OK:
                assert lines == b"".join(r + b"\n" for r in recs)
                assert bases == sum(len(r) for r in recs)

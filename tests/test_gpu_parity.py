"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle.
Integer / byte work => bit-exact: identical hash lists, counts, extra_counts, k-mers, totals."""
import os

import numpy as np
import pytest

import gen

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def assert_same(fbres, ora, k):
    h, c, x, km, seq_len, n_kmers, fmt = fbres
    assert len(h) == len(ora["hashes"]), (len(h), len(ora["hashes"]))
    assert np.array_equal(h, ora["hashes"])
    assert np.array_equal(c, ora["counts"])
    assert np.array_equal(x, ora["extras"])
    assert [km[i, :k].tobytes() for i in range(len(h))] == ora["kmers"]


def oracle_sketch(oracle, data, kind, size, k, seed, scale=0.001):
    """records through fo_process; returns to_vec + totals."""
    s = oracle.Sketcher.mash(size, k, seed) if kind == "mash" else oracle.Sketcher.scaled(size, scale, k, seed)
    rc, fmt, recs = oracle.parse_fastx(data)
    assert rc == oracle.OK
    for r in recs:
        s.process(r)
    return s.to_vec(), s.total_bases_and_kmers(), fmt


def gpu_sketch(fb, data, kind, size, k, seed, scale=0.001, pieces=None):
    sp = fb.SketchParams.mash(size, size, False, k, seed) if kind == "mash" else fb.SketchParams.scaled(size, k, scale, seed)
    with sp.create_sketcher() as s:
        if pieces is None:
            s.feed_fastx(data, final=True)
        else:
            pos = 0
            for p in pieces:
                s.feed_fastx(data[pos:pos + p], final=False)
                pos += p
            s.feed_fastx(data[pos:], final=True)
        return s.to_arrays(), s.total_bases_and_kmers(), s.format()


# ---- the reference's own unit tests, through the mirrored API ----------------------------------
def _push4(q):
    q.push(b"ca", 0); q.push(b"cc", 1); q.push(b"ac", 0); q.push(b"ac", 1)
    return q.to_vec(2)


@pytest.mark.parametrize("mk", ["mash", "scaled1", "scaled1000"])
def test_minhashkmers(fb, mk):  # mash.rs:115-134, scaled.rs:118-161
    q = {"mash": lambda: fb.MashSketcher(3, 2, 42), "scaled1": lambda: fb.ScaledSketcher(3, 1.0, 2, 42),
         "scaled1000": lambda: fb.ScaledSketcher(3, 0.001, 2, 42)}[mk]()
    a = _push4(q)
    assert [e.kmer for e in a] == [b"cc", b"ca", b"ac"]
    assert [(e.count, e.extra_count) for e in a] == [(1, 1), (1, 0), (2, 1)]
    assert a[0].hash < a[1].hash < a[2].hash
    assert q.total_bases_and_kmers() == (0, 4)


def test_minhashkmers_eviction(fb):  # scaled.rs:163-176
    q = fb.ScaledSketcher(1, 0.01, 4, 42)
    q.push(b"AAAA", 0); q.push(b"AGTA", 0); q.push(b"CCCC", 1); q.push(b"ATAA", 0)
    a = q.to_vec(4)
    assert len(a) == 3 and all(e.kmer != b"AAAA" for e in a)


def test_minhashkmers_pure_scaled_empty(fb):  # scaled.rs:178-200
    assert _push4(fb.ScaledSketcher(0, 0.001, 2, 42)) == []


def test_pure_scaled_property(fb):  # scaled.rs:202-213
    rng = np.random.default_rng(11)
    seq = gen.rand_seq(rng, 700)
    q = fb.ScaledSketcher(0, 1.0 / 100.0, 2, 42)
    for i in range(len(seq) - 3):
        q.push(seq[i:i + 4], 0)
    assert all(e.hash <= (2**64 - 1) // 100 for e in q.to_vec(4))


def test_counts_saturate_at_u32_max(fb):  # mash.rs:48-49: count.0.saturating_add(1), count.1.saturating_add(extra)
    for mk in ("mash", "scaled"):
        q = fb.MashSketcher(3, 2, 42) if mk == "mash" else fb.ScaledSketcher(3, 1.0, 2, 42)
        q.push(b"ac", 1); q.push(b"ac", 1); q.push(b"ca", 0)
        v = {e.kmer: e for e in q.to_vec(2)}
        assert (v[b"ac"].count, v[b"ac"].extra_count) == (2, 2)
        q.debug_bump(v[b"ac"].hash, 2**32 - 4, 2**32 - 5)            # totals now (2^32 - 2, 2^32 - 3)
        v = {e.kmer: e for e in q.to_vec(2)}
        assert (v[b"ac"].count, v[b"ac"].extra_count) == (2**32 - 2, 2**32 - 3)
        q.push(b"ac", 1)                                             # (2^32 - 1, 2^32 - 2): the largest u32, exact
        v = {e.kmer: e for e in q.to_vec(2)}
        assert (v[b"ac"].count, v[b"ac"].extra_count) == (2**32 - 1, 2**32 - 2)
        for _ in range(3):
            q.push(b"ac", 1)                                         # saturated: stays at u32::MAX
        v = {e.kmer: e for e in q.to_vec(2)}
        assert (v[b"ac"].count, v[b"ac"].extra_count) == (2**32 - 1, 2**32 - 1)
        assert (v[b"ca"].count, v[b"ca"].extra_count) == (1, 0)
        q.close()


def test_cpp_mirror_runs_the_reference_unit_tests(fb):
    """finch_rs_b200/host/finch_b200.hpp through tests/hpp_mirror_test.cpp: mash.rs:115-134, scaled.rs:118-176."""
    import subprocess
    from test_abi_cpu import build_hpp_mirror_test
    out = subprocess.run([build_hpp_mirror_test()], capture_output=True, text=True)
    assert out.returncode == 0 and "hpp mirror ok" in out.stdout, out.stdout + out.stderr


def test_pushed_kmers_keep_their_length(fb):
    q = fb.MashSketcher(5, 2, 0)
    q.push(b"ACGTACGT", 0); q.push(b"ac", 1)
    h, c, x, km, *_ = q.to_arrays()
    got = {bytes(km[i]).rstrip(b"\0") for i in range(len(h))}
    assert got == {b"ACGTACGT", b"ac"}
    q.close()


def test_longer_sequence_seed42(fb):  # mash.rs:136-154
    q = fb.MashSketcher(100, 21, 42)
    q.process(b"ACACGGAAATCCTCACGTCGCGGCGCCGGGC")
    assert [e.hash for e in q.to_vec()] == [
        3186265289206375993, 3197567229193635484, 5157287830980272133, 7515070071080094037,
        9123665698461883699, 9650810550987401968, 10462414310441547028, 12872951831549606632,
        13584836512372089324, 14093285637546356047, 16069721578136260683]
    assert q.total_bases_and_kmers() == (31, 11)


KC_KMERS = [b"ATGCTAGCTACGTAACGTCGC", b"CAGTCGATCGATCGTAGCTGA", b"CTCAGATGCTGAGCCGGTCTA",
            b"GCTAGCTAGCATCGCTAGCTA", b"GACTAGCTAGCTAGCTAGCGA", b"CGCTAGCTACGATCGATCGAC",
            b"TAATTTATACGGGCCTATTAA", b"GCATCAGCTAGCATCGCTGTA", b"AGCCGGTCTACTACTACACAT",
            b"AAGGCCTAACTTAATAGGCCC"]


@pytest.mark.parametrize("kind", ["mash", "scaled"])
def test_cli_golden_query_fa(fb, kind):  # cli/tests/test_cli.rs:80-149
    data = open(os.path.join(GOLD, "query.fa"), "rb").read()
    sp = fb.SketchParams.from_cli(kind, n_hashes=10, scale=0.001)
    fp = fb.FilterParams(None, (None, None), 1.0 * 21 / 100, 0.1)
    sk = fb.sketch_stream(data, "tests/data/query.fa", sp, fp)
    assert [sk.kmer_bytes(i) for i in range(len(sk))] == KC_KMERS
    assert sk.num_valid_kmers == 339 and sk.seq_length == 405
    assert sk.format == fb.FORMAT_FASTA and sk.filter_params.filter_on is False


# ---- randomized parity against the oracle ------------------------------------------------------
CASES = []
for i, (kind, size, k) in enumerate([("mash", 50, 21), ("mash", 1000, 21), ("mash", 7, 5), ("mash", 300, 31),
                                     ("mash", 100, 32), ("mash", 100, 16), ("mash", 64, 1), ("scaled", 20, 21),
                                     ("scaled", 0, 11), ("scaled", 500, 31), ("mash", 100000, 13),
                                     # k > 32 (the reference takes any u8, mod.rs:59): exact multi-word kernel
                                     ("mash", 200, 33), ("mash", 100, 48), ("scaled", 50, 64), ("mash", 300, 255),
                                     ("mash", 100, 100)]):
    CASES.append((i, kind, size, k))


@pytest.mark.parametrize("case,kind,size,k", CASES)
@pytest.mark.parametrize("fmt", ["fasta", "fasta_crlf_ragged", "fastq", "fastq_crlf"])
def test_random_small(fb, oracle, case, kind, size, k, fmt):
    rng = np.random.default_rng(100 * case + len(fmt))
    if fmt == "fasta":
        data = gen.fasta(rng, n_records=5, max_len=3000, width=70, messy=0.01)
    elif fmt == "fasta_crlf_ragged":
        data = gen.fasta(rng, n_records=7, max_len=800, width=25, messy=0.03, crlf=True, blank_lines=True,
                         ragged=True, final_newline=bool(case % 2))
    elif fmt == "fastq":
        data = gen.fastq(rng, n_records=60, max_len=250, messy=0.01, final_newline=bool(case % 2))
    else:
        data = gen.fastq(rng, n_records=40, max_len=120, messy=0.05, crlf=True)
    scale = 0.05
    ovec, ototals, ofmt = oracle_sketch(oracle, data, kind, size, k, 7, scale)
    gres, gtotals, gfmt = gpu_sketch(fb, data, kind, size, k, 7, scale)
    assert gfmt == ofmt
    assert gtotals == ototals
    assert_same(gres, ovec, k)


@pytest.mark.parametrize("fmt", ["fasta", "fastq"])
def test_chunk_seams_anywhere(fb, oracle, fmt):
    rng = np.random.default_rng(77)
    data = (gen.fasta(rng, n_records=4, max_len=1500, width=40, messy=0.02, crlf=True, ragged=True)
            if fmt == "fasta" else gen.fastq(rng, n_records=30, max_len=150, crlf=False))
    ovec, ototals, _ = oracle_sketch(oracle, data, "mash", 200, 21, 0)
    for trial in range(4):
        cuts = sorted(rng.integers(1, 40, size=int(rng.integers(1, 60))).tolist())
        pieces = [c for c in cuts if c > 0]
        if sum(pieces) >= len(data):
            pieces = [1, 2, 3]
        gres, gtotals, _ = gpu_sketch(fb, data, "mash", 200, 21, 0, pieces=pieces)
        assert gtotals == ototals
        assert_same(gres, ovec, 21)


def test_gt_inside_fasta_sequence_line(fb, oracle):
    """'>' only starts a record at a line start; inside a sequence line it is just a non-base byte."""
    rng = np.random.default_rng(9)
    lines = []
    for r in range(40):
        lines.append(b">rec%d" % r)
        for _ in range(int(rng.integers(1, 6))):
            s = bytearray(gen.rand_seq(rng, int(rng.integers(30, 90)), 0.0))
            for pos in rng.integers(1, len(s), size=int(rng.integers(0, 3))):
                s[int(pos)] = ord(">")
            lines.append(bytes(s))
    data = b"\n".join(lines) + b"\n"
    for k in (5, 21):
        ovec, ototals, ofmt = oracle_sketch(oracle, data, "mash", 400, k, 0)
        gres, gtotals, gfmt = gpu_sketch(fb, data, "mash", 400, k, 0)
        assert (gfmt, gtotals) == (ofmt, ototals)
        assert_same(gres, ovec, k)


def test_one_byte_first_piece_is_sniffed_with_the_next(fb, oracle):
    """Format sniffing waits for two bytes: a gzip magic split over two pieces is still reported as compressed
    input, and a 1-byte first piece of a plain file changes nothing."""
    sp = fb.SketchParams.mash(50, 50, True, 11, 0)
    with sp.create_sketcher() as s:
        s.feed_fastx(b"\x1f", final=False)
        with pytest.raises(fb.FinchError) as e:
            s.feed_fastx(b"\x8b\x08\x00", final=True)
        assert e.value.code == fb.EUNSUPPORTED
    data = gen.fasta(np.random.default_rng(3), n_records=3, max_len=400)
    ovec, ototals, _ = oracle_sketch(oracle, data, "mash", 50, 11, 0)
    gres, gtotals, _ = gpu_sketch(fb, data, "mash", 50, 11, 0, pieces=[1, 1, 5])
    assert gtotals == ototals
    assert_same(gres, ovec, 11)


def test_process_records_api(fb, oracle):
    """SketchScheme::process per record (mash.rs:67-80), records with newlines / lowercase / N."""
    rng = np.random.default_rng(5)
    recs = [gen.rand_seq(rng, int(rng.integers(0, 400)), 0.03) for _ in range(50)]
    recs[3] = recs[3][:50] + b"\n" + recs[3][50:] + b"\r\n"
    recs[7] = b""
    o = oracle.Sketcher.mash(150, 21, 3)
    g = fb.MashSketcher(150, 21, 3)
    for r in recs:
        o.process(r); g.process(r)
    assert g.total_bases_and_kmers() == o.total_bases_and_kmers()
    ov = o.to_vec()
    assert_same(g.to_arrays(), ov, 21)
    # to_vec is non-destructive: keep going
    more = gen.rand_seq(rng, 5000)
    o.process(more); g.process(more)
    assert_same(g.to_arrays(), o.to_vec(), 21)
    assert g.total_bases_and_kmers() == o.total_bases_and_kmers()


@pytest.mark.parametrize("kind,size,k,scale", [("mash", 1000, 21, 0), ("mash", 200000, 21, 0),
                                               ("scaled", 1000, 31, 0.001), ("scaled", 0, 21, 0.01)])
def test_medium_fasta(fb, synth, oracle, kind, size, k, scale):
    """~3 Mbp multi-record FASTA with lowercase and N runs (C1/C4-shaped, reduced)."""
    data = synth.synth_fasta(3_000_000, n_records=3, line_width=80, lower_frac=0.02, n_frac=0.01, seed=4).tobytes()
    ovec, ototals, _ = oracle_sketch(oracle, data, kind, size, k, 0, scale or 0.001)
    gres, gtotals, _ = gpu_sketch(fb, data, kind, size, k, 0, scale or 0.001)
    assert gtotals == ototals
    assert_same(gres, ovec, k)


def test_medium_fastq_with_coverage(fb, synth, oracle):
    """C2-shaped, reduced: 40k x 150 bp reads from a 20 kbp genome (300x), 0.5% errors, heap 200000."""
    genome = synth.synth_genome(20_000, 2)
    data, nb = synth.synth_fastq(genome, 40_000, 150, 0.005, 3)
    data = data.tobytes()
    ovec, ototals, _ = oracle_sketch(oracle, data, "mash", 200000, 21, 0)
    gres, gtotals, _ = gpu_sketch(fb, data, "mash", 200000, 21, 0)
    assert ototals[0] == nb and gtotals == ototals
    assert_same(gres, ovec, 21)
    # ...and through sketch_stream with the CLI's filter defaults (-f): identical filtered sketch
    sp_o = oracle.mash_params(200000, 1000, False, 21, 0)
    rc, osk = oracle.sketch_stream(data, sp_o, oracle.make_filter(True, (None, None), 0.21, 0.1))
    assert rc == oracle.OK
    sk = fb.sketch_stream(data, "reads.fq", fb.SketchParams.mash(200000, 1000, False, 21, 0),
                          fb.FilterParams(True, (None, None), 0.21, 0.1))
    assert np.array_equal(sk.hashes_u64, osk["hashes"]) and np.array_equal(sk.counts, osk["counts"])
    assert np.array_equal(sk.extra_counts, osk["extras"])
    assert sk.filter_params.abun_filter[0] == osk["min_copies"]
    assert len(sk) == 1000


def test_multi_chunk_stream(fb, synth, oracle, monkeypatch):
    """Chunks of 1 MiB so one stream crosses many chunk seams inside the engine."""
    monkeypatch.setenv("FB2_CHUNK_MB", "1")
    genome = synth.synth_genome(300_000, 9)
    data, _ = synth.synth_fastq(genome, 25_000, 100, 0.01, 5)
    data = data.tobytes()
    assert len(data) > 5 * (1 << 20)
    ovec, ototals, _ = oracle_sketch(oracle, data, "mash", 5000, 21, 0)
    gres, gtotals, _ = gpu_sketch(fb, data, "mash", 5000, 21, 0)
    assert gtotals == ototals
    assert_same(gres, ovec, 21)
    # device-resident feed of the same bytes
    import torch
    t = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    with fb.SketchParams.mash(5000, 5000, False, 21, 0).create_sketcher() as s:
        s.feed_device(t.data_ptr(), t.numel(), final=True)
        assert s.total_bases_and_kmers() == ototals
        assert_same(s.to_arrays(), ovec, 21)


@pytest.mark.parametrize("kind,size,scale", [("mash", 20000, 0.0), ("scaled", 1000, 0.01), ("scaled", 0, 0.02)])
def test_many_chunks_steady_state(fb, synth, oracle, monkeypatch, kind, size, scale):
    """~40 chunks of 1 MiB: most of the stream runs through the asynchronous steady-state path (absorb on
    its own stream under the next chunk's parse kernels, soft threshold updates from the live table
    histogram after every chunk, occasional rebuilds).  High coverage so counts, strand counts and
    first-occurrence k-mers of hot keys accumulate across many chunks."""
    monkeypatch.setenv("FB2_CHUNK_MB", "1")
    genome = synth.synth_genome(150_000, 21)
    data, nb = synth.synth_fastq(genome, 130_000, 150, 0.01, 23)
    data = data.tobytes()
    assert len(data) > 38 * (1 << 20)
    ovec, ototals, _ = oracle_sketch(oracle, data, kind, size, 21, 0, scale or 0.001)
    gres, gtotals, _ = gpu_sketch(fb, data, kind, size, 21, 0, scale or 0.001)
    assert ototals[0] == nb and gtotals == ototals
    assert_same(gres, ovec, 21)
    assert int(ovec["counts"].max()) > 50
    # the A/B switches used for the measurements give the same sketch
    for var in ("FB2_NO_SOFT_THRESHOLD", "FB2_NO_ABSORB_STREAM"):
        monkeypatch.setenv(var, "1")
        gres2, gtotals2, _ = gpu_sketch(fb, data, kind, size, 21, 0, scale or 0.001)
        monkeypatch.delenv(var)
        assert gtotals2 == ototals
        assert_same(gres2, ovec, 21)


def test_sketch_files_gzip(fb, oracle, tmp_path):
    """gzip input (needletail sniffs 1f 8b): single member, concatenated members, and a truncated stream."""
    import gzip
    rng = np.random.default_rng(91)
    fa = gen.fasta(rng, n_records=3, max_len=50000, width=70)
    fq = gen.fastq(rng, n_records=800, max_len=250)
    p1, p2, p3, p4 = (tmp_path / n for n in ("a.fa.gz", "b.fq.gz", "c.fa", "bad.fq.gz"))
    p1.write_bytes(gzip.compress(fa))
    half = len(fq) // 2
    p2.write_bytes(gzip.compress(fq[:half]) + gzip.compress(fq[half:]))       # two members, cut mid-record
    p3.write_bytes(fa)
    p4.write_bytes(gzip.compress(fq)[:-200])
    sp = fb.SketchParams.mash(3000, 200, True, 21, 0)
    fp = fb.FilterParams(False, (None, None), 0.21, 0.1)
    sks = fb.sketch_files([str(p1), str(p2), str(p3)], sp, fp)
    for sk, data in zip(sks, (fa, fq, fa)):
        rc, osk = oracle.sketch_stream(data, oracle.mash_params(3000, 200, True, 21, 0), oracle.make_filter(False, (None, None), 0.21, 0.1))
        assert rc == oracle.OK
        assert np.array_equal(sk.hashes_u64, osk["hashes"]) and np.array_equal(sk.counts, osk["counts"])
        assert (sk.seq_length, sk.num_valid_kmers) == (osk["seq_length"], osk["num_valid_kmers"])
    assert np.array_equal(sks[0].hashes_u64, sks[2].hashes_u64)
    with pytest.raises(fb.FinchError) as ei:
        fb.sketch_files([str(p4)], sp, fp)
    assert "gzip" in str(ei.value)


@pytest.mark.parametrize("crlf", [False, True])
def test_large_fastq_file_framed_from_the_mapping(fb, oracle, tmp_path, monkeypatch, crlf):
    """Large plain FASTQ files are mapped and framed by the host cores straight from the page cache (FB2_BIG_FILE_KB
    lowers "large" for the test): same sketch as the oracle and as the copying path, same errors for malformed files;
    FASTA files of that size keep the parallel-read path."""
    rng = np.random.default_rng(23)
    fq = gen.fastq(rng, n_records=6000, min_len=30, max_len=250, crlf=crlf, final_newline=not crlf)
    fa = gen.fasta(rng, n_records=4, max_len=300000, width=70)
    bad = bytearray(fq)
    cut = fq.index(b"\n+", len(fq) // 2)                  # drop a '+' line's first byte deep inside the file
    del bad[cut + 1]
    pq, pa, pb = tmp_path / "reads.fq", tmp_path / "genome.fa", tmp_path / "bad.fq"
    pq.write_bytes(fq); pa.write_bytes(fa); pb.write_bytes(bytes(bad))
    sp = fb.SketchParams.mash(3000, 200, True, 21, 0)
    fp = fb.FilterParams(False, (None, None), 0.21, 0.1)
    monkeypatch.setenv("FB2_BIG_FILE_KB", "64")
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("FB2_FILE_MMAP", mode)
        res[mode] = fb.sketch_files([str(pq), str(pa)], sp, fp)
        with pytest.raises(fb.FinchError) as ei:
            fb.sketch_files([str(pb)], sp, fp)
        assert ei.value.code == fb.ERECORD, str(ei.value)
    for data, a, b in zip((fq, fa), res["1"], res["0"]):
        rc, osk = oracle.sketch_stream(data, oracle.mash_params(3000, 200, True, 21, 0), oracle.make_filter(False, (None, None), 0.21, 0.1))
        assert rc == oracle.OK
        for sk in (a, b):
            assert np.array_equal(sk.hashes_u64, osk["hashes"]) and np.array_equal(sk.counts, osk["counts"])
            assert np.array_equal(sk.extra_counts, osk["extras"])
            assert (sk.seq_length, sk.num_valid_kmers) == (osk["seq_length"], osk["num_valid_kmers"])
            assert [sk.kmers[i, :21].tobytes() for i in range(len(sk))] == osk["kmers"]


def test_sketch_stream_takes_compressed_bytes(fb, oracle):
    """sketch_stream reads through needletail, which sniffs gzip / bzip2 / xz (lib.rs:58-60): so does fb2_sketch_stream
    for bytes in memory (two gzip members; one bzip2 / xz stream); the handle-level feed calls still refuse them."""
    import bz2, gzip, lzma
    rng = np.random.default_rng(92)
    fq = gen.fastq(rng, n_records=700, max_len=250)
    sp = fb.SketchParams.mash(3000, 200, True, 21, 0)
    fp = fb.FilterParams(False, (None, None), 0.21, 0.1)
    rc, osk = oracle.sketch_stream(fq, oracle.mash_params(3000, 200, True, 21, 0), oracle.make_filter(False, (None, None), 0.21, 0.1))
    assert rc == oracle.OK
    half = len(fq) // 2
    for packed in (gzip.compress(fq[:half]) + gzip.compress(fq[half:]), bz2.compress(fq), lzma.compress(fq)):
        sk = fb.sketch_stream(packed, "reads.fq.z", sp, fp)
        assert np.array_equal(sk.hashes_u64, osk["hashes"]) and np.array_equal(sk.counts, osk["counts"])
        assert (sk.seq_length, sk.num_valid_kmers) == (osk["seq_length"], osk["num_valid_kmers"])
    with pytest.raises(fb.FinchError) as ei:
        fb.sketch_stream(gzip.compress(fq)[:-300], "cut.fq.gz", sp, fp)
    assert ei.value.code == fb.EIO and "truncated gzip" in str(ei.value)
    with sp.create_sketcher() as s, pytest.raises(fb.FinchError) as ei:
        s.feed_fastx(gzip.compress(fq), final=True)
    assert ei.value.code == fb.EUNSUPPORTED


def test_sketch_files_bz2_xz(fb, oracle, tmp_path):
    """bzip2 ("BZ") and xz (fd 37) input, decompressed on the host through the system's libbz2 / liblzma (loaded at
    first use); one stream each, as needletail's BzDecoder / XzDecoder read them; truncated streams are errors."""
    import bz2
    import lzma
    rng = np.random.default_rng(92)
    fa = gen.fasta(rng, n_records=3, max_len=60000, width=70)
    fq = gen.fastq(rng, n_records=900, max_len=250)
    p1, p2, p3, p4 = (tmp_path / n for n in ("a.fa.bz2", "b.fq.xz", "bad.fa.bz2", "bad.fq.xz"))
    p1.write_bytes(bz2.compress(fa))
    p2.write_bytes(lzma.compress(fq))
    p3.write_bytes(bz2.compress(fa)[:-100])
    p4.write_bytes(lzma.compress(fq)[:-100])
    sp = fb.SketchParams.mash(3000, 200, True, 21, 0)
    fp = fb.FilterParams(False, (None, None), 0.21, 0.1)
    sks = fb.sketch_files([str(p1), str(p2)], sp, fp)
    for sk, data in zip(sks, (fa, fq)):
        rc, osk = oracle.sketch_stream(data, oracle.mash_params(3000, 200, True, 21, 0), oracle.make_filter(False, (None, None), 0.21, 0.1))
        assert rc == oracle.OK
        assert np.array_equal(sk.hashes_u64, osk["hashes"]) and np.array_equal(sk.counts, osk["counts"])
        assert (sk.seq_length, sk.num_valid_kmers) == (osk["seq_length"], osk["num_valid_kmers"])
    for bad, word in ((p3, "bzip2"), (p4, "xz")):
        with pytest.raises(fb.FinchError) as ei:
            fb.sketch_files([str(bad)], sp, fp)
        assert word in str(ei.value)


def test_provisional_first_threshold(fb, synth, oracle, monkeypatch):
    """A first chunk too large for one infinite-threshold launch starts from a provisional finite threshold
    (16 * size expected candidates) and is hashed in one launch; if fewer than `size` distinct keys lie
    below it the chunk is redone with the exact ramp.  Smallest log so that small inputs take this path."""
    monkeypatch.setenv("FB2_LOG_M", "0")
    genome = synth.synth_genome(400_000, 31)
    data, nb = synth.synth_fastq(genome, 14_000, 150, 0.01, 33)
    data = data.tobytes()
    for kind, size, scale in (("mash", 2000, 0.0), ("scaled", 500, 0.002)):
        ovec, ototals, _ = oracle_sketch(oracle, data, kind, size, 21, 0, scale or 0.001)
        sp = fb.SketchParams.mash(size, size, False, 21, 0) if kind == "mash" else fb.SketchParams.scaled(size, 21, scale, 0)
        with sp.create_sketcher() as s:
            s.feed_fastx(data, final=True)
            assert s.total_bases_and_kmers() == ototals
            assert_same(s.to_arrays(), ovec, 21)
            st = s.stats()
            assert st["provisional_redos"] == 0 and st["hash_launches"] <= 2, st
        monkeypatch.setenv("FB2_NO_PROVISIONAL", "1")
        with sp.create_sketcher() as s:
            s.feed_fastx(data, final=True)
            assert_same(s.to_arrays(), ovec, 21)
            assert s.stats()["hash_launches"] > 2
        monkeypatch.delenv("FB2_NO_PROVISIONAL")
    # very repetitive input: 30 + 1 distinct k-mers in 1.5 M positions -> the provisional threshold cannot hold
    unit = b"ACGTTGCAAGGCTTAACCGGATATCGCGTA"
    rep = b">rep\n" + unit * 40000 + b"\n>polyA\n" + b"A" * 300000 + b"\n"
    ovec, ototals, _ = oracle_sketch(oracle, rep, "mash", 1000, 21, 0)
    with fb.SketchParams.mash(1000, 1000, True, 21, 0).create_sketcher() as s:
        s.feed_fastx(rep, final=True)
        assert s.total_bases_and_kmers() == ototals
        assert_same(s.to_arrays(), ovec, 21)
        assert s.stats()["provisional_redos"] == 1


def test_provisional_threshold_needs_empty_table(fb, oracle, monkeypatch):
    """Second chunk arrives while the threshold is still infinite and the table already holds keys of the
    first chunk: the provisional threshold (whose fallback clears the table) must not be used then."""
    monkeypatch.setenv("FB2_LOG_M", "0")
    unit = b"ACGTTGCAAGGCTTAACCGGATATCGCGTA"
    piece1 = b">polyA\n" + b"A" * 1_300_000 + b"\n"                   # >= 1 MiB: fed as its own chunk, 1 distinct k-mer
    piece2 = b">rep\n" + unit * 60000 + b"\n>polyC\n" + b"C" * 200000 + b"\n"
    data = piece1 + piece2
    ovec, ototals, _ = oracle_sketch(oracle, data, "mash", 1000, 21, 0)
    gres, gtotals, _ = gpu_sketch(fb, data, "mash", 1000, 21, 0, pieces=[len(piece1)])
    assert gtotals == ototals
    assert_same(gres, ovec, 21)
    assert len(ovec["hashes"]) >= 31


def test_sketch_files_many_workers(fb, oracle, tmp_path, monkeypatch):
    """fb2_sketch_files with more files than worker handles, mixed formats and sizes, twice (the second
    call re-uses the pooled handles): results in input order, identical to one-by-one sketching."""
    rng = np.random.default_rng(77)
    paths, datas = [], []
    for i in range(21):
        if i % 3 == 2:
            d = gen.fastq(rng, n_records=int(rng.integers(20, 400)), max_len=200)
        else:
            d = gen.fasta(rng, n_records=int(rng.integers(1, 4)), max_len=int(rng.integers(2000, 60000)), width=int(rng.integers(40, 90)))
        p = tmp_path / f"w{i}.{'fq' if i % 3 == 2 else 'fa'}"
        p.write_bytes(d)
        paths.append(str(p)); datas.append(d)
    sp = fb.SketchParams.mash(2000, 100, True, 21, 0)
    fp = fb.FilterParams(None, (None, None), 0.21, 0.1)
    want = []
    for d in datas:
        rc, osk = oracle.sketch_stream(d, oracle.mash_params(2000, 100, True, 21, 0), oracle.make_filter(None, (None, None), 0.21, 0.1))
        assert rc == oracle.OK
        want.append(osk)
    for workers in ("4", "4", "1", "16"):
        monkeypatch.setenv("FB2_FILE_WORKERS", workers)
        sks = fb.sketch_files(paths, sp, fp)
        for pth, osk, sk in zip(paths, want, sks):
            assert sk.name == pth
            assert np.array_equal(sk.hashes_u64, osk["hashes"]) and np.array_equal(sk.counts, osk["counts"])
            assert np.array_equal(sk.extra_counts, osk["extras"])
            assert (sk.seq_length, sk.num_valid_kmers) == (osk["seq_length"], osk["num_valid_kmers"])
    # one bad file fails the whole call with its message (lib.rs:29-49 returns the first error)
    with pytest.raises(fb.FinchError) as ei:
        fb.sketch_files(paths[:5] + [str(tmp_path / "missing.fa")] + paths[5:], sp, fp)
    assert "No such file or directory" in str(ei.value)
    fb.lib().fb2_sketch_files_release_pool()


def test_errors(fb):
    sp = fb.SketchParams.mash(10, 10, False, 21, 0)
    with pytest.raises(fb.FinchError) as e:
        fb.sketch_stream(b"ACGT\n", "x", sp, fb.FilterParams())
    assert e.value.code == fb.EFORMAT
    with pytest.raises(fb.FinchError) as e:
        fb.sketch_stream(b"", "x", sp, fb.FilterParams())
    assert e.value.code == fb.EEMPTY
    with pytest.raises(fb.FinchError) as e:
        fb.sketch_stream(b"@r1\nACGT\n+\n", "x", sp, fb.FilterParams())       # truncated record
    assert e.value.code == fb.ERECORD
    with pytest.raises(fb.FinchError) as e:
        fb.sketch_stream(b"@r1\nACGT\n-\nIIII\n", "x", sp, fb.FilterParams())  # bad separator
    assert e.value.code == fb.ERECORD
    with pytest.raises(fb.FinchError) as e:                                      # strict: too few k-mers
        fb.sketch_stream(b">a\nACGTACGTACGTACGTACGTACGTAAAC\n", "few", sp, fb.FilterParams())
    assert e.value.code == fb.ETOOFEW and "few had too few kmers (" in e.value.message
    with pytest.raises(fb.FinchError) as e:
        fb.SketchParams.mash(10, 10, False, 0, 0).create_sketcher()
    assert e.value.code == fb.EINVAL
    with pytest.raises(fb.FinchError) as e:
        fb.sketch_files(["/nonexistent/file.fa"], sp, fb.FilterParams())
    assert e.value.code == fb.EIO and "No such file or directory" in e.value.message  # test_cli.rs:9-18


@pytest.mark.parametrize("k,fmt", [(40, "fasta"), (64, "fastq"), (255, "fasta")])
def test_big_k_multi_chunk(fb, synth, oracle, monkeypatch, k, fmt):
    """k > 32 across chunk and region seams (1 MiB chunks: the 256-symbol halo carries k - 1 symbols), with the
    CLI's filter tail on the FASTQ case."""
    monkeypatch.setenv("FB2_CHUNK_MB", "1")
    if fmt == "fasta":
        data = synth.synth_fasta(2_600_000, n_records=3, line_width=60, lower_frac=0.03, n_frac=0.004, seed=40 + k).tobytes()
    else:
        genome = synth.synth_genome(60_000, 5)
        data = synth.synth_fastq(genome, 9_000, 150, 0.01, 7)[0].tobytes()
    for kind, size, scale in (("mash", 3000, 0.0), ("scaled", 100, 0.01)):
        ovec, ototals, _ = oracle_sketch(oracle, data, kind, size, k, 0, scale or 0.001)
        gres, gtotals, _ = gpu_sketch(fb, data, kind, size, k, 0, scale or 0.001)
        assert gtotals == ototals
        assert_same(gres, ovec, k)


def _fastq_records(rng, n, crlf):
    nl = b"\r\n" if crlf else b"\n"
    recs = []
    for r in range(n):
        m = int(rng.integers(0, 260))
        s = gen.rand_seq(rng, m, 0.01)
        q = bytes(rng.integers(33, 74, size=m).astype(np.uint8))
        recs.append([b"@r%d" % r, s, b"+", q])
    return recs, nl


def _join(recs, nl, final_newline=True):
    data = b"".join(nl.join(r) + nl for r in recs)
    return data if final_newline else data[:-len(nl)]


@pytest.mark.parametrize("st_tiles", ["1", "8"])
@pytest.mark.parametrize("crlf", [False, True])
def test_fastq_seq_qual_length_mismatch(fb, oracle, monkeypatch, st_tiles, crlf):
    """needletail rejects a record whose sequence and quality lengths differ (the reference panics, lib.rs:63);
    the oracle returns E_RECORD.  Records are broken one at a time: at chunk seams (1 MiB chunks), supertile
    seams, batch seams inside a supertile, the first and the last record, with and without a final newline."""
    monkeypatch.setenv("FB2_CHUNK_MB", "1")
    monkeypatch.setenv("FB2_ST_TILES", st_tiles)
    rng = np.random.default_rng(5 + int(st_tiles) + 2 * crlf)
    recs, nl = _fastq_records(rng, 9000, crlf)
    sp = fb.SketchParams.mash(500, 500, True, 21, 0)
    osp = oracle.mash_params(500, 500, True, 21, 0)
    good = _join(recs, nl)
    assert len(good) > (2 << 20)
    rc, osk = oracle.sketch_stream(good, osp, oracle.make_filter(False))
    assert rc == oracle.OK
    sk = fb.sketch_stream(good, "good", sp, fb.FilterParams(False))
    assert np.array_equal(sk.hashes_u64, osk["hashes"]) and np.array_equal(sk.counts, osk["counts"])
    # record offsets, to aim at the seams
    offs = np.cumsum([0] + [sum(len(x) for x in r) + 4 * len(nl) for r in recs])
    targets = {0, 1, len(recs) - 1, len(recs) - 2}
    for seam in (1 << 20, 2 << 20, 4096 * 7, 16384 * 3, 32768 * 5):
        i = int(np.searchsorted(offs, seam))
        targets.update({max(0, i - 2), max(0, i - 1), min(len(recs) - 1, i)})
    targets.update(int(x) for x in rng.integers(0, len(recs), size=6))
    for trial, i in enumerate(sorted(targets)):
        bad = [list(r) for r in recs]
        how = trial % 4
        if how == 0: bad[i][3] = bad[i][3] + b"I"                # quality one longer
        elif how == 1: bad[i][1] = bad[i][1] + b"A"              # sequence one longer
        elif how == 2: bad[i][3] = bad[i][3][:-1] if bad[i][3] else b"II"
        else: bad[i][1] = bad[i][1][:-1] if bad[i][1] else b"AC"
        for final_newline in ((True, False) if i >= len(recs) - 2 else (True,)):
            data = _join(bad, nl, final_newline)
            rc, _ = oracle.sketch_stream(data, osp, oracle.make_filter(False))
            assert rc == oracle.E_RECORD, (i, how)
            with pytest.raises(fb.FinchError) as e:
                fb.sketch_stream(data, "bad", sp, fb.FilterParams(False))
            assert e.value.code == fb.ERECORD, (i, how, final_newline, e.value.message)
    # no false positives without a final newline / with a CR-less last line / trailing blank lines
    for tail in (b"", nl, nl + nl + b"\r\n\n"):
        data = _join(recs, nl, final_newline=False) + tail
        rc, osk = oracle.sketch_stream(data, osp, oracle.make_filter(False))
        assert rc == oracle.OK
        sk = fb.sketch_stream(data, "good", sp, fb.FilterParams(False))
        assert np.array_equal(sk.hashes_u64, osk["hashes"]) and sk.seq_length == osk["seq_length"]


def test_fastq_end_of_input_corner_cases(fb, oracle):
    """What the reader accepts / rejects at the end of a FASTQ stream, against the oracle's verdict: empty last
    quality line, CR-only lines, unterminated lines, trailing blank lines."""
    sp = fb.SketchParams.mash(50, 50, True, 3, 0)
    osp = oracle.mash_params(50, 50, True, 3, 0)
    body = b"@a\nACGTAC\n+\nIIIIII\n"
    cases = [b"@r\n\n+\n", b"@r\n\n+\n\n", b"@r\n\n+\n\n\n\r\n", body + b"@r\n\n+\n\n", body + b"@r\n\n+\n",
             b"@r\nA\n+\n", b"@r\nA\n+\n\n", b"@r\nA\n+\n\r\r\n", b"@r\nA\n+\n\r\r", b"@r\nA\n+\n\r\n", b"@r\nAC\n+\n\r\r\n",
             b"@r\n\r\n+\r\n\r\n", b"@r\r\n\r\n+\r\n", body + b"@r\nACG\n+\nII", body + b"@r\nACG\n+\nIII", body + b"@r\nACG\n+\nIII\r",
             body + b"@r\nACG\n+\nIIII\r", body + b"@r\nACG\n+", body + b"@r\nACG\n+\r", body + b"@r\nACG\n", body + b"\n\n\n",
             body + b"@r\nACG\r\n+\r\nIII", body + b"@r\nACG\r\n+\r\nIIII"]
    for data in cases:
        rc, osk = oracle.sketch_stream(data, osp, oracle.make_filter(False))
        if rc == oracle.OK:
            sk = fb.sketch_stream(data, "x", sp, fb.FilterParams(False))
            assert np.array_equal(sk.hashes_u64, osk["hashes"]) and np.array_equal(sk.counts, osk["counts"]), data
            assert (sk.seq_length, sk.num_valid_kmers) == (osk["seq_length"], osk["num_valid_kmers"]), data
        else:
            assert rc == oracle.E_RECORD, data
            with pytest.raises(fb.FinchError) as e:
                fb.sketch_stream(data, "x", sp, fb.FilterParams(False))
            assert e.value.code == fb.ERECORD, (data, e.value.message)


@pytest.mark.parametrize("fmt,kind,size,k,scale", [("fastq", "mash", 20000, 21, 0.0), ("fasta", "mash", 1000, 21, 0.0),
                                                   ("fasta", "scaled", 100, 31, 0.01), ("fasta", "mash", 500, 64, 0.0),
                                                   ("fastq_crlf", "scaled", 0, 21, 0.02)])
def test_one_stream_split_over_gpus(fb, synth, oracle, monkeypatch, fmt, kind, size, k, scale):
    """fb2_sketch_stream_multi: one file cut into byte ranges, one range per GPU, tables united exactly over peer
    memory (SURVEY 8e second mode).  On a one-GPU box the ranges share the device (FB2_MULTI_OVERSUBSCRIBE): the
    cuts, the carried parser state / halo symbols, the position-id bases and the merge are the same code."""
    monkeypatch.setenv("FB2_MULTI_OVERSUBSCRIBE", "1")
    monkeypatch.setenv("FB2_MIN_RANGE_KB", "64")
    monkeypatch.setenv("FB2_CHUNK_MB", "1")
    rng = np.random.default_rng(len(fmt) + size + k)
    if fmt == "fasta":
        data = gen.fasta(rng, n_records=4, min_len=100_000, max_len=400_000, width=61, messy=0.004, crlf=bool(size % 2 == 0 and k == 31))
        data += synth.synth_fasta(700_000, n_records=2, line_width=80, lower_frac=0.05, n_frac=0.01, seed=9).tobytes()
    else:
        genome = synth.synth_genome(30_000, 4)
        data = synth.synth_fastq(genome, 8_000, 150, 0.01, 11)[0].tobytes()
        if fmt == "fastq_crlf":
            data = data.replace(b"\n", b"\r\n")
        data += gen.fastq(rng, n_records=300, max_len=300, messy=0.01, crlf=fmt == "fastq_crlf", final_newline=False)
    sp = fb.SketchParams.mash(size, min(size, 1000), True, k, 0) if kind == "mash" else fb.SketchParams.scaled(size, k, scale, 0)
    osp = oracle.mash_params(size, min(size, 1000), True, k, 0) if kind == "mash" else oracle.scaled_params(size, k, scale, 0)
    for filt in ((True, 0.21, 0.1), (None, 0.21, 0.1)):
        fp = fb.FilterParams(filt[0], (None, None), filt[1], filt[2])
        rc, osk = oracle.sketch_stream(data, osp, oracle.make_filter(filt[0], (None, None), filt[1], filt[2]))
        assert rc == oracle.OK
        for ngpus in (2, 3, 5):
            sk = fb.sketch_stream_multi(data, "split", sp, fp, ngpus)
            assert np.array_equal(sk.hashes_u64, osk["hashes"]), (ngpus, filt)
            assert np.array_equal(sk.counts, osk["counts"]) and np.array_equal(sk.extra_counts, osk["extras"])
            assert [sk.kmer_bytes(i) for i in range(len(sk))] == osk["kmers"]
            assert (sk.seq_length, sk.num_valid_kmers) == (osk["seq_length"], osk["num_valid_kmers"])


def test_library_leaves_the_callers_device_alone(fb, synth, monkeypatch):
    """Multi-GPU calls work on several devices from several threads; the calling thread's current CUDA device must
    be what it was (a caller mixing this library with torch would otherwise allocate on the wrong GPU)."""
    import ctypes as C
    rt = C.CDLL("libcudart.so.12")
    ndev = fb.lib().fb2_device_count()
    monkeypatch.setenv("FB2_MULTI_OVERSUBSCRIBE", "1")
    monkeypatch.setenv("FB2_MIN_RANGE_KB", "64")
    data = synth.synth_fasta(900_000, n_records=2, line_width=70, seed=3).tobytes()
    sp = fb.SketchParams.mash(500, 500, True, 21, 0)
    for want in range(min(ndev, 2)):
        assert rt.cudaSetDevice(want) == 0
        fb.sketch_stream_multi(data, "x", sp, fb.FilterParams(False), 3)
        fb.dist_all_pairs_cut(np.sort(np.random.default_rng(1).integers(0, 2**60, size=(40, 50), dtype=np.uint64), axis=1),
                              np.full(40, 50, np.uint32), 21, 0.5, ngpus=0)
        with fb.SketchParams.mash(10, 10, True, 5, 0, device=ndev - 1).create_sketcher() as s:
            s.feed_fastx(b">a\nACGTACGTAC\n", final=True)
            s.to_arrays()
        cur = C.c_int(-1)
        assert rt.cudaGetDevice(C.byref(cur)) == 0 and cur.value == want
    rt.cudaSetDevice(0)


def test_split_stream_reports_record_errors(fb, oracle, monkeypatch):
    monkeypatch.setenv("FB2_MULTI_OVERSUBSCRIBE", "1")
    monkeypatch.setenv("FB2_MIN_RANGE_KB", "64")
    rng = np.random.default_rng(4)
    recs, nl = _fastq_records(rng, 3000, False)
    sp = fb.SketchParams.mash(500, 500, True, 21, 0)
    for i in (5, 1500, 2995):
        bad = [list(r) for r in recs]
        bad[i][3] = bad[i][3] + b"I"
        with pytest.raises(fb.FinchError) as e:
            fb.sketch_stream_multi(_join(bad, nl), "bad", sp, fb.FilterParams(False), 4)
        assert e.value.code == fb.ERECORD
    # a quality line that looks like a header ('@' first) right where a cut would go must not fool the split
    tricky = [list(r) for r in recs]
    for r in tricky:
        if len(r[3]) > 2:
            r[3] = b"@" + r[3][1:]
    data = _join(tricky, nl)
    rc, osk = oracle.sketch_stream(data, oracle.mash_params(500, 500, True, 21, 0), oracle.make_filter(False))
    sk = fb.sketch_stream_multi(data, "tricky", sp, fb.FilterParams(False), 4)
    assert rc == oracle.OK and np.array_equal(sk.hashes_u64, osk["hashes"]) and np.array_equal(sk.counts, osk["counts"])


def test_sketch_files_multi_gpu_assignment(fb, oracle, tmp_path, monkeypatch):
    """fb2_sketch_files_multi: LPT over the devices, results in input order."""
    monkeypatch.setenv("FB2_MULTI_OVERSUBSCRIBE", "1")
    rng = np.random.default_rng(8)
    paths, datas = [], []
    for i in range(11):
        d = gen.fasta(rng, n_records=2, max_len=int(rng.integers(500, 30000)), width=80) if i % 3 else gen.fastq(rng, n_records=int(rng.integers(5, 200)))
        p = tmp_path / f"m{i}.fx"
        p.write_bytes(d)
        paths.append(str(p)); datas.append(d)
    sp = fb.SketchParams.mash(300, 30, True, 21, 0)
    fp = fb.FilterParams(None, (None, None), 0.21, 0.1)
    for ngpus in (1, 3):
        sks = fb.sketch_files(paths, sp, fp, ngpus=ngpus)
        for p, d, sk in zip(paths, datas, sks):
            rc, osk = oracle.sketch_stream(d, oracle.mash_params(300, 30, True, 21, 0), oracle.make_filter(None, (None, None), 0.21, 0.1))
            assert rc == oracle.OK and sk.name == p
            assert np.array_equal(sk.hashes_u64, osk["hashes"]) and np.array_equal(sk.counts, osk["counts"])
            assert (sk.seq_length, sk.num_valid_kmers) == (osk["seq_length"], osk["num_valid_kmers"])


def test_host_strip_matches_device_parse(fb, synth, oracle, monkeypatch):
    """FB2_HOST_STRIP=1: FASTQ record framing on the host (strip.cpp), sequence lines only over PCIe.  Same sketches,
    same totals, same verdict on malformed input as the device parse -- i.e. as the oracle."""
    monkeypatch.setenv("FB2_HOST_STRIP", "1")
    monkeypatch.setenv("FB2_STRIP_THREADS", "5")
    monkeypatch.setenv("FB2_CHUNK_MB", "1")
    sp = fb.SketchParams.mash(2000, 2000, True, 21, 0)
    osp = oracle.mash_params(2000, 2000, True, 21, 0)
    rng = np.random.default_rng(12)
    genome = synth.synth_genome(40_000, 6)
    big = synth.synth_fastq(genome, 30_000, 150, 0.01, 8)[0].tobytes()          # 9 MB: parallel blocks, several chunks
    inputs = [big, big.replace(b"\n", b"\r\n"), big[:-1],
              gen.fastq(rng, n_records=400, max_len=300, messy=0.03, final_newline=False),
              gen.fastq(rng, n_records=300, max_len=120, messy=0.05, crlf=True),
              gen.fastq(rng, n_records=50, max_len=50_000, messy=0.001),          # long reads: records larger than a range
              b"@r\n\n+\n", b"@r\nACGTTGCA\n+\nIIIIIIII", big + b"\n\n\r\n"]
    for data in inputs:
        rc, osk = oracle.sketch_stream(data, osp, oracle.make_filter(False))
        assert rc == oracle.OK
        for pieces in (None, [3, 1, 700_001, 5], [1 << 20, 1 << 20]):
            with sp.create_sketcher() as s:
                if pieces is None:
                    s.feed_fastx(data, final=True)
                else:
                    pos = 0
                    for pc in pieces:
                        s.feed_fastx(data[pos:pos + pc], final=False)
                        pos += pc
                    s.feed_fastx(data[pos:], final=True)
                sk = s.sketch("x", fb.FilterParams(False))
                assert s.stats()["h2d_bytes"] < 0.62 * len(data) + 4096          # the quality / header lines stayed on the host
            assert np.array_equal(sk.hashes_u64, osk["hashes"]) and np.array_equal(sk.counts, osk["counts"])
            assert np.array_equal(sk.extra_counts, osk["extras"]) and [sk.kmer_bytes(i) for i in range(len(sk))] == osk["kmers"]
            assert (sk.seq_length, sk.num_valid_kmers) == (osk["seq_length"], osk["num_valid_kmers"])
    # malformed input: the verdicts of the reader
    recs, nl = _fastq_records(rng, 4000, False)
    bad_inputs = []
    for i, how in ((0, 0), (1999, 1), (3999, 2), (2500, 3)):
        bad = [list(r) for r in recs]
        if how == 0: bad[i][3] += b"I"
        elif how == 1: bad[i][1] += b"A"
        elif how == 2: bad[i][2] = b"-"
        else: bad[i][0] = b"r" + bad[i][0][1:]
        bad_inputs.append(_join(bad, nl))
    good = _join(recs, nl)
    bad_inputs += [good + b"@r\nACG\n+\n", good + b"@r\nACG\n+\nII", good[:len(good) // 2] + b"\n\n\n\n" + good[len(good) // 2:],
                   good + b"@r\nACG\n", b"@r\nA\n+\n"]
    for data in bad_inputs:
        rc, _ = oracle.sketch_stream(data, osp, oracle.make_filter(False))
        assert rc == oracle.E_RECORD
        with pytest.raises(fb.FinchError) as e:
            fb.sketch_stream(data, "bad", sp, fb.FilterParams(False))
        assert e.value.code == fb.ERECORD, e.value.message


def test_sketch_files_large_file_is_read_in_parallel(fb, synth, tmp_path):
    """Files of 64 MiB and more are read by several threads (pread slices of 32 MiB pieces, double-buffered against
    the copies to the GPU): same sketch as the bytes handed over in one piece."""
    genome = synth.synth_genome(200_000, 3)
    data = synth.synth_fastq(genome, 240_000, 150, 0.005, 9)[0]
    assert data.size > (70 << 20)
    p = tmp_path / "big.fq"
    data.tofile(p)
    sp = fb.SketchParams.from_cli("mash", n_hashes=500, kmer_length=21, filters_enabled=True)
    fp = fb.FilterParams(True, (None, None), 0.21, 0.1)
    a = fb.sketch_files([str(p)], sp, fp)[0]
    b = fb.sketch_stream(data, str(p), sp, fp)
    assert np.array_equal(a.hashes_u64, b.hashes_u64) and np.array_equal(a.counts, b.counts) and np.array_equal(a.extra_counts, b.extra_counts)
    assert (a.seq_length, a.num_valid_kmers) == (b.seq_length, b.num_valid_kmers) == (240_000 * 150, 240_000 * 130)


def test_sketch_files(fb, oracle, tmp_path):
    rng = np.random.default_rng(21)
    paths, datas = [], []
    for i in range(5):
        d = gen.fasta(rng, n_records=2, max_len=4000, width=80) if i % 2 == 0 else gen.fastq(rng, n_records=50)
        p = tmp_path / f"in{i}.{'fa' if i % 2 == 0 else 'fq'}"
        p.write_bytes(d)
        paths.append(str(p)); datas.append(d)
    sp = fb.SketchParams.mash(400, 20, True, 21, 0)
    fp = fb.FilterParams(None, (None, None), 0.21, 0.1)
    sks = fb.sketch_files(paths, sp, fp)
    for p, d, sk in zip(paths, datas, sks):
        rc, osk = oracle.sketch_stream(d, oracle.mash_params(400, 20, True, 21, 0),
                                       oracle.make_filter(None, (None, None), 0.21, 0.1))
        assert rc == oracle.OK and sk.name == p
        assert np.array_equal(sk.hashes_u64, osk["hashes"]) and np.array_equal(sk.counts, osk["counts"])
        assert (sk.seq_length, sk.num_valid_kmers) == (osk["seq_length"], osk["num_valid_kmers"])
        assert sk.filter_params.filter_on == osk["filter_on"]


# ---- dist -----------------------------------------------------------------------------------------
def test_raw_distance_kats(fb):  # distance.rs:188-242
    rd = fb.raw_distance
    assert rd([0, 1, 2], [1, 2]) == (2. / 2., 2. / 3., 2, 3)
    assert rd([0, 2], [1, 2]) == (1. / 2., 1. / 3., 1, 3)
    assert rd([0, 1], [2, 3]) == (0., 0., 0, 2)
    assert rd([], []) == (0., 1., 0, 0)
    assert rd([], [5]) == (0., 1., 0, 0)
    assert rd([10, 15, 20], [15, 20], 1e-18) == (1., 2. / 3., 2, 3)
    assert rd([5, 10, 15], [5, 10], 1e-18) == (1., 2. / 3., 2, 3)
    assert rd([5, 10, 15, 20], [5, 10], 1e-18) == (1., 2. / 3., 2, 3)
    assert rd([5, 10], [5, 10, 15, 20], 1e-18) == (2. / 3., 2. / 3., 2, 3)


def test_dist_batch_random(fb, oracle):
    rng = np.random.default_rng(8)
    pool = np.unique(rng.integers(0, 2**63, size=4000, dtype=np.uint64))
    sk = [np.sort(rng.choice(pool, size=int(rng.integers(0, 1000)), replace=False)) for _ in range(24)]
    sk.append(np.array([2**64 - 1], np.uint64)); sk.append(np.zeros(0, np.uint64))
    n = len(sk)
    q, r = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    for scale in (0.0, 0.3):
        out = fb.dist_batch(sk, q.ravel(), r.ravel(), scale)
        mat, lens, stride = fb._pack(sk)
        allp = fb.dist_all_pairs(mat, lens, scale)
        for p, (a, b) in enumerate(zip(q.ravel(), r.ravel())):
            cont, jac, com, tot = oracle.raw_distance(sk[a], sk[b], scale)
            c, i, j = (int(v) for v in out[p])
            assert (c, i - c + j) == (com, tot), (a, b, scale)
            assert fb._finish_pair(out[p], 21)[:2] == (cont, jac)
            assert tuple(allp[a, b]) == (c, i, j)


def test_raw_distance_commutes(fb, oracle):
    """distance.rs:176-185 (proptest): raw_distance(a, b, 0.) == raw_distance(b, a, 0.) for arbitrary sorted lists
    (strictly ascending here: what a sketch holds), small values and u64::MAX included; every case also against the
    oracle's literal loop."""
    from hypothesis import given, settings, strategies as st, HealthCheck
    u64 = st.one_of(st.integers(0, 40), st.integers(0, 2**64 - 1), st.sampled_from([0, 1, 2**63, 2**64 - 2, 2**64 - 1]))
    lists = st.lists(u64, max_size=60, unique=True).map(sorted)

    @settings(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck))
    @given(st.lists(st.tuples(lists, lists), min_size=1, max_size=8))
    def run(pairs):
        sk = [np.array(x, np.uint64) for ab in pairs for x in ab]
        q = np.arange(0, len(sk), 2); r = q + 1
        fwd = fb.dist_batch(sk, q, r, 0.0)
        rev = fb.dist_batch(sk, r, q, 0.0)
        for t, (a, b) in enumerate(pairs):
            lhs, rhs = fb._finish_pair(fwd[t], 21), fb._finish_pair(rev[t], 21)
            assert (lhs[1], lhs[3], lhs[4]) == (rhs[1], rhs[3], rhs[4])       # jaccard, common, total commute
            cont, jac, com, tot = oracle.raw_distance(sk[2 * t], sk[2 * t + 1], 0.0)
            assert (lhs[0], lhs[1], lhs[3], lhs[4]) == (cont, jac, com, tot)
            cont, jac, com, tot = oracle.raw_distance(sk[2 * t + 1], sk[2 * t], 0.0)
            assert (rhs[0], rhs[1], rhs[3], rhs[4]) == (cont, jac, com, tot)
    run()


def _closed_form_pairs(sk, scale=0.0):
    """(common, i, j) per ordered pair from the closed form of the merge loop, numpy."""
    n = len(sk)
    out = np.zeros((n, n, 3), np.uint32)
    mh = None
    if scale > 0.0:
        mh = (2**64 - 1) // int(1.0 / scale)
    for a in range(n):
        A = sk[a]
        for b in range(n):
            B = sk[b]
            c = i = j = 0
            if len(A) and len(B):
                c = len(np.intersect1d(A, B, assume_unique=True))
                t = min(int(A[-1]), int(B[-1]))
                i = int(np.searchsorted(A, np.uint64(t), side="right"))
                j = int(np.searchsorted(B, np.uint64(t), side="right"))
            if mh is not None:
                i = max(i, int(np.searchsorted(A, np.uint64(mh), side="left")))
                j = max(j, int(np.searchsorted(B, np.uint64(mh), side="left")))
            out[a, b] = (c, i, j)
    return out


def test_dist_all_pairs_tiled_edge_cases(fb, oracle):
    """Tiled shared-memory kernel: colliding table slots, lengths up to its limit (1023), longer
    reference rows, query sub-ranges, more queries than one tile, and the warp-per-pair fallback."""
    rng = np.random.default_rng(21)
    pool = np.unique(rng.integers(0, 2**63, size=6000, dtype=np.uint64))
    sk = []
    for n in (1023, 1000, 1, 0, 517, 1023, 999, 1000, 64, 1000, 1000, 333):     # 12 queries -> 2 tiles
        sk.append(np.sort(rng.choice(pool, size=n, replace=False)))
    # identical low 13 bits (every key lands in the same table slot) and identical fingerprints
    sk.append(np.sort((rng.choice(1 << 20, size=300, replace=False).astype(np.uint64) << np.uint64(19)) | np.uint64(0x155)))
    sk.append(sk[-1][::2].copy())
    sk.append(np.arange(1, 1001, dtype=np.uint64))                                # consecutive small integers
    sk.append(np.array([0, 2**64 - 1], np.uint64))
    n = len(sk)
    mat, lens, stride = fb._pack(sk)
    for scale in (0.0, 0.3):
        want = _closed_form_pairs(sk, scale)
        got = fb.dist_all_pairs(mat, lens, scale)
        assert np.array_equal(got, want), np.argwhere((got != want).any(axis=2))[:5]
        sub = fb.dist_all_pairs(mat, lens, scale, 3, 14)
        assert np.array_equal(sub, want[3:14])
    # spot-check the closed form itself against the oracle's literal merge loop
    for a, b in ((0, 1), (12, 13), (13, 12), (14, 2), (15, 0), (3, 5)):
        cont, jac, com, tot = oracle.raw_distance(sk[a], sk[b], 0.0)
        c, i, j = (int(v) for v in _closed_form_pairs([sk[a], sk[b]])[0, 1])
        assert (c, i - c + j) == (com, tot)
    # a query longer than the tiled kernel's limit: whole call falls back to the warp-per-pair kernel
    sk2 = sk[:6] + [np.sort(rng.choice(pool, size=1500, replace=False)), np.sort(rng.choice(pool, size=1024, replace=False))]
    mat2, lens2, _ = fb._pack(sk2)
    want2 = _closed_form_pairs(sk2)
    assert np.array_equal(fb.dist_all_pairs(mat2, lens2, 0.0), want2)
    # long rows as references only (queries 0..5 are short): tiled kernel with nb > 1023
    assert np.array_equal(fb.dist_all_pairs(mat2, lens2, 0.0, 0, 6), want2[:6])


def test_dist_all_pairs_multi_slab(fb):
    """More pairs than one output slab (2^22): the kernel / D2H / host-copy pipeline keeps row order."""
    rng = np.random.default_rng(22)
    n, m = 2304, 64
    base = np.sort(rng.choice(1 << 40, size=(n // 16, m), replace=False).astype(np.uint64), axis=1)
    mat = np.repeat(base, 16, axis=0)                       # 16 identical sketches per cluster
    noise = rng.integers(0, 1 << 40, size=(n, 8), dtype=np.uint64)
    mat[:, -8:] = noise + np.uint64(1 << 41)                # 8 private hashes above the shared ones
    mat.sort(axis=1)
    lens = np.full(n, m, np.uint32)
    got = fb.dist_all_pairs(mat, lens, 0.0)
    assert got.shape == (n, n, 3)
    idx = rng.integers(0, n, size=(400, 2))
    for a, b in idx:
        A, B = mat[a], mat[b]
        c = len(np.intersect1d(A, B))
        t = min(int(A[-1]), int(B[-1]))
        assert tuple(got[a, b]) == (c, int(np.searchsorted(A, np.uint64(t), side="right")),
                                    int(np.searchsorted(B, np.uint64(t), side="right"))), (a, b)
    # a pinned result array takes the device copies directly: same bits
    import torch
    pin = torch.empty((n * n, 3), dtype=torch.int32, pin_memory=True)
    got_p = fb.dist_all_pairs(mat, lens, 0.0, out=pin.numpy().view(np.uint32))
    assert np.array_equal(got_p, got)
    # diagonal: a sketch against itself
    assert np.array_equal(got[np.arange(n), np.arange(n)], np.tile(np.array([m, m, m], np.uint32), (n, 1)))


@pytest.mark.parametrize("scale", [0.0, 0.002])
def test_dist_all_pairs_cut(fb, synth, oracle, monkeypatch, scale):
    """fb2_dist_all_pairs_cut == the dense all-pairs result filtered with main.rs:328's exact test, in (q, r) order;
    small launches / tiny hit buffers force the retry paths; q sub-ranges; the > 1023-hash fallback."""
    n = 600
    mat = synth.synth_sketches(n, 1000, 6, 5)
    lens = np.full(n, 1000, np.uint32)
    lens[7] = 0; lens[11] = 3; lens[500] = 999
    if scale:
        mat = mat >> np.uint64(8)        # about half of every row below u64::MAX * scale: the scaled tail matters
        assert np.all(np.diff(mat.astype(np.int64), axis=1) > 0)
    dense = fb.dist_all_pairs(mat, lens, scale)
    k = 21
    for max_d, rows, (q0, q1) in ((0.05, "27", (0, n)), (0.3, "9", (100, 350)), (1.0, "45", (0, 40)), (0.0, "", (0, n))):
        if rows:
            monkeypatch.setenv("FB2_DIST_ROWS", rows)
        else:
            monkeypatch.delenv("FB2_DIST_ROWS", raising=False)
        hits = fb.dist_all_pairs_cut(mat, lens, k, max_d, scale, q0, q1, skip_self=True, cap=64)
        key = hits["q"].astype(np.int64) * n + hits["r"]
        assert np.all(np.diff(key) > 0)                                   # ascending by (q, r), no duplicates
        got = {}
        cont, jac, md, com, tot = fb.distance_of_hits(hits, k)
        for t in range(len(hits)):
            assert tuple(dense[hits["q"][t], hits["r"][t]]) == (hits["common"][t], hits["i"][t], hits["j"][t])
            if md[t] <= max_d:
                got[(int(hits["q"][t]), int(hits["r"][t]))] = md[t]
        want = {}
        for q in range(q0, q1):
            for r in range(n):
                if q == r:
                    continue
                m = fb._finish_pair(dense[q, r], k)[2]
                if m <= max_d:
                    want[(q, r)] = m
        assert got == want, (max_d, len(got), len(want))
    # sketches longer than the tiled kernel takes: dense fallback with the same cut
    big = np.sort(np.random.default_rng(1).integers(0, 2**62, size=(12, 1500), dtype=np.uint64), axis=1)
    big[5, :700] = big[4, :700]; big[5] = np.sort(big[5])
    bl = np.full(12, 1500, np.uint32)
    hits = fb.dist_all_pairs_cut(big, bl, 21, 0.1, 0.0)
    d2 = fb.dist_all_pairs(big, bl, 0.0)
    want = {(q, r) for q in range(12) for r in range(12) if q != r and fb._finish_pair(d2[q, r], 21)[2] <= 0.1}
    md = fb.distance_of_hits(hits, 21)[2]
    assert {(int(h["q"]), int(h["r"])) for h, m in zip(hits, md) if m <= 0.1} == want and len(want) >= 2


@pytest.mark.parametrize("scale", [0.0, 0.5])
@pytest.mark.parametrize("cb", ["", "64", "250"])
def test_dist_cut_through_inverted_index(fb, synth, monkeypatch, scale, cb):
    """The cut's second engine (sorted (hash, sketch) postings, 16-bit counters per query block) returns exactly what the
    tile kernel returns: same hits, same (common, i, j), same order; column blocks smaller than the collection, ragged
    lengths (incl. > 1023 hashes, which the tile kernel cannot take), row sub-ranges, tiny hit buffers, skip_self off."""
    n = 500
    mat = synth.synth_sketches(n, 1000, 6, 5)
    lens = np.full(n, 1000, np.uint32)
    lens[11] = 3; lens[12] = 1; lens[400] = 999
    if scale:
        mat = mat >> np.uint64(8)
        assert np.all(np.diff(mat.astype(np.int64), axis=1) > 0)
    if cb:
        monkeypatch.setenv("FB2_DIST_CB", cb)
    dense = fb.dist_all_pairs(mat, lens, scale)
    k = 21
    for max_d, rows, (q0, q1), skip in ((0.05, "27", (0, n), True), (0.3, "9", (100, 350), True), (0.6, "", (0, n), False)):
        if rows:
            monkeypatch.setenv("FB2_DIST_ROWS", rows)
        else:
            monkeypatch.delenv("FB2_DIST_ROWS", raising=False)
        monkeypatch.setenv("FB2_DIST_INVERTED", "0")
        want = fb.dist_all_pairs_cut(mat, lens, k, max_d, scale, q0, q1, skip_self=skip, cap=64)
        monkeypatch.setenv("FB2_DIST_INVERTED", "1")
        got = fb.dist_all_pairs_cut(mat, lens, k, max_d, scale, q0, q1, skip_self=skip, cap=64)
        assert len(got) == len(want) and len(want) > 0
        assert np.array_equal(got, want)
        for t in range(0, len(got), 7):
            assert tuple(dense[got["q"][t], got["r"][t]]) == (got["common"][t], got["i"][t], got["j"][t])
        if not skip:
            assert np.any(got["q"] == got["r"])
    # an empty sketch has jaccard 1 with everything (distance.rs:119-123): the index is not applicable, the answer the same
    lens2 = lens.copy(); lens2[7] = 0
    monkeypatch.setenv("FB2_DIST_INVERTED", "0")
    want = fb.dist_all_pairs_cut(mat, lens2, k, 0.05, scale)
    monkeypatch.setenv("FB2_DIST_INVERTED", "1")
    assert np.array_equal(fb.dist_all_pairs_cut(mat, lens2, k, 0.05, scale), want)
    # sketches longer than the tile kernel takes: the index handles them (the other engine falls back to dense slabs)
    big = np.sort(np.random.default_rng(1).integers(0, 2**62, size=(12, 1500), dtype=np.uint64), axis=1)
    big[5, :700] = big[4, :700]; big[5] = np.sort(big[5])
    bl = np.full(12, 1500, np.uint32)
    monkeypatch.setenv("FB2_DIST_INVERTED", "0")
    want = fb.dist_all_pairs_cut(big, bl, 21, 0.1, 0.0)
    monkeypatch.setenv("FB2_DIST_INVERTED", "1")
    got = fb.dist_all_pairs_cut(big, bl, 21, 0.1, 0.0)
    assert np.array_equal(got, want) and len(want) >= 2


def test_distance_scaled_end_to_end(fb):  # distance.rs:312-337
    def mk():
        q = fb.ScaledSketcher(3, 0.001, 2, 42)
        q.push(b"ca", 0); q.push(b"cc", 1); q.push(b"ac", 0); q.push(b"ac", 1)
        return q.to_sketch()
    d = fb.distance(mk(), mk(), False)
    assert (d.jaccard, d.containment, d.common_hashes) == (1.0, 1.0, 3)


def test_distance_old_mode(fb, oracle):
    """distance(.., old_mode=True) = old_distance (distance.rs:136-157) + the mash distance of distance.rs:35-41."""
    rng = np.random.default_rng(41)
    g = rand = gen.rand_seq(rng, 20000)
    mut = bytearray(rand)
    for pos in rng.choice(len(mut), size=400, replace=False):
        mut[pos] = b"ACGT"[int(rng.integers(0, 4))]
    sp = fb.SketchParams.mash(300, 300, True, 21, 0)
    fp = fb.FilterParams(False, (None, None), 0.21, 0.1)
    a = fb.sketch_stream(b">a\n" + g + b"\n", "a", sp, fp)
    b = fb.sketch_stream(b">b\n" + bytes(mut) + b"\n", "b", sp, fp)
    d = fb.distance(a, b, True)
    cont, jac, com, tot = oracle.old_distance(a.hashes_u64, b.hashes_u64)
    assert (d.containment, d.jaccard, d.common_hashes, d.total_hashes) == (cont, jac, com, tot)
    assert d.mash_distance == oracle.mash_distance(jac, 21)
    assert 0 < com < 300 and tot == 300


def test_symbol_regions_match_normalize(fb, oracle):
    """pack_kernel output == per-record BREAK + normalize(false) (mash.rs:73), region by region."""
    rng = np.random.default_rng(99)
    lut = np.full(256, 4, np.uint8)
    for ch, v in ((b"A", 0), (b"C", 1), (b"G", 2), (b"T", 3)):
        lut[ch[0]] = v
    for fmt in ("fasta", "fastq"):
        data = (gen.fasta(rng, n_records=9, max_len=9000, width=61, messy=0.02, crlf=True)
                if fmt == "fasta" else gen.fastq(rng, n_records=200, max_len=300, messy=0.02))
        rc, f, recs = oracle.parse_fastx(data)
        exp = []
        for r in recs:
            norm = lut[np.frombuffer(oracle.normalize(r), np.uint8)]
            exp += ([np.array([4], np.uint8), norm] if fmt == "fasta" else [norm, np.array([4], np.uint8)])
        exp = np.concatenate(exp)
        with fb.SketchParams.mash(10, 10, False, 3, 0).create_sketcher() as s:
            s.feed_fastx(data, final=True)
            geom, counts, regs, buf = s.debug_symbols()
        assert geom["n_st"] > 1
        assert np.array_equal(np.concatenate(regs), exp)


@pytest.mark.parametrize("kind,size,k,scale", [("scaled", 1000, 21, 0.05), ("scaled", 0, 31, 0.2),
                                               ("mash", 5_000_000, 21, 0)])
def test_table_growth(fb, synth, oracle, kind, size, k, scale):
    """Sketches far larger than the initial table (Scaled keeps ~scale * distinct k-mers; a Mash heap larger
    than the input keeps everything): the table must grow and stay exact."""
    data = synth.synth_fasta(2_000_000, n_records=2, line_width=70, lower_frac=0.05, n_frac=0.002, seed=11).tobytes()
    ovec, ototals, _ = oracle_sketch(oracle, data, kind, size, k, 0, scale or 0.001)
    gres, gtotals, _ = gpu_sketch(fb, data, kind, size, k, 0, scale or 0.001)
    assert gtotals == ototals
    assert len(ovec["hashes"]) > 90_000
    assert_same(gres, ovec, k)


def test_repetitive_input(fb, oracle):
    """Low-complexity input (few distinct k-mers, huge counts): the threshold never becomes finite."""
    unit = b"ACGTTGCAAGGCTTAACCGGATATCGCGTA"
    data = b">rep\n" + unit * 40000 + b"\n>polyA\n" + b"A" * 300000 + b"\n"
    ovec, ototals, _ = oracle_sketch(oracle, data, "mash", 1000, 21, 0)
    gres, gtotals, _ = gpu_sketch(fb, data, "mash", 1000, 21, 0)
    assert gtotals == ototals
    assert_same(gres, ovec, 21)
    assert int(ovec["counts"].max()) >= 40000 - 1


# ---- fused single-pass parse: the workers' guess of a supertile's start state ---------------------
@pytest.mark.parametrize("st_tiles", ["1", "8"])
def test_fused_parse_wrong_fastq_guess_is_redone(fb, oracle, monkeypatch, st_tiles):
    """parse_fused_kernel guesses the line phase at the start of a supertile from the first '@' line whose second next
    line starts with '+', and checks the guess against the look-back.  Here every quality line starts with '@' and
    every sequence line with '+' (a kept non-base), so the first candidate is a QUALITY line for about half of the
    supertiles: the guess is wrong there and the supertile must be redone."""
    monkeypatch.setenv("FB2_CHUNK_MB", "1")
    monkeypatch.setenv("FB2_ST_TILES", st_tiles)
    rng = np.random.default_rng(77 + int(st_tiles))
    recs = []
    for r in range(14000):
        m = int(rng.integers(30, 120))
        s = b"+" + gen.rand_seq(rng, m, 0.0)
        q = b"@" + bytes(rng.integers(33, 74, size=m).astype(np.uint8))
        recs.append(b"@r%d\n" % r + s + b"\n+\n" + q + b"\n")
    data = b"".join(recs)
    assert len(data) > (2 << 20)
    want, totals, _ = oracle_sketch(oracle, data, "mash", 5000, 21, 0)
    got = gpu_sketch(fb, data, "mash", 5000, 21, 0)
    assert_same(got[0], want, 21)
    assert got[1] == totals


@pytest.mark.parametrize("st_tiles", ["1", "8"])
def test_fused_parse_header_lines_across_supertiles(fb, oracle, monkeypatch, st_tiles):
    """FASTA: the guess is 'the supertile does not start inside a header line'; header lines of several KiB make it
    wrong (and some supertiles hold no line start at all: identity in the look-back)."""
    monkeypatch.setenv("FB2_CHUNK_MB", "1")
    monkeypatch.setenv("FB2_ST_TILES", st_tiles)
    rng = np.random.default_rng(91 + int(st_tiles))
    parts = []
    for r in range(60):
        hdr = b">" + bytes(rng.integers(65, 91, size=int(rng.integers(10, 70000))).astype(np.uint8))   # ACGT in headers too
        seq = gen.rand_seq(rng, int(rng.integers(100, 90000)), 0.001)
        width = int(rng.integers(40, 50000))
        lines = [seq[i:i + width] for i in range(0, len(seq), width)]
        parts.append(hdr + b"\n" + b"\n".join(lines) + b"\n")
    data = b"".join(parts)
    assert len(data) > (2 << 20)
    want, totals, _ = oracle_sketch(oracle, data, "mash", 5000, 21, 0)
    got = gpu_sketch(fb, data, "mash", 5000, 21, 0)
    assert_same(got[0], want, 21)
    assert got[1] == totals


@pytest.mark.parametrize("fmt", ["fasta", "fastq"])
def test_three_kernel_parse_still_matches(fb, synth, oracle, monkeypatch, fmt):
    """FB2_PARSE_V1=1: the phase / scan / pack pipeline the fused kernel replaced (kept as the A/B baseline)."""
    monkeypatch.setenv("FB2_CHUNK_MB", "1")
    monkeypatch.setenv("FB2_PARSE_V1", "1")
    genome = synth.synth_genome(300_000, 5)
    if fmt == "fastq":
        data = synth.synth_fastq(genome, 20000, 150, 0.005, 6)[0].tobytes()
    else:
        data = synth.synth_fasta(400_000, n_records=7, line_width=70, lower_frac=0.05, n_frac=0.001, seed=9).tobytes()
    want, totals, _ = oracle_sketch(oracle, data, "mash", 2000, 21, 0)
    got = gpu_sketch(fb, data, "mash", 2000, 21, 0)
    assert_same(got[0], want, 21)
    assert got[1] == totals


# ---- hash pieces planned by the parse kernel ---------------------------------------------------------
@pytest.mark.parametrize("k", [21, 31, 7, 32])
@pytest.mark.parametrize("pieces_env", ["1", "0"])
def test_hash_pieces_of_records(fb, oracle, monkeypatch, k, pieces_env):
    """The hash kernel walks the runs of k-mer end positions the parse kernel plans per record (no run starts inside
    the k - 1 positions after a record break).  Read lengths around every boundary of the plan (shorter than k, exactly
    k, one run / two runs / many runs), long stretches of reads without any k-mer between normal ones (their pieces lie
    further apart than a staged group may span), Ns inside reads, CRLF.  FB2_PIECES=0: uniform pieces, same result."""
    monkeypatch.setenv("FB2_CHUNK_MB", "1")
    monkeypatch.setenv("FB2_PIECES", pieces_env)
    rng = np.random.default_rng(1000 + k)
    pm = 67 - k
    lens = [0, 1, k - 1, k, k + 1, pm, pm + k - 2, pm + k - 1, pm + k, 2 * pm + k - 1, 2 * pm + k, 150, 151, 300, 5000, 40000]
    recs = []
    r = 0
    while sum(len(x) for x in recs) < (3 << 20):
        if rng.random() < 0.02:                       # a stretch of reads too short for any k-mer
            for _ in range(int(rng.integers(50, 900))):
                m = int(rng.integers(0, k))
                recs.append(b"@t%d\n" % r + gen.rand_seq(rng, m, 0.0) + b"\n+\n" + b"I" * m + b"\n"); r += 1
        m = int(lens[int(rng.integers(0, len(lens)))]) if rng.random() < 0.5 else int(rng.integers(0, 400))
        nl = b"\r\n" if rng.random() < 0.1 else b"\n"
        recs.append(b"@r%d" % r + nl + gen.rand_seq(rng, m, 0.02) + nl + b"+" + nl + b"I" * m + nl); r += 1
    data = b"".join(recs)
    want, totals, _ = oracle_sketch(oracle, data, "mash", 4000, k, 0)
    got = gpu_sketch(fb, data, "mash", 4000, k, 0)
    assert_same(got[0], want, k)
    assert got[1] == totals


# ---- one FASTQ stream from both ends: host-framed front, raw back (FB2_HOST_STRIP=2) ----------------------
@pytest.mark.parametrize("crlf", [False, True])
def test_two_ended_stream(fb, synth, oracle, monkeypatch, crlf):
    """sketch_stream_two_ended (hostlogic.cpp): handle A frames records on the host from the front of the stream,
    handle B takes raw ranges from its back, the tables are united exactly.  Same sketch as the oracle's, and the
    same errors as the plain path for malformed input (which the two-ended path hands back to it)."""
    monkeypatch.setenv("FB2_CHUNK_MB", "1")
    monkeypatch.setenv("FB2_HOST_STRIP", "2")
    monkeypatch.setenv("FB2_TWO_ENDED_UNIT_KB", "512")
    monkeypatch.setenv("FB2_TWO_ENDED_MIN_KB", "1024")
    monkeypatch.setenv("FB2_TRACE_TWO_ENDED", "1")
    genome = synth.synth_genome(200_000, 11)
    data = synth.synth_fastq(genome, 30000, 150, 0.005, 12)[0].tobytes()
    if crlf:
        data = data.replace(b"\n", b"\r\n")
    assert len(data) > (8 << 20)
    sp = fb.SketchParams.from_cli("mash", n_hashes=1000, kmer_length=21, filters_enabled=True)
    fp = fb.FilterParams(True, (None, None), 0.21, 0.1)
    sk = fb.sketch_stream(data, "two.fq", sp, fp)
    rc, osk = oracle.sketch_stream(data, oracle.mash_params(200000, 1000, False, 21, 0), oracle.make_filter(True, (None, None), 0.21, 0.1))
    assert rc == oracle.OK
    assert np.array_equal(sk.hashes_u64, osk["hashes"]) and np.array_equal(sk.counts, osk["counts"])
    assert np.array_equal(sk.extra_counts, osk["extras"])
    assert [sk.kmers[i, :21].tobytes() for i in range(len(sk))] == osk["kmers"]
    assert (sk.seq_length, sk.num_valid_kmers) == (osk["seq_length"], osk["num_valid_kmers"])
    st = fb.last_stream_stats()
    assert st["h2d_bytes"] > 0 and st["kernel_launches"] > 0
    # a record with unequal sequence / quality lengths somewhere in the middle: the reader's verdict, as always
    nl = b"\r\n" if crlf else b"\n"
    lines = data.split(nl)
    lines[4 * 15000 + 3] += b"I"
    bad = nl.join(lines)
    rc, _ = oracle.sketch_stream(bad, oracle.mash_params(200000, 1000, False, 21, 0), oracle.make_filter(True, (None, None), 0.21, 0.1))
    assert rc == oracle.E_RECORD
    with pytest.raises(fb.FinchError) as ei:
        fb.sketch_stream(bad, "bad.fq", sp, fp)
    assert ei.value.code == fb.ERECORD


# ---- AllCountsSketcher (`--sketch-type none`, counts.rs) and minmer_matrix (distance.rs:344-364) -----------------
@pytest.mark.parametrize("k", [1, 3, 4, 8, 11])
@pytest.mark.parametrize("fmt", ["fasta", "fastq"])
def test_allcounts_sketcher(fb, oracle, monkeypatch, k, fmt):
    """Counts of all 4^k forward k-mers, folded with their reverse complements by to_vec exactly as counts.rs does
    (palindromes count twice; an index whose reverse complement came first is skipped)."""
    monkeypatch.setenv("FB2_CHUNK_MB", "1")
    rng = np.random.default_rng(300 + k)
    data = gen.fasta(rng, n_records=8, min_len=200000, max_len=400000, width=61, messy=0.01) if fmt == "fasta" else \
        gen.fastq(rng, n_records=12000, max_len=300, messy=0.01)
    assert len(data) > (1 << 20)
    o = oracle.AllCountsSketcher(k)
    rc, _, recs = oracle.parse_fastx(data)
    assert rc == oracle.OK
    for r in recs:
        o.process(r)
    want = o.to_vec()
    with fb.AllCountsSketcher(k) as s:
        s.feed_fastx(data, final=True)
        h, c, x, km, seq_len, n_kmers, _ = s.to_arrays()
        assert s.total_bases_and_kmers() == o.total_bases_and_kmers() == (0, n_kmers)
        assert s.parameters().kind == fb.KIND_ALLCOUNTS and s.parameters().expected_size() == 4 ** k
        with pytest.raises(fb.FinchError):
            s.push(b"A" * k, 0)
    assert np.array_equal(h, want["hashes"]) and np.array_equal(c, want["counts"]) and np.array_equal(x, want["extras"])
    assert [km[i, :k].tobytes() for i in range(len(h))] == want["kmers"]
    assert seq_len == 0
    # the same through process(): one record at a time
    with fb.AllCountsSketcher(k) as s:
        for r in recs[:200]:
            s.process(r)
        o2 = oracle.AllCountsSketcher(k)
        for r in recs[:200]:
            o2.process(r)
        h2, c2, x2, *_ = s.to_arrays()
        w2 = o2.to_vec()
        assert np.array_equal(h2, w2["hashes"]) and np.array_equal(c2, w2["counts"]) and np.array_equal(x2, w2["extras"])
    # sketch_stream with the type taken from the command line (cli.rs:336)
    sk = fb.sketch_stream(data, "all.fx", fb.SketchParams.from_cli("none", kmer_length=k), fb.FilterParams(False))
    assert np.array_equal(sk.hashes_u64, want["hashes"]) and np.array_equal(sk.counts, want["counts"])


def test_allcounts_refuses_large_k(fb):
    with pytest.raises(fb.FinchError) as e:
        fb.AllCountsSketcher(17)
    assert e.value.code == fb.ENOMEM


def test_minmer_matrix(fb, oracle):
    rng = np.random.default_rng(17)
    ref = np.unique(rng.integers(0, 1 << 62, size=5000, dtype=np.uint64))
    sketches = []
    for i in range(40):
        own = np.unique(rng.integers(0, 1 << 62, size=int(rng.integers(0, 3000)), dtype=np.uint64))
        shared = rng.choice(ref, size=int(rng.integers(0, 2000)), replace=False)
        h = np.unique(np.concatenate([own, shared]))
        sketches.append((h, rng.integers(1, 1000, size=len(h), dtype=np.uint32)))
    sketches.append((np.zeros(0, np.uint64), np.zeros(0, np.uint32)))                  # an empty sketch
    sketches.append((ref.copy(), np.full(len(ref), 7, np.uint32)))                     # the reference itself
    got = fb.minmer_matrix(ref, sketches)
    want = oracle.minmer_matrix(ref, sketches)
    assert got.dtype == np.int32 and got.shape == (len(sketches), len(ref))
    assert np.array_equal(got, want)
    assert (got[-1] == 7).all() and not got[-2].any()

#!/usr/bin/env python
"""A small tour of every kernel family, meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_small.py
    compute-sanitizer --tool racecheck python tools/sanitize_small.py
    compute-sanitizer --tool initcheck python tools/sanitize_small.py     (optional: --part N runs one part)

Inputs are a few hundred KB so that the instrumented kernels finish in minutes; every result is still compared with the
oracle (test infrastructure), so a run also says the instrumented kernels computed the right thing."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
os.environ.setdefault("FB2_CHUNK_MB", "1")        # several chunks, seams, the asynchronous path
import numpy as np

import finch_rs_b200 as fb
import gen
import oracle

oracle.build()
only = int(sys.argv[sys.argv.index("--part") + 1]) if "--part" in sys.argv else None


def check(sk, osk, k, what):
    assert np.array_equal(sk.hashes_u64, osk["hashes"]), what
    assert np.array_equal(sk.counts, osk["counts"]) and np.array_equal(sk.extra_counts, osk["extras"]), what
    assert [sk.kmers[i, :k].tobytes() for i in range(len(sk))] == osk["kmers"], what
    assert (sk.seq_length, sk.num_valid_kmers) == (osk["seq_length"], osk["num_valid_kmers"]), what
    print("ok:", what, len(sk), "hashes", flush=True)


def part(n):
    return only is None or only == n


rng = np.random.default_rng(7)
if part(1):   # FASTQ, several chunks, filters on (the metric path in small)
    genome = gen.rand_seq(rng, 20000)
    reads = []
    for i in range(9000):
        p = int(rng.integers(0, len(genome) - 150))
        seq = genome[p:p + 150]
        if rng.random() < 0.5:
            seq = seq[::-1].translate(bytes.maketrans(b"ACGT", b"TGCA"))   # both strands, or the strand filter drops everything
        reads.append(b"@r%d\n" % i + seq + b"\n+\n" + b"I" * 150 + b"\n")
    data = b"".join(reads)
    sp = fb.SketchParams.mash(2000, 100, True, 21, 0)
    fp = fb.FilterParams(True, (None, None), 0.21, 0.1)
    rc, osk = oracle.sketch_stream(data, oracle.mash_params(2000, 100, True, 21, 0), oracle.make_filter(True, (None, None), 0.21, 0.1))
    assert rc == oracle.OK
    for mode in ("0", "1", "2"):
        os.environ["FB2_HOST_STRIP"] = mode
        check(fb.sketch_stream(data, "reads.fq", sp, fp), osk, 21, f"FASTQ {len(data)} bytes, strip mode {mode}")
    os.environ.pop("FB2_HOST_STRIP")
    os.environ["FB2_PARSE_V1"] = "1"
    check(fb.sketch_stream(data, "reads.fq", sp, fp), osk, 21, "FASTQ, three-kernel parse")
    os.environ.pop("FB2_PARSE_V1")
    os.environ["FB2_MULTI_OVERSUBSCRIBE"] = "1"
    check(fb.sketch_stream_multi(data, "reads.fq", sp, fp, ngpus=3), osk, 21, "FASTQ cut into 3 ranges, tables merged")
    os.environ.pop("FB2_MULTI_OVERSUBSCRIBE")

if part(2):   # messy FASTA / FASTQ, small k and k = 31 scaled, k = 48 (multi-word kernel)
    for k, kind in ((7, "mash"), (31, "scaled"), (48, "mash"), (32, "mash")):
        for maker in (gen.fasta, gen.fastq):
            data = maker(rng, 60, 50, 3000, messy=0.03) if maker is gen.fasta else maker(rng, 400, 1, 400, messy=0.03, crlf=True)
            sp = fb.SketchParams.mash(500, 500, True, k, 42) if kind == "mash" else fb.SketchParams.scaled(100, k, 0.01, 42)
            osp = oracle.mash_params(500, 500, True, k, 42) if kind == "mash" else oracle.scaled_params(100, k, 0.01, 42)
            fpo = fb.FilterParams(False, (None, None), 0.0, 0.0)
            rc, osk = oracle.sketch_stream(data, osp, oracle.make_filter(False))
            assert rc == oracle.OK
            check(fb.sketch_stream(data, "x", sp, fpo), osk, k, f"{maker.__name__} k={k} {kind}")

if part(3):   # process() / push(), AllCounts
    with fb.SketchParams.mash(300, 300, True, 21, 0).create_sketcher() as s:
        o = oracle.Sketcher.mash(300, 21, 0)
        for _ in range(50):
            r = gen.rand_seq(rng, int(rng.integers(0, 2000)), 0.02)
            s.process(r); o.process(r)
        for _ in range(20):
            km = gen.rand_seq(rng, 21)
            s.push(km, 1); o.push(km, 1)
        got, want = s.to_vec(), o.to_vec()
        assert [g.hash for g in got] == [int(h) for h in want["hashes"]] and [g.kmer for g in got] == want["kmers"]
        assert [g.count for g in got] == [int(c) for c in want["counts"]] and [g.extra_count for g in got] == [int(x) for x in want["extras"]]
        print("ok: process / push", len(got), flush=True)
    with fb.SketchParams.allcounts(4).create_sketcher() as s:
        o = oracle.AllCountsSketcher(4)
        for _ in range(10):
            r = gen.rand_seq(rng, 500, 0.02)
            s.process(r); o.process(r)
        got, want = s.to_vec(), o.to_vec()
        assert [g.hash for g in got] == [int(h) for h in want["hashes"]] and [g.count for g in got] == [int(c) for c in want["counts"]]
        print("ok: AllCounts", len(got), flush=True)

if part(4):   # dist: pair list, tiled all pairs, the cut, minmer_matrix
    base = np.unique(rng.integers(0, 1 << 62, size=3000, dtype=np.uint64))
    sks = [np.sort(rng.choice(base, size=int(rng.integers(1, 1000)), replace=False)) for _ in range(40)] + [np.zeros(0, np.uint64)]
    q = rng.integers(0, len(sks), size=200); r = rng.integers(0, len(sks), size=200)
    out = fb.dist_batch(sks, q, r, 0.0)
    for t in range(200):
        cont, jac, com, tot = oracle.raw_distance(sks[q[t]], sks[r[t]], 0.0)
        assert fb._finish_pair(out[t], 21)[:2] == (cont, jac)
    mat, lens, _ = fb._pack(sks)
    allp = fb.dist_all_pairs(mat, lens)
    pairs = fb.dist_batch(sks, np.repeat(np.arange(len(sks)), len(sks)), np.tile(np.arange(len(sks)), len(sks)), 0.0)
    assert np.array_equal(allp.reshape(-1, 3), pairs)
    hits = fb.dist_all_pairs_cut(mat, lens, 21, 0.2)
    assert len(hits) and (hits["q"] != hits["r"]).all()
    m = fb.minmer_matrix(sks[0], [(s, np.ones(len(s), np.uint32)) for s in sks[1:6]])
    assert np.array_equal(m, oracle.minmer_matrix(sks[0], [(s, np.ones(len(s), np.uint32)) for s in sks[1:6]]))
    print("ok: dist", len(hits), "hits", flush=True)
if part(5):   # the cut through the inverted index (postings, tile radix sort, counting kernel) against the tile kernel
    import synth
    n = 300
    mat = synth.synth_sketches(n, 1000, 6, 5)
    lens = np.full(n, 1000, np.uint32); lens[11] = 3; lens[200] = 999
    os.environ["FB2_DIST_INVERTED"] = "0"
    want = fb.dist_all_pairs_cut(mat, lens, 21, 0.3)
    os.environ["FB2_DIST_INVERTED"] = "1"
    for cb in ("", "64"):
        if cb: os.environ["FB2_DIST_CB"] = cb
        got = fb.dist_all_pairs_cut(mat, lens, 21, 0.3)
        assert np.array_equal(got, want) and len(want) > 0
    os.environ.pop("FB2_DIST_CB"); os.environ.pop("FB2_DIST_INVERTED")
    print("ok: dist through the inverted index", len(want), "hits", flush=True)
print("all parts done", flush=True)

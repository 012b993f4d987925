"""The reference's Python module (lib/src/python.rs) served by the engine: finch_rs_b200/pyfinch.py, importable as
`finch` from <repo>/python.  Host logic (merge, compare_counts, counts, the Multisketch container, sketch files) is
checked on the CPU against the literal restatements in oracle/oracle.py; everything that computes on the GPU
(sketch_file, compare, best_match, filter_to_matches, compare_matrix) is gpu-marked and checked against the oracle."""
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python"))

import finch  # noqa: E402  (python/finch -> finch_rs_b200.pyfinch)
from finch_rs_b200 import pyfinch  # noqa: E402

QUERY_FA = os.path.join(ROOT, "tests", "golden", "query.fa")


def mk(name, hashes, counts=None, extras=None, kmers=None, fb=None, scaled=None, k=21, seed=0):
    """a Sketch with the given entries (what sketch_file / Multisketch.open would have produced)"""
    import finch_rs_b200 as m
    s = finch.Sketch(name)
    h = np.asarray(hashes, np.uint64)
    s.s.hashes = h
    s.s.counts = np.asarray(counts if counts is not None else np.ones(len(h)), np.uint32)
    s.s.extra_counts = np.asarray(extras if extras is not None else np.zeros(len(h)), np.uint32)
    s.s.kmers = list(kmers) if kmers is not None else [b"K%d" % i for i in range(len(h))]
    s.s.sketch_params = m.SketchParams.scaled(1000, k, scaled, seed) if scaled is not None else m.SketchParams.mash(1000, 1000, True, k, seed)
    return s


def rand_entries(rng, n, space=1 << 20):
    h = np.unique(rng.integers(0, space, size=n, dtype=np.uint64))
    return h, rng.integers(1, 50, size=len(h), dtype=np.uint32), rng.integers(0, 5, size=len(h), dtype=np.uint32)


def same_float(a, b):
    return (math.isnan(a) and math.isnan(b)) or a == b


# ---- module surface -------------------------------------------------------------------------------------------
def test_module_surface():
    assert finch.sketch_file is pyfinch.sketch_file and finch.Sketch is pyfinch.Sketch
    assert issubclass(finch.FinchError, Exception)
    s = finch.Sketch("abc")
    assert repr(s) == '<Sketch "abc">' and len(s) == 0 and s.name == "abc"
    assert s.seq_length == 0 and s.num_valid_kmers == 0 and s.comment == "" and s.hashes == []
    assert s.sketch_params == {"sketch_type": "mash", "kmers_to_sketch": 1000, "final_size": 1000, "no_strict": True,
                               "kmer_length": 21, "hash_seed": 0}                   # python.rs:321-327
    s.name = "x"; s.comment = "c"
    assert (s.name, s.comment) == ("x", "c")
    with pytest.raises(AttributeError):
        s.seq_length = 3                                                            # getter only
    with pytest.raises(TypeError):
        finch.Multisketch()                                                         # no #[new]
    c = s.copy()
    c.name = "y"
    assert s.name == "x"


# ---- merge ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("size", [None, 0, 7, 100000])
@pytest.mark.parametrize("scale", [None, 0.5, 0.001])
def test_merge_matches_reference_walk(oracle, seed, size, scale):
    rng = np.random.default_rng(seed)
    n1, n2 = int(rng.integers(0, 400)), int(rng.integers(0, 400))
    space = (1 << 64) - 1 if scale is not None else 1 << 12
    h1, c1, x1 = rand_entries(rng, n1, space)
    h2, c2, x2 = rand_entries(rng, n2, space)
    if seed == 3 and len(h1) and len(h2):        # equal maxima: nothing is dropped
        h2[-1] = h1[-1] = max(h1[-1], h2[-1])
    if seed == 4 and len(h1):
        c1[0] = 0xFFFFFFFF                       # the u32 addition wraps (release build)
        if len(h2):
            h2[0] = h1[0] = min(h1[0], h2[0])
    a = mk("a", h1, c1, x1, scaled=scale)
    b = mk("b", h2, c2, x2, kmers=[b"B%d" % i for i in range(len(h2))], scaled=scale)
    a.s.seq_length, a.s.num_valid_kmers, b.s.seq_length, b.s.num_valid_kmers = 10, 7, 5, 3
    want = oracle.py_merge_sketches(a.hashes, b.hashes, size, scale)
    b_before = b.hashes
    a.merge(b, size)
    assert a.hashes == want
    assert (a.seq_length, a.num_valid_kmers) == (15, 10)
    assert b.hashes == b_before                  # the argument is not touched
    if len(h1) and len(h2) and size is None and scale is None:
        # the quirk in words: nothing above the smaller of the two maxima survives
        assert max(h for h, *_ in a.hashes) == min(int(h1[-1]), int(h2[-1]))


def test_merge_incompatible_params():
    a, b = mk("a", [1, 2, 3]), mk("b", [2, 3, 4], k=31)
    a.s.seq_length, b.s.seq_length = 4, 6
    with pytest.raises(finch.FinchError, match="^First sketch has k 21, but second sketch has k 31$"):
        a.merge(b)
    assert a.seq_length == 10 and len(a) == 3    # the totals were added before the check (python.rs:26-27)
    c = mk("c", [1], seed=5)
    with pytest.raises(finch.FinchError, match="First sketch has hash seed 0, but second sketch has hash seed 5"):
        a.merge(c)
    with pytest.raises(TypeError):
        a.merge("not a sketch")


# ---- compare_counts -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(12))
def test_compare_counts_matches_reference_walk(oracle, seed):
    rng = np.random.default_rng(100 + seed)
    space = 1 << (6 + seed % 7)
    h1, c1, x1 = rand_entries(rng, int(rng.integers(0, 300)), space)
    h2, c2, x2 = rand_entries(rng, int(rng.integers(0, 300)), space)
    if seed == 5 and len(h1) and len(h2):
        h2[-1] = h1[-1] = max(h1[-1], h2[-1]) + 1
    ref, qry = mk("r", h1, c1, x1), mk("q", h2, c2, x2)
    got = ref.compare_counts(qry)
    want = oracle.py_compare_counts(ref.hashes, qry.hashes)
    assert got[:5] == want[:5]
    for g, w in zip(got[5:], want[5:]):
        assert same_float(g, w), (got, want)     # bit-exact f64: same operations in the same order


def test_compare_counts_empty_and_disjoint(oracle):
    e, a, b = mk("e", []), mk("a", [1, 3, 5]), mk("b", [2, 4, 6, 8])
    for ref, qry in ((e, a), (a, e), (a, b), (b, a), (e, e)):
        got, want = ref.compare_counts(qry), oracle.py_compare_counts(ref.hashes, qry.hashes)
        assert got[:5] == want[:5] and all(same_float(g, w) for g, w in zip(got[5:], want[5:]))
        assert got[0] == 0 and math.isnan(got[5])


# ---- counts ---------------------------------------------------------------------------------------------------
def test_counts_getter_and_setter(oracle):
    rng = np.random.default_rng(5)
    h, c, x = rand_entries(rng, 200)
    s = mk("s", h, c, x)
    assert s.counts.dtype == np.int32 and np.array_equal(s.counts, c.astype(np.int32))
    new = rng.integers(0, 4, size=len(h)).astype(np.int32)
    want = oracle.py_set_counts(s.hashes, new.tolist())
    s.counts = new
    assert s.hashes == want and len(s) == int((new > 0).sum())
    with pytest.raises(finch.FinchError, match="counts must be same length as sketch"):
        s.counts = np.zeros(3, np.int32)
    bad = np.ones(len(s), np.int32)
    bad[1] = -4
    before = s.hashes
    with pytest.raises(finch.FinchError, match="Negative count -4 not supported"):
        s.counts = bad
    assert s.hashes == before
    assert oracle.py_set_counts(before, bad.tolist()) == "Negative count -4 not supported"


# ---- Multisketch container ------------------------------------------------------------------------------------
def test_multisketch_container():
    sk = [mk(n, [i + 1, i + 10]) for i, n in enumerate(["a", "b", "c", "b"])]
    ms = finch.Multisketch.from_sketches(sk)
    assert repr(ms) == "<Multisketch (4 sketches)>" and len(ms) == 4
    assert repr(finch.Multisketch.from_sketches(sk[:1])) == "<Multisketch (1 sketch)>"
    assert [s.name for s in ms] == ["a", "b", "c", "b"]
    assert ms[0].name == "a" and ms[3].name == "b" and ms[np.int64(2)].name == "c" and ms[True].name == "b"
    assert ms["b"].hashes == sk[1].hashes                    # the first sketch of that name
    got = ms[1]
    got.name = "changed"
    assert ms[1].name == "b"                                 # items are copies (python.rs:155)
    with pytest.raises(IndexError, match="index out of range"):
        ms[4]
    with pytest.raises(IndexError):
        ms[-5]
    with pytest.raises(KeyError):
        ms["zzz"]
    with pytest.raises(finch.FinchError, match="key is not a string or integer"):
        ms[1.5]
    # the reference maps -1 to len + 1 (python.rs:284-286) and panics on the access
    with pytest.raises(finch.PanicException, match="index out of bounds: the len is 4 but the index is 5"):
        ms[-1]
    with pytest.raises(finch.PanicException):
        del ms[-4]
    assert not issubclass(finch.PanicException, Exception)
    assert "a" in ms and "zzz" not in ms
    with pytest.raises(TypeError):
        3 in ms
    del ms["b"]
    assert [s.name for s in ms] == ["a", "c", "b"]
    del ms[0]
    assert [s.name for s in ms] == ["c", "b"]
    ms.add(mk("d", [5]))
    ms.filter_to_names(["d", "c", "nope"])
    assert [s.name for s in ms] == ["c", "d"]
    with pytest.raises(TypeError):
        ms.filter_to_names(("c",))
    it = iter(ms)
    assert next(it).name == "c" and next(it).name == "d"
    with pytest.raises(StopIteration):
        next(it)


# ---- sketch files ---------------------------------------------------------------------------------------------
def sketch_set(fb):
    rng = np.random.default_rng(9)
    out = []
    for i in range(3):
        h, c, x = rand_entries(rng, 50 + 30 * i, (1 << 64) - 1)
        kmers = ["".join("ACGT"[int(q)] for q in rng.integers(0, 4, size=21)).encode() for _ in range(len(h))]
        s = mk(f"file{i}.fa", h, c, x, kmers=kmers)
        s.s.seq_length, s.s.num_valid_kmers, s.comment = 1000 + i, 900 + i, "" if i else "first"
        s.s.sketch_params = fb.SketchParams.mash(len(h), len(h), False, 21, 0)
        out.append(s)
    return out


def test_multisketch_save_and_open(fb, tmp_path):
    sks = sketch_set(fb)
    ms = finch.Multisketch.from_sketches(sks)
    p = str(tmp_path / "set.bsk")
    ms.save(p)
    back = finch.Multisketch.open(p)
    assert len(back) == 3
    for a, b in zip(sks, back):
        assert (a.name, a.seq_length, a.num_valid_kmers, a.comment) == (b.name, b.seq_length, b.num_valid_kmers, b.comment)
        assert a.hashes == b.hashes and a.sketch_params == b.sketch_params
    # save() writes the finch binary format whatever the name says (python.rs:180-186)
    q = str(tmp_path / "named.sk")
    ms.save(q)
    assert open(q, "rb").read() == open(p, "rb").read()
    with pytest.raises(finch.FinchError, match="Error parsing"):
        finch.Multisketch.open(q)
    # the other two formats (what the command line writes) open as well
    j, m = str(tmp_path / "set.sk"), str(tmp_path / "set.msh")
    ms._save_as(j, pyfinch.FILE_SK)
    ms._save_as(m, pyfinch.FILE_MSH)
    bj, bm = finch.Multisketch.open(j), finch.Multisketch.open(m)
    for a, b, c in zip(sks, bj, bm):
        assert [(h, k, n) for h, k, n, _ in a.hashes] == [(h, k, n) for h, k, n, _ in b.hashes]   # .sk: extra_count is not stored
        assert [h for h, *_ in a.hashes] == [h for h, *_ in c.hashes] and c.name == a.name         # .msh: hashes (+ counts)
        assert b.seq_length == a.seq_length and c.seq_length == a.seq_length


def test_open_errors(tmp_path):
    with pytest.raises(finch.FinchError, match=r"File suffix is not \*\.bsk, \*\.msh, or \*\.sk"):
        p = tmp_path / "x.txt"
        p.write_text("{}")
        finch.Multisketch.open(str(p))
    with pytest.raises(finch.FinchError, match="Error opening"):
        finch.Multisketch.open(str(tmp_path / "missing.sk"))
    bad = tmp_path / "bad.sk"
    bad.write_text("not json")
    with pytest.raises(finch.FinchError, match="Error parsing"):
        finch.Multisketch.open(str(bad))
    with pytest.raises(finch.FinchError, match="Could not create"):
        finch.Multisketch.from_sketches([mk("a", [1])]).save(str(tmp_path / "no_such_dir" / "a.bsk"))


# ---- GPU: sketch_file, compare, best_match, filter_to_matches, compare_matrix -------------------------------------
def oracle_sketch(oracle, path, n_hashes, final_size, k, filt, seed, no_strict):
    sp = oracle.mash_params(n_hashes, n_hashes if final_size is None else final_size, no_strict, k, seed)
    fp = oracle.make_filter(filt, (None, None), 1.0, 0.1)
    return oracle.sketch_stream(open(path, "rb").read(), sp, fp)


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(), dict(n_hashes=10, filter=False), dict(n_hashes=50, final_size=20, kmer_length=11, seed=42, filter=False),
                                dict(n_hashes=100, kmer_length=31, filter=True, no_strict=True)])
def test_sketch_file_matches_oracle(oracle, kw, tmp_path):
    path = QUERY_FA
    if kw.get("kmer_length") == 31:              # a FASTQ with coverage, so that the filters have something to do
        path = str(tmp_path / "reads.fq")
        open(path, "wb").write(make_fastq(np.random.default_rng(3)))
    args = dict(n_hashes=1000, final_size=None, kmer_length=21, filter=True, seed=0, no_strict=False)
    args.update(kw)
    rc, want = oracle_sketch(oracle, path, args["n_hashes"], args["final_size"], args["kmer_length"], args["filter"], args["seed"], args["no_strict"])
    if rc != oracle.OK:
        with pytest.raises(finch.FinchError, match="had too few kmers"):
            finch.sketch_file(path, **kw)
        return
    s = finch.sketch_file(path, **kw)
    assert s.name == path and repr(s) == f'<Sketch "{path}">'
    assert (s.seq_length, s.num_valid_kmers) == (want["seq_length"], want["num_valid_kmers"])
    assert s.hashes == [(int(h), k, int(c), int(x)) for h, k, c, x in zip(want["hashes"], want["kmers"], want["counts"], want["extras"])]
    assert s.sketch_params["kmers_to_sketch"] == args["n_hashes"] and s.sketch_params["kmer_length"] == args["kmer_length"]


def make_fastq(rng):
    genome = rng.integers(0, 4, size=3000)
    out = []
    for i in range(800):
        p = int(rng.integers(0, 2900))
        seq = "".join("ACGT"[int(q)] for q in genome[p:p + 100])
        out.append(f"@r{i}\n{seq}\n+\n{'I' * len(seq)}\n")
    return "".join(out).encode()


@pytest.mark.gpu
def test_sketch_file_golden_kmers():
    """cli/tests/test_cli.rs:80-149 through the Python surface: the ten k-mers of query.fa, in hash order"""
    s = finch.sketch_file(QUERY_FA, n_hashes=10, filter=False)
    assert [k.decode() for _, k, _, _ in s.hashes] == [
        "ATGCTAGCTACGTAACGTCGC", "CAGTCGATCGATCGTAGCTGA", "CTCAGATGCTGAGCCGGTCTA", "GCTAGCTAGCATCGCTAGCTA", "GACTAGCTAGCTAGCTAGCGA",
        "CGCTAGCTACGATCGATCGAC", "TAATTTATACGGGCCTATTAA", "GCATCAGCTAGCATCGCTGTA", "AGCCGGTCTACTACTACACAT", "AAGGCCTAACTTAATAGGCCC"]
    with pytest.raises(finch.FinchError, match="No such file"):
        finch.sketch_file("/nonexistent/file.fa")


@pytest.mark.gpu
def test_compare_best_match_filter_and_matrix(oracle):
    rng = np.random.default_rng(21)
    base = np.unique(rng.integers(0, 1 << 62, size=1500, dtype=np.uint64))
    query = mk("q", np.sort(rng.choice(base, size=800, replace=False)), rng.integers(1, 9, size=800, dtype=np.uint32))
    refs = []
    for i in range(12):
        shared = rng.choice(query.s.hashes, size=int(rng.integers(0, 700)), replace=False)
        own = rng.integers(0, 1 << 62, size=int(rng.integers(1, 900)), dtype=np.uint64)
        h = np.unique(np.concatenate([shared, own]))
        refs.append(mk(f"ref{i}", h, rng.integers(1, 100, size=len(h), dtype=np.uint32)))
    refs.append(mk("empty", []))
    # compare: `other` is the query, `self` the reference
    want = []
    for r in refs:
        cont, jac, _, _ = oracle.raw_distance(query.s.hashes, r.s.hashes, 0.0)
        assert r.compare(query) == (cont, jac)
        want.append(cont)
        od = oracle.old_distance(query.s.hashes, r.s.hashes)
        got_old = r.compare(query, old_mode=True)
        assert same_float(got_old[0], od[0]) and same_float(got_old[1], od[1])   # (an empty reference: 0 / 0)
    # scaled sketches: min of the two scales
    a = mk("a", query.s.hashes, scaled=0.25)
    b = mk("b", refs[0].s.hashes, scaled=0.5)
    cont, jac, _, _ = oracle.raw_distance(b.s.hashes, a.s.hashes, 0.25)
    assert a.compare(b) == (cont, jac)
    ms = finch.Multisketch.from_sketches(refs)
    ix, best = ms.best_match(query)
    first_max = max(range(len(want)), key=lambda i: (want[i], -i)) if max(want) > 0 else 0
    assert ix == first_max and best.name == refs[ix].name
    thr = sorted(want)[len(want) // 2]
    ms.filter_to_matches(query, thr)
    assert [s.name for s in ms] == [r.name for r, c in zip(refs, want) if c >= thr]
    with pytest.raises(finch.PanicException):
        finch.Multisketch.from_sketches([]).best_match(query)
    # compare_matrix: counts of the other sketches aligned to this sketch's hashes
    got = query.compare_matrix(*refs)
    assert got.dtype == np.int32 and got.shape == (len(refs), len(query))
    assert np.array_equal(got, oracle.minmer_matrix(query.s.hashes, [(r.s.hashes, r.s.counts) for r in refs]))


@pytest.mark.gpu
def test_open_files_written_by_the_command_line(tmp_path):
    """`finch sketch` (-o .sk, -b .bsk, -B .msh) -> Multisketch.open: the same ten entries sketch_file returns"""
    import subprocess
    exe = os.path.join(ROOT, "finch_rs_b200", "finch")
    want = finch.sketch_file(QUERY_FA, n_hashes=10, filter=False)
    for flag, ext in ((None, ".sk"), ("-b", ".bsk"), ("-B", ".msh")):
        out = tmp_path / ("q" + ext)
        cmd = [exe, "sketch", "--n-hashes", "10"] + ([flag] if flag else []) + ["-O", QUERY_FA]
        p = subprocess.run(cmd, capture_output=True)
        assert p.returncode == 0, p.stderr
        out.write_bytes(p.stdout)
        ms = finch.Multisketch.open(str(out))
        assert len(ms) == 1 and QUERY_FA in ms
        got = ms[QUERY_FA]
        assert [h for h, *_ in got.hashes] == [h for h, *_ in want.hashes]
        assert (got.seq_length, got.num_valid_kmers) == (want.seq_length, want.num_valid_kmers) or ext == ".msh"
        if ext != ".msh":                      # the Mash format keeps no k-mers
            assert [k for _, k, _, _ in got.hashes] == [k for _, k, _, _ in want.hashes]
            assert [c for _, _, c, _ in got.hashes] == [c for _, _, c, _ in want.hashes]
        assert got.compare(want) == (1.0, 1.0) and want.compare(got) == (1.0, 1.0)
        assert got.sketch_params["sketch_type"] == "mash" and got.sketch_params["kmer_length"] == 21


def test_open_minimal_json_and_empty_collections(tmp_path):
    """`.sk` files may hold hashes only (json.rs:91-139: counts default to 1, extra_count = count / 2, no k-mers); empty
    collections and empty sketches survive a save / open round trip."""
    import json
    base = {"kmer": 21, "alphabet": "ACGT", "preserveCase": False, "canonical": True, "sketchSize": 3,
            "hashType": "MurmurHash3_x64_128", "hashBits": 64, "hashSeed": 0,
            "sketches": [{"name": "x", "seqLength": 5, "numValidKmers": 4, "comment": "c", "hashes": ["1", "5", "18446744073709551615"]}]}
    p = tmp_path / "a.sk"
    p.write_text(json.dumps(base))
    s = finch.Multisketch.open(str(p))[0]
    assert s.hashes == [(1, b"", 1, 0), (5, b"", 1, 0), (2**64 - 1, b"", 1, 0)]
    assert (s.name, s.seq_length, s.num_valid_kmers, s.comment) == ("x", 5, 4, "c")
    assert s.sketch_params == {"sketch_type": "mash", "kmers_to_sketch": 3, "final_size": 3, "no_strict": True, "kmer_length": 21, "hash_seed": 0}
    base["sketches"][0]["kmers"] = ["AAA", "CCC", "GGG"]
    base["sketches"][0]["counts"] = [3, 4, 5]
    p.write_text(json.dumps(base))
    assert finch.Multisketch.open(str(p))[0].hashes == [(1, b"AAA", 3, 1), (5, b"CCC", 4, 2), (2**64 - 1, b"GGG", 5, 2)]
    base["sketches"] = []
    p.write_text(json.dumps(base))
    assert len(finch.Multisketch.open(str(p))) == 0
    e = tmp_path / "e.bsk"
    finch.Multisketch.from_sketches([]).save(str(e))
    assert len(finch.Multisketch.open(str(e))) == 0
    finch.Multisketch.from_sketches([finch.Sketch("empty")]).save(str(e))
    back = finch.Multisketch.open(str(e))
    assert len(back) == 1 and back["empty"].hashes == [] and back[0].sketch_params == finch.Sketch("q").sketch_params

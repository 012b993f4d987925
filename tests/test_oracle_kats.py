"""Pins the CPU oracle on every golden vector the reference's own tests hold for this path
(KATs K-A .. K-G).  CPU only."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


# ---- public MurmurHash3_x64_128 vectors (not from the reference; cross-check of S8) --------
def test_murmur_public_vectors(oracle):
    h1, h2 = oracle.murmur3_x64_128(b"foo", 0)
    assert (h2 << 64) | h1 == 168394135621993849475852668931176482145
    h1, h2 = oracle.murmur3_x64_128(b"The quick brown fox jumps over the lazy dog", 0)
    assert (h1, h2) == (0xe34bbc7bbc071b6c, 0x7a433ca9c49a9347)
    assert oracle.murmur3_x64_128(b"", 0) == (0, 0)


# ---- K-A: mash.rs:136-154 (commented-out test; still a valid KAT) ---------------------------
def test_kat_longer_sequence_seed42(oracle):
    seq = b"ACACGGAAATCCTCACGTCGCGGCGCCGGGC"
    want = [3186265289206375993, 3197567229193635484, 5157287830980272133, 7515070071080094037,
            9123665698461883699, 9650810550987401968, 10462414310441547028, 12872951831549606632,
            13584836512372089324, 14093285637546356047, 16069721578136260683]
    h, rc = oracle.kmer_stream(seq, 21, 42)
    assert sorted(int(x) for x in h) == want
    s = oracle.Sketcher.mash(100, 21, 42)
    s.process(seq)
    assert [int(x) for x in s.to_vec()["hashes"]] == want
    assert s.total_bases_and_kmers() == (31, 11)


# ---- K-B: mash.rs:115-134, scaled.rs:118-161 ------------------------------------------------
def _push4(s):
    s.push(b"ca", 0); s.push(b"cc", 1); s.push(b"ac", 0); s.push(b"ac", 1)
    return s.to_vec(2)


@pytest.mark.parametrize("mk", ["mash", "scaled1", "scaled1000"])
def test_kat_minhashkmers(oracle, mk):
    s = {"mash": lambda: oracle.Sketcher.mash(3, 2, 42),
         "scaled1": lambda: oracle.Sketcher.scaled(3, 1.0, 2, 42),
         "scaled1000": lambda: oracle.Sketcher.scaled(3, 0.001, 2, 42)}[mk]()
    v = _push4(s)
    assert v["kmers"] == [b"cc", b"ca", b"ac"]
    assert list(v["counts"]) == [1, 1, 2]
    assert list(v["extras"]) == [1, 0, 1]
    assert v["hashes"][0] < v["hashes"][1] < v["hashes"][2]


# ---- K-D: scaled.rs:163-176 eviction; scaled.rs:178-200 pure scaled ---------------------------
def test_kat_scaled_eviction(oracle):
    s = oracle.Sketcher.scaled(1, 0.01, 4, 42)
    for km in (b"AAAA", b"AGTA", b"CCCC", b"ATAA"):
        s.push(km, 1 if km == b"CCCC" else 0)
    v = s.to_vec(4)
    assert len(v["kmers"]) == 3 and b"AAAA" not in v["kmers"]
    assert oracle.hash_f(b"AAAA", 42) == 17832910516274425539
    assert sorted(oracle.hash_f(k, 42) for k in (b"AGTA", b"CCCC", b"ATAA")) == \
        [24933659310187264, 73459868045630124, 179996601836427478]
    assert s.max_hash() == (2**64 - 1) // 100


def test_kat_pure_scaled_empty(oracle):
    s = oracle.Sketcher.scaled(0, 0.001, 2, 42)
    assert _push4(s)["kmers"] == []


# ---- K-E: scaled.rs:202-213 property ---------------------------------------------------------
def test_kat_pure_scaled_property(oracle):
    rng = np.random.default_rng(7)
    for _ in range(20):
        seq = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), size=int(rng.integers(500, 900))))
        s = oracle.Sketcher.scaled(0, 1.0 / 100.0, 2, 42)
        for i in range(len(seq) - 3):
            s.push(seq[i:i + 4], 0)
        assert all(int(h) <= (2**64 - 1) // 100 for h in s.to_vec(4)["hashes"])


# ---- K-C: cli/tests/test_cli.rs:80-149 on cli/tests/data/query.fa ----------------------------
KC_KMERS = [b"ATGCTAGCTACGTAACGTCGC", b"CAGTCGATCGATCGTAGCTGA", b"CTCAGATGCTGAGCCGGTCTA",
            b"GCTAGCTAGCATCGCTAGCTA", b"GACTAGCTAGCTAGCTAGCGA", b"CGCTAGCTACGATCGATCGAC",
            b"TAATTTATACGGGCCTATTAA", b"GCATCAGCTAGCATCGCTGTA", b"AGCCGGTCTACTACTACACAT",
            b"AAGGCCTAACTTAATAGGCCC"]


@pytest.mark.parametrize("kind", ["mash", "scaled"])
def test_kat_query_fa(oracle, kind):
    data = open(os.path.join(GOLD, "query.fa"), "rb").read()
    # CLI resolution (cli.rs:277-340): FASTA + auto filter => kmers_to_sketch = 10*200, final 10
    sp = (oracle.mash_params(2000, 10, False, 21, 0) if kind == "mash"
          else oracle.scaled_params(10, 21, 0.001, 0))
    fp = oracle.make_filter(None, (None, None), 1.0 * 21 / 100, 0.1)
    rc, sk = oracle.sketch_stream(data, sp, fp)
    assert rc == oracle.OK
    assert sk["kmers"] == KC_KMERS
    assert sk["format"] == oracle.FMT_FASTA and sk["filter_on"] is False
    assert sk["num_valid_kmers"] == 339
    assert sk["seq_length"] == 405  # believed (raw bytes incl. interior newlines); unpinned upstream
    assert [int(h) for h in sk["hashes"]] == [
        933085113509804, 8582128962097342, 12581283643378369, 13388215406653903,
        59671498055219043, 85163822212241463, 196329111101504065, 240583695071237384,
        241465901919730030, 256930375650047524]


# ---- K-F: distance.rs:176-242 ------------------------------------------------------------------
def test_kat_raw_distance(oracle):
    rd = oracle.raw_distance
    assert rd([0, 1, 2], [1, 2]) == (2. / 2., 2. / 3., 2, 3)
    assert rd([0, 2], [1, 2]) == (1. / 2., 1. / 3., 1, 3)
    assert rd([0, 1], [2, 3]) == (0., 0., 0, 2)
    assert rd([], []) == (0., 1., 0, 0)
    assert rd([], [5]) == (0., 1., 0, 0)
    # scaled: 1e-18 => max_hash 18
    assert rd([10, 15, 20], [15, 20], 1e-18) == (1., 2. / 3., 2, 3)
    assert rd([5, 10, 15], [5, 10], 1e-18) == (1., 2. / 3., 2, 3)
    assert rd([5, 10, 15, 20], [5, 10], 1e-18) == (1., 2. / 3., 2, 3)
    assert rd([5, 10], [5, 10, 15, 20], 1e-18) == (2. / 3., 2. / 3., 2, 3)


def test_kat_raw_distance_commutes(oracle):
    rng = np.random.default_rng(3)
    for _ in range(200):
        a = np.sort(rng.integers(0, 50, size=int(rng.integers(0, 30))).astype(np.uint64))
        b = np.sort(rng.integers(0, 50, size=int(rng.integers(0, 30))).astype(np.uint64))
        a, b = np.unique(a), np.unique(b)
        x, y = oracle.raw_distance(a, b), oracle.raw_distance(b, a)
        # the proptest compares whole tuples on random u64 vectors (almost never intersecting);
        # jaccard/common/total commute in general
        assert x[1:] == y[1:]


def test_kat_distance_scaled_end_to_end(oracle):  # distance.rs:312-337
    a = _push4(oracle.Sketcher.scaled(3, 0.001, 2, 42))["hashes"]
    cont, jac, com, tot = oracle.raw_distance(a, a, 0.001)
    assert (cont, jac, com) == (1.0, 1.0, 3)
    assert oracle.mash_distance(1.0, 2) == 0.0
    assert oracle.mash_distance(0.0, 21) == 1.0


# ---- K-G: filtering.rs:197-327, :345-407, :434-505 ; statistics.rs:53-129 ---------------------
def test_kat_guess_filter_threshold(oracle):
    g = oracle.guess_filter_threshold
    assert g([], 0.2) == 1
    assert g([1], 0.2) == 1
    assert g([1, 1], 0.2) == 1
    assert g([1, 9], 0.2) == 8
    assert g([1, 10, 10, 9], 0.1) == 8
    assert g([1, 1, 2, 4], 0.1) == 1
    assert g([2], 1.0) == 2


def test_kat_filter_abundance_and_strands(oracle):
    # filtering.rs:345-407 (inclusive bounds); returned values are kept indices (hash = idx+1)
    assert list(oracle.filter_abundance([1, 1], 1, None)) == [0, 1]
    assert list(oracle.filter_abundance([1, 10, 10, 9], 9, None)) == [1, 2, 3]
    assert list(oracle.filter_abundance([1, 10, 10, 9], 2, 9)) == [3]
    # filtering.rs:434-505: count < 16 passes; otherwise min-strand ratio >= cutoff
    assert list(oracle.filter_strands([10, 10, 10, 10], [1, 2, 8, 9], 0.15)) == [0, 1, 2, 3]
    assert list(oracle.filter_strands([16, 16, 16, 16], [1, 2, 8, 9], 0.15)) == [2, 3]


def test_kat_hist(oracle):  # statistics.rs:53-129 incl. issue #63's sparse count
    assert list(oracle.hist([1, 1, 1])) == [3]
    h = oracle.hist([4, 2, 4, 3, 126497])
    assert len(h) == 126497
    assert (h[0], h[1], h[2], h[3], h[126497 - 1]) == (0, 1, 1, 2, 1)
    assert list(oracle.hist([])) == []


def test_old_distance_closed_form(oracle):
    """distance.rs:136-157 has no test upstream (unpinned): the literal pointer walk equals |Q n R| over |R|
    for sorted distinct lists, which is what the product computes from the GPU's intersection count."""
    import numpy as np
    rng = np.random.default_rng(3)
    pool = np.unique(rng.integers(0, 2**63, size=400, dtype=np.uint64))
    for _ in range(300):
        q = np.sort(rng.choice(pool, size=int(rng.integers(1, 120)), replace=False))
        r = np.sort(rng.choice(pool, size=int(rng.integers(0, 120)), replace=False))
        got = oracle.old_distance(q, r)
        com, tot = len(np.intersect1d(q, r)), len(r)
        assert got[2:] == (com, tot)
        if tot:
            assert got[0] == com / tot and got[1] == com / (com + 2 * (tot - com))
    assert oracle.old_distance(np.zeros(0, np.uint64), np.array([1], np.uint64)) is None      # the reference panics here

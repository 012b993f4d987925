"""`finch` command line (finch_rs_b200/cli/finch_cli.cpp) and the `.sk` JSON wire format.

CPU part: flag rules, JSON reader/writer, `hist` / `info` on sketch files, number formatting.
GPU part (`-m gpu`): the reference's own CLI tests (cli/tests/test_cli.rs) run against this binary:
same arguments, same assertions, including the golden k-mer lists for tests/data/query.fa.
"""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FINCH = os.path.join(ROOT, "finch_rs_b200", "finch")
QUERY = os.path.join(ROOT, "tests", "golden", "query.fa")

GOLDEN_KMERS = [  # cli/tests/test_cli.rs:97-106 and :134-143 (identical for mash and scaled)
    "ATGCTAGCTACGTAACGTCGC", "CAGTCGATCGATCGTAGCTGA", "CTCAGATGCTGAGCCGGTCTA", "GCTAGCTAGCATCGCTAGCTA",
    "GACTAGCTAGCTAGCTAGCGA", "CGCTAGCTACGATCGATCGAC", "TAATTTATACGGGCCTATTAA", "GCATCAGCTAGCATCGCTGTA",
    "AGCCGGTCTACTACTACACAT", "AAGGCCTAACTTAATAGGCCC"]


@pytest.fixture(scope="session")
def finch():
    if not os.path.exists(FINCH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "finch_rs_b200", "csrc"), "-j8"])

    def run(*args, cwd=None, check=True, env=None):
        p = subprocess.run([FINCH, *args], capture_output=True, text=True, cwd=cwd,
                           env=None if env is None else {**os.environ, **env})
        if check:
            assert p.returncode == 0, p.stderr
        return p
    return run


SK = {"kmer": 21, "alphabet": "ACGT", "preserveCase": False, "canonical": True, "sketchSize": 3,
      "hashType": "MurmurHash3_x64_128", "hashBits": 64, "hashSeed": 0, "scale": None,
      "sketches": [{"name": "a \"q\"\n.fa", "seqLength": 405, "numValidKmers": 339, "comment": "",
                    "filters": {"strandFilter": "0.1", "errFilter": "0.21", "minCopies": "2"},
                    "hashes": ["933085113509804", "8582128962097342", "18446744073709551615"],
                    "kmers": ["ATGCTAGCTACGTAACGTCGC", "CAGTCGATCGATCGTAGCTGA", "CTCAGATGCTGAGCCGGTCTA"],
                    "counts": [1, 3, 3]}]}


def test_number_formatting_matches_serde_and_rust_display(finch):
    vals = ["1", "0.5", "0.001", "1e-5", "1e-6", "1e16", "1e15", "123456.789", "1.5e-7", "0.30000000000000004",
            "0.21", "1e21", "-2.5", "0", "33.333332"]
    want = ["1.0 1 1", "0.5 0.5 0.5", "0.001 0.001 0.001", "0.00001 0.00001 0.00001", "1e-6 0.000001 0.000001",
            "1e16 10000000000000000 10000000000000000", "1000000000000000.0 1000000000000000 1000000000000000",
            "123456.789 123456.789 123456.79", "1.5e-7 0.00000015 0.00000015",
            "0.30000000000000004 0.30000000000000004 0.3", "0.21 0.21 0.21",
            "1e21 1000000000000000000000 1000000000000000000000", "-2.5 -2.5 -2.5", "0.0 0 0",
            "33.333332 33.333332 33.333332"]
    assert finch("fmt-f64", *vals).stdout.split("\n")[:-1] == want


def _ryu_like(v):
    """serde_json / Ryu layout from Python's shortest round-trip digits (repr)."""
    import decimal
    if v == 0:
        return "-0.0" if str(v).startswith("-") else "0.0"
    sign, digits, exp = decimal.Decimal(repr(abs(v))).as_tuple()
    d = "".join(map(str, digits)).rstrip("0") or "0"
    kk = len(digits) + exp                      # v = 0.d1d2.. * 10^kk
    neg = "-" if v < 0 else ""
    if len(d) <= kk <= 16:
        return neg + d + "0" * (kk - len(d)) + ".0"
    if 0 < kk <= 16:
        return neg + d[:kk] + "." + d[kk:]
    if -5 < kk <= 0:
        return neg + "0." + "0" * (-kk) + d
    return neg + d[0] + ("." + d[1:] if len(d) > 1 else "") + "e" + str(kk - 1)


def test_json_f64_random_values(finch):
    """1000 random doubles over many magnitudes: same text as Ryu would print, and it parses back exactly."""
    import random
    import struct
    rnd = random.Random(11)
    vals = []
    for _ in range(1000):
        m = rnd.choice([rnd.random(), rnd.random() * 10 ** rnd.randint(-12, 20), rnd.randint(0, 10 ** 9) / 10 ** rnd.randint(0, 9)])
        vals.append(-m if rnd.random() < 0.2 else m)
    vals += [struct.unpack("<d", struct.pack("<Q", rnd.getrandbits(62)))[0] for _ in range(200)]   # arbitrary bit patterns
    out = finch("fmt-f64", *[repr(v) for v in vals]).stdout.split("\n")[:-1]
    assert len(out) == len(vals)
    for v, line in zip(vals, out):
        got = line.split(" ")[0]
        assert float(got) == v, (v, got)
        assert got == _ryu_like(v), (v, got, _ryu_like(v))


def test_sk_json_round_trip_is_byte_identical(finch, tmp_path):
    """json.rs:64-89,141-158: field order, quoted-decimal hashes, escaped names, `scale: null`."""
    f = tmp_path / "a.sk"
    text = json.dumps(SK, separators=(",", ":"))
    f.write_text(text)
    out = finch("sketch", "-O", str(f)).stdout
    assert json.loads(out) == SK
    # byte-identical except for the key order inside `filters` (a HashMap in the reference)
    a, b = json.loads(out), json.loads(text)
    a["sketches"][0]["filters"] = b["sketches"][0]["filters"] = {}
    assert json.dumps(a, separators=(",", ":")) == json.dumps(b, separators=(",", ":"))
    assert out.startswith('{"kmer":21,"alphabet":"ACGT","preserveCase":false,"canonical":true,"sketchSize":3,'
                          '"hashType":"MurmurHash3_x64_128","hashBits":64,"hashSeed":0,"scale":null,"sketches":[{"name":')
    # -o adds the extension when missing (main.rs:29-37)
    finch("sketch", "-o", str(tmp_path / "out"), str(f))
    assert json.loads((tmp_path / "out.sk").read_text()) == SK


def test_scaled_sk_and_optional_fields(finch, tmp_path):
    sk = dict(SK, scale=0.001, sketchSize=1000)
    sk["sketches"] = [{"name": "x", "seqLength": None, "numValidKmers": None, "comment": None, "filters": None,
                       "hashes": ["5", "7"]}]                      # kmers / counts absent (json.rs:105-121)
    f = tmp_path / "s.json"
    f.write_text(json.dumps(sk))
    out = json.loads(finch("sketch", "-s", "scaled", "-O", str(f)).stdout)
    assert out["scale"] == 0.001 and out["sketchSize"] == 1000
    s = out["sketches"][0]
    assert (s["seqLength"], s["numValidKmers"], s["comment"], s["filters"]) == (0, 0, "", {})
    assert s["hashes"] == ["5", "7"] and s["counts"] == [1, 1] and s["kmers"] == ["", ""]
    # sketch type of the file and of the command line must agree (main.rs:345-349)
    p = finch("sketch", "-O", str(f), check=False)
    assert p.returncode == 1 and "Sketch types are not the same" in p.stderr


def test_hist_and_info_on_sketch_files(finch, tmp_path):
    f = tmp_path / "a.sk"
    f.write_text(json.dumps(SK))
    assert json.loads(finch("hist", str(f)).stdout) == {SK["sketches"][0]["name"]: [1, 0, 2]}   # statistics.rs:30-47
    info = finch("info", str(f)).stdout.split("\n")
    assert info[0] == 'a "q"' and info[1] == ".fa (from 405bp)"
    # cardinality in f32 against usize::MAX (statistics.rs:19-22): hash == u64::MAX -> (3-1)/1.0
    assert info[2] == "  Estimated # of Unique Kmers: 2"
    assert info[3] == "  Estimated Average Depth: 2.3333333x"      # (1 + 3 + 3) / 3 in f32
    gc = sum(k.count("G") + k.count("C") for k in SK["sketches"][0]["kmers"][0:1]) * 1 + \
        sum(k.count("G") + k.count("C") for k in SK["sketches"][0]["kmers"][1:]) * 3
    assert info[4].startswith("  Estimated % GC: ") and abs(float(info[4].split(": ")[1][:-1]) - 100.0 * gc / (7 * 21)) < 1e-3


def test_flag_rules(finch, tmp_path):
    f = tmp_path / "a.sk"
    f.write_text(json.dumps(SK))
    cases = [
        (["sketch", "--scale", "0.1", "-O", str(f)], "`scale` can not be specified for `mash` sketch types"),     # cli.rs:293
        (["sketch", "-s", "scaled", "--oversketch", "3", "-O", str(f)], "`oversketch` can not be specified"),   # cli.rs:314
        (["sketch", "-s", "scaled", "-N", "-O", str(f)], "`no_strict` can not be specified"),                    # cli.rs:317
        (["sketch", "--err-filter", "5", "-O", str(f)], "err-filter must be between 0 and 4.761904761904762"),   # cli.rs:264
        (["sketch", "-n", "x", "-O", str(f)], "n-hashes must be a positive integer"),
        (["sketch", "-k", "20", "-O", str(f)], "Specified kmer length 20 does not match 21 from sketch"),        # main.rs:367
        (["sketch", "--seed", "4", "-O", str(f)], "Specified hash seed 4 does not match 0 from sketch"),
        (["sketch", "-f", "--no-filter", "-O", str(f)], "cannot be used with"),
        (["sketch", "-o", "x", "-O", str(f)], "cannot be used with"),
        (["dist", str(f), "-p", "-q", "a"], "cannot be used with"),
        (["sketch", "-b", "-B", "-O", str(f)], "cannot be used with"),                                          # cli.rs: conflicts_with
        (["sketch", str(f)], "is not a sequence file?"),                                                          # main.rs:213-219
        (["sketch", "-O", str(tmp_path / "missing.sk")], "Error opening"),
        (["frobnicate", "x"], "wasn't expected"),
    ]
    for args, msg in cases:
        p = finch(*args, check=False)
        assert p.returncode == 1 and msg in p.stderr, (args, p.stderr)
    bad = tmp_path / "bad.sk"
    bad.write_text('{"kmer": 21, "sketches": [')
    p = finch("hist", str(bad), check=False)
    assert p.returncode == 1 and "Error parsing" in p.stderr
    # a loaded sketch with -f only gets its filter METADATA updated (filtering.rs:20-52)
    out = json.loads(finch("sketch", "-f", "--min-abun-filter", "3", "-O", str(f)).stdout)
    s = out["sketches"][0]
    assert s["hashes"] == SK["sketches"][0]["hashes"] and s["counts"] == [1, 3, 3]
    assert s["filters"] == {"strandFilter": "0.1", "errFilter": "0.21", "minCopies": "3"}


# ---- the reference's CLI tests (cli/tests/test_cli.rs), against this binary -----------------------------
@pytest.mark.gpu
def test_file_doesnt_exist(finch):                       # test_cli.rs:9-18
    p = finch("sketch", "test/file/doesnt/exist", check=False)
    assert p.returncode != 0 and "No such file or directory" in p.stderr


@pytest.mark.gpu
def test_finch_sketch(finch):                            # test_cli.rs:20-37
    sk = json.loads(finch("sketch", "--n-hashes", "10", "-O", QUERY).stdout)
    assert (sk["kmer"], sk["alphabet"], sk["sketchSize"], sk["hashSeed"]) == (21, "ACGT", 10, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("extra", [["--sketch-type", "scaled", "--scale", ".001"], ["--sketch-type", "mash"]])
def test_finch_sketch_scaled_and_mash(finch, oracle, extra):   # test_cli.rs:80-149
    sk = json.loads(finch("sketch", "--n-hashes", "10", *extra, QUERY, "-O").stdout)
    assert (sk["kmer"], sk["alphabet"], sk["sketchSize"], sk["hashSeed"]) == (21, "ACGT", 10, 0)
    s = sk["sketches"][0]
    assert s["kmers"] == GOLDEN_KMERS
    assert sk["scale"] == (0.001 if "scaled" in extra else None)
    # the rest of the document against the oracle (derived, not pinned upstream: SURVEY 8c)
    data = open(QUERY, "rb").read()
    if "scaled" in extra:
        rc, osk = oracle.sketch_stream(data, oracle.scaled_params(10, 21, 0.001, 0), oracle.make_filter(None, (None, None), 0.21, 0.1))
    else:
        rc, osk = oracle.sketch_stream(data, oracle.mash_params(2000, 10, False, 21, 0), oracle.make_filter(None, (None, None), 0.21, 0.1))
    assert rc == oracle.OK
    assert [int(h) for h in s["hashes"]] == [int(h) for h in osk["hashes"]]
    assert s["counts"] == [int(c) for c in osk["counts"]]
    assert (s["name"], s["seqLength"], s["numValidKmers"], s["comment"], s["filters"]) == \
        (QUERY, osk["seq_length"], osk["num_valid_kmers"], "", {})


@pytest.mark.gpu
def test_sketch_in_place_then_dist(finch, tmp_path):
    """generate_sketch_files (main.rs:201-235) + dist over a sketch file and a sequence file
    (parse_mash_files main.rs:237-313, calc_sketch_distances :315-334, distance.rs:9-47)."""
    rng = np.random.default_rng(7)
    genome = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=30000)
    mut = genome.copy()
    pos = rng.choice(len(mut), size=300, replace=False)
    mut[pos] = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=300)

    def fasta(path, seq, name):
        with open(path, "wb") as fh:
            fh.write(b">" + name + b"\n")
            for i in range(0, len(seq), 70):
                fh.write(seq[i:i + 70].tobytes() + b"\n")
    a, b = tmp_path / "a.fa", tmp_path / "b.fa"
    fasta(a, genome, b"a")
    fasta(b, mut, b"b")
    finch("sketch", "-n", "500", str(a))
    ska = json.loads((tmp_path / "a.fa.sk").read_text())
    assert ska["sketchSize"] == 500 and len(ska["sketches"][0]["hashes"]) == 500
    assert ska["sketches"][0]["name"] == str(a)
    # first sketch is the query by default; the pair (a, a) is skipped; a sequence file joins the sketch file.
    # n-hashes comes from the sketch file (update_sketch_params), so b.fa is sketched with n = 500 too.
    d = json.loads(finch("dist", str(tmp_path / "a.fa.sk"), str(b)).stdout)
    assert len(d) == 1 and list(d[0].keys()) == ["containment", "jaccard", "mashDistance", "commonHashes", "totalHashes",
                                                  "query", "reference"]
    assert (d[0]["query"], d[0]["reference"]) == (str(a), str(b))
    ha = np.array([int(h) for h in ska["sketches"][0]["hashes"]], np.uint64)
    skb = json.loads(finch("sketch", "-n", "500", "-O", str(b)).stdout)
    hb = np.array([int(h) for h in skb["sketches"][0]["hashes"]], np.uint64)
    import oracle as o
    cont, jac, com, tot = o.raw_distance(ha, hb, 0.0)
    assert (d[0]["commonHashes"], d[0]["totalHashes"]) == (com, tot)
    assert d[0]["containment"] == cont and d[0]["jaccard"] == jac
    md = min(1.0, max(0.0, -np.log(2 * jac / (1 + jac)) / 21))
    assert abs(d[0]["mashDistance"] - md) < 1e-15
    # pairwise: both ordered pairs, reference-major (main.rs:321-331); max-dist filters
    dp = json.loads(finch("dist", "-p", str(tmp_path / "a.fa.sk"), str(b)).stdout)
    assert [(x["query"], x["reference"]) for x in dp] == [(str(b), str(a)), (str(a), str(b))]
    assert json.loads(finch("dist", "-p", "-d", "0.0", str(tmp_path / "a.fa.sk"), str(b)).stdout) == []
    # --old-dist: old_distance (distance.rs:136-157) = |Q n R| over the whole reference
    do = json.loads(finch("dist", "--old-dist", str(tmp_path / "a.fa.sk"), str(b)).stdout)
    ocont, ojac, ocom, otot = o.old_distance(ha, hb)
    assert (do[0]["commonHashes"], do[0]["totalHashes"], do[0]["containment"], do[0]["jaccard"]) == (ocom, otot, ocont, ojac)
    assert abs(do[0]["mashDistance"] - o.mash_distance(ojac, 21)) < 1e-15
    # queries by name
    dq = json.loads(finch("dist", "-q", str(b), "--", str(tmp_path / "a.fa.sk"), str(b)).stdout)
    assert [(x["query"], x["reference"]) for x in dq] == [(str(b), str(a))]


@pytest.mark.gpu
def test_dist_pairwise_cut_equals_pair_list(finch, tmp_path):
    """`finch dist -p -d X` over a collection: the tiled all-pairs kernel with the device-side max_distance cut
    (fb2_dist_all_pairs_cut) prints byte for byte what the pair-list kernel + host filter prints
    (calc_sketch_distances, cli/src/main.rs:315-334: reference-major order, equal-by-value pairs skipped)."""
    rng = np.random.default_rng(17)
    base = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=20000)
    files = []
    for i in range(14):
        g = base.copy() if i % 2 else rng.choice(np.frombuffer(b"ACGT", np.uint8), size=20000)
        pos = rng.choice(len(g), size=40 * (i + 1), replace=False)
        g[pos] = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=len(pos))
        p = tmp_path / f"g{i}.fa"
        with open(p, "wb") as fh:
            fh.write(b">g%d\n" % i)
            for a in range(0, len(g), 60):
                fh.write(g[a:a + 60].tobytes() + b"\n")
        files.append(str(p))
    files.append(files[3])                       # the same file twice: equal by value, skipped (Q12)
    finch("sketch", "-n", "300", "-o", str(tmp_path / "all"), *files)
    sk = str(tmp_path / "all.sk")
    for extra in (["-p"], ["-p", "-d", "0.05"], ["-p", "-d", "0.0"], [], ["-d", "0.2"]):
        fast = finch("dist", *extra, sk).stdout
        slow = finch("dist", *extra, sk, env={"FINCH_DIST_PAIR_LIST": "1"}).stdout
        assert fast == slow, extra
    d = json.loads(finch("dist", "-p", "-d", "0.05", sk).stdout)
    assert 0 < len(d) < 15 * 14 and all(x["mashDistance"] <= 0.05 for x in d)
    assert all(x["query"] != x["reference"] for x in d)


# ---- `.bsk` / `.msh`: Cap'n Proto sketch files (finch_rs_b200/host/sketch_capnp.hpp) --------------------
def _capnp_words(raw):
    """capnp::serialize framing -> list of segments (lists of u64 words)."""
    import struct
    nseg = struct.unpack_from("<I", raw, 0)[0] + 1
    sizes = struct.unpack_from("<%dI" % nseg, raw, 4)
    pos = (4 + 4 * nseg + 7) // 8 * 8
    segs = []
    for n in sizes:
        segs.append(list(struct.unpack_from("<%dQ" % n, raw, pos)))
        pos += 8 * n
    assert pos == len(raw)
    return segs


class _Walk:
    """An independent reader of the published wire format, just enough for single-segment messages: the test's
    yardstick for the C++ writer (the reference has no golden .bsk / .msh files)."""
    def __init__(self, words):
        self.w = words

    def struct(self, p):
        v = self.w[p]
        assert v & 3 == 0, "struct pointer expected"
        off = ((v & 0xFFFFFFFF) ^ 0x80000000) - 0x80000000 >> 2
        return p + 1 + off, (v >> 32) & 0xFFFF, v >> 48

    def list(self, p):
        v = self.w[p]
        if v == 0:
            return 0, 0, 0, 0, 0
        assert v & 3 == 1, "list pointer expected"
        off = ((v & 0xFFFFFFFF) ^ 0x80000000) - 0x80000000 >> 2
        start, code, count = p + 1 + off, (v >> 32) & 7, v >> 35
        if code == 7:
            tag = self.w[start]
            return start + 1, 7, (tag & 0xFFFFFFFF) >> 2, (tag >> 32) & 0xFFFF, tag >> 48
        return start, code, count, 0, 0

    def bytes(self, p, text):
        start, code, count, _, _ = self.list(p)
        if count == 0:
            return b""
        assert code == 2
        import struct
        raw = struct.pack("<%dQ" % ((count + 7) // 8), *self.w[start:start + (count + 7) // 8])[:count]
        if text:
            assert raw[-1] == 0
            raw = raw[:-1]
        return raw


def test_bsk_round_trip_and_wire_layout(finch, tmp_path):
    import struct
    f = tmp_path / "in.sk"
    f.write_text(json.dumps(SK))
    finch("sketch", "-b", "-o", str(tmp_path / "out"), str(f))
    raw = (tmp_path / "out.bsk").read_bytes()
    # JSON -> .bsk -> JSON: nothing lost (filters, k-mers, counts, the u64::MAX hash, the quoted name)
    back = json.loads(finch("sketch", "-O", str(tmp_path / "out.bsk")).stdout)
    assert back == SK
    # the bytes, read by the test's own walker with the layouts capnpc generated for finch.capnp
    segs = _capnp_words(raw)
    assert len(segs) == 1
    w = _Walk(segs[0])
    root, d, p = w.struct(0)
    assert (d, p) == (0, 1)                                           # Multisketch
    sk0, code, n, sd, sp = w.list(root)
    assert (code, n, sd, sp) == (7, 1, 2, 5)                          # List(Sketch), one element of 2 data + 5 pointer words
    assert w.w[sk0] == 405 and w.w[sk0 + 1] == 339                    # seqLength, numValidKmers
    assert w.bytes(sk0 + 2, True) == 'a "q"\n.fa'.encode() and w.bytes(sk0 + 3, True) == b""
    h0, code, n, hd, hp = w.list(sk0 + 4)
    assert (code, n, hd, hp) == (7, 3, 2, 2)                          # List(KmerCount)
    for j in range(3):
        h = h0 + 4 * j
        assert w.w[h] == int(SK["sketches"][0]["hashes"][j])
        count, extra = w.w[h + 1] & 0xFFFFFFFF, w.w[h + 1] >> 32
        assert count == SK["sketches"][0]["counts"][j] and extra == count // 2    # JSON reader: extra_count = count / 2
        assert w.bytes(h + 2, False) == SK["sketches"][0]["kmers"][j].encode()
        assert w.w[h + 3] == 0                                        # label: None -> null pointer
    fp, d, p = w.struct(sk0 + 5)
    assert (d, p) == (4, 0)
    assert w.w[fp] & 1 == 1 and (w.w[fp] >> 32) == 2 and (w.w[fp + 1] & 0xFFFFFFFF) == 0xFFFFFFFF   # filtered, low 2, high None
    assert struct.unpack("<d", struct.pack("<Q", w.w[fp + 2]))[0] == 0.21 and struct.unpack("<d", struct.pack("<Q", w.w[fp + 3]))[0] == 0.1
    spp, d, p = w.struct(sk0 + 6)
    assert (d, p) == (5, 0)
    assert w.w[spp] & 0xFFFF == 0 and (w.w[spp] >> 16) & 0xFF == 21   # murmurHash3, k
    assert w.w[spp + 3] == 3                                          # finalSize (the JSON's sketchSize)


def test_msh_round_trip(finch, tmp_path):
    f = tmp_path / "in.sk"
    f.write_text(json.dumps(SK))
    finch("sketch", "-B", "-o", str(tmp_path / "out"), str(f))
    raw = (tmp_path / "out.msh").read_bytes()
    back = json.loads(finch("sketch", "-O", str(tmp_path / "out.msh")).stdout)
    # what the Mash schema keeps (serialization/mash.rs:60-132, quirk Q10): hashes and counts; no k-mers, no filters,
    # kmers_to_sketch = final_size = 0
    want = json.loads(json.dumps(SK))
    want["sketchSize"] = 0
    want["sketches"][0]["filters"] = {}
    want["sketches"][0]["kmers"] = ["", "", ""]
    assert back == want
    w = _Walk(_capnp_words(raw)[0])
    root, d, p = w.struct(0)
    assert (d, p) == (3, 4)                                           # MinHash
    assert w.w[root] & 0xFFFFFFFF == 21 and w.w[root] >> 32 == 21     # kmerSize, windowSize
    assert w.w[root + 1] & 0xFFFFFFFF == 3 and (w.w[root + 1] >> 32) & 7 == 1   # minHashesPerWindow, concatenated only
    assert w.w[root + 2] >> 32 == 42                                  # hashSeed 0, stored XOR its default 42
    assert w.bytes(root + 3 + 2, True) == b"ACGT"
    rl, d, p = w.struct(root + 3 + 3)
    r0, code, n, rd, rp = w.list(rl)
    assert (code, n, rd, rp) == (7, 1, 3, 7)
    assert w.w[r0 + 1] == 405 and w.w[r0 + 2] == 339
    hs, code, n, _, _ = w.list(r0 + 3 + 5)
    assert (code, n) == (5, 3) and w.w[hs:hs + 3] == [int(x) for x in SK["sketches"][0]["hashes"]]
    cs, code, n, _, _ = w.list(r0 + 3 + 6)
    assert (code, n) == (4, 3) and w.w[cs] == 1 | (3 << 32) and w.w[cs + 1] & 0xFFFFFFFF == 3


def test_capnp_reader_takes_segments_and_far_pointers(finch, tmp_path):
    """The reference's builder spreads large messages over several segments; its files then hold far pointers."""
    import struct
    f = tmp_path / "in.sk"
    f.write_text(json.dumps(SK))
    finch("sketch", "-b", "-o", str(tmp_path / "one"), str(f))
    words = _capnp_words((tmp_path / "one.bsk").read_bytes())[0]

    def frame(segs):
        hdr = struct.pack("<I%dI" % len(segs), len(segs) - 1, *[len(s) for s in segs])
        hdr += b"\0" * (-len(hdr) % 8)
        return hdr + b"".join(struct.pack("<%dQ" % len(s), *s) for s in segs)
    # (1) root = far pointer to a landing pad in segment 1 (the original message, whose word 0 is the root pointer)
    far = 2 | (0 << 3) | (1 << 32)
    (tmp_path / "far.bsk").write_bytes(frame([[far], words]))
    assert json.loads(finch("sketch", "-O", str(tmp_path / "far.bsk")).stdout) == SK
    # (2) double-far: pad in segment 2 = far pointer to the struct's first word (segment 1, word 1) + its tag word
    dfar = 2 | (1 << 2) | (0 << 3) | (2 << 32)
    pad = [2 | (1 << 3) | (1 << 32), (0 << 32) | (1 << 48)]
    (tmp_path / "dfar.bsk").write_bytes(frame([[dfar], words, pad]))
    assert json.loads(finch("sketch", "-O", str(tmp_path / "dfar.bsk")).stdout) == SK
    # damaged files are parse errors, not crashes
    raw = (tmp_path / "one.bsk").read_bytes()
    for bad in (raw[:40], raw[:8] + b"\xff" * 8 + raw[16:], b"", raw[:-9]):
        (tmp_path / "bad.bsk").write_bytes(bad)
        p = finch("info", str(tmp_path / "bad.bsk"), check=False)
        assert p.returncode == 1 and "Error parsing" in p.stderr, p.stderr


@pytest.mark.gpu
def test_finch_sketch_bin_and_msh(finch, tmp_path):                  # test_cli.rs:39-78
    for flag, ext in (("-b", ".bsk"), ("-B", ".msh")):
        out = tmp_path / ("q" + ext)
        p = subprocess.run([FINCH, "sketch", "--n-hashes", "10", flag, "-O", QUERY], capture_output=True)
        assert p.returncode == 0, p.stderr
        out.write_bytes(p.stdout)
        sk = json.loads(finch("sketch", "-O", str(out)).stdout)       # read_finch_file / read_mash_file, shown as JSON
        assert len(sk["sketches"]) == 1 and sk["kmer"] == 21
        assert len(sk["sketches"][0]["hashes"]) == 10
        if ext == ".bsk":
            assert sk["sketchSize"] == 10 and sk["sketches"][0]["kmers"] == GOLDEN_KMERS
    # dist reads them like .sk files
    a = tmp_path / "q.bsk"
    d = json.loads(finch("dist", str(a), str(tmp_path / "q.msh")).stdout)
    assert len(d) == 1 and d[0]["jaccard"] == 1.0 and d[0]["mashDistance"] == 0.0


@pytest.mark.gpu
def test_finch_sketch_type_none(finch, oracle):                      # cli.rs:336, counts.rs
    sk = json.loads(finch("sketch", "--sketch-type", "none", "-k", "4", "-O", QUERY).stdout)
    assert (sk["kmer"], sk["hashType"], sk["hashBits"], sk["hashSeed"], sk["sketchSize"], sk["scale"]) == (4, "None", 0, 0, 256, None)
    o = oracle.AllCountsSketcher(4)
    rc, _, recs = oracle.parse_fastx(open(QUERY, "rb").read())
    for r in recs:
        o.process(r)
    want = o.to_vec()
    s = sk["sketches"][0]
    assert [int(h) for h in s["hashes"]] == [int(h) for h in want["hashes"]]
    assert s["counts"] == [int(c) for c in want["counts"]] and s["kmers"] == [k.decode() for k in want["kmers"]]
    assert s["seqLength"] == 0 and s["numValidKmers"] == o.total_bases_and_kmers()[1]

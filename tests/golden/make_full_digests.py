#!/usr/bin/env python
"""Run the CPU oracle (oracle/, the restatement of the reference's CPU path) over the EXACT inputs of the
BASELINE.json configs at their full stated size and commit sha256 digests of the resulting sketches:

    python tests/golden/make_full_digests.py [--only c1,c2,c3,c4,c5] [--procs 8]

writes tests/golden/full_digests.json.  bench.py / bench_configs.py compare the GPU results of every rank and
every config with these digests (no oracle run on the GPU box is needed for full-size bit-exactness).
CPU-only; about 10 minutes on 8 cores.  Inputs are defined in tools/workloads.py.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
OUT = os.path.join(ROOT, "tests", "golden", "full_digests.json")


def _sketch(data, sp, fp):
    import oracle
    import workloads as W
    t0 = time.time()
    rc, sk = oracle.sketch_stream(data, sp, fp, kmers_array=True)
    assert rc == oracle.OK, rc
    d = W.sketch_digest(sk["hashes"], sk["counts"], sk["extras"], sk["kmers"], sk["seq_length"], sk["num_valid_kmers"])
    d["oracle_s"] = round(time.time() - t0, 2)
    return d


def task(t):
    import oracle
    import workloads as W
    kind = t[0]
    if kind == "c2":       # 10 M x 150 bp FASTQ, `finch sketch -k 21 -n 1000 -f`
        rank, reads = t[1], t[2]
        buf, nbytes, nbases = W.c2_fastq(rank, reads, threads=1)
        d = _sketch(buf, oracle.mash_params(W.N_HASHES * W.OVERSKETCH, W.N_HASHES, False, W.K_C2, 0),
                    oracle.make_filter(True, (None, None), 1.0 * W.K_C2 / 100.0, 0.1))
        d.update(input_bytes=int(nbytes), input_sha256_head=None)
        return f"c2/reads={reads}/rank={rank}", d
    if kind == "c1":
        data = W.c1_fasta()
        return "c1", _sketch(data, oracle.mash_params(200000, 1000, False, 21, 0), oracle.make_filter(None, (None, None), 0.21, 0.1))
    if kind == "c3":
        i = t[1]
        data = W.c3_fasta(i)
        return f"c3/file={i}", _sketch(data, oracle.mash_params(200000, 1000, False, 21, 0),
                                       oracle.make_filter(None, (None, None), 0.21, 0.1))
    if kind == "c4":
        nb = t[1]
        data = W.c4_fasta(nb)
        return f"c4/bases={nb}", _sketch(data, oracle.scaled_params(1000, 31, 0.001, 0),
                                         oracle.make_filter(None, (None, None), 0.31, 0.1))
    if kind == "c5":       # the oracle's literal merge loop on the parity sample of pairs
        import hashlib
        n_sk, n_pairs = t[1], t[2]
        q, r = W.c5_sample_pairs(n_sk, n_pairs)
        need = np.unique(np.concatenate([q, r]))
        rows = {int(i): W.c5_rows(int(i), 1, n_sk)[0] for i in need}
        out = np.zeros((n_pairs, 2), np.uint64)
        t0 = time.time()
        for p in range(n_pairs):
            _, _, com, tot = oracle.raw_distance(rows[int(q[p])], rows[int(r[p])], 0.0)
            out[p] = (com, tot)
        # the per-pair values too (bench_configs.py checks the hit list of the cut kernel against them)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"c5_sample_n{n_sk}.npz"),
                            common=out[:, 0].astype(np.uint16), total=out[:, 1].astype(np.uint16))
        return f"c5/n={n_sk}/pairs={n_pairs}", {"sha256_common_total": hashlib.sha256(out.tobytes()).hexdigest(),
                                                 "sum_common": int(out[:, 0].sum()), "oracle_s": round(time.time() - t0, 2)}
    raise ValueError(kind)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="c2,c4,c1,c3,c5")
    ap.add_argument("--procs", type=int, default=8)
    args = ap.parse_args()
    import oracle
    import synth
    import workloads as W
    oracle.build()
    synth.build()
    only = args.only.split(",")
    tasks = []
    if "c4" in only:
        tasks += [("c4", W.C4_BASES), ("c4", 300_000_000)]
    if "c2" in only:
        tasks += [("c2", r, W.C2_READS) for r in range(8)]
        tasks += [("c2", 0, 1_000_000)]      # --reads 1000000: quick runs of bench.py
    if "c5" in only:
        tasks += [("c5", W.C5_SKETCHES, 100_000), ("c5", 8192, 100_000)]
    if "c1" in only:
        tasks += [("c1",)]
    if "c3" in only:
        tasks += [("c3", i) for i in range(W.C3_FILES)]
    digests = json.load(open(OUT)) if os.path.exists(OUT) else {}
    t0 = time.time()
    with mp.get_context("spawn").Pool(args.procs, maxtasksperchild=4) as pool:
        for key, d in pool.imap_unordered(task, tasks):
            digests[key] = d
            print(f"[{time.time() - t0:7.1f}s] {key}: {d}", flush=True)
            json.dump(dict(sorted(digests.items())), open(OUT, "w"), indent=0)
    print("wrote", OUT)


if __name__ == "__main__":
    main()

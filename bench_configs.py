#!/usr/bin/env python
"""bench_configs.py -- the other BASELINE.json configs (C1, C3, C4, C5) on ONE GPU, with the CPU
oracle port timed beside them on a bounded sample and a bit-exactness check each.

    python bench_configs.py [--quick] [--out profiles/r01_configs.json]

bench.py remains the driver's contract (config C2); this script fills BASELINE.md's table.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def oracle_sketch_stream(oracle, data, sp, fp):
    t0 = time.perf_counter()
    rc, sk = oracle.sketch_stream(data, sp, fp)
    return rc, sk, time.perf_counter() - t0


def same(sk, osk):
    return bool(np.array_equal(sk.hashes_u64, osk["hashes"]) and np.array_equal(sk.counts, osk["counts"])
                and np.array_equal(sk.extra_counts, osk["extras"])
                and (sk.seq_length, sk.num_valid_kmers) == (osk["seq_length"], osk["num_valid_kmers"]))


def gpu_time(torch, fn, iters, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = None
    for _ in range(iters):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    import finch_rs_b200 as fb
    import oracle
    oracle.build()
    assert torch.cuda.is_available() and fb.lib().fb2_device_count() > 0, "needs a GPU (no CPU fallback)"
    dev = torch.device("cuda", 0)
    rows = []

    # ---- C1: finch sketch on one 5 Mbp FASTA, k=21, n=1000 (FASTA => filter off; CLI heap 200 000) ----
    data = fb.synth_fasta(5_000_000, n_records=1, line_width=80, seed=1)
    sp = fb.SketchParams.from_cli("mash", n_hashes=1000, kmer_length=21)            # heap 200 000 (auto filter)
    fp = fb.FilterParams(None, (None, None), 0.21, 0.1)
    host = torch.from_numpy(data).pin_memory()
    sk_h = sp.create_sketcher()

    def c1():
        sk_h.reset()
        sk_h.feed_fastx_ptr(host.data_ptr(), host.numel(), final=True)
        return sk_h.sketch("c1.fa", fp)
    dt, sk = gpu_time(torch, c1, 20)
    rc, osk, odt = oracle_sketch_stream(oracle, data.tobytes(), oracle.mash_params(200000, 1000, False, 21, 0),
                                        oracle.make_filter(None, (None, None), 0.21, 0.1))
    rows.append({"config": "C1 5 Mbp FASTA k21 n1000 (heap 200000, filter off)", "gpu_e2e_ms": dt * 1e3,
                 "gbases_per_s_e2e": 5e6 / dt / 1e9, "cpu_port_s": odt, "cpu_threads": 1,
                 "cpu_gbases_per_s": 5e6 / odt / 1e9, "bit_exact": same(sk, osk)})
    print(json.dumps(rows[-1]), flush=True)

    # ---- C3: batch of ~5 Mbp FASTAs through one re-used handle (per-GPU share of the 1024-file batch) ----
    nfiles = 8 if args.quick else 32
    files = [fb.synth_fasta(int(4.5e6 + (i * 7919 % 1000) * 1e3), n_records=1 + i % 3, line_width=80, seed=1000 + i)
             for i in range(nfiles)]
    pinned = [torch.from_numpy(f).pin_memory() for f in files]
    total_bases = sum(int(4.5e6 + (i * 7919 % 1000) * 1e3) for i in range(nfiles))

    def c3():
        out = []
        for t in pinned:
            sk_h.reset()
            sk_h.feed_fastx_ptr(t.data_ptr(), t.numel(), final=True)
            out.append(sk_h.sketch("f", fp))
        return out
    dt, sks = gpu_time(torch, c3, 3, warmup=1)
    t0 = time.perf_counter()
    ok = True
    for i in (0, nfiles - 1):
        rc, osk = oracle.sketch_stream(files[i].tobytes(), oracle.mash_params(200000, 1000, False, 21, 0),
                                       oracle.make_filter(None, (None, None), 0.21, 0.1))
        ok = ok and same(sks[i], osk)
    odt = (time.perf_counter() - t0) / 2
    rows.append({"config": f"C3 batch of {nfiles} x ~5 Mbp FASTA (one GPU's share), k21 n1000", "gpu_e2e_ms": dt * 1e3,
                 "ms_per_file": dt * 1e3 / nfiles, "gbases_per_s_e2e": total_bases / dt / 1e9,
                 "cpu_port_s_per_file": odt, "cpu_threads": 1, "cpu_gbases_per_s": 5e6 / odt / 1e9, "bit_exact": ok})
    print(json.dumps(rows[-1]), flush=True)
    sk_h.close()

    # ---- C3 through sketch_files (lib.rs:29-49): files on tmpfs, worker handles overlap on the GPU ----
    import shutil
    import tempfile
    tdir = tempfile.mkdtemp(prefix="fb2c3_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        nrep = 1 if args.quick else 4                  # 128 files = one GPU's share of the 1024-file batch
        paths = []
        for rep in range(nrep):
            for i, f in enumerate(files):
                pth = os.path.join(tdir, f"g{rep}_{i}.fa")
                f.tofile(pth)
                paths.append(pth)
        for workers in (1, 8, 16):
            os.environ["FB2_FILE_WORKERS"] = str(workers)
            fb.sketch_files(paths[:nfiles], sp, fp)    # warm-up (handles, page cache)
            t0 = time.perf_counter()
            sks_f = fb.sketch_files(paths, sp, fp)
            dtf = time.perf_counter() - t0
            okf = all(np.array_equal(sks_f[j].hashes_u64, sks[j % nfiles].hashes_u64) and
                      np.array_equal(sks_f[j].counts, sks[j % nfiles].counts) for j in range(len(paths)))
            rows.append({"config": f"C3 sketch_files({len(paths)} x ~5 Mbp FASTA on tmpfs), {workers} worker handle(s), 1 GPU",
                         "gpu_e2e_ms": dtf * 1e3, "ms_per_file": dtf * 1e3 / len(paths),
                         "gbases_per_s_e2e": total_bases * nrep / dtf / 1e9, "bit_exact": bool(okf),
                         "bit_exact_on": "every file vs the single-handle result (itself checked against the oracle)"})
            print(json.dumps(rows[-1]), flush=True)
        os.environ.pop("FB2_FILE_WORKERS", None)
    finally:
        shutil.rmtree(tdir, ignore_errors=True)

    # ---- C4: 3 Gbp FASTA (24 records, 60-col, 2% lowercase, 1% N), k=31, scaled 0.001, n=1000 ----------
    nb = 300_000_000 if args.quick else 3_000_000_000
    big = fb.synth_fasta(nb, n_records=24, line_width=60, lower_frac=0.02, n_frac=0.01, seed=4)
    sp4 = fb.SketchParams.scaled(1000, 31, 0.001, 0)
    fp4 = fb.FilterParams(None, (None, None), 0.31, 0.1)
    hbig = torch.from_numpy(big).pin_memory()
    dbig = hbig.to(dev)
    sk4 = sp4.create_sketcher()

    def c4_res():
        sk4.reset()
        sk4.feed_device(dbig.data_ptr(), dbig.numel(), final=True)
        return sk4.sketch("c4.fa", fp4)

    def c4_e2e():
        sk4.reset()
        sk4.feed_fastx_ptr(hbig.data_ptr(), hbig.numel(), final=True)
        return sk4.sketch("c4.fa", fp4)
    dt_res, skr = gpu_time(torch, c4_res, 3, warmup=1)
    dt_e2e, ske = gpu_time(torch, c4_e2e, 2, warmup=1)
    assert np.array_equal(skr.hashes_u64, ske.hashes_u64) and np.array_equal(skr.counts, ske.counts)
    assert skr.seq_length >= nb and np.all(skr.hashes_u64[1:] > skr.hashes_u64[:-1])
    assert int(skr.hashes_u64.max()) <= (2**64 - 1) // 1000                      # scaled.rs:202-213 property
    sample = fb.synth_fasta(30_000_000, n_records=3, line_width=60, lower_frac=0.02, n_frac=0.01, seed=4)
    rc, osk, odt = oracle_sketch_stream(oracle, sample.tobytes(), oracle.scaled_params(1000, 31, 0.001, 0),
                                        oracle.make_filter(None, (None, None), 0.31, 0.1))
    gsk = fb.sketch_stream(sample, "s.fa", sp4, fp4)
    rows.append({"config": f"C4 {nb / 1e9:.1f} Gbp FASTA k31 scaled 0.001", "n_hashes": len(skr),
                 "gpu_resident_s": dt_res, "gpu_e2e_s": dt_e2e, "gbases_per_s_resident": nb / dt_res / 1e9,
                 "gbases_per_s_e2e": nb / dt_e2e / 1e9, "cpu_port_sample_s": odt, "cpu_threads": 1,
                 "cpu_gbases_per_s": 30e6 / odt / 1e9, "bit_exact": same(gsk, osk), "bit_exact_on": "30 Mbp sample"})
    print(json.dumps(rows[-1]), flush=True)
    sk4.close()
    del dbig, hbig, big

    # ---- C5: dist all-vs-all on n=1000 sketches (1000 clusters sharing 50-95% of their hashes) --------
    n_sk = 2048 if args.quick else 8192
    rng = np.random.default_rng(5)
    n_clusters = max(1, n_sk // 100)
    base = [np.unique(rng.integers(0, 2**63, size=1400, dtype=np.uint64))[:1000] for _ in range(n_clusters)]
    mat = np.zeros((n_sk, 1000), np.uint64)
    for i in range(n_sk):
        b = base[i % n_clusters]
        share = int(rng.integers(500, 951))
        own = np.unique(rng.integers(0, 2**63, size=1200 - share + 400, dtype=np.uint64))[:1000 - share]
        mat[i] = np.sort(np.concatenate([rng.choice(b, size=share, replace=False), own]))[:1000]
    lens = np.full(n_sk, 1000, np.uint32)
    q1 = n_sk
    obuf = np.zeros((q1 * n_sk, 3), np.uint32)     # result array allocated (and its pages touched) once
    dt, kms = 1e30, 1e30
    for _ in range(3):
        t0 = time.perf_counter()
        out = fb.dist_all_pairs(mat, lens, 0.0, 0, q1, out=obuf)
        dt = min(dt, time.perf_counter() - t0)
        kms = min(kms, fb.lib().fb2_dist_last_kernel_ms())
    npairs = q1 * n_sk
    pinned_out = torch.empty((q1 * n_sk, 3), dtype=torch.int32, pin_memory=True)   # same bits as uint32
    pbuf = pinned_out.numpy().view(np.uint32)
    dtp = 1e30
    for _ in range(3):
        t0 = time.perf_counter()
        outp = fb.dist_all_pairs(mat, lens, 0.0, 0, q1, out=pbuf)
        dtp = min(dtp, time.perf_counter() - t0)
    same_pinned = bool(np.array_equal(outp, out))
    # CPU port on a bounded sample of pairs
    nq = 4
    t0 = time.perf_counter()
    ok = True
    for q in range(nq):
        for r in range(0, n_sk, 4):
            cont, jac, com, tot = oracle.raw_distance(mat[q], mat[r], 0.0)
            c, i, j = (int(v) for v in out[q, r])
            ok = ok and (c == com and i - c + j == tot)
    odt = time.perf_counter() - t0
    cpu_pairs = nq * (n_sk // 4)
    rows.append({"config": f"C5 dist all-vs-all {n_sk} x {n_sk} sketches of 1000 hashes (API call incl. H2D/D2H)",
                 "pairs": npairs, "gpu_s": dt, "pairs_per_s": npairs / dt, "kernel_ms": kms,
                 "kernel_pairs_per_s": npairs / (kms * 1e-3), "pinned_out_s": dtp, "pinned_out_pairs_per_s": npairs / dtp,
                 "pinned_out_identical": same_pinned, "cpu_port_pairs_per_s": cpu_pairs / odt,
                 "cpu_threads": 1, "bit_exact": ok, "bit_exact_on": f"{cpu_pairs} sampled pairs"})
    print(json.dumps(rows[-1]), flush=True)
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench_configs.py -- the BASELINE.json configs other than the driver's bench line, at their STATED size, through
the C ABI, on 1..N GPUs of this process, each result checked bit for bit against the committed CPU-oracle digests
(tests/golden/full_digests.json, made by tests/golden/make_full_digests.py on the same seeded inputs).

    python bench_configs.py [--gpus N] [--only c1,c2split,c3,c4,c5] [--quick] [--out profiles/r02_configs.json]

  c1      configs[0]  one 5 Mbp FASTA, k=21, n=1000 (CLI defaults: heap 200 000, FASTA => filter off)
  c2split configs[1]  the 10 M x 150 bp FASTQ cut over N GPUs by fb2_sketch_stream_multi (exact peer-memory merge)
  c3      configs[2]  1024 x ~5 Mbp FASTA files, fb2_sketch_files_multi over N GPUs (LPT), files on tmpfs
  c4      configs[3]  3 Gbp FASTA, k=31, scaled 0.001: resident / end to end on 1 GPU, and cut over N GPUs
  c5      configs[4]  dist all-vs-all on 100 000 sketches x 1000 hashes with the max-distance cut on the device,
                      query rows cut over N GPUs; parity on 100 000 sampled pairs vs the oracle's literal merge loop
bench.py remains the driver's contract (configs[1], one file per GPU); this script fills BASELINE.md's table.
No oracle code runs here: full-size bit-exactness comes from the committed digests.
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import workloads as W  # noqa: E402


def digest_of(sk, k):
    return W.sketch_digest(sk.hashes_u64, sk.counts, sk.extra_counts, sk.kmers[:, :k], sk.seq_length, sk.num_valid_kmers)


def matches(d, want):
    return want is not None and all(d[x] == want[x] for x in ("n", "sha256", "seq_length", "num_valid_kmers"))


def timed(torch, fn, iters, warmup=1):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts, out = [], None
    for _ in range(iters):
        t0 = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), float(min(ts)), out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--only", default="c1,c3,c2split,c4,c5")
    ap.add_argument("--quick", action="store_true", help="reduced sizes (no digests exist for those: bit_exact null)")
    ap.add_argument("--out", default="")
    ap.add_argument("--max-dist", type=float, default=0.05)
    ap.add_argument("--c3-workers", default="4,8,16", help="worker handles per GPU to try for C3")
    ap.add_argument("--c3-files", type=int, default=0, help="C3 with the first N of its 1024 files (digests exist for all)")
    args = ap.parse_args()
    only = args.only.split(",")
    import torch
    import finch_rs_b200 as fb
    W.synth.build()
    assert torch.cuda.is_available() and fb.lib().fb2_device_count() > 0, "needs a GPU (no CPU fallback)"
    G = min(args.gpus, fb.lib().fb2_device_count())
    digests = json.load(open(os.path.join(ROOT, "tests", "golden", "full_digests.json")))
    rows = []

    def emit(row):
        row["n_gpus"] = row.get("n_gpus", G)
        rows.append(row)
        print(json.dumps(row), flush=True)
        if args.out:
            json.dump(rows, open(args.out, "w"), indent=1)

    fp_auto = fb.FilterParams(None, (None, None), 0.21, 0.1)
    sp_cli = fb.SketchParams.from_cli("mash", n_hashes=1000, kmer_length=21)            # heap 200 000 (auto filter)

    # ---- C1 ------------------------------------------------------------------------------------------------
    if "c1" in only:
        data = W.c1_fasta()
        host = torch.from_numpy(data).pin_memory()
        med, best, sk = timed(torch, lambda: fb.sketch_stream(host.numpy(), "c1.fa", sp_cli, fp_auto), 20, 3)
        emit({"config": "C1 one 5 Mbp FASTA, k21 n1000 (CLI defaults: heap 200000, FASTA => filter off), fb2_sketch_stream",
              "n_gpus": 1, "ms": med * 1e3, "ms_best": best * 1e3, "gbases_per_s_e2e": 5e6 / med / 1e9,
              "bit_exact": matches(digest_of(sk, 21), digests.get("c1")), "bit_exact_on": "whole result vs oracle digest"})

    # ---- C3 ------------------------------------------------------------------------------------------------
    if "c3" in only:
        nfiles = 64 if args.quick else (args.c3_files or W.C3_FILES)
        tdir = tempfile.mkdtemp(prefix="fb2c3_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        try:
            paths = [os.path.join(tdir, f"g{i:04d}.fa") for i in range(nfiles)]

            def gen(lo, hi):
                for i in range(lo, hi):
                    W.c3_fasta(i).tofile(paths[i])
            nt = min(32, os.cpu_count() or 1)
            per = (nfiles + nt - 1) // nt
            ths = [threading.Thread(target=gen, args=(t * per, min(nfiles, (t + 1) * per))) for t in range(nt)]
            [t.start() for t in ths]
            [t.join() for t in ths]
            total_bases = sum(W.c3_nbases(i) for i in range(nfiles))
            for workers in ((8,) if args.quick else tuple(int(x) for x in args.c3_workers.split(",")) + (0,)):
                if workers:
                    os.environ["FB2_FILE_WORKERS"] = str(workers)
                else:                                    # the library's own choice (by host cores and GPUs)
                    os.environ.pop("FB2_FILE_WORKERS", None)
                # warm-up: every worker's handle exists (a worker creates its handle before it takes a file, so the call needs
                # at least `workers` files per GPU), page cache.  Creating handles is the expensive part of a FIRST call:
                # tools/c3_trace.py, tools/alloc_cost.cu
                fb.sketch_files(paths[:min(nfiles, 4 * (workers or 16) * G)], sp_cli, fp_auto, ngpus=G)
                t0 = time.perf_counter()
                sks = fb.sketch_files(paths, sp_cli, fp_auto, ngpus=G)
                dt = time.perf_counter() - t0
                ok = all(matches(digest_of(sks[i], 21), digests.get(f"c3/file={i}")) for i in range(nfiles))
                emit({"config": f"C3 fb2_sketch_files_multi({nfiles} x ~5 Mbp FASTA on tmpfs), {workers or 'default'} worker handles per GPU",
                      "s": dt, "ms_per_file": dt * 1e3 / nfiles, "gbases_per_s_e2e": total_bases / dt / 1e9,
                      "bit_exact": bool(ok), "bit_exact_on": f"every one of the {nfiles} files vs its oracle digest"})
            os.environ.pop("FB2_FILE_WORKERS", None)
        finally:
            shutil.rmtree(tdir, ignore_errors=True)
            fb.lib().fb2_sketch_files_release_pool()

    # ---- C2 cut over the GPUs ---------------------------------------------------------------------------------
    if "c2split" in only:
        reads = 1_000_000 if args.quick else W.C2_READS
        need = W.synth.fastq_nbytes(reads, W.READ_LEN, 0)
        host = torch.empty(need, dtype=torch.uint8, pin_memory=True)
        _, nbytes, nbases = W.c2_fastq(0, reads, out_ptr=host.data_ptr())
        sp = fb.SketchParams.from_cli("mash", n_hashes=1000, kmer_length=21, filters_enabled=True)
        fp = fb.FilterParams(True, (None, None), 0.21, 0.1)
        for g in sorted({1, G}):
            med, best, sk = timed(torch, lambda: fb.sketch_stream_multi_ptr(host.data_ptr(), nbytes, "c2.fq", sp, fp, g), 5, 2)
            emit({"config": f"C2 one FASTQ {reads} x 150 bp cut over {g} GPU(s): fb2_sketch_stream_multi from pinned host memory",
                  "n_gpus": g, "ms": med * 1e3, "ms_best": best * 1e3, "gbases_per_s_e2e": nbases / med / 1e9,
                  "h2d_gbs": nbytes / med / 1e9,
                  "bit_exact": matches(digest_of(sk, 21), digests.get(f"c2/reads={reads}/rank=0")),
                  "bit_exact_on": "whole result vs oracle digest of the full file"})
        # the same FASTQ as a FILE (tmpfs) through fb2_sketch_files: what `finch sketch reads.fq` does
        if os.path.isdir("/dev/shm") and not args.quick:
            path = os.path.join(tempfile.mkdtemp(prefix="fb2c2_", dir="/dev/shm"), "c2.fq")
            try:
                host.numpy().tofile(path)
                med, best, sks = timed(torch, lambda: fb.sketch_files([path], sp, fp), 3, 1)
                emit({"config": f"C2 the same FASTQ as one file on tmpfs: fb2_sketch_files (parallel pread, double-buffered pieces)",
                      "n_gpus": 1, "ms": med * 1e3, "gbases_per_s_e2e": nbases / med / 1e9, "file_gbs": nbytes / med / 1e9,
                      "bit_exact": matches(digest_of(sks[0], 21), digests.get(f"c2/reads={reads}/rank=0")),
                      "bit_exact_on": "whole result vs oracle digest of the full file"})
            finally:
                shutil.rmtree(os.path.dirname(path), ignore_errors=True)
        del host
        fb.lib().fb2_sketch_files_release_pool()

    # ---- C4 ------------------------------------------------------------------------------------------------
    if "c4" in only:
        nb = 300_000_000 if args.quick else W.C4_BASES
        big = W.c4_fasta(nb)
        sp4 = fb.SketchParams.scaled(1000, 31, 0.001, 0)
        fp4 = fb.FilterParams(None, (None, None), 0.31, 0.1)
        hbig = torch.from_numpy(big).pin_memory()
        del big
        want = digests.get(f"c4/bases={nb}")
        dbig = hbig.to("cuda:0")
        sk4 = sp4.create_sketcher()

        def c4_res():
            sk4.reset()
            sk4.feed_device(dbig.data_ptr(), dbig.numel(), final=True)
            return sk4.sketch("c4.fa", fp4)

        def c4_e2e():
            sk4.reset()
            sk4.feed_fastx_ptr(hbig.data_ptr(), hbig.numel(), final=True)
            return sk4.sketch("c4.fa", fp4)
        med_r, best_r, skr = timed(torch, c4_res, 3, 1)
        ok_r = matches(digest_of(skr, 31), want)
        del skr
        med_e, best_e, ske = timed(torch, c4_e2e, 3, 1)
        emit({"config": f"C4 {nb / 1e9:.1f} Gbp FASTA k31 scaled 0.001, 1 GPU", "n_gpus": 1, "n_hashes": len(ske),
              "resident_s": med_r, "e2e_s": med_e, "gbases_per_s_resident": nb / med_r / 1e9, "gbases_per_s_e2e": nb / med_e / 1e9,
              "bit_exact": bool(ok_r and matches(digest_of(ske, 31), want)), "bit_exact_on": "whole result (both paths) vs oracle digest of the full file"})
        del ske
        sk4.close()
        del dbig
        torch.cuda.empty_cache()
        if G > 1:
            med, best, skm = timed(torch, lambda: fb.sketch_stream_multi_ptr(hbig.data_ptr(), hbig.numel(), "c4.fa", sp4, fp4, G), 3, 1)
            emit({"config": f"C4 {nb / 1e9:.1f} Gbp FASTA k31 scaled 0.001 cut over {G} GPUs: fb2_sketch_stream_multi",
                  "e2e_s": med, "gbases_per_s_e2e": nb / med / 1e9, "h2d_gbs": hbig.numel() / med / 1e9,
                  "bit_exact": matches(digest_of(skm, 31), want), "bit_exact_on": "whole result vs oracle digest of the full file"})
            del skm
        del hbig
        fb.lib().fb2_sketch_files_release_pool()

    # ---- C5 ------------------------------------------------------------------------------------------------
    if "c5" in only:
        n_sk = 8192 if args.quick else W.C5_SKETCHES
        nt = min(32, os.cpu_count() or 1)
        mat = np.empty((n_sk, W.C5_HASHES), np.uint64)
        per = (n_sk + nt - 1) // nt

        def genrows(lo, hi):
            if lo < hi:
                mat[lo:hi] = W.c5_rows(lo, hi - lo, n_sk)
        ths = [threading.Thread(target=genrows, args=(t * per, min(n_sk, (t + 1) * per))) for t in range(nt)]
        [t.start() for t in ths]
        [t.join() for t in ths]
        lens = np.full(n_sk, W.C5_HASHES, np.uint32)
        k, max_d = 21, args.max_dist
        npairs = n_sk * n_sk
        for g in sorted({1, G}) if not args.quick else (G,):
            t0 = time.perf_counter()
            hits = fb.dist_all_pairs_cut(mat, lens, k, max_d, 0.0, 0, n_sk, skip_self=True, ngpus=g, cap=1 << 25)
            dt = time.perf_counter() - t0
            kms = fb.lib().fb2_dist_last_kernel_ms()
            # parity on the committed sample of pairs: values from the oracle's literal merge loop
            q, r = W.c5_sample_pairs(n_sk, 100_000)
            gold = np.load(os.path.join(ROOT, "tests", "golden", f"c5_sample_n{n_sk}.npz"))
            gcom, gtot = gold["common"].astype(np.int64), gold["total"].astype(np.int64)
            key = hits["q"].astype(np.int64) * n_sk + hits["r"]
            skey = q.astype(np.int64) * n_sk + r
            pos = np.searchsorted(key, skey)
            found = (pos < len(key)) & (key[np.minimum(pos, len(key) - 1)] == skey)
            jac = np.where(gtot == 0, 1.0, gcom / np.maximum(gtot, 1))
            with np.errstate(divide="ignore"):
                md = np.clip(-np.log(2 * jac / (1 + jac)) / k, 0.0, 1.0)
            must = (md <= max_d) & (q != r)
            h = hits[np.minimum(pos, len(key) - 1)]
            ok = bool(np.all(found[must]))                                            # nothing the exact test keeps is missing
            ok = ok and bool(np.all((h["common"][found] == gcom[found]) & ((h["i"].astype(np.int64) - h["common"] + h["j"])[found] == gtot[found])))
            ok = ok and bool(np.all(md[found & ~must] <= max_d * 1.001 + 1e-6))       # extras only within the conservative margin
            # the pair-list kernel on the same sample: every value
            pb = None
            if g == 1:
                pb = np.zeros((len(q), 3), np.uint32)
                rc = fb.lib().fb2_dist_batch(mat.ctypes.data, lens.ctypes.data, n_sk, W.C5_HASHES, 0.0, q.ctypes.data, r.ctypes.data,
                                             len(q), pb.ctypes.data, -1)
                assert rc == 0, fb.lib().fb2_last_error()
            ok_batch = None if pb is None else bool(np.all(pb[:, 0] == gcom) and np.all(pb[:, 1].astype(np.int64) - pb[:, 0] + pb[:, 2] == gtot))
            emit({"config": f"C5 dist all-vs-all {n_sk} x {n_sk} sketches of 1000 hashes, max-dist {max_d} cut on the device, "
                            f"query rows over {g} GPU(s) (API call incl. matrix H2D / peer copies, hit sort, D2H)",
                  "n_gpus": g, "pairs": npairs, "hits": int(len(hits)), "s": dt, "pairs_per_s": npairs / dt, "kernel_ms_slowest_gpu": kms,
                  "kernel_pairs_per_s": npairs / (kms * 1e-3) if kms else None,
                  "bit_exact": ok, "bit_exact_on": f"{int(must.sum())} of 100000 sampled pairs pass the cut in the oracle: all present with the oracle's (common, total); no hit outside the margin",
                  "pair_list_kernel_bit_exact_on_100000_pairs": ok_batch})
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""C3-shaped files through fb2_sketch_files with FB2_TRACE_FILES=1: first call (handles created) and second call (pooled)
for several worker counts, spinning and polite waits.  usage: c3_trace.py [nfiles] [workers,comma] [modes 0,1,auto]"""
import os, shutil, sys, tempfile, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
os.environ["FB2_TRACE_FILES"] = "1"
import finch_rs_b200 as fb
import workloads as W

nfiles = int(sys.argv[1]) if len(sys.argv) > 1 else 256
workers = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "4,8,16,32").split(",")]
modes = (sys.argv[3] if len(sys.argv) > 3 else "0,1").split(",")
sp = fb.SketchParams.from_cli("mash", 1000, 21, 0)
fp = fb.FilterParams(None, (None, None), 0.21, 0.1)
tdir = tempfile.mkdtemp(prefix="fb2c3_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
try:
    paths = [os.path.join(tdir, f"g{i:04d}.fa") for i in range(nfiles)]
    def gen(lo, hi):
        for i in range(lo, hi):
            W.c3_fasta(i).tofile(paths[i])
    nt = min(32, os.cpu_count() or 1); per = (nfiles + nt - 1) // nt
    ths = [threading.Thread(target=gen, args=(t * per, min(nfiles, (t + 1) * per))) for t in range(nt)]
    [t.start() for t in ths]; [t.join() for t in ths]
    fb.sketch_files(paths[:4], sp, fp)   # context, kernels loaded
    fb.lib().fb2_sketch_files_release_pool()
    for mode in modes:
        if mode == "auto": os.environ.pop("FB2_POLITE_SYNC", None)
        else: os.environ["FB2_POLITE_SYNC"] = mode
        for w in workers:
            os.environ["FB2_FILE_WORKERS"] = str(w)
            for rep in range(3):
                t0 = time.perf_counter(); fb.sketch_files(paths, sp, fp); dt = time.perf_counter() - t0
                print(f"polite={mode} workers={w} call {rep}: {dt * 1e3:.1f} ms = {dt * 1e3 / nfiles:.3f} ms/file", flush=True)
            fb.lib().fb2_sketch_files_release_pool()
finally:
    shutil.rmtree(tdir, ignore_errors=True)

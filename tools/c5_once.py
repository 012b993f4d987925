"""C5-shaped dist (100 000 x 100 000 sketches of 1000 hashes, max-dist 0.05 cut) through fb2_dist_all_pairs_cut, three
calls; used under ncu for a launch list / captures of the inverted-index kernels.  usage: c5_once.py [n_sketches]"""
import os, sys, threading, time
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import numpy as np
import finch_rs_b200 as fb
import workloads as W
n_sk = int(float(sys.argv[1])) if len(sys.argv) > 1 else W.C5_SKETCHES
nt = min(32, os.cpu_count() or 1)
mat = np.empty((n_sk, W.C5_HASHES), np.uint64)
per = (n_sk + nt - 1) // nt
def genrows(lo, hi):
    if lo < hi:
        mat[lo:hi] = W.c5_rows(lo, hi - lo, n_sk)
ths = [threading.Thread(target=genrows, args=(t * per, min(n_sk, (t + 1) * per))) for t in range(nt)]
[t.start() for t in ths]; [t.join() for t in ths]
lens = np.full(n_sk, W.C5_HASHES, np.uint32)
for i in range(int(os.environ.get("C5_ITERS", "3"))):
    t0 = time.perf_counter()
    hits = fb.dist_all_pairs_cut(mat, lens, 21, 0.05, 0.0, 0, n_sk, skip_self=True, cap=1 << 25)
    dt = time.perf_counter() - t0
    print(f"iter {i}: {dt * 1e3:.1f} ms  {len(hits)} hits  kernels {fb.lib().fb2_dist_last_kernel_ms():.1f} ms  {n_sk * n_sk / dt:.3e} pairs/s", flush=True)

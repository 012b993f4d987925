"""Sketch one 5 Mbp FASTA a few times (C1); used under ncu to list the launches of one call."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import finch_rs_b200 as fb
sys.path.insert(0, "tools"); import synth
data = synth.synth_fasta(5_000_000, n_records=1, line_width=80, seed=1)
sp = fb.SketchParams.from_cli("mash", n_hashes=1000, kmer_length=21)
fp = fb.FilterParams(None, (None, None), 0.21, 0.1)
host = torch.from_numpy(data).pin_memory()
h = sp.create_sketcher()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for i in range(n):
    h.reset()
    t0 = time.perf_counter()
    h.feed_fastx_ptr(host.data_ptr(), host.numel(), final=True)
    t1 = time.perf_counter()
    sk = h.sketch("c1.fa", fp)
    t2 = time.perf_counter()
    print(f"iter {i}: feed {1e3*(t1-t0):.3f} ms, sketch {1e3*(t2-t1):.3f} ms, n={len(sk)}", flush=True)
    st = h.stats() if hasattr(h, "stats") else None
    if st: print(st)

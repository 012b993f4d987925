"""Sketch a 0.3 Gbp FASTA (C4 shape: k=31, scaled 0.001) from device memory; used under ncu for a launch list."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import finch_rs_b200 as fb
sys.path.insert(0, "tools"); import synth
nb = int(float(sys.argv[1])) if len(sys.argv) > 1 else 300_000_000
big = synth.synth_fasta(nb, n_records=24, line_width=60, lower_frac=0.02, n_frac=0.01, seed=4)
sp4 = fb.SketchParams.scaled(1000, 31, 0.001, 0)
fp4 = fb.FilterParams(None, (None, None), 0.31, 0.1)
d = torch.from_numpy(big).cuda()
h = sp4.create_sketcher()
for i in range(3):
    h.reset()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    h.feed_device(d.data_ptr(), d.numel(), final=True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    sk = h.sketch("c4.fa", fp4)
    t2 = time.perf_counter()
    print(f"iter {i}: feed {1e3*(t1-t0):.2f} ms  sketch {1e3*(t2-t1):.2f} ms  n={len(sk)}  {nb/(t2-t0)/1e9:.1f} Gbases/s", flush=True)
    print(h.stats())

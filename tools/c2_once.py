"""C2-shaped resident run, feed vs sketch() time split (where the end-of-stream cost goes); also the ncu target."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import numpy as np, torch
import finch_rs_b200 as fb
import workloads as W
n_reads = int(float(sys.argv[1])) if len(sys.argv) > 1 else W.C2_READS
buf, need, nbases = W.c2_fastq(0, n_reads)
d = torch.from_numpy(buf).cuda()
sp = fb.SketchParams.from_cli("mash", n_hashes=1000, kmer_length=21, filters_enabled=True)
fp = fb.FilterParams(True, (None, None), 0.21, 0.1)
h = sp.create_sketcher()
for i in range(4):
    h.reset()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    h.feed_device(d.data_ptr(), d.numel(), final=True)
    t1 = time.perf_counter()
    sk = h.sketch("c2.fq", fp)
    t2 = time.perf_counter()
    print(f"iter {i}: feed {1e3*(t1-t0):.2f} ms  sketch {1e3*(t2-t1):.2f} ms  total {1e3*(t2-t0):.2f}  n={len(sk)}  {nbases/(t2-t0)/1e9:.1f} Gbases/s", flush=True)

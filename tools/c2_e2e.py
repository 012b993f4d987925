"""C2 end to end from pinned host memory, a few steps: ms per step and H2D bytes (A/B of FB2_HOST_STRIP / threads)."""
import os, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import torch
import finch_rs_b200 as fb
import workloads as W
reads = int(float(sys.argv[1])) if len(sys.argv) > 1 else W.C2_READS
need = W.synth.fastq_nbytes(reads, W.READ_LEN, 0)
host = torch.empty(need, dtype=torch.uint8, pin_memory=True)
_, nbytes, nbases = W.c2_fastq(0, reads, out_ptr=host.data_ptr())
sp = fb.SketchParams.from_cli("mash", n_hashes=1000, kmer_length=21, filters_enabled=True)
fp = fb.FilterParams(True, (None, None), 0.21, 0.1)
h = sp.create_sketcher()
for i in range(5):
    h.reset()
    s0 = h.stats()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    h.feed_fastx_ptr(host.data_ptr(), nbytes, final=True)
    sk = h.sketch("c2.fq", fp)
    dt = time.perf_counter() - t0
    s1 = h.stats()
    print(f"strip={os.environ.get('FB2_HOST_STRIP','0')} threads={os.environ.get('FB2_STRIP_THREADS','-')} iter {i}: {dt*1e3:.2f} ms  "
          f"{nbases/dt/1e9:.1f} Gbases/s  h2d {(s1['h2d_bytes']-s0['h2d_bytes'])/1e9:.2f} GB  n={len(sk)}", flush=True)

"""C2 through fb2_sketch_stream from pinned host memory (the call the two-ended mode FB2_HOST_STRIP=2 lives behind)."""
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
for i in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sk = fb.sketch_stream_ptr(host.data_ptr(), nbytes, "c2.fq", sp, fp)
    dt = time.perf_counter() - t0
    st = fb.last_stream_stats()
    print(f"strip={os.environ.get('FB2_HOST_STRIP','0')} iter {i}: {dt*1e3:.2f} ms  {nbases/dt/1e9:.1f} Gbases/s  "
          f"h2d {st['h2d_bytes']/1e9:.2f} GB  n={len(sk)}", flush=True)

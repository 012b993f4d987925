"""The C2 FASTQ as ONE FILE on tmpfs through fb2_sketch_files (what `finch sketch reads.fq` does), for several
FB2_READ_THREADS.  usage: c2_file.py [reads] [threads,comma]"""
import os, shutil, sys, tempfile, time
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import numpy as np, torch
import finch_rs_b200 as fb
import workloads as W
reads = int(float(sys.argv[1])) if len(sys.argv) > 1 else W.C2_READS
threads = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "4,8,12,16").split(",")]
need = W.synth.fastq_nbytes(reads, W.READ_LEN, 0)
host = torch.empty(need, dtype=torch.uint8, pin_memory=True)
_, nbytes, nbases = W.c2_fastq(0, reads, out_ptr=host.data_ptr())
sp = fb.SketchParams.from_cli("mash", n_hashes=1000, kmer_length=21, filters_enabled=True)
fp = fb.FilterParams(True, (None, None), 0.21, 0.1)
path = os.path.join(tempfile.mkdtemp(prefix="fb2c2_", dir="/dev/shm"), "c2.fq")
try:
    host.numpy().tofile(path)
    del host
    for mode in os.environ.get("C2_MODES", "default").split(","):
        if mode == "default": os.environ.pop("FB2_HOST_STRIP", None)
        else: os.environ["FB2_HOST_STRIP"] = mode
        for t in threads:
            os.environ["FB2_READ_THREADS"] = str(t)
            for it in range(3):
                t0 = time.perf_counter(); sk = fb.sketch_files([path], sp, fp)[0]; dt = time.perf_counter() - t0
                print(f"strip={mode} read threads {t} iter {it}: {dt * 1e3:.1f} ms  {nbases / dt / 1e9:.2f} Gbases/s  file {nbytes / dt / 1e9:.1f} GB/s  n={len(sk)}", flush=True)
finally:
    shutil.rmtree(os.path.dirname(path), ignore_errors=True)

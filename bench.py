#!/usr/bin/env python
"""bench.py -- Gbases/s sketched (k=21, n=1000, 150 bp FASTQ, --filter on) on N B200s.

One "step" = one pass of the sketching hot path over one synthetic FASTQ of `--reads` x 150 bp
reads (BASELINE.json configs[1]: 10 M reads, 1.5 Gbases, ~3.1 GB of FASTQ text; 300x coverage of a
5 Mbp genome, 0.5 % substitution errors), with the CLI's resolved parameters for
`finch sketch -k 21 -n 1000 -f`:  MashSketcher heap 200 000, strand filter 0.1, err filter 0.21,
final size 1000, strict.

  value  : whole-job Gbases/s, FASTQ bytes already resident in HBM when the timed region starts
           (fb2_sketcher_feed_device), result read back to the host and filtered every step.
  e2e    : the same through the host-facing C-ABI call (fb2_sketcher_feed_fastx) from a PINNED
           host buffer, host->device copies inside the timed region.
  N > 1  : one rank per GPU (torchrun), every rank sketches its own file (independent files shard
           with no data-path collective) and one NCCL gather per step brings the finished
           1000-entry sketches to rank 0; weak scaling.
  --impl reference : the CPU path of the reference (oracle port; the Rust reference cannot be built
           here) on all host cores, a bounded sample per step.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K, N_HASHES, OVERSKETCH, READ_LEN = 21, 1000, 200, 150
JSON_OUT = sys.stdout   # replaced in main(): the real stdout, kept apart from library chatter on fd 1
GENOME_LEN, ERR_RATE = 5_000_000, 0.005
ERR_FILTER, STRAND_FILTER = 1.0 * K / 100.0, 0.1   # cli.rs:264-265, :142


def load_oracle():
    """CPU oracle = test/baseline infrastructure (only the cpu_baseline / reference legs use it)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    oracle.build()
    return oracle


def gen_fastq(fb, genome, n_reads, seed, first_id=0, out_ptr=None, threads=None):
    """Deterministic synthetic FASTQ, generated in parallel slices (per-read RNG streams)."""
    threads = threads or min(32, os.cpu_count() or 1)
    need = fb.fastq_nbytes(n_reads, READ_LEN, first_id)
    if out_ptr is None:
        buf = np.empty(need, np.uint8)
        base = buf.ctypes.data
    else:
        buf, base = None, out_ptr
    per = (n_reads + threads - 1) // threads
    jobs, off = [], 0
    for t in range(threads):
        a, b = t * per, min(n_reads, (t + 1) * per)
        if a >= b:
            break
        nb = fb.fastq_nbytes(b - a, READ_LEN, first_id + a)
        jobs.append((a, b - a, off, nb))
        off += nb
    assert off == need

    def work(a, n, o, nb):
        got = fb.lib().fb2_synth_fastq(base + o, nb, genome.ctypes.data, genome.size, n, READ_LEN, ERR_RATE, seed,
                                       first_id + a, None)
        assert got == nb

    ths = [threading.Thread(target=work, args=j) for j in jobs]
    [t.start() for t in ths]
    [t.join() for t in ths]
    return buf, need, n_reads * READ_LEN


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms.  The sampler is started BEFORE the warm-up
    (nvidia-smi needs a few hundred ms to deliver its first sample, the timed region is ~150 ms) and the
    samples are cut to the timed region by their timestamps; if none falls inside, the samples of the
    warm-up (same workload, same load) are used and `window` says so."""
    QUERY = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.path = gpu_index, None, None
        self.t0 = self.t1 = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def region_start(self):
        self.t0 = time.time()

    def region_end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.05)   # let the last samples of the region reach the file
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[2]), float(f[3]), f[6:10]))
            except ValueError:
                continue
        os.unlink(self.path)
        inside = [r for r in rows if self.t0 is not None and self.t0 - 0.01 <= r[0] <= (self.t1 or 1e30) + 0.01]
        window = "timed region"
        if not inside:
            inside, window = rows, "warm-up + timed region (no sample fell inside the timed region)"
        reasons = set()
        for _, _, _, flags in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), flags):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if inside:
            out.update(sm_mhz=float(np.median([r[1] for r in inside])), sm_max_mhz=float(max(r[2] for r in inside)),
                       reasons=sorted(reasons), samples=len(inside), window=window)
        return out


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_baseline(fb, oracle, genome, seed, n_reads, threads):
    """Oracle port timed on the host cores on a bounded sample of the same workload.
    threads > 1 mirrors rayon's only parallelism: one sketcher per file (lib.rs:34-36)."""
    per = n_reads // threads
    bufs = [gen_fastq(fb, genome, per, seed, first_id=i * per, threads=1)[0].tobytes() for i in range(threads)]
    sp = oracle.mash_params(N_HASHES * OVERSKETCH, N_HASHES, True, K, 0)
    res = [None] * threads

    def work(i):
        res[i] = oracle.sketch_stream(bufs[i], sp, oracle.make_filter(True, (None, None), ERR_FILTER, STRAND_FILTER))

    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    dt = time.perf_counter() - t0
    assert all(r[0] == oracle.OK for r in res)
    bases = per * threads * READ_LEN
    return bases / dt / 1e9, dt, bufs[0], res[0][1]


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port) on all host
    cores; each step sketches a bounded sample (cores x sample_reads reads) of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import finch_rs_b200 as fb   # only for the synthetic generator (host code in the product lib)
    oracle = load_oracle()
    cores = os.cpu_count() or 1
    genome = fb.synth_genome(GENOME_LEN, 2)
    per_core = max(2000, args.ref_reads // cores)
    vals = []
    for it in range(args.warmup + args.steps):
        v, dt, _, _ = cpu_baseline(fb, oracle, genome, 3, per_core * cores, cores)
        if it >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals])) * 1e3
    sample = f"{per_core * cores} reads x {READ_LEN} bp per step ({per_core} per core, one file per core as rayon would)"
    line = {
        "impl": "reference", "metric": "Gbases/s sketched (k=21, n=1000, 150bp FASTQ)", "value": value,
        "unit": "Gbases/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(args), "note": "CPU reference arm: bounded sample of the workload"},
        "cpu_baseline": {"value": value, "unit": "Gbases/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=JSON_OUT, flush=True)


def workload_name(args):
    return (f"configs[1]: single synthetic FASTQ {args.reads} x {READ_LEN} bp reads (5 Mbp genome, 0.5% errors), "
            f"k={K}, n={N_HASHES}, --filter on (heap {N_HASHES * OVERSKETCH}), per GPU")


def bind_to_gpu_numa_node(torch, local):
    """Multi-rank runs: keep this rank's threads (and so its pinned host buffers, first-touch) on the NUMA
    node its GPU hangs off, like `numactl --cpunodebind --membind`.  Best effort: None when unknown."""
    try:
        bdf = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), "pci_bus_id") else None
        if bdf is None:
            import ctypes as C
            buf = C.create_string_buffer(32)
            if C.CDLL("libcudart.so.12").cudaDeviceGetPCIBusId(buf, 32, local) != 0:
                return None
            bdf = buf.value.decode()
        node = int(open(f"/sys/bus/pci/devices/{bdf.lower()}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, z = part.partition("-")
            cpus.update(range(int(a), int(z or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--cpu-reads", type=int, default=150_000, help="cpu_baseline sample (reads, 1 core)")
    ap.add_argument("--ref-reads", type=int, default=6_000_000, help="--impl reference sample per step (all cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    # The contract is ONE JSON line on stdout.  Libraries write banners to fd 1 (NCCL prints its version
    # there): keep the real stdout aside for the JSON line and send everything else to stderr.
    global JSON_OUT
    sys.stdout.flush()
    JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import finch_rs_b200 as fb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or fb.lib().fb2_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: finch_rs_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(torch, local) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    # ---- synthetic input: pinned host copy + HBM-resident copy -------------------------------
    genome = fb.synth_genome(GENOME_LEN, 2)
    need = fb.fastq_nbytes(args.reads, READ_LEN, 0)
    host = torch.empty(need, dtype=torch.uint8, pin_memory=True)
    _, nbytes, nbases = gen_fastq(fb, genome, args.reads, seed=3 + 1000 * rank, out_ptr=host.data_ptr())
    devbuf = host.to(dev, non_blocking=False)
    torch.cuda.synchronize()

    sp = fb.SketchParams.from_cli("mash", n_hashes=N_HASHES, kmer_length=K, seed=0, oversketch=OVERSKETCH,
                                  filters_enabled=True, device=local)
    fp = fb.FilterParams(True, (None, None), ERR_FILTER, STRAND_FILTER)
    stream = torch.cuda.Stream(dev)   # a real (non-legacy) stream: kernels and the timing events share it
    torch.cuda.set_stream(stream)
    sk = sp.create_sketcher(stream=stream.cuda_stream)
    L = fb.lib()

    def finish(h, c, x):
        """host filter + truncate (filter_counts + process_post_filter), as sketch_stream does"""
        hh, cc, xx, _ = fb.filter_counts(fp, h, c, x, fb.FORMAT_FASTQ)
        assert len(hh) >= N_HASHES, "strict: too few kmers"
        return hh[:N_HASHES], cc[:N_HASHES], xx[:N_HASHES]

    gathered = []

    def step(resident):
        sk.reset()
        if resident:
            sk.feed_device(devbuf.data_ptr(), nbytes, final=True)
        else:
            sk.feed_fastx_ptr(host.data_ptr(), nbytes, final=True)
        res = sk.sketch("bench.fq", fp)   # to_vec + filter_counts + process_post_filter (lib.rs:78-82)
        hh, cc, xx, seq_len, n_kmers = res.hashes_u64, res.counts, res.extra_counts, res.seq_length, res.num_valid_kmers
        if dist is not None:  # one NCCL gather of the finished sketch (hash, count, extra) to rank 0
            t = torch.from_numpy(np.stack([hh.view(np.int64), cc.astype(np.int64), xx.astype(np.int64)])).to(dev)
            outl = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
            dist.gather(t, outl, dst=0)
            if rank == 0:
                gathered[:] = outl
        return hh, cc, xx, seq_len, n_kmers

    def timed(resident, steps, warmup, sample_clocks=False):
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()
            time.sleep(0.5)      # nvidia-smi start-up; the warm-up below keeps the GPU under the same load
        for _ in range(warmup):
            step(resident)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.region_start()
        st0 = sk.stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        last = None
        for _ in range(steps):
            last = step(resident)
        e1.record(stream)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.region_end()
        clocks = sampler.stop() if sampler else None
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        st1 = sk.stats()
        return ms, last, {k: st1[k] - st0[k] for k in st1}, clocks

    ms_res, last_res, stats_res, clocks = timed(True, args.steps, args.warmup, sample_clocks=True)
    ms_e2e, last_e2e, stats_e2e, _ = timed(False, max(3, args.steps // 2), 1)
    e2e_steps = max(3, args.steps // 2)

    # ---- roofline of the dominant kernel (k-mer hash kernel), CUDA events on its own stream ----
    sk.enable_timing(True)
    s0 = sk.stats()
    for _ in range(2):
        step(True)
    s1 = sk.stats()
    sk.enable_timing(False)
    hash_ms = s1["hash_kernel_ms"] - s0["hash_kernel_ms"]
    hash_syms = s1["hash_symbols"] - s0["hash_symbols"]
    hash_launches = s1["hash_launches"] - s0["hash_launches"]
    parse_ms = s1["parse_kernel_ms"] - s0["parse_kernel_ms"]
    peak, peak_src = measured_peak()
    achieved = hash_syms * 1.0 / (hash_ms * 1e-3) / 1e9 if hash_ms > 0 else 0.0   # 1 B per base

    # ---- correctness at full size: size-independent properties ----------------------------------
    hh, cc, xx, seq_len, n_kmers = last_res
    assert seq_len == nbases, (seq_len, nbases)
    assert n_kmers == args.reads * (READ_LEN - K + 1), n_kmers           # no N in the synthetic reads
    assert np.all(hh[1:] > hh[:-1]) and len(hh) == N_HASHES               # strictly ascending, final size
    assert np.all(xx <= cc) and np.all(cc >= 1)
    assert all(np.array_equal(a, b) for a, b in zip(last_res[:3], last_e2e[:3])), "resident and e2e paths disagree"

    total_bases = nbases * world
    value = total_bases * args.steps / (ms_res * 1e-3) / 1e9
    e2e_value = total_bases * e2e_steps / (ms_e2e * 1e-3) / 1e9

    line = {
        "metric": "Gbases/s sketched (k=21, n=1000, 150bp FASTQ)", "value": value, "unit": "Gbases/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_res / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(args), "fastq_bytes_per_gpu": nbytes, "bases_per_gpu": nbases,
                   "l2": "inputs (3.1 GB per step) are far larger than the 126 MB L2; no explicit flush",
                   "parallelism": f"files x {world} (one file per GPU, NCCL gather of finished sketches)" if world > 1 else "1 GPU",
                   "numa_node_rank0": numa,
                   "chunk_mb": int(os.environ.get("FB2_CHUNK_MB", "128"))},
        "e2e": {"value": e2e_value, "unit": "Gbases/s", "h2d_bytes_per_step": int(stats_e2e["h2d_bytes"] // e2e_steps),
                "d2h_bytes_per_step": int(stats_e2e["d2h_bytes"] // e2e_steps), "ms_per_step": ms_e2e / e2e_steps,
                "steps": e2e_steps},
        "gpu_launches": int(stats_res["kernel_launches"]),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if peak else None,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one steady-state launch (61.0 M symbols of a
                     # 128 MiB chunk), ncu --set full: profiles/r01_ncu_full_v10_hash_fullsize.txt  => 1.08 B/symbol
                     "traffic": 66.04e6, "traffic_unit": "bytes per launch (61.0e6 algorithmic bytes)",
                     "kernel": "fb2::hash_kernel<21>", "peak_source": peak_src,
                     "algorithmic_bytes_per_unit": "1 B per base (symbol) walked by the hash kernel",
                     "avg_launch_ms": hash_ms / max(1, hash_launches), "launches_per_step": hash_launches / 2,
                     "hash_kernel_share_of_step": (hash_ms / 2) / (ms_res / args.steps),
                     "parse_kernels_ms_per_step": parse_ms / 2,
                     "note": "integer-issue-bound, not HBM-bound: ~100 SASS instr per k-mer split over the two half-rate integer pipes (ALU ~60%, IMAD ~60% busy, issue ~70% in ncu); see DESIGN.md"},
        "bit_exact": None,
        "aux": {"prunes_per_step": stats_res["prunes"] / args.steps, "chunks_per_step": stats_res["chunks"] / args.steps,
                "hash_launches_per_step": stats_res["hash_launches"] / args.steps,
                "kernel_launches_per_step": stats_res["kernel_launches"] / args.steps},
    }

    # ---- CPU baseline (oracle port, 1 core, bounded sample) + bit-exactness on that sample ------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        oracle = load_oracle()
        v, dt, sample_bytes, osk = cpu_baseline(fb, oracle, genome, 3, args.cpu_reads, 1)
        line["cpu_baseline"] = {"value": v, "unit": "Gbases/s", "cores": 1, "kind": "port",
                                "sample": f"first {args.cpu_reads} reads x {READ_LEN} bp of the workload, {dt:.1f} s"}
        gsk = fb.sketch_stream(sample_bytes, "sample.fq", fb.SketchParams.mash(N_HASHES * OVERSKETCH, N_HASHES, True, K, 0, local), fp)
        line["bit_exact"] = bool(np.array_equal(gsk.hashes_u64, osk["hashes"]) and np.array_equal(gsk.counts, osk["counts"])
                                 and np.array_equal(gsk.extra_counts, osk["extras"])
                                 and (gsk.seq_length, gsk.num_valid_kmers) == (osk["seq_length"], osk["num_valid_kmers"]))
    if rank == 0:
        if world > 1:
            assert len(gathered) == world and np.array_equal(gathered[0][0].cpu().numpy().view(np.uint64), hh)
        print(json.dumps(line), file=JSON_OUT, flush=True)
    sk.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

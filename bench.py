#!/usr/bin/env python
"""bench.py -- Gbases/s sketched (k=21, n=1000, 150 bp FASTQ, --filter on) on N B200s.

One "step" = one pass of the sketching hot path over one synthetic FASTQ of `--reads` x 150 bp
reads (BASELINE.json configs[1]: 10 M reads, 1.5 Gbases, ~3.1 GB of FASTQ text; 300x coverage of a
5 Mbp genome, 0.5 % substitution errors; tools/workloads.py), with the CLI's resolved parameters for
`finch sketch -k 21 -n 1000 -f`:  MashSketcher heap 200 000, strand filter 0.1, err filter 0.21,
final size 1000, strict.

  value  : whole-job Gbases/s, FASTQ bytes already resident in HBM when the timed region starts
           (fb2_sketcher_feed_device), result read back to the host and filtered every step.
  e2e    : the same through the host-facing C-ABI call (fb2_sketcher_feed_fastx) from a PINNED
           host buffer, host->device copies inside the timed region; `h2d_ceiling_gbs` is the
           platform's measured H2D rate for the same buffers at the same N with no kernels running.
  N > 1  : one rank per GPU (torchrun), every rank sketches its own file (independent files shard
           with no data-path collective) and one NCCL gather per step brings the finished
           1000-entry sketches to rank 0; weak scaling.
  bit_exact : EVERY rank's full-size result (hashes, counts, extra counts, k-mer bytes, totals) is compared
           with the sha256 digest of the CPU oracle's result on the same bytes, committed under
           tests/golden/full_digests.json (made by tests/golden/make_full_digests.py).
  --impl reference : the CPU path of the reference (oracle port; the Rust reference cannot be built
           here) on all host cores, on the same workload (one file slice per core, as rayon would).
"""
import argparse
import hashlib
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
sys.path.insert(0, os.path.join(ROOT, "tools"))

import workloads as W   # noqa: E402  (tools/workloads.py: the configs as seeded synthetic inputs)

K, N_HASHES, OVERSKETCH, READ_LEN = W.K_C2, W.N_HASHES, W.OVERSKETCH, W.READ_LEN
JSON_OUT = sys.stdout   # replaced in main(): the real stdout, kept apart from library chatter on fd 1
ERR_FILTER, STRAND_FILTER = 1.0 * K / 100.0, 0.1   # cli.rs:264-265, :142
METRIC = "Gbases/s sketched (k=21, n=1000, 150bp FASTQ)"


def load_oracle():
    """CPU oracle = test/baseline infrastructure (only the cpu_baseline / reference legs use it)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    oracle.build()
    return oracle


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def config_of(args, world):
    """The `config` object: identical in the b200 arm and the reference arm for the same command line."""
    nbytes = W.synth.fastq_nbytes(args.reads, READ_LEN, 0)
    return {"workload": (f"configs[1]: single synthetic FASTQ {args.reads} x {READ_LEN} bp reads (5 Mbp genome, 0.5% errors), "
                         f"k={K}, n={N_HASHES}, --filter on (heap {N_HASHES * OVERSKETCH}), per GPU"),
            "reads_per_gpu": args.reads, "read_len": READ_LEN, "fastq_bytes_per_gpu": nbytes,
            "bases_per_gpu": args.reads * READ_LEN,
            "l2": "inputs (3.1 GB per step) are far larger than the 126 MB L2; no explicit flush",
            "parallelism": (f"files x {world} (one file per GPU, NCCL gather of finished sketches)" if world > 1 else "1 GPU")}


def load_digests():
    try:
        return json.load(open(os.path.join(ROOT, "tests", "golden", "full_digests.json")))
    except Exception:
        return {}


def kernel_src_sha16():
    """Identity of the hash kernel's source: profiles/*_ncu_hash.json entries are only valid for the same source."""
    h = hashlib.sha256()
    for f in ("hash.cu", "common.cuh", "device_types.cuh"):
        h.update(open(os.path.join(ROOT, "finch_rs_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def profiled_traffic():
    """dram bytes per launch of the hash kernel from the newest committed ncu capture of THIS kernel source."""
    pdir = os.path.join(ROOT, "profiles")
    sha = kernel_src_sha16()
    best = None
    for name in sorted(os.listdir(pdir)) if os.path.isdir(pdir) else []:
        if not (name.endswith(".json") and "ncu_hash" in name):
            continue
        try:
            d = json.load(open(os.path.join(pdir, name)))
        except Exception:
            continue
        if d.get("kernel_src_sha16") == sha:
            best = (d, name)
    if not best:
        return None, f"no profiles/*ncu_hash*.json for kernel source {sha}"
    d, name = best
    return d, name


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


def cpu_sketch_slices(oracle, slices):
    """Oracle port over `slices` (one FASTQ byte buffer per thread) concurrently -- rayon's only parallelism is
    one sketcher per file (lib.rs:34-36).  -> (seconds, results)"""
    sp = oracle.mash_params(N_HASHES * OVERSKETCH, N_HASHES, True, K, 0)
    res = [None] * len(slices)

    def work(i):
        res[i] = oracle.sketch_stream(slices[i], sp, oracle.make_filter(True, (None, None), ERR_FILTER, STRAND_FILTER))

    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(i,)) for i in range(len(slices))]
    [t.start() for t in ths]
    [t.join() for t in ths]
    dt = time.perf_counter() - t0
    assert all(r[0] == oracle.OK for r in res)
    return dt, res


def fastq_slices(genome, seed, n_reads, parts):
    """The workload's FASTQ cut into `parts` files of whole reads (same reads, same bytes per read)."""
    per = (n_reads + parts - 1) // parts
    out, first = [], 0
    while first < n_reads:
        n = min(per, n_reads - first)
        out.append(W.synth.synth_fastq(genome, n, READ_LEN, W.ERR_RATE, seed, first_read_id=first)[0])
        first += n
    return out


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port) on all host cores over
    the SAME workload as the b200 arm (args.reads reads per step), cut into one file per core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    oracle = load_oracle()
    cores = host_cores()
    genome = W.c2_genome()
    n_reads = args.ref_reads or args.reads
    slices = fastq_slices(genome, W.c2_seed(0), n_reads, cores)     # generated once, outside the timed steps
    vals = []
    for it in range(args.warmup + args.steps):
        dt, _ = cpu_sketch_slices(oracle, slices)
        if it >= args.warmup:
            vals.append(dt)
    bases = n_reads * READ_LEN
    ms = float(np.mean(vals)) * 1e3
    value = bases / (ms * 1e-3) / 1e9
    sample = (f"{n_reads} reads x {READ_LEN} bp per step = the whole workload of one GPU, as {len(slices)} files of "
              f"{(n_reads + cores - 1) // cores} reads, one per core (rayon's file-level parallelism; upstream runs ONE file on one core)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value,
        "unit": "Gbases/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": config_of(args, world),
        "cpu_baseline": {"value": value, "unit": "Gbases/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=JSON_OUT, flush=True)


def bind_to_gpu_numa_node(torch, local):
    """Multi-rank runs: keep this rank's threads (and so its pinned host buffers, first-touch) on the NUMA
    node its GPU hangs off, like `numactl --cpunodebind --membind`.  Best effort: None when unknown."""
    try:
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
    ap.add_argument("--reads", type=int, default=W.C2_READS)
    ap.add_argument("--cpu-reads", type=int, default=600_000, help="cpu_baseline sample (reads, 1 core)")
    ap.add_argument("--ref-reads", type=int, default=0, help="--impl reference reads per step (0 = --reads: same config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-host-strip", action="store_true", help="do not try the FB2_HOST_STRIP=1 mode of the e2e call")
    args = ap.parse_args()
    # The contract is ONE JSON line on stdout.  Libraries write banners to fd 1 (NCCL prints its version
    # there): keep the real stdout aside for the JSON line and send everything else to stderr.
    global JSON_OUT
    sys.stdout.flush()
    JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.warmup < 3:
        args.warmup = 3
    W.synth.build()
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
    genome = W.c2_genome()
    need = W.synth.fastq_nbytes(args.reads, READ_LEN, 0)
    host = torch.empty(need, dtype=torch.uint8, pin_memory=True)
    _, nbytes, nbases = W.c2_fastq(rank, args.reads, genome, out_ptr=host.data_ptr(),
                                   threads=max(1, min(32, host_cores() // max(1, world))))
    devbuf = host.to(dev, non_blocking=False)
    torch.cuda.synchronize()

    sp = fb.SketchParams.from_cli("mash", n_hashes=N_HASHES, kmer_length=K, seed=0, oversketch=OVERSKETCH,
                                  filters_enabled=True, device=local)
    fp = fb.FilterParams(True, (None, None), ERR_FILTER, STRAND_FILTER)
    stream = torch.cuda.Stream(dev)   # a real (non-legacy) stream: kernels and the timing events share it
    torch.cuda.set_stream(stream)
    sk = sp.create_sketcher(stream=stream.cuda_stream)

    gathered = []

    def step(resident):
        sk.reset()
        if resident:
            sk.feed_device(devbuf.data_ptr(), nbytes, final=True)
        else:
            sk.feed_fastx_ptr(host.data_ptr(), nbytes, final=True)
        res = sk.sketch("bench.fq", fp)   # to_vec + filter_counts + process_post_filter (lib.rs:78-82)
        if dist is not None:  # one NCCL gather of the finished sketch (hash, count, extra) to rank 0
            t = torch.from_numpy(np.stack([res.hashes_u64.view(np.int64), res.counts.astype(np.int64),
                                           res.extra_counts.astype(np.int64)])).to(dev)
            outl = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
            dist.gather(t, outl, dst=0)
            if rank == 0:
                gathered[:] = outl
        return res

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
    e2e_steps = max(3, args.steps // 2)
    os.environ["FB2_HOST_STRIP"] = "0"
    ms_e2e, last_e2e, stats_e2e, _ = timed(False, e2e_steps, 1)
    e2e_plain = {"value": nbases * world * e2e_steps / (ms_e2e * 1e-3) / 1e9, "ms_per_step": ms_e2e / e2e_steps,
                 "h2d_bytes_per_step": int(stats_e2e["h2d_bytes"] // e2e_steps)}
    # The same call with the host pre-strip (FB2_HOST_STRIP=1, strip.cpp): worker threads frame the FASTQ records and
    # only the sequence lines cross PCIe.  Needs host cores: used when this rank has >= 8 of them to itself.
    strip_threads = min(16, host_cores() // max(1, world))
    e2e_strip, used_strip = None, False
    if strip_threads >= 8 and not args.no_host_strip:
        os.environ["FB2_HOST_STRIP"] = "1"
        os.environ["FB2_STRIP_THREADS"] = str(strip_threads)
        ms_s, last_s, stats_s, _ = timed(False, e2e_steps, 2)
        os.environ["FB2_HOST_STRIP"] = "0"
        e2e_strip = {"value": nbases * world * e2e_steps / (ms_s * 1e-3) / 1e9, "ms_per_step": ms_s / e2e_steps,
                     "h2d_bytes_per_step": int(stats_s["h2d_bytes"] // e2e_steps), "threads": strip_threads}
        if ms_s < ms_e2e:      # the headline e2e is the best of the modes of the same public call
            ms_e2e, last_e2e, stats_e2e, used_strip = ms_s, last_s, stats_s, True
    # Both at once (FB2_HOST_STRIP=2, hostlogic.cpp sketch_stream_two_ended): host cores frame records from the front
    # of the stream while the link carries raw bytes from its back; fb2_sketch_stream on the same pinned buffer.
    e2e_two, used_two = None, False
    if not args.no_host_strip:
        os.environ["FB2_HOST_STRIP"] = "2"
        os.environ["FB2_STRIP_THREADS"] = str(max(1, strip_threads))

        def step_stream():
            res = fb.sketch_stream_ptr(host.data_ptr(), nbytes, "bench.fq", sp, fp)
            if dist is not None:
                t = torch.from_numpy(np.stack([res.hashes_u64.view(np.int64), res.counts.astype(np.int64),
                                               res.extra_counts.astype(np.int64)])).to(dev)
                outl = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
                dist.gather(t, outl, dst=0)
            return res
        for _ in range(2):
            step_stream()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        acc = {"h2d_bytes": 0, "d2h_bytes": 0, "kernel_launches": 0}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        last_t = None
        for _ in range(e2e_steps):
            last_t = step_stream()
            stt = fb.last_stream_stats()
            for kx in acc:
                acc[kx] += stt[kx]
        e1.record(stream)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ms_t = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms_t], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_t = float(t.item())
        os.environ["FB2_HOST_STRIP"] = "0"
        e2e_two = {"value": nbases * world * e2e_steps / (ms_t * 1e-3) / 1e9, "ms_per_step": ms_t / e2e_steps,
                   "h2d_bytes_per_step": int(acc["h2d_bytes"] // e2e_steps), "threads": max(1, strip_threads)}
        if ms_t < ms_e2e:
            ms_e2e, last_e2e, stats_e2e, used_strip, used_two = ms_t, last_t, acc, False, True

    # ---- platform H2D ceiling: the same pinned buffers, all N ranks at once, no kernels ---------------------
    def h2d_ceiling(reps=3):
        for _ in range(1):
            devbuf.copy_(host, non_blocking=True)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            devbuf.copy_(host, non_blocking=True)
        e1.record(stream)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return nbytes * reps * world / (ms * 1e-3) / 1e9     # aggregate GB/s over all ranks
    h2d_gbs = h2d_ceiling()

    # ---- roofline of the dominant kernel (k-mer hash kernel): CUDA events around every launch of the SAME
    # asynchronous path the timed region ran (events are read back when the chunk is settled) -------------------
    ROOF_STEPS = 3
    sk.enable_timing(True)
    step(True)
    torch.cuda.synchronize()
    s0 = sk.stats()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record(stream)
    for _ in range(ROOF_STEPS):
        step(True)
    r1.record(stream)
    torch.cuda.synchronize()
    s1 = sk.stats()
    sk.enable_timing(False)
    roof_ms_per_step = r0.elapsed_time(r1) / ROOF_STEPS
    hash_ms = s1["hash_kernel_ms"] - s0["hash_kernel_ms"]
    hash_syms = s1["hash_symbols"] - s0["hash_symbols"]
    hash_launches = s1["hash_launches"] - s0["hash_launches"]
    parse_ms = s1["parse_kernel_ms"] - s0["parse_kernel_ms"]
    peak, peak_src = measured_peak()
    achieved = hash_syms * 1.0 / (hash_ms * 1e-3) / 1e9 if hash_ms > 0 else 0.0   # 1 B per base
    prof, prof_name = profiled_traffic()

    # ---- correctness at FULL size: oracle digest of this rank's input + size-independent properties --------------
    res = last_res
    hh, cc, xx = res.hashes_u64, res.counts, res.extra_counts
    assert res.seq_length == nbases, (res.seq_length, nbases)
    assert res.num_valid_kmers == args.reads * (READ_LEN - K + 1), res.num_valid_kmers   # no N in the synthetic reads
    assert np.all(hh[1:] > hh[:-1]) and len(hh) == N_HASHES               # strictly ascending, final size
    assert np.all(xx <= cc) and np.all(cc >= 1)
    dig = lambda r: W.sketch_digest(r.hashes_u64, r.counts, r.extra_counts, r.kmers[:, :K], r.seq_length, r.num_valid_kmers)
    mine, mine_e2e = dig(last_res), dig(last_e2e)
    assert mine == mine_e2e, "resident and e2e paths disagree"
    want = load_digests().get(f"c2/reads={args.reads}/rank={rank}")
    exact = None
    if want is not None:
        exact = all(mine[k] == want[k] for k in ("n", "sha256", "seq_length", "num_valid_kmers"))
    if dist is not None:
        t = torch.tensor([-1 if exact is None else int(exact)], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        exact = None if int(t.item()) < 0 else bool(int(t.item()))

    total_bases = nbases * world
    value = total_bases * args.steps / (ms_res * 1e-3) / 1e9
    e2e_value = total_bases * e2e_steps / (ms_e2e * 1e-3) / 1e9
    e2e_gbs = int(stats_e2e["h2d_bytes"] // e2e_steps) * world * e2e_steps / (ms_e2e * 1e-3) / 1e9

    line = {
        "metric": METRIC, "value": value, "unit": "Gbases/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_res / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": config_of(args, world),
        "e2e": {"value": e2e_value, "unit": "Gbases/s", "h2d_bytes_per_step": int(stats_e2e["h2d_bytes"] // e2e_steps),
                "d2h_bytes_per_step": int(stats_e2e["d2h_bytes"] // e2e_steps), "ms_per_step": ms_e2e / e2e_steps,
                "steps": e2e_steps, "h2d_gbs": e2e_gbs, "h2d_ceiling_gbs": h2d_gbs,
                "frac_of_h2d_ceiling": e2e_gbs / h2d_gbs if h2d_gbs else None,
                "h2d_ceiling_how": f"{world} rank(s) copying the same pinned buffers concurrently, no kernels, max over ranks",
                "mode": ("two-ended (FB2_HOST_STRIP=2): host cores frame FASTQ records from the front of the stream while raw bytes "
                         "from its back cross PCIe; the two tables are united exactly on the device") if used_two else
                        ("host pre-strip (FB2_HOST_STRIP=1): FASTQ framing on the host cores, sequence lines only over PCIe"
                         if used_strip else "raw FASTQ bytes over PCIe, parsed on the GPU"),
                "raw_bytes": e2e_plain, "host_strip": e2e_strip, "two_ended": e2e_two},
        "gpu_launches": int(stats_res["kernel_launches"]),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if peak else None,
                     "traffic": (prof or {}).get("dram_bytes_per_launch"), "traffic_source": prof_name,
                     "traffic_unit": "dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full)",
                     "kernel": "fb2::hash_kernel<21>", "kernel_src_sha16": kernel_src_sha16(), "peak_source": peak_src,
                     "algorithmic_bytes_per_unit": "1 B per base (symbol) walked by the hash kernel",
                     "algorithmic_bytes_per_launch": hash_syms / max(1, hash_launches),
                     "avg_launch_ms": hash_ms / max(1, hash_launches), "launches_per_step": hash_launches / ROOF_STEPS,
                     "hash_kernel_share_of_step": (hash_ms / ROOF_STEPS) / roof_ms_per_step,
                     "parse_kernels_ms_per_step": parse_ms / ROOF_STEPS, "ms_per_step_with_events": roof_ms_per_step,
                     "how": "CUDA events around every hash / parse launch of the asynchronous (timed) path, read back at settle",
                     "note": "integer-issue-bound, not HBM-bound (~100 SASS instr per k-mer on two half-rate integer pipes); see DESIGN.md"},
        "bit_exact": exact,
        "bit_exact_how": "sha256(hashes|counts|extras|kmers) + totals of every rank's FULL-size result vs tests/golden/full_digests.json (CPU oracle on the same bytes)",
        "aux": {"prunes_per_step": stats_res["prunes"] / args.steps, "chunks_per_step": stats_res["chunks"] / args.steps,
                "hash_launches_per_step": stats_res["hash_launches"] / args.steps,
                "kernel_launches_per_step": stats_res["kernel_launches"] / args.steps,
                "numa_node_rank0": numa, "chunk_mb": int(os.environ.get("FB2_CHUNK_MB", "128")), "host_cores": host_cores()},
    }

    # ---- CPU baseline (oracle port, 1 core, bounded sample) + bit-exactness on that sample too ------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        oracle = load_oracle()
        n_s = min(args.cpu_reads, args.reads)
        sample = fastq_slices(genome, W.c2_seed(0), n_s, 1)[0]
        dt, ores = cpu_sketch_slices(oracle, [sample])
        osk = ores[0][1]
        line["cpu_baseline"] = {"value": n_s * READ_LEN / dt / 1e9, "unit": "Gbases/s", "cores": 1, "kind": "port",
                                "sample": f"first {n_s} reads x {READ_LEN} bp of the workload, {dt:.1f} s"}
        gsk = fb.sketch_stream(sample, "sample.fq", fb.SketchParams.mash(N_HASHES * OVERSKETCH, N_HASHES, True, K, 0, local), fp)
        line["bit_exact_sample"] = bool(np.array_equal(gsk.hashes_u64, osk["hashes"]) and np.array_equal(gsk.counts, osk["counts"])
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

"""Kernel timeline of one C2-shaped resident step (torch.profiler / CUPTI: every kernel of the process, with start times):
busy time per kernel, idle gaps of the device, and the first chunk's share.  Diagnostic only."""
import sys, json, collections
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import torch
from torch.profiler import profile, ProfilerActivity
import finch_rs_b200 as fb
import workloads as W
n_reads = int(float(sys.argv[1])) if len(sys.argv) > 1 else W.C2_READS
buf, need, nbases = W.c2_fastq(0, n_reads)
d = torch.from_numpy(buf).cuda()
sp = fb.SketchParams.from_cli("mash", n_hashes=1000, kmer_length=21, filters_enabled=True)
fp = fb.FilterParams(True, (None, None), 0.21, 0.1)
h = sp.create_sketcher()
def step():
    h.reset()
    h.feed_device(d.data_ptr(), d.numel(), final=True)
    return h.sketch("c2.fq", fp)
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
prof.export_chrome_trace("gpurun_out/c2_trace.json")
ev = [e for e in json.load(open("gpurun_out/c2_trace.json"))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]; t1 = max(e["ts"] + e["dur"] for e in ev)
print(f"span {(t1-t0)/1e3:.3f} ms, {len(ev)} device activities")
by = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    n = e["name"].split("(")[0].replace("void ", "").replace("fb2::", "")
    by[n][0] += 1; by[n][1] += e["dur"]
for n, (c, t) in sorted(by.items(), key=lambda x: -x[1][1])[:25]:
    print(f"  {n[:60]:60s} n={c:4d} total={t/1e3:8.3f} ms avg={t/c:8.1f} us")
# device idle time (no activity on any stream)
cur = t0; idle = 0.0; gaps = []
for e in ev:
    if e["ts"] > cur: idle += e["ts"] - cur; gaps.append((e["ts"] - cur, cur - t0, e["name"][:40]))
    cur = max(cur, e["ts"] + e["dur"])
print(f"idle {idle/1e3:.3f} ms in {len(gaps)} gaps; largest:")
for g in sorted(gaps, reverse=True)[:12]: print(f"   {g[0]:8.1f} us at +{g[1]/1e3:7.3f} ms before {g[2]}")
# timeline of the first 1.6 ms and the last 1 ms
print("first 60 activities:")
for e in ev[:60]: print(f"   +{(e['ts']-t0)/1e3:7.3f} ms {e['dur']:8.1f} us  {e['name'][:70]}")
print("last 40 activities:")
for e in ev[-40:]: print(f"   +{(e['ts']-t0)/1e3:7.3f} ms {e['dur']:8.1f} us  {e['name'][:70]}")

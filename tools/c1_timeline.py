"""Kernel timeline of one C1 call (5 Mbp FASTA, Mash n = 1000, fb2 handle API from pinned host memory): every device
activity with its start time, idle gaps.  Diagnostic only (torch.profiler / CUPTI)."""
import sys, json, time
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import torch
from torch.profiler import profile, ProfilerActivity
import finch_rs_b200 as fb
import synth
data = synth.synth_fasta(5_000_000, n_records=1, line_width=80, seed=1)
sp = fb.SketchParams.from_cli("mash", n_hashes=1000, kmer_length=21)
fp = fb.FilterParams(None, (None, None), 0.21, 0.1)
host = torch.from_numpy(data).pin_memory()
def step():
    return fb.sketch_stream_ptr(host.data_ptr(), host.numel(), "c1.fa", sp, fp)
for _ in range(5): step()
ts = []
for _ in range(20):
    t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
print(f"sketch_stream: median {1e3*sorted(ts)[len(ts)//2]:.3f} ms, best {1e3*min(ts):.3f} ms")
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
prof.export_chrome_trace("gpurun_out/c1_trace.json")
tr = json.load(open("gpurun_out/c1_trace.json"))["traceEvents"]
ev = sorted([e for e in tr if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")], key=lambda e: e["ts"])
t0 = ev[0]["ts"]; t1 = max(e["ts"] + e["dur"] for e in ev)
busy = sum(e["dur"] for e in ev)
print(f"device span {(t1-t0)/1e3:.3f} ms, busy {busy/1e3:.3f} ms, {len(ev)} activities")
prev = t0
for e in ev:
    print(f"  +{(e['ts']-t0):8.1f} us (gap {e['ts']-prev:6.1f}) {e['dur']:7.1f} us  {e['name'][:64]}")
    prev = e["ts"] + e["dur"]
rt = sorted([e for e in tr if e.get("cat") == "cuda_runtime"], key=lambda e: e["ts"])
import collections
c = collections.Counter(e["name"] for e in rt)
print("runtime calls:", dict(c))

#!/usr/bin/env python
"""Pretty-print the interesting parts of a bench.py JSON line (last line of the file / stdin)."""
import json
import sys

txt = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
d = json.loads([l for l in txt.strip().splitlines() if l.startswith("{")][-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "bit_exact", "gpu_launches")})
print("e2e", d.get("e2e"))
print("aux", d.get("aux"))
r = d.get("roofline") or {}
print("roofline", {k: r.get(k) for k in ("achieved", "frac", "avg_launch_ms", "launches_per_step",
                                         "hash_kernel_share_of_step", "parse_kernels_ms_per_step")})
print("cpu", d.get("cpu_baseline"), "clocks", d.get("clocks"))

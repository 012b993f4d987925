#!/usr/bin/env python
"""profiles/make_ncu_hash_json.py <hash.ncu-rep> <out.json> -- the hash kernel's DRAM traffic per launch from an
`ncu --set full` capture, tagged with the identity of the kernel source it profiled (bench.py prints
`roofline.traffic` only from a capture of the same source)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    v = float(r[idx[name]].replace(',', ''))
    u = units[idx[name]]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


launches = []
for r in rows[2:]:
    if 'hash_kernel' not in r[idx['Kernel Name']]:
        continue
    launches.append({"kernel": r[idx['Kernel Name']][:60], "dram_read": val(r, 'dram__bytes_read.sum'),
                     "dram_write": val(r, 'dram__bytes_write.sum'),
                     "duration_us_under_ncu": float(r[idx['gpu__time_duration.sum']].replace(',', '')) *
                     {"ns": 1e-3, "us": 1, "ms": 1e3}.get(units[idx['gpu__time_duration.sum']], 1)})
import bench
out = {"kernel_src_sha16": bench.kernel_src_sha16(), "launches": launches,
       "dram_bytes_per_launch": sum(l["dram_read"] + l["dram_write"] for l in launches) / max(1, len(launches)),
       "source": os.path.basename(sys.argv[1]), "how": "ncu --set full --clock-control none, steady-state launches of C2 (128 MiB chunks)"}
json.dump(out, open(sys.argv[2], 'w'), indent=1)
print(json.dumps(out)[:400])

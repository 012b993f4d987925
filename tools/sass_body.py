#!/usr/bin/env python
"""Opcode histogram of the per-position body of hash_kernel<K> (between consecutive VOTE.ANY P0 lines)."""
import collections, re, subprocess, sys
obj = sys.argv[1]
k = sys.argv[2] if len(sys.argv) > 2 else "21"
var = sys.argv[3] if len(sys.argv) > 3 and sys.argv[3].isdigit() else "0"
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
# split per function
funcs = re.split(r"\n\s*Function : ", out)
for f in funcs:
    name = f.split("\n", 1)[0]
    if f"hash_kernelILi{k}ELb1ELi{var}E" not in name:
        continue
    ins = []
    for line in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    votes = [i for i, (a, t) in enumerate(ins) if re.search(r"VOTE\.ANY P\d, P\d", t)]
    print(name[:60], "total instr", len(ins), "votes at", votes)
    if len(votes) >= 3:
        a, b = votes[1], votes[2]
        # body: from the fallthrough target after vote a .. vote b.  Use the BRA target after vote a
        tgt = None
        for j in range(a, a + 3):
            m = re.search(r"BRA (0x[0-9a-f]+)", ins[j][1])
            if m:
                tgt = int(m.group(1), 16); break
        start = next(i for i, (ad, t) in enumerate(ins) if ad == tgt)
        body = ins[start:b + 2]
        h = collections.Counter()
        for ad, t in body:
            op = t.split()[1] if t.startswith("@") else t.split()[0]
            h[op.split(".")[0]] += 1
        alu = sum(v for k_, v in h.items() if k_ in ("LOP3", "SHF", "ISETP", "IADD3", "LEA", "SEL", "VIADD", "PRMT", "PLOP3", "IABS", "FLO", "POPC", "MOV", "VOTE"))
        fma = sum(v for k_, v in h.items() if k_ in ("IMAD", "FFMA", "FMUL", "FADD"))
        print("body len", len(body), "ALU", alu, "FMA", fma, dict(h.most_common()))
        if len(sys.argv) > 4:
            for ad, t in body:
                print(f"{ad:05x} {t}")

"""Messy FASTA/FASTQ generators for parity tests (numpy, seeded)."""
import numpy as np

ACGT = np.frombuffer(b"ACGT", np.uint8)


def rand_seq(rng, n, messy=0.0):
    s = ACGT[rng.integers(0, 4, size=n)].copy()
    if messy > 0 and n:
        m = rng.random(n)
        other = np.frombuffer(b"acgtNnRYKMuU-.*xX", np.uint8)
        idx = np.nonzero(m < messy)[0]
        s[idx] = other[rng.integers(0, len(other), size=len(idx))]
    return s.tobytes()


def fasta(rng, n_records=3, min_len=0, max_len=500, width=60, messy=0.02, crlf=False, blank_lines=False,
          final_newline=True, ragged=False):
    out = []
    nl = b"\r\n" if crlf else b"\n"
    for r in range(n_records):
        out.append(b">rec%d some description > with gt" % r + nl)
        n = int(rng.integers(min_len, max_len + 1))
        s = rand_seq(rng, n, messy)
        i = 0
        while i < n:
            w = int(rng.integers(1, width + 1)) if ragged else width
            out.append(s[i:i + w] + nl)
            i += w
            if blank_lines and rng.random() < 0.1:
                out.append(nl)
    data = b"".join(out)
    if not final_newline and data.endswith(nl):
        data = data[:-len(nl)]
    return data


def fastq(rng, n_records=10, min_len=1, max_len=200, messy=0.02, crlf=False, final_newline=True,
          at_in_qual=True):
    out = []
    nl = b"\r\n" if crlf else b"\n"
    for r in range(n_records):
        n = int(rng.integers(min_len, max_len + 1))
        s = rand_seq(rng, n, messy)
        q = rng.integers(33, 74, size=n).astype(np.uint8)
        if at_in_qual and n:
            q[0] = ord("@")          # '@' is a legal quality character, even at line start
            if n > 1:
                q[-1] = ord("+")
        out.append(b"@read%d/1 x" % r + nl + s + nl + b"+" + (b"read%d" % r if r % 2 else b"") + nl + q.tobytes() + nl)
    data = b"".join(out)
    if not final_newline and data.endswith(nl):
        data = data[:-len(nl)]
    return data

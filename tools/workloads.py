"""The BASELINE.json configs as concrete, seeded synthetic inputs -- one definition shared by bench.py,
bench_configs.py, the parity tests and tests/golden/make_full_digests.py, so the oracle digests committed under
tests/golden/ are digests of EXACTLY the bytes the benchmarks sketch."""
import hashlib

import numpy as np

import synth

K_C2, N_HASHES, OVERSKETCH, READ_LEN = 21, 1000, 200, 150
GENOME_LEN, GENOME_SEED, ERR_RATE = 5_000_000, 2, 0.005
C2_READS = 10_000_000
C3_FILES = 1024
C4_BASES = 3_000_000_000
C5_SKETCHES, C5_HASHES, C5_CLUSTERS, C5_SEED = 100_000, 1000, 1000, 5


def c2_seed(rank):
    return 3 + 1000 * rank


def c2_genome():
    return synth.synth_genome(GENOME_LEN, GENOME_SEED)


def c2_fastq(rank, n_reads=C2_READS, genome=None, out_ptr=None, threads=None):
    """configs[1]: rank's FASTQ of n_reads x 150 bp reads.  -> (numpy buffer | None, nbytes, nbases)"""
    genome = c2_genome() if genome is None else genome
    return synth.synth_fastq_parallel(genome, n_reads, READ_LEN, ERR_RATE, c2_seed(rank), 0, out_ptr, threads)


def c1_fasta():
    """configs[0]: one 5 Mbp FASTA record, 80-column lines."""
    return synth.synth_fasta(5_000_000, n_records=1, line_width=80, seed=1)


def c3_nbases(i):
    return int(4.5e6 + (i * 7919 % 1000) * 1e3)


def c3_fasta(i):
    """configs[2]: file i of the 1024-file batch: 4.5-5.5 Mbp, 1-3 contigs, 80-column lines."""
    return synth.synth_fasta(c3_nbases(i), n_records=1 + i % 3, line_width=80, seed=1000 + i)


def c4_fasta(n_bases=C4_BASES):
    """configs[3]: 24 records, 60-column lines, 2 % lowercase runs, 1 % N runs."""
    return synth.synth_fasta(n_bases, n_records=24, line_width=60, lower_frac=0.02, n_frac=0.01, seed=4)


def c5_rows(first=0, count=None, n_sk=C5_SKETCHES):
    """configs[4]: rows of the n_sk x 1000 sorted-hash matrix (1000 clusters sharing 50-95 % of their hashes)."""
    return synth.synth_sketches(n_sk, C5_HASHES, min(C5_CLUSTERS, max(1, n_sk // 100)), C5_SEED, first, count)


def c5_sample_pairs(n_sk=C5_SKETCHES, n_pairs=100_000, seed=55):
    """The (q, r) pairs the parity sample of C5 checks: half inside clusters (dense intersections), half anywhere."""
    rng = np.random.default_rng(seed)
    ncl = min(C5_CLUSTERS, max(1, n_sk // 100))
    q = rng.integers(0, n_sk, size=n_pairs, dtype=np.int64)
    r = rng.integers(0, n_sk, size=n_pairs, dtype=np.int64)
    half = n_pairs // 2
    per = max(1, n_sk // ncl)
    r[:half] = (q[:half] % ncl) + ncl * rng.integers(0, per, size=half)       # same cluster as q
    r = np.minimum(r, n_sk - 1)
    return q.astype(np.uint32), r.astype(np.uint32)


def sketch_digest(hashes, counts, extras, kmers, seq_length, num_valid_kmers):
    """sha256 over hashes (u64 LE) || counts (u32) || extras (u32) || k-mer bytes (n x k), plus the totals."""
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(hashes, np.uint64).tobytes())
    h.update(np.ascontiguousarray(counts, np.uint32).tobytes())
    h.update(np.ascontiguousarray(extras, np.uint32).tobytes())
    h.update(np.ascontiguousarray(kmers, np.uint8).tobytes())
    return {"n": int(len(hashes)), "sha256": h.hexdigest(), "seq_length": int(seq_length),
            "num_valid_kmers": int(num_valid_kmers)}

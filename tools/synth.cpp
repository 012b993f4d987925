// synth.cpp -- deterministic synthetic inputs for the benchmarks and tests: genomes, FASTA
// files and FASTQ read sets.  Built into tools/libfb2_synth.so -- NOT part of the product library, so the
// reference arm of bench.py and the oracle digests can generate the exact bench inputs without loading
// libfinch_b200.so.  Return bytes written, or the required size when out == NULL.
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>

// ---- synthetic inputs ------------------------------------------------------------------
static inline uint64_t splitmix64(uint64_t &x) {
    uint64_t z = (x += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() { return splitmix64(s); }
    double unif() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

extern "C" size_t fb2_synth_genome(uint8_t *out, size_t n_bases, uint64_t seed) {
    if (!out) return n_bases;
    Rng r(seed);
    size_t i = 0;
    while (i < n_bases) {
        uint64_t w = r.next();
        for (int b = 0; b < 32 && i < n_bases; ++b, w >>= 2) out[i++] = (uint8_t)"ACGT"[w & 3];
    }
    return n_bases;
}

extern "C" size_t fb2_synth_fasta(uint8_t *out, size_t cap, size_t n_bases, uint32_t n_records,
                                  uint32_t line_width, double lower_frac, double n_frac, uint64_t seed) {
    if (n_records == 0) n_records = 1;
    if (line_width == 0) line_width = 80;
    Rng r(seed);
    size_t o = 0;
    auto put = [&](uint8_t c) { if (out && o < cap) out[o] = c; ++o; };
    const size_t per = n_bases / n_records;
    for (uint32_t rec = 0; rec < n_records; ++rec) {
        const size_t len = rec + 1 == n_records ? n_bases - per * (n_records - 1) : per;
        char hdr[64];
        const int hl = snprintf(hdr, sizeof hdr, ">contig%u len=%zu\n", rec + 1, len);
        for (int i = 0; i < hl; ++i) put((uint8_t)hdr[i]);
        size_t i = 0, col = 0;
        int run_kind = 0;        // 0 plain, 1 lowercase run, 2 N run
        size_t run_left = 0;
        uint64_t w = 0; int wb = 0;
        while (i < len) {
            if (run_left == 0) {
                const double u = r.unif();
                // runs of ~200 bases; fractions are of bases
                if (u < n_frac) run_kind = 2; else if (u < n_frac + lower_frac) run_kind = 1; else run_kind = 0;
                run_left = 100 + (size_t)(r.next() % 200);
            }
            if (wb == 0) { w = r.next(); wb = 32; }
            uint8_t c = (uint8_t)"ACGT"[w & 3]; w >>= 2; --wb;
            if (run_kind == 1) c = (uint8_t)(c | 0x20);
            else if (run_kind == 2) c = 'N';
            put(c); ++i; --run_left;
            if (++col == line_width) { put('\n'); col = 0; }
        }
        if (col) put('\n');
    }
    return o;
}

extern "C" size_t fb2_synth_fastq(uint8_t *out, size_t cap, const uint8_t *genome, size_t genome_len,
                                  uint64_t n_reads, uint32_t read_len, double err_rate, uint64_t seed,
                                  uint64_t first_read_id, uint64_t *n_bases_out) {
    if (!genome || genome_len < read_len) return 0;
    size_t o = 0;
    auto put = [&](uint8_t c) { if (out && o < cap) out[o] = c; ++o; };
    const uint32_t err_thr = (uint32_t)(err_rate * 4294967296.0);
    for (uint64_t i = 0; i < n_reads; ++i) {
        Rng r(seed ^ ((first_read_id + i) * 0xD1342543DE82EF95ULL + 0x2545F4914F6CDD1DULL));  // per-read stream
        const uint64_t id = first_read_id + i;
        char hdr[32];
        const int hl = snprintf(hdr, sizeof hdr, "@r%llu\n", (unsigned long long)id);
        for (int j = 0; j < hl; ++j) put((uint8_t)hdr[j]);
        const size_t start = (size_t)(r.next() % (genome_len - read_len + 1));
        const bool rev = r.next() & 1;
        for (uint32_t j = 0; j < read_len; ++j) {
            uint8_t c = rev ? genome[start + read_len - 1 - j] : genome[start + j];
            if (rev) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A';
            const uint64_t w = r.next();
            if ((uint32_t)w < err_thr) {  // substitution by one of the three other bases
                const uint8_t alt = (uint8_t)"ACGT"[(w >> 32) & 3];
                c = alt != c ? alt : (uint8_t)"ACGT"[((w >> 32) + 1) & 3];
            }
            put(c);
        }
        put('\n'); put('+'); put('\n');
        for (uint32_t j = 0; j < read_len; ++j) put('I');
        put('\n');
    }
    if (n_bases_out) *n_bases_out = n_reads * read_len;
    return o;
}

// ---- C5: clustered sketches -------------------------------------------------------------------------
// Sketch i (global index) = sorted union of `share` hashes drawn from its cluster's base set
// (cluster = i % n_clusters, base set = n_hashes hashes from the cluster's own RNG) and n_hashes - share of its
// own, share uniform in [50 %, 95 %] of n_hashes.  Values are 63-bit (like the upper part of a murmur range).
// Duplicates inside a sketch (probability ~1e-13 per pair) are bumped to keep the rows strictly ascending.
#include <algorithm>
#include <vector>
extern "C" void fb2_synth_sketches(uint64_t *out, uint64_t count, uint32_t n_hashes, uint32_t n_clusters,
                                   uint64_t seed, uint64_t first, uint64_t /*n_sk*/) {
    if (n_clusters == 0) n_clusters = 1;
    std::vector<uint64_t> base(n_hashes);
    std::vector<uint32_t> perm(n_hashes);
    for (uint64_t r = 0; r < count; ++r) {
        const uint64_t i = first + r;
        const uint64_t c = i % n_clusters;
        Rng rb(seed * 0x9E3779B97F4A7C15ULL + 0xC1u + c * 0xD1342543DE82EF95ULL);
        for (uint32_t j = 0; j < n_hashes; ++j) base[j] = rb.next() >> 1;
        Rng ro(seed * 0x9E3779B97F4A7C15ULL + 0x5EEDu + (i + 1) * 0x2545F4914F6CDD1DULL);
        const uint32_t lo = n_hashes / 2, hi = n_hashes - n_hashes / 20;
        const uint32_t share = lo + (uint32_t)(ro.next() % (uint64_t)(hi - lo + 1));
        for (uint32_t j = 0; j < n_hashes; ++j) perm[j] = j;
        for (uint32_t j = 0; j < share; ++j) {   // partial Fisher-Yates: `share` distinct base indices
            const uint32_t t = j + (uint32_t)(ro.next() % (uint64_t)(n_hashes - j));
            std::swap(perm[j], perm[t]);
        }
        uint64_t *row = out + r * (uint64_t)n_hashes;
        for (uint32_t j = 0; j < share; ++j) row[j] = base[perm[j]];
        for (uint32_t j = share; j < n_hashes; ++j) row[j] = ro.next() >> 1;
        std::sort(row, row + n_hashes);
        for (uint32_t j = 1; j < n_hashes; ++j) if (row[j] <= row[j - 1]) row[j] = row[j - 1] + 1;
    }
}

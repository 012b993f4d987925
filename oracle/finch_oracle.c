/*
 * finch_oracle.c -- CPU ORACLE (test infrastructure only; see finch_oracle.h).
 *
 * Literal, single-threaded restatement of the finch-rs CPU hot path.  Each function cites
 * the reference file:line it follows (paths relative to the reference root).
 * Shape is deliberately the reference's: per-record normalize + reverse-complement buffers,
 * byte-compare canonicalisation, murmur3 per k-mer, binary max-heap + hash map `push`.
 */
#include "finch_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ====================================================================================== */
/* murmurhash3 0.0.5 :: murmurhash3_x64_128  (call site lib/src/sketch_schemes/hashing.rs:11) */
/* ====================================================================================== */
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t fmix64(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}
static inline uint64_t load_le64(const uint8_t *p) {
    uint64_t v = 0;
    for (int i = 7; i >= 0; --i) v = (v << 8) | p[i];
    return v;
}

void fo_murmur3_x64_128(const uint8_t *data, size_t len, uint64_t seed, uint64_t out[2]) {
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = seed, h2 = seed;
    size_t nblocks = len / 16;
    for (size_t i = 0; i < nblocks; ++i) {
        uint64_t k1 = load_le64(data + 16 * i), k2 = load_le64(data + 16 * i + 8);
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const uint8_t *tail = data + 16 * nblocks;
    size_t t = len & 15;
    uint64_t k1 = 0, k2 = 0;
    if (t > 8) {
        for (size_t i = t; i > 8; --i) k2 = (k2 << 8) | tail[i - 1];
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
    }
    if (t > 0) {
        size_t m = t > 8 ? 8 : t;
        for (size_t i = m; i > 0; --i) k1 = (k1 << 8) | tail[i - 1];
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
    }
    h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2; h2 += h1;
    out[0] = h1; out[1] = h2;
}

/* hashing.rs:9-12 */
uint64_t fo_hash_f(const uint8_t *item, size_t len, uint64_t seed) {
    uint64_t o[2];
    fo_murmur3_x64_128(item, len, seed, o);
    return o[0];
}

/* ====================================================================================== */
/* needletail 0.5.0 :: Sequence::normalize(false) / reverse_complement / canonical_kmers    */
/* (call sites mash.rs:73-76, scaled.rs:71-74).  SURVEY 8a rows S5-S7.                      */
/* ====================================================================================== */
size_t fo_normalize(const uint8_t *in, size_t n, uint8_t *out) {
    size_t m = 0;
    for (size_t i = 0; i < n; ++i) {
        uint8_t c = in[i], o;
        switch (c) {
        case 'A': case 'C': case 'G': case 'T': o = c; break;
        case 'a': o = 'A'; break;
        case 'c': o = 'C'; break;
        case 'g': o = 'G'; break;
        case 't': case 'u': case 'U': o = 'T'; break;       /* uridine -> thymine */
        case '-': case '.': case '~': o = '-'; break;       /* gaps */
        case ' ': case '\t': case '\r': case '\n': continue; /* whitespace removed */
        default: o = 'N'; break;                             /* everything else (incl. IUPAC) */
        }
        out[m++] = o;
    }
    return m;
}

static inline uint8_t complement(uint8_t c) {
    switch (c) {
    case 'A': return 'T'; case 'T': return 'A';
    case 'C': return 'G'; case 'G': return 'C';
    case 'a': return 't'; case 't': return 'a';
    case 'c': return 'g'; case 'g': return 'c';
    default: return c;  /* others pass through */
    }
}
void fo_reverse_complement(const uint8_t *in, size_t n, uint8_t *out) {
    for (size_t i = 0; i < n; ++i) out[i] = complement(in[n - 1 - i]);
}
static inline int is_good_base(uint8_t c) {
    return c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'a' || c == 'c' || c == 'g' || c == 't';
}

/* Generic driver shared by fo_kmer_stream and fo_process: calls f(ctx, kmer, is_rc) for each
 * canonical k-mer.  fwd < rc (byte-lexicographic) ? (fwd,false) : (rc,true)  -- a palindrome
 * therefore reports the rc slice with is_rc = true. */
typedef void (*kmer_fn)(void *ctx, const uint8_t *kmer, int is_rc);
static void for_canonical_kmers(const uint8_t *raw, size_t len, uint8_t k, kmer_fn f, void *ctx) {
    if (k == 0) return;
    uint8_t *norm = (uint8_t *)malloc(len ? len : 1);
    size_t L = fo_normalize(raw, len, norm);
    uint8_t *rc = (uint8_t *)malloc(L ? L : 1);
    fo_reverse_complement(norm, L, rc);
    size_t good = 0; /* length of the run of good bases ending at position i */
    for (size_t i = 0; i < L; ++i) {
        good = is_good_base(norm[i]) ? good + 1 : 0;
        if (good >= k) {
            size_t p = i + 1 - k;
            const uint8_t *fw = norm + p, *rv = rc + (L - p - k);
            if (memcmp(fw, rv, k) < 0) f(ctx, fw, 0); else f(ctx, rv, 1);
        }
    }
    free(norm); free(rc);
}

struct stream_ctx { uint64_t *h; uint8_t *rc; uint8_t *kmers; size_t cap, n; uint8_t k; uint64_t seed; };
static void stream_cb(void *vctx, const uint8_t *kmer, int is_rc) {
    struct stream_ctx *c = (struct stream_ctx *)vctx;
    if (c->n < c->cap) {
        if (c->h) c->h[c->n] = fo_hash_f(kmer, c->k, c->seed);
        if (c->rc) c->rc[c->n] = (uint8_t)is_rc;
        if (c->kmers) memcpy(c->kmers + c->n * c->k, kmer, c->k);
    }
    c->n++;
}
size_t fo_kmer_stream(const uint8_t *seq, size_t len, uint8_t k, uint64_t seed, uint64_t *hashes,
                      uint8_t *is_rc, uint8_t *kmers, size_t cap) {
    struct stream_ctx c = {hashes, is_rc, kmers, cap, 0, k, seed};
    for_canonical_kmers(seq, len, k, stream_cb, &c);
    return c.n;
}

/* ====================================================================================== */
/* Sketchers: BinaryHeap<HashedItem<Vec<u8>>> + HashMap<ItemHash,(u32,u32)>                 */
/* (mash.rs:10-18, scaled.rs:10-19)                                                         */
/* ====================================================================================== */
typedef struct { uint64_t hash; uint8_t *kmer; } heap_item;
typedef struct { uint64_t key; uint32_t count, extra; uint8_t used; } map_slot;

struct fo_sketcher {
    int kind; /* 0 mash, 1 scaled */
    heap_item *heap; size_t heap_len, heap_cap;        /* max-heap ordered by hash only */
    map_slot *map; size_t map_cap, map_len;            /* open addressing, linear probing */
    uint8_t k; uint64_t total_kmers, total_bases; size_t size; uint64_t max_hash, seed;
};

static inline size_t map_home(uint64_t key, size_t cap) {
    return (size_t)((key * 0x9E3779B97F4A7C15ULL) >> 20) & (cap - 1);
}
static void map_grow(fo_sketcher *s);
static map_slot *map_find(fo_sketcher *s, uint64_t key) {
    size_t i = map_home(key, s->map_cap);
    while (s->map[i].used) {
        if (s->map[i].key == key) return &s->map[i];
        i = (i + 1) & (s->map_cap - 1);
    }
    return NULL;
}
static void map_insert(fo_sketcher *s, uint64_t key, uint32_t count, uint32_t extra) {
    if ((s->map_len + 1) * 2 > s->map_cap) map_grow(s);
    size_t i = map_home(key, s->map_cap);
    while (s->map[i].used) i = (i + 1) & (s->map_cap - 1);
    s->map[i].key = key; s->map[i].count = count; s->map[i].extra = extra; s->map[i].used = 1;
    s->map_len++;
}
static void map_grow(fo_sketcher *s) {
    map_slot *old = s->map; size_t oc = s->map_cap;
    s->map_cap = oc * 2; s->map = (map_slot *)calloc(s->map_cap, sizeof(map_slot)); s->map_len = 0;
    for (size_t i = 0; i < oc; ++i) if (old[i].used) map_insert(s, old[i].key, old[i].count, old[i].extra);
    free(old);
}
static void map_remove(fo_sketcher *s, uint64_t key) { /* backward-shift deletion */
    size_t cap = s->map_cap, i = map_home(key, cap);
    while (s->map[i].used && s->map[i].key != key) i = (i + 1) & (cap - 1);
    if (!s->map[i].used) return;
    s->map[i].used = 0; s->map_len--;
    size_t j = i;
    for (;;) {
        j = (j + 1) & (cap - 1);
        if (!s->map[j].used) break;
        size_t h = map_home(s->map[j].key, cap);
        /* can slot j move to the hole i?  yes iff h is cyclically outside (i, j] */
        int between = (i <= j) ? (i < h && h <= j) : (i < h || h <= j);
        if (!between) { s->map[i] = s->map[j]; s->map[j].used = 0; i = j; }
    }
}

static void heap_push(fo_sketcher *s, uint64_t hash, const uint8_t *kmer, size_t klen) {
    if (s->heap_len == s->heap_cap) {
        s->heap_cap = s->heap_cap ? s->heap_cap * 2 : 16;
        s->heap = (heap_item *)realloc(s->heap, s->heap_cap * sizeof(heap_item));
    }
    uint8_t *copy = (uint8_t *)malloc(klen ? klen : 1); /* kmer.to_owned() */
    memcpy(copy, kmer, klen);
    size_t i = s->heap_len++;
    while (i > 0) {
        size_t p = (i - 1) / 2;
        if (s->heap[p].hash >= hash) break;
        s->heap[i] = s->heap[p]; i = p;
    }
    s->heap[i].hash = hash; s->heap[i].kmer = copy;
}
static heap_item heap_pop(fo_sketcher *s) {
    heap_item top = s->heap[0], last = s->heap[--s->heap_len];
    size_t i = 0, n = s->heap_len;
    for (;;) {
        size_t c = 2 * i + 1;
        if (c >= n) break;
        if (c + 1 < n && s->heap[c + 1].hash > s->heap[c].hash) c++;
        if (s->heap[c].hash <= last.hash) break;
        s->heap[i] = s->heap[c]; i = c;
    }
    if (n) s->heap[i] = last;
    return top;
}

static fo_sketcher *sk_new(int kind, size_t size, uint8_t k, uint64_t seed, uint64_t max_hash) {
    fo_sketcher *s = (fo_sketcher *)calloc(1, sizeof(*s));
    s->kind = kind; s->size = size; s->k = k; s->seed = seed; s->max_hash = max_hash;
    s->map_cap = 1024; s->map = (map_slot *)calloc(s->map_cap, sizeof(map_slot));
    return s;
}
/* mash.rs:21-32 */
fo_sketcher *fo_mash_new(size_t size, uint8_t k, uint64_t seed) { return sk_new(0, size, k, seed, 0); }
/* scaled.rs:22-34: iscale = (1./scale) as u64 (saturating float->int cast); max_hash = u64::MAX / iscale */
fo_sketcher *fo_scaled_new(size_t size, double scale, uint8_t k, uint64_t seed) {
    double inv = 1.0 / scale;
    uint64_t iscale;
    if (!(inv == inv)) iscale = 0;                              /* NaN -> 0 (Rust `as`) */
    else if (inv >= 18446744073709551616.0) iscale = UINT64_MAX; /* saturate */
    else if (inv <= 0.0) iscale = 0;
    else iscale = (uint64_t)inv;
    /* Rust panics on division by zero (scale > 1 => iscale == 0); CLI limits scale to [0,1]. */
    uint64_t max_hash = iscale ? UINT64_MAX / iscale : UINT64_MAX;
    return sk_new(1, size, k, seed, max_hash);
}
void fo_sketcher_free(fo_sketcher *s) {
    if (!s) return;
    for (size_t i = 0; i < s->heap_len; ++i) free(s->heap[i].kmer);
    free(s->heap); free(s->map); free(s);
}
uint64_t fo_scaled_max_hash(const fo_sketcher *s) { return s->max_hash; }

static inline uint32_t sat_add_u32(uint32_t a, uint32_t b) {
    uint32_t r = a + b; return r < a ? UINT32_MAX : r;
}

/* mash.rs:34-63 and scaled.rs:37-61 */
void fo_push(fo_sketcher *s, const uint8_t *kmer, size_t klen, uint8_t extra_count) {
    s->total_kmers += 1;
    uint64_t new_hash = fo_hash_f(kmer, klen, s->seed);
    int add_hash;
    if (s->kind == 0) {
        if (s->heap_len == 0) add_hash = 1;                                    /* peek() == None */
        else add_hash = (new_hash <= s->heap[0].hash) || (s->heap_len < s->size);
    } else {
        add_hash = new_hash <= s->max_hash || (s->heap_len <= s->size && s->size != 0);
    }
    if (!add_hash) return;
    map_slot *slot = map_find(s, new_hash);
    if (slot) {
        slot->count = sat_add_u32(slot->count, 1);
        slot->extra = sat_add_u32(slot->extra, (uint32_t)extra_count);
    } else {
        heap_push(s, new_hash, kmer, klen);
        map_insert(s, new_hash, 1, (uint32_t)extra_count);
        int evict = (s->kind == 0) ? (s->heap_len > s->size)
                                   : (s->heap_len > s->size && s->heap[0].hash > s->max_hash);
        if (evict) {
            heap_item it = heap_pop(s);
            map_remove(s, it.hash);
            free(it.kmer);
        }
    }
}

static void process_cb(void *ctx, const uint8_t *kmer, int is_rc) {
    fo_sketcher *s = (fo_sketcher *)ctx;
    fo_push(s, kmer, s->k, (uint8_t)is_rc);
}
/* mash.rs:67-80 / scaled.rs:65-78 */
void fo_process(fo_sketcher *s, const uint8_t *raw_seq, size_t len) {
    s->total_bases += (uint64_t)len;   /* seq.sequence().len(): RAW record sequence bytes */
    for_canonical_kmers(raw_seq, len, s->k, process_cb, s);
}
void fo_totals(const fo_sketcher *s, uint64_t *tb, uint64_t *tk) { *tb = s->total_bases; *tk = s->total_kmers; }
size_t fo_result_len(const fo_sketcher *s) { return s->heap_len; }

static int cmp_heap_item(const void *a, const void *b) {
    uint64_t x = ((const heap_item *)a)->hash, y = ((const heap_item *)b)->hash;
    return x < y ? -1 : (x > y ? 1 : 0);
}
/* mash.rs:86-102 / scaled.rs:84-100: clone heap, into_sorted_vec (ascending), attach counts */
size_t fo_result(const fo_sketcher *s, uint64_t *hashes, uint32_t *counts, uint32_t *extras,
                 uint8_t *kmers, size_t kmer_stride) {
    size_t n = s->heap_len;
    heap_item *v = (heap_item *)malloc((n ? n : 1) * sizeof(heap_item));
    memcpy(v, s->heap, n * sizeof(heap_item));
    qsort(v, n, sizeof(heap_item), cmp_heap_item);
    for (size_t i = 0; i < n; ++i) {
        map_slot *slot = map_find((fo_sketcher *)s, v[i].hash);
        if (hashes) hashes[i] = v[i].hash;
        if (counts) counts[i] = slot->count;
        if (extras) extras[i] = slot->extra;
        if (kmers) {
            memset(kmers + i * kmer_stride, 0, kmer_stride);
            memcpy(kmers + i * kmer_stride, v[i].kmer, s->k < kmer_stride ? s->k : kmer_stride);
        }
    }
    free(v);
    return n;
}

/* ====================================================================================== */
/* FASTX records (needletail 0.5.0 parse_fastx_reader; call site lib/src/lib.rs:60-68).     */
/* sequence() of a record = RAW slice:                                                     */
/*   FASTA: header-newline+1 .. the record's last newline (exclusive), one trailing CR      */
/*          trimmed, interior line terminators KEPT;   FASTQ: the sequence line, CR trimmed.*/
/* ====================================================================================== */
static const uint8_t *find_nl(const uint8_t *p, const uint8_t *end) {
    return (const uint8_t *)memchr(p, '\n', (size_t)(end - p));
}
int fo_parse_fastx(const uint8_t *data, size_t len, fo_record_cb cb, void *ctx, int *format,
                   uint64_t *n_records) {
    uint64_t nrec = 0;
    if (format) *format = 0;
    if (n_records) *n_records = 0;
    if (len == 0) return FO_E_EMPTY;
    const uint8_t *end = data + len;
    if (data[0] == '>') {
        if (format) *format = FO_FMT_FASTA;
        const uint8_t *p = data;
        while (p < end) {
            const uint8_t *hnl = find_nl(p, end);
            if (!hnl) { cb(ctx, p, 0); nrec++; break; }        /* header only, no sequence */
            const uint8_t *seq = hnl + 1, *q = seq;
            /* next record: a '>' that directly follows a '\n' */
            const uint8_t *rec_end = end;
            while (q < end) {
                const uint8_t *g = (const uint8_t *)memchr(q, '>', (size_t)(end - q));
                if (!g) break;
                if (g[-1] == '\n') { rec_end = g; break; }
                q = g + 1;
            }
            const uint8_t *e = rec_end;
            if (e > seq && e[-1] == '\n') e--;
            if (e > seq && e[-1] == '\r') e--;
            cb(ctx, seq, (size_t)(e - seq)); nrec++;
            p = rec_end;
        }
    } else if (data[0] == '@') {
        if (format) *format = FO_FMT_FASTQ;
        const uint8_t *p = data;
        while (p < end) {
            int blank = 1;                                     /* trailing blank lines tolerated */
            for (const uint8_t *t = p; t < end; ++t) if (*t != '\n' && *t != '\r') { blank = 0; break; }
            if (blank) break;
            if (*p != '@') return FO_E_RECORD;
            const uint8_t *l1 = find_nl(p, end); if (!l1) return FO_E_RECORD;
            const uint8_t *seq = l1 + 1;
            const uint8_t *l2 = find_nl(seq, end); if (!l2) return FO_E_RECORD;
            const uint8_t *sep = l2 + 1;
            if (sep >= end || *sep != '+') return FO_E_RECORD;
            const uint8_t *l3 = find_nl(sep, end); if (!l3) return FO_E_RECORD;
            const uint8_t *qual = l3 + 1;
            const uint8_t *l4 = find_nl(qual, end); if (!l4) l4 = end;
            const uint8_t *se = l2, *qe = l4;
            if (se > seq && se[-1] == '\r') se--;
            if (qe > qual && qe[-1] == '\r') qe--;
            if ((se - seq) != (qe - qual)) return FO_E_RECORD;
            cb(ctx, seq, (size_t)(se - seq)); nrec++;
            p = (l4 < end) ? l4 + 1 : end;
        }
    } else {
        return FO_E_FORMAT;
    }
    if (n_records) *n_records = nrec;
    return nrec ? FO_OK : FO_E_EMPTY;
}

/* ====================================================================================== */
/* statistics.rs:30-47 hist ; filtering.rs                                                  */
/* ====================================================================================== */
uint64_t fo_hist(const uint32_t *counts, size_t n, uint64_t *out, size_t cap) {
    uint64_t max_count = 0;
    for (size_t i = 0; i < n; ++i) if (counts[i] > max_count) max_count = counts[i];
    if (out) {
        for (size_t i = 0; i < cap && i < max_count; ++i) out[i] = 0;
        for (size_t i = 0; i < n; ++i) {
            size_t idx = (size_t)counts[i] - 1;   /* count == 0 would underflow in the reference */
            if (counts[i] && idx < cap) out[idx]++;
        }
    }
    return max_count;
}

/* filtering.rs:154-195 */
uint32_t fo_guess_filter_threshold(const uint32_t *counts, size_t n, double filter_level) {
    uint64_t max_count = fo_hist(counts, n, NULL, 0);
    size_t hl = (size_t)max_count;
    uint64_t *hist = (uint64_t *)calloc(hl ? hl : 1, sizeof(uint64_t));
    fo_hist(counts, n, hist, hl);
    uint64_t total = 0;
    for (size_t i = 0; i < hl; ++i) total += ((uint64_t)i + 1) * hist[i];
    double total_counts = (double)total;
    double cutoff_amt = filter_level * total_counts;

    size_t wgt_cutoff = 0; uint64_t cum_count = 0;
    for (size_t i = 0; i < hl; ++i) {
        cum_count += (uint64_t)wgt_cutoff * hist[i];
        if ((double)cum_count > cutoff_amt) break;
        wgt_cutoff += 1;
    }
    if (wgt_cutoff == 0) { free(hist); return 1; }
    size_t win_size = wgt_cutoff / 20; if (win_size < 1) win_size = 1;
    uint64_t sum = 0;
    for (size_t i = 0; i < win_size; ++i) sum += hist[i];
    uint64_t lowest_val = sum; size_t lowest_idx = win_size - 1;
    /* (0..wgt_cutoff - win_size).zip(win_size..wgt_cutoff) */
    for (size_t i = 0, j = win_size; i < wgt_cutoff - win_size && j < wgt_cutoff; ++i, ++j) {
        if (sum <= lowest_val) { lowest_val = sum; lowest_idx = j; }
        sum -= hist[i]; sum += hist[j];
    }
    free(hist);
    return (uint32_t)lowest_idx + 1;
}

/* filtering.rs:413-432 */
size_t fo_filter_strands(const uint32_t *counts, const uint32_t *extras, size_t n,
                         double ratio_cutoff, uint32_t *keep) {
    size_t m = 0;
    for (size_t i = 0; i < n; ++i) {
        if (counts[i] < 16) { keep[m++] = (uint32_t)i; continue; }
        uint32_t other = counts[i] - extras[i];   /* u32 subtraction; extra <= count always */
        uint32_t lowest = extras[i] < other ? extras[i] : other;
        if ((double)lowest / (double)counts[i] >= ratio_cutoff) keep[m++] = (uint32_t)i;
    }
    return m;
}
/* filtering.rs:329-343 */
size_t fo_filter_abundance(const uint32_t *counts, size_t n, int has_low, uint32_t low,
                           int has_high, uint32_t high, uint32_t *keep) {
    uint32_t lo = has_low ? low : 0u, hi = has_high ? high : UINT32_MAX;
    size_t m = 0;
    for (size_t i = 0; i < n; ++i) if (lo <= counts[i] && counts[i] <= hi) keep[m++] = (uint32_t)i;
    return m;
}
/* filtering.rs:60-87 */
size_t fo_filter_counts(fo_filter_params *fp, const uint32_t *counts, const uint32_t *extras,
                        size_t n, uint32_t *keep) {
    int filter_on = fp->filter_on == 1;
    uint32_t *idx = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    uint32_t *c = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    uint32_t *x = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    uint32_t *tmp = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    size_t m = n;
    for (size_t i = 0; i < n; ++i) { idx[i] = (uint32_t)i; c[i] = counts[i]; x[i] = extras[i]; }
#define APPLY_KEEP(mm)                                                             \
    do { for (size_t i_ = 0; i_ < (mm); ++i_) { uint32_t s_ = tmp[i_];             \
             idx[i_] = idx[s_]; c[i_] = c[s_]; x[i_] = x[s_]; } m = (mm); } while (0)
    if (filter_on && fp->strand_filter > 0.0) {
        size_t mm = fo_filter_strands(c, x, m, fp->strand_filter, tmp);
        APPLY_KEEP(mm);
    }
    if (filter_on && fp->err_filter > 0.0) {
        uint32_t cutoff = fo_guess_filter_threshold(c, m, fp->err_filter);
        if (fp->has_abun_low) { if (cutoff > fp->abun_low) fp->abun_low = cutoff; }
        else { fp->has_abun_low = 1; fp->abun_low = cutoff; }
    }
    if (filter_on && (fp->has_abun_low || fp->has_abun_high)) {
        size_t mm = fo_filter_abundance(c, m, fp->has_abun_low, fp->abun_low, fp->has_abun_high,
                                        fp->abun_high, tmp);
        APPLY_KEEP(mm);
    }
#undef APPLY_KEEP
    for (size_t i = 0; i < m; ++i) keep[i] = idx[i];
    free(idx); free(c); free(x); free(tmp);
    return m;
}

/* ====================================================================================== */
/* lib.rs:51-94 sketch_stream ; mod.rs:86-128 create_sketcher / process_post_filter          */
/* ====================================================================================== */
static void stream_record_cb(void *ctx, const uint8_t *raw, size_t len) { fo_process((fo_sketcher *)ctx, raw, len); }

int fo_sketch_stream(const uint8_t *data, size_t len, const fo_sketch_params *sp,
                     const fo_filter_params *fp_in, fo_sketch *out) {
    memset(out, 0, sizeof(*out));
    fo_filter_params fp = *fp_in;
    fo_sketcher *s = sp->kind == 0 ? fo_mash_new((size_t)sp->kmers_to_sketch, sp->kmer_length, sp->hash_seed)
                                   : fo_scaled_new((size_t)sp->kmers_to_sketch, sp->scale, sp->kmer_length, sp->hash_seed);
    int fmt = 0; uint64_t nrec = 0;
    int rc = fo_parse_fastx(data, len, stream_record_cb, s, &fmt, &nrec);
    if (rc != FO_OK) { fo_sketcher_free(s); return rc; }
    /* lib.rs:71-76: FASTA -> off, FASTQ -> on, unless explicit */
    if (fp.filter_on < 0) fp.filter_on = (fmt == FO_FMT_FASTQ) ? 1 : 0;
    fo_totals(s, &out->seq_length, &out->num_valid_kmers);
    size_t n = fo_result_len(s), k = sp->kmer_length;
    uint64_t *h = (uint64_t *)malloc((n ? n : 1) * 8);
    uint32_t *c = (uint32_t *)malloc((n ? n : 1) * 4), *x = (uint32_t *)malloc((n ? n : 1) * 4);
    uint8_t *km = (uint8_t *)malloc((n ? n : 1) * (k ? k : 1));
    fo_result(s, h, c, x, km, k);
    fo_sketcher_free(s);
    uint32_t *keep = (uint32_t *)malloc((n ? n : 1) * 4);
    size_t m = fo_filter_counts(&fp, c, x, n, keep);
    /* process_post_filter (mod.rs:115-128): Mash truncates to final_size, strictness check */
    if (sp->kind == 0) {
        if (m > sp->final_size) m = (size_t)sp->final_size;
        if (!sp->no_strict && m < sp->final_size) {
            free(h); free(c); free(x); free(km); free(keep);
            return FO_E_TOO_FEW;
        }
    }
    out->n = m;
    out->hashes = (uint64_t *)malloc((m ? m : 1) * 8);
    out->counts = (uint32_t *)malloc((m ? m : 1) * 4);
    out->extras = (uint32_t *)malloc((m ? m : 1) * 4);
    out->kmers = (uint8_t *)malloc((m ? m : 1) * (k ? k : 1));
    for (size_t i = 0; i < m; ++i) {
        uint32_t s_ = keep[i];
        out->hashes[i] = h[s_]; out->counts[i] = c[s_]; out->extras[i] = x[s_];
        memcpy(out->kmers + i * k, km + (size_t)s_ * k, k);
    }
    out->filters = fp; out->format = fmt;
    free(h); free(c); free(x); free(km); free(keep);
    return FO_OK;
}
void fo_sketch_free(fo_sketch *sk) {
    free(sk->hashes); free(sk->counts); free(sk->extras); free(sk->kmers);
    memset(sk, 0, sizeof(*sk));
}

/* ====================================================================================== */
/* distance.rs:66-126 raw_distance ; distance.rs:35-41 mash distance                        */
/* ====================================================================================== */
void fo_raw_distance(const uint64_t *q, size_t nq, const uint64_t *r, size_t nr, double scale,
                     double *containment, double *jaccard, uint64_t *common_out, uint64_t *total_out) {
    size_t i = 0, j = 0; uint64_t common = 0;
    while (i < nq && j < nr) {
        if (q[i] < r[j]) i++;
        else if (q[i] > r[j]) j++;
        else { common++; i++; j++; }
    }
    if (scale > 0.0) {
        double rec = 1.0 / scale;                       /* scale.recip() as u64 (saturating) */
        uint64_t d = rec >= 18446744073709551616.0 ? UINT64_MAX : (uint64_t)rec;
        uint64_t max_hash = d ? UINT64_MAX / d : UINT64_MAX;
        while (i < nq && q[i] < max_hash) i++;
        while (j < nr && r[j] < max_hash) j++;
    }
    *containment = (j == 0) ? 0.0 : (double)common / (double)j;
    uint64_t total = (uint64_t)i - common + (uint64_t)j;
    *jaccard = (total == 0) ? 1.0 : (double)common / (double)total;
    *common_out = common; *total_out = total;
}
/* distance.rs:136-157 old_distance, literally (the pointer walk with its `i < len - 1` guard).
 * Returns -1 where the reference panics (empty query: index out of bounds). */
int fo_old_distance(const uint64_t *q, size_t nq, const uint64_t *r, size_t nr,
                    double *containment, double *jaccard, uint64_t *common_out, uint64_t *total_out) {
    size_t i = 0; uint64_t common = 0, total = 0;
    for (size_t x = 0; x < nr; ++x) {
        if (nq == 0) return -1;
        while (q[i] < r[x] && i < nq - 1) i++;
        if (q[i] == r[x]) common++;
        total++;
    }
    *containment = (double)common / (double)total;                               /* NaN for an empty reference */
    *jaccard = (double)common / (double)(common + 2 * (total - common));
    *common_out = common; *total_out = total;
    return 0;
}
double fo_mash_distance(double jaccard, uint8_t k) {
    double md = -1.0 * log((2.0 * jaccard) / (1.0 + jaccard)) / (double)k;
    /* f64::min(1, f64::max(0, md)) -- Rust's max/min ignore NaN operands */
    double m = (md != md) ? 0.0 : (md > 0.0 ? md : 0.0);
    return m < 1.0 ? m : 1.0;
}

/* ---- AllCountsSketcher (lib/src/sketch_schemes/counts.rs:7-70) ------------------------------------------------
 * counts has 4^k entries.  process (counts.rs:24-36): seq.normalize(false).bit_kmers(k, false) -- needletail 0.5
 * BitNuclKmer, canonical = false: one (position, forward BitKmer) per window of k bases that are all ACGT, the
 * BitKmer's integer holding the first base in its highest pair (A, C, G, T = 0..3); counts[kmer] saturating += 1. */
void fo_allcounts_process(uint32_t *counts, uint8_t k, const uint8_t *raw_seq, size_t len) {
    uint8_t *norm = (uint8_t *)malloc(len ? len : 1);
    const size_t n = fo_normalize(raw_seq, len, norm);
    const uint64_t mask = k >= 32 ? ~0ULL : ((1ULL << (2 * k)) - 1ULL);
    uint64_t fwd = 0;
    size_t run = 0;
    for (size_t i = 0; i < n; ++i) {
        int c;
        switch (norm[i]) { case 'A': c = 0; break; case 'C': c = 1; break; case 'G': c = 2; break; case 'T': c = 3; break; default: c = -1; }
        if (c < 0) { run = 0; continue; }
        fwd = ((fwd << 2) | (uint64_t)c) & mask;
        if (++run >= k) { if (counts[fwd] != 0xFFFFFFFFu) counts[fwd]++; }
    }
    free(norm);
}
static uint64_t fo_bit_revcomp(uint64_t ix, uint8_t k) {            /* needletail bitkmer::reverse_complement */
    uint64_t r = 0;
    for (uint8_t i = 0; i < k; ++i) { r = (r << 2) | (3 - (ix & 3)); ix >>= 2; }
    return r;
}
/* to_vec (counts.rs:45-63), literally: a working copy is zeroed at the reverse complement of every emitted index.
 * Returns the number of entries; kmers: k bytes each (bitmer_to_bytes: first base from the highest pair). */
size_t fo_allcounts_to_vec(const uint32_t *self_counts, uint8_t k, uint64_t *hashes, uint32_t *cnt, uint32_t *ext,
                           uint8_t *kmers, size_t cap) {
    const uint64_t n = 1ULL << (2 * k);
    uint32_t *counts = (uint32_t *)malloc((size_t)n * 4);
    memcpy(counts, self_counts, (size_t)n * 4);
    size_t m = 0;
    for (uint64_t ix = 0; ix < n; ++ix) {
        uint32_t count = counts[ix];
        if (count == 0) continue;
        const uint64_t rc = fo_bit_revcomp(ix, k);
        const uint32_t extra = self_counts[rc];
        counts[rc] = 0;
        count += extra;                                             /* wraps in a release build */
        if (m < cap) {
            hashes[m] = ix; cnt[m] = count; ext[m] = extra;
            for (uint8_t i = 0; i < k; ++i) kmers[m * k + i] = (uint8_t)"ACGT"[(ix >> (2 * (k - 1 - i))) & 3];
        }
        ++m;
    }
    free(counts);
    return m;
}
uint64_t fo_allcounts_total(const uint32_t *counts, uint8_t k) {    /* total_bases_and_kmers().1 (counts.rs:38-43) */
    uint64_t t = 0;
    for (uint64_t ix = 0; ix < (1ULL << (2 * k)); ++ix) t += counts[ix];
    return t;
}

/* ---- minmer_matrix (lib/src/distance.rs:344-364), literally: one row per sketch, one column per reference hash, the
 * sketch's COUNT where it holds the hash (the pointer walk never passes the last reference entry). */
void fo_minmer_matrix(const uint64_t *ref_hashes, size_t n_ref, const uint64_t *const *sk_hashes, const uint32_t *const *sk_counts,
                      const size_t *sk_len, size_t n_sk, int32_t *result /* n_sk x n_ref, zeroed here */) {
    memset(result, 0, n_sk * n_ref * sizeof(int32_t));
    if (n_ref == 0) return;                                         /* (the reference would index out of bounds) */
    for (size_t i = 0; i < n_sk; ++i) {
        size_t ref_pos = 0;
        for (size_t j = 0; j < sk_len[i]; ++j) {
            const uint64_t h = sk_hashes[i][j];
            while (h > ref_hashes[ref_pos] && ref_pos < n_ref - 1) ref_pos++;
            if (h == ref_hashes[ref_pos]) result[i * n_ref + ref_pos] = (int32_t)sk_counts[i][j];
        }
    }
}

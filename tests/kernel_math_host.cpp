// TEST HELPER (not product code): host build of the inline arithmetic in
// finch_rs_b200/csrc/common.cuh so the exact functions the kernels call can be checked against
// the oracle on a machine without a GPU.  Built by tests/conftest.py with g++.
#include <cstddef>
#include <cstdint>
#include "../finch_rs_b200/csrc/common.cuh"

extern "C" {
// Walk a symbol stream the way hash_kernel does: roll, canonical choice, murmur.
size_t km_stream(const uint8_t *sym, size_t n, int k, uint64_t seed, uint64_t *h, uint8_t *rc, uint64_t *codes) {
    fb2::Roll r; r.fwd = 0; r.rc = 0; r.run = 0;
    const uint64_t mask = fb2::kmer_mask(k);
    size_t m = 0;
    for (size_t i = 0; i < n; ++i) {
        fb2::roll_push(r, sym[i], k, mask);
        if (r.run >= (uint32_t)k) {
            bool is_rc;
            const uint64_t c = fb2::roll_canonical_lsb(r, mask, is_rc);
            uint64_t hv;
            if (k == 21) hv = fb2::murmur_kmer_h1<21>(c, k, seed);
            else if (k == 31) hv = fb2::murmur_kmer_h1<31>(c, k, seed);
            else hv = fb2::murmur_kmer_h1<0>(c, k, seed);
            h[m] = hv; rc[m] = is_rc; codes[m] = c; ++m;
        }
    }
    return m;
}
uint8_t km_classify(uint8_t c) { return fb2::classify_byte(c); }
uint64_t km_murmur_bytes(const uint8_t *p, uint32_t n, uint64_t seed) { return fb2::murmur_bytes_h1(p, n, seed); }
void km_codes_to_ascii(uint64_t codes, int k, uint8_t *out) { fb2::codes_to_ascii(codes, k, out); }
}

// finch_b200.hpp -- header-only C++ mirror of the finch-rs sketching surface over the C ABI
// (include/finch_b200.h).  Same names / argument meaning as the reference:
//   SketchParams::create_sketcher  (lib/src/sketch_schemes/mod.rs:86-113)
//   SketchScheme::{process,total_bases_and_kmers,to_vec}  (mod.rs:24-51)
//   MashSketcher::push / ScaledSketcher::push             (mash.rs:34, scaled.rs:37)
//   sketch_stream                                          (lib/src/lib.rs:51-94)
// Compiled and run by tests/hpp_mirror_test.cpp (tests/test_abi_cpu.py builds it, tests/test_gpu_parity.py runs it).
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/finch_b200.h"

namespace finch {

struct FinchError : std::runtime_error {   // lib/src/errors.rs:5-23
    int code;
    FinchError(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) { if (rc != FB2_OK) throw FinchError(rc, fb2_last_error()); }

struct KmerCount {                         // mod.rs:15-22
    uint64_t hash;
    std::vector<uint8_t> kmer;
    uint32_t count, extra_count;
};

class SketchScheme {                       // mod.rs:24-51
public:
    explicit SketchScheme(const fb2_params &p) : params_(p) { check(fb2_sketcher_create(&p, &h_)); }
    ~SketchScheme() { fb2_sketcher_destroy(h_); }
    SketchScheme(const SketchScheme &) = delete;
    SketchScheme &operator=(const SketchScheme &) = delete;

    void process(const uint8_t *raw_seq, size_t len) { check(fb2_sketcher_process(h_, raw_seq, len)); }
    void push(const uint8_t *kmer, size_t k, uint8_t extra_count) { check(fb2_sketcher_push(h_, kmer, k, extra_count)); }
    void feed_fastx(const uint8_t *bytes, size_t len, bool last) { check(fb2_sketcher_feed_fastx(h_, bytes, len, last)); }
    std::pair<uint64_t, uint64_t> total_bases_and_kmers() const {
        uint64_t b = 0, k = 0;
        check(fb2_sketcher_totals(h_, &b, &k));
        return {b, k};
    }
    std::vector<KmerCount> to_vec() const {
        fb2_result r;
        check(fb2_sketcher_result(h_, &r));
        std::vector<KmerCount> v(r.n);
        for (uint64_t i = 0; i < r.n; ++i) {
            v[i].hash = r.hashes[i]; v[i].count = r.counts[i]; v[i].extra_count = r.extras[i];
            const size_t kl = r.kmer_lens ? r.kmer_lens[i] : params_.kmer_length;   // pushed k-mers keep their own length
            v[i].kmer.assign(r.kmers + i * r.kmer_stride, r.kmers + i * r.kmer_stride + kl);
        }
        fb2_result_free(&r);
        return v;
    }
    // SketchScheme::parameters as the reference's sketchers answer it (quirks Q7 / Q8): MashSketcher reports
    // final_size = size, no_strict = false (mash.rs:104-112); ScaledSketcher recomputes the scale from its integer
    // max_hash (scaled.rs:23,31,102-109).
    fb2_params parameters() const {
        fb2_params p = params_;
        if (p.kind == FB2_KIND_MASH) { p.final_size = p.kmers_to_sketch; p.no_strict = 0; }
        else {
            const uint64_t iscale = (uint64_t)(1.0 / p.scale);
            const uint64_t max_hash = UINT64_MAX / iscale;
            p.scale = 1.0 / ((double)UINT64_MAX / (double)max_hash);
        }
        return p;
    }

private:
    fb2_sketcher *h_ = nullptr;
    fb2_params params_;
};

struct SketchParams {
    static fb2_params Mash(uint64_t kmers_to_sketch, uint64_t final_size, bool no_strict, uint8_t k, uint64_t seed) {
        return fb2_params{FB2_KIND_MASH, kmers_to_sketch, final_size, no_strict, k, seed, 0.0, -1, nullptr};
    }
    static fb2_params Scaled(uint64_t kmers_to_sketch, uint8_t k, double scale, uint64_t seed) {
        return fb2_params{FB2_KIND_SCALED, kmers_to_sketch, 0, 0, k, seed, scale, -1, nullptr};
    }
    static std::unique_ptr<SketchScheme> create_sketcher(const fb2_params &p) { return std::make_unique<SketchScheme>(p); }
};
inline std::unique_ptr<SketchScheme> MashSketcher(size_t size, uint8_t k, uint64_t seed) {
    return SketchParams::create_sketcher(SketchParams::Mash(size, size, false, k, seed));
}
inline std::unique_ptr<SketchScheme> ScaledSketcher(size_t size, double scale, uint8_t k, uint64_t seed) {
    return SketchParams::create_sketcher(SketchParams::Scaled(size, k, scale, seed));
}

}  // namespace finch

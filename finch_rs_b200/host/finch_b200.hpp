// finch_b200.hpp -- header-only C++ mirror of the finch-rs sketching surface over the C ABI
// (include/finch_b200.h).  Same names / argument meaning as the reference:
//   SketchParams::create_sketcher  (lib/src/sketch_schemes/mod.rs:86-113)
//   SketchScheme::{process,total_bases_and_kmers,to_vec}  (mod.rs:24-51)
//   MashSketcher::push / ScaledSketcher::push             (mash.rs:34, scaled.rs:37)
//   sketch_stream                                          (lib/src/lib.rs:51-94)
//   open_sketch_file / write_finch_file / write_mash_file  (lib/src/lib.rs:96-117, serialization/)
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

// ---- sketch files: Sketch (serialization/mod.rs:45-55), open_sketch_file (lib.rs:96-117), write_finch_file
// (serialization/mod.rs:123-176), write_mash_file (mash.rs:12-71), the `.sk` JSON (json.rs:64-89).  Host only. ----
struct Sketch {
    std::string name, comment;
    uint64_t seq_length = 0, num_valid_kmers = 0;
    std::vector<KmerCount> hashes;
    fb2_params sketch_params{};
    fb2_filter filter_params{};
};
inline std::vector<Sketch> open_sketch_file(const std::string &path) {
    fb2_sketch_set *set = nullptr;
    check(fb2_sketch_set_open(path.c_str(), &set));
    std::unique_ptr<fb2_sketch_set, void (*)(fb2_sketch_set *)> guard(set, fb2_sketch_set_close);
    std::vector<Sketch> out(fb2_sketch_set_len(set));
    for (size_t i = 0; i < out.size(); ++i) {
        fb2_sketch_view v;
        check(fb2_sketch_set_get(set, i, &v));
        Sketch &s = out[i];
        s.name = v.name; s.comment = v.comment; s.seq_length = v.seq_length; s.num_valid_kmers = v.num_valid_kmers;
        s.sketch_params = v.params; s.filter_params = v.filter;
        s.hashes.resize(v.n);
        for (uint64_t q = 0; q < v.n; ++q) {
            s.hashes[q].hash = v.hashes[q]; s.hashes[q].count = v.counts[q]; s.hashes[q].extra_count = v.extras[q];
            s.hashes[q].kmer.assign(v.kmers + v.kmer_offs[q], v.kmers + v.kmer_offs[q + 1]);
        }
    }
    return out;
}
inline void write_sketch_file(const std::string &path, const std::vector<Sketch> &sketches, int file_format) {
    fb2_sketch_set *set = nullptr;
    check(fb2_sketch_set_new(&set));
    std::unique_ptr<fb2_sketch_set, void (*)(fb2_sketch_set *)> guard(set, fb2_sketch_set_close);
    for (const Sketch &s : sketches) {
        std::vector<uint64_t> h(s.hashes.size()), offs(s.hashes.size() + 1, 0);
        std::vector<uint32_t> c(s.hashes.size()), x(s.hashes.size());
        std::vector<uint8_t> bytes;
        for (size_t q = 0; q < s.hashes.size(); ++q) {
            h[q] = s.hashes[q].hash; c[q] = s.hashes[q].count; x[q] = s.hashes[q].extra_count;
            offs[q] = bytes.size();
            bytes.insert(bytes.end(), s.hashes[q].kmer.begin(), s.hashes[q].kmer.end());
        }
        offs[s.hashes.size()] = bytes.size();
        fb2_sketch_view v{};
        v.name = s.name.c_str(); v.comment = s.comment.c_str(); v.seq_length = s.seq_length; v.num_valid_kmers = s.num_valid_kmers;
        v.params = s.sketch_params; v.filter = s.filter_params; v.n = h.size();
        v.hashes = h.data(); v.counts = c.data(); v.extras = x.data(); v.kmers = bytes.data(); v.kmer_offs = offs.data();
        check(fb2_sketch_set_add(set, &v));
    }
    check(fb2_sketch_set_save(set, path.c_str(), file_format));
}
inline void write_finch_file(const std::string &path, const std::vector<Sketch> &s) { write_sketch_file(path, s, FB2_FILE_BSK); }
inline void write_mash_file(const std::string &path, const std::vector<Sketch> &s) { write_sketch_file(path, s, FB2_FILE_MSH); }

}  // namespace finch

// sketchfile.cpp -- sketch files behind the C ABI: collections of finished sketches read from / written to the three
// formats of the reference.  Host only (no device is touched); the codecs are the ones the command line uses
// (host/sketch_json.hpp, host/sketch_capnp.hpp).
//
//   fb2_sketch_set_open   <- open_sketch_file                (lib/src/lib.rs:96-117: format by file suffix)
//   fb2_sketch_set_save   <- write_finch_file                (lib/src/serialization/mod.rs:123-176; Multisketch.save, python.rs:180-186)
//                            write_mash_file                 (lib/src/serialization/mash.rs:12-71)
//                            MultiSketch::from_sketches JSON (lib/src/serialization/json.rs:64-89)
//   fb2_sketch_set_get / _add: the `Sketch` struct           (lib/src/serialization/mod.rs:45-55) as a view of plain arrays
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/finch_b200.h"
#include "../host/sketch_capnp.hpp"
#include "../host/sketch_json.hpp"

int fb2_fail(int code, const std::string &msg);  // engine.cu

using fb2host::Sketch;

struct fb2_sketch_set {
    std::vector<Sketch> sketches;
    // flattened k-mer bytes of the sketch last handed out by _get (the view points into these)
    mutable std::vector<uint8_t> kmer_bytes;
    mutable std::vector<uint64_t> kmer_offs;
};

namespace {

bool ends_with(const std::string &s, const char *suf) {
    const size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}
std::string file_name_of(const std::string &path) {
    const size_t p = path.find_last_of('/');
    return p == std::string::npos ? path : path.substr(p + 1);
}
void params_to_c(const fb2host::SketchParams &p, fb2_params *o) {
    memset(o, 0, sizeof(*o));
    o->kind = p.kind == fb2host::Kind::Mash ? FB2_KIND_MASH : (p.kind == fb2host::Kind::Scaled ? FB2_KIND_SCALED : FB2_KIND_ALLCOUNTS);
    o->kmers_to_sketch = p.kmers_to_sketch; o->final_size = p.final_size; o->no_strict = p.no_strict ? 1 : 0;
    o->kmer_length = p.kmer_length; o->hash_seed = p.hash_seed; o->scale = p.scale;
    o->device = -1; o->stream = nullptr;
}
void params_from_c(const fb2_params &c, fb2host::SketchParams *p) {
    *p = fb2host::SketchParams();
    p->kind = c.kind == FB2_KIND_MASH ? fb2host::Kind::Mash : (c.kind == FB2_KIND_SCALED ? fb2host::Kind::Scaled : fb2host::Kind::AllCounts);
    p->kmer_length = c.kmer_length;
    if (c.kind == FB2_KIND_MASH) { p->kmers_to_sketch = c.kmers_to_sketch; p->final_size = c.final_size; p->no_strict = c.no_strict != 0; p->hash_seed = c.hash_seed; }
    else if (c.kind == FB2_KIND_SCALED) { p->kmers_to_sketch = c.kmers_to_sketch; p->scale = c.scale; p->hash_seed = c.hash_seed; }
}
void filter_to_c(const fb2host::FilterParams &f, fb2_filter *o) {
    memset(o, 0, sizeof(*o));
    o->filter_on = f.filter_on;
    o->has_abun_low = f.has_lo ? 1 : 0; o->abun_low = f.lo;
    o->has_abun_high = f.has_hi ? 1 : 0; o->abun_high = f.hi;
    o->err_filter = f.err_filter; o->strand_filter = f.strand_filter;
}
void filter_from_c(const fb2_filter &c, fb2host::FilterParams *f) {
    *f = fb2host::FilterParams();
    f->filter_on = c.filter_on < 0 ? -1 : (c.filter_on ? 1 : 0);
    f->has_lo = c.has_abun_low != 0; f->lo = f->has_lo ? c.abun_low : 0;
    f->has_hi = c.has_abun_high != 0; f->hi = f->has_hi ? c.abun_high : 0;
    f->err_filter = c.err_filter; f->strand_filter = c.strand_filter;
}

}  // namespace

extern "C" int fb2_sketch_set_new(fb2_sketch_set **out) {
    if (!out) return fb2_fail(FB2_EINVAL, "null argument");
    *out = new (std::nothrow) fb2_sketch_set();
    return *out ? FB2_OK : fb2_fail(FB2_ENOMEM, "out of memory");
}

extern "C" int fb2_sketch_set_open(const char *path, fb2_sketch_set **out) {
    if (!path || !out) return fb2_fail(FB2_EINVAL, "null argument");
    *out = nullptr;
    const std::string p = path, fname = file_name_of(p);
    if (fname.empty()) return fb2_fail(FB2_EINVAL, "Path does not have a filename: \"" + p + "\"");
    std::ifstream in(p, std::ios::binary);
    if (!in) return fb2_fail(FB2_EIO, "Error opening \"" + p + "\"");
    // the suffix decides (lib.rs:104-116); anything else is refused before a byte is parsed
    const bool msh = ends_with(fname, ".msh"), bsk = ends_with(fname, ".bsk"), sk = ends_with(fname, ".sk") || ends_with(fname, ".json");
    if (!msh && !bsk && !sk) return fb2_fail(FB2_EFORMAT, "File suffix is not *.bsk, *.msh, or *.sk");
    std::ostringstream ss;
    ss << in.rdbuf();
    const std::string data = ss.str();
    fb2_sketch_set *set = new (std::nothrow) fb2_sketch_set();
    if (!set) return fb2_fail(FB2_ENOMEM, "out of memory");
    try {
        if (msh) set->sketches = fb2host::read_mash_file(data.data(), data.size());
        else if (bsk) set->sketches = fb2host::read_finch_file(data.data(), data.size());
        else set->sketches = fb2host::read_multisketch_json(data.data(), data.size());
    } catch (const std::exception &e) {
        delete set;
        return fb2_fail(FB2_EFORMAT, "Error parsing \"" + p + "\" (" + e.what() + ")");
    }
    *out = set;
    return FB2_OK;
}

extern "C" uint64_t fb2_sketch_set_len(const fb2_sketch_set *set) { return set ? set->sketches.size() : 0; }

extern "C" int fb2_sketch_set_get(const fb2_sketch_set *set, uint64_t i, fb2_sketch_view *out) {
    if (!set || !out) return fb2_fail(FB2_EINVAL, "null argument");
    if (i >= set->sketches.size()) return fb2_fail(FB2_EINVAL, "sketch index out of range");
    const Sketch &s = set->sketches[i];
    memset(out, 0, sizeof(*out));
    out->name = s.name.c_str(); out->comment = s.comment.c_str();
    out->seq_length = s.seq_length; out->num_valid_kmers = s.num_valid_kmers;
    params_to_c(s.sketch_params, &out->params);
    filter_to_c(s.filter_params, &out->filter);
    out->n = s.hashes.size();
    out->hashes = s.hashes.data(); out->counts = s.counts.data(); out->extras = s.extras.data();
    set->kmer_offs.assign(s.hashes.size() + 1, 0);
    size_t total = 0;
    for (size_t q = 0; q < s.hashes.size(); ++q) { set->kmer_offs[q] = total; total += q < s.kmers.size() ? s.kmers[q].size() : 0; }
    set->kmer_offs[s.hashes.size()] = total;
    set->kmer_bytes.resize(total + 1);
    for (size_t q = 0; q < s.hashes.size() && q < s.kmers.size(); ++q)
        if (!s.kmers[q].empty()) memcpy(set->kmer_bytes.data() + set->kmer_offs[q], s.kmers[q].data(), s.kmers[q].size());
    out->kmers = set->kmer_bytes.data(); out->kmer_offs = set->kmer_offs.data();
    return FB2_OK;
}

extern "C" int fb2_sketch_set_add(fb2_sketch_set *set, const fb2_sketch_view *v) {
    if (!set || !v) return fb2_fail(FB2_EINVAL, "null argument");
    if (v->n && (!v->hashes || !v->counts || !v->extras)) return fb2_fail(FB2_EINVAL, "null hash / count arrays");
    Sketch s;
    s.name = v->name ? v->name : ""; s.comment = v->comment ? v->comment : "";
    s.seq_length = v->seq_length; s.num_valid_kmers = v->num_valid_kmers;
    params_from_c(v->params, &s.sketch_params);
    filter_from_c(v->filter, &s.filter_params);
    s.hashes.assign(v->hashes, v->hashes + v->n);
    s.counts.assign(v->counts, v->counts + v->n);
    s.extras.assign(v->extras, v->extras + v->n);
    s.kmers.resize(v->n);
    if (v->kmers && v->kmer_offs)
        for (uint64_t q = 0; q < v->n; ++q) {
            if (v->kmer_offs[q + 1] < v->kmer_offs[q]) return fb2_fail(FB2_EINVAL, "k-mer offsets must not decrease");
            s.kmers[q].assign(reinterpret_cast<const char *>(v->kmers) + v->kmer_offs[q], v->kmer_offs[q + 1] - v->kmer_offs[q]);
        }
    set->sketches.push_back(std::move(s));
    return FB2_OK;
}

extern "C" int fb2_sketch_set_remove(fb2_sketch_set *set, uint64_t i) {
    if (!set) return fb2_fail(FB2_EINVAL, "null argument");
    if (i >= set->sketches.size()) return fb2_fail(FB2_EINVAL, "sketch index out of range");
    set->sketches.erase(set->sketches.begin() + (ptrdiff_t)i);
    return FB2_OK;
}

extern "C" int fb2_sketch_set_save(const fb2_sketch_set *set, const char *path, int file_format) {
    if (!set || !path) return fb2_fail(FB2_EINVAL, "null argument");
    std::string bytes;
    try {
        if (file_format == FB2_FILE_BSK) bytes = fb2host::write_finch_file(set->sketches);
        else if (file_format == FB2_FILE_MSH) bytes = fb2host::write_mash_file(set->sketches, fb2host::params_from_sketches(set->sketches));
        else if (file_format == FB2_FILE_SK) bytes = fb2host::write_multisketch_json(set->sketches);
        else return fb2_fail(FB2_EINVAL, "unknown sketch file format");
    } catch (const std::exception &e) {
        return fb2_fail(FB2_EINVAL, e.what());
    }
    std::ofstream o(path, std::ios::binary | std::ios::trunc);
    if (!o) return fb2_fail(FB2_EIO, std::string("Could not create ") + path);
    o.write(bytes.data(), (std::streamsize)bytes.size());
    o.close();
    if (!o) return fb2_fail(FB2_EIO, std::string("Could not write ") + path);
    return FB2_OK;
}

extern "C" void fb2_sketch_set_close(fb2_sketch_set *set) { delete set; }

// sketch_capnp.hpp -- `.bsk` (finch schema) and `.msh` (Mash schema) sketch files: Cap'n Proto messages, written and
// read without the capnp library or its code generator (neither is in this image).  SURVEY 8f row N4.
//
// Mirrors
//   write_finch_file / read_finch_file     lib/src/serialization/mod.rs:123-224   (schema finch.capnp:1-56)
//   write_mash_file  / read_mash_file      lib/src/serialization/mash.rs:12-132   (schema mash.capnp)
// Struct layouts are the ones capnpc generated for the reference (finch_capnp.rs / mash_capnp.rs):
//   Multisketch 0/1            ptr0 sketches: List(Sketch)
//   Sketch      2/5            u64@0 seqLength, u64@1 numValidKmers; ptr0 name, ptr1 comment, ptr2 hashes: List(KmerCount),
//                              ptr3 filterParams, ptr4 sketchParams
//   KmerCount   2/2            u64@0 hash, u32@2 count, u32@3 extraCount; ptr0 kmer: Data, ptr1 label: Data
//   FilterParams 4/0           bool@0 filtered, u32@1 lowAbunFilter, u32@2 highAbunFilter, f64@2 errFilter, f64@3 strandFilter
//   SketchParams 5/0           u16@0 sketchMethod, u8@2 kmerLength, u64@1 kmersToSketch, u64@2 hashSeed, u64@3 finalSize,
//                              bool@24 noStrict, f64@4 scale
//   MinHash     3/4            u32@0 kmerSize, u32@1 windowSize, u32@2 minHashesPerWindow, bool@96 concatenated, f32@4 error,
//                              bool@97 noncanonical, bool@98 preserveCase, u32@5 hashSeed (stored XOR 42); ptr0 referenceListOld,
//                              ptr1 locusList, ptr2 alphabet, ptr3 referenceList
//   ReferenceList 0/1          ptr0 references: List(Reference)
//   Reference   3/7            u32@0 length, u64@1 length64, u64@2 numValidKmers; ptr0 sequence, ptr1 quality, ptr2 name,
//                              ptr3 comment, ptr4 hashes32, ptr5 hashes64, ptr6 counts32
// Encoding: the published Cap'n Proto wire format, unpacked, with the stream framing of capnp::serialize (segment
// table: u32 count - 1, u32 words per segment, padded to 8 bytes).  The WRITER emits one segment (any reader accepts
// that; the reference's builder would spread a large message over several, so bytes are not expected to match its
// output -- there are no golden files upstream to pin them against); the READER takes any number of segments, far
// and double-far pointers, and structs shorter or longer than the layouts above (absent fields read as their default).
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "sketch_json.hpp"

namespace fb2host {
namespace capnp_lite {

// ---- building: one segment ------------------------------------------------------------------------------------
class Builder {
public:
    Builder() { w_.push_back(0); }                        // word 0: the root pointer
    size_t alloc(size_t n) { const size_t at = w_.size(); w_.resize(at + n, 0); return at; }
    uint64_t &word(size_t i) { return w_[i]; }
    void set_struct_ptr(size_t p, size_t target, uint32_t dwords, uint32_t pwords) {
        const int64_t off = (int64_t)target - (int64_t)p - 1;
        w_[p] = ((uint64_t)(uint32_t)(off << 2) & 0xFFFFFFFFull) | ((uint64_t)dwords << 32) | ((uint64_t)pwords << 48);
    }
    void set_list_ptr(size_t p, size_t target, uint32_t elem_code, uint32_t count) {
        const int64_t off = (int64_t)target - (int64_t)p - 1;
        w_[p] = (((uint64_t)(uint32_t)(off << 2) & 0xFFFFFFFFull) | 1ull) | ((uint64_t)elem_code << 32) | ((uint64_t)count << 35);
    }
    size_t init_struct(size_t p, uint32_t dwords, uint32_t pwords) {
        const size_t t = alloc(dwords + pwords);
        set_struct_ptr(p, t, dwords, pwords);
        return t;
    }
    // List(struct): a tag word, then `n` elements of dwords + pwords words each; returns the first element
    size_t init_struct_list(size_t p, uint32_t n, uint32_t dwords, uint32_t pwords) {
        const uint64_t words = (uint64_t)n * (dwords + pwords);
        if (words >= (1ull << 29)) throw std::runtime_error("capnp: list too long");
        const size_t t = alloc(1 + (size_t)words);
        set_list_ptr(p, t, 7, (uint32_t)words);
        w_[t] = ((uint64_t)n << 2) | ((uint64_t)dwords << 32) | ((uint64_t)pwords << 48);
        return t + 1;
    }
    void set_bytes(size_t p, const void *data, size_t n, bool text) {
        const size_t count = n + (text ? 1 : 0);
        if (count >= (1ull << 29)) throw std::runtime_error("capnp: blob too long");
        const size_t t = alloc((count + 7) / 8);
        if (n) memcpy(reinterpret_cast<uint8_t *>(&w_[t]), data, n);
        set_list_ptr(p, t, 2, (uint32_t)count);
    }
    void set_text(size_t p, const std::string &s) { set_bytes(p, s.data(), s.size(), true); }
    void set_data(size_t p, const std::string &s) { set_bytes(p, s.data(), s.size(), false); }
    template <class T>
    size_t init_prim_list(size_t p, uint32_t n) {           // List(UInt32) / List(UInt64)
        static_assert(sizeof(T) == 4 || sizeof(T) == 8, "element size");
        const size_t t = alloc(((size_t)n * sizeof(T) + 7) / 8);
        set_list_ptr(p, t, sizeof(T) == 4 ? 4 : 5, n);
        return t;
    }
    template <class T> void put(size_t struct_at, size_t index, T v) {       // data field at offset `index` in units of T
        memcpy(reinterpret_cast<uint8_t *>(&w_[struct_at]) + index * sizeof(T), &v, sizeof(T));
    }
    void put_bit(size_t struct_at, size_t bit, bool v) {
        uint8_t *b = reinterpret_cast<uint8_t *>(&w_[struct_at]) + bit / 8;
        *b = (uint8_t)((*b & ~(1u << (bit % 8))) | ((v ? 1u : 0u) << (bit % 8)));
    }
    std::string finish() const {                            // capnp::serialize::write_message framing
        std::string out;
        const uint32_t hdr[2] = {0u, (uint32_t)w_.size()};
        out.append(reinterpret_cast<const char *>(hdr), 8);
        out.append(reinterpret_cast<const char *>(w_.data()), w_.size() * 8);
        return out;
    }
private:
    std::vector<uint64_t> w_;
};

// ---- reading: any segments, far pointers, bounds checked -------------------------------------------------------
struct Message {
    std::vector<std::vector<uint64_t>> seg;
    uint64_t at(uint32_t s, uint64_t i) const {
        if (s >= seg.size() || i >= seg[s].size()) throw std::runtime_error("capnp: pointer out of bounds");
        return seg[s][(size_t)i];
    }
};
inline Message parse_message(const char *data, size_t len) {
    Message m;
    if (len < 8) throw std::runtime_error("capnp: message too short");
    uint32_t nseg1;
    memcpy(&nseg1, data, 4);
    const uint64_t nseg = (uint64_t)nseg1 + 1;
    if (nseg > 512) throw std::runtime_error("capnp: too many segments");
    const size_t table = (size_t)((4 + 4 * nseg + 7) / 8 * 8);
    if (len < table) throw std::runtime_error("capnp: truncated segment table");
    size_t pos = table;
    for (uint64_t s = 0; s < nseg; ++s) {
        uint32_t words;
        memcpy(&words, data + 4 + 4 * s, 4);
        if ((uint64_t)words * 8 > len - pos) throw std::runtime_error("capnp: truncated segment");
        std::vector<uint64_t> w(words);
        if (words) memcpy(w.data(), data + pos, (size_t)words * 8);
        pos += (size_t)words * 8;
        m.seg.push_back(std::move(w));
    }
    return m;
}
struct Loc { uint32_t seg; uint64_t idx; };                 // where a pointer word lives
struct Resolved { uint64_t ptr; uint32_t seg; uint64_t target; bool null; };   // the object pointer and the word it refers to
inline Resolved resolve(const Message &m, Loc l) {
    uint64_t p = m.at(l.seg, l.idx);
    if (p == 0) return {0, 0, 0, true};
    if ((p & 3u) != 2u) {
        const int64_t off = (int32_t)(uint32_t)(p & 0xFFFFFFFFull) >> 2;
        return {p, l.seg, (uint64_t)((int64_t)l.idx + 1 + off), false};
    }
    // far pointer: landing pad in another segment
    const bool dbl = (p >> 2) & 1u;
    const uint64_t pad = (uint32_t)(p & 0xFFFFFFFFull) >> 3;
    const uint32_t seg = (uint32_t)(p >> 32);
    if (!dbl) {
        const uint64_t q = m.at(seg, pad);
        if ((q & 3u) == 2u) throw std::runtime_error("capnp: far pointer to a far pointer");
        if (q == 0) return {0, 0, 0, true};
        const int64_t off = (int32_t)(uint32_t)(q & 0xFFFFFFFFull) >> 2;
        return {q, seg, (uint64_t)((int64_t)pad + 1 + off), false};
    }
    const uint64_t far = m.at(seg, pad), tag = m.at(seg, pad + 1);
    if ((far & 3u) != 2u || ((far >> 2) & 1u)) throw std::runtime_error("capnp: bad double-far landing pad");
    return {tag, (uint32_t)(far >> 32), (uint64_t)((uint32_t)(far & 0xFFFFFFFFull) >> 3), false};
}
struct StructRef {
    const Message *m = nullptr;
    uint32_t seg = 0; uint64_t data = 0; uint32_t dwords = 0, pwords = 0;
    template <class T> T get(size_t index) const {          // 0 when the struct (an older writer's) is shorter
        if ((index + 1) * sizeof(T) > (size_t)dwords * 8) return T(0);
        const uint64_t w = m->at(seg, data + index * sizeof(T) / 8);
        T v;
        memcpy(&v, reinterpret_cast<const uint8_t *>(&w) + (index * sizeof(T)) % 8, sizeof(T));
        return v;
    }
    bool bit(size_t b) const { return (get<uint8_t>(b / 8) >> (b % 8)) & 1u; }
    bool has_ptr(uint32_t i) const { return m && i < pwords && m->at(seg, data + dwords + i) != 0; }
    Loc ptr(uint32_t i) const { return {seg, data + dwords + i}; }
};
inline StructRef read_struct(const Message &m, Loc l) {
    StructRef s;
    s.m = &m;
    const Resolved r = resolve(m, l);
    if (r.null) return s;                                   // all defaults
    if ((r.ptr & 3u) != 0u) throw std::runtime_error("capnp: expected a struct pointer");
    s.seg = r.seg; s.data = r.target; s.dwords = (uint32_t)(r.ptr >> 32) & 0xFFFFu; s.pwords = (uint32_t)(r.ptr >> 48);
    if (s.dwords + s.pwords) (void)m.at(s.seg, s.data + s.dwords + s.pwords - 1);
    return s;
}
struct ListRef {
    const Message *m = nullptr;
    uint32_t seg = 0; uint64_t start = 0; uint32_t code = 0, count = 0, dwords = 0, pwords = 0;
    StructRef element(uint32_t i) const {                   // List(struct)
        if (code != 7 || i >= count) throw std::runtime_error("capnp: bad struct list access");
        StructRef s;
        s.m = m; s.seg = seg; s.data = start + (uint64_t)i * (dwords + pwords); s.dwords = dwords; s.pwords = pwords;
        return s;
    }
    template <class T> T prim(uint32_t i) const {
        if (i >= count || (code != 4 && code != 5) || (code == 4) != (sizeof(T) == 4)) throw std::runtime_error("capnp: bad primitive list access");
        const uint64_t w = m->at(seg, start + (uint64_t)i * sizeof(T) / 8);
        T v;
        memcpy(&v, reinterpret_cast<const uint8_t *>(&w) + ((uint64_t)i * sizeof(T)) % 8, sizeof(T));
        return v;
    }
};
inline ListRef read_list(const Message &m, Loc l) {
    ListRef r;
    r.m = &m;
    const Resolved p = resolve(m, l);
    if (p.null) return r;                                   // empty
    if ((p.ptr & 3u) != 1u) throw std::runtime_error("capnp: expected a list pointer");
    r.seg = p.seg; r.start = p.target; r.code = (uint32_t)(p.ptr >> 32) & 7u; r.count = (uint32_t)(p.ptr >> 35);
    if (r.code == 7) {
        const uint64_t tag = m.at(r.seg, r.start), words = r.count;
        if ((tag & 3u) != 0u) throw std::runtime_error("capnp: bad composite list tag");
        r.count = (uint32_t)(tag & 0xFFFFFFFFull) >> 2; r.dwords = (uint32_t)(tag >> 32) & 0xFFFFu; r.pwords = (uint32_t)(tag >> 48);
        r.start += 1;
        if ((uint64_t)r.count * (r.dwords + r.pwords) > words) throw std::runtime_error("capnp: composite list overruns its words");
        if (words) (void)m.at(r.seg, r.start + words - 1);
    } else {
        static const uint32_t bits[7] = {0, 1, 8, 16, 32, 64, 64};
        const uint64_t words = ((uint64_t)r.count * bits[r.code] + 63) / 64;
        if (words) (void)m.at(r.seg, r.start + words - 1);
    }
    return r;
}
inline std::string read_bytes(const Message &m, Loc l, bool text) {
    const ListRef r = read_list(m, l);
    if (!r.count) return std::string();
    if (r.code != 2) throw std::runtime_error("capnp: expected a byte list");
    std::string out((size_t)r.count, '\0');
    for (uint32_t i = 0; i < r.count; i += 8) {
        const uint64_t w = m.at(r.seg, r.start + i / 8);
        memcpy(&out[i], &w, std::min<size_t>(8, r.count - i));
    }
    if (text) {                                             // NUL-terminated
        if (out.back() != '\0') throw std::runtime_error("capnp: text without terminator");
        out.pop_back();
    }
    return out;
}

}  // namespace capnp_lite

// ---- .bsk: finch.capnp -----------------------------------------------------------------------------------------
// write_finch_file (serialization/mod.rs:123-165)
inline std::string write_finch_file(const std::vector<Sketch> &sketches) {
    using namespace capnp_lite;
    Builder b;
    const size_t root = b.init_struct(0, 0, 1);
    const size_t first = b.init_struct_list(root, (uint32_t)sketches.size(), 2, 5);
    for (size_t i = 0; i < sketches.size(); ++i) {
        const Sketch &s = sketches[i];
        const size_t at = first + i * 7;
        b.set_text(at + 2, s.name);
        b.put<uint64_t>(at, 0, s.seq_length);
        b.put<uint64_t>(at, 1, s.num_valid_kmers);
        b.set_text(at + 3, s.comment);
        const size_t h0 = b.init_struct_list(at + 4, (uint32_t)s.hashes.size(), 2, 2);
        for (size_t j = 0; j < s.hashes.size(); ++j) {
            const size_t h = h0 + j * 4;
            b.put<uint64_t>(h, 0, s.hashes[j]);
            b.set_data(h + 2, j < s.kmers.size() ? s.kmers[j] : std::string());
            b.put<uint32_t>(h, 2, j < s.counts.size() ? s.counts[j] : 0u);
            b.put<uint32_t>(h, 3, j < s.extras.size() ? s.extras[j] : 0u);
            // label: None -> the pointer stays null (mod.rs:146-148)
        }
        const size_t fp = b.init_struct(at + 5, 4, 0);
        b.put_bit(fp, 0, s.filter_params.filter_on == 1);                                   // unwrap_or(false)
        b.put<uint32_t>(fp, 1, s.filter_params.has_lo ? s.filter_params.lo : 0u);            // unwrap_or(0)
        b.put<uint32_t>(fp, 2, s.filter_params.has_hi ? s.filter_params.hi : UINT32_MAX);    // unwrap_or(u32::MAX)
        b.put<double>(fp, 2, s.filter_params.err_filter);
        b.put<double>(fp, 3, s.filter_params.strand_filter);
        const size_t sp = b.init_struct(at + 6, 5, 0);                                       // set_sketch_params (mod.rs:67-100)
        const SketchParams &p = s.sketch_params;
        b.put<uint16_t>(sp, 0, (uint16_t)(p.kind == Kind::Mash ? 0 : (p.kind == Kind::Scaled ? 1 : 2)));
        b.put<uint8_t>(sp, 2, p.kmer_length);
        if (p.kind != Kind::AllCounts) { b.put<uint64_t>(sp, 1, p.kmers_to_sketch); b.put<uint64_t>(sp, 2, p.hash_seed); }
        if (p.kind == Kind::Mash) { b.put<uint64_t>(sp, 3, p.final_size); b.put_bit(sp, 24, p.no_strict); }
        if (p.kind == Kind::Scaled) b.put<double>(sp, 4, p.scale);
    }
    return b.finish();
}
// read_finch_file (serialization/mod.rs:167-224)
inline std::vector<Sketch> read_finch_file(const char *data, size_t len) {
    using namespace capnp_lite;
    const Message m = parse_message(data, len);
    if (m.seg.empty() || m.seg[0].empty()) throw std::runtime_error("capnp: empty message");
    const StructRef root = read_struct(m, Loc{0, 0});
    std::vector<Sketch> out;
    if (!root.has_ptr(0)) return out;
    const ListRef sk = read_list(m, root.ptr(0));
    for (uint32_t i = 0; i < sk.count; ++i) {
        const StructRef cs = sk.element(i);
        Sketch s;
        if (cs.has_ptr(0)) s.name = read_bytes(m, cs.ptr(0), true);
        if (cs.has_ptr(1)) s.comment = read_bytes(m, cs.ptr(1), true);
        s.seq_length = cs.get<uint64_t>(0);
        s.num_valid_kmers = cs.get<uint64_t>(1);
        if (cs.has_ptr(2)) {
            const ListRef hs = read_list(m, cs.ptr(2));
            for (uint32_t j = 0; j < hs.count; ++j) {
                const StructRef h = hs.element(j);
                s.hashes.push_back(h.get<uint64_t>(0));
                s.kmers.push_back(h.has_ptr(0) ? read_bytes(m, h.ptr(0), false) : std::string());
                s.counts.push_back(h.get<uint32_t>(2));
                s.extras.push_back(h.get<uint32_t>(3));
            }
        }
        const StructRef sp = cs.has_ptr(4) ? read_struct(m, cs.ptr(4)) : StructRef{&m};
        const uint16_t method = sp.get<uint16_t>(0);                                        // get_sketch_params (mod.rs:102-121)
        if (method > 2) throw std::runtime_error("capnp: unknown sketch method");
        SketchParams &p = s.sketch_params;
        p.kind = method == 0 ? Kind::Mash : (method == 1 ? Kind::Scaled : Kind::AllCounts);
        p.kmer_length = sp.get<uint8_t>(2);
        if (p.kind != Kind::AllCounts) { p.kmers_to_sketch = sp.get<uint64_t>(1); p.hash_seed = sp.get<uint64_t>(2); }
        if (p.kind == Kind::Mash) { p.final_size = sp.get<uint64_t>(3); p.no_strict = sp.bit(24); }
        if (p.kind == Kind::Scaled) p.scale = sp.get<double>(4);
        const StructRef fp = cs.has_ptr(3) ? read_struct(m, cs.ptr(3)) : StructRef{&m};
        FilterParams &f = s.filter_params;
        f.filter_on = fp.bit(0) ? 1 : 0;                                                    // Some(get_filtered())
        const uint32_t lo = fp.get<uint32_t>(1), hi = fp.get<uint32_t>(2);
        f.has_lo = lo != 0; f.lo = lo;                                                      // 0 => None
        f.has_hi = hi != UINT32_MAX; f.hi = f.has_hi ? hi : 0;                              // u32::MAX => None
        f.err_filter = fp.get<double>(2);
        f.strand_filter = fp.get<double>(3);
        out.push_back(std::move(s));
    }
    return out;
}

// SketchParams::from_sketches (sketch_schemes/mod.rs:158-178): the first sketch's parameters, all others compatible
inline SketchParams params_from_sketches(const std::vector<Sketch> &sketches) {
    if (sketches.empty()) throw std::runtime_error("no sketches to serialize");
    const SketchParams &p = sketches[0].sketch_params;
    for (size_t i = 1; i < sketches.size(); ++i) {
        std::string nm, v1, v2;
        if (!p.compatible(sketches[i].sketch_params, nm, v1, v2))
            throw std::runtime_error("First sketch has " + nm + " " + v1 + ", but sketch " + std::to_string(i + 1) + " has " + nm + " " + v2);
    }
    return p;
}

// ---- .msh: mash.capnp ------------------------------------------------------------------------------------------
// write_mash_file (serialization/mash.rs:12-58).  `params` = SketchParams::from_sketches(sketches) (the caller checks
// that the sketches agree, as the reference does before writing).
inline std::string write_mash_file(const std::vector<Sketch> &sketches, const SketchParams &params) {
    using namespace capnp_lite;
    Builder b;
    const size_t root = b.init_struct(0, 3, 4);
    b.put<uint32_t>(root, 0, params.kmer_length);                                           // kmerSize
    b.put<uint32_t>(root, 5, (uint32_t)params.seed() ^ 42u);                                // hashSeed: default 42 is XORed in
    b.put<float>(root, 4, 0.0f);                                                            // error
    b.put_bit(root, 97, false);                                                             // noncanonical
    b.put_bit(root, 98, false);                                                             // preserveCase
    b.set_text(root + 3 + 2, "ACGT");                                                       // alphabet
    size_t largest = sketches.empty() ? 1 : 0;                                              // .max().unwrap_or(1)
    for (auto &s : sketches) largest = std::max(largest, s.hashes.size());
    b.put<uint32_t>(root, 1, params.kmer_length);                                           // windowSize
    b.put<uint32_t>(root, 2, (uint32_t)largest);                                            // minHashesPerWindow
    b.put_bit(root, 96, true);                                                              // concatenated
    const size_t rl = b.init_struct(root + 3 + 3, 0, 1);                                    // referenceList
    const size_t first = b.init_struct_list(rl, (uint32_t)sketches.size(), 3, 7);
    for (size_t i = 0; i < sketches.size(); ++i) {
        const Sketch &s = sketches[i];
        const size_t at = first + i * 10;
        b.set_text(at + 3 + 2, s.name);
        b.set_text(at + 3 + 3, s.comment);
        b.put<uint64_t>(at, 1, s.seq_length);                                               // length64
        b.put<uint64_t>(at, 2, s.num_valid_kmers);
        const size_t h = b.init_prim_list<uint64_t>(at + 3 + 5, (uint32_t)s.hashes.size());
        for (size_t j = 0; j < s.hashes.size(); ++j) b.word(h + j) = s.hashes[j];
        const size_t c = b.init_prim_list<uint32_t>(at + 3 + 6, (uint32_t)s.hashes.size());
        for (size_t j = 0; j < s.hashes.size(); ++j) {
            const uint32_t v = j < s.counts.size() ? s.counts[j] : 0u;
            memcpy(reinterpret_cast<uint8_t *>(&b.word(c + j / 2)) + 4 * (j % 2), &v, 4);
        }
    }
    return b.finish();
}
// read_mash_file (serialization/mash.rs:60-132)
inline std::vector<Sketch> read_mash_file(const char *data, size_t len) {
    using namespace capnp_lite;
    const Message m = parse_message(data, len);
    if (m.seg.empty() || m.seg[0].empty()) throw std::runtime_error("capnp: empty message");
    const StructRef root = read_struct(m, Loc{0, 0});
    SketchParams params;
    params.kind = Kind::Mash;
    params.kmers_to_sketch = 0; params.final_size = 0; params.no_strict = true;
    params.hash_seed = (uint64_t)(root.get<uint32_t>(5) ^ 42u);
    params.kmer_length = (uint8_t)root.get<uint32_t>(0);
    ListRef refs;
    refs.m = &m;
    bool have = false;
    if (root.has_ptr(3)) {                                   // referenceList, if it has references
        const StructRef rl = read_struct(m, root.ptr(3));
        if (rl.has_ptr(0)) { refs = read_list(m, rl.ptr(0)); have = true; }
    }
    if (!have && root.has_ptr(0)) {                          // else referenceListOld
        const StructRef rl = read_struct(m, root.ptr(0));
        if (rl.has_ptr(0)) refs = read_list(m, rl.ptr(0));
    }
    std::vector<Sketch> out;
    for (uint32_t i = 0; i < refs.count; ++i) {
        const StructRef r = refs.element(i);
        Sketch s;
        ListRef hs, cs;
        hs.m = cs.m = &m;
        if (r.has_ptr(5)) hs = read_list(m, r.ptr(5));
        if (r.has_ptr(6)) cs = read_list(m, r.ptr(6));
        if (hs.count && hs.code != 5) throw std::runtime_error("capnp: hashes64 is not a list of 64-bit values");
        if (cs.count && cs.code != 4) throw std::runtime_error("capnp: counts32 is not a list of 32-bit values");
        const uint32_t n = cs.count == 0 ? hs.count : std::min(hs.count, cs.count);         // zip
        for (uint32_t j = 0; j < n; ++j) {
            s.hashes.push_back(hs.prim<uint64_t>(j));
            s.kmers.emplace_back();
            const uint32_t c = cs.count == 0 ? 1u : cs.prim<uint32_t>(j);                   // old lists carry no counts
            s.counts.push_back(c);
            s.extras.push_back(cs.count == 0 ? 0u : c / 2);
        }
        if (r.has_ptr(2)) s.name = read_bytes(m, r.ptr(2), true);
        if (r.has_ptr(3)) s.comment = read_bytes(m, r.ptr(3), true);
        s.seq_length = r.get<uint64_t>(1);
        s.num_valid_kmers = r.get<uint64_t>(2);
        s.sketch_params = params;
        s.filter_params = FilterParams();                    // FilterParams::default()
        out.push_back(std::move(s));
    }
    return out;
}

}  // namespace fb2host

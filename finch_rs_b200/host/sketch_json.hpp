// sketch_json.hpp -- host-side model of finch's `Sketch` / `MultiSketch` and the `.sk` JSON wire format
// (SURVEY 8f row N2).  Header-only, no CUDA; used by the CLI (finch_rs_b200/cli/finch_cli.cpp).
//
// Mirrors, field for field and in the same order:
//   MultiSketch                       lib/src/serialization/json.rs:141-158
//   JsonSketch::serialize             json.rs:64-89   (hashes as decimal STRINGS, kmers, counts)
//   JsonSketch::deserialize           json.rs:91-139  (kmers / counts optional; extra_count = count / 2)
//   MultiSketch::{get_params,from_sketches,to_sketches}   json.rs:160-239
//   FilterParams::{to_serialized,from_serialized}         lib/src/filtering.rs:89-135
//   SketchDistance                    lib/src/serialization/mod.rs:31-43
// Numbers are printed the way serde_json (Ryu) and Rust's `Display` print them, so the bytes match the
// reference wherever its own field order is deterministic (`filters` is a HashMap there: compare parsed).
#pragma once
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace fb2host {

// ---- number formatting -------------------------------------------------------------------------------
// shortest round-trip digits of |v| and the decimal exponent such that v = 0.d1d2... * 10^kk
inline void shortest_digits(double v, std::string &digits, int &kk) {
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, std::fabs(v), std::chars_format::scientific);
    std::string s(buf, r.ptr);                       // d.ddddde[+-]XX
    const size_t e = s.find('e');
    const int exp10 = std::stoi(s.substr(e + 1));
    digits.clear();
    for (size_t i = 0; i < e; ++i) if (s[i] != '.') digits.push_back(s[i]);
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    kk = exp10 + 1;
}
// serde_json f64 (ryu::Buffer::format_finite, pretty.rs): 1.0, 0.001, 1e-7, 1.5e16, 123456.789
inline std::string json_f64(double v) {
    if (!std::isfinite(v)) return "null";            // serde_json writes null for NaN / inf
    if (v == 0.0) return std::signbit(v) ? "-0.0" : "0.0";
    std::string d; int kk;
    shortest_digits(v, d, kk);
    const int len = (int)d.size();
    std::string out = v < 0 ? "-" : "";
    if (len <= kk && kk <= 16) { out += d; out.append((size_t)(kk - len), '0'); out += ".0"; }
    else if (0 < kk && kk <= 16) { out += d.substr(0, (size_t)kk); out += '.'; out += d.substr((size_t)kk); }
    else if (-5 < kk && kk <= 0) { out += "0."; out.append((size_t)(-kk), '0'); out += d; }
    else {
        out += d[0];
        if (len > 1) { out += '.'; out += d.substr(1); }
        out += 'e'; out += std::to_string(kk - 1);
    }
    return out;
}
// Rust `Display` for f64 / f32 (`to_string()`): shortest round-trip digits, never an exponent, no ".0"
inline std::string rust_display_digits(const std::string &d, int kk, bool neg) {
    const int len = (int)d.size();
    std::string out = neg ? "-" : "";
    if (kk <= 0) { out += "0."; out.append((size_t)(-kk), '0'); out += d; }
    else if (kk >= len) { out += d; out.append((size_t)(kk - len), '0'); }
    else { out += d.substr(0, (size_t)kk); out += '.'; out += d.substr((size_t)kk); }
    return out;
}
inline std::string rust_display_f64(double v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
    if (v == 0.0) return std::signbit(v) ? "-0" : "0";
    std::string d; int kk;
    shortest_digits(v, d, kk);
    return rust_display_digits(d, kk, v < 0);
}
inline std::string rust_display_f32(float v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
    if (v == 0.0f) return std::signbit(v) ? "-0" : "0";
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, std::fabs(v), std::chars_format::scientific);
    std::string s(buf, r.ptr);
    const size_t e = s.find('e');
    const int exp10 = std::stoi(s.substr(e + 1));
    std::string d;
    for (size_t i = 0; i < e; ++i) if (s[i] != '.') d.push_back(s[i]);
    while (d.size() > 1 && d.back() == '0') d.pop_back();
    return rust_display_digits(d, exp10 + 1, v < 0);
}
inline std::string json_string(const std::string &s) {      // serde_json string escaping
    std::string o = "\"";
    for (unsigned char c : s) {
        switch (c) {
        case '"': o += "\\\""; break;
        case '\\': o += "\\\\"; break;
        case '\n': o += "\\n"; break;
        case '\r': o += "\\r"; break;
        case '\t': o += "\\t"; break;
        case '\b': o += "\\b"; break;
        case '\f': o += "\\f"; break;
        default:
            if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); o += b; }
            else o += (char)c;
        }
    }
    return o + "\"";
}

// ---- a small JSON value + recursive-descent parser (enough for .sk files) --------------------------
struct JValue {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    std::string text;                                   // Num: the literal as written; Str: decoded
    std::vector<JValue> arr;
    std::vector<std::pair<std::string, JValue>> obj;    // insertion order kept
    const JValue *get(const std::string &k) const {
        for (auto &kv : obj) if (kv.first == k) return &kv.second;
        return nullptr;
    }
};
class JParser {
public:
    JParser(const char *p, size_t n) : p_(p), e_(p + n) {}
    JValue parse() { JValue v = value(); ws(); if (p_ != e_) fail("trailing characters"); return v; }
private:
    const char *p_, *e_;
    [[noreturn]] void fail(const char *m) { throw std::runtime_error(std::string("JSON: ") + m); }
    void ws() { while (p_ < e_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r')) ++p_; }
    JValue value() {
        ws();
        if (p_ >= e_) fail("unexpected end");
        JValue v;
        const char c = *p_;
        if (c == '{') {
            v.kind = JValue::Obj; ++p_; ws();
            if (p_ < e_ && *p_ == '}') { ++p_; return v; }
            while (true) {
                ws();
                if (p_ >= e_ || *p_ != '"') fail("expected object key");
                std::string k = str();
                ws();
                if (p_ >= e_ || *p_ != ':') fail("expected ':'");
                ++p_;
                v.obj.emplace_back(std::move(k), value());
                ws();
                if (p_ < e_ && *p_ == ',') { ++p_; continue; }
                if (p_ < e_ && *p_ == '}') { ++p_; break; }
                fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            v.kind = JValue::Arr; ++p_; ws();
            if (p_ < e_ && *p_ == ']') { ++p_; return v; }
            while (true) {
                v.arr.push_back(value());
                ws();
                if (p_ < e_ && *p_ == ',') { ++p_; continue; }
                if (p_ < e_ && *p_ == ']') { ++p_; break; }
                fail("expected ',' or ']'");
            }
        } else if (c == '"') { v.kind = JValue::Str; v.text = str(); }
        else if (c == 't' && e_ - p_ >= 4 && !memcmp(p_, "true", 4)) { v.kind = JValue::Bool; v.b = true; p_ += 4; }
        else if (c == 'f' && e_ - p_ >= 5 && !memcmp(p_, "false", 5)) { v.kind = JValue::Bool; v.b = false; p_ += 5; }
        else if (c == 'n' && e_ - p_ >= 4 && !memcmp(p_, "null", 4)) { v.kind = JValue::Null; p_ += 4; }
        else if (c == '-' || (c >= '0' && c <= '9')) {
            v.kind = JValue::Num;
            const char *s = p_;
            while (p_ < e_ && (*p_ == '-' || *p_ == '+' || *p_ == '.' || *p_ == 'e' || *p_ == 'E' || (*p_ >= '0' && *p_ <= '9'))) ++p_;
            v.text.assign(s, p_);
        } else fail("unexpected character");
        return v;
    }
    static void utf8(std::string &o, uint32_t cp) {
        if (cp < 0x80) o += (char)cp;
        else if (cp < 0x800) { o += (char)(0xC0 | (cp >> 6)); o += (char)(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) { o += (char)(0xE0 | (cp >> 12)); o += (char)(0x80 | ((cp >> 6) & 0x3F)); o += (char)(0x80 | (cp & 0x3F)); }
        else { o += (char)(0xF0 | (cp >> 18)); o += (char)(0x80 | ((cp >> 12) & 0x3F)); o += (char)(0x80 | ((cp >> 6) & 0x3F)); o += (char)(0x80 | (cp & 0x3F)); }
    }
    uint32_t hex4() {
        if (e_ - p_ < 4) fail("bad \\u escape");
        uint32_t v = 0;
        for (int i = 0; i < 4; ++i) {
            const char c = *p_++;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= (uint32_t)(c - '0');
            else if (c >= 'a' && c <= 'f') v |= (uint32_t)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= (uint32_t)(c - 'A' + 10);
            else fail("bad \\u escape");
        }
        return v;
    }
    std::string str() {
        std::string o;
        ++p_;  // opening quote
        while (true) {
            if (p_ >= e_) fail("unterminated string");
            const char c = *p_++;
            if (c == '"') break;
            if (c != '\\') { o += c; continue; }
            if (p_ >= e_) fail("unterminated escape");
            const char x = *p_++;
            switch (x) {
            case '"': o += '"'; break; case '\\': o += '\\'; break; case '/': o += '/'; break;
            case 'b': o += '\b'; break; case 'f': o += '\f'; break; case 'n': o += '\n'; break;
            case 'r': o += '\r'; break; case 't': o += '\t'; break;
            case 'u': {
                uint32_t cp = hex4();
                if (cp >= 0xD800 && cp < 0xDC00 && e_ - p_ >= 6 && p_[0] == '\\' && p_[1] == 'u') {
                    p_ += 2;
                    const uint32_t lo = hex4();
                    cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                }
                utf8(o, cp);
                break;
            }
            default: fail("bad escape");
            }
        }
        return o;
    }
};
inline uint64_t parse_u64(const std::string &s, const char *what) {
    if (s.empty()) throw std::runtime_error(std::string("bad ") + what);
    uint64_t v = 0;
    auto r = std::from_chars(s.data(), s.data() + s.size(), v);
    if (r.ec != std::errc() || r.ptr != s.data() + s.size()) throw std::runtime_error(std::string("bad ") + what + ": " + s);
    return v;
}
inline double parse_f64(const std::string &s, const char *what) {
    double v = 0;
    auto r = std::from_chars(s.data(), s.data() + s.size(), v);
    if (r.ec != std::errc() || r.ptr != s.data() + s.size()) throw std::runtime_error(std::string("bad ") + what + ": " + s);
    return v;
}

// ---- SketchParams / FilterParams / Sketch -----------------------------------------------------------
enum class Kind { Mash, Scaled, AllCounts };
struct SketchParams {                                   // sketch_schemes/mod.rs:53-84
    Kind kind = Kind::Mash;
    uint64_t kmers_to_sketch = 0, final_size = 0;
    bool no_strict = false;
    uint8_t kmer_length = 21;
    uint64_t hash_seed = 0;
    double scale = 0.0;
    const char *hash_type() const { return kind == Kind::AllCounts ? "None" : "MurmurHash3_x64_128"; }   // mod.rs:138-146
    uint16_t hash_bits() const { return kind == Kind::AllCounts ? 0 : 64; }
    uint64_t seed() const { return kind == Kind::AllCounts ? 0 : hash_seed; }
    bool has_scale() const { return kind == Kind::Scaled; }
    uint64_t expected_size() const {                    // mod.rs:148-156
        if (kind == Kind::Mash) return final_size;
        if (kind == Kind::Scaled) return kmers_to_sketch;
        uint64_t v = 1; for (int i = 0; i < kmer_length; ++i) v *= 4; return v;
    }
    bool operator==(const SketchParams &o) const {      // #[derive(PartialEq)] on the enum
        if (kind != o.kind || kmer_length != o.kmer_length) return false;
        if (kind == Kind::Mash) return kmers_to_sketch == o.kmers_to_sketch && final_size == o.final_size && no_strict == o.no_strict && hash_seed == o.hash_seed;
        if (kind == Kind::Scaled) return kmers_to_sketch == o.kmers_to_sketch && scale == o.scale && hash_seed == o.hash_seed;
        return true;
    }
    // check_compatibility (mod.rs:186-214): "" when compatible, else "<param>\t<mine>\t<theirs>"
    bool compatible(const SketchParams &o, std::string &name, std::string &v1, std::string &v2) const {
        if (kmer_length != o.kmer_length) { name = "k"; v1 = std::to_string(kmer_length); v2 = std::to_string(o.kmer_length); return false; }
        if (strcmp(hash_type(), o.hash_type())) { name = "hash type"; v1 = hash_type(); v2 = o.hash_type(); return false; }
        if (hash_bits() != o.hash_bits()) { name = "hash bits"; v1 = std::to_string(hash_bits()); v2 = std::to_string(o.hash_bits()); return false; }
        if (seed() != o.seed()) { name = "hash seed"; v1 = std::to_string(seed()); v2 = std::to_string(o.seed()); return false; }
        return true;
    }
};
struct FilterParams {                                   // filtering.rs:10-16
    int filter_on = -1;                                 // -1 None, 0 Some(false), 1 Some(true)
    bool has_lo = false, has_hi = false;
    uint32_t lo = 0, hi = 0;
    double err_filter = 0.0, strand_filter = 0.0;
    bool operator==(const FilterParams &o) const {
        return filter_on == o.filter_on && has_lo == o.has_lo && has_hi == o.has_hi && (!has_lo || lo == o.lo) &&
               (!has_hi || hi == o.hi) && err_filter == o.err_filter && strand_filter == o.strand_filter;
    }
    // to_serialized (filtering.rs:89-108); the reference's HashMap order is arbitrary, ours is fixed
    std::vector<std::pair<std::string, std::string>> to_serialized() const {
        std::vector<std::pair<std::string, std::string>> m;
        if (filter_on != 1) return m;
        if (strand_filter > 0.0) m.emplace_back("strandFilter", rust_display_f64(strand_filter));
        if (err_filter > 0.0) m.emplace_back("errFilter", rust_display_f64(err_filter));
        if (has_lo) m.emplace_back("minCopies", std::to_string(lo));
        if (has_hi) m.emplace_back("maxCopies", std::to_string(hi));
        return m;
    }
    static FilterParams from_serialized(const JValue *filters) {     // filtering.rs:110-135
        FilterParams f;
        size_t n = 0;
        f.filter_on = 0;
        if (filters && filters->kind == JValue::Obj) {
            n = filters->obj.size();
            for (auto &kv : filters->obj) {
                if (kv.second.kind != JValue::Str) throw std::runtime_error("filters values must be strings");
                if (kv.first == "minCopies") { f.has_lo = true; f.lo = (uint32_t)parse_u64(kv.second.text, "minCopies"); }
                else if (kv.first == "maxCopies") { f.has_hi = true; f.hi = (uint32_t)parse_u64(kv.second.text, "maxCopies"); }
                else if (kv.first == "errFilter") f.err_filter = parse_f64(kv.second.text, "errFilter");
                else if (kv.first == "strandFilter") f.strand_filter = parse_f64(kv.second.text, "strandFilter");
            }
        }
        f.filter_on = n ? 1 : 0;
        return f;
    }
};
struct Sketch {                                         // serialization/mod.rs:45-55
    std::string name, comment;
    uint64_t seq_length = 0, num_valid_kmers = 0;
    std::vector<uint64_t> hashes;                       // KmerCount SoA, ascending by hash
    std::vector<std::string> kmers;
    std::vector<uint32_t> counts, extras;
    FilterParams filter_params;
    SketchParams sketch_params;
    bool operator==(const Sketch &o) const {            // #[derive(PartialEq)]: every field (main.rs:324 relies on it)
        return name == o.name && comment == o.comment && seq_length == o.seq_length && num_valid_kmers == o.num_valid_kmers &&
               hashes == o.hashes && kmers == o.kmers && counts == o.counts && extras == o.extras &&
               filter_params == o.filter_params && sketch_params == o.sketch_params;
    }
};

// MultiSketch::from_sketches + serde_json::to_writer (json.rs:64-89,141-158,200-217)
inline std::string write_multisketch_json(const std::vector<Sketch> &sketches) {
    if (sketches.empty()) throw std::runtime_error("no sketches to serialize");
    const SketchParams &p = sketches[0].sketch_params;   // SketchParams::from_sketches (mod.rs:158-178)
    for (size_t i = 1; i < sketches.size(); ++i) {
        std::string nm, v1, v2;
        if (!p.compatible(sketches[i].sketch_params, nm, v1, v2))
            throw std::runtime_error("First sketch has " + nm + " " + v1 + ", but sketch " + std::to_string(i + 1) + " has " + nm + " " + v2);
    }
    std::string o;
    o.reserve(256 + sketches.size() * 64);
    o += "{\"kmer\":" + std::to_string(p.kmer_length);
    o += ",\"alphabet\":\"ACGT\",\"preserveCase\":false,\"canonical\":true";
    o += ",\"sketchSize\":" + std::to_string((uint32_t)p.expected_size());   // `as u32`
    o += std::string(",\"hashType\":") + json_string(p.hash_type());
    o += ",\"hashBits\":" + std::to_string(p.hash_bits());
    o += ",\"hashSeed\":" + std::to_string(p.seed());
    o += ",\"scale\":" + (p.has_scale() ? json_f64(p.scale) : std::string("null"));
    o += ",\"sketches\":[";
    for (size_t s = 0; s < sketches.size(); ++s) {
        const Sketch &k = sketches[s];
        if (s) o += ',';
        o += "{\"name\":" + json_string(k.name);
        o += ",\"seqLength\":" + std::to_string(k.seq_length);
        o += ",\"numValidKmers\":" + std::to_string(k.num_valid_kmers);
        o += ",\"comment\":" + json_string(k.comment);
        o += ",\"filters\":{";
        const auto f = k.filter_params.to_serialized();
        for (size_t i = 0; i < f.size(); ++i) { if (i) o += ','; o += json_string(f[i].first) + ":" + json_string(f[i].second); }
        o += "},\"hashes\":[";
        for (size_t i = 0; i < k.hashes.size(); ++i) { if (i) o += ','; o += '"'; o += std::to_string(k.hashes[i]); o += '"'; }
        o += "],\"kmers\":[";
        for (size_t i = 0; i < k.hashes.size(); ++i) { if (i) o += ','; o += json_string(i < k.kmers.size() ? k.kmers[i] : std::string()); }
        o += "],\"counts\":[";
        for (size_t i = 0; i < k.hashes.size(); ++i) { if (i) o += ','; o += std::to_string(i < k.counts.size() ? k.counts[i] : 1u); }
        o += "]}";
    }
    o += "]}";
    return o;
}

// serde_json::from_slice::<MultiSketch> + to_sketches (json.rs:91-139,160-239)
inline std::vector<Sketch> read_multisketch_json(const char *data, size_t len) {
    JValue root = JParser(data, len).parse();
    if (root.kind != JValue::Obj) throw std::runtime_error("not a MultiSketch object");
    auto need = [&](const char *k, JValue::Kind kind) -> const JValue & {
        const JValue *v = root.get(k);
        if (!v || v->kind != kind) throw std::runtime_error(std::string("missing or mistyped field `") + k + "`");
        return *v;
    };
    SketchParams p;
    p.kmer_length = (uint8_t)parse_u64(need("kmer", JValue::Num).text, "kmer");
    need("alphabet", JValue::Str); need("preserveCase", JValue::Bool); need("canonical", JValue::Bool);
    const uint64_t sketch_size = parse_u64(need("sketchSize", JValue::Num).text, "sketchSize");
    const std::string hash_type = need("hashType", JValue::Str).text;
    const uint64_t hash_bits = parse_u64(need("hashBits", JValue::Num).text, "hashBits");
    const uint64_t hash_seed = parse_u64(need("hashSeed", JValue::Num).text, "hashSeed");
    const JValue *scale = root.get("scale");              // Option<f64>: absent or null = None
    const bool has_scale = scale && scale->kind == JValue::Num;
    if (hash_type == "MurmurHash3_x64_128") {              // get_params (json.rs:160-198)
        if (hash_bits != 64) throw std::runtime_error("Multisketch has incompatible hash size (" + std::to_string(hash_bits) + " != 64)");
        p.hash_seed = hash_seed; p.kmers_to_sketch = sketch_size;
        if (has_scale) { p.kind = Kind::Scaled; p.scale = parse_f64(scale->text, "scale"); }
        else { p.kind = Kind::Mash; p.final_size = sketch_size; p.no_strict = true; }
    } else if (hash_type == "None") p.kind = Kind::AllCounts;
    else throw std::runtime_error(hash_type + " sketch type is not supported");
    std::vector<Sketch> out;
    for (const JValue &js : need("sketches", JValue::Arr).arr) {
        if (js.kind != JValue::Obj) throw std::runtime_error("sketch entry is not an object");
        Sketch s;
        const JValue *nm = js.get("name");
        if (!nm || nm->kind != JValue::Str) throw std::runtime_error("sketch without a name");
        s.name = nm->text;
        if (const JValue *v = js.get("seqLength")) if (v->kind == JValue::Num) s.seq_length = parse_u64(v->text, "seqLength");
        if (const JValue *v = js.get("numValidKmers")) if (v->kind == JValue::Num) s.num_valid_kmers = parse_u64(v->text, "numValidKmers");
        if (const JValue *v = js.get("comment")) if (v->kind == JValue::Str) s.comment = v->text;
        const JValue *f = js.get("filters");
        s.filter_params = FilterParams::from_serialized(f && f->kind == JValue::Obj ? f : nullptr);
        const JValue *h = js.get("hashes");
        if (!h || h->kind != JValue::Arr) throw std::runtime_error("sketch without hashes");
        const JValue *km = js.get("kmers"), *ct = js.get("counts");
        const bool has_km = km && km->kind == JValue::Arr, has_ct = ct && ct->kind == JValue::Arr;
        const size_t n = h->arr.size();
        if ((has_km && km->arr.size() < n) || (has_ct && ct->arr.size() < n)) throw std::runtime_error("kmers/counts shorter than hashes");
        s.hashes.resize(n); s.kmers.resize(n); s.counts.resize(n); s.extras.resize(n);
        for (size_t i = 0; i < n; ++i) {
            if (h->arr[i].kind != JValue::Str) throw std::runtime_error("usize as a json string expected in `hashes`");
            s.hashes[i] = parse_u64(h->arr[i].text, "hash");
            if (has_km) s.kmers[i] = km->arr[i].text;
            s.counts[i] = has_ct ? (uint32_t)parse_u64(ct->arr[i].text, "count") : 1u;
            s.extras[i] = s.counts[i] / 2;               // json.rs:126
        }
        s.sketch_params = p;
        out.push_back(std::move(s));
    }
    return out;
}

// statistics.rs:30-47
inline std::vector<uint64_t> hist(const std::vector<uint32_t> &counts) {
    uint32_t mx = 0;
    for (uint32_t c : counts) mx = std::max(mx, c);
    std::vector<uint64_t> h(mx, 0);
    for (uint32_t c : counts) if (c) ++h[c - 1];
    return h;
}

struct SketchDistance {                                 // serialization/mod.rs:31-43
    double containment, jaccard, mash_distance;
    uint64_t common_hashes, total_hashes;
    std::string query, reference;
};
inline std::string write_distances_json(const std::vector<SketchDistance> &d) {
    std::string o = "[";
    for (size_t i = 0; i < d.size(); ++i) {
        if (i) o += ',';
        o += "{\"containment\":" + json_f64(d[i].containment) + ",\"jaccard\":" + json_f64(d[i].jaccard) +
             ",\"mashDistance\":" + json_f64(d[i].mash_distance) + ",\"commonHashes\":" + std::to_string(d[i].common_hashes) +
             ",\"totalHashes\":" + std::to_string(d[i].total_hashes) + ",\"query\":" + json_string(d[i].query) +
             ",\"reference\":" + json_string(d[i].reference) + "}";
    }
    return o + "]";
}

}  // namespace fb2host

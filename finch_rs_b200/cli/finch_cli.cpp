// finch_cli.cpp -- `finch` command line over the B200 C ABI.
//
// Mirrors the reference CLI's surface for the sketching / distance path:
//   flags and defaults        cli/src/cli.rs:7-232      (sketch / dist / hist / info sub-commands)
//   flag -> parameter rules   cli/src/cli.rs:241-340    (parse_filter_options, parse_sketch_options)
//   command bodies            cli/src/main.rs:48-200    (output naming, -o/-O, in-place sketching)
//   parse_mash_files          cli/src/main.rs:237-313   (sketch files + sequence files, update_sketch_params)
//   calc_sketch_distances     cli/src/main.rs:315-334   (ref-major order, skip equal sketches, max-dist)
//   update_sketch_params      cli/src/main.rs:336-441
// The sequence work (sketch_files) and the sorted-hash intersections (raw_distance) run on the GPU
// through libfinch_b200.so; there is no CPU fallback.  `.bsk` / `.msh` Cap'n Proto files: host/sketch_capnp.hpp.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/finch_b200.h"
#include "../host/sketch_json.hpp"
#include "../host/sketch_capnp.hpp"

using namespace fb2host;

namespace {

struct Bail : std::runtime_error { using std::runtime_error::runtime_error; };
[[noreturn]] void bail(const std::string &m) { throw Bail(m); }

const char *FINCH_EXT = ".sk", *FINCH_BIN_EXT = ".bsk", *MASH_EXT = ".msh";
bool ends_with(const std::string &s, const std::string &suf) {
    return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0;
}
bool is_sketch_file(const std::string &f) {
    return ends_with(f, ".json") || ends_with(f, FINCH_EXT) || ends_with(f, FINCH_BIN_EXT) || ends_with(f, MASH_EXT);
}

// ---- a small clap-like matcher -------------------------------------------------------------------
struct OptSpec { const char *name; char short_name; const char *long_name; bool takes_value; bool multiple; const char *dflt; };
const OptSpec OPTS[] = {
    {"binary_format", 'b', "finch-binary-format", false, false, nullptr},
    {"mash_binary_format", 'B', "mash-binary-format", false, false, nullptr},
    {"output_file", 'o', "output", true, false, nullptr},
    {"std_out", 'O', "std-out", false, false, nullptr},
    {"no_filter", 0, "no-filter", false, false, nullptr},
    {"filter", 'f', "filter", false, false, nullptr},
    {"min_abun_filter", 0, "min-abun-filter", true, false, nullptr},
    {"max_abun_filter", 0, "max-abun-filter", true, false, nullptr},
    {"strand_filter", 0, "strand-filter", true, false, "0.1"},
    {"err_filter", 0, "err-filter", true, false, "1"},
    {"sketch_type", 's', "sketch-type", true, false, "mash"},
    {"kmer_length", 'k', "kmer-length", true, false, "21"},
    {"n_hashes", 'n', "n-hashes", true, false, "1000"},
    {"scale", 0, "scale", true, false, "0.001"},
    {"seed", 0, "seed", true, false, "0"},
    {"oversketch", 0, "oversketch", true, false, "200"},
    {"no_strict", 'N', "no-strict", false, false, nullptr},
    {"pairwise", 'p', "pairwise", false, false, nullptr},
    {"queries", 'q', "queries", true, true, nullptr},
    {"max_distance", 'd', "max-dist", true, false, "1.0"},
    {"old_dist_mode", 0, "old-dist", false, false, nullptr},
};
struct Matches {
    std::string sub;
    std::map<std::string, std::vector<std::string>> vals;   // occurrences on the command line
    std::vector<std::string> inputs;
    bool is_present(const char *k) const { return vals.count(k) != 0; }
    size_t occurrences_of(const char *k) const { auto it = vals.find(k); return it == vals.end() ? 0 : std::max<size_t>(1, it->second.size()); }
    const char *value_of(const char *k) const {
        auto it = vals.find(k);
        if (it != vals.end() && !it->second.empty()) return it->second[0].c_str();
        for (const OptSpec &o : OPTS) if (!strcmp(o.name, k)) {
            if (!strcmp(k, "kmer_length") && !strcmp(value_of("sketch_type"), "none")) return "4";   // cli.rs:174
            return o.dflt;
        }
        return nullptr;
    }
};
bool allowed(const std::string &sub, const char *name) {
    const bool out = !strcmp(name, "output_file") || !strcmp(name, "std_out");
    const bool bin = !strcmp(name, "binary_format") || !strcmp(name, "mash_binary_format");
    const bool dist = !strcmp(name, "pairwise") || !strcmp(name, "queries") || !strcmp(name, "max_distance") || !strcmp(name, "old_dist_mode");
    if (bin) return sub == "sketch";
    if (dist) return sub == "dist";
    if (out) return sub != "info";
    return true;   // filter + sketch options: every sub-command
}
void usage(FILE *f) {
    fprintf(f,
            "finch (B200 build)\nTool for working with genomic MinHash sketches\n\nUSAGE:\n    finch <SUBCOMMAND>\n\n"
            "SUBCOMMANDS:\n    info      Display basic statistics\n    sketch    Create sketches from FASTA/Q file(s)\n"
            "    dist      Compute distances between sketches\n    hist      Display histograms of kmer abundances\n\n"
            "OPTIONS (see cli/src/cli.rs of finch-rs for the full help text):\n"
            "    -o, --output <file>   -O, --std-out   -f, --filter   --no-filter   --min-abun-filter <n>\n"
            "    --max-abun-filter <n>   --strand-filter <0.1>   --err-filter <1>   -s, --sketch-type <mash|scaled|none>\n"
            "    -k, --kmer-length <21>   -n, --n-hashes <1000>   --scale <0.001>   --seed <0>   --oversketch <200>\n"
            "    -N, --no-strict   [dist] -p, --pairwise   -q, --queries <name>...   -d, --max-dist <1.0>   --old-dist\n"
            "    [sketch] -b, --finch-binary-format (.bsk)   -B, --mash-binary-format (.msh)\n");
}
Matches parse_args(int argc, char **argv) {
    Matches m;
    if (argc < 2) { usage(stderr); exit(1); }
    m.sub = argv[1];
    if (m.sub == "-h" || m.sub == "--help" || m.sub == "help") { usage(stdout); exit(0); }
    if (m.sub == "-V" || m.sub == "--version") { printf("finch %s\n", fb2_version()); exit(0); }
    if (m.sub == "fmt-f64") {   // hidden: number formatting probe used by tests/test_cli_cpu.py
        for (int i = 2; i < argc; ++i) {
            const double v = strtod(argv[i], nullptr);
            printf("%s %s %s\n", json_f64(v).c_str(), rust_display_f64(v).c_str(), rust_display_f32((float)v).c_str());
        }
        exit(0);
    }
    if (m.sub != "sketch" && m.sub != "dist" && m.sub != "hist" && m.sub != "info")
        bail("Found argument '" + m.sub + "' which wasn't expected, or isn't valid in this context");
    const OptSpec *open_multi = nullptr;
    bool only_positional = false;
    for (int i = 2; i < argc; ++i) {
        std::string a = argv[i];
        if (only_positional || a == "-" || a.empty() || a[0] != '-') {
            if (open_multi && !only_positional) { m.vals[open_multi->name].push_back(a); continue; }   // -q a b c
            m.inputs.push_back(a);
            continue;
        }
        open_multi = nullptr;
        if (a == "--") { only_positional = true; continue; }
        if (a == "-h" || a == "--help") { usage(stdout); exit(0); }
        const OptSpec *spec = nullptr;
        std::string inline_val; bool has_inline = false;
        if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
            std::string name = a.substr(2);
            const size_t eq = name.find('=');
            if (eq != std::string::npos) { inline_val = name.substr(eq + 1); name = name.substr(0, eq); has_inline = true; }
            for (const OptSpec &o : OPTS) if (name == o.long_name) spec = &o;
        } else {
            for (const OptSpec &o : OPTS) if (o.short_name && a[1] == o.short_name) spec = &o;
            if (spec && a.size() > 2) {
                if (spec->takes_value) { inline_val = a.substr(a[2] == '=' ? 3 : 2); has_inline = true; }
                else {   // combined short flags: -bO
                    for (size_t j = 1; j < a.size(); ++j) {
                        const OptSpec *s2 = nullptr;
                        for (const OptSpec &o : OPTS) if (o.short_name && a[j] == o.short_name && !o.takes_value) s2 = &o;
                        if (!s2 || !allowed(m.sub, s2->name)) bail("Found argument '" + a + "' which wasn't expected, or isn't valid in this context");
                        m.vals[s2->name];
                    }
                    continue;
                }
            }
        }
        if (!spec || !allowed(m.sub, spec->name)) bail("Found argument '" + a + "' which wasn't expected, or isn't valid in this context");
        if (!spec->takes_value) { m.vals[spec->name]; continue; }
        if (has_inline) m.vals[spec->name].push_back(inline_val);
        else {
            if (i + 1 >= argc) bail(std::string("The argument '--") + spec->long_name + " <" + spec->name + ">' requires a value but none was supplied");
            m.vals[spec->name].push_back(argv[++i]);
        }
        if (spec->multiple) open_multi = spec;
    }
    if (m.inputs.empty()) bail("The following required arguments were not provided:\n    <INPUT>...");
    // conflicts (cli.rs:53,77,88,138,219; sketch_type possible values cli.rs:166)
    auto conflict = [&](const char *a, const char *b) {
        if (m.is_present(a) && m.is_present(b)) bail(std::string("The argument '--") + a + "' cannot be used with '--" + b + "'");
    };
    conflict("binary_format", "mash_binary_format"); conflict("pairwise", "queries");
    conflict("no_filter", "filter"); conflict("output_file", "std_out");
    const std::string st = m.value_of("sketch_type");
    if (st != "mash" && st != "scaled" && st != "none") bail("'" + st + "' isn't a valid value for '--sketch-type <sketch_type>'\n\t[possible values: mash, none, scaled]");
    return m;
}

// get_int_arg / get_float_arg (cli.rs:234-262)
std::string display_key(const char *key) { std::string k = key; std::replace(k.begin(), k.end(), '_', '-'); return k; }
uint64_t get_int_arg(const Matches &m, const char *key, uint64_t max = UINT64_MAX) {
    const char *v = m.value_of(key);
    if (!v) bail("Bad " + display_key(key));
    const std::string s = v;
    uint64_t out = 0;
    const char *b = s.data() + (!s.empty() && s[0] == '+' ? 1 : 0);
    auto r = std::from_chars(b, s.data() + s.size(), out);
    if (s.empty() || r.ec != std::errc() || r.ptr != s.data() + s.size() || out > max) bail(display_key(key) + " must be a positive integer");
    return out;
}
double get_float_arg(const Matches &m, const char *key, double limit) {
    const char *v = m.value_of(key);
    if (!v) bail("Bad " + display_key(key));
    char *end = nullptr;
    const double r = strtod(v, &end);
    if (end == v || *end) bail(display_key(key) + " must be a number");
    if (0.0 <= r && r <= limit) return r;
    bail(display_key(key) + " must be between 0 and " + rust_display_f64(limit));
}
FilterParams parse_filter_options(const Matches &m, uint8_t k) {   // cli.rs:241-283
    FilterParams f;
    f.filter_on = m.is_present("filter") ? 1 : (m.is_present("no_filter") ? 0 : -1);
    if (m.occurrences_of("min_abun_filter")) { f.has_lo = true; f.lo = (uint32_t)get_int_arg(m, "min_abun_filter", UINT32_MAX); }
    if (m.occurrences_of("max_abun_filter")) { f.has_hi = true; f.hi = (uint32_t)get_int_arg(m, "max_abun_filter", UINT32_MAX); }
    double err = get_float_arg(m, "err_filter", 100.0 / (double)k);
    err *= (double)k / 100.0;
    f.err_filter = err;
    f.strand_filter = get_float_arg(m, "strand_filter", 1.0);
    return f;
}
SketchParams parse_sketch_options(const Matches &m, uint8_t k, int filters_enabled) {   // cli.rs:285-340
    SketchParams p;
    p.kmer_length = k;
    const std::string st = m.value_of("sketch_type");
    if (st == "mash") {
        if (m.occurrences_of("scale")) bail("`scale` can not be specified for `mash` sketch types");
        const uint64_t final_size = get_int_arg(m, "n_hashes"), oversketch = get_int_arg(m, "oversketch");
        p.kind = Kind::Mash;
        p.final_size = final_size;
        p.kmers_to_sketch = filters_enabled == 0 ? final_size : final_size * oversketch;
        p.no_strict = m.is_present("no_strict");
        p.hash_seed = get_int_arg(m, "seed");
    } else if (st == "scaled") {
        if (m.occurrences_of("oversketch")) bail("`oversketch` can not be specified for `scaled` sketch types");
        if (m.occurrences_of("no_strict")) bail("`no_strict` can not be specified for `scaled` sketch types");
        p.kind = Kind::Scaled;
        p.kmers_to_sketch = get_int_arg(m, "n_hashes");
        p.scale = get_float_arg(m, "scale", 1.0);
        p.hash_seed = get_int_arg(m, "seed");
    } else {
        for (const char *k2 : {"n_hashes", "seed", "oversketch", "no_strict", "scale"})
            if (m.occurrences_of(k2)) bail(std::string("`") + k2 + "` can not be specified for `none` sketch types");
        p.kind = Kind::AllCounts;
    }
    return p;
}

// ---- the GPU calls ----------------------------------------------------------------------------------
fb2_params to_c(const SketchParams &p) {
    fb2_params c; memset(&c, 0, sizeof c);
    c.kind = p.kind == Kind::Scaled ? FB2_KIND_SCALED : (p.kind == Kind::AllCounts ? FB2_KIND_ALLCOUNTS : FB2_KIND_MASH);
    c.kmers_to_sketch = p.kmers_to_sketch; c.final_size = p.final_size; c.no_strict = p.no_strict;
    c.kmer_length = p.kmer_length; c.hash_seed = p.hash_seed; c.scale = p.scale; c.device = -1; c.stream = nullptr;
    return c;
}
fb2_filter to_c(const FilterParams &f) {
    fb2_filter c; memset(&c, 0, sizeof c);
    c.filter_on = f.filter_on; c.has_abun_low = f.has_lo; c.abun_low = f.lo; c.has_abun_high = f.has_hi; c.abun_high = f.hi;
    c.err_filter = f.err_filter; c.strand_filter = f.strand_filter;
    return c;
}
FilterParams from_c(const fb2_filter &c) {
    FilterParams f;
    f.filter_on = c.filter_on; f.has_lo = c.has_abun_low != 0; f.lo = c.abun_low; f.has_hi = c.has_abun_high != 0; f.hi = c.abun_high;
    f.err_filter = c.err_filter; f.strand_filter = c.strand_filter;
    return f;
}
// finch::sketch_files (lib/src/lib.rs:29-49) through the C ABI
std::vector<Sketch> sketch_files(const std::vector<std::string> &files, const SketchParams &p, const FilterParams &f) {
    std::vector<Sketch> out;
    if (files.empty()) return out;
    std::vector<const char *> paths;
    for (auto &s : files) paths.push_back(s.c_str());
    std::vector<fb2_result> res(files.size());
    const fb2_params cp = to_c(p);
    const fb2_filter cf = to_c(f);
    if (fb2_sketch_files(paths.data(), paths.size(), &cp, &cf, res.data()) != FB2_OK) bail(fb2_last_error());
    for (size_t i = 0; i < files.size(); ++i) {
        const fb2_result &r = res[i];
        Sketch s;
        s.name = files[i]; s.seq_length = r.seq_length; s.num_valid_kmers = r.num_valid_kmers;
        s.hashes.assign(r.hashes, r.hashes + r.n); s.counts.assign(r.counts, r.counts + r.n); s.extras.assign(r.extras, r.extras + r.n);
        s.kmers.resize(r.n);
        for (uint64_t j = 0; j < r.n; ++j) s.kmers[j].assign((const char *)r.kmers + j * r.kmer_stride, p.kmer_length);
        s.filter_params = from_c(r.filters);
        s.sketch_params = p;                      // lib.rs:92: the caller's params, cloned
        out.push_back(std::move(s));
        fb2_result_free(&res[i]);
    }
    return out;
}

std::string read_file(const std::string &path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) bail("Error opening \"" + path + "\"");
    std::ostringstream ss; ss << in.rdbuf();
    return ss.str();
}
std::vector<Sketch> open_sketch_file(const std::string &path) {          // lib/src/lib.rs:96-117: by extension
    const std::string data = read_file(path);
    try {
        if (ends_with(path, MASH_EXT)) return read_mash_file(data.data(), data.size());
        if (ends_with(path, FINCH_BIN_EXT)) return read_finch_file(data.data(), data.size());
        return read_multisketch_json(data.data(), data.size());
    } catch (const std::exception &e) { bail("Error parsing \"" + path + "\" (" + e.what() + ")"); }
}
// the three sketch file formats (main.rs:52-84): by the extension the flags picked
std::string serialize_sketches(const std::vector<Sketch> &sk, const std::string &ext) {
    if (ext == FINCH_BIN_EXT) return write_finch_file(sk);
    if (ext == MASH_EXT) return write_mash_file(sk, params_from_sketches(sk));
    return write_multisketch_json(sk);
}

// FilterParams::filter_sketch (filtering.rs:20-52): the filtered hashes are computed and DROPPED by
// the reference; only the sketch's filter metadata changes.
void filter_sketch(const FilterParams &self, Sketch &s) {
    FilterParams &t = s.filter_params;
    t.filter_on = self.filter_on;
    const uint32_t cur_lo = t.has_lo ? t.lo : 0u, cur_hi = t.has_hi ? t.hi : UINT32_MAX;
    if (self.has_lo) { t.has_lo = true; t.lo = std::max(self.lo, cur_lo); } else t.has_lo = false;
    if (self.has_hi) { t.has_hi = true; t.hi = std::min(self.hi, cur_hi); } else t.has_hi = false;
    t.err_filter = std::max(t.err_filter, self.err_filter);
    t.strand_filter = std::max(t.strand_filter, self.strand_filter);
}

void update_sketch_params(const Matches &m, SketchParams &p, const Sketch &first, const std::string &name) {   // main.rs:336-441
    const SketchParams &n = first.sketch_params;
    if (p.kind != n.kind) bail("Sketch types are not the same");
    auto check_k = [&]() {
        if (!m.occurrences_of("kmer_length")) p.kmer_length = n.kmer_length;
        else if (p.kmer_length != n.kmer_length)
            bail("Specified kmer length " + std::to_string(p.kmer_length) + " does not match " + std::to_string(n.kmer_length) + " from sketch " + name);
    };
    auto check_seed = [&]() {
        if (!m.occurrences_of("seed")) p.hash_seed = n.seed();
        else if (p.hash_seed != n.seed())
            bail("Specified hash seed " + std::to_string(p.hash_seed) + " does not match " + std::to_string(n.seed()) + " from sketch " + name);
    };
    if (p.kind == Kind::Mash) {
        if (!m.occurrences_of("n_hashes")) p.final_size = n.expected_size();
        check_k(); check_seed();
    } else if (p.kind == Kind::Scaled) {
        check_k(); check_seed();
        if (n.has_scale()) {
            if (!m.occurrences_of("scale")) p.scale = n.scale;
            else if (std::fabs(p.scale - n.scale) < 2.220446049250313e-16)   // sic: the reference bails when they are EQUAL (Q11)
                bail("Specified scale " + rust_display_f64(p.scale) + " does not match " + rust_display_f64(n.scale) + " from sketch " + name);
        }
    } else check_k();
}

std::vector<Sketch> parse_mash_files(const Matches &m) {                  // main.rs:237-313
    std::vector<std::string> sketch_files_, seq_files;
    for (auto &f : m.inputs) (is_sketch_file(f) ? sketch_files_ : seq_files).push_back(f);
    const uint8_t k = (uint8_t)get_int_arg(m, "kmer_length", 255);
    FilterParams filters = parse_filter_options(m, k);
    SketchParams params = parse_sketch_options(m, k, filters.filter_on);
    std::vector<Sketch> sketches;
    if (!sketch_files_.empty()) {
        sketches = open_sketch_file(sketch_files_[0]);
        if (sketches.empty()) bail("index out of bounds: sketch file " + sketch_files_[0] + " holds no sketches");
        update_sketch_params(m, params, sketches[0], sketch_files_[0]);
        if (!m.occurrences_of("kmer_length")) filters = parse_filter_options(m, params.kmer_length);
        if (filters.filter_on == 1) for (auto &s : sketches) filter_sketch(filters, s);
        for (size_t i = 1; i < sketch_files_.size(); ++i) {
            std::vector<Sketch> extra = open_sketch_file(sketch_files_[i]);
            for (auto &s : extra) {
                std::string nm, v1, v2;
                if (!params.compatible(s.sketch_params, nm, v1, v2))
                    bail("Sketch " + s.name + " has " + nm + " " + v2 + ", but working value is " + v1);
            }
            for (auto &s : extra) sketches.push_back(std::move(s));
            if (filters.filter_on == 1) for (auto &s : sketches) filter_sketch(filters, s);
        }
    }
    std::vector<Sketch> extra = sketch_files(seq_files, params, filters);
    for (auto &s : extra) sketches.push_back(std::move(s));
    return sketches;
}

// output_to (main.rs:21-46)
void output_to(const std::string &payload, const char *output, const std::string &ext) {
    if (!output) { fwrite(payload.data(), 1, payload.size(), stdout); fflush(stdout); return; }
    std::string fn = output;
    if (!ends_with(fn, ext)) fn += ext;
    std::ofstream out(fn, std::ios::binary);
    if (!out) bail("unable to create '" + fn + "'");
    out.write(payload.data(), (std::streamsize)payload.size());
}

// distance() for every (query, reference) pair that calc_sketch_distances keeps (main.rs:315-334, distance.rs:9-47).
// Queries are always elements of `refs` (main.rs:92-113 picks them out of the parsed sketches), so every sketch is
// one row of the hash matrix.  Usual case (one scale for every pair, contiguous query rows: --pairwise, the default
// first-sketch query): the tiled all-pairs kernel with the max_distance cut applied on the device
// (fb2_dist_all_pairs_cut); the host finishes the f64 fields of the survivors and applies main.rs:328 exactly.
// Mixed scales, scattered --queries and --old-dist take the pair-list kernel (fb2_dist_batch).
std::vector<SketchDistance> calc_sketch_distances(const std::vector<const Sketch *> &queries, const std::vector<Sketch> &refs, bool old_mode, double max_dist) {
    std::vector<SketchDistance> out;
    if (queries.empty() || refs.empty()) return out;
    std::unordered_map<const Sketch *, uint32_t> row_of;
    for (size_t r = 0; r < refs.size(); ++r) row_of.emplace(&refs[r], (uint32_t)r);
    std::vector<uint32_t> qrow(queries.size());
    for (size_t i = 0; i < queries.size(); ++i) {
        auto it = row_of.find(queries[i]);
        if (it == row_of.end()) bail("internal: query sketch is not one of the parsed sketches");
        qrow[i] = it->second;
    }
    size_t stride = 1;
    for (auto &s : refs) stride = std::max(stride, s.hashes.size());
    std::vector<uint64_t> mat(refs.size() * stride, 0);
    std::vector<uint32_t> lens(refs.size());
    for (size_t i = 0; i < refs.size(); ++i) {
        lens[i] = (uint32_t)refs[i].hashes.size();
        std::copy(refs[i].hashes.begin(), refs[i].hashes.end(), mat.begin() + (long)(i * stride));
    }
    auto pair_scale = [&](const Sketch &q, const Sketch &r) -> double {   // distance.rs:23-28 (old mode: no scale)
        if (!old_mode && q.sketch_params.has_scale() && r.sketch_params.has_scale()) return std::min(q.sketch_params.scale, r.sketch_params.scale);
        return 0.0;
    };
    auto emit = [&](const Sketch &q, const Sketch &r, const fb2_pair_out &po) {
        SketchDistance d;
        if (old_mode) {                                                   // old_distance (distance.rs:136-157)
            if (fb2_old_distance_finish(po.common, q.hashes.size(), r.hashes.size(), q.sketch_params.kmer_length,
                                        &d.containment, &d.jaccard, &d.mash_distance, &d.common_hashes, &d.total_hashes) != FB2_OK)
                bail(fb2_last_error());
        } else
            fb2_distance_finish(&po, q.sketch_params.kmer_length, &d.containment, &d.jaccard, &d.mash_distance, &d.common_hashes, &d.total_hashes);
        if (!(d.mash_distance <= max_dist)) return;                       // main.rs:328
        d.query = q.name; d.reference = r.name;
        out.push_back(std::move(d));
    };
    // ---- fast path ----
    bool uniform = !old_mode, contiguous = true;
    const bool sc0 = refs[0].sketch_params.has_scale();
    for (auto &s : refs)
        if (s.sketch_params.has_scale() != sc0 || (sc0 && s.sketch_params.scale != refs[0].sketch_params.scale) ||
            s.sketch_params.kmer_length != refs[0].sketch_params.kmer_length) { uniform = false; break; }
    for (size_t i = 1; i < qrow.size(); ++i) if (qrow[i] != qrow[i - 1] + 1) contiguous = false;
    if (uniform && contiguous && !getenv("FINCH_DIST_PAIR_LIST")) {
        const double scale = sc0 ? refs[0].sketch_params.scale : 0.0;
        const size_t q0 = qrow.front(), q1 = (size_t)qrow.back() + 1;
        std::vector<fb2_pair_hit> hits((size_t)1 << 16);
        uint64_t n_hits = 0;
        int rc;
        while ((rc = fb2_dist_all_pairs_cut(mat.data(), lens.data(), refs.size(), stride, scale, q0, q1, refs[0].sketch_params.kmer_length,
                                            max_dist, 1, hits.data(), hits.size(), &n_hits, -1, 0)) == FB2_ENOMEM && n_hits > hits.size())
            hits.resize((size_t)n_hits);
        if (rc != FB2_OK) bail(fb2_last_error());
        hits.resize((size_t)n_hits);
        // reference-major, query-minor (main.rs:321-323): stable counting sort of the (q, r)-ordered hits by r
        std::vector<uint64_t> start(refs.size() + 1, 0);
        for (auto &h : hits) start[h.r + 1]++;
        for (size_t r = 0; r < refs.size(); ++r) start[r + 1] += start[r];
        std::vector<fb2_pair_hit> by_ref(hits.size());
        for (auto &h : hits) by_ref[start[h.r]++] = h;
        for (auto &h : by_ref) {
            const Sketch &q = refs[h.q], &r = refs[h.r];
            if (q == r) continue;                                         // main.rs:324: equal BY VALUE (Q12); q == r by index never comes back
            emit(q, r, fb2_pair_out{h.common, h.i, h.j});
        }
        return out;
    }
    // ---- pair list ----
    struct Pair { uint32_t q, r; double scale; };
    std::vector<Pair> pairs;
    for (size_t r = 0; r < refs.size(); ++r)
        for (size_t i = 0; i < queries.size(); ++i) {
            if (qrow[i] == r || *queries[i] == refs[r]) continue;         // main.rs:324: equal BY VALUE (Q12)
            pairs.push_back({qrow[i], (uint32_t)r, pair_scale(*queries[i], refs[r])});
        }
    if (pairs.empty()) return out;
    std::vector<fb2_pair_out> po(pairs.size());
    std::set<double> scales;
    for (auto &p : pairs) scales.insert(p.scale);
    for (double sc : scales) {
        std::vector<uint32_t> qi, ri, where;
        for (size_t i = 0; i < pairs.size(); ++i) if (pairs[i].scale == sc) { qi.push_back(pairs[i].q); ri.push_back(pairs[i].r); where.push_back((uint32_t)i); }
        std::vector<fb2_pair_out> tmp(qi.size());
        if (fb2_dist_batch(mat.data(), lens.data(), refs.size(), stride, sc, qi.data(), ri.data(), qi.size(), tmp.data(), -1) != FB2_OK)
            bail(fb2_last_error());
        for (size_t i = 0; i < where.size(); ++i) po[where[i]] = tmp[i];
    }
    for (size_t i = 0; i < pairs.size(); ++i) emit(refs[pairs[i].q], refs[pairs[i].r], po[i]);
    return out;
}

int run(int argc, char **argv) {
    const Matches m = parse_args(argc, argv);
    if (m.sub == "sketch") {
        const std::string file_ext = m.is_present("binary_format") ? FINCH_BIN_EXT : (m.is_present("mash_binary_format") ? MASH_EXT : FINCH_EXT);
        if (m.is_present("output_file") || m.is_present("std_out")) {
            const std::vector<Sketch> sk = parse_mash_files(m);
            output_to(serialize_sketches(sk, file_ext), m.is_present("output_file") ? m.value_of("output_file") : nullptr, file_ext);
        } else {                                                          // generate_sketch_files (main.rs:201-235)
            const uint8_t k = (uint8_t)get_int_arg(m, "kmer_length", 255);
            const FilterParams filters = parse_filter_options(m, k);
            const SketchParams params = parse_sketch_options(m, k, filters.filter_on);
            for (auto &fn : m.inputs) {
                if (is_sketch_file(fn)) bail("Filename " + fn + " is not a sequence file?");
                const std::vector<Sketch> sk = sketch_files({fn}, params, filters);
                std::ofstream out(fn + file_ext, std::ios::binary);
                if (!out) bail("Could not open " + fn + file_ext);
                const std::string js = serialize_sketches(sk, file_ext);
                out.write(js.data(), (std::streamsize)js.size());
            }
        }
    } else if (m.sub == "dist") {
        const bool old_mode = m.is_present("old_dist_mode");
        const double max_dist = get_float_arg(m, "max_distance", 1.0);
        const std::vector<Sketch> all = parse_mash_files(m);
        std::vector<const Sketch *> queries;
        if (m.is_present("pairwise")) for (auto &s : all) queries.push_back(&s);
        else if (m.is_present("queries")) {
            std::set<std::string> names;
            auto it = m.vals.find("queries");
            for (auto &n : it->second) names.insert(n);
            for (auto &s : all) if (names.count(s.name)) queries.push_back(&s);
        } else {
            if (all.empty()) bail("No sketches present!");
            queries.push_back(&all[0]);
        }
        output_to(write_distances_json(calc_sketch_distances(queries, all, old_mode, max_dist)),
                  m.is_present("output_file") ? m.value_of("output_file") : nullptr, ".json");
    } else if (m.sub == "hist") {
        const std::vector<Sketch> all = parse_mash_files(m);
        std::vector<std::pair<std::string, std::vector<uint64_t>>> hm;    // HashMap<String, Vec<u64>>: later duplicates win
        for (auto &s : all) {
            bool found = false;
            for (auto &kv : hm) if (kv.first == s.name) { kv.second = hist(s.counts); found = true; }
            if (!found) hm.emplace_back(s.name, hist(s.counts));
        }
        std::string o = "{";
        for (size_t i = 0; i < hm.size(); ++i) {
            if (i) o += ',';
            o += json_string(hm[i].first) + ":[";
            for (size_t j = 0; j < hm[i].second.size(); ++j) { if (j) o += ','; o += std::to_string(hm[i].second[j]); }
            o += ']';
        }
        o += '}';
        output_to(o, m.is_present("output_file") ? m.value_of("output_file") : nullptr, ".json");
    } else {   // info (main.rs:148-195); all arithmetic in f32 as the reference does it
        const std::vector<Sketch> all = parse_mash_files(m);
        for (auto &s : all) {
            printf("%s (from %llubp)\n", s.name.c_str(), (unsigned long long)s.seq_length);
            unsigned long long card = 0;                                  // statistics.rs:8-23
            if (!s.hashes.empty()) {
                const float v = (float)(s.hashes.size() - 1) / ((float)s.hashes.back() / (float)UINT64_MAX);
                card = std::isnan(v) || v <= 0.0f ? 0ULL : (v >= 18446744073709551616.0f ? UINT64_MAX : (unsigned long long)v);   // Rust `as u64` saturates
            }
            printf("  Estimated # of Unique Kmers: %llu\n", card);
            const std::vector<uint64_t> h = hist(s.counts);
            float m0 = 0.0f, m1 = 0.0f;
            for (size_t i = 0; i < h.size(); ++i) { m0 += ((float)i + 1.0f) * (float)h[i]; m1 += (float)h[i]; }
            printf("  Estimated Average Depth: %sx\n", rust_display_f32(m0 / m1).c_str());
            uint64_t total_gc = 0;
            for (size_t i = 0; i < s.kmers.size(); ++i)
                for (char c : s.kmers[i]) if (c == 'G' || c == 'g' || c == 'C' || c == 'c') total_gc += s.counts[i];
            const float total_bases = s.hashes.empty() ? 0.0f : m0 * (float)s.kmers[0].size();
            printf("  Estimated %% GC: %s%%\n", rust_display_f32(100.0f * (float)total_gc / total_bases).c_str());
        }
    }
    return 0;
}

}  // namespace

int main(int argc, char **argv) {
    try { return run(argc, argv); }
    catch (const std::exception &e) { fprintf(stderr, "Error: %s\n", e.what()); return 1; }
}

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../finch_rs_b200/csrc/strip.h"
// usage: strip_host_test <threads> < fastq   -> prints records bases bad len_bad first_blank last_nonblank tail_len, then the lines
int main(int argc, char **argv) {
    const unsigned T = argc > 1 ? (unsigned)atoi(argv[1]) : 1;
    std::vector<uint8_t> in;
    uint8_t buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, stdin)) > 0) in.insert(in.end(), buf, buf + n);
    std::vector<fb2::StripOut> outs(T);
    std::vector<std::vector<uint8_t>> store(T);
    const size_t each = in.size() / T / 2 + 4096;
    for (unsigned t = 0; t < T; ++t) { store[t].resize(each); outs[t].out = store[t].data(); outs[t].out_cap = each; }
    const uint8_t *e = fb2::strip_parallel(in.data(), in.data() + in.size(), 0, T, outs);
    fb2::StripOut fin;
    fin.origin = in.data();
    fin.spill.resize((size_t)(in.data() + in.size() - e) / 2 + 16);
    fin.spilled = true;
    const int bad_tail = fb2::strip_final(e, in.data() + in.size(), (uint64_t)(e - in.data()), fin);
    unsigned long long recs = fin.records, bases = fin.bases, bad = ~0ull, lbad = ~0ull, fblank = ~0ull, lnb = fin.last_nonblank;
    for (auto &o : outs) { recs += o.records; bases += o.bases; if (o.bad_pos < bad) bad = o.bad_pos; if (o.len_bad_pos < lbad) lbad = o.len_bad_pos;
                           if (o.first_blank < fblank) fblank = o.first_blank;
                           if (o.last_nonblank > lnb) lnb = o.last_nonblank; }
    const int invalid = bad != ~0ull || lbad != ~0ull || (fblank != ~0ull && fblank < lnb) || bad_tail;
    printf("%llu %llu %d\n", recs, bases, invalid);
    for (auto &o : outs) fwrite(o.data(), 1, o.out_len, stdout);
    fwrite(fin.spill.data(), 1, fin.out_len, stdout);
    return 0;
}

// The reference's own unit tests, through the header-only C++ mirror (finch_rs_b200/host/finch_b200.hpp):
//   mash.rs:115-134 / scaled.rs:118-161 (test_minhashkmers), scaled.rs:163-176 (eviction), parameters() quirks Q7/Q8.
// Built by tests/test_abi_cpu.py (links libfinch_b200.so), run on the GPU box by tests/test_gpu_parity.py.
#include <cstdio>
#include <cstring>
#include <string>

#include "../finch_rs_b200/host/finch_b200.hpp"

#define REQUIRE(c) do { if (!(c)) { fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

static std::string str(const std::vector<uint8_t> &v) { return std::string(v.begin(), v.end()); }

int main(int argc, char **argv) {
    if (argc > 1 && !strcmp(argv[1], "--link-only")) { printf("%s\n", fb2_version()); return 0; }
    if (argc > 2 && !strcmp(argv[1], "--files")) {   // host only: sketch files through the mirror (no device needed)
        try {
            const std::string dir = argv[2];
            std::vector<finch::Sketch> sk(2);
            sk[0].name = "a.fa"; sk[0].comment = "first"; sk[0].seq_length = 100; sk[0].num_valid_kmers = 80;
            sk[0].sketch_params = finch::SketchParams::Mash(3, 3, false, 4, 0);
            sk[0].filter_params.filter_on = 0;
            const char *kmers[3] = {"ACGT", "CCCC", "TTGA"};
            for (int i = 0; i < 3; ++i) {
                finch::KmerCount k;
                k.hash = 10u + 7u * (uint64_t)i; k.count = 1 + i; k.extra_count = i; k.kmer.assign(kmers[i], kmers[i] + 4);
                sk[0].hashes.push_back(k);
            }
            sk[1] = sk[0]; sk[1].name = "b.fa"; sk[1].comment = ""; sk[1].hashes.pop_back();
            for (int fmt : {FB2_FILE_SK, FB2_FILE_BSK, FB2_FILE_MSH}) {
                const std::string path = dir + (fmt == FB2_FILE_SK ? "/m.sk" : fmt == FB2_FILE_BSK ? "/m.bsk" : "/m.msh");
                finch::write_sketch_file(path, sk, fmt);
                const auto back = finch::open_sketch_file(path);
                REQUIRE(back.size() == 2 && back[0].name == "a.fa" && back[1].name == "b.fa");
                REQUIRE(back[0].hashes.size() == 3 && back[1].hashes.size() == 2);
                REQUIRE(back[0].seq_length == 100 && back[0].sketch_params.kmer_length == 4);
                for (int i = 0; i < 3; ++i) REQUIRE(back[0].hashes[i].hash == sk[0].hashes[i].hash);
                if (fmt == FB2_FILE_BSK) {
                    REQUIRE(back[0].comment == "first" && back[0].num_valid_kmers == 80);
                    for (int i = 0; i < 3; ++i)
                        REQUIRE(str(back[0].hashes[i].kmer) == kmers[i] && back[0].hashes[i].count == (uint32_t)(1 + i) &&
                                back[0].hashes[i].extra_count == (uint32_t)i);
                }
            }
            bool threw = false;
            try { finch::open_sketch_file(dir + "/m.txt"); } catch (const finch::FinchError &e) { threw = std::string(e.what()).find("suffix") != std::string::npos || e.code == FB2_EIO; }
            REQUIRE(threw);
        } catch (const finch::FinchError &e) {
            fprintf(stderr, "FinchError %d: %s\n", e.code, e.what());
            return 1;
        }
        printf("files ok\n");
        return 0;
    }
    try {
        for (int which = 0; which < 2; ++which) {
            auto q = which == 0 ? finch::MashSketcher(3, 2, 42) : finch::ScaledSketcher(3, 1.0, 2, 42);
            q->push((const uint8_t *)"ca", 2, 0); q->push((const uint8_t *)"cc", 2, 1);
            q->push((const uint8_t *)"ac", 2, 0); q->push((const uint8_t *)"ac", 2, 1);
            const auto v = q->to_vec();
            REQUIRE(v.size() == 3);
            REQUIRE(str(v[0].kmer) == "cc" && v[0].count == 1 && v[0].extra_count == 1);
            REQUIRE(str(v[1].kmer) == "ca" && v[1].count == 1 && v[1].extra_count == 0);
            REQUIRE(str(v[2].kmer) == "ac" && v[2].count == 2 && v[2].extra_count == 1);
            REQUIRE(v[0].hash < v[1].hash && v[1].hash < v[2].hash);
        }
        {   // scaled.rs:163-176: AAAA hashes above u64::MAX / 100 and is evicted
            auto q = finch::ScaledSketcher(1, 0.01, 4, 42);
            for (const char *k : {"AAAA", "AGTA", "CCCC", "ATAA"}) q->push((const uint8_t *)k, 4, 0);
            const auto v = q->to_vec();
            REQUIRE(v.size() == 3);
            for (auto &e : v) REQUIRE(str(e.kmer) != "AAAA");
        }
        {   // pushed k-mers longer than kmer_length keep their bytes (mash.rs:52-55 stores kmer.to_owned())
            auto q = finch::MashSketcher(5, 2, 0);
            q->push((const uint8_t *)"ACGTACGT", 8, 0);
            const auto v = q->to_vec();
            REQUIRE(v.size() == 1 && str(v[0].kmer) == "ACGTACGT");
        }
        {   // Q8 / Q7
            finch::SketchScheme m(finch::SketchParams::Mash(200000, 1000, true, 21, 7));
            const fb2_params pm = m.parameters();
            REQUIRE(pm.final_size == 200000 && pm.no_strict == 0 && pm.kmers_to_sketch == 200000 && pm.hash_seed == 7);
            finch::SketchScheme s(finch::SketchParams::Scaled(10, 21, 0.001, 0));
            const fb2_params ps = s.parameters();
            REQUIRE(ps.scale == 1.0 / ((double)UINT64_MAX / (double)(UINT64_MAX / 1000)));
        }
        {   // process(): total_bases counts the raw record bytes (mash.rs:72)
            auto q = finch::MashSketcher(10, 3, 0);
            const char *rec = "ACGTN\nacgu";
            q->process((const uint8_t *)rec, strlen(rec));
            const auto t = q->total_bases_and_kmers();
            REQUIRE(t.first == 10 && t.second == 2 + 2);
        }
    } catch (const finch::FinchError &e) {
        fprintf(stderr, "FinchError [%d] %s\n", e.code, e.what());
        return 2;
    }
    printf("hpp mirror ok\n");
    return 0;
}

// build_gesa -- the index-construction step of the pipeline on the GPU: reads X (FASTA) and writes X.gesa (or the BCR
// triple X.out / X.out.lcp / X.out.pairSA with -b), the files ebwt2clust / clust2snp -i X look for
// (ref:include.hpp:42-81).  Stands in for the external `egsa` / BCR run of ref:pipeline.sh:98-109 and
// ref:README.md:46-60; record layout as egsa_stream reads it: text(y) suff(z) lcp(x) bwt(1), little endian
// (ref:include.hpp:126-155).  -x / -y / -z have the meaning and the defaults (1 / 4 / 1) of the two tools, so the same
// flags can be passed to all three.  Reads may have any lengths below 65536 (empty ones included, as the reference's
// FASTA parser keeps them, ref:clust2snp.cpp:147-212); bases must be ACGT / acgt.
#include <getopt.h>

#include <string>
#include <thread>
#include <vector>

#include "host_io.hpp"

static void help() {
    printf("build_gesa [options]\nOptions:\n"
           "-h          Print this help.\n"
           "-i <arg>    Input fasta file (REQUIRED). Output: <arg>.gesa\n"
           "-b          Write the BCR triple <arg>.out, <arg>.out.lcp, <arg>.out.pairSA instead.\n"
           "-x <arg>    Byte-size of LCP values in the output (1, 2, 4 or 8; default 1).\n"
           "-y <arg>    Byte-size of DA values (read number) in the output (default 4).\n"
           "-z <arg>    Byte-size of pos values (position in read) in the output (default 1).\n");
    exit(0);
}

static void put_le(uint8_t* p, uint64_t v, int nb) {
    for (int b = 0; b < nb; ++b) p[b] = uint8_t(v >> (8 * b));
}

int main(int argc, char** argv) {
    std::string input;
    int x = 0, y = 0, z = 0;
    bool bcr = false;
    if (argc < 2) help();
    int opt;
    while ((opt = getopt(argc, argv, "hbi:x:y:z:")) != -1) {
        switch (opt) {
            case 'i': input = optarg; break;
            case 'b': bcr = true; break;
            case 'x': x = atoi(optarg); break;
            case 'y': y = atoi(optarg); break;
            case 'z': z = atoi(optarg); break;
            default: help();
        }
    }
    x = x == 0 ? 1 : x;
    y = y == 0 ? 4 : y;
    z = z == 0 ? 1 : z;
    auto ok = [](int v) { return v == 1 || v == 2 || v == 4 || v == 8; };
    if (input.empty() || !ok(x) || !ok(y) || !ok(z)) help();
    printf("This is build_gesa. Input file: %s\n", input.c_str());
    host::Reads reads;
    if (!reads.load(input) || reads.n_reads() == 0) {
        printf("Error: could not read %s\n", input.c_str());
        return 1;
    }
    const uint64_t R = reads.n_reads();
    uint64_t L = 0;  // the longest read
    bool equal = true;
    for (uint64_t r = 0; r < R; ++r) {
        const uint64_t l = reads.off[r + 1] - reads.off[r];
        L = l > L ? l : L;
        equal = equal && l == reads.off[1] - reads.off[0];
    }
    if (L == 0) {
        printf("Error: every read of %s is empty.\n", input.c_str());
        return 2;
    }
    const uint64_t n = reads.bases.size() + R;
    auto fits = [](uint64_t v, int nb) { return nb >= 8 || v < (uint64_t(1) << (8 * nb)); };
    if (!fits(R - 1, y) || !fits(L, z) || !fits(L, x))
        printf("Warning: values do not fit the requested field widths and will be truncated (%llu reads, longest %llu bases).\n",
               (unsigned long long)R, (unsigned long long)L);
    e2s_ctx* ctx = nullptr;
    if (e2s_ctx_create(0, &ctx) != E2S_OK) {
        printf("Error: %s\n", e2s_last_error(nullptr));
        return 3;
    }
    std::vector<uint32_t> lcp(n), text(n), suff(n);
    std::vector<uint8_t> bwt(n);
    const int brc = equal ? e2s_build_egsa(ctx, reads.bases.data(), R, uint32_t(L), lcp.data(), text.data(), suff.data(), bwt.data())
                          : e2s_build_egsa_ragged(ctx, reads.bases.data(), reads.off.data(), R, lcp.data(), text.data(), suff.data(), bwt.data());
    if (brc != E2S_OK) {
        printf("Error: %s\n", e2s_last_error(ctx));
        e2s_ctx_destroy(ctx);
        return 3;
    }
    e2s_ctx_destroy(ctx);
    // the records in file layout: disjoint ranges of the index, one host thread each
    unsigned nt = std::thread::hardware_concurrency();
    nt = nt < 1 ? 1 : (nt > 32 ? 32 : nt);
    if (n < (uint64_t(1) << 20)) nt = 1;
    auto in_ranges = [&](auto&& body) {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t) {
            const uint64_t a = n / nt * t + (t < n % nt ? t : n % nt), b = a + n / nt + (t < n % nt ? 1 : 0);
            th.emplace_back([&body, a, b]() { body(a, b); });
        }
        for (auto& x : th) x.join();
    };
    bool wrote;
    if (!bcr) {
        const size_t rs = size_t(x + y + z + 1);
        std::vector<uint8_t> rec(n * rs);
        in_ranges([&](uint64_t a, uint64_t b) {
            for (uint64_t i = a; i < b; ++i) {
                uint8_t* p = rec.data() + i * rs;
                put_le(p, text[i], y);
                put_le(p + y, suff[i], z);
                put_le(p + y + z, lcp[i], x);
                p[y + z + x] = bwt[i];
            }
        });
        wrote = host::write_all(input + ".gesa", rec.data(), rec.size());
    } else {
        std::vector<uint8_t> l(n * size_t(x)), g(n * size_t(z + y));
        in_ranges([&](uint64_t a, uint64_t b) {
            for (uint64_t i = a; i < b; ++i) {
                put_le(l.data() + i * x, lcp[i], x);
                put_le(g.data() + i * (z + y), suff[i], z);  // suff(z) then text(y): ref:include.hpp:159-175
                put_le(g.data() + i * (z + y) + z, text[i], y);
            }
        });
        wrote = host::write_all(input + ".out", bwt.data(), bwt.size()) && host::write_all(input + ".out.lcp", l.data(), l.size()) &&
                host::write_all(input + ".out.pairSA", g.data(), g.size());
    }
    if (!wrote) {
        printf("Error: could not write the index files next to %s\n", input.c_str());
        return 1;
    }
    printf("Done. %llu suffixes of %llu reads indexed.\n", (unsigned long long)n, (unsigned long long)R);
    return 0;
}

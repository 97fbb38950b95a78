// clust2snp -- drop-in for the reference CLI (ref:clust2snp.cpp:969-1085): same options (including the
// dead ones), same inputs (index, X.clusters, the FASTA), same X.snp output and result-carrying stdout
// lines; statistics, per-cluster analysis and SNP/indel calling run on the GPU(s) through the C ABI.
#include <getopt.h>

#include <iomanip>
#include <iostream>
#include <thread>

#include "host_io.hpp"

// the reference's help text, byte for byte (ref:clust2snp.cpp:64-94 as its binary prints it with the default values;
// tests/golden/help_clust2snp.txt is that output)
static const char* const HELP_TEXT = R"HELP(clust2snp [options]
Options:
-h          Print this help.
-i <arg>    Input fasta file containing the samples' reads (REQUIRED).
-n <arg>    Number of reads in the first sample (REQUIRED).
-L <arg>    Length of left-context, SNP included (default: 31).
-R <arg>    Length of right context, SNP excluded (default: 30).
-g <arg>    Maximum allowed gap length in indel (default: 10). If 0, indels are disabled.
-v <arg>    Maximum number of non-isolated SNPs in left-contexts. The central SNP/indel is excluded from this count (default: 3).
-c <arg>    Extract this maximum number of reads per individual to compute consensus of left-context (default: 20).
-e <arg>    Mismatches allowed between DNA fragments forming consensus of left-context (default: 2).
-m <arg>    Minimum cluster length per individual (default: 5). The minimum cluster length (for the 2 individuals) is 2*<arg>.
-p <arg>    Automatically choose max cluster length so that this fraction of bases is analyzed (default: 
            0.99). In any case, the maximum cluster length will not exceed the value specified with -M.
-M <arg>    Maximum cluster length. Read the description of option -p.
-x <arg>    Byte size of LCP integers in input EGSA/BCR file (default: 1).
-y <arg>    Byte size of DA integers (read number) in input EGSA/BCR file (default: 4).
-z <arg>    Byte size of pos integers (position in read) in input EGSA/BCR file (default: 1).


To run clust2snp, you must first build (1) the Enhanced Generalized Suffix Array of the input sequences
and the  cluster file built with ebwt2snp. Output is stored in reads.snp (this  is actually a fasta
file), where reads.fasta is the input fasta file.

Output:  SNPs are output in KisSNP2 format as a fasta file. IMPORTANT: in many cases, each SNP/indel is
reported twice: one time on the forward strand and one on the reverse strand. 
)HELP";

static void help(const e2s_snp_params&) {
    std::cout << HELP_TEXT << std::flush;
    exit(0);  // ref:clust2snp.cpp:93
}

int main(int argc, char** argv) {
    e2s_snp_params def;
    e2s_snp_default_params(&def);
    if (argc < 3) help(def);
    e2s_snp_params p;
    memset(&p, 0, sizeof p);
    int lcp = 0, da = 0, pos = 0, max_snvs = 0;
    std::string input;
    int opt;
    // 'b' and 'M' are not in the option string: they reach `default` exactly as in the reference
    while ((opt = getopt(argc, argv, "hi:n:p:v:L:R:m:g:c:x:y:z:e:")) != -1) {
        switch (opt) {
            case 'h': help(def); break;
            case 'i': input = optarg; break;
            case 'n': p.nr_reads1 = uint64_t(int64_t(atoi(optarg))); break;
            case 'm': p.mcov_out = atoi(optarg); break;
            case 'g': p.max_gap = atoi(optarg); break;
            case 'L': p.k_left = atoi(optarg); break;
            case 'c': p.consensus_reads = atoi(optarg); break;
            case 'R': p.k_right = atoi(optarg); break;
            case 'p': p.pval = atof(optarg); break;
            case 'v': max_snvs = atoi(optarg); break;
            case 'e': p.max_err = int(atof(optarg)); break;
            case 'x': lcp = atoi(optarg); break;
            case 'y': da = atoi(optarg); break;
            case 'z': pos = atoi(optarg); break;
            default: help(def); return -1;
        }
    }
    (void)max_snvs;  // parsed, never used: the test is against the default 3 (ref:clust2snp.cpp:648)
    lcp = lcp == 0 ? 1 : lcp;
    da = da == 0 ? 4 : da;
    pos = pos == 0 ? 1 : pos;
    p.max_err = p.max_err == 0 ? def.max_err : p.max_err;
    p.consensus_reads = p.consensus_reads == 0 ? def.consensus_reads : p.consensus_reads;
    p.max_gap = p.max_gap == 0 ? def.max_gap : p.max_gap;
    p.k_left = p.k_left == 0 ? def.k_left : p.k_left;
    p.k_right = p.k_right == 0 ? def.k_right : p.k_right;
    p.pval = p.pval == 0 ? def.pval : p.pval;
    p.mcov_out = p.mcov_out == 0 ? def.mcov_out : p.mcov_out;
    p.max_snvs = def.max_snvs;
    if (input.empty() || p.nr_reads1 == 0) help(def);

    host::Index idx;
    if (!idx.open(input, lcp, da, pos)) {
        std::cout << "Error: missing index files." << std::endl;
        return 1;
    }
    std::cout << "This is clust2snp." << std::endl
              << "Input file: " << input << std::endl
              << "Left-extending GSA ranges by " << p.k_left << " bases." << std::endl
              << "Right context length: at most " << p.k_right << " bases." << std::endl;
    const std::string clusters_path = input + ".clusters";
    host::MappedFile cl;
    if (!cl.open(clusters_path)) {
        std::cout << "\nERROR: Could not find BWT clusters file \"" << clusters_path << "\"" << std::endl << std::endl;
        help(def);
    }
    const size_t cut = input.rfind(".fast");
    const std::string out_path = input.substr(0, cut) + ".snp";
    std::cout << "Output events will be stored in " << out_path << std::endl;

    const uint64_t m = cl.size / 10;
    if (m == 0) {
        std::cerr << "clust2snp: empty cluster file (the reference divides by zero here)" << std::endl;
        return 3;
    }
    auto rec_start = [&](uint64_t i) { uint64_t s; memcpy(&s, cl.data + i * 10, 8); return s; };
    auto rec_len = [&](uint64_t i) { uint16_t l; memcpy(&l, cl.data + i * 10 + 8, 2); return l; };

    // ---- shards: positions by contiguous ranges, records by the position of their START ----
    const std::vector<uint64_t> cuts = host::shard_cuts(idx.n, host::gpu_count_from_env());
    const int G = int(cuts.size()) - 1;
    std::vector<uint64_t> rcut(size_t(G) + 1, m);
    rcut[0] = 0;
    for (int g = 1; g < G; ++g) {  // first record with start >= cuts[g]
        uint64_t lo = 0, hi = m;
        while (lo < hi) {
            uint64_t mid = (lo + hi) / 2;
            if (rec_start(mid) < cuts[size_t(g)]) lo = mid + 1;
            else hi = mid;
        }
        rcut[size_t(g)] = lo;
    }
    host::stamp("index + .clusters opened");
    host::Reads reads;
    if (!reads.load(input)) {
        std::cerr << "clust2snp: cannot read " << input << std::endl;
        return 2;
    }

    host::stamp("FASTA parsed");
    std::vector<e2s_ctx*> ctx(size_t(G), nullptr);
    std::vector<e2s_shard*> sh(size_t(G), nullptr);
    std::vector<e2s_stats> stats(static_cast<size_t>(G));
    std::vector<int> rc(size_t(G), 0);
    std::vector<std::string> errs(static_cast<size_t>(G));
    auto fail_of = [&](int g, int r) {
        rc[size_t(g)] = r;
        const char* e = ctx[size_t(g)] ? e2s_last_error(ctx[size_t(g)]) : "";
        errs[size_t(g)] = e[0] ? e : e2s_last_error(nullptr);
    };
    // A shard that does not fit the device (or E2S_CLI_STREAM=1) is STREAMED through a chunked shard of E2S_CHUNK_POSITIONS
    // positions (default 2^28), as the reference streams the index past the records (ref:clust2snp.cpp:806-857): statistics()
    // needs the .clusters file only, each chunk's records are staged with it, the prefilter runs per chunk and phase 2 on the
    // records of its survivors.  A BCR triple then sends X.out.lcp + X.out only; X.out.pairSA is read for the survivors.
    std::vector<char> streamed(size_t(G), 0);
    uint64_t chunk = uint64_t(1) << 28;
    if (const char* e = getenv("E2S_CHUNK_POSITIONS"))
        if (strtoull(e, nullptr, 10)) chunk = strtoull(e, nullptr, 10);
    auto stage = [&](int g) {
        int r = e2s_ctx_create(g, &ctx[size_t(g)]);
        if (r) return fail_of(g, r);
        if (g == 0) host::stamp("context created");
        const uint64_t lo = cuts[size_t(g)], hi = cuts[size_t(g) + 1];
        {
            uint64_t free_b = 0, total_b = 0;
            e2s_ctx_mem_info(ctx[size_t(g)], &free_b, &total_b);
            const uint64_t need = (hi - lo) * 15 + (rcut[size_t(g) + 1] - rcut[size_t(g)]) * 12 + reads.bases.size() + (uint64_t(1) << 30);
            const char* fs = getenv("E2S_CLI_STREAM");
            streamed[size_t(g)] = (fs && atoi(fs) != 0) || need > free_b / 10 * 9;
        }
        if (streamed[size_t(g)]) {  // statistics() of this shard's records on the host; the index is streamed in analyse()
            e2s_stats& s_g = stats[size_t(g)];
            memset(&s_g, 0, sizeof s_g);
            for (uint64_t i = rcut[size_t(g)]; i < rcut[size_t(g) + 1]; ++i) {
                const uint16_t l = rec_len(i);
                if (l <= E2S_MAX_C_LEN) s_g.hist[l]++;
                s_g.n_bases += l;
                s_g.n_clust++;
            }
            r = e2s_reads_stage(ctx[size_t(g)], reads.bases.data(), reads.off.data(), reads.n_reads());
            if (r) fail_of(g, r);
            return;
        }
        r = e2s_shard_create(ctx[size_t(g)], hi - lo, lo, idx.n, &sh[size_t(g)]);
        const uint64_t a = lo >= 2 ? lo - 2 : 0, b = hi + E2S_MAX_C_LEN + 1 < idx.n ? hi + E2S_MAX_C_LEN + 1 : idx.n;
        if (!r) r = idx.load(sh[size_t(g)], a, b - a);
        if (!r) r = e2s_shard_set_layout(sh[size_t(g)], idx.x, idx.y, idx.z, idx.bcr ? 1 : 0);
        if (!r) r = e2s_shard_seal(sh[size_t(g)]);
        if (g == 0) host::stamp("index loaded + sealed");
        if (!r) r = e2s_clusters_stage_packed(sh[size_t(g)], cl.data + rcut[size_t(g)] * 10, rcut[size_t(g) + 1] - rcut[size_t(g)]);
        if (!r) r = e2s_statistics(sh[size_t(g)], &stats[size_t(g)]);
        if (!r) r = e2s_reads_stage(ctx[size_t(g)], reads.bases.data(), reads.off.data(), reads.n_reads());
        if (r) fail_of(g, r);
    };
    {
        std::vector<std::thread> th;
        for (int g = 0; g < G; ++g) th.emplace_back(stage, g);
        for (auto& t : th) t.join();
    }
    host::stamp(".clusters + reads staged, statistics");
    for (int g = 0; g < G; ++g)
        if (rc[size_t(g)]) {
            std::cerr << "clust2snp: GPU " << g << ": " << errs[size_t(g)] << std::endl;
            return 2;
        }

    // ---- statistics(): ref:clust2snp.cpp:877-966 ----
    e2s_stats st;
    memset(&st, 0, sizeof st);
    for (int g = 0; g < G; ++g) {
        for (int i = 0; i < E2S_HIST_BINS; ++i) st.hist[i] += stats[size_t(g)].hist[i];
        st.n_clust += stats[size_t(g)].n_clust;
        st.n_bases += stats[size_t(g)].n_bases;
    }
    if (e2s_statistics_finish(&st, rec_len(m - 1), p.mcov_out, p.pval)) {
        std::cerr << "clust2snp: " << e2s_last_error(nullptr) << std::endl;
        return 3;
    }
    {
        uint64_t mx = 0;
        for (int i = 1; i <= E2S_MAX_C_LEN; ++i) mx = st.hist[i] * uint64_t(i) > mx ? st.hist[i] * uint64_t(i) : mx;
        uint64_t cumulative = 0;
        std::cout << "\nDistribution of base coverage: " << std::endl;
        std::cout << "\ncluster length\t# bases in a cluster with this length\t cumulative fraction (from 2m = " << 2 * p.mcov_out << ")" << std::endl;
        for (uint64_t i = 0; i <= st.max_len; ++i) {
            std::cout << i << "\t";
            if (mx) for (uint64_t j = 0; j < (100 * st.hist[i] * i) / mx; ++j) std::cout << "-";
            std::cout << "\t" << st.hist[i] * i;
            if (i >= uint64_t(2 * p.mcov_out)) {
                cumulative += st.hist[i] * i;
                std::cout << "\t" << double(cumulative) / double(st.n_bases);
            }
            std::cout << std::endl;
        }
        std::cout << "\nCluster sizes allowed: [" << p.mcov_out * 2 << "," << st.max_clust_length << "]" << std::endl;
    }

    // ---- find_events(): ref:clust2snp.cpp:788-872 ----
    std::cout << "(1/4) Filtering relevant clusters ... " << std::endl;
    std::vector<e2s_snp_counts> counts(static_cast<size_t>(G));
    auto analyse = [&](int g) {
        int r = 0;
        if (streamed[size_t(g)]) {
            const uint64_t lo = cuts[size_t(g)], hi = cuts[size_t(g) + 1];
            e2s_shard*& s_g = sh[size_t(g)];
            r = e2s_shard_create_chunked(ctx[size_t(g)], hi - lo, lo, idx.n, chunk, &s_g);
            if (!r) r = e2s_shard_set_layout(s_g, idx.x, idx.y, idx.z, idx.bcr ? 1 : 0);
            if (!r && idx.bcr) r = e2s_shard_host_gsa(s_g, idx.gsa.data, idx.y, idx.z);
            const uint64_t cp = r ? 0 : e2s_shard_chunk_positions(s_g);
            uint64_t r0 = rcut[size_t(g)];
            for (uint64_t clo = lo; !r && clo < hi; clo += cp) {
                const uint64_t cn = hi - clo < cp ? hi - clo : cp;
                r = e2s_chunk_begin(s_g, clo, cn);
                const uint64_t a = clo >= 176 ? clo - 176 : 0, b = clo + cn + E2S_MAX_C_LEN + 1 < idx.n ? clo + cn + E2S_MAX_C_LEN + 1 : idx.n;
                if (!r) r = idx.load(s_g, a, b - a, !idx.bcr);
                if (!r) r = e2s_shard_set_layout(s_g, idx.x, idx.y, idx.z, idx.bcr ? 1 : 0);
                uint64_t r1 = r0, top = rcut[size_t(g) + 1];  // first record of the shard that starts at or after the chunk's end
                while (r1 < top) {
                    const uint64_t mid = (r1 + top) / 2;
                    if (rec_start(mid) < clo + cn) r1 = mid + 1;
                    else top = mid;
                }
                if (!r) r = e2s_chunk_stage_clusters(s_g, cl.data + r0 * 10, r1 - r0, p.mcov_out, st.max_clust_length, nullptr);
                r0 = r1;
            }
            if (!r) r = e2s_chunked_clusters_finish(s_g);
            if (g == 0) host::stamp("index streamed, survivors captured");
        }
        if (!r) r = e2s_find_events(sh[size_t(g)], &p, st.max_clust_length, &counts[size_t(g)]);
        if (r) fail_of(g, r);
    };
    {
        std::vector<std::thread> th;
        for (int g = 0; g < G; ++g) th.emplace_back(analyse, g);
        for (auto& t : th) t.join();
    }
    uint64_t n_cand = 0;
    bool saw_n = false;
    for (int g = 0; g < G; ++g) {
        if (rc[size_t(g)]) {
            std::cerr << "clust2snp: GPU " << g << ": " << errs[size_t(g)] << std::endl;
            return 2;
        }
        n_cand += counts[size_t(g)].n_candidates;
        saw_n |= counts[size_t(g)].saw_n != 0;
    }
    host::stamp("find_events done");
    std::cout << " 100% done." << std::endl;
    std::cout << "Done. " << n_cand << " potential variants detected (some might be detected twice: on fw and rev strands)" << std::endl;
    if (saw_n)
        std::cerr << "clust2snp: warning: N bases met; the reference maps them through rand() (ref:include.hpp:273), here they count as A" << std::endl;
    if (n_cand == 0) {
        // the reference dereferences an empty vector here (ref:clust2snp.cpp:531) and dies without writing a .snp
        std::cerr << "clust2snp: no candidate variants; no .snp written" << std::endl;
        return 3;
    }
    std::cout << "(2/4) Extracting reads from fasta file ..." << std::endl << " 100% done." << std::endl;
    std::cout << "(3/4) Filtering " << n_cand << " candidates and computing consensus of left-contexts ... " << std::endl << " 100% done." << std::endl;
    std::cout << "(4/4) Computing edit distances and saving SNPs/indels to file ... " << std::endl;
    FILE* out = fopen(out_path.c_str(), "wb");
    if (!out) {
        std::cerr << "clust2snp: cannot write " << out_path << std::endl;
        return 2;
    }
    uint64_t next_id = 1;
    for (int g = 0; g < G; ++g) {
        uint64_t nv = 0;
        int r = e2s_events_fetch(sh[size_t(g)], nullptr, 0, &nv);
        std::vector<e2s_event> ev(static_cast<size_t>(nv));
        if (!r && nv) r = e2s_events_fetch(sh[size_t(g)], ev.data(), nv, &nv);
        char* text = nullptr;
        size_t len = 0;
        if (!r) r = e2s_events_format(ev.data(), nv, next_id, &p, &text, &len);
        if (r) {
            std::cerr << "clust2snp: " << e2s_last_error(nullptr) << std::endl;
            return 2;
        }
        if (len && fwrite(text, 1, len, out) != len) {
            std::cerr << "clust2snp: short write" << std::endl;
            return 2;
        }
        e2s_free(text);
        next_id += counts[size_t(g)].n_events;
    }
    fclose(out);
    host::stamp(".snp written");
    std::cout << " 100% done." << std::endl;
    std::cout << "Done. " << std::endl;
    host::quick_exit_unless_asked(0);  // (everything is on disk: skip the teardown of the CUDA contexts unless E2S_CLI_CLEAN_EXIT is set)
    for (int g = 0; g < G; ++g) {
        e2s_shard_destroy(sh[size_t(g)]);
        e2s_ctx_destroy(ctx[size_t(g)]);
    }
    host::stamp("contexts destroyed");
    return 0;
}

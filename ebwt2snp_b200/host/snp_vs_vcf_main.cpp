// snp_vs_vcf -- scores the SNPs of a .snp file against a ground-truth VCF (replaces the reference
// tool of the same name, ref:snp_vs_vcf.cpp:103-573, same options, same stdout).  No GPU work: a validation harness
// beside the hot path.  What it does, in the reference's terms:
//   * the reference FASTA of sample 1 is read contig by contig, upper-cased (ref:snp_vs_vcf.cpp:159-181);
//   * every single-base REF/ALT line of the VCF whose contig exists counts as a true SNP; those at least -l bases away
//     from both contig ends yield two "truth" entries -- forward and reverse complement -- holding the -l bases to the
//     right of the SNP and the reversed -l bases to its left (ref:snp_vs_vcf.cpp:216-283);
//   * entries whose neighbours in VCF order are closer than -k are non-isolated (ref:snp_vs_vcf.cpp:294-313);
//   * the truth entries are sorted by right context; every mismatching column of every >SNP pair of the calls file is
//     looked up (right context must prefix the truth's, alleles must agree in either order, reversed left context must
//     prefix the truth's), first with the higher path's contexts, then with the lower path's
//     (ref:snp_vs_vcf.cpp:349-497);
//   * TP = distinct true SNPs hit, FN = true SNPs - TP, FP = call columns without a hit, TN = (N - columns) - FN.
#include <getopt.h>
#include <stdint.h>

#include <algorithm>
#include <fstream>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>

namespace {

struct Truth {
    std::string right, left_rev;  // bases after the SNP; bases before it, nearest first
    char ref, alt;
    uint64_t id;                  // index of the SNP in the VCF (both strands share it)
    bool isolated;
    int pos;
};

const int K_NONISOLATED_DEFAULT = 31, READ_LENGTH_DEFAULT = 100;

void usage() {
    std::cout << "snp_vs_vcf [options]" << std::endl
              << "Options:" << std::endl
              << "-h          Print this help" << std::endl
              << "-v <arg>    VCF file with the ground-truth SNPs (REQUIRED)" << std::endl
              << "-c <arg>    Calls in KisSNP2 format (REQUIRED)" << std::endl
              << "-f <arg>    Reference fasta file of first sample (REQUIRED)" << std::endl
              << "-k <arg>    Value to define non-isolated SNPs (default: " << K_NONISOLATED_DEFAULT << ")" << std::endl
              << "-l <arg>    Max read length (default: " << READ_LENGTH_DEFAULT << ")" << std::endl;
    exit(0);
}

char complement(char c) {
    switch (c) {
        case 'A': return 'T';
        case 'C': return 'G';
        case 'G': return 'C';
        case 'T': return 'A';
    }
    return c;
}
std::string reversed(const std::string& s) { return std::string(s.rbegin(), s.rend()); }
std::string revcomp(const std::string& s) {
    std::string r = reversed(s);
    for (char& c : r) c = complement(c);
    return r;
}
bool prefixes(const std::string& a, const std::string& b) { return a.size() <= b.size() && b.compare(0, a.size(), a) == 0; }
bool single_base(const std::string& s) { return s == "A" || s == "C" || s == "G" || s == "T"; }
bool by_right(const Truth& a, const Truth& b) { return a.right < b.right; }

}  // namespace

int main(int argc, char** argv) {
    std::string vcf_path, calls_path, ref_path;
    int k_nonis = 0, rlength = 0;
    if (argc < 4) usage();
    int opt;
    while ((opt = getopt(argc, argv, "hv:c:f:l:k:")) != -1) {
        switch (opt) {
            case 'v': vcf_path = optarg; break;
            case 'c': calls_path = optarg; break;
            case 'f': ref_path = optarg; break;
            case 'l': rlength = atoi(optarg); break;
            case 'k': k_nonis = atoi(optarg); break;
            default: usage();
        }
    }
    if (rlength == 0) rlength = READ_LENGTH_DEFAULT;
    if (k_nonis == 0) k_nonis = K_NONISOLATED_DEFAULT;
    if (vcf_path.empty() || calls_path.empty() || ref_path.empty()) usage();

    // ---- reference contigs ----
    std::cout << "Loading reference ... " << std::flush;
    std::map<std::string, std::string> genome;
    std::vector<std::string> order;
    {
        std::ifstream in(ref_path);
        std::string line, name;
        while (!in.eof()) {  // (a final empty read after the last newline appends nothing)
            std::getline(in, line);
            if (!line.empty() && line[0] == '>') {
                name = line.substr(1);
                order.push_back(name);
                genome[name] = std::string();
            } else {
                for (char& c : line) c = char(toupper(c));
                genome[name].append(line);
            }
        }
    }
    std::cout << "done." << std::endl;
    uint64_t total_bases = 0;
    std::cout << "Contig\tlength" << std::endl;
    for (const std::string& name : order) {
        std::cout << name << "\t" << genome[name].length() << std::endl;
        total_bases += genome[name].length();
    }

    // ---- truth entries from the VCF ----
    std::cout << "Loading VCF ... " << std::flush;
    std::vector<Truth> truth;
    uint64_t n_true = 0, next_id = 0;
    int n_nonisolated = 0;
    {
        std::ifstream in(vcf_path);
        std::string line;
        while (!in.eof()) {
            std::getline(in, line);
            if (line.empty() || line[0] == '#') continue;
            std::istringstream fields(line);
            std::string chr, id, ref, alt;
            int pos;
            fields >> chr >> pos >> id >> ref >> alt;
            pos--;  // VCF coordinates start at 1
            if (!single_base(ref) || !single_base(alt)) continue;
            const std::string& seq = genome[chr];
            if (seq.empty()) {
                std::cout << "WARNING: chromosome " << chr << " not found. " << std::endl;
                continue;
            }
            n_true++;
            if (size_t(pos) >= seq.size())  // (the comparison is unsigned, as in the reference: pos = -1 warns too)
                std::cout << "WARNING: position " << pos << " larger than chromosome " << chr << "'s length " << seq.size() << std::endl;
            if (pos >= rlength && size_t(pos + rlength) < seq.size()) {
                const std::string after = seq.substr(size_t(pos) + 1, size_t(rlength)), before = seq.substr(size_t(pos - rlength), size_t(rlength));
                truth.push_back(Truth{after, reversed(before), ref[0], alt[0], next_id, true, pos});
                truth.push_back(Truth{revcomp(before), reversed(revcomp(after)), complement(ref[0]), complement(alt[0]), next_id, true, pos});
            }
            ++next_id;
        }
    }
    if (truth.empty())
        std::cout << "WARNING: no variants found. Check that chromosome names are the same in the fasta and vcf files. " << std::endl;
    // forward entries sit at even indices; the first and the last SNP stay "isolated"
    for (size_t i = 2; truth.size() > 1 && i < truth.size() - 2; i += 2) {
        const bool iso = truth[i].pos - truth[i - 2].pos >= k_nonis && truth[i + 2].pos - truth[i].pos >= k_nonis;
        truth[i].isolated = truth[i + 1].isolated = iso;
        if (!iso) n_nonisolated++;
    }
    std::cout << "done." << std::endl;
    std::cout << "Sorting VCF by context ... " << std::flush;
    std::sort(truth.begin(), truth.end(), by_right);
    std::cout << "done." << std::endl;

    // ---- calls ----
    std::cout << "Checking calls ... " << std::flush;
    uint64_t n_columns = 0, FP = 0;
    std::vector<char> hit(truth.size(), 0);
    auto lookup = [&](const std::string& right, const std::string& left_rev, char ref, char alt) {
        bool found = false;
        Truth probe{right, std::string(), 0, 0, 0, false, 0};
        size_t i = size_t(std::lower_bound(truth.begin(), truth.end(), probe, by_right) - truth.begin());
        for (; i < truth.size() && prefixes(right, truth[i].right); ++i) {
            const bool alleles = (truth[i].alt == alt && truth[i].ref == ref) || (truth[i].alt == ref && truth[i].ref == alt);
            if (alleles && prefixes(left_rev, truth[i].left_rev)) {
                found = true;
                hit[i] = 1;
            }
        }
        return found;
    };
    {
        std::ifstream in(calls_path);
        std::string header, skip;
        std::getline(in, header);
        while (!in.eof()) {
            if (header.substr(0, header.find('|')).compare(0, 4, ">SNP") == 0) {  // indels are not scored
                std::string dna_hi, header_lo, dna_lo;  // (fresh per record: a truncated record leaves them empty)
                std::getline(in, dna_hi);
                std::getline(in, header_lo);
                std::getline(in, dna_lo);
                if (dna_hi.length() != dna_lo.length()) {
                    std::cout << "Error: malformed SNP file. Two reads with different length in a SNP:\n";
                    std::cout << header << std::endl << dna_hi << std::endl << header_lo << std::endl << dna_lo << std::endl;
                    exit(1);
                }
                const size_t len = dna_hi.size();
                for (size_t back = 0; back < len; ++back) {  // every mismatching column, from the right end
                    const size_t col = len - back - 1;
                    if (dna_hi[col] == dna_lo[col]) continue;
                    n_columns++;
                    const char ref = dna_hi[col], alt = dna_lo[col];
                    if (!lookup(dna_hi.substr(col + 1), reversed(dna_hi.substr(0, col)), ref, alt) &&
                        !lookup(dna_lo.substr(col + 1), reversed(dna_lo.substr(0, col)), ref, alt))
                        FP++;
                }
            } else {
                std::getline(in, skip);
                std::getline(in, skip);
                std::getline(in, skip);
            }
            std::getline(in, header);
        }
    }
    std::cout << "done." << std::endl;

    std::set<int> found, found_nonisolated;
    for (size_t i = 0; i < truth.size(); ++i) {
        if (!hit[i]) continue;
        found.insert(int(truth[i].id));
        if (!truth[i].isolated) found_nonisolated.insert(int(truth[i].id));
    }
    const uint64_t TP = found.size(), FN = n_true - TP, TN = (total_bases - n_columns) - FN;  // TN: a lower bound
    std::cout << std::endl << "Non-isolated SNPs detected: " << found_nonisolated.size() << "/" << n_nonisolated << std::endl;
    std::cout << std::endl;
    std::cout << "TP = " << TP << std::endl;
    std::cout << "TN = " << TN << std::endl;
    std::cout << "FP = " << FP << std::endl;
    std::cout << "FN = " << FN << std::endl;
    std::cout << "sensitivity = TP/(TP+FN) = " << 100 * double(TP) / (double(TP) + double(FN)) << "%" << std::endl;
    std::cout << "specificity = TN/(TN+FP) = " << 100 * double(TN) / (double(TN) + double(FP)) << "%" << std::endl;
    std::cout << "precision   = TP/(TP+FP) = " << 100 * double(TP) / (double(TP) + double(FP)) << "%" << std::endl;
    return 0;
}

// Text post-processors of the .snp format: filter_snp (ref:filter_snp.cpp:24-117) and
// snp2fastq (ref:snp2fastq.cpp:28-150), as functions over streams so that the CLIs and the tests share them.
// Both are line-driven state machines over groups of four lines; their variables deliberately persist from one
// group to the next, because the reference's do (its "-i" mode of snp2fastq depends on it).
#pragma once

#include <cstdlib>
#include <istream>
#include <ostream>
#include <string>

namespace snptext {

// std::getline(istringstream, out, delim) semantics, including what happens to `out` at the end of the input:
// a field is assigned while the stream is good; once the end was hit, further reads leave `out` untouched.
struct Fields {
    const std::string& s;
    size_t pos = 0;
    bool at_end = false;
    explicit Fields(const std::string& str) : s(str) {}
    void next(std::string& out, char delim) {
        if (at_end) return;
        const size_t hit = s.find(delim, pos);
        if (hit == std::string::npos) {
            out = s.substr(pos);
            at_end = true;
        } else {
            out = s.substr(pos, hit - pos);
            pos = hit + 1;
        }
    }
};

struct Header {  // >{type}_{higher|lower}_path_{number}|P_1:{pos}_{event}|{coverage}|nb_pol_1
    std::string type, number, pos, event, coverage;
    // fields that are missing keep their previous value
    void parse(const std::string& line) {
        Fields bars(line);
        std::string tok;
        bars.next(tok, '|');
        tok = tok.empty() ? tok : tok.substr(1);  // drop '>'
        {   // the skipped fields land in the same scratch string the '|' fields use: when the line has no second
            // '|' field, the P_1 part is parsed from whatever was skipped last (the reference reuses one `token`)
            const std::string first = tok;
            Fields us(first);
            us.next(type, '_');
            us.next(tok, '_');
            us.next(tok, '_');
            us.next(number, '_');
        }
        bars.next(tok, '|');
        {
            const std::string second = tok;
            Fields colon(second);
            colon.next(tok, ':');
            colon.next(tok, ':');
            const std::string third = tok;
            Fields us(third);
            us.next(pos, '_');
            us.next(event, '_');
        }
        bars.next(coverage, '|');
    }
};

inline void third_bar_field(const std::string& line, std::string& out) {
    Fields bars(line);
    bars.next(out, '|');
    bars.next(out, '|');
    bars.next(out, '|');
}

// keep the 4-line groups whose two coverages are both >= M
inline void filter_snp(std::istream& in, int M, std::ostream& out) {
    std::string line, l1, l2, l3, cov0, cov1;
    Header h;
    unsigned idx = 0;
    while (std::getline(in, line)) {
        switch (idx % 4) {
            case 0: l1 = line; h.parse(line); cov0 = h.coverage; break;
            case 1: l2 = line; break;
            case 2: l3 = line; third_bar_field(line, cov1); break;
            default:
                if (atoi(cov0.c_str()) >= M && atoi(cov1.c_str()) >= M) out << l1 << '\n' << l2 << '\n' << l3 << '\n' << line << '\n';
        }
        ++idx;
    }
    out.flush();
}

// one FASTQ read per call: name = fields of the first individual + coverage and DNA of the second, DNA = the first
// individual's, fake qualities; swapped = the reference's -i
inline void snp2fastq(std::istream& in, bool swapped, std::ostream& out) {
    std::string line, header, dna, cov1;
    Header h;
    unsigned idx = 0;
    const unsigned head_a = swapped ? 2 : 0, dna_a = swapped ? 3 : 1, head_b = swapped ? 0 : 2, dna_b = swapped ? 1 : 3;
    while (std::getline(in, line)) {
        const unsigned r = idx % 4;
        if (r == head_a) {
            h.parse(line);
            header = h.type + "_" + h.number + "_" + h.pos + "_" + h.event + "_" + h.coverage;
        }
        if (r == dna_a) dna = line;
        if (r == head_b) {
            third_bar_field(line, cov1);
            header += "_" + cov1 + "_";
        }
        if (r == dna_b) {
            header += line;
            out << '@' << header << '\n' << dna << '\n' << "+\n" << std::string(dna.size(), 'I') << '\n';
        }
        ++idx;
    }
    out.flush();
}

}  // namespace snptext

// snp2fastq -- drop-in for ref:snp2fastq.cpp: `snp2fastq calls.snp [-i]` writes calls.snp.fastq.
#include <fstream>
#include <iostream>
#include <string>

#include "snp_text.hpp"

static void help() {
    std::cout << "snp2fastq calls.snp [-i]\n\n"
              << "Converts clust2snp's calls 'calls.snp' into a fastq file 'calls.snp.fastq': one read per call, the second\n"
              << "individual's DNA in the read's name, the first individual's DNA as the read. Base qualities are fake (all\n"
              << "maximum). With -i the individuals are switched." << std::endl;
    exit(0);  // ref:snp2fastq.cpp:26
}

int main(int argc, char** argv) {
    if (argc != 2 && argc != 3) help();
    bool swapped = false;
    if (argc == 3) {
        if (std::string(argv[2]) == "-i") swapped = true;
        else help();
    }
    const std::string infile = argv[1];
    std::ifstream is(infile);
    std::ofstream os(infile + ".fastq");
    snptext::snp2fastq(is, swapped, os);
    return 0;
}

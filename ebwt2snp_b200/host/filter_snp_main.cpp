// filter_snp -- drop-in for ref:filter_snp.cpp: `filter_snp calls.snp M` prints the calls with coverage >= M in both variants.
#include <fstream>
#include <iostream>

#include "snp_text.hpp"

int main(int argc, char** argv) {
    if (argc != 3) {
        std::cout << "filter_snp calls.snp M\n\n"
                  << "Input: a .snp file. Filters out only pairs with at least coverage M (in both variants). Output to stdout." << std::endl;
        return 0;  // ref:filter_snp.cpp:20
    }
    std::ifstream is(argv[1]);
    snptext::filter_snp(is, atoi(argv[2]), std::cout);
    return 0;
}

#!/bin/bash
# builds the timing tools that are not part of the library (they may use library sorts for comparison)
set -e
cd "$(dirname "$0")/.."
mkdir -p ebwt2snp_b200/bin
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -cudart static \
    -I include scripts/bench_build_egsa.cu -o ebwt2snp_b200/bin/bench_build_egsa

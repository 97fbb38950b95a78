// Host-side file handling shared by the two CLIs: index discovery (EGSA .gesa or the BCR triple, as
// egsa_stream does: ref:include.hpp:42-81), memory-mapped inputs, FASTA reader (multi-line records as
// ref:clust2snp.cpp:147-212 accepts them), .clusters reader/writer (ref:ebwt2clust.cpp:58-59) and the
// sharding of an eBWT over several GPUs of one box.
#pragma once

#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <string>
#include <vector>

#include "ebwt2snp_b200.h"

namespace host {

struct MappedFile {
    const uint8_t* data = nullptr;
    size_t size = 0;
    int fd = -1;
    bool open(const std::string& path) {
        fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) return false;
        size = size_t(st.st_size);
        if (size == 0) {
            data = nullptr;
            return true;
        }
        void* p = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (p == MAP_FAILED) return false;
        madvise(p, size, MADV_SEQUENTIAL);
        data = static_cast<const uint8_t*>(p);
        return true;
    }
    ~MappedFile() {
        if (data) munmap(const_cast<uint8_t*>(data), size);
        if (fd >= 0) close(fd);
    }
};

// The index next to the FASTA: X.gesa, else X.out + X.out.lcp + X.out.pairSA.
struct Index {
    bool egsa = false, bcr = false;
    MappedFile gesa, bwt, lcp, gsa;
    int x = 1, y = 4, z = 1;  // byte sizes of lcp, text (DA), suff (pos)
    uint64_t n = 0;

    bool open(const std::string& input, int x_, int y_, int z_) {
        x = x_; y = y_; z = z_;
        if (gesa.open(input + ".gesa")) {
            egsa = true;
            n = gesa.size / uint64_t(x + y + z + 1);
            return true;
        }
        if (lcp.open(input + ".out.lcp") && bwt.open(input + ".out") && gsa.open(input + ".out.pairSA")) {
            bcr = true;
            n = bwt.size;
            return true;
        }
        return false;
    }

    static uint32_t le(const uint8_t* p, int nb) {
        uint32_t v = 0;
        for (int b = 0; b < (nb < 4 ? nb : 4); ++b) v |= uint32_t(p[b]) << (8 * b);
        return v;
    }

    // Loads the global range [first, first+count) into the shard (which keeps what it needs).  need_gsa = false (ebwt2clust: the
    // cluster scan reads the LCP and the BWT only) leaves X.out.pairSA of a BCR triple on the disk.
    int load(e2s_shard* sh, uint64_t first, uint64_t count, bool need_gsa = true) const {
        if (count == 0) return E2S_OK;
        if (egsa) {
            // straight from the file through the library's pinned ring (E2S_CLI_MMAP=1: hand the mapping over instead -- pageable copies)
            static const bool use_map = getenv("E2S_CLI_MMAP") != nullptr;
            if (!use_map) return e2s_shard_load_gesa_fd(sh, gesa.fd, first, count, x, y, z);
            const size_t rs = size_t(x + y + z + 1);
            return e2s_shard_load_gesa(sh, gesa.data + first * rs, first, count, x, y, z);
        }
        // the BCR triple is already structure-of-arrays: the bytes go to the device as they are in the files and are widened there
        if (!need_gsa) return e2s_shard_load_lcp_bwt(sh, lcp.data + first * uint64_t(x), x, bwt.data + first, first, count);
        return e2s_shard_load_bcr(sh, lcp.data + first * uint64_t(x), x, bwt.data + first, gsa.data + first * uint64_t(y + z), y, z, first, count);
    }
};

// E2S_CLI_TIMING=1: wall-clock stamps of the CLI's phases on stderr (seconds since the first stamp)
inline void stamp(const char* what) {
    static const bool on = getenv("E2S_CLI_TIMING") != nullptr;
    if (!on) return;
    static const auto t0 = std::chrono::steady_clock::now();
    const double t = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "[e2s timing] %8.3f s  %s\n", t, what);
}

// A CLI whose outputs are on disk has nothing left to do: tearing the CUDA contexts down costs 50-200 ms of wall time per process.
inline void quick_exit_unless_asked(int code) {
    if (getenv("E2S_CLI_CLEAN_EXIT")) return;
    stamp("exit");
    fflush(stdout);
    fflush(stderr);
    _exit(code);
}

// contiguous, nearly equal shards; every shard has >= 2 positions
inline std::vector<uint64_t> shard_cuts(uint64_t n, int parts) {
    if (parts < 1) parts = 1;
    while (parts > 1 && n / uint64_t(parts) < 2) --parts;
    std::vector<uint64_t> cuts(size_t(parts) + 1);
    for (int g = 0; g <= parts; ++g) cuts[size_t(g)] = n / uint64_t(parts) * uint64_t(g) + (uint64_t(g) < n % uint64_t(parts) ? uint64_t(g) : n % uint64_t(parts));
    cuts[size_t(parts)] = n;
    return cuts;
}

inline int gpu_count_from_env() {
    const char* e = getenv("E2S_GPUS");
    int g = e ? atoi(e) : 1;
    return g < 1 ? 1 : g;
}

// FASTA: header lines start with '>', the sequence may span several lines.
struct Reads {
    std::vector<uint8_t> bases;
    std::vector<uint64_t> off;  // n_reads + 1
    bool load(const std::string& path) {
        MappedFile f;
        if (!f.open(path)) return false;
        bases.clear();
        off.clear();
        bases.reserve(f.size);
        const uint8_t* p = f.data;
        const uint8_t* end = f.data + f.size;
        bool first_line = true, in_read = false;
        while (p < end) {
            const uint8_t* nl = static_cast<const uint8_t*>(memchr(p, '\n', size_t(end - p)));
            const uint8_t* le = nl ? nl : end;
            if (first_line || (le > p && *p == '>')) {  // the first line is a header whatever it holds (ref:clust2snp.cpp:159)
                off.push_back(bases.size());
                in_read = true;
                first_line = false;
            } else if (in_read) {
                bases.insert(bases.end(), p, le);
            }
            p = nl ? nl + 1 : end;
        }
        off.push_back(bases.size());
        return true;
    }
    uint64_t n_reads() const { return off.empty() ? 0 : off.size() - 1; }
};

inline bool write_all(const std::string& path, const void* data, size_t bytes) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    bool ok = bytes == 0 || fwrite(data, 1, bytes, f) == bytes;
    return fclose(f) == 0 && ok;
}

}  // namespace host

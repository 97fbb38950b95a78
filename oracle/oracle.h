/*
 * oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded restatement of the reference's ebwt2clust + clust2snp algorithm
 * (nicolaprezza/ebwt2snp).  It exists to CHECK the CUDA path; nothing under ebwt2snp_b200/
 * may include, link or call it.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it.
 *
 * Parity status: PINNED.  The restatement is validated byte-for-byte against the reference
 * itself, compiled unmodified from /root/reference into oracle/_ref/ by oracle/Makefile
 * (tests/test_oracle_golden.py fuzzes it live where oracle/_ref exists; tests/golden/ holds outputs of oracle/_ref
 * committed together with the script that produced them, tests/golden/make_golden.py).
 * The reference tree ships no golden vectors of its own except the two distance() examples
 * in ref:clust2snp.cpp:250-251, which tests/test_oracle_golden.py checks.
 *
 * Known unpinned corner: an 'N'/'n' in the BWT is mapped by the reference through rand()%4
 * seeded with time() (ref:include.hpp:273, ref:clust2snp.cpp:971) -- not reproducible by
 * anything; the oracle counts it as 'A' and raises ORACLE_FLAG_SAW_N.
 */
#ifndef EBWT2SNP_ORACLE_H
#define EBWT2SNP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_MAX_C_LEN 150 /* ref:clust2snp.cpp:32 (max_clust_length_def; -M is unreachable) */
#define ORACLE_FLAG_SAW_N 1u
#define ORACLE_FLAG_BAD_READ_REF 2u /* a candidate referenced a read / offset outside the FASTA */

typedef struct {
    int k_left;          /* -L, default 31  ref:clust2snp.cpp:17   */
    int k_right;         /* -R, default 30  ref:clust2snp.cpp:20   */
    int mcov_out;        /* -m, default 5   ref:clust2snp.cpp:29   */
    int max_gap;         /* -g, default 10  ref:clust2snp.cpp:37   */
    int consensus_reads; /* -c, default 20  ref:clust2snp.cpp:40   */
    int max_err;         /* -e, default 2   ref:clust2snp.cpp:43   */
    int max_snvs;        /* always 3: the code tests max_snvs_def, -v is dead (ref:clust2snp.cpp:648) */
    double pval;         /* -p, default 0.99 ref:clust2snp.cpp:23  */
    uint64_t nr_reads1;  /* -n */
} oracle_params;

typedef struct {
    uint64_t n_written;     /* records appended to .clusters (length >= min_len) */
    uint32_t n_clust_out;   /* every closure, 32-bit counter as printed (ref:ebwt2clust.cpp:88,137) */
    uint32_t phantom_lcp;   /* the post-EOF record's lcp value P used */
} oracle_cluster_result;

typedef struct {
    uint64_t hist[ORACLE_MAX_C_LEN + 1];
    uint64_t n_clust;          /* incl. the double-counted last record (ref:clust2snp.cpp:889-909) */
    uint64_t n_bases;
    uint64_t max_len;          /* largest length <= 150 seen */
    int max_clust_length;      /* result of the pval loop (ref:clust2snp.cpp:938-946) */
} oracle_stats;

typedef struct {
    uint64_t n_candidates;  /* "Done. C potential variants detected" (ref:clust2snp.cpp:859) */
    uint64_t n_variants;    /* candidates with supp0>0 and supp1>0 (ref:clust2snp.cpp:595) */
    uint64_t n_events;      /* written to .snp (D <= 3)   (ref:clust2snp.cpp:648) */
    uint64_t n_analysed;    /* clusters passing the [2m, max] length filter */
    uint32_t flags;
} oracle_snp_result;

void oracle_default_params(oracle_params *p);

/* ebwt2clust: cluster_lm + append_entry, ref:ebwt2clust.cpp:54-139.  lcp/bwt: n records.
 * start_out/len_out: capacity cap (n+1 is always enough).  Returns 0, or -1 if cap too small. */
int oracle_cluster_lm(const uint32_t *lcp, const uint8_t *bwt, uint64_t n, uint32_t k, int min_len,
                      uint64_t *start_out, uint16_t *len_out, uint64_t cap, oracle_cluster_result *res);

int oracle_cluster_lm_x(const uint32_t *lcp, const uint8_t *bwt, uint64_t n, uint32_t k, int min_len, int lcp_bytes,
                        uint64_t *start_out, uint16_t *len_out, uint64_t cap, oracle_cluster_result *res);

/* the phantom record consumed after EOF, exposed for tests */
uint32_t oracle_phantom_field(const uint32_t *lcp, const uint8_t *bwt, uint64_t n);

/* clust2snp statistics(), ref:clust2snp.cpp:877-966.  m >= 1. */
int oracle_statistics(const uint64_t *start, const uint16_t *len, uint64_t m, int mcov_out, double pval,
                      oracle_stats *st);

/* clust2snp find_events() = find_variants + extract_variants + to_file,
 * ref:clust2snp.cpp:367-872.  reads: concatenated bases + offsets[R+1].
 * *snp_text is malloc'ed (oracle_free).  Returns 0, or -1 on an input the reference would crash on. */
int oracle_find_events(const uint32_t *lcp, const uint32_t *text, const uint32_t *suff, const uint8_t *bwt,
                       uint64_t n, const uint64_t *start, const uint16_t *len, uint64_t m,
                       const oracle_params *p, int max_clust_length,
                       const uint8_t *read_bases, const uint64_t *read_off, uint64_t n_reads,
                       char **snp_text, size_t *snp_len, oracle_snp_result *res);

/* same for index files with other field widths / the BCR triple (only the phantom record depends on them) */
int oracle_find_events_w(const uint32_t *lcp, const uint32_t *text, const uint32_t *suff, const uint8_t *bwt,
                         uint64_t n, const uint64_t *start, const uint16_t *len, uint64_t m,
                         const oracle_params *p, int max_clust_length,
                         const uint8_t *read_bases, const uint64_t *read_off, uint64_t n_reads,
                         int x, int y, int z, int bcr,
                         char **snp_text, size_t *snp_len, oracle_snp_result *res);
uint64_t oracle_phantom_slot(uint32_t lcp_last, uint32_t text_last, uint32_t suff_last, uint8_t bwt_last,
                             int x, int y, int z, int bcr);

/* distance(), ref:clust2snp.cpp:254-302 (equal-length strings) */
void oracle_distance(const char *a, const char *b, int len, int max_gap, int *D, int *gap);

/* EGSA of n_reads reads of read_len ACGT bases (row-major): n = n_reads * (read_len + 1) records.  Checks
 * e2s_build_egsa_dev; parity UNPINNED against the external egsa / BCR tools (absent; see oracle.c). */
int oracle_build_egsa(const uint8_t *reads, uint64_t n_reads, uint32_t read_len, uint32_t *lcp, uint32_t *text,
                      uint32_t *suff, uint8_t *bwt);

/* reads of any lengths: bases back to back, off = n_reads + 1 offsets; n = off[n_reads] + n_reads records */
int oracle_build_egsa_ragged(const uint8_t *bases, const uint64_t *off, uint64_t n_reads, uint32_t *lcp, uint32_t *text,
                             uint32_t *suff, uint8_t *bwt);

void oracle_free(void *p);

#ifdef __cplusplus
}
#endif
#endif

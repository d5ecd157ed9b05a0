/* oracle/port/oracle.h -- CPU restatement of BWBBLE's read-mapping hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by or executed
 * from the product (bwbble_b200/, include/).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it, and only as the checker.
 *
 * Parity pinning: this restatement is compared byte-for-byte (.aln streams, rank values,
 * interval lists, D arrays) against the UNMODIFIED reference compiled by `make -C oracle ref`
 * (tests/test_oracle_vs_ref.py, runs wherever /root/reference exists) and against the golden
 * files under tests/golden/ that the real reference produced (tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line (under /root/reference/mg-aligner) it follows.
 */
#ifndef BWBBLE_ORACLE_H
#define BWBBLE_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_ALPHABET 16
#define ORC_OCC_INTERVAL 128   /* bwt.h:14 */
#define ORC_SA_INTERVAL 32     /* bwt.h:16 */
#define ORC_READ_BATCH 0x40000 /* align.h:14 */
#define ORC_PATH_ALLOC 256     /* align.h:21 */
#define ORC_PRECALC_LEN 12     /* PRECALC_INTERVAL_LENGTH, align.h:31 */
#define ORC_NUM_PRECALC 16777216u /* NUM_PRECALC = 4^12, align.h:30 */

/* FM-index container; field meaning as bwt_t (bwt.h:19-40), file layout as store_bwt (bwt.c:66-82) */
typedef struct {
    uint64_t length, num_words, num_sa, num_occ, sa0_index;
    uint64_t C[ORC_ALPHABET + 1];
    uint32_t *bwt;   /* 8 symbols per word, first symbol in the top nibble (bwt.c:337-345) */
    uint64_t *O;     /* num_occ rows x 16, row k = counts in BWT[0..128k] inclusive (bwt.c:280-291) */
    uint64_t *SA;    /* every 32nd SA value, may be NULL */
} orc_bwt;

/* same 15 ints, same order, as aln_params_t (align.h:48-79) */
typedef struct {
    int max_diff, max_gapo, max_gape, max_entries;
    int mm_score, gapo_score, gape_score;
    int seed_length, max_diff_seed, max_best, no_indel_length;
    int matched_Ncontig, use_precalc, is_multiref, n_threads;
} orc_params;

/* workload counters; rank queries exclude the i==-1 / i==length-1 shortcuts */
typedef struct {
    uint64_t n_O, n_O_shortcut, n_Oalpha, n_Oalpha_shortcut;
    uint64_t pops, pushes, exact_tail_calls, max_heap, hits;
    uint64_t max_list;
} orc_stats;

typedef struct { int z; int w; } orc_dbound;   /* diff_lower_bound_t, inexact_match.h:11-14 */

void orc_default_params(orc_params *p);                       /* align.c:22-38 */

orc_bwt *orc_bwt_load(const char *path, int load_sa);         /* bwt.c:90-125 */
orc_bwt *orc_bwt_wrap(uint64_t length, uint64_t sa0_index, const uint64_t C[17],
                      const uint32_t *bwt, uint64_t num_words, const uint64_t *O, uint64_t num_occ,
                      const uint64_t *SA, uint64_t num_sa);   /* copies the arrays */
void orc_bwt_free(orc_bwt *b);

unsigned orc_B(const orc_bwt *b, uint64_t i);                                  /* bwt.c:337-345 */
uint64_t orc_O(const orc_bwt *b, unsigned c, uint64_t i);                      /* bwt.c:348-372 */
void orc_O_alphabet(const orc_bwt *b, uint64_t i, uint64_t occ[16], int inc);  /* bwt.c:374-438 */
void orc_O_actg(const orc_bwt *b, uint64_t i, uint64_t occ[5], int inc);       /* bwt.c:440-463 */
uint64_t orc_invPsi(const orc_bwt *b, uint64_t i);                             /* bwt.c:311-317 */
uint64_t orc_SA(const orc_bwt *b, uint64_t i);                                 /* bwt.c:320-329 */

/* interval list as a growable array (the reference uses a linked list, align.h:34-46) */
typedef struct { uint64_t L, U; } orc_intv;
typedef struct { orc_intv *v; int n, cap; } orc_list;
void orc_list_add(orc_list *l, uint64_t L, uint64_t U);                        /* align.c:93-110 */

/* exact_match.c:66-119 (multiref) / :196-222 (single-genome).  Returns 1 if any interval survives. */
int orc_exact_match_bounded(const orc_bwt *b, const uint8_t *read, int i, uint64_t l, uint64_t u,
                            const orc_params *p, orc_list *out);
/* inexact_match.c:171-254.  D has len+1 entries. */
/* -P: row index of a read (align.c:174-186) and one row of the .pre table (align.c:200-224) */
long orc_read2index(const uint8_t *read, int len);
int orc_precalc_entry(const orc_bwt *b, const orc_params *p, uint32_t index, orc_list *out);
void orc_calculate_d(const orc_bwt *b, const uint8_t *read, int len, orc_dbound *D, const orc_params *p);

/* Whole-batch driver, inexact_match.c:25-168.  seq = nt4 codes of the FORWARD reads, concatenated;
 * offsets has n_reads+1 entries.  Produces the byte stream alns2alnf_bin (align.c:345-382) would
 * append to the .aln file, in input order.  n_threads<=1 follows the serial driver (one D_seed for
 * the whole run, Q6), >1 the OpenMP driver (fresh D_seed per thread per 262144-read batch, static
 * chunks).  *aln is malloc'ed; free with orc_free. Returns 0, or <0 on unsupported params. */
int orc_align(const orc_bwt *b, const orc_params *p, const uint8_t *seq, const uint64_t *offsets,
              uint64_t n_reads, uint8_t **aln, uint64_t *aln_len, orc_stats *stats);
void orc_free(void *p);
/* test hook: bit sc of out[sc/64] = an entry with score sc was pushed since the last reset (scores < 1024) */
void orc_score_mask(uint64_t out[16], int reset);

#ifdef __cplusplus
}
#endif
#endif

/* oracle/port/rank.c -- FM-index container + rank primitives (TEST INFRASTRUCTURE ONLY).
 * Restates bwt.c:90-125 (load), :311-329 (invPsi/SA), :337-345 (B), :348-372 (O),
 * :374-438 + :689-781 (O_alphabet), :440-463 + :647-687 (O_actg_alphabet).
 * The reference counts zero nibbles of (word XOR c*0x11111111) through a 64 Ki LUT
 * (bwt.c:525-536, :575-600); here the same count is a popcount.  */
#include "oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

_Thread_local orc_stats orc_tls_stats;

static void die(const char *what, const char *path) {
    fprintf(stderr, "oracle: %s: %s\n", what, path);
    exit(1);
}

orc_bwt *orc_bwt_load(const char *path, int load_sa) {
    FILE *f = fopen(path, "rb");
    if (!f) die("cannot open .bwt", path);
    orc_bwt *b = calloc(1, sizeof *b);
    uint64_t hdr[5];
    if (fread(hdr, 8, 5, f) != 5) die("short .bwt header", path);
    b->length = hdr[0]; b->num_words = hdr[1]; b->num_sa = hdr[2]; b->num_occ = hdr[3]; b->sa0_index = hdr[4];
    if (fread(b->C, 8, 17, f) != 17) die("short .bwt C[]", path);
    b->bwt = malloc(b->num_words * 4 + 4);
    b->O = malloc(b->num_occ * 16 * 8);
    if (fread(b->bwt, 4, b->num_words, f) != b->num_words) die("short .bwt words", path);
    if (fread(b->O, 8, b->num_occ * 16, f) != b->num_occ * 16) die("short .bwt O[]", path);
    if (load_sa) {
        b->SA = malloc(b->num_sa * 8);
        if (fread(b->SA, 8, b->num_sa, f) != b->num_sa) die("short .bwt SA[]", path);
    }
    fclose(f);
    return b;
}

orc_bwt *orc_bwt_wrap(uint64_t length, uint64_t sa0_index, const uint64_t C[17], const uint32_t *bwt,
                      uint64_t num_words, const uint64_t *O, uint64_t num_occ, const uint64_t *SA,
                      uint64_t num_sa) {
    orc_bwt *b = calloc(1, sizeof *b);
    b->length = length; b->sa0_index = sa0_index; b->num_words = num_words; b->num_occ = num_occ;
    b->num_sa = num_sa;
    memcpy(b->C, C, sizeof b->C);
    b->bwt = malloc(num_words * 4 + 4); memcpy(b->bwt, bwt, num_words * 4);
    b->O = malloc(num_occ * 16 * 8); memcpy(b->O, O, num_occ * 16 * 8);
    if (SA) { b->SA = malloc(num_sa * 8); memcpy(b->SA, SA, num_sa * 8); }
    return b;
}

void orc_bwt_free(orc_bwt *b) {
    if (!b) return;
    free(b->bwt); free(b->O); free(b->SA); free(b);
}

void orc_free(void *p) { free(p); }

/* symbol at row i: nibble (i mod 8) from the top of word i/8 */
unsigned orc_B(const orc_bwt *b, uint64_t i) {
    return (b->bwt[i >> 3] >> (28 - 4 * (unsigned)(i & 7))) & 15u;
}

/* number of the first `n` nibbles (from the top) of w that equal c; n in 0..8 */
static inline unsigned nib_eq_prefix(uint32_t w, unsigned c, unsigned n) {
    uint32_t x = w ^ (0x11111111u * c);
    x |= x >> 1; x |= x >> 2;          /* low bit of each nibble = nibble != 0 */
    x = ~x & 0x11111111u;              /* low bit of each nibble = nibble == c */
    if (n < 8) x &= ~(0xFFFFFFFFu >> (4 * n));   /* keep the top n nibbles (n==0 -> none) */
    return (unsigned)__builtin_popcount(x);
}

/* #{p in (start, end] : bwt[p]==c}, start = 128*(end/128).  Same value as get_occ_count_opt
 * (bwt.c:575-600): whole words + masked partial word, minus the checkpoint symbol itself. */
static inline uint64_t block_count(const orc_bwt *b, unsigned c, uint64_t start, uint64_t end) {
    uint64_t w0 = start >> 3, nsym = end - start + 1, full = nsym >> 3, cnt = 0;
    for (uint64_t w = 0; w < full; w++) cnt += nib_eq_prefix(b->bwt[w0 + w], c, 8);
    unsigned rem = (unsigned)(nsym & 7);
    if (rem) cnt += nib_eq_prefix(b->bwt[w0 + full], c, rem);
    if ((b->bwt[w0] >> 28) == c) cnt--;
    return cnt;
}

uint64_t orc_O(const orc_bwt *b, unsigned c, uint64_t i) {
    if (i == b->length - 1) { orc_tls_stats.n_O_shortcut++; return b->C[c + 1] - b->C[c]; }
    if (i == (uint64_t)-1)  { orc_tls_stats.n_O_shortcut++; return 0; }
    orc_tls_stats.n_O++;
    uint64_t k = i / ORC_OCC_INTERVAL;
    uint64_t o = b->O[k * 16 + c];
    if (c != 0) return o + block_count(b, c, k * ORC_OCC_INTERVAL, i);
    /* code 0: walk the symbols, the sentinel row does not count (bwt.c:362-370) */
    for (uint64_t j = k * ORC_OCC_INTERVAL + 1; j <= i; j++)
        if (j != b->sa0_index && orc_B(b, j) == 0) o++;
    return o;
}

/* Q1: codes 5,9,11,13 get neither the in-block count nor the checkpoint value, but
 * still the "checkpoint symbol" decrement; all arithmetic is wrapping u64 (bwt.c:423-437,780). */
void orc_O_alphabet(const orc_bwt *b, uint64_t i, uint64_t occ[16], int inc) {
    if (i == b->length - 1) {
        orc_tls_stats.n_Oalpha_shortcut++;
        for (int j = 1; j < 16; j++) occ[j] = b->C[j + 1] + (uint64_t)inc;
        return;
    }
    if (i == (uint64_t)-1) {
        orc_tls_stats.n_Oalpha_shortcut++;
        for (int j = 1; j < 16; j++) occ[j] = b->C[j] + (uint64_t)inc;
        return;
    }
    orc_tls_stats.n_Oalpha++;
    uint64_t k = i / ORC_OCC_INTERVAL, start = k * ORC_OCC_INTERVAL;
    unsigned first = b->bwt[start >> 3] >> 28;
    for (unsigned j = 1; j < 16; j++) {
        int skipped = (j == 5 || j == 9 || j == 11 || j == 13);
        uint64_t v = occ[j];            /* caller pre-zeroes (inexact_match.c:377-378) */
        if (!skipped) {
            v += block_count(b, j, start, i) + b->O[k * 16 + j];
        } else if (first == j) {
            v -= 1;
        }
        occ[j] = v + b->C[j] + (uint64_t)inc;
    }
}

/* single-genome mode: occ[1..4] = A,G,C,T i.e. codes 15,3,7,1 (bwt.c:440-463) */
void orc_O_actg(const orc_bwt *b, uint64_t i, uint64_t occ[5], int inc) {
    static const unsigned code[5] = {0, 15, 3, 7, 1};
    if (i == b->length - 1) {
        orc_tls_stats.n_Oalpha_shortcut++;
        for (int j = 1; j < 5; j++) occ[j] = b->C[code[j] + 1] + (uint64_t)inc;
        return;
    }
    if (i == (uint64_t)-1) {
        orc_tls_stats.n_Oalpha_shortcut++;
        for (int j = 1; j < 5; j++) occ[j] = b->C[code[j]] + (uint64_t)inc;
        return;
    }
    orc_tls_stats.n_Oalpha++;
    uint64_t k = i / ORC_OCC_INTERVAL, start = k * ORC_OCC_INTERVAL;
    /* bwt.c:653-658: the decrement lands on the 16-wide scratch BEFORE the remap
     * occ[4]=occ[1]; occ[1]=occ[15]; occ[2]=occ[3]; occ[3]=occ[7]; with occ[] only 5 wide in the
     * caller's eyes the reference indexes a 16-wide array (inexact_match.c:377), so emulate it. */
    uint64_t scratch[16];
    memset(scratch, 0, sizeof scratch);
    for (int j = 0; j < 5; j++) scratch[j] = occ[j];
    unsigned first = b->bwt[start >> 3] >> 28;
    scratch[first]--;
    scratch[4] = scratch[1]; scratch[1] = scratch[15]; scratch[2] = scratch[3]; scratch[3] = scratch[7];
    for (int j = 1; j < 5; j++) {
        uint64_t c = block_count(b, code[j], start, i);
        if (first == code[j]) c++;   /* block_count already removed the checkpoint symbol; the
                                        reference removes it via scratch[first]-- instead */
        occ[j] = scratch[j] + c + b->C[code[j]] + b->O[k * 16 + code[j]] + (uint64_t)inc;
    }
}

uint64_t orc_invPsi(const orc_bwt *b, uint64_t i) {
    if (i == b->sa0_index) return 0;
    unsigned c = orc_B(b, i);
    return b->C[c] + orc_O(b, c, i);
}

uint64_t orc_SA(const orc_bwt *b, uint64_t i) {
    uint64_t j = 0;
    while (i % ORC_SA_INTERVAL) { i = orc_invPsi(b, i); j++; }
    return (b->SA[i / ORC_SA_INTERVAL] + j) % b->length;
}

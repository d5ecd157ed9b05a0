/* bwbble_shim.c -- drop-in replacement of the reference's inexact_match.o.
 *
 * Defines the two symbols align.o imports from inexact_match.o (mg-aligner/align.c:72-76,
 * inexact_match.h:39-40, bodies at inexact_match.c:25-168) on top of the C ABI of
 * libbwbble_b200.so, so that the reference's own main.o / align.o / bwt.o / io.o / is.o /
 * exact_match.o link unchanged into a `bwbble` whose CLI, index files, .aln format and SAM output
 * are untouched.  Observable behaviour kept: the .aln file is opened in append
 * mode (the caller removed it, align.c:48), one alns2alnf_bin record per read in input order, per
 * batch of READ_BATCH_SIZE reads the progress lines, and seq/rc/qual of every read are freed and
 * NULLed (inexact_match.c:71-79) so that free_reads() skips them.  Fatal errors print and exit(1)
 * like the reference's; there is no CPU fallback.
 *
 * The structs below mirror the reference's layouts (bwt.h:19-40, io.h:151-194, align.h:48-79) --
 * only the fields this file touches are named; offsets are asserted at compile time against the
 * values measured on the reference build (x86-64, gcc: SURVEY.md 8b).
 */
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bwbble_b200.h"

#define SHIM_READ_BATCH 0x800000         /* 32 x READ_BATCH_SIZE (align.h:14): the device needs millions of
                                           reads per launch to amortise the per-launch tail; the records
                                           and their order in the .aln file do not depend on the batching */

typedef struct {
    uint64_t length;                     /* bwt.h:21 */
    uint64_t num_words;
    uint32_t *bwt;
    uint64_t C[17];
    uint64_t *O;
    uint64_t num_occ;
    uint8_t occ_count_table[1 << 16];
    uint64_t *SA;
    uint64_t num_sa;
    uint64_t sa0_index;
} shim_bwt_t;

typedef struct {
    int len;                             /* io.h:153 */
    char name[257];
    char *seq;
    char *rc;
    char *qual;
    unsigned char rest[408 - 288];
} shim_read_t;

typedef struct {
    unsigned int count;                  /* io.h:188 */
    unsigned int max_len;
    shim_read_t *reads;
} shim_reads_t;

typedef struct shim_sa_intv_ {           /* align.h:34-38 */
    uint64_t L, U;
    struct shim_sa_intv_ *next_intv;
} shim_sa_intv_t;

typedef struct {                         /* align.h:42-46 */
    int size;
    shim_sa_intv_t *first_intv;
    shim_sa_intv_t *last_intv;
} shim_sa_intv_list_t;

#define SHIM_NUM_PRECALC 16777216        /* align.h:30 */

_Static_assert(sizeof(shim_sa_intv_t) == 24 && sizeof(shim_sa_intv_list_t) == 24 &&
               offsetof(shim_sa_intv_list_t, first_intv) == 8, "sa_intv_list_t layout");
_Static_assert(offsetof(shim_bwt_t, bwt) == 16 && offsetof(shim_bwt_t, C) == 24 && offsetof(shim_bwt_t, O) == 160 &&
               offsetof(shim_bwt_t, num_occ) == 168 && offsetof(shim_bwt_t, SA) == 65712 &&
               offsetof(shim_bwt_t, sa0_index) == 65728 && sizeof(shim_bwt_t) == 65736, "bwt_t layout");
_Static_assert(offsetof(shim_read_t, seq) == 264 && offsetof(shim_read_t, rc) == 272 &&
               offsetof(shim_read_t, qual) == 280 && sizeof(shim_read_t) == 408, "read_t layout");
_Static_assert(sizeof(bwb_params) == 60, "aln_params_t layout");

static void die(bwb_ctx *ctx, const char *what) {
    printf("bwbble_b200: %s: %s\n", what, bwb_last_error(ctx));
    exit(1);
}

/* -P: the table load_precalc_sa_intervals built (align.c:226-238), 4^12 linked lists -> sizes + flat (L,U) pairs */
static void upload_precalc(bwb_ctx *ctx, const shim_sa_intv_list_t *table, const bwb_params *params) {
    int32_t *sizes = (int32_t *)malloc((size_t)SHIM_NUM_PRECALC * sizeof(int32_t));
    uint64_t total = 0;
    if (!sizes) { printf("bwbble_b200: out of host memory\n"); exit(1); }
    for (size_t x = 0; x < SHIM_NUM_PRECALC; x++) { sizes[x] = table[x].size; total += (uint64_t)(table[x].size > 0 ? table[x].size : 0); }
    uint64_t *lu = (uint64_t *)malloc((size_t)(total + 1) * 16);
    if (!lu) { printf("bwbble_b200: out of host memory\n"); exit(1); }
    uint64_t w = 0;
    for (size_t x = 0; x < SHIM_NUM_PRECALC; x++) {
        const shim_sa_intv_t *iv = table[x].size > 0 ? table[x].first_intv : NULL;
        for (int k = 0; k < table[x].size; k++, iv = iv->next_intv) { lu[2 * w] = iv->L; lu[2 * w + 1] = iv->U; w++; }
    }
    if (bwb_precalc_upload(ctx, sizes, lu, total, params->is_multiref)) die(ctx, "seed table upload failed");
    free(sizes);
    free(lu);
}

static int run(shim_bwt_t *BWT, shim_reads_t *reads, const shim_sa_intv_list_t *precalc, bwb_params *params, char *alnFname,
               const char *banner) {
    printf("%s", banner);
    FILE *alnFile = fopen(alnFname, "a+b");
    if (alnFile == NULL) {
        printf("align_reads_inexact: Cannot open ALN file: %s!\n", alnFname);
        perror(alnFname);
        exit(1);
    }
    fclose(alnFile);

    int ndev = 0;
    const char *env = getenv("BWBBLE_GPUS");
    if (env) ndev = atoi(env);
    int devs[64];
    bwb_ctx *ctx = NULL;
    if (ndev > 0) {
        if (ndev > 64) ndev = 64;
        for (int i = 0; i < ndev; i++) devs[i] = i;
        ctx = bwb_create(devs, ndev);
    } else {
        ctx = bwb_create(NULL, 0);
    }
    if (!ctx) die(NULL, "cannot create the device context");
    if (bwb_index_upload(ctx, BWT->length, BWT->sa0_index, BWT->C, BWT->bwt, BWT->num_words, BWT->O, BWT->num_occ))
        die(ctx, "index upload failed");
    if (params->use_precalc) {
        if (!precalc) { printf("bwbble_b200: -P without a pre-calculated interval table\n"); exit(1); }
        upload_precalc(ctx, precalc, params);
    }

    /* a read no longer than the seed consults the D_seed of the last longer read before it.  The serial
     * driver's chain runs through all reads (inexact_match.c:36), so it is carried across the launches below; the OpenMP
     * driver's restarts per thread chunk of every 262144-read batch (:115-121) -- SHIM_READ_BATCH is a multiple. */
    bwb_set_option(ctx, "seed_carry", params->n_threads > 1 ? 0 : 1);
    uint8_t *seq = NULL;
    uint64_t *off = NULL;
    size_t seq_cap = 0;
    off = (uint64_t *)malloc((SHIM_READ_BATCH + 1) * sizeof(uint64_t));
    unsigned int done = 0;
    while (done < reads->count) {
        const unsigned int bs = reads->count - done > SHIM_READ_BATCH ? SHIM_READ_BATCH : reads->count - done;
        size_t total = 0;
        for (unsigned int i = 0; i < bs; i++) total += (size_t)reads->reads[done + i].len;
        if (total + 1 > seq_cap) {
            seq_cap = total + 1;
            seq = (uint8_t *)realloc(seq, seq_cap);
        }
        if (!seq || !off) { printf("bwbble_b200: out of host memory\n"); exit(1); }
        size_t p = 0;
        for (unsigned int i = 0; i < bs; i++) {
            shim_read_t *r = &reads->reads[done + i];
            off[i] = p;
            memcpy(seq + p, r->seq, (size_t)r->len);
            p += (size_t)r->len;
        }
        off[bs] = p;
        bwb_results *res = NULL;
        if (bwb_align(ctx, params, seq, off, bs, &res)) die(ctx, "alignment failed");
        printf("Processed %d reads. Inexact matching time: n/a (device).", (int)(done + bs));
        if (bwb_results_write_aln(res, alnFname, 1)) die(ctx, "cannot write the ALN file");
        bwb_results_free(res);
        for (unsigned int i = 0; i < bs; i++) {
            shim_read_t *r = &reads->reads[done + i];
            free(r->seq); free(r->rc); free(r->qual);
            r->seq = r->rc = r->qual = NULL;
        }
        printf("Storing results time: n/a\n");
        done += bs;
    }
    free(seq);
    free(off);
    bwb_destroy(ctx);
    return 0;
}

int align_reads_inexact(void *BWT, void *reads, void *precalc_sa_intervals_table, void *params, char *alnFname) {
    bwb_params p = *(const bwb_params *)params;
    p.n_threads = 1;                         /* this entry point IS the serial driver, whatever -t says */
    return run((shim_bwt_t *)BWT, (shim_reads_t *)reads, (const shim_sa_intv_list_t *)precalc_sa_intervals_table,
               &p, alnFname, "BWBBLE Inexact Alignment...\n");
}

int align_reads_inexact_parallel(void *BWT, void *reads, void *precalc_sa_intervals_table, void *params, char *alnFname) {
    return run((shim_bwt_t *)BWT, (shim_reads_t *)reads, (const shim_sa_intv_list_t *)precalc_sa_intervals_table,
               (bwb_params *)params, alnFname, "BWT-SNP Inexact Alignment...\n");
}

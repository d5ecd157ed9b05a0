/* bwbble_b200.h -- C ABI of the B200-native BWBBLE read-mapping hot path.
 *
 * Drop-in boundary: the reference's only callers of the hot path are
 *     align_reads()            mg-aligner/align.c:72-76
 * which calls
 *     int align_reads_inexact         (bwt_t*, reads_t*, sa_intv_list_t*, aln_params_t*, char*)
 *     int align_reads_inexact_parallel(bwt_t*, reads_t*, sa_intv_list_t*, aln_params_t*, char*)
 *                              mg-aligner/inexact_match.h:39-40, inexact_match.c:25-168
 * bwbble_b200/csrc/shim/bwbble_shim.c defines those two symbols on top of this ABI, so the
 * reference's own main.o/align.o/bwt.o/io.o link against it unchanged.
 *
 * Plain pointers and sizes only; no torch/CUDA types.  Every function returns 0 on success or a
 * negative bwb_status; the text of the last failure is bwb_last_error().  Nothing here calls
 * exit(), and nothing falls back to the CPU: without a usable CUDA device bwb_create() fails.
 */
#ifndef BWBBLE_B200_H
#define BWBBLE_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BWB_ABI_VERSION 1

typedef enum {
    BWB_OK = 0,
    BWB_ERR_ARG = -1,          /* bad argument / unsupported parameter combination */
    BWB_ERR_CUDA = -2,         /* CUDA runtime failure (no device, OOM, launch error) */
    BWB_ERR_IO = -3,           /* file could not be read / written */
    BWB_ERR_NO_INDEX = -4,     /* bwb_align before bwb_index_upload */
    BWB_ERR_CAPACITY = -5,     /* a device pool (heap chunks, interval lists, hits) overflowed */
    BWB_ERR_UNSUPPORTED = -6   /* parameter combination outside the device path (max_gapo > 4; > 128 reachable scores) */
} bwb_status;

/* Mirror of aln_params_t (mg-aligner/align.h:48-79): the same 15 ints in the same order, so a
 * reference aln_params_t* can be passed as a bwb_params*. */
typedef struct {
    int32_t max_diff;         /* -n */
    int32_t max_gapo;         /* -o */
    int32_t max_gape;         /* -e */
    int32_t max_entries;      /* -m */
    int32_t mm_score;         /* -M */
    int32_t gapo_score;       /* -O */
    int32_t gape_score;       /* -E */
    int32_t seed_length;      /* -l */
    int32_t max_diff_seed;    /* -k */
    int32_t max_best;
    int32_t no_indel_length;
    int32_t matched_Ncontig;  /* unused by the reference's hot path */
    int32_t use_precalc;      /* -P: BWB_ERR_UNSUPPORTED */
    int32_t is_multiref;      /* 0 = -S single-genome mode */
    int32_t n_threads;        /* -t: only selects which driver's D_seed semantics short reads get (see bwb_align) */
} bwb_params;

/* set_default_aln_params, align.c:22-38 */
void bwb_default_params(bwb_params *p);

/* The score buckets the device search keeps for these parameters.  The reference's heap has
 * nb = (n+1)*M + (o+1)*O + (e+1)*E buckets, one per score (heap_init, inexact_match.c:510-528); an
 * entry can only score m*M + o*O + e*E with m + o + e <= max_diff, o <= max_gapo, e <= max_gape
 * (e > 0 only after an opening), and only those scores get a bucket on the device, in ascending
 * order.  Returns their number (bwb_align needs it <= 128) or a negative bwb_status; when
 * bucket_of is not NULL, bucket_of[s] for s < min(cap, nb) is the bucket of score s or 0xff.
 * Host only: needs no device. */
int bwb_score_buckets(const bwb_params *p, uint8_t *bucket_of, int cap);

/* One gap run of an alignment path.  A path (aln_entry_t.aln_path, align.h:118) is all STATE_M
 * except for at most num_gapo runs of STATE_I(1)/STATE_D(2); `start` is the index of the run's
 * first element in search order (path[0] = first step = last read base). */
typedef struct {
    uint8_t start, len, state, pad;
} bwb_gap_run;

#define BWB_MAX_GAP_RUNS 4     /* max_gapo above this is BWB_ERR_UNSUPPORTED */

/* One alignment hit = one aln_t (align.h:81-91) without num_snps (never serialised). */
typedef struct {
    uint64_t L, U;             /* SA interval */
    int32_t score;
    uint8_t num_mm, num_gapo, num_gape, aln_length;
    uint8_t n_runs, pad[3];
    uint32_t read_id;          /* index of the read inside the bwb_align call */
    bwb_gap_run runs[BWB_MAX_GAP_RUNS];
} bwb_hit;                     /* 48 bytes */

typedef struct bwb_ctx bwb_ctx;
typedef struct bwb_results bwb_results;
typedef struct bwb_reads bwb_reads;

/* ---- context ------------------------------------------------------------------------ */
/* devices==NULL / ndev<=0: CUDA device 0 only.  Reads of one bwb_align call are sharded over the
 * context's devices in contiguous ranges (the reference's static OpenMP chunks,
 * inexact_match.c:115-116); the index is replicated.  Returns NULL on failure
 * (bwb_last_error(NULL) has the reason). */
bwb_ctx *bwb_create(const int *devices, int ndev);
void bwb_destroy(bwb_ctx *ctx);
const char *bwb_last_error(const bwb_ctx *ctx);
int bwb_device_count(const bwb_ctx *ctx);

/* Options (before bwb_index_upload / first bwb_align):
 *   "heap_pool_mb"     device MB for the search arena (heap entries, tail lists, hits), per device
 *                      (default 0 = auto: half of the free device memory, at most 64 GB)
 *   "list_cap"         max SA intervals per list per read-slot (default 4096)
 *   "hits_per_read"    staging capacity for hits of one read (default 512)
 *   "warps_per_block"  search kernel block shape (default 8)
 *   "blocks_per_sm"    persistent blocks per SM (default: occupancy query)
 *   "engine"           1 = warp-per-read kernel k_align (A/B only; 2 = back to the default 8-lane groups)
 *   "kmer_table"       2 = do not use the 10-mer table of calculate_d's top of tree (A/B, tests; default on)
 *   "force_wide"       1 = 64-bit coordinates / 32-byte heap entries even on a small index (tests)  */
int bwb_set_option(bwb_ctx *ctx, const char *key, long long value);
/* Launch on a caller-owned cudaStream_t (e.g. torch's current stream) instead of the context's own. */
int bwb_set_stream(bwb_ctx *ctx, int dev_slot, void *cuda_stream);

/* ---- index (replaces load_bwt's in-memory result, bwt.c:90-125, bwt.h:19-40) ---------- */
/* bwt: 8 symbols per word, first symbol in the top nibble; O: num_occ rows of 16 inclusive counts.
 * Re-laid-out on the device into 128-byte blocks (K0).  Host arrays are not referenced afterwards. */
int bwb_index_upload(bwb_ctx *ctx, uint64_t length, uint64_t sa0_index, const uint64_t C[17],
                     const uint32_t *bwt, uint64_t num_words, const uint64_t *O, uint64_t num_occ);
/* Read "<path>" in store_bwt's layout (bwt.c:66-82) and upload it (without the sampled SA). */
int bwb_index_load_file(bwb_ctx *ctx, const char *bwt_path);
/* Same, and upload the sampled SA stored at the end of the file as well (load_bwt(path, 1)). */
int bwb_index_load_file_sa(bwb_ctx *ctx, const char *bwt_path);
/* Download the device blocks of device slot 0 (tests): out must hold bwb_index_num_blocks()*128 bytes. */
uint64_t bwb_index_num_blocks(const bwb_ctx *ctx);
int bwb_index_download_blocks(bwb_ctx *ctx, void *out);

/* ---- rank primitives (K1; parity tests + the Occ-gather micro-benchmark) -------------- */
/* out[q] = O(code[q], pos[q])  (bwt.c:348-372), code in 1..15 */
int bwb_occ(bwb_ctx *ctx, const uint8_t *code, const uint64_t *pos, uint64_t n, uint64_t *out);
/* out[q*16+j] = occ[j] of O_alphabet(pos[q], inc) with occ pre-zeroed (bwt.c:374-438), j=1..15 */
int bwb_occ_alphabet(bwb_ctx *ctx, const uint64_t *pos, uint64_t n, int inc, uint64_t *out);
/* n uniform random rank queries (device-generated positions, whole index), `iters` timed launches.
 * mode 0: one code per query (16 B counters+planes by one lane); mode 1: all 15 codes (16 lanes).
 * Returns average ms per launch and a checksum (so the work cannot be elided). */
int bwb_occ_bench(bwb_ctx *ctx, uint64_t n, uint64_t seed, int mode, int iters, float *ms_per_launch,
                  uint64_t *checksum);

/* ---- -P seed table -------------------------------------------------------- */
/* `bwbble align -P` starts every search from the exact-match intervals of the read's last 12
 * (reverse-complement) bases, looked up in a 4^12-row table that precalc_sa_intervals()
 * (align.c:200-224) computes with exact_match() and stores as <fasta>.pre.  A read with an N among
 * those bases, or shorter than 12, gets no alignments (inexact_match.c:50-57; the reference reads
 * out of bounds for the latter).  bwb_params.use_precalc = 1 needs one of these first: */
/* K0c: compute the table on the device (is_multiref = 0 for the -S flavour) */
int bwb_precalc_build(bwb_ctx *ctx, int is_multiref);
/* the table as the reference holds it after load_precalc_sa_intervals (align.c:226-238):
 * sizes[4^12] and the rows' (L,U) pairs concatenated in row order */
int bwb_precalc_upload(bwb_ctx *ctx, const int32_t *sizes, const uint64_t *intervals_LU, uint64_t n_intervals,
                       int is_multiref);
/* read / write a .pre file, byte-identical to the reference's (align.c:144-172) */
int bwb_precalc_load_file(bwb_ctx *ctx, const char *pre_path, int is_multiref);
int bwb_precalc_write(bwb_ctx *ctx, const char *pre_path);
uint64_t bwb_precalc_num_intervals(const bwb_ctx *ctx);
/* one row (tests): *n = its size, up to cap (L,U) pairs copied */
int bwb_precalc_row(const bwb_ctx *ctx, uint32_t row, uint64_t *intervals_LU, uint32_t cap, uint32_t *n);

/* ---- search ---------------------------------------------------------------------------- */
/* seq: nt4 codes (A0 G1 C2 T3, anything else = N) of the FORWARD reads, concatenated;
 * offsets: n_reads+1 entries.  Reads longer than 255 are rejected (8-bit positions, align.h:104). */

/* exact_match_bounded(read, len-1, 0, length-1) per read (exact_match.c:58-60,66-119): interval
 * lists in reference order.  counts[n_reads]; intervals returned flat as (L,U) pairs. */
int bwb_exact_match(bwb_ctx *ctx, const uint8_t *seq, const uint64_t *offsets, uint64_t n_reads,
                    uint32_t *counts, uint64_t **intervals_LU, uint64_t *n_intervals);
/* calculate_d (inexact_match.c:171-254) per read on read[0..dlen) where dlen = use_len>0 ?
 * min(use_len,len) : len.  out holds, per read, (dlen+1) pairs {num_diff, sa_intv_width} packed at
 * 2*(offsets[r]+r) int32s. */
int bwb_calculate_d(bwb_ctx *ctx, const uint8_t *seq, const uint64_t *offsets, uint64_t n_reads,
                    int use_len, int32_t *out);

/* K3 of the production engine: for every read D = calculate_d(read, len) and, if seed_len > 0,
 * D_seed = calculate_d(read, seed_len) (inexact_match.c:61-64).  d_main: (len+1) pairs per read at
 * 2*(offsets[r]-offsets[0]+r) int32s; d_seed: (seed_len+1) pairs per read at 2*r*(seed_len+1); all
 * zero for reads with len <= seed_len (the reference leaves a stale array there). */
int bwb_lower_bounds(bwb_ctx *ctx, const uint8_t *seq, const uint64_t *offsets, uint64_t n_reads, int seed_len,
                     int32_t *d_main, int32_t *d_seed);

/* The hot path: calculate_d + inexact_match for every read (inexact_match.c:25-168), results in
 * input order.  Host buffers in, host-readable results out (H2D/D2H inside).
 *
 * One call = one call of the reference's driver on these reads, including its one order-dependent
 * behaviour: a read no longer than seed_length gets no D_seed of its own and consults
 * the bounds of the last longer read its driver thread aligned before it (zeros if none).
 * params->n_threads <= 1 follows align_reads_inexact (one array for the whole call,
 * inexact_match.c:36,62-64); n_threads > 1 follows align_reads_inexact_parallel (a fresh array per
 * thread and 262144-read batch, static chunks, :115-121,141-143).  Nothing else depends on
 * n_threads; reads longer than the seed are unaffected. */
int bwb_align(bwb_ctx *ctx, const bwb_params *params, const uint8_t *seq, const uint64_t *offsets,
              uint64_t n_reads, bwb_results **out);

/* A run split over several bwb_align calls (serial driver only): the last read longer than the seed
 * that the earlier calls saw, whose D_seed short reads at the start of the next call inherit.
 * bwb_set_option(ctx, "seed_carry", 1) starts a run in which every call updates it by itself
 * (what bwb_align_fastq and the drop-in shim do); this function sets it explicitly (len = 0: none)
 * -- for callers that shard a read set over processes (bwbble_b200/dist.py).  Host only. */
int bwb_set_seed_carry(bwb_ctx *ctx, const uint8_t *read_seq, int len);

/* Which read's D_seed does read r consult in one bwb_align call with these parameters (see above)?
 * donor_of[r] = r (its own: the read is longer than the seed), the index of an earlier read, -1 (none:
 * zeros; also for reads -P skips) or -2 (the run's carry, when have_carry).  Host only: the plan the
 * device path follows, exposed for tests and for callers that shard a run themselves. */
int bwb_seed_donor_plan(const bwb_params *p, const uint8_t *seq, const uint64_t *offsets, uint64_t n_reads,
                        int have_carry, int64_t *donor_of);

/* Device-resident variant: upload once, align many times (bench `value`; no PCIe in the loop). */
int bwb_reads_upload(bwb_ctx *ctx, const uint8_t *seq, const uint64_t *offsets, uint64_t n_reads,
                     bwb_reads **out);
void bwb_reads_free(bwb_reads *r);   /* before bwb_destroy() of the context the reads were uploaded to */
/* Runs the kernels on the context's stream(s); results stay on the device until
 * bwb_results_fetch() (fetch==0: no bulk D2H, only the 256-byte status block is read back).
 * Lifetime: un-fetched results live in per-context device buffers that the NEXT bwb_align /
 * bwb_align_resident on the same context overwrites; bwb_results_fetch() on results of an older
 * launch fails with BWB_ERR_ARG instead of returning another batch's hits. */
int bwb_align_resident(bwb_ctx *ctx, const bwb_params *params, const bwb_reads *reads, int fetch,
                       bwb_results **out);
int bwb_results_fetch(bwb_results *r);

/* ---- results --------------------------------------------------------------------------- */
uint64_t bwb_results_num_reads(const bwb_results *r);
uint64_t bwb_results_num_hits(const bwb_results *r);
const uint32_t *bwb_results_counts(const bwb_results *r);   /* hits per read, input order */
const bwb_hit *bwb_results_hits(const bwb_results *r);      /* flat, grouped by read, input order */
/* kernel-side counters of the last call: [0] pops [1] pushes [2] exact-tail calls [3] block loads
 * (physical 128-B rank gathers) [4] max heap entries of any read [5] max interval-list length
 * [6], [7] reads deferred to K4's 2nd / 3rd pass (their heap outgrew the arena share of a lane; those passes run
 * with 8x / 64x the arena per read) */
int bwb_results_counters(const bwb_results *r, uint64_t out[8]);
/* Duration of the search kernel (K4) of the call that produced r, in ms: CUDA events on the launch
 * stream, max over the context's devices. */
double bwb_results_kernel_ms(const bwb_results *r);
/* Same for the lower-bound kernel K3 (0 for the fused warp engine). */
double bwb_results_k3_ms(const bwb_results *r);
/* Serialise exactly as alns2alnf_bin (align.c:345-382) does for each read in order. */
int bwb_results_aln_bytes(const bwb_results *r, uint8_t **buf, uint64_t *len);   /* free with bwb_free */
int bwb_results_write_aln(const bwb_results *r, const char *path, int append);
void bwb_results_free(bwb_results *r);
void bwb_free(void *p);

/* ---- SA locate + SAM (the step right after the path; bwbble aln2sam) ------------ */
/* Sampled suffix array of the index (every 32nd value, bwt.h:16,34; the tail of the .bwt file).  Once
 * uploaded, every bwb_align call also runs K6: per read the text position of its first hit,
 * SA(L) = (SA[i/32] + j) mod length after j applications of invPsi (bwt.c:311-329), and the
 * aln_top1_count / aln_top2_count sums of eval_aln (align.c:760-812). */
int bwb_sa_upload(bwb_ctx *ctx, const uint64_t *SA, uint64_t num_sa);
typedef struct {
    uint64_t ref_pos;          /* SA(L of hit 0); ~0 for a read without hits */
    int32_t top1, top2;        /* wrapped int sums of interval widths: score <= best / score > best */
} bwb_loc;
const bwb_loc *bwb_results_locations(const bwb_results *r);     /* NULL if no SA was uploaded */
/* Write what `bwbble aln2sam -n max_mm` writes for these reads (alns2sam, align.c:494-652): @SQ/@PG
 * header (if write_header), one line per read; strand/POS from the located position, MAPQ by mapq()
 * (align.c:738-746), CIGAR from the edit path.  ann_path = <fasta>.ann; names/quals = n_reads C strings
 * (quals may be NULL), seq/offsets as given to bwb_align. */
int bwb_results_write_sam(const bwb_results *r, const char *ann_path, const char *const *names, const uint8_t *seq,
                          const uint64_t *offsets, const char *const *quals, uint64_t index_length, int max_mm,
                          const char *sam_path, int write_header, int append);

/* ---- streaming ingest (replaces the all-resident fastq2reads, io.c:410-515) ------- */
/* Parse `fastq_path` in batches of `batch_reads` (0 = 4 Mi) reads, align every batch and append its
 * records to `aln_path` (binary .aln, removed first like align.c:48) and/or `sam_path` (needs the
 * sampled SA uploaded, `ann_path` = <fasta>.ann, index_length, max_mm as for bwb_results_write_sam).
 * Three threads: the parser (memchr over 4 MB blocks) works on batch k+1 and the writer on batch k-1 while
 * batch k is on the device.  Returns the number of reads, or a negative bwb_status -- in which case the
 * output written so far is renamed to <path>.partial rather than left under the final name. */
/* Host-only FASTQ reader with fastq2reads' record grammar (io.c:410-515): nt4 codes of all reads concatenated,
 * n_reads+1 offsets; optionally the names / quality strings, each followed by a NUL.  malloc'ed: bwb_free(). */
int bwb_fastq_parse(const char *fastq_path, uint8_t **seq, uint64_t **offsets, uint64_t *n_reads,
                    char **names, uint64_t *names_bytes, char **quals, uint64_t *quals_bytes);
long long bwb_align_fastq(bwb_ctx *ctx, const bwb_params *params, const char *fastq_path, const char *aln_path,
                          const char *sam_path, const char *ann_path, uint64_t index_length, int max_mm,
                          uint64_t batch_reads);

/* ---- host-side index construction (bwbble index, bwt.c:29-63; SURVEY 8f "next") ----------- */
/* Builds <fasta>.bwt and <fasta>.ann byte-identical to the reference's `bwbble index <fasta>`
 * (io.c:190-321 fasta2ref, bwt.c:161-218 construct_bwt, is.c:214-243) with an own SA-IS. */
int bwb_index_build(const char *fasta_path, int write_ref_file);
/* K7: the same files with the suffix sort (prefix doubling) and the BWT / checkpoint / SA-sample passes on
 * device 0 of ctx (bwt.c:161-218 on the GPU); indexes of < 2^31-16 rows.  *sort_rounds (optional) = number of
 * doubling rounds the suffix sort took. */
int bwb_index_build_device(bwb_ctx *ctx, const char *fasta_path, int write_ref_file, int *sort_rounds);

#ifdef __cplusplus
}
#endif
#endif

/* oracle/port/main.c -- command line around the CPU restatement (TEST INFRASTRUCTURE ONLY).
 *   bwbble_oracle align [-M -O -E -n -k -o -e -l -m -t -S] <ref.fasta> <reads.fastq> <out.aln>
 * Same options and defaults as the reference CLI (main.c:91-120); reads <ref.fasta>.bwt, parses
 * FASTQ the way fastq2reads does (io.c:410-515: 4-line records, nt4 codes, '@' resync), writes the
 * binary .aln stream and prints one JSON line of workload counters on stdout. */
#include "oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

static uint8_t nt4_of(int c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'G': case 'g': return 1;
        case 'C': case 'c': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}

static int load_fastq(const char *path, uint8_t **seq, uint64_t **off, uint64_t *n) {
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    size_t cap = 1 << 20, len = 0, rcap = 1 << 16, nr = 0;
    uint8_t *s = malloc(cap);
    uint64_t *o = malloc((rcap + 1) * 8);
    o[0] = 0;
    int c;
    for (;;) {
        while ((c = getc(f)) != EOF && c != '@') {}
        if (c == EOF) break;
        while ((c = getc(f)) != EOF && c != '\n') {}          /* name */
        while ((c = getc(f)) != EOF && c != '\n') {           /* bases */
            if (len == cap) { cap *= 2; s = realloc(s, cap); }
            s[len++] = nt4_of(c);
        }
        while ((c = getc(f)) != EOF && c != '+') {}
        while ((c = getc(f)) != EOF && c != '\n') {}          /* + line */
        while ((c = getc(f)) != EOF && c != '\n') {}          /* qualities */
        if (nr == rcap) { rcap *= 2; o = realloc(o, (rcap + 1) * 8); }
        o[++nr] = len;
    }
    fclose(f);
    *seq = s; *off = o; *n = nr;
    return 0;
}

static double now(void) {
    struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

int main(int argc, char **argv) {
    if (argc < 5 || strcmp(argv[1], "align")) {
        fprintf(stderr, "usage: bwbble_oracle align [opts] <ref.fasta> <reads.fastq> <out.aln>\n");
        return 1;
    }
    orc_params p; orc_default_params(&p);
    int c;
    while ((c = getopt(argc - 1, argv + 1, "M:O:E:n:k:o:e:l:m:t:SP")) >= 0) {
        switch (c) {
            case 'M': p.mm_score = atoi(optarg); break;
            case 'O': p.gapo_score = atoi(optarg); break;
            case 'E': p.gape_score = atoi(optarg); break;
            case 'n': p.max_diff = atoi(optarg); break;
            case 'k': p.max_diff_seed = atoi(optarg); break;
            case 'o': p.max_gapo = atoi(optarg); break;
            case 'e': p.max_gape = atoi(optarg); break;
            case 'l': p.seed_length = atoi(optarg); break;
            case 'm': p.max_entries = atoi(optarg); break;
            case 't': p.n_threads = atoi(optarg); break;
            case 'S': p.is_multiref = 0; break;
            case 'P': p.use_precalc = 1; break;
            default: return 1;
        }
    }
    char **pos = argv + 1 + optind;
    if (argc - 1 - optind < 3) { fprintf(stderr, "missing arguments\n"); return 1; }
    char *bwtname = malloc(strlen(pos[0]) + 5);
    sprintf(bwtname, "%s.bwt", pos[0]);
    double t0 = now();
    orc_bwt *b = orc_bwt_load(bwtname, 0);
    uint8_t *seq; uint64_t *off, n;
    if (load_fastq(pos[1], &seq, &off, &n)) { fprintf(stderr, "cannot open %s\n", pos[1]); return 1; }
    double t1 = now();
    uint8_t *aln; uint64_t alen; orc_stats st;
    int rc = orc_align(b, &p, seq, off, n, &aln, &alen, &st);
    if (rc) { fprintf(stderr, "orc_align failed: %d\n", rc); return 1; }
    double t2 = now();
    FILE *f = fopen(pos[2], "wb");
    if (!f) { fprintf(stderr, "cannot open %s\n", pos[2]); return 1; }
    fwrite(aln, 1, alen, f);
    fclose(f);
    printf("{\"reads\": %llu, \"load_s\": %.4f, \"align_s\": %.4f, \"threads\": %d, \"n_O\": %llu, "
           "\"n_Oalpha\": %llu, \"pops\": %llu, \"pushes\": %llu, \"exact_tail\": %llu, "
           "\"max_heap\": %llu, \"max_list\": %llu, \"hits\": %llu}\n",
           (unsigned long long)n, t1 - t0, t2 - t1, p.n_threads, (unsigned long long)st.n_O,
           (unsigned long long)st.n_Oalpha, (unsigned long long)st.pops, (unsigned long long)st.pushes,
           (unsigned long long)st.exact_tail_calls, (unsigned long long)st.max_heap,
           (unsigned long long)st.max_list, (unsigned long long)st.hits);
    orc_free(aln); free(seq); free(off); orc_bwt_free(b); free(bwtname);
    return 0;
}

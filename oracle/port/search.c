/* oracle/port/search.c -- exact / inexact backward search (TEST INFRASTRUCTURE ONLY).
 * Restates exact_match.c:66-119,196-222; inexact_match.c:25-168 (drivers), :171-254
 * (calculate_d), :256-506 (inexact_match), :510-610 (bucket heap); align.c:93-110
 * (add_sa_interval), :271-298 (add_alignment), :345-382 (alns2alnf_bin).
 * Data structures are arrays instead of the reference's linked lists; order of every list,
 * of every push and of every pop is the reference's.  */
#include "oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

extern _Thread_local orc_stats orc_tls_stats;

/* io.h:28-33,102-111 */
static const uint8_t GRAY_VAL[16] = {0, 1, 3, 2, 6, 7, 5, 4, 12, 13, 15, 14, 10, 11, 9, 8};
static const uint8_t IS_SNP[16]   = {0, 0, 1, 0, 1, 1, 1, 0, 1, 1, 1, 1, 1, 1, 1, 0};
static const uint8_t BASES[4][7] = {{8, 9, 11, 12, 13, 14, 15}, {2, 3, 4, 5, 11, 12, 13},
                                    {4, 5, 6, 7, 8, 9, 11},     {1, 2, 5, 6, 9, 13, 14}};
static const uint8_t NT4_GRAY[5]     = {15, 3, 7, 1, 10};
static const uint8_t NT4_GRAY_VAL[5] = {8, 2, 4, 1, 15};
static const uint8_t NT4_COMPL[5]    = {3, 2, 1, 0, 4};

void orc_default_params(orc_params *p) {
    memset(p, 0, sizeof *p);
    p->gape_score = 4; p->gapo_score = 11; p->mm_score = 3;
    p->max_diff = 0; p->max_gape = 6; p->max_gapo = 1;
    p->seed_length = 32; p->max_diff_seed = 2; p->max_entries = 3000000;
    p->use_precalc = 0; p->matched_Ncontig = 0; p->is_multiref = 1;
    p->max_best = 30; p->no_indel_length = 5; p->n_threads = 1;
}

/* ---- interval lists ---------------------------------------------------------------- */
void orc_list_add(orc_list *l, uint64_t L, uint64_t U) {
    if (l->n && L == l->v[l->n - 1].U + 1) { l->v[l->n - 1].U = U; return; }   /* merge adjoining */
    if (l->n == l->cap) {
        l->cap = l->cap ? 2 * l->cap : 16;
        l->v = realloc(l->v, (size_t)l->cap * sizeof *l->v);
    }
    l->v[l->n].L = L; l->v[l->n].U = U; l->n++;
    if ((uint64_t)l->n > orc_tls_stats.max_list) orc_tls_stats.max_list = (uint64_t)l->n;
}

/* one backward-extension step of every interval of `cur` by read base c (multi-genome):
 * interval order x 7 compatible codes ascending (exact_match.c:88-101, inexact_match.c:219-232).
 * Returns the wrapped int sum of widths. */
static int extend_list(const orc_bwt *b, const orc_list *cur, orc_list *next, unsigned c) {
    int num = 0;
    next->n = 0;
    for (int s = 0; s < cur->n; s++) {
        for (int k = 0; k < 7; k++) {
            unsigned base = BASES[c][k];
            uint64_t L = b->C[base] + orc_O(b, base, cur->v[s].L - 1) + 1;
            uint64_t U = b->C[base] + orc_O(b, base, cur->v[s].U);
            if (L <= U) {
                num = (int)((unsigned)num + (unsigned)(U - L + 1));
                orc_list_add(next, L, U);
            }
        }
    }
    return num;
}

/* single-genome step, exact_match.c:196-222 */
static int step_1to1(const orc_bwt *b, unsigned code, uint64_t *L, uint64_t *U) {
    uint64_t oL, oU;
    if (*L - 1 == *U) { oL = orc_O(b, code, *L - 1); oU = oL; }
    else { oL = orc_O(b, code, *L - 1); oU = orc_O(b, code, *U); }
    *L = b->C[code] + oL + 1;
    *U = b->C[code] + oU;
    return *L <= *U;
}

int orc_exact_match_bounded(const orc_bwt *b, const uint8_t *read, int i, uint64_t l, uint64_t u,
                            const orc_params *p, orc_list *out) {
    out->n = 0;
    if (!p->is_multiref) {
        uint64_t L = l, U = u;
        for (int j = i; j >= 0; j--) {
            if (read[j] > 3) return 0;
            if (!step_1to1(b, NT4_GRAY[read[j]], &L, &U)) return 0;
        }
        orc_list_add(out, L, U);
        return 1;
    }
    orc_list tmp = {0};
    orc_list *cur = out, *nxt = &tmp;
    orc_list_add(cur, l, u);
    for (int r = i; r >= 0; r--) {
        unsigned c = read[r];
        if (c == 4) { cur->n = 0; break; }           /* N in the read never matches */
        extend_list(b, cur, nxt, c);
        orc_list *t = cur; cur = nxt; nxt = t;
        if (cur->n == 0) break;
    }
    if (cur != out) {                                 /* move the result into `out` */
        orc_list t = *out; *out = *cur; *cur = t;
    }
    free(tmp.v);
    return out->n != 0;
}

void orc_calculate_d(const orc_bwt *b, const uint8_t *read, int len, orc_dbound *D, const orc_params *p) {
    int z = 0;
    const uint64_t L0 = 0, U0 = b->length - 1;
    if (!p->is_multiref) {                            /* inexact_match.c:176-206 */
        uint64_t L = L0, U = U0;
        for (int i = len - 1; i >= 0; i--) {
            unsigned code = NT4_GRAY[read[i]];
            if (code == 10 || !step_1to1(b, code, &L, &U)) { L = L0; U = U0; z++; }
            D[len - 1 - i].z = z;
            D[len - 1 - i].w = (int)(U - L + 1);
        }
        D[len].w = 0; D[len].z = ++z;
        return;
    }
    orc_list a = {0}, c2 = {0};
    orc_list *cur = &a, *nxt = &c2;
    orc_list_add(cur, L0, U0);
    for (int i = len - 1; i >= 0; i--) {
        unsigned c = read[i];
        int num = 0;
        if (c > 3) nxt->n = 0;
        else num = extend_list(b, cur, nxt, c);
        orc_list *t = cur; cur = nxt; nxt = t;
        if (cur->n == 0) {                            /* restart from the full range */
            orc_list_add(cur, L0, U0);
            z++;
            num = (int)(U0 - L0 + 1);
        }
        D[len - 1 - i].z = z;
        D[len - 1 - i].w = num;
    }
    D[len].w = 0; D[len].z = ++z;
    free(a.v); free(c2.v);
}

/* ---- partial alignments + bucket heap ---------------------------------------------- */
typedef struct {
    uint64_t L, U;
    uint8_t mm, go, ge, snps;        /* 8-bit fields, align.h:100-104 (Q4) */
    uint8_t score, i, state, alen;
    uint8_t path[ORC_PATH_ALLOC];
} entry_t;

typedef struct { int n, cap; entry_t *e; } bucket_t;
typedef struct { int best, nb, n; bucket_t *b; } heap_t;

static inline int score_of(int m, int o, int e, const orc_params *p) {
    return m * p->mm_score + o * p->gapo_score + e * p->gape_score;
}

static heap_t *heap_new(const orc_params *p) {
    heap_t *h = calloc(1, sizeof *h);
    h->nb = score_of(p->max_diff + 1, p->max_gapo + 1, p->max_gape + 1, p);
    h->b = calloc((size_t)h->nb, sizeof *h->b);
    for (int i = 0; i < h->nb; i++) { h->b[i].cap = 4; h->b[i].e = calloc(4, sizeof(entry_t)); }
    h->best = h->nb;
    return h;
}
static void heap_del(heap_t *h) {
    for (int i = 0; i < h->nb; i++) free(h->b[i].e);
    free(h->b); free(h);
}
static uint64_t g_score_mask[16];   /* bit sc: an entry of score sc was pushed (process-wide, test hook) */
static void heap_clear(heap_t *h) {
    for (int i = 0; i < h->nb; i++) h->b[i].n = 0;
    h->best = h->nb; h->n = 0;
}
/* inexact_match.c:548-591.  parent==NULL is the root push (path pointer NULL in the reference:
 * length 0 and the slot's old path bytes stay, Q5). */
static void heap_push(heap_t *h, int i, uint64_t L, uint64_t U, int mm, int go, int ge, int state,
                      int snps, const entry_t *parent, const orc_params *p) {
    int sc = score_of(mm, go, ge, p);
    bucket_t *bk = &h->b[sc];
    if (bk->n == bk->cap) { bk->cap *= 2; bk->e = realloc(bk->e, (size_t)bk->cap * sizeof(entry_t)); }
    entry_t *q = &bk->e[bk->n];
    q->i = (uint8_t)i; q->score = (uint8_t)sc; q->L = L; q->U = U;
    q->mm = (uint8_t)mm; q->go = (uint8_t)go; q->ge = (uint8_t)ge;
    q->state = (uint8_t)state; q->snps = (uint8_t)snps; q->alen = 0;
    if (parent) {
        memset(q->path, 0, ORC_PATH_ALLOC);
        memcpy(q->path, parent->path, parent->alen);
        q->path[parent->alen] = (uint8_t)state;
        q->alen = (uint8_t)(parent->alen + 1);
    }
    bk->n++; h->n++;
    if (h->best > sc) h->best = sc;
    orc_tls_stats.pushes++;
    if ((uint64_t)h->n > orc_tls_stats.max_heap) orc_tls_stats.max_heap = (uint64_t)h->n;
    if (sc >= 0 && sc < 1024) __atomic_fetch_or(&g_score_mask[sc >> 6], 1ull << (sc & 63), __ATOMIC_RELAXED);
}
/* which scores were ever pushed since the last reset (tests: the device's bucket map must cover them) */
void orc_score_mask(uint64_t out[16], int reset) {
    for (int k = 0; k < 16; k++) {
        out[k] = __atomic_load_n(&g_score_mask[k], __ATOMIC_RELAXED);
        if (reset) __atomic_store_n(&g_score_mask[k], 0ull, __ATOMIC_RELAXED);
    }
}
/* inexact_match.c:594-610: last entry of the lowest non-empty bucket */
static void heap_pop(heap_t *h, entry_t *out) {
    bucket_t *bk = &h->b[h->best];
    const entry_t *top = &bk->e[bk->n - 1];
    bk->n--; h->n--;
    if (bk->n == 0 && h->n) {
        int i;
        for (i = h->best + 1; i < h->nb; i++) if (h->b[i].n) break;
        h->best = i;
    } else if (h->n == 0) {
        h->best = h->nb;
    }
    memcpy(out, top, sizeof *out);
    orc_tls_stats.pops++;
}

/* ---- hits --------------------------------------------------------------------------- */
typedef struct { int score; uint64_t L, U; int mm, go, ge, alen; uint8_t *path; } hit_t;
typedef struct { int n, cap; hit_t *h; } hits_t;

/* align.c:271-298 */
static void add_hit(hits_t *hs, const entry_t *e, uint64_t L, uint64_t U, int score) {
    if (e->go) for (int j = 0; j < hs->n; j++) if (hs->h[j].L == L && hs->h[j].U == U) return;
    if (hs->n == hs->cap) { hs->cap = hs->cap ? 2 * hs->cap : 4; hs->h = realloc(hs->h, (size_t)hs->cap * sizeof *hs->h); }
    hit_t *t = &hs->h[hs->n++];
    t->score = score; t->L = L; t->U = U; t->mm = e->mm; t->go = e->go; t->ge = e->ge;
    t->alen = e->alen;
    t->path = malloc(e->alen ? e->alen : 1);
    memcpy(t->path, e->path, e->alen);
    orc_tls_stats.hits++;
}

/* growable byte sink for the .aln stream */
typedef struct { uint8_t *p; size_t n, cap; } sink_t;
static void sink_put(sink_t *s, const void *src, size_t n) {
    if (s->n + n > s->cap) { s->cap = (s->cap ? 2 * s->cap : 4096) + n; s->p = realloc(s->p, s->cap); }
    memcpy(s->p + s->n, src, n); s->n += n;
}
static void put_i32(sink_t *s, int v) { sink_put(s, &v, 4); }
static void put_u64(sink_t *s, uint64_t v) { sink_put(s, &v, 8); }

/* align.c:345-382: path scanned from its last element to its first, runs as state|(count<<2),
 * count is a uint16_t (Q11) */
static void write_hits(sink_t *s, const hits_t *hs) {
    put_i32(s, hs->n);
    for (int k = 0; k < hs->n; k++) {
        const hit_t *t = &hs->h[k];
        put_i32(s, t->score); put_u64(s, t->L); put_u64(s, t->U);
        put_i32(s, t->mm); put_i32(s, t->go); put_i32(s, t->ge); put_i32(s, t->alen);
        if (t->alen <= 0) { put_i32(s, 0); continue; }
        int pairs[ORC_PATH_ALLOC], np = 0;
        int state = t->path[t->alen - 1];
        uint16_t run = 1;
        for (int j = t->alen - 2; j >= 0; j--) {
            if (t->path[j] == state) run++;
            else { pairs[np++] = state | (run << 2); state = t->path[j]; run = 1; }
        }
        pairs[np++] = state | (run << 2);
        put_i32(s, np);
        for (int j = 0; j < np; j++) put_i32(s, pairs[j]);
    }
}
static void free_hits(hits_t *hs) {
    for (int k = 0; k < hs->n; k++) free(hs->h[k].path);
    free(hs->h); hs->h = NULL; hs->n = hs->cap = 0;
}

/* ---- inexact_match, inexact_match.c:256-506 ---------------------------------------- */
static void inexact_match(const orc_bwt *b, const uint8_t *rc, int len, heap_t *heap, const orc_list *pre,
                          const orc_params *p, const orc_dbound *D, const orc_dbound *Ds, hits_t *hits) {
    int nN = 0;
    for (int i = 0; i < len; i++) nN += rc[i] > 3;
    if (nN > p->max_diff) return;

    heap_clear(heap);
    if (pre) {                                      /* -P seeds, inexact_match.c:269-279 */
        if (pre->n == 0) return;
        entry_t seed; memset(&seed, 0, sizeof seed);
        seed.alen = ORC_PRECALC_LEN - 1;            /* 11 matches; the push appends the 12th */
        for (int k = 0; k < pre->n; k++)
            heap_push(heap, len - ORC_PRECALC_LEN, pre->v[k].L, pre->v[k].U, 0, 0, 0, 0, 0, &seed, p);
    } else {
        heap_push(heap, len, 0, b->length - 1, 0, 0, 0, 0, 0, NULL, p);
    }

    int best_score = score_of(p->max_diff + 1, p->max_gapo + 1, p->max_gape + 1, p);
    int max_diff = p->max_diff;
    int num_best = 0;
    orc_list tail = {0};
    entry_t e;

    while (heap->n != 0) {
        if (heap->n > p->max_entries) break;
        heap_pop(heap, &e);
        if (e.score > best_score + p->mm_score) break;
        int used = e.mm + e.go + e.ge;
        int dl = max_diff - used;
        if (dl < 0) continue;
        if (e.i > 0 && dl < D[e.i - 1].z) continue;
        int dls = p->max_diff_seed - used;
        int si = e.i - (len - p->seed_length);
        if (si > 0 && dls < Ds[si - 1].z) continue;

        if (e.i == 0) {
            int sc = score_of(e.mm, e.go, e.ge, p);
            if (hits->n == 0) {
                best_score = sc;
                max_diff = (used + 1 > p->max_diff) ? p->max_diff : used + 1;
            }
            if (sc == best_score) num_best = (int)((unsigned)num_best + (unsigned)(e.U - e.L + 1));
            else if (num_best > p->max_best) break;
            add_hit(hits, &e, e.L, e.U, sc);
            continue;
        }
        if (dl == 0) {
            orc_tls_stats.exact_tail_calls++;
            if (orc_exact_match_bounded(b, rc, e.i - 1, e.L, e.U, p, &tail)) {
                int sc = score_of(e.mm, e.go, e.ge, p);
                if (hits->n == 0) {
                    best_score = sc;
                    max_diff = (used + 1 > p->max_diff) ? p->max_diff : used + 1;
                }
                if (sc == best_score) {
                    for (int k = 0; k < tail.n; k++)
                        num_best = (int)((unsigned)num_best + (unsigned)(tail.v[k].U - tail.v[k].L + 1));
                } else if (num_best > p->max_best) break;
                e.alen = (uint8_t)(e.alen + e.i);          /* rest of the path is M (=0) */
                for (int k = 0; k < tail.n; k++) add_hit(hits, &e, tail.v[k].L, tail.v[k].U, sc);
            }
            continue;
        }

        uint64_t Lo[16] = {0}, Up[16] = {0};
        int nsym = 16;
        if (p->is_multiref) {
            orc_O_alphabet(b, e.L - 1, Lo, 1);
            orc_O_alphabet(b, e.U, Up, 0);
        } else {
            orc_O_actg(b, e.L - 1, Lo, 1);
            orc_O_actg(b, e.U, Up, 0);
            nsym = 5;
        }

        int allow_diff = 1, allow_indels = 1, allow_mm = 1, allow_open = 1, allow_ext = 1;
        int i1 = e.i - 1;
        if (i1 > 0) {
            if (dl - 1 < D[i1 - 1].z) allow_diff = 0;
            else if (D[i1].z == dl - 1 && D[i1 - 1].z == dl - 1 && D[i1].w == D[i1 - 1].w) allow_mm = 0;
        }
        if (si - 1 > 0) {
            if (dls - 1 < Ds[si - 2].z) allow_diff = 0;
            else if (Ds[si - 1].z == dls - 1 && Ds[si - 2].z == dls - 1 && Ds[si - 1].w == Ds[si - 2].w) allow_mm = 0;
        }
        int g = e.go + e.ge;
        if (i1 < p->no_indel_length + g || len - i1 < p->no_indel_length + g) allow_indels = 0;
        if (e.go >= p->max_gapo && e.ge >= p->max_gape) allow_indels = 0;
        if (e.go >= p->max_gapo) allow_open = 0;
        if (e.ge >= p->max_gape) allow_ext = 0;

        if (allow_diff && allow_indels) {
            if (e.state == 1) {
                if (allow_ext) heap_push(heap, i1, e.L, e.U, e.mm, e.go, e.ge + 1, 1, e.snps, &e, p);
            } else {
                if (allow_open && e.state == 0)
                    heap_push(heap, i1, e.L, e.U, e.mm, e.go + 1, e.ge, 1, e.snps, &e, p);
                for (int j = 1; j < nsym; j++) {
                    if (Lo[j] > Up[j]) continue;
                    if (e.state == 0) {
                        if (allow_open) heap_push(heap, e.i, Lo[j], Up[j], e.mm, e.go + 1, e.ge, 2, e.snps, &e, p);
                    } else if (allow_ext) {
                        heap_push(heap, e.i, Lo[j], Up[j], e.mm, e.go, e.ge + 1, 2, e.snps, &e, p);
                    }
                }
            }
        }

        unsigned c = rc[i1];
        if (allow_diff && allow_mm) {
            for (int j = 1; j < nsym; j++) {
                if (Lo[j] > Up[j]) continue;
                int is_mm;
                if (p->is_multiref) is_mm = (c > 3) || j == 10 || ((NT4_GRAY_VAL[c] & GRAY_VAL[j]) == 0);
                else is_mm = (c > 3) || ((int)c != j - 1);
                heap_push(heap, i1, Lo[j], Up[j], e.mm + is_mm, e.go, e.ge, 0,
                          e.snps + (p->is_multiref && IS_SNP[j]), &e, p);
            }
        } else if (c < 4) {
            if (p->is_multiref) {
                for (int k = 0; k < 7; k++) {
                    unsigned j = BASES[c][k];
                    if (Lo[j] <= Up[j]) heap_push(heap, i1, Lo[j], Up[j], e.mm, e.go, e.ge, 0, e.snps + IS_SNP[j], &e, p);
                }
            } else if (Lo[c + 1] <= Up[c + 1]) {
                heap_push(heap, i1, Lo[c + 1], Up[c + 1], e.mm, e.go, e.ge, 0, e.snps, &e, p);
            }
        }
    }
    free(tail.v);
}

/* ---- batch drivers, inexact_match.c:25-168 ------------------------------------------ */
typedef struct { orc_dbound *D, *Ds; heap_t *heap; uint8_t *rc; orc_list pre; } work_t;

static void work_init(work_t *w, int max_len, const orc_params *p) {
    w->D = calloc((size_t)max_len + 1, sizeof *w->D);
    w->Ds = calloc((size_t)p->seed_length + 1, sizeof *w->Ds);
    w->heap = heap_new(p);
    w->rc = malloc((size_t)max_len + 1);
    memset(&w->pre, 0, sizeof w->pre);
}
static void work_free(work_t *w) { free(w->D); free(w->Ds); heap_del(w->heap); free(w->rc); free(w->pre.v); }

/* read2index, align.c:174-186: the last 12 bases as a base-4 number, <0 if any of them is N */
long orc_read2index(const uint8_t *read, int len) {
    long idx = 0;
    for (int i = len - ORC_PRECALC_LEN; i < len; i++) {
        if (read[i] > 3) return -1;
        idx = idx * 4 + read[i];
    }
    return idx;
}

/* entry `index` of the .pre table: exact_match() of the 12-mer whose base-4 digits are `index`
 * (most significant first; precalc_sa_intervals + next_read, align.c:188-224) */
int orc_precalc_entry(const orc_bwt *b, const orc_params *p, uint32_t index, orc_list *out) {
    uint8_t kmer[ORC_PRECALC_LEN];
    for (int k = 0; k < ORC_PRECALC_LEN; k++) kmer[k] = (uint8_t)((index >> (2 * (ORC_PRECALC_LEN - 1 - k))) & 3u);
    return orc_exact_match_bounded(b, kmer, ORC_PRECALC_LEN - 1, 0, b->length - 1, p, out);
}

static void align_one(const orc_bwt *b, const orc_params *p, const uint8_t *seq, int len, work_t *w, hits_t *hits) {
    for (int i = 0; i < len; i++) w->rc[len - 1 - i] = NT4_COMPL[seq[i] > 4 ? 4 : seq[i]];
    const orc_list *pre = NULL;
    if (p->use_precalc) {                           /* inexact_match.c:50-57: the table row of rc's last 12 bases */
        long idx = len >= ORC_PRECALC_LEN ? orc_read2index(w->rc, len) : -1;
        if (idx < 0) return;                        /* N among them: the read is skipped altogether */
        orc_precalc_entry(b, p, (uint32_t)idx, &w->pre);   /* same list the table holds for this row */
        pre = &w->pre;
    }
    orc_calculate_d(b, seq, len, w->D, p);
    if (p->seed_length && len > p->seed_length) orc_calculate_d(b, seq, p->seed_length, w->Ds, p);
    inexact_match(b, w->rc, len, w->heap, pre, p, w->D, w->Ds, hits);
}

static void stats_add(orc_stats *a, const orc_stats *t) {
    a->n_O += t->n_O; a->n_O_shortcut += t->n_O_shortcut; a->n_Oalpha += t->n_Oalpha;
    a->n_Oalpha_shortcut += t->n_Oalpha_shortcut; a->pops += t->pops; a->pushes += t->pushes;
    a->exact_tail_calls += t->exact_tail_calls; a->hits += t->hits;
    if (t->max_heap > a->max_heap) a->max_heap = t->max_heap;
    if (t->max_list > a->max_list) a->max_list = t->max_list;
}

int orc_align(const orc_bwt *b, const orc_params *p, const uint8_t *seq, const uint64_t *offsets,
              uint64_t n_reads, uint8_t **aln, uint64_t *aln_len, orc_stats *stats) {
    int nb = score_of(p->max_diff + 1, p->max_gapo + 1, p->max_gape + 1, p);
    if (nb <= 0) return -2;
    int max_len = 0;
    for (uint64_t r = 0; r < n_reads; r++) {
        uint64_t l = offsets[r + 1] - offsets[r];
        if (l > 255) return -3;                     /* 8-bit read position, align.h:104 */
        if ((int)l > max_len) max_len = (int)l;
    }
    sink_t out = {0};
    orc_stats total; memset(&total, 0, sizeof total);
    int nthr = p->n_threads > 1 ? p->n_threads : 1;

    if (nthr == 1) {                                /* serial driver: state lives for the whole run */
        memset(&orc_tls_stats, 0, sizeof orc_tls_stats);
        work_t w; work_init(&w, max_len, p);
        hits_t hs = {0};
        for (uint64_t r = 0; r < n_reads; r++) {
            align_one(b, p, seq + offsets[r], (int)(offsets[r + 1] - offsets[r]), &w, &hs);
            write_hits(&out, &hs);
            free_hits(&hs);
        }
        work_free(&w);
        stats_add(&total, &orc_tls_stats);
    } else {
        for (uint64_t done = 0; done < n_reads; done += ORC_READ_BATCH) {
            uint64_t bs = n_reads - done > ORC_READ_BATCH ? ORC_READ_BATCH : n_reads - done;
            hits_t *res = calloc(bs, sizeof *res);
            #pragma omp parallel num_threads(nthr)
            {
                int tid = omp_get_thread_num(), nt = omp_get_num_threads();
                uint64_t lo = (uint64_t)tid * bs / (uint64_t)nt, hi = (uint64_t)(tid + 1) * bs / (uint64_t)nt;
                memset(&orc_tls_stats, 0, sizeof orc_tls_stats);
                work_t w; work_init(&w, max_len, p);
                for (uint64_t r = lo; r < hi; r++) {
                    uint64_t g = done + r;
                    align_one(b, p, seq + offsets[g], (int)(offsets[g + 1] - offsets[g]), &w, &res[r]);
                }
                work_free(&w);
                #pragma omp critical
                stats_add(&total, &orc_tls_stats);
            }
            for (uint64_t r = 0; r < bs; r++) { write_hits(&out, &res[r]); free_hits(&res[r]); }
            free(res);
        }
    }
    *aln = out.p; *aln_len = out.n;
    if (stats) *stats = total;
    return 0;
}

// bwb_group.cuh -- the production search kernels: 8 lanes per read, 4 reads per warp.
//
// Why sub-warp groups: the per-pop work of inexact_match (unpack the entry, prune, BWA heuristics,
// bucket bookkeeping) is uniform per read.  With a whole warp per read those ~400 instructions are
// issued for one read at a time (ncu: 2.7 M warp-instructions per read, issue-bound at IPC 1.7).
// Here a read owns 8 lanes, so one warp instruction advances 4 independent reads; the 32 rank
// tasks of an expansion (16 codes x {L-1, U}) become 4 per lane, the <=31 children 4 per lane.
//
// Groups in different phases would serialise if they simply diverged, so every warp iteration is
// a fixed sequence of predicated phases (NEED_READ -> list pass -> pop step -> flush) and every
// warp-synchronous intrinsic sits at phase level with the full mask; per-group results are taken
// from the group's byte of a ballot / a width-8 shuffle.
//
//   K3 k_calc_d_g   calculate_d (inexact_match.c:171-254) for D and D_seed -> HBM (int2 per base)
//   K4 k_search_g   inexact_match (inexact_match.c:256-506) incl. the exact tails and the bucket heap
#pragma once
#include "bwb_kernels.cuh"

namespace bwb {

constexpr int GL = 8;        // lanes per read group
constexpr int SLG = 16;      // intervals of each list kept in shared memory per group
constexpr int G_LIST_SMEM = 2 * SLG * (int)sizeof(ulonglong2);

__device__ __forceinline__ uint32_t g_lane() { return threadIdx.x & 7u; }
__device__ __forceinline__ uint32_t g_shift() { return threadIdx.x & 24u; }
__device__ __forceinline__ uint32_t gballot(bool p) { return (__ballot_sync(FULL, p) >> g_shift()) & 0xffu; }
__device__ __forceinline__ uint32_t gshfl(uint32_t v, int src) { return __shfl_sync(FULL, v, src, GL); }
__device__ __forceinline__ uint64_t gshfl(uint64_t v, int src) {
    uint32_t lo = __shfl_sync(FULL, (uint32_t)v, src, GL);
    uint32_t hi = __shfl_sync(FULL, (uint32_t)(v >> 32), src, GL);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint32_t gsum(uint32_t v) {
    v += __shfl_xor_sync(FULL, v, 1, GL);
    v += __shfl_xor_sync(FULL, v, 2, GL);
    v += __shfl_xor_sync(FULL, v, 4, GL);
    return v;
}

// ---- per-group interval lists -----------------------------------------------------------------
template <class T>
struct GList {
    typedef typename Pair<T>::type P;
    P *s;     // [2][SLG] shared
    P *g;     // [2][cap] HBM
    int cap;
};
template <class T>
__device__ __forceinline__ typename Pair<T>::type glget(const GList<T> &ls, int which, int k) {
    return k < SLG ? ls.s[which * SLG + k] : ls.g[(size_t)which * ls.cap + k];
}
template <class T>
__device__ __forceinline__ void glset(const GList<T> &ls, int which, int k, T L, T U) {
    typename Pair<T>::type v;
    v.x = L; v.y = U;
    if (k < SLG) ls.s[which * SLG + k] = v;
    else ls.g[(size_t)which * ls.cap + k] = v;
}

// state of one backward-extension step in progress (one interval of the current list per pass)
template <class T>
struct StepState {
    int cur;          // list holding the current intervals
    int n_cur;        // its length
    int s;            // next interval to extend
    int n_next;       // intervals written to list cur^1 so far
    bool tail_valid;
    T tailU;          // U of the interval written last
    uint32_t acc;     // wrapped sum of widths (per lane partial)
};

// One pass = interval st.s of list st.cur x the 7 codes compatible with read base c, lanes 0..6.
// Same ordered adjacent-merge as extend_step (align.c:93-110), on the group's byte of the ballots.
// `on` is uniform per group.  Returns false (for the group) if the list overflows ls.cap.
template <class T>
__device__ __forceinline__ bool list_pass(bool on, const IndexView &ix, const T *sC, const GList<T> &ls,
                                          StepState<T> &st, uint32_t c, uint32_t &nloads, bool multiref = true) {
    const uint32_t gl = g_lane();
    // multi-genome: the 7 codes containing base c on lanes 0..6; single-genome (-S): nt4_gray[c] on lane 0
    const bool active = on && (multiref ? gl < 7u : gl == 0u) && st.s < st.n_cur;
    const uint32_t code = multiref ? ((compat_codes(c & 3u) >> (4u * (gl < 7u ? gl : 0u))) & 15u)
                                   : ((0x173Fu >> (4u * (c & 3u))) & 15u);
    typename Pair<T>::type iv;
    iv.x = 1; iv.y = 0;
    if (active) iv = glget<T>(ls, st.cur, st.s);
    T oL, oU;
    occ_pair<T>(ix, sC, code, (T)(iv.x - 1), iv.y, oL, oU);
    const T Cc = sC[code];
    const T nL = (T)(Cc + oL + 1), nU = (T)(Cc + oU);
    const bool valid = active && (nL <= nU);
    nloads += active ? 2u : 0u;

    const uint32_t lt = (1u << gl) - 1u;
    const uint32_t V = gballot(valid);
    const uint32_t below = V & lt;
    const T prevU = gshfl(nU, below ? (31 - __clz(below)) : 0);
    const bool cmp_ok = below ? true : st.tail_valid;
    const T cmpU = below ? prevU : st.tailU;
    const bool head = valid && !(cmp_ok && nL == (T)(cmpU + 1));
    const uint32_t H = gballot(head);
    const uint32_t above = H & ~((2u << gl) - 1u);
    const uint32_t lim = above ? ((1u << (__ffs(above) - 1)) - 1u) : 0xffu;
    const uint32_t runV = V & lim;
    const T endU = gshfl(nU, runV ? (31 - __clz(runV)) : 0);
    const uint32_t leadV = V & (H ? ((1u << (__ffs(H) - 1)) - 1u) : 0xffu);
    const T leadU = gshfl(nU, leadV ? (31 - __clz(leadV)) : 0);
    const T lastU = gshfl(nU, V ? (31 - __clz(V)) : 0);
    const int nxt = st.cur ^ 1;
    bool ok = true;
    if (on) {
        if (leadV && gl == 0) {
            typename Pair<T>::type t = glget<T>(ls, nxt, st.n_next - 1);
            glset<T>(ls, nxt, st.n_next - 1, t.x, leadU);
        }
        const int nh = __popc(H);
        if (st.n_next + nh > ls.cap) ok = false;
        else if (head) glset<T>(ls, nxt, st.n_next + __popc(H & lt), nL, endU);
        st.n_next += nh;
        if (V) { st.tailU = lastU; st.tail_valid = true; }
        st.acc += valid ? (uint32_t)(nU - nL + 1) : 0u;
        st.s++;
    }
    return ok;
}

template <class T>
__device__ __forceinline__ void step_begin(StepState<T> &st) {
    st.s = 0; st.n_next = 0; st.tail_valid = false; st.tailU = 0; st.acc = 0;
}

// stage a read into the group's shared memory; returns the number of N bases (uniform per group)
__device__ __forceinline__ uint32_t g_stage_read(bool on, const uint8_t *__restrict__ g, int len, uint8_t *s) {
    uint32_t nN = 0;
    if (on)
        for (int k = g_lane(); k < len; k += GL) {
            uint8_t c = g[k];
            if (c > 4) c = 4;
            s[k] = c;
            nN += (c == 4);
        }
    return gsum(nN);
}

// ---------------------------------------------------------------------------------------------
// K3: calculate_d for every read (D over the whole read, D_seed over the first seed_len bases)
// ---------------------------------------------------------------------------------------------
struct CalcArgs {
    IndexView ix;
    const uint8_t *seq;
    const uint64_t *offsets;
    uint32_t n_reads;
    int seed_len;            // 0: no seed array
    int max_len;
    int is_multiref;         // 0 = -S: one code per base (inexact_match.c:176-206)
    uint32_t *queue;
    void *glists;            // [n_groups][2][list_cap] pairs (allocated 16 B each)
    int list_cap;
    int2 *d_main;            // per read (len+1) entries at offsets[r] + r
    int2 *d_seed;            // per read (seed_len+1) entries at r*(seed_len+1); reads with len <= seed_len: see seed_src
    uint16_t *pk_main, *pk_seed;   // same arrays packed for K4: num_diff | (width == previous width) << 15
    uint16_t *n_count;             // number of N bases per read (inexact_match.c:259-263)
    uint32_t *status;
    unsigned long long *counters;   // [3] rank queries, [5] max list
    int smem_per_group;
    // k-mer table of the top of the search tree (k_kmer_table); ktab_w == nullptr disables it
    const unsigned long long *ktab_w;
    const uint32_t *ktab_off, *ktab_cnt;
    const void *ktab_iv;
    // a read with len <= seed_len gets no D_seed of its own in the reference -- it consults the array the
    // previous longer read of its thread left behind (inexact_match.c:36,62-64,121,141-143).  The host works out which
    // read that is (seed_donors, bwb_abi.cu): seed_src[r] = offset in `seq` of that read's first base, SEED_SRC_EXT =
    // the donor is not in this shard (its first seed_len bases are in seed_ext), SEED_SRC_NONE = none (the calloc'ed
    // array: zeros).  Only looked at for reads with len <= seed_len; null = no such read has a donor.
    const uint32_t *seed_src;
    const uint8_t *seed_ext;
};
constexpr uint32_t SEED_SRC_NONE = 0xffffffffu, SEED_SRC_EXT = 0xfffffffeu;

template <bool WIDE>
__global__ void __launch_bounds__(256) k_calc_d_g(const __grid_constant__ CalcArgs a) {
    typedef typename Coord<WIDE>::type T;
    typedef typename Pair<T>::type P;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ T sC[17];
    stage_C<T>(a.ix, sC);
    const uint32_t gl = g_lane();
    const uint32_t gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    unsigned char *gbase = smem + (size_t)(threadIdx.x >> 3) * a.smem_per_group;
    GList<T> ls;
    ls.s = reinterpret_cast<P *>(gbase);
    ls.g = reinterpret_cast<P *>(a.glists) + (size_t)gid * 2 * a.list_cap;
    ls.cap = a.list_cap;
    uint8_t *sseq = gbase + G_LIST_SMEM;
    const T fullU = (T)(a.ix.length - 1);

    enum { NEED = 0, RUN = 1, DONE = 2 };
    int mode = NEED;
    uint32_t r = 0;
    int len = 0, dlen = 0, phase = 0, i = 0, z = 0;
    uint64_t off = 0;
    int2 *D = nullptr;
    uint16_t *PK = nullptr;
    uint32_t prev_w = 0;
    StepState<T> st;
    st.cur = 0; st.n_cur = 0;
    step_begin(st);
    uint32_t c = 0;
    bool in_step = false;
    bool fresh = false;           // the current list is the full range (start of an array / after a restart)
    const bool use_tab = a.ktab_w != nullptr && a.is_multiref != 0;
    uint32_t nloads = 0, maxlist = 0;

    for (;;) {
        if (__any_sync(FULL, mode == NEED)) {
            uint32_t rr = 0;
            if (mode == NEED && gl == 0) rr = atomicAdd(a.queue, 1u);
            rr = gshfl(rr, 0);
            const bool got = (mode == NEED) && rr < a.n_reads;
            if (mode == NEED && !got) mode = DONE;
            if (got) {
                r = rr;
                off = a.offsets[r];
                len = (int)(a.offsets[r + 1] - off);
            }
            const uint32_t nN = g_stage_read(got, a.seq + off, len, sseq);
            if (got) {
                if (a.n_count && gl == 0) a.n_count[r] = (uint16_t)nN;
                phase = 0; dlen = len; D = a.d_main ? a.d_main + off + r : nullptr;
                PK = a.pk_main ? a.pk_main + off + r : nullptr;
                i = dlen - 1; z = 0; st.cur = 0; st.n_cur = 1; in_step = false; fresh = true; prev_w = 0;
                if (gl == 0) glset<T>(ls, 0, 0, (T)0, fullU);
                mode = RUN;
            }
            __syncwarp();
        }
        if (__all_sync(FULL, mode == DONE)) break;

        // ---- table step: from the full range, the next KTAB steps are a function of the next KTAB bases
        bool run = (mode == RUN);
        {
            const bool cand = use_tab && run && !in_step && fresh && i >= KTAB - 1;
            if (__any_sync(FULL, cand)) {
                // lane l serves steps s = l+1 and l+9 (<= KTAB); base of step s is seq[i-s+1]
                const int s1 = (int)gl + 1, s2 = (int)gl + 9;
                const uint32_t b1 = cand ? sseq[i - s1 + 1] : 0u;
                const uint32_t b2 = (cand && s2 <= KTAB) ? sseq[i - s2 + 1] : 0u;
                const bool hasN = gballot(b1 > 3u || b2 > 3u) != 0u;
                const bool go = cand && !hasN;
                const uint32_t X = gsum(((b1 & 3u) << (2 * (s1 - 1))) | (s2 <= KTAB ? ((b2 & 3u) << (2 * (s2 - 1))) : 0u));
                unsigned long long e1 = 0, e2 = 0, p1 = 0, p2 = 0;     // entries of steps s1, s2 and of the steps before them
                if (go) {
                    e1 = a.ktab_w[ktab_level_off(s1) + (X & ((1u << (2 * s1)) - 1u))];
                    if (s1 > 1) p1 = a.ktab_w[ktab_level_off(s1 - 1) + (X & ((1u << (2 * (s1 - 1))) - 1u))];
                    if (s2 <= KTAB) {
                        e2 = a.ktab_w[ktab_level_off(s2) + (X & ((1u << (2 * s2)) - 1u))];
                        p2 = a.ktab_w[ktab_level_off(s2 - 1) + (X & ((1u << (2 * (s2 - 1))) - 1u))];
                    }
                }
                // first step whose list is empty (KTAB+1 if none)
                const uint32_t em1 = gballot(go && !((e1 >> 32) & 1ull));
                const uint32_t em2 = gballot(go && s2 <= KTAB && !((e2 >> 32) & 1ull));
                const int first_empty = em1 ? __ffs(em1) : (em2 ? 8 + __ffs(em2) : KTAB + 1);
                const uint32_t wlast = (uint32_t)gshfl((uint32_t)e2, 1);                 // width after step KTAB (= s2 of lane 1)
                const uint32_t lenw = (uint32_t)a.ix.length;
                uint32_t ml = 0;
                if (go) {
                    const int k0 = dlen - 1 - i;                       // D index of step 1
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int sx = h ? s2 : s1;
                        const unsigned long long ee = h ? e2 : e1, pp = h ? p2 : p1;
                        if (sx <= KTAB && sx <= first_empty) {
                            const bool emptying = (sx == first_empty);
                            const uint32_t num = emptying ? lenw : (uint32_t)ee;
                            const uint32_t before = (sx == 1) ? prev_w : (uint32_t)pp;
                            const int zz = emptying ? z + 1 : z;
                            const int k = k0 + sx - 1;
                            if (D) D[k] = make_int2(zz, (int)num);
                            if (PK) PK[k] = (uint16_t)((zz & 0x1ff) | ((k && num == before) ? 0x8000 : 0));
                            if (!emptying) ml = max(ml, (uint32_t)(ee >> 33));
                        }
                    }
                }
                ml = max(ml, __shfl_xor_sync(FULL, ml, 1, GL));
                ml = max(ml, __shfl_xor_sync(FULL, ml, 2, GL));
                ml = max(ml, __shfl_xor_sync(FULL, ml, 4, GL));
                if (go) {
                    if (ml > maxlist) maxlist = ml;
                    if (first_empty <= KTAB) {                         // restart inside the window: stay fresh
                        i -= first_empty;
                        z++;
                        prev_w = lenw;
                    } else {                                           // adopt the tabulated list
                        const uint32_t n = a.ktab_cnt[X], o = a.ktab_off[X];
                        const P *src = reinterpret_cast<const P *>(a.ktab_iv) + o;
                        for (uint32_t k = gl; k < n; k += GL) { const P v = src[k]; glset<T>(ls, 0, (int)k, v.x, v.y); }
                        st.cur = 0; st.n_cur = (int)n;
                        i -= KTAB;
                        prev_w = wlast;
                        fresh = false;
                    }
                }
                __syncwarp();
            }
        }

        // begin a step: read base i of the forward read; an N (or a finished array) needs no pass
        if (run && !in_step && i >= 0) {
            c = sseq[i];
            step_begin(st);
            in_step = true;
            if (c > 3u) st.s = st.n_cur;             // no extension: the step ends with an empty list
        }
        const bool pass_on = run && in_step && st.s < st.n_cur;
        if (__any_sync(FULL, pass_on)) {
            if (!list_pass<T>(pass_on, a.ix, sC, ls, st, c, nloads, a.is_multiref != 0)) {
                if (gl == 0) atomicExch(a.status, (uint32_t)(-BWB_ERR_CAPACITY));
                st.n_next = 0;
            }
            __syncwarp();
        }
        // end of step
        const bool fin = run && in_step && st.s >= st.n_cur;
        const uint32_t num_all = gsum(st.acc);
        if (fin) {
            uint32_t num = (c <= 3u) ? num_all : 0u;
            int nn = (c <= 3u) ? st.n_next : 0;
            if (c <= 3u) st.cur ^= 1;
            if ((uint32_t)nn > maxlist) maxlist = (uint32_t)nn;
            fresh = false;
            if (nn == 0) {                           // restart from the full range, one more difference
                if (gl == 0) glset<T>(ls, st.cur, 0, (T)0, fullU);
                nn = 1;
                z++;
                num = (uint32_t)a.ix.length;
                fresh = true;
            }
            st.n_cur = nn;
            if (gl == 0) {
                const int k = dlen - 1 - i;
                if (D) D[k] = make_int2(z, (int)num);
                if (PK) PK[k] = (uint16_t)((z & 0x1ff) | ((k && num == prev_w) ? 0x8000 : 0));
            }
            prev_w = num;
            i--;
            in_step = false;
        }
        if (run && !in_step && i < 0) {              // array complete
            if (gl == 0) {
                if (D) D[dlen] = make_int2(z + 1, 0);
                if (PK) PK[dlen] = (uint16_t)(((z + 1) & 0x1ff) | ((dlen && prev_w == 0u) ? 0x8000 : 0));
            }
            if (phase == 0 && a.seed_len > 0) {
                int2 *Ds = a.d_seed ? a.d_seed + (size_t)r * (a.seed_len + 1) : nullptr;
                uint16_t *PKs = a.pk_seed ? a.pk_seed + (size_t)r * (a.seed_len + 1) : nullptr;
                if (len > a.seed_len) {
                    phase = 1; dlen = a.seed_len; D = Ds; PK = PKs;
                    i = dlen - 1; z = 0; st.cur = 0; st.n_cur = 1; fresh = true; prev_w = 0;
                    if (gl == 0) glset<T>(ls, 0, 0, (T)0, fullU);
                } else {
                    // Q6: the array of the donor read (computed here from the donor's first seed_len bases, which
                    // replace the finished read in the group's staging buffer), or the calloc'ed zeros
                    const uint32_t src = a.seed_src ? a.seed_src[r] : SEED_SRC_NONE;
                    if (src != SEED_SRC_NONE) {
                        const uint8_t *g = src == SEED_SRC_EXT ? a.seed_ext : a.seq + src;
                        for (int k = gl; k < a.seed_len; k += GL) {
                            uint8_t cc = g[k];
                            sseq[k] = cc > 4 ? (uint8_t)4 : cc;
                        }
                        phase = 1; dlen = a.seed_len; D = Ds; PK = PKs;
                        i = dlen - 1; z = 0; st.cur = 0; st.n_cur = 1; fresh = true; prev_w = 0;
                        if (gl == 0) glset<T>(ls, 0, 0, (T)0, fullU);
                    } else {
                        for (int k = gl; k <= a.seed_len; k += GL) {
                            if (Ds) Ds[k] = make_int2(0, 0);
                            if (PKs) PKs[k] = (uint16_t)(k ? 0x8000 : 0);
                        }
                        mode = NEED;
                    }
                }
            } else {
                mode = NEED;
            }
        }
        __syncwarp();
    }
    const uint32_t tot = gsum(nloads);
    if (gl == 0) {
        atomicAdd(a.counters + 3, (unsigned long long)tot);
        atomicMax(a.counters + 5, (unsigned long long)maxlist);
    }
}

// ---------------------------------------------------------------------------------------------
// K4: inexact_match, 8 lanes per read
// ---------------------------------------------------------------------------------------------
constexpr int POOL_SHARDS = 128;
constexpr int POOL_BATCH = 8;
struct PoolState {
    unsigned long long head[POOL_SHARDS];
    uint32_t bump[POOL_SHARDS];
    uint32_t limit[POOL_SHARDS];
    unsigned long long n_borrowed;      // diagnostics
};

struct SearchArgs {
    IndexView ix;
    const uint8_t *seq;
    const uint64_t *offsets;
    uint32_t n_reads;
    uint32_t read_id_base;
    int max_diff, max_gapo, max_gape, max_entries, mm_score, gapo_score, gape_score;
    int seed_len, max_diff_seed, max_best, no_indel_len;
    int nb;
    int max_len;
    uint32_t *queue;
    const int2 *d_main, *d_seed;     // from K3
    void *glists;
    int list_cap;
    uint4 *chunks;
    uint32_t *chunk_link;
    uint32_t chunks_per_group;
    uint32_t n_chunks;
    uint32_t priv_total;             // chunk ids below this belong to private ranges
    PoolState *pool;                 // shared pool above them
    bwb_hit *stage;
    int hits_cap;
    bwb_hit *out_hits;
    unsigned long long out_cap;
    unsigned long long *out_cursor;
    unsigned long long *read_off;
    uint32_t *read_cnt;
    uint32_t *status;
    unsigned long long *counters;
    int smem_per_group, off_D, off_Ds, off_bk, off_seq;
};

// lower-bound arrays packed to 16 bits in shared memory: low 9 bits num_diff, bit 15 = width equals
// the previous entry's width (the only way sa_intv_width is used, inexact_match.c:402-403,411-412)
__device__ __forceinline__ void g_load_bounds(bool on, const int2 *__restrict__ src, int n, uint16_t *dst) {
    if (on)
        for (int k = g_lane(); k < n; k += GL) {
            const int2 v = src[k];
            const int pw = k ? src[k - 1].y : ~v.y;
            dst[k] = (uint16_t)((v.x & 0x1ff) | ((k && pw == v.y) ? 0x8000 : 0));
        }
}

struct GHeap {
    uint32_t *cnt, *top, *bot;    // shared memory, nb each (per group)
    uint32_t priv_hi, bump, free_head;
    bool took_shared;             // the group holds chunks of the shared pool (returned at flush)
    int n, best;
};

// Shared chunk pool = every chunk above the private ranges, split into POOL_SHARDS regions so that no
// single address serialises the atomics.  Per shard: a lock-free LIFO of returned chunks
// (head = tag<<32 | chunk id; the tag defeats ABA) in front of a bump cursor.  Chunks move in
// batches: one CAS pops up to POOL_BATCH chunks or returns a whole chain.  Group leaders only.
// pops up to POOL_BATCH chunks as a chain first -> ... -> last (links intact); returns the count
__device__ __forceinline__ int pool_pop_chain(unsigned long long *head, uint32_t *link, uint32_t &first, uint32_t &last) {
    unsigned long long old = atomicAdd(head, 0ull);
    for (;;) {
        first = (uint32_t)old;
        if (first == NO_CHUNK) return 0;
        uint32_t cur = first;
        int n = 1;
        while (n < POOL_BATCH) {
            const uint32_t nxt = *reinterpret_cast<volatile uint32_t *>(link + cur);
            if (nxt == NO_CHUNK) break;
            cur = nxt;
            n++;
        }
        last = cur;
        const uint32_t after = *reinterpret_cast<volatile uint32_t *>(link + last);
        const unsigned long long nw = (((old >> 32) + 1ull) << 32) | after;
        const unsigned long long prev = atomicCAS(head, old, nw);
        if (prev == old) return n;
        old = prev;
    }
}
__device__ __forceinline__ void pool_push_chain(unsigned long long *head, uint32_t *link, uint32_t first, uint32_t last) {
    unsigned long long old = atomicAdd(head, 0ull);
    for (;;) {
        *reinterpret_cast<volatile uint32_t *>(link + last) = (uint32_t)old;
        __threadfence();
        const unsigned long long nw = (((old >> 32) + 1ull) << 32) | first;
        const unsigned long long prev = atomicCAS(head, old, nw);
        if (prev == old) return;
        old = prev;
    }
}

// leader: refill the group's free list from the shared pool (own shard first); false if exhausted
__device__ __forceinline__ bool pool_borrow(PoolState *ps, uint32_t *link, uint32_t gid, uint32_t &free_head) {
    for (int t = 0; t < POOL_SHARDS; t++) {
        const uint32_t sh = (gid + (uint32_t)t) % POOL_SHARDS;
        uint32_t first, last;
        if (pool_pop_chain(&ps->head[sh], link, first, last) > 0) {
            link[last] = free_head;
            free_head = first;
            return true;
        }
        if (*reinterpret_cast<volatile uint32_t *>(&ps->bump[sh]) + POOL_BATCH <= ps->limit[sh]) {
            const uint32_t o = atomicAdd(&ps->bump[sh], (uint32_t)POOL_BATCH);
            if (o + POOL_BATCH <= ps->limit[sh]) {
                for (int k = 0; k < POOL_BATCH - 1; k++) link[o + k] = o + k + 1;
                link[o + POOL_BATCH - 1] = free_head;
                free_head = o;
                return true;
            }
        }
    }
    return false;
}

// group-level chunk allocation for the groups with `need`; every lane of such a group gets the id.
// Order: the group's free list, its private range, then a batch from the shared pool.
__device__ __forceinline__ uint32_t g_chunk_alloc(bool need, GHeap &h, uint32_t *link, PoolState *ps, uint32_t gid) {
    uint32_t id = NO_CHUNK, nx = 0, how = 0;
    if (need && g_lane() == 0) {
        if (h.free_head != NO_CHUNK) {
            id = h.free_head; nx = link[id]; how = 1;
        } else if (h.bump < h.priv_hi) {
            id = h.bump; how = 2;
        } else {
            uint32_t fh = NO_CHUNK;
            if (pool_borrow(ps, link, gid, fh)) { id = fh; nx = link[id]; how = 3; }
        }
    }
    id = gshfl(id, 0);
    nx = gshfl(nx, 0);
    how = gshfl(how, 0);
    if (need) {
        if (how == 1) h.free_head = nx;
        else if (how == 2) h.bump++;
        else if (how == 3) { h.free_head = nx; h.took_shared = true; }
    }
    return id;
}

#ifdef BWB_AB_ENGINES   // round-1 A/B baseline (8 lanes per read); not in the default build
template <bool WIDE>
__global__ void __launch_bounds__(256, BWB_K4_MIN_BLOCKS) k_search_g(const __grid_constant__ SearchArgs a) {
    typedef typename Coord<WIDE>::type T;
    typedef typename Pair<T>::type P;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ T sC[17];
    stage_C<T>(a.ix, sC);

    const uint32_t gl = g_lane();
    const uint32_t gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    unsigned char *gbase = smem + (size_t)(threadIdx.x >> 3) * a.smem_per_group;
    GList<T> ls;
    ls.s = reinterpret_cast<P *>(gbase);
    ls.g = reinterpret_cast<P *>(a.glists) + (size_t)gid * 2 * a.list_cap;
    ls.cap = a.list_cap;
    uint16_t *D = reinterpret_cast<uint16_t *>(gbase + a.off_D);
    uint16_t *Ds = reinterpret_cast<uint16_t *>(gbase + a.off_Ds);
    uint8_t *sseq = gbase + a.off_seq;
    GHeap h;
    h.cnt = reinterpret_cast<uint32_t *>(gbase + a.off_bk);
    h.top = h.cnt + a.nb;
    h.bot = h.top + a.nb;
    h.bump = gid * a.chunks_per_group;
    h.priv_hi = h.bump + a.chunks_per_group;
    h.free_head = NO_CHUNK;
    h.took_shared = false;
    h.n = 0; h.best = a.nb;
    bwb_hit *stage = a.stage + (size_t)gid * a.hits_cap;
    const T lastrow = (T)(a.ix.length - 1);

    // lane l of a group serves codes l and l+8; grayVal (io.h:29) of both
    const uint32_t symA = gl, symB = gl + 8u;
    const uint32_t grayA = (uint32_t)(0x89BAEFDC45762310ull >> (4u * symA)) & 15u;
    const uint32_t grayB = (uint32_t)(0x89BAEFDC45762310ull >> (4u * symB)) & 15u;

    enum { NEED = 0, SEARCH = 1, TAIL = 2, FLUSH = 3, DONE = 4 };
    int mode = NEED;
    uint32_t r = 0, read_id = 0;
    int len = 0, err = 0;
    // search state (uniform per group)
    bool have_next = false;
    PE<T> nx;
    nx.L = 0; nx.U = 0; nx.z = 0; nx.w = 0; nx.r1 = nx.r2 = nx.r3 = 0;
    int nx_bucket = 0, best_score = 0, max_diff = 0, num_best = 0, n_hits = 0;
    // exact tail in progress
    StepState<T> st;
    st.cur = 0; st.n_cur = 0;
    step_begin(st);
    PE<T> te = nx;               // the entry whose tail is being matched
    int t_bucket = 0, t_r = 0;
    uint32_t t_c = 0;
    bool t_in_step = false;

    uint64_t c_pops = 0, c_push = 0, c_tails = 0;
    uint32_t nloads = 0, c_maxheap = 0, c_maxlist = 0;

    for (;;) {
        // ================= A: take the next read =================
        if (__any_sync(FULL, mode == NEED)) {
            uint32_t rr = 0;
            if (mode == NEED && gl == 0) rr = atomicAdd(a.queue, 1u);
            rr = gshfl(rr, 0);
            const bool got = (mode == NEED) && rr < a.n_reads;
            if (mode == NEED && !got) mode = DONE;
            uint64_t off = 0;
            if (got) {
                r = rr;
                off = a.offsets[r];
                len = (int)(a.offsets[r + 1] - off);
                read_id = a.read_id_base + r;
            }
            const uint32_t nN = g_stage_read(got, a.seq + off, len, sseq);
            g_load_bounds(got, a.d_main + off + r, len + 1, D);
            g_load_bounds(got && a.seed_len > 0, a.d_seed + (size_t)r * (a.seed_len + 1), a.seed_len + 1, Ds);
            if (got) {
                for (int b = gl; b < a.nb; b += GL) { h.cnt[b] = 0; h.top[b] = NO_CHUNK; h.bot[b] = NO_CHUNK; }
                h.n = 0; h.best = a.nb;
                n_hits = 0; err = 0;
                // root entry (inexact_match.c:281) goes straight to the "next pop" registers
                have_next = true;
                nx.L = 0; nx.U = lastrow; nx.z = (uint32_t)len; nx.w = 0; nx.r1 = nx.r2 = nx.r3 = 0;
                nx_bucket = 0;
                best_score = a.nb; max_diff = a.max_diff; num_best = 0;
                mode = ((int)nN <= a.max_diff) ? SEARCH : FLUSH;      // N pre-check, inexact_match.c:259-266
                if (mode == SEARCH) c_push++;
            }
            __syncwarp();
        }
        if (__all_sync(FULL, mode == DONE)) break;

        // ================= B: one list pass of the exact tails in progress =================
        if (__any_sync(FULL, mode == TAIL)) {
            bool on = (mode == TAIL);
            bool nomatch = false;
            if (on && !t_in_step) {                    // begin the step for rc[t_r]
                t_c = nt4_compl(sseq[len - 1 - t_r]);
                step_begin(st);
                t_in_step = true;
                if (t_c > 3u) nomatch = true;          // N in the read never matches (exact_match.c:84-87)
            }
            const bool pass_on = on && !nomatch && st.s < st.n_cur;
            if (!list_pass<T>(pass_on, a.ix, sC, ls, st, t_c, nloads)) { err = BWB_ERR_CAPACITY; nomatch = true; }
            __syncwarp();
            bool complete = false;
            if (on && !nomatch && st.s >= st.n_cur) {   // step finished
                st.cur ^= 1;
                st.n_cur = st.n_next;
                if ((uint32_t)st.n_cur > c_maxlist) c_maxlist = (uint32_t)st.n_cur;
                t_in_step = false;
                t_r--;
                if (st.n_cur == 0) nomatch = true;
                else if (t_r < 0) complete = true;
            }
            if (on && nomatch) { mode = err ? FLUSH : SEARCH; t_in_step = false; }
            // tail matched: same bookkeeping as a hit, once per interval (inexact_match.c:347-371)
            if (__any_sync(FULL, complete)) {
                const uint32_t z = te.z;
                const int ei = (int)(z & 0xffu);
                const int used = (int)((z >> 8) & 0xffu) + (int)((z >> 24) & 15u) + (int)((z >> 16) & 0xffu);
                const uint32_t go = (z >> 24) & 15u;
                bool stop = false;
                uint32_t wsum = 0;
                if (complete) {
                    if (n_hits == 0) {
                        best_score = t_bucket;
                        max_diff = (used + 1 > a.max_diff) ? a.max_diff : used + 1;
                    }
                    if (t_bucket == best_score)
                        for (int k = gl; k < st.n_cur; k += GL) {
                            const P iv = glget<T>(ls, st.cur, k);
                            wsum += (uint32_t)(iv.y - iv.x + 1);
                        }
                }
                wsum = gsum(wsum);
                if (complete) {
                    if (t_bucket == best_score) num_best = (int)((uint32_t)num_best + wsum);
                    else if (num_best > a.max_best) stop = true;
                }
                const uint32_t alen2 = ((uint32_t)(len - ei) + (te.w & 0xffu) + (uint32_t)ei) & 0xffu;
                // add the intervals 8 at a time, in list order
                int base = 0;
                bool adding = complete && !stop;
                while (__any_sync(FULL, adding && base < st.n_cur)) {
                    const bool round = adding && base < st.n_cur;
                    const int k = base + (int)gl;
                    const bool have = round && k < st.n_cur;
                    P iv;
                    iv.x = 0; iv.y = 0;
                    if (have) iv = glget<T>(ls, st.cur, k);
                    bool keep = have;
                    {                                  // with gaps: dedupe against earlier hits (align.c:273-280)
                        const int jmax = __reduce_max_sync(FULL, (round && go) ? n_hits : 0);
                        for (int q = 0; q < jmax; q++)
                            if (have && go && q < n_hits && stage[q].L == (uint64_t)iv.x && stage[q].U == (uint64_t)iv.y) keep = false;
                    }
                    const uint32_t K = gballot(keep);
                    if (round) {
                        const int nk = __popc(K);
                        if (n_hits + nk > a.hits_cap) { err = BWB_ERR_CAPACITY; adding = false; }
                        else {
                            if (keep) {
                                bwb_hit ht;
                                ht.L = (uint64_t)iv.x; ht.U = (uint64_t)iv.y; ht.score = t_bucket;
                                ht.num_mm = (uint8_t)((z >> 8) & 0xffu); ht.num_gapo = (uint8_t)go;
                                ht.num_gape = (uint8_t)((z >> 16) & 0xffu); ht.aln_length = (uint8_t)alen2;
                                ht.n_runs = (uint8_t)go; ht.pad[0] = ht.pad[1] = ht.pad[2] = 0;
                                ht.read_id = read_id;
                                const uint32_t rr4[BWB_MAX_GAP_RUNS] = {te.w >> 8, te.r1, te.r2, te.r3};
#pragma unroll
                                for (int q = 0; q < BWB_MAX_GAP_RUNS; q++) {
                                    const uint32_t v = (uint32_t)q < go ? rr4[q] : 0u;
                                    ht.runs[q].start = (uint8_t)(v & 0xffu); ht.runs[q].len = (uint8_t)((v >> 8) & 0xffu);
                                    ht.runs[q].state = (uint8_t)((v >> 16) & 0xffu); ht.runs[q].pad = 0;
                                }
                                stage[n_hits + __popc(K & ((1u << gl) - 1u))] = ht;
                            }
                            n_hits += nk;
                        }
                        base += GL;
                    }
                    __syncwarp();
                }
                if (complete) mode = (stop || err) ? FLUSH : SEARCH;
            }
            __syncwarp();
        }

        // ================= C: one pop of the bucket heap per searching group =================
        if (__any_sync(FULL, mode == SEARCH)) {
            bool on = (mode == SEARCH);
            // loop head of inexact_match (inexact_match.c:293-301)
            const int nvirt = h.n + (have_next ? 1 : 0);
            if (on) {
                if ((uint32_t)nvirt > c_maxheap) c_maxheap = (uint32_t)nvirt;
                if (nvirt == 0 || nvirt > a.max_entries) { mode = FLUSH; on = false; }
            }
            // ---- entry: from the registers or from the heap (heap_pop, inexact_match.c:594-610)
            PE<T> e = nx;
            int b = nx_bucket;
            const bool need_pop = on && !have_next;
            if (__any_sync(FULL, need_pop)) {
                const int pb = need_pop ? h.best : 0;
                const uint32_t cnt = need_pop ? h.cnt[pb] : 1u;
                const uint32_t ch = need_pop ? h.top[pb] : 0u;
                const uint32_t slot = (cnt - 1u) & 31u;
                PE<T> pe;
                if (need_pop) load_entry<WIDE>(a.chunks, ch, slot, pe);
                else pe = nx;
                __syncwarp();
                if (need_pop) {
                    e = pe; b = pb;
                    h.n--;
                    if (slot == 0u) {
                        if (gl == 0) {
                            const uint32_t prev = a.chunk_link[ch];
                            a.chunk_link[ch] = h.free_head;
                            h.top[pb] = prev;
                        }
                        h.free_head = ch;
                    }
                    if (gl == 0) h.cnt[pb] = cnt - 1u;
                }
                __syncwarp();
                // bucket drained: next non-empty one
                bool scan = need_pop && cnt == 1u;
                if (scan) h.best = a.nb;
                scan = scan && h.n != 0;
                int sb = pb + 1;
                while (__any_sync(FULL, scan)) {
                    const int q = sb + (int)gl;
                    const uint32_t m = gballot(scan && q < a.nb && h.cnt[q] != 0u);
                    if (scan) {
                        if (m) { h.best = sb + __ffs(m) - 1; scan = false; }
                        else { sb += GL; if (sb >= a.nb) scan = false; }
                    }
                }
            }
            have_next = on ? false : have_next;
            if (on) c_pops++;

            // ---- pruning (inexact_match.c:309-328)
            const uint32_t z = e.z;
            const int ei = (int)(z & 0xffu);
            const int go = (int)((z >> 24) & 15u), ge = (int)((z >> 16) & 0xffu);
            const int used = (int)((z >> 8) & 0xffu) + go + ge;
            const uint32_t state = (z >> 28) & 3u;
            const int dl = max_diff - used;
            const int dls = a.max_diff_seed - used;
            const int si = ei - (len - a.seed_len);
            if (on && (b & 0xff) > best_score + a.mm_score) { mode = FLUSH; on = false; }
            if (on) {
                if (dl < 0) on = false;
                else if (ei > 0 && dl < (int)(D[ei - 1] & 0x1ff)) on = false;
                else if (si > 0 && dls < (int)(Ds[si - 1] & 0x1ff)) on = false;
            }
            const uint32_t alen = ((uint32_t)(len - ei) + (e.w & 0xffu)) & 0xffu;

            // ---- a hit (inexact_match.c:331-344)
            const bool is_hit = on && ei == 0;
            if (__any_sync(FULL, is_hit)) {
                bool add = false;
                if (is_hit) {
                    if (n_hits == 0) {
                        best_score = b;
                        max_diff = (used + 1 > a.max_diff) ? a.max_diff : used + 1;
                    }
                    if (b == best_score) { num_best = (int)((uint32_t)num_best + (uint32_t)(e.U - e.L + 1)); add = true; }
                    else if (num_best > a.max_best) mode = FLUSH;
                    else add = true;
                }
                bool keep = add && gl == 0;
                {                                      // with gaps: dedupe against earlier hits (align.c:273-280)
                    const int jmax = __reduce_max_sync(FULL, (add && go) ? n_hits : 0);
                    for (int q = 0; q < jmax; q++)
                        if (keep && go && q < n_hits && stage[q].L == (uint64_t)e.L && stage[q].U == (uint64_t)e.U) keep = false;
                }
                if (keep) {
                    if (n_hits + 1 > a.hits_cap) err = BWB_ERR_CAPACITY;
                    else {
                        bwb_hit ht;
                        ht.L = (uint64_t)e.L; ht.U = (uint64_t)e.U; ht.score = b;
                        ht.num_mm = (uint8_t)((z >> 8) & 0xffu); ht.num_gapo = (uint8_t)go; ht.num_gape = (uint8_t)ge;
                        ht.aln_length = (uint8_t)alen; ht.n_runs = (uint8_t)go; ht.pad[0] = ht.pad[1] = ht.pad[2] = 0;
                        ht.read_id = read_id;
                        const uint32_t rr4[BWB_MAX_GAP_RUNS] = {e.w >> 8, e.r1, e.r2, e.r3};
#pragma unroll
                        for (int q = 0; q < BWB_MAX_GAP_RUNS; q++) {
                            const uint32_t v = q < go ? rr4[q] : 0u;
                            ht.runs[q].start = (uint8_t)(v & 0xffu); ht.runs[q].len = (uint8_t)((v >> 8) & 0xffu);
                            ht.runs[q].state = (uint8_t)((v >> 16) & 0xffu); ht.runs[q].pad = 0;
                        }
                        stage[n_hits] = ht;
                    }
                }
                const uint32_t kept = gballot(keep) & 1u;
                err = __shfl_sync(FULL, err, 0, GL);
                if (is_hit) {
                    n_hits += (int)kept;
                    if (err) mode = FLUSH;
                }
                __syncwarp();
            }
            if (is_hit) on = false;

            // ---- no differences left: exact tail (inexact_match.c:345-375), runs in phase B
            if (on && dl == 0) {
                c_tails++;
                te = e; t_bucket = b; t_r = ei - 1; t_in_step = false;
                st.cur = 0; st.n_cur = 1;
                if (gl == 0) glset<T>(ls, 0, 0, e.L, e.U);
                mode = TAIL;
                on = false;
            }

            // ---- expansion
            if (__any_sync(FULL, on)) {
                // the two 16-code rank gathers (inexact_match.c:377-383); lane serves codes gl, gl+8
                const T iL = on ? (T)(e.L - 1) : (T)0, iU = on ? e.U : (T)0;
                const T none = (T)~(T)0;
                const bool topL = (iL == lastrow), negL = (iL == none), topU = (iU == lastrow), negU = (iU == none);
                const T aL = (topL || negL) ? (T)0 : iL, aU = (topU || negU) ? (T)0 : iU;
                const uint4 *blkU = a.ix.blocks + (size_t)(aU >> 7) * 8;
                const uint4 *blkL = a.ix.blocks + (size_t)(aL >> 7) * 8;
                const uint32_t rU = (uint32_t)(aU & 127u), rL = (uint32_t)(aL & 127u);
                const Planes pu = load_planes(blkU);
                const BlockBits buA = match_code(pu, blkU, symA), buB = match_code(pu, blkU, symB);
                const uint32_t vUA = rank_in_block(buA, rU), vUB = rank_in_block(buB, rU);
                uint32_t vLA, vLB, fLA, fLB;
                if ((aL >> 7) == (aU >> 7)) {                 // narrow interval: one cache line serves both ends
                    vLA = rank_in_block(buA, rL); vLB = rank_in_block(buB, rL);
                    fLA = buA.m0 & 1u; fLB = buB.m0 & 1u;
                } else {
                    const Planes pl = load_planes(blkL);
                    const BlockBits blA = match_code(pl, blkL, symA), blB = match_code(pl, blkL, symB);
                    vLA = rank_in_block(blA, rL); vLB = rank_in_block(blB, rL);
                    fLA = blA.m0 & 1u; fLB = blB.m0 & 1u;
                }
                nloads += (on && gl == 0) ? 2u : 0u;
                // O_alphabet values incl. quirk Q1 (codes 5,9,11,13) and the two shortcuts (bwt.c:374-438)
                const bool qA = (0x2A20u >> symA) & 1u, qB = (0x2A20u >> symB) & 1u;
                const T cA = sC[symA], cB = sC[symB], cA1 = sC[symA + 1], cB1 = sC[symB + 1];
                const T LA = (T)((topL ? cA1 : (negL ? cA : (qA ? (T)(cA - fLA) : (T)(cA + vLA)))) + 1);
                const T LB = (T)((topL ? cB1 : (negL ? cB : (qB ? (T)(cB - fLB) : (T)(cB + vLB)))) + 1);
                const T UA = topU ? cA1 : (negU ? cA : (qA ? (T)(cA - (buA.m0 & 1u)) : (T)(cA + vUA)));
                const T UB = topU ? cB1 : (negU ? cB : (qB ? (T)(cB - (buB.m0 & 1u)) : (T)(cB + vUB)));
                const bool okA = (symA != 0u) && (LA <= UA), okB = (LB <= UB);

                // BWA heuristics (inexact_match.c:391-430)
                bool allow_diff = true, allow_mm = true;
                const int i1 = on ? ei - 1 : 1;
                if (i1 > 0) {
                    const uint32_t d1 = D[i1], d0 = D[i1 - 1];
                    if (dl - 1 < (int)(d0 & 0x1ff)) allow_diff = false;
                    else if ((int)(d1 & 0x1ff) == dl - 1 && (int)(d0 & 0x1ff) == dl - 1 && (d1 & 0x8000u)) allow_mm = false;
                }
                if (on && si - 1 > 0) {
                    const uint32_t s1 = Ds[si - 1], s0 = Ds[si - 2];
                    if (dls - 1 < (int)(s0 & 0x1ff)) allow_diff = false;
                    else if ((int)(s1 & 0x1ff) == dls - 1 && (int)(s0 & 0x1ff) == dls - 1 && (s1 & 0x8000u)) allow_mm = false;
                }
                const int gaps = go + ge;
                const bool allow_indels = !(i1 < a.no_indel_len + gaps || len - i1 < a.no_indel_len + gaps) &&
                                          !(go >= a.max_gapo && ge >= a.max_gape);
                const bool opening = (state == 0u);
                const bool gap_allowed = allow_diff && allow_indels && (opening ? (go < a.max_gapo) : (ge < a.max_gape));
                const bool full = allow_diff && allow_mm;

                // children: virtual lane v = 16*kind + code, kind 0 = gap (code 0 = insertion), 1 = match/mismatch;
                // this lane owns v = gl (q0), gl+8 (q1), 16+gl (q2), 24+gl (q3); v order = reference push order
                const uint32_t cbase = nt4_compl(sseq[on ? len - 1 - i1 : 0]);          // rc[i-1]
                const uint32_t cmask = (0x01428u >> (4u * cbase)) & 15u;                // nt4_gray_val; 0 for N
                const bool mmA = (symA == 10u) || ((cmask & grayA) == 0u);
                const bool mmB = (symB == 10u) || ((cmask & grayB) == 0u);
                const bool v0 = on && gap_allowed && (symA == 0u ? (state != 2u) : (state != 1u && okA));
                const bool v1 = on && gap_allowed && (state != 1u && okB);
                const bool v2 = on && okA && (full || !mmA);
                const bool v3 = on && okB && (full || !mmB);
                const uint32_t Vall = gballot(v0) | (gballot(v1) << 8) | (gballot(v2) << 16) | (gballot(v3) << 24);
                const uint32_t MM = (gballot(mmA) << 16) | (gballot(mmB) << 24);
                if (on) c_push += __popc(Vall);

                // packed children (see PE): gap children share everything but L,U,state
                const uint32_t zg = (z & ~(3u << 28)) + (opening ? (1u << 24) : (1u << 16));
                const uint32_t newrunI = alen | (1u << 8) | (1u << 16), newrunD = alen | (1u << 8) | (2u << 16);
                uint32_t wI = e.w, wD = e.w + 1u, r1I = e.r1, r2I = e.r2, r3I = e.r3, r1D = e.r1, r2D = e.r2, r3D = e.r3;
                if (opening) {
                    if (go == 0) { wI = (wI & 0xffu) | (newrunI << 8); wD = (wD & 0xffu) | (newrunD << 8); }
                    else if (WIDE && go == 1) { r1I = newrunI; r1D = newrunD; }
                    else if (WIDE && go == 2) { r2I = newrunI; r2D = newrunD; }
                    else if (WIDE) { r3I = newrunI; r3D = newrunD; }
                } else {
                    if (go == 1) { wI += 1u << 16; wD += 1u << 16; }
                    else if (WIDE && go == 2) { r1I += 1u << 8; r1D += 1u << 8; }
                    else if (WIDE && go == 3) { r2I += 1u << 8; r2D += 1u << 8; }
                    else if (WIDE && go == 4) { r3I += 1u << 8; r3D += 1u << 8; }
                }
                const bool insA = (symA == 0u);
                PE<T> c0, c1, c2, c3;
                c0.L = insA ? e.L : LA; c0.U = insA ? e.U : UA;
                c0.z = insA ? ((zg | (1u << 28)) - 1u) : (zg | (2u << 28));
                c0.w = insA ? wI : wD; c0.r1 = insA ? r1I : r1D; c0.r2 = insA ? r2I : r2D; c0.r3 = insA ? r3I : r3D;
                c1.L = LB; c1.U = UB; c1.z = zg | (2u << 28); c1.w = wD; c1.r1 = r1D; c1.r2 = r2D; c1.r3 = r3D;
                const uint32_t zm = (z - 1u) & ~(3u << 28);
                c2.L = LA; c2.U = UA; c2.z = zm + (mmA ? 0x100u : 0u); c2.w = e.w; c2.r1 = e.r1; c2.r2 = e.r2; c2.r3 = e.r3;
                c3.L = LB; c3.U = UB; c3.z = zm + (mmB ? 0x100u : 0u); c3.w = e.w; c3.r1 = e.r1; c3.r2 = e.r2; c3.r3 = e.r3;

                // score classes: match -> bucket b, mismatch -> b+M, gap -> b+O (open) / b+E (extend);
                // classes sharing a bucket form one push group (virtual-lane order = push order)
                uint32_t g0 = Vall & 0xffff0000u & ~MM, g1 = Vall & 0xffff0000u & MM, g2 = Vall & 0x0000ffffu;
                const int b0 = b, b1 = b + a.mm_score, b2 = b + (opening ? a.gapo_score : a.gape_score);
                if (b2 == b1) { g1 |= g2; g2 = 0; }
                if (b1 == b0) { g0 |= g1; g1 = 0; }
                if (b2 == b0) { g0 |= g2; g2 = 0; }
                // the child the next heap_pop would return stays in registers
                {
                    uint32_t km = g0;
                    int kb = b0, which = 0;
                    if (!km) {
                        if (g1 && (!g2 || b1 < b2)) { km = g1; kb = b1; which = 1; }
                        else if (g2) { km = g2; kb = b2; which = 2; }
                    }
                    const bool keepn = on && km && kb <= h.best;
                    const int kl = km ? 31 - __clz(km) : 0;
                    const int q = kl >> 3, owner = kl & 7;
                    const PE<T> &s01 = (q & 1) ? c1 : c0;
                    const PE<T> &s23 = (q & 1) ? c3 : c2;
                    const T sL = (q & 2) ? s23.L : s01.L, sU = (q & 2) ? s23.U : s01.U;
                    const uint32_t sz = (q & 2) ? s23.z : s01.z, sw = (q & 2) ? s23.w : s01.w;
                    const T kL = gshfl(sL, owner), kU = gshfl(sU, owner);
                    const uint32_t kz = gshfl(sz, owner), kw = gshfl(sw, owner);
                    uint32_t k1 = 0, k2 = 0, k3 = 0;
                    if constexpr (WIDE) {
                        k1 = gshfl((q & 2) ? s23.r1 : s01.r1, owner);
                        k2 = gshfl((q & 2) ? s23.r2 : s01.r2, owner);
                        k3 = gshfl((q & 2) ? s23.r3 : s01.r3, owner);
                    }
                    if (keepn) {
                        nx.L = kL; nx.U = kU; nx.z = kz; nx.w = kw; nx.r1 = k1; nx.r2 = k2; nx.r3 = k3;
                        nx_bucket = kb;
                        have_next = true;
                        const uint32_t bit = ~(1u << kl);
                        if (which == 0) g0 &= bit; else if (which == 1) g1 &= bit; else g2 &= bit;
                    }
                }
                // pushes (heap_push, inexact_match.c:548-591), one group of lanes per bucket
#pragma unroll
                for (int cls = 0; cls < 3; cls++) {
                    const uint32_t grp = on ? (cls == 0 ? g0 : (cls == 1 ? g1 : g2)) : 0u;
                    const int sc = cls == 0 ? b0 : (cls == 1 ? b1 : b2);
                    if (__any_sync(FULL, grp != 0u)) {
                        const bool pg = grp != 0u;
                        const uint32_t k = __popc(grp);
                        const uint32_t cnt = pg ? h.cnt[sc] : 1u;
                        const uint32_t oldtop = pg ? h.top[sc] : 0u;
                        const uint32_t topidx = (cnt - 1u) >> 5;
                        const bool need_new = pg && ((cnt == 0u) || (((cnt + k - 1u) >> 5) != topidx));
                        uint32_t newc = NO_CHUNK;
                        if (__any_sync(FULL, need_new)) {
                            newc = g_chunk_alloc(need_new, h, a.chunk_link, a.pool, gid);
                            if (need_new) {
                                if (newc == NO_CHUNK) err = BWB_ERR_CAPACITY;
                                else if (gl == 0) {
                                    a.chunk_link[newc] = oldtop;
                                    h.top[sc] = newc;
                                    if (cnt == 0u) h.bot[sc] = newc;
                                }
                            }
                        }
                        if (pg && !err) {
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                const uint32_t vb = gl + 8u * q;
                                if ((grp >> vb) & 1u) {
                                    const uint32_t pos = cnt + __popc(grp & ((1u << vb) - 1u));
                                    const bool in_old = (cnt != 0u) && ((pos >> 5) == topidx);
                                    store_entry<WIDE>(a.chunks, in_old ? oldtop : newc, pos & 31u,
                                                      q == 0 ? c0 : (q == 1 ? c1 : (q == 2 ? c2 : c3)));
                                }
                            }
                            if (gl == 0) h.cnt[sc] = cnt + k;
                            h.n += (int)k;
                            h.best = min(h.best, sc);
                        }
                        __syncwarp();
                    }
                }
                if (on && err) mode = FLUSH;
            }
            __syncwarp();
        }

        // ================= D: hand the read's hit group over (K5 restores input order) =================
        if (__any_sync(FULL, mode == FLUSH)) {
            const bool on = (mode == FLUSH);
            if (on && err) {
                if (gl == 0 && atomicCAS(a.status, 0u, (uint32_t)(-err)) == 0u) a.status[1] = read_id;
                n_hits = 0;
            }
            unsigned long long base = 0;
            if (on && gl == 0) base = atomicAdd(a.out_cursor, (unsigned long long)n_hits);
            base = gshfl((uint64_t)base, 0);
            if (on) {
                if (base + n_hits <= a.out_cap) {
                    const uint4 *src = reinterpret_cast<const uint4 *>(stage);
                    uint4 *dst = reinterpret_cast<uint4 *>(a.out_hits + base);
                    for (int k = gl; k < n_hits * 3; k += GL) dst[k] = src[k];
                }
                if (gl == 0) { a.read_off[r] = base; a.read_cnt[r] = (uint32_t)n_hits; }
                // give the chunks still held by buckets back to the group's free list; after a read that
                // borrowed from the shared pool, hand every shared chunk back (walk of the free list)
                uint32_t fh = h.free_head;
                if (gl == 0) {
                    for (int b = 0; b < a.nb; b++)
                        if (h.cnt[b]) { a.chunk_link[h.bot[b]] = fh; fh = h.top[b]; }
                    if (h.took_shared) {
                        uint32_t keep = NO_CHUNK, cur = fh, sfirst = NO_CHUNK, slast = NO_CHUNK;
                        while (cur != NO_CHUNK) {
                            const uint32_t nxt = a.chunk_link[cur];
                            if (cur >= a.priv_total) {
                                if (sfirst == NO_CHUNK) slast = cur;
                                else a.chunk_link[cur] = sfirst;
                                sfirst = cur;
                            } else { a.chunk_link[cur] = keep; keep = cur; }
                            cur = nxt;
                        }
                        if (sfirst != NO_CHUNK) pool_push_chain(&a.pool->head[gid % POOL_SHARDS], a.chunk_link, sfirst, slast);
                        fh = keep;
                    }
                }
                h.free_head = fh;
                h.took_shared = false;
            }
            h.free_head = gshfl(h.free_head, 0);
            if (on) { have_next = false; mode = NEED; }
            __syncwarp();
        }
    }

    const uint32_t tot = gsum(nloads);
    if (gl == 0) {
        atomicAdd(a.counters + 0, (unsigned long long)c_pops);
        atomicAdd(a.counters + 1, (unsigned long long)c_push);
        atomicAdd(a.counters + 2, (unsigned long long)c_tails);
        atomicAdd(a.counters + 3, (unsigned long long)tot);
        atomicMax(a.counters + 4, (unsigned long long)c_maxheap);
        atomicMax(a.counters + 5, (unsigned long long)c_maxlist);
    }
}

#endif  // BWB_AB_ENGINES

}  // namespace bwb

// bwb_kernels.cuh -- the sm_100a kernels of the BWBBLE hot path (included by bwb_abi.cu).
//
//   K0 k_relayout        .bwt arrays -> 128-B index blocks                     (bwt.c:90-125 product)
//   K1 k_occ / k_occ_alphabet / k_occ_bench   rank primitives, gather bench    (bwt.c:348-438)
//   K2 k_exact           exact_match_bounded from the full range, warp / read  (exact_match.c:66-119)
//   K3 k_calc_d          calculate_d, warp / read                              (inexact_match.c:171-254)
//   K4 k_align           calculate_d + inexact_match, persistent warp / read   (inexact_match.c:25-168,256-610)
//   K5 k_emit            per-read hit groups -> one array in input order       (inexact_match.c:152-164)
//
// No tensor cores: nothing on this path is a dense contraction (integer rank + popcount only).
#pragma once
#include "bwb_device.cuh"
#include "bwbble_b200.h"

namespace bwb {

// ---------------------------------------------------------------------------------------------
// K0: one thread per 128-row block
// ---------------------------------------------------------------------------------------------
__global__ void k_relayout(const uint32_t *__restrict__ bwt, uint64_t num_words, const uint64_t *__restrict__ O,
                           uint64_t num_occ, uint64_t sa0_index, uint32_t *__restrict__ blocks,
                           uint64_t num_blocks, uint32_t *__restrict__ err) {
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= num_blocks) return;
    uint32_t plane[4][4];
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int w = 0; w < 4; w++) plane[k][w] = 0;
    uint32_t first = 0;
#pragma unroll
    for (int wi = 0; wi < 16; wi++) {
        const uint64_t widx = b * 16 + wi;
        const uint32_t w = widx < num_words ? bwt[widx] : 0u;
        if (wi == 0) first = w >> 28;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t sym = (w >> (28 - 4 * j)) & 15u;
            const int bit = 8 * (wi & 3) + j;
#pragma unroll
            for (int k = 0; k < 4; k++) plane[k][wi >> 2] |= ((sym >> k) & 1u) << bit;
        }
    }
    uint32_t *out = blocks + b * 32;
    for (uint32_t c = 0; c < 16; c++) {
        // reference checkpoint row is inclusive of row 128*b; make it exclusive.  The sentinel row
        // holds nibble 0 but is never counted in O[.][0] (bwt.c:284-286).
        uint64_t v = b < num_occ ? O[b * 16 + c] : 0ull;
        if (c == first && !(c == 0 && b * 128 == sa0_index)) v -= 1;
        if (v >> 32) atomicExch(err, 1u);
        out[c] = (uint32_t)v;
    }
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int w = 0; w < 4; w++) out[16 + 4 * k + w] = plane[k][w];
}

// ---------------------------------------------------------------------------------------------
// K1: rank primitives (tests + the Occ-gather micro-benchmark).  T = SA-coordinate type.
// ---------------------------------------------------------------------------------------------
template <class T>
__global__ void k_occ(IndexView ix, const uint8_t *__restrict__ code, const uint64_t *__restrict__ pos, uint64_t n,
                      uint64_t *__restrict__ out) {
    __shared__ T sC[17];
    stage_C<T>(ix, sC);
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const uint64_t p = pos[q];
    out[q] = (uint64_t)occ1<T>(ix, sC, code[q] & 15u, p == ~0ull ? (T)~(T)0 : (T)p);
}

// 16 lanes per query, lane j computes occ[j]; results widened the way the reference's u64 wraps
template <class T>
__global__ void k_occ_alphabet(IndexView ix, const uint64_t *__restrict__ pos, uint64_t n, uint32_t inc,
                               uint64_t *__restrict__ out) {
    __shared__ T sC[17];
    stage_C<T>(ix, sC);
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t q = t >> 4;
    const uint32_t j = (uint32_t)(t & 15u);
    if (q >= n) return;
    const uint64_t p = pos[q];
    out[q * 16 + j] = j ? (uint64_t)occ_alpha<T>(ix, sC, j, p == ~0ull ? (T)~(T)0 : (T)p, inc) : 0ull;
}

__device__ __forceinline__ uint64_t mix64(uint64_t x) {   // splitmix64 finaliser
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// Occ-gather micro-benchmark: uniform random positions over the whole index.
// MODE 0: one thread = one O(c,i) query.  MODE 1: 16 lanes = one O_alphabet(i) query.
// `chain` dependent queries per thread (next position derived from the previous result) model the
// search's dependent gathers; chain=1 is the pure throughput case.
template <int MODE>
__global__ void k_occ_bench(IndexView ix, uint64_t n, uint64_t seed, int chain, unsigned long long *__restrict__ sink) {
    __shared__ uint64_t sC[17];
    stage_C<uint64_t>(ix, sC);
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t q = MODE ? (t >> 4) : t;
    if (q >= n) return;
    uint64_t h = mix64(seed ^ (q * 0xD6E8FEB86659FD93ull));
    uint64_t acc = 0;
    for (int s = 0; s < chain; s++) {
        const uint64_t pos = h % (ix.length - 1);
        uint64_t v;
        if (MODE) v = occ_alpha<uint64_t>(ix, sC, (uint32_t)(t & 15u), pos, 0);
        else v = occ1<uint64_t>(ix, sC, 1u + (uint32_t)((h >> 58) % 15u), pos);
        acc += v;
        h = mix64(h ^ (MODE ? __shfl_sync(FULL, (uint32_t)v, (threadIdx.x & 16u) + 7) : v));
    }
    if (acc == 0x123456789ABCDEFull) atomicAdd(sink, acc);      // keeps the loads alive
    if ((t & 0xFFFFF) == 0) atomicAdd(sink, acc);
}

// ---------------------------------------------------------------------------------------------
// shared helpers for the warp-per-read kernels
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t nt4_compl(uint32_t c) { return c > 3u ? 4u : 3u - c; }

// exact_match_bounded(read, i0, L, U) (exact_match.c:66-119) on the warp's lists.
// use_rc: read base r is the complement of seq[len-1-r] (the search runs on the reverse complement).
// Returns list length (0 = no match, -1 = list capacity exceeded); `cur` says which list holds it.
template <class T>
__device__ __forceinline__ int exact_from(const IndexView &ix, const T *sC, const ListStore<T> &ls,
                                          const uint8_t *seq, int len, bool use_rc, int i0, T L, T U, int &cur,
                                          uint32_t &nloads, uint32_t &maxlist) {
    if (lane_id() == 0) lset<T>(ls, 0, 0, L, U);
    __syncwarp();
    cur = 0;
    int n = 1;
    for (int r = i0; r >= 0; r--) {
        const uint32_t c = use_rc ? nt4_compl(seq[len - 1 - r]) : (uint32_t)seq[r];
        if (c > 3u) return 0;                        // N in the read never matches
        uint32_t sumw;
        n = extend_step<T>(ix, sC, ls, cur, n, c, sumw, nloads);
        if (n < 0) return -1;
        cur ^= 1;
        if ((uint32_t)n > maxlist) maxlist = (uint32_t)n;
        if (n == 0) return 0;
    }
    return n;
}

// calculate_d (inexact_match.c:208-254) on seq[0..dlen): D[k] = {num_diff, sa_intv_width}.
template <class T>
__device__ __forceinline__ bool calc_d(const IndexView &ix, const T *sC, const ListStore<T> &ls, const uint8_t *seq,
                                       int dlen, int2 *D, uint32_t &nloads, uint32_t &maxlist) {
    const uint32_t lane = lane_id();
    const T fullU = (T)(ix.length - 1);
    if (lane == 0) lset<T>(ls, 0, 0, (T)0, fullU);
    __syncwarp();
    int cur = 0, n = 1, z = 0;
    for (int i = dlen - 1; i >= 0; i--) {
        const uint32_t c = seq[i];
        uint32_t num = 0;
        int nn = 0;
        if (c <= 3u) {
            nn = extend_step<T>(ix, sC, ls, cur, n, c, num, nloads);
            if (nn < 0) return false;
            cur ^= 1;
            if ((uint32_t)nn > maxlist) maxlist = (uint32_t)nn;
        }
        if (nn == 0) {                               // restart from the full range, one more difference
            if (lane == 0) lset<T>(ls, cur, 0, (T)0, fullU);
            __syncwarp();
            nn = 1;
            z++;
            num = (uint32_t)ix.length;               // (int)(U-L+1) of the outer L,U (inexact_match.c:243)
        }
        n = nn;
        if (lane == 0) D[dlen - 1 - i] = make_int2(z, (int)num);
    }
    if (lane == 0) D[dlen] = make_int2(z + 1, 0);
    __syncwarp();
    return true;
}

// stage one read's bases into shared memory (codes > 4 clamp to 4 = N)
__device__ __forceinline__ uint32_t stage_read(const uint8_t *__restrict__ g, int len, uint8_t *s) {
    uint32_t nN = 0;
    for (int k = lane_id(); k < len; k += 32) {
        uint8_t c = g[k];
        if (c > 4) c = 4;
        s[k] = c;
        nN += (c == 4);
    }
    __syncwarp();
    return __reduce_add_sync(FULL, nN);
}

// ---------------------------------------------------------------------------------------------
// K2 / K3: parity kernels, one warp per read
// ---------------------------------------------------------------------------------------------
struct ListArgs {
    IndexView ix;
    const uint8_t *seq;
    const uint64_t *offsets;
    uint32_t n_reads;
    void *glists;            // [n_warps][2][list_cap] of Pair<T> (allocated for 16-byte pairs)
    int list_cap;
    int max_len;
    int use_len;             // K3: prefix length (0 = whole read)
    // K2 outputs
    ulonglong2 *out_iv;
    unsigned long long out_cap;
    unsigned long long *out_cursor;
    unsigned long long *read_off;
    uint32_t *read_cnt;
    // K3 output
    int32_t *out_d;
    uint32_t *status;
};

constexpr int LIST_SMEM_BYTES = 2 * SL * (int)sizeof(ulonglong2);   // sized for the 64-bit pairs

template <class T>
__device__ __forceinline__ void warp_lists(unsigned char *wbase, void *glists, uint32_t gw, int cap, ListStore<T> &ls) {
    typedef typename Pair<T>::type P;
    ls.s = reinterpret_cast<P *>(wbase);
    ls.g = reinterpret_cast<P *>(glists) + (size_t)gw * 2 * cap;
    ls.cap = cap;
}

template <class T>
__global__ void k_exact(ListArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ T sC[17];
    stage_C<T>(a.ix, sC);
    const int wpb = blockDim.x >> 5;
    const int per_warp = LIST_SMEM_BYTES + ((a.max_len + 15) & ~15);
    const uint32_t gw = blockIdx.x * wpb + (threadIdx.x >> 5);
    const uint32_t nw = gridDim.x * wpb;
    unsigned char *wbase = smem + (size_t)(threadIdx.x >> 5) * per_warp;
    ListStore<T> ls;
    warp_lists<T>(wbase, a.glists, gw, a.list_cap, ls);
    uint8_t *sseq = wbase + LIST_SMEM_BYTES;
    const uint32_t lane = lane_id();
    for (uint32_t r = gw; r < a.n_reads; r += nw) {
        const uint64_t off = a.offsets[r];
        const int len = (int)(a.offsets[r + 1] - off);
        stage_read(a.seq + off, len, sseq);
        int cur;
        uint32_t nl = 0, ml = 0;
        int n = exact_from<T>(a.ix, sC, ls, sseq, len, false, len - 1, (T)0, (T)(a.ix.length - 1), cur, nl, ml);
        if (n < 0) { if (lane == 0) atomicExch(a.status, (uint32_t)(-BWB_ERR_CAPACITY)); n = 0; }
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(a.out_cursor, (unsigned long long)n);
        base = shfl64(base, 0);
        if (base + n <= a.out_cap)
            for (int k = lane; k < n; k += 32) {
                const typename Pair<T>::type iv = lget<T>(ls, cur, k);
                a.out_iv[base + k] = make_ulonglong2((unsigned long long)iv.x, (unsigned long long)iv.y);
            }
        if (lane == 0) { a.read_off[r] = base; a.read_cnt[r] = (uint32_t)n; }
        __syncwarp();
    }
}

template <class T>
__global__ void k_calc_d(ListArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ T sC[17];
    stage_C<T>(a.ix, sC);
    const int wpb = blockDim.x >> 5;
    const int dbytes = ((a.max_len + 1) * 8 + 15) & ~15;
    const int per_warp = LIST_SMEM_BYTES + dbytes + ((a.max_len + 15) & ~15);
    const uint32_t gw = blockIdx.x * wpb + (threadIdx.x >> 5);
    const uint32_t nw = gridDim.x * wpb;
    unsigned char *wbase = smem + (size_t)(threadIdx.x >> 5) * per_warp;
    ListStore<T> ls;
    warp_lists<T>(wbase, a.glists, gw, a.list_cap, ls);
    int2 *D = reinterpret_cast<int2 *>(wbase + LIST_SMEM_BYTES);
    uint8_t *sseq = wbase + LIST_SMEM_BYTES + dbytes;
    const uint32_t lane = lane_id();
    for (uint32_t r = gw; r < a.n_reads; r += nw) {
        const uint64_t off = a.offsets[r];
        const int len = (int)(a.offsets[r + 1] - off);
        const int dlen = (a.use_len > 0 && a.use_len < len) ? a.use_len : len;
        stage_read(a.seq + off, len, sseq);
        uint32_t nl = 0, ml = 0;
        if (!calc_d<T>(a.ix, sC, ls, sseq, dlen, D, nl, ml)) {
            if (lane == 0) atomicExch(a.status, (uint32_t)(-BWB_ERR_CAPACITY));
            continue;
        }
        int32_t *o = a.out_d + 2 * (off + r);
        for (int k = lane; k <= dlen; k += 32) { o[2 * k] = D[k].x; o[2 * k + 1] = D[k].y; }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// K4: calculate_d + inexact_match, persistent warp per read
// ---------------------------------------------------------------------------------------------
// Partial alignment (aln_entry_t, align.h:100-119) kept PACKED, in registers and in the heap.  The
// reference's 256-byte edit path is all STATE_M except for <= max_gapo runs of I/D, so only the runs
// are kept; aln_length = readLen - i + (#D steps); the score is the bucket index; num_snps is never
// read by a decision nor serialised; the number of runs equals num_gapo.
//   z = i | mm<<8 | ge<<16 | go<<24 (4 bits) | state<<28 (2 bits)
//   w = nD | run0.start<<8 | run0.len<<16 | run0.state<<24
//   r1..r3 = runs 1..3 as start | len<<8 | state<<16                      (WIDE format only)
// Compact format (T=uint32_t, max_gapo<=1): one uint4 {L, U, z, w}             = 16 bytes
// Wide format    (T=uint64_t or max_gapo>1): {Llo, Ulo, Lhi|Uhi<<8, z} {w,r1,r2,r3} = 32 bytes
constexpr int CHUNK_ENTRIES = 32;
constexpr uint32_t NO_CHUNK = 0xffffffffu;

struct AlignArgs {
    IndexView ix;
    const uint8_t *seq;
    const uint64_t *offsets;
    uint32_t n_reads;
    uint32_t read_id_base;
    // aln_params_t
    int max_diff, max_gapo, max_gape, max_entries, mm_score, gapo_score, gape_score;
    int seed_len, max_diff_seed, max_best, no_indel_len;
    int nb;                   // number of score buckets, heap_init (inexact_match.c:513)
    int max_len;
    uint32_t *queue;          // next read to take
    // per-warp scratch
    void *glists;
    int list_cap;
    uint4 *chunks;            // [n_chunks][CHUNK_ENTRIES] entries of 16 or 32 bytes
    uint32_t *chunk_link;     // [n_chunks]
    uint32_t chunks_per_warp;
    uint32_t n_chunks;
    uint32_t *overflow_cursor;   // starts at n_warps*chunks_per_warp
    bwb_hit *stage;
    int hits_cap;
    // outputs (unordered groups; K5 orders them)
    bwb_hit *out_hits;
    unsigned long long out_cap;
    unsigned long long *out_cursor;
    unsigned long long *read_off;
    uint32_t *read_cnt;
    uint32_t *status;            // [0] error code, [1] read id
    unsigned long long *counters;   // [0] pops [1] pushes [2] exact tails [3] rank queries [4] max heap [5] max list
    // shared-memory layout (bytes per warp)
    int smem_per_warp, off_D, off_Ds, off_bk, off_seq;
};

template <class T>
struct PE {                   // packed entry
    T L, U;
    uint32_t z, w, r1, r2, r3;
};

template <bool WIDE> struct Coord { typedef uint32_t type; };
template <> struct Coord<true> { typedef uint64_t type; };

template <bool WIDE>
__device__ __forceinline__ void store_entry(uint4 *chunks, uint32_t ch, uint32_t slot, const PE<typename Coord<WIDE>::type> &e) {
    if constexpr (WIDE) {
        uint4 *dst = chunks + ((size_t)ch * CHUNK_ENTRIES + slot) * 2;
        dst[0] = make_uint4((uint32_t)e.L, (uint32_t)e.U,
                            ((uint32_t)(e.L >> 32) & 0xffu) | (((uint32_t)(e.U >> 32) & 0xffu) << 8), e.z);
        dst[1] = make_uint4(e.w, e.r1, e.r2, e.r3);
    } else {
        chunks[(size_t)ch * CHUNK_ENTRIES + slot] = make_uint4(e.L, e.U, e.z, e.w);
    }
}
template <bool WIDE>
__device__ __forceinline__ void load_entry(const uint4 *chunks, uint32_t ch, uint32_t slot, PE<typename Coord<WIDE>::type> &e) {
    if constexpr (WIDE) {
        const uint4 *src = chunks + ((size_t)ch * CHUNK_ENTRIES + slot) * 2;
        const uint4 a = src[0], b = src[1];
        e.L = (uint64_t)a.x | ((uint64_t)(a.z & 0xffu) << 32);
        e.U = (uint64_t)a.y | ((uint64_t)((a.z >> 8) & 0xffu) << 32);
        e.z = a.w; e.w = b.x; e.r1 = b.y; e.r2 = b.z; e.r3 = b.w;
    } else {
        const uint4 a = chunks[(size_t)ch * CHUNK_ENTRIES + slot];
        e.L = a.x; e.U = a.y; e.z = a.z; e.w = a.w; e.r1 = e.r2 = e.r3 = 0;
    }
}

// per-warp bucket heap (priority_heap_t, inexact_match.h:16-34): bucket = LIFO stack of chunks
struct Heap {
    uint32_t *cnt, *top, *bot;    // shared memory, nb each
    uint4 *chunks;
    uint32_t *link;
    uint32_t priv_hi;             // end of this warp's private chunk range
    uint32_t bump;                // next never-used private chunk
    uint32_t free_head;           // recycled chunks (linked through `link`)
    uint32_t *overflow_cursor;
    uint32_t n_chunks;
    int nb;
    int n;                        // entries in all buckets
    int best;                     // lowest non-empty bucket (nb if none)
};

// all lanes call; returns the same chunk id on every lane (NO_CHUNK when the pool is exhausted)
__device__ __forceinline__ uint32_t chunk_alloc(Heap &h) {
    uint32_t id = NO_CHUNK, nx = 0;
    if (lane_id() == 0) {
        if (h.free_head != NO_CHUNK) {
            id = h.free_head;
            nx = h.link[id];
        } else if (h.bump < h.priv_hi) {
            id = h.bump;
        } else {
            const uint32_t o = atomicAdd(h.overflow_cursor, 1u);
            id = o < h.n_chunks ? o : NO_CHUNK;
        }
    }
    id = __shfl_sync(FULL, id, 0);
    if (h.free_head != NO_CHUNK) h.free_head = __shfl_sync(FULL, nx, 0);     // uniform branch
    else if (h.bump < h.priv_hi) h.bump++;
    return id;
}

__device__ __forceinline__ void heap_reset(Heap &h) {
    for (int b = lane_id(); b < h.nb; b += 32) { h.cnt[b] = 0; h.top[b] = NO_CHUNK; h.bot[b] = NO_CHUNK; }
    h.n = 0;
    h.best = h.nb;
    __syncwarp();
}

// give every chunk still held by a bucket back to the warp's free list (O(1) per bucket)
__device__ __forceinline__ void heap_release(Heap &h) {
    __syncwarp();
    uint32_t fh = h.free_head;
    if (lane_id() == 0) {
        for (int b = 0; b < h.nb; b++) {
            if (h.cnt[b]) {
                h.link[h.bot[b]] = fh;
                fh = h.top[b];
            }
        }
    }
    h.free_head = __shfl_sync(FULL, fh, 0);
    __syncwarp();
}

// Push the entries of the lanes in `grp` (all with score `sc`) in lane order (heap_push,
// inexact_match.c:548-591).  Returns false if no chunk could be allocated.
template <bool WIDE>
__device__ __forceinline__ bool heap_push_group(Heap &h, uint32_t grp, int sc, const PE<typename Coord<WIDE>::type> &e) {
    const uint32_t lane = lane_id();
    const uint32_t k = __popc(grp);
    const uint32_t cnt = h.cnt[sc];
    const uint32_t oldtop = h.top[sc];
    const uint32_t topidx = (cnt - 1u) >> 5;                   // chunk ordinal of the current top (cnt>0)
    const bool need_new = (cnt == 0u) || (((cnt + k - 1u) >> 5) != topidx);
    uint32_t newc = NO_CHUNK;
    if (need_new) {
        newc = chunk_alloc(h);
        if (newc == NO_CHUNK) return false;
        if (lane == 0) {
            h.link[newc] = oldtop;
            h.top[sc] = newc;
            if (cnt == 0u) h.bot[sc] = newc;
        }
    }
    if ((grp >> lane) & 1u) {
        const uint32_t pos = cnt + __popc(grp & ((1u << lane) - 1u));
        const bool in_old = (cnt != 0u) && ((pos >> 5) == topidx);
        store_entry<WIDE>(h.chunks, in_old ? oldtop : newc, pos & 31u, e);
    }
    if (lane == 0) h.cnt[sc] = cnt + k;
    h.n += (int)k;
    h.best = min(h.best, sc);
    __syncwarp();
    return true;
}

// heap_pop (inexact_match.c:594-610): last entry of the lowest non-empty bucket; every lane gets it
template <bool WIDE>
__device__ __forceinline__ int heap_pop(Heap &h, PE<typename Coord<WIDE>::type> &e) {
    const uint32_t lane = lane_id();
    const int b = h.best;
    const uint32_t cnt = h.cnt[b];
    const uint32_t ch = h.top[b];
    const uint32_t slot = (cnt - 1u) & 31u;
    load_entry<WIDE>(h.chunks, ch, slot, e);
    __syncwarp();
    h.n--;
    if (slot == 0u) {                                 // chunk is empty now: recycle it
        if (lane == 0) {
            const uint32_t prev = h.link[ch];
            h.link[ch] = h.free_head;
            h.top[b] = prev;
        }
        h.free_head = ch;
    }
    if (lane == 0) h.cnt[b] = cnt - 1u;
    __syncwarp();
    if (cnt == 1u) {                                  // bucket drained: find the next non-empty one
        int nbst = h.nb;
        if (h.n) {
            for (int s = b + 1; s < h.nb; s += 32) {
                const int q = s + (int)lane;
                const uint32_t m = __ballot_sync(FULL, q < h.nb && h.cnt[q] != 0u);
                if (m) { nbst = s + __ffs(m) - 1; break; }
            }
        }
        h.best = nbst;
    }
    return b;
}

static_assert(sizeof(bwb_hit) == 48, "bwb_hit must be 3 x 16 bytes");

struct HitSink {
    bwb_hit *stage;
    int cap;
    int n;
};

// add_alignment (align.c:271-298) for intervals held one per lane (`have` lanes), in lane order.
// With gaps, an interval equal to an already recorded hit is dropped (align.c:273-280).
template <class T>
__device__ __forceinline__ bool add_hits(HitSink &hs, const PE<T> &e, int score, uint32_t alen, uint32_t read_id,
                                         bool have, T L, T U) {
    const uint32_t lane = lane_id();
    const uint32_t go = (e.z >> 24) & 15u;
    bool keep = have;
    if (go) {
        for (int j = 0; j < hs.n; j++) {
            const uint64_t hl = hs.stage[j].L, hu = hs.stage[j].U;
            if (hl == (uint64_t)L && hu == (uint64_t)U) keep = false;
        }
    }
    const uint32_t K = __ballot_sync(FULL, keep);
    const int nk = __popc(K);
    if (hs.n + nk > hs.cap) return false;
    if (keep) {
        bwb_hit h;
        h.L = (uint64_t)L; h.U = (uint64_t)U; h.score = score;
        h.num_mm = (uint8_t)((e.z >> 8) & 0xffu); h.num_gapo = (uint8_t)go; h.num_gape = (uint8_t)((e.z >> 16) & 0xffu);
        h.aln_length = (uint8_t)alen;
        h.n_runs = (uint8_t)go; h.pad[0] = h.pad[1] = h.pad[2] = 0;
        h.read_id = read_id;
        const uint32_t rr[BWB_MAX_GAP_RUNS] = {e.w >> 8, e.r1, e.r2, e.r3};
#pragma unroll
        for (int r = 0; r < BWB_MAX_GAP_RUNS; r++) {
            const uint32_t v = (uint32_t)r < go ? rr[r] : 0u;
            h.runs[r].start = (uint8_t)(v & 0xffu);
            h.runs[r].len = (uint8_t)((v >> 8) & 0xffu);
            h.runs[r].state = (uint8_t)((v >> 16) & 0xffu);
            h.runs[r].pad = 0;
        }
        hs.stage[hs.n + __popc(K & ((1u << lane) - 1u))] = h;
    }
    hs.n += nk;
    __syncwarp();
    return true;
}

#ifndef BWB_K4_MIN_BLOCKS
#define BWB_K4_MIN_BLOCKS 3
#endif

#ifdef BWB_AB_ENGINES   // round-1 A/B baseline (warp per read); not in the default build
template <bool WIDE>
__global__ void __launch_bounds__(256, BWB_K4_MIN_BLOCKS) k_align(const __grid_constant__ AlignArgs a) {
    typedef typename Coord<WIDE>::type T;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ T sC[17];
    stage_C<T>(a.ix, sC);

    const uint32_t lane = lane_id();
    const int wpb = blockDim.x >> 5;
    const uint32_t gw = blockIdx.x * wpb + (threadIdx.x >> 5);
    unsigned char *wbase = smem + (size_t)(threadIdx.x >> 5) * a.smem_per_warp;

    ListStore<T> ls;
    warp_lists<T>(wbase, a.glists, gw, a.list_cap, ls);
    int2 *D = reinterpret_cast<int2 *>(wbase + a.off_D);
    int2 *Ds = reinterpret_cast<int2 *>(wbase + a.off_Ds);
    uint8_t *sseq = wbase + a.off_seq;

    Heap h;
    h.cnt = reinterpret_cast<uint32_t *>(wbase + a.off_bk);
    h.top = h.cnt + a.nb;
    h.bot = h.top + a.nb;
    h.chunks = a.chunks;
    h.link = a.chunk_link;
    h.bump = gw * a.chunks_per_warp;
    h.priv_hi = h.bump + a.chunks_per_warp;
    h.free_head = NO_CHUNK;
    h.overflow_cursor = a.overflow_cursor;
    h.n_chunks = a.n_chunks;
    h.nb = a.nb;

    HitSink hs;
    hs.stage = a.stage + (size_t)gw * a.hits_cap;
    hs.cap = a.hits_cap;

    // lane roles in an expansion: lanes 0..15 = rank side L-1 / indel children, 16..31 = side U /
    // match+mismatch children, symbol j = lane & 15
    const uint32_t j = lane & 15u;
    const bool hi = lane >= 16u;
    const uint32_t gray_j = (uint32_t)(0x89BAEFDC45762310ull >> (4u * j)) & 15u;   // grayVal, io.h:29
    const T lastrow = (T)(a.ix.length - 1);

    uint64_t c_pops = 0, c_push = 0, c_tails = 0, c_rank = 0;
    uint32_t c_maxheap = 0, c_maxlist = 0;

    for (;;) {
        uint32_t r = 0;
        if (lane == 0) r = atomicAdd(a.queue, 1u);
        r = __shfl_sync(FULL, r, 0);
        if (r >= a.n_reads) break;
        const uint64_t off = a.offsets[r];
        const int len = (int)(a.offsets[r + 1] - off);
        const uint32_t read_id = a.read_id_base + r;
        const uint32_t nN = stage_read(a.seq + off, len, sseq);
        hs.n = 0;
        int err = 0;
        uint32_t nloads = 0;

        // lower bounds (inexact_match.c:61-64 / 140-143)
        if (!calc_d<T>(a.ix, sC, ls, sseq, len, D, nloads, c_maxlist)) err = BWB_ERR_CAPACITY;
        if (!err && a.seed_len > 0) {
            if (len > a.seed_len) {
                if (!calc_d<T>(a.ix, sC, ls, sseq, a.seed_len, Ds, nloads, c_maxlist)) err = BWB_ERR_CAPACITY;
            } else {
                // Q6: the reference consults a stale per-thread D_seed here; the defined behaviour
                // of this implementation is the freshly calloc'ed one (all zero).
                for (int k = lane; k <= a.seed_len; k += 32) Ds[k] = make_int2(0, 0);
                __syncwarp();
            }
        }

        if (!err && (int)nN <= a.max_diff) {
            heap_reset(h);
            // The entry that the next heap_pop would return is kept in registers whenever it is one
            // of the children just generated (it would be pushed last into the lowest non-empty
            // bucket and popped straight back); the root (inexact_match.c:281) starts there.
            bool have_next = true;
            PE<T> nx;
            nx.L = 0; nx.U = lastrow; nx.z = (uint32_t)len; nx.w = 0; nx.r1 = nx.r2 = nx.r3 = 0;
            int nx_bucket = 0;
            c_push++;
            int best_score = a.nb;            // aln_score(max_diff+1, max_gapo+1, max_gape+1)
            int max_diff = a.max_diff;
            int num_best = 0;

            while (!err) {
                const int nvirt = h.n + (have_next ? 1 : 0);       // heap->num_entries of the reference
                if (nvirt == 0) break;
                if ((uint32_t)nvirt > c_maxheap) c_maxheap = (uint32_t)nvirt;
                if (nvirt > a.max_entries) break;
                PE<T> e;
                int b;
                if (have_next) { e = nx; b = nx_bucket; have_next = false; }
                else b = heap_pop<WIDE>(h, e);
                c_pops++;
                if ((b & 0xff) > best_score + a.mm_score) break;           // 8-bit score field (Q4)
                const uint32_t z = e.z;
                const int ei = (int)(z & 0xffu);
                const int go = (int)((z >> 24) & 15u), ge = (int)((z >> 16) & 0xffu);
                const int used = (int)((z >> 8) & 0xffu) + go + ge;
                const uint32_t state = (z >> 28) & 3u;
                const int dl = max_diff - used;
                if (dl < 0) continue;
                if (ei > 0 && dl < D[ei - 1].x) continue;
                const int dls = a.max_diff_seed - used;
                const int si = ei - (len - a.seed_len);
                if (si > 0 && dls < Ds[si - 1].x) continue;
                const uint32_t alen = ((uint32_t)(len - ei) + (e.w & 0xffu)) & 0xffu;

                if (ei == 0) {                                          // a hit (inexact_match.c:331-344)
                    if (hs.n == 0) {
                        best_score = b;
                        max_diff = (used + 1 > a.max_diff) ? a.max_diff : used + 1;
                    }
                    if (b == best_score) num_best = (int)((uint32_t)num_best + (uint32_t)(e.U - e.L + 1));
                    else if (num_best > a.max_best) break;
                    if (!add_hits<T>(hs, e, b, alen, read_id, lane == 0, e.L, e.U)) err = BWB_ERR_CAPACITY;
                    continue;
                }
                if (dl == 0) {                                          // exact tail (inexact_match.c:345-375)
                    c_tails++;
                    int cur;
                    const int n = exact_from<T>(a.ix, sC, ls, sseq, len, true, ei - 1, e.L, e.U, cur, nloads, c_maxlist);
                    if (n < 0) { err = BWB_ERR_CAPACITY; break; }
                    if (n > 0) {
                        if (hs.n == 0) {
                            best_score = b;
                            max_diff = (used + 1 > a.max_diff) ? a.max_diff : used + 1;
                        }
                        if (b == best_score) {
                            uint32_t wsum = 0;
                            for (int k = lane; k < n; k += 32) {
                                const typename Pair<T>::type iv = lget<T>(ls, cur, k);
                                wsum += (uint32_t)(iv.y - iv.x + 1);
                            }
                            num_best = (int)((uint32_t)num_best + __reduce_add_sync(FULL, wsum));
                        } else if (num_best > a.max_best) break;
                        const uint32_t alen2 = (alen + (uint32_t)ei) & 0xffu;   // rest of the path is M
                        for (int base = 0; base < n && !err; base += 32) {
                            const int k = base + (int)lane;
                            const bool have = k < n;
                            typename Pair<T>::type iv;
                            iv.x = 0; iv.y = 0;
                            if (have) iv = lget<T>(ls, cur, k);
                            if (!add_hits<T>(hs, e, b, alen2, read_id, have, iv.x, iv.y)) err = BWB_ERR_CAPACITY;
                        }
                    }
                    continue;
                }

                // ---- expansion: the two 16-code rank gathers (inexact_match.c:377-383) ----
                const T pos = hi ? e.U : (T)(e.L - 1);
                const T mine = occ_alpha<T>(a.ix, sC, j, pos, hi ? 0u : 1u);
                const T other = shfl_xor(mine, 16);
                const T Lj = hi ? other : mine;
                const T Uj = hi ? mine : other;
                const bool ok = (j != 0u) && (Lj <= Uj);
                nloads += (j == 0u) ? 1u : 0u;

                // BWA heuristics (inexact_match.c:391-430)
                bool allow_diff = true, allow_mm = true;
                const int i1 = ei - 1;
                if (i1 > 0) {
                    const int2 d1 = D[i1], d0 = D[i1 - 1];
                    if (dl - 1 < d0.x) allow_diff = false;
                    else if (d1.x == dl - 1 && d0.x == dl - 1 && d1.y == d0.y) allow_mm = false;
                }
                if (si - 1 > 0) {
                    const int2 s1 = Ds[si - 1], s0 = Ds[si - 2];
                    if (dls - 1 < s0.x) allow_diff = false;
                    else if (s1.x == dls - 1 && s0.x == dls - 1 && s1.y == s0.y) allow_mm = false;
                }
                const int gaps = go + ge;
                const bool allow_indels = !(i1 < a.no_indel_len + gaps || len - i1 < a.no_indel_len + gaps) &&
                                          !(go >= a.max_gapo && ge >= a.max_gape);
                const bool opening = (state == 0u);
                const bool gap_allowed = allow_diff && allow_indels && (opening ? (go < a.max_gapo) : (ge < a.max_gape));
                const bool full = allow_diff && allow_mm;

                // ---- children, one per lane, lane order = reference push order (:433-504) ----
                const uint32_t c = nt4_compl(sseq[len - 1 - i1]);        // rc[i-1]
                const uint32_t cmask = (0x01428u >> (4u * c)) & 15u;     // nt4_gray_val; 0 for N
                const bool is_mm = (j == 10u) || ((cmask & gray_j) == 0u);
                const bool isD = (j != 0u);
                // lanes < 16: lane 0 = insertion (not from state D), lanes 1..15 = deletion of code j
                // (not from state I); lanes >= 16: match / mismatch with code j
                const bool gap_lane = isD ? (state != 1u && ok) : (state != 2u);
                const bool valid = hi ? (ok && (full || !is_mm)) : (gap_allowed && gap_lane);

                PE<T> ch;
                ch.r1 = e.r1; ch.r2 = e.r2; ch.r3 = e.r3;
                if (hi) {
                    ch.L = Lj; ch.U = Uj;
                    ch.z = ((z - 1u) & ~(3u << 28)) + (is_mm ? 0x100u : 0u);
                    ch.w = e.w;
                } else {
                    const uint32_t st = isD ? 2u : 1u;
                    ch.L = isD ? Lj : e.L;
                    ch.U = isD ? Uj : e.U;
                    ch.z = ((z & ~(3u << 28)) | (st << 28)) - (isD ? 0u : 1u) + (opening ? (1u << 24) : (1u << 16));
                    uint32_t cw = e.w + (isD ? 1u : 0u);
                    const uint32_t newrun = alen | (1u << 8) | (st << 16);
                    if (opening) {
                        if (go == 0) cw = (cw & 0xffu) | (newrun << 8);
                        else if (WIDE && go == 1) ch.r1 = newrun;
                        else if (WIDE && go == 2) ch.r2 = newrun;
                        else if (WIDE) ch.r3 = newrun;
                    } else {
                        if (go == 1) cw += 1u << 16;
                        else if (WIDE && go == 2) ch.r1 += 1u << 8;
                        else if (WIDE && go == 3) ch.r2 += 1u << 8;
                        else if (WIDE && go == 4) ch.r3 += 1u << 8;
                    }
                    ch.w = cw;
                }

                // score classes: match -> bucket b, mismatch -> b+M, gap -> b+O (open) or b+E (extend)
                const uint32_t Vall = __ballot_sync(FULL, valid);
                const uint32_t MM = __ballot_sync(FULL, is_mm);
                c_push += __popc(Vall);
                uint32_t g0 = Vall & 0xffff0000u & ~MM, g1 = Vall & 0xffff0000u & MM, g2 = Vall & 0x0000ffffu;
                const int b0 = b, b1 = b + a.mm_score, b2 = b + (opening ? a.gapo_score : a.gape_score);
                // classes that share a bucket are one push group (lane order = push order)
                if (b2 == b1) { g1 |= g2; g2 = 0; }
                if (b1 == b0) { g0 |= g1; g1 = 0; }
                if (b2 == b0) { g0 |= g2; g2 = 0; }
                // next pop = last child of the lowest child bucket, if that bucket is <= the heap's best
                {
                    uint32_t km = g0;
                    int kb = b0, which = 0;
                    if (!km) {
                        if (g1 && (!g2 || b1 < b2)) { km = g1; kb = b1; which = 1; }
                        else if (g2) { km = g2; kb = b2; which = 2; }
                    }
                    if (km && kb <= h.best) {
                        const int kl = 31 - __clz(km);
                        nx.L = shfl(ch.L, kl); nx.U = shfl(ch.U, kl);
                        nx.z = shfl(ch.z, kl); nx.w = shfl(ch.w, kl);
                        if constexpr (WIDE) { nx.r1 = shfl(ch.r1, kl); nx.r2 = shfl(ch.r2, kl); nx.r3 = shfl(ch.r3, kl); }
                        nx_bucket = kb;
                        have_next = true;
                        const uint32_t bit = ~(1u << kl);
                        if (which == 0) g0 &= bit; else if (which == 1) g1 &= bit; else g2 &= bit;
                    }
                }
                if (g0 && !heap_push_group<WIDE>(h, g0, b0, ch)) { err = BWB_ERR_CAPACITY; break; }
                if (g1 && !heap_push_group<WIDE>(h, g1, b1, ch)) { err = BWB_ERR_CAPACITY; break; }
                if (g2 && !heap_push_group<WIDE>(h, g2, b2, ch)) { err = BWB_ERR_CAPACITY; break; }
            }
            heap_release(h);
        }

        // ---- hand the read's hit group over (unordered; K5 restores input order) ----
        if (err) {
            if (lane == 0 && atomicCAS(a.status, 0u, (uint32_t)(-err)) == 0u) a.status[1] = read_id;
            hs.n = 0;
        }
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(a.out_cursor, (unsigned long long)hs.n);
        base = shfl64(base, 0);
        if (base + hs.n <= a.out_cap) {
            const uint4 *src = reinterpret_cast<const uint4 *>(hs.stage);
            uint4 *dst = reinterpret_cast<uint4 *>(a.out_hits + base);
            for (int k = lane; k < hs.n * 3; k += 32) dst[k] = src[k];
        }
        if (lane == 0) { a.read_off[r] = base; a.read_cnt[r] = (uint32_t)hs.n; }
        c_rank += __reduce_add_sync(FULL, nloads);
        __syncwarp();
    }

    if (lane == 0) {
        atomicAdd(a.counters + 0, (unsigned long long)c_pops);
        atomicAdd(a.counters + 1, (unsigned long long)c_push);
        atomicAdd(a.counters + 2, (unsigned long long)c_tails);
        atomicAdd(a.counters + 3, (unsigned long long)c_rank);
        atomicMax(a.counters + 4, (unsigned long long)c_maxheap);
        atomicMax(a.counters + 5, (unsigned long long)c_maxlist);
    }
}

#endif  // BWB_AB_ENGINES

// ---------------------------------------------------------------------------------------------
// K5: ordered emit.  ordered_off = exclusive scan of read_cnt, then a gather in input order.
// ---------------------------------------------------------------------------------------------
// single-block exclusive scan (n is a few million at most; <1 ms), out has n+1 entries
__global__ void __launch_bounds__(1024) k_scan_counts(const uint32_t *__restrict__ cnt, uint32_t n,
                                                      unsigned long long *__restrict__ out) {
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long carry_s;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 4096) {
        const uint32_t i0 = base + tid * 4;
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = (i0 + k < n) ? cnt[i0 + k] : 0u;
        unsigned long long mine = (unsigned long long)v[0] + v[1] + v[2] + v[3];
        unsigned long long inc = mine;                                   // inclusive warp scan
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = shfl64(inc, (int)lane - o >= 0 ? (int)lane - o : 0);
            if ((int)lane - o >= 0) inc += t;
        }
        if (lane == 31) warp_tot[wid] = inc;
        __syncthreads();
        unsigned long long wbase = 0;
        for (uint32_t w = 0; w < wid; w++) wbase += warp_tot[w];
        const unsigned long long carry = carry_s;
        unsigned long long run = carry + wbase + inc - mine;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (i0 + k < n) out[i0 + k] = run;
            run += v[k];
        }
        __syncthreads();
        if (tid == 1023) carry_s = carry + wbase + inc;
        __syncthreads();
    }
    if (tid == 0) out[n] = carry_s;
}

__global__ void k_emit(const bwb_hit *__restrict__ unordered, const unsigned long long *__restrict__ read_off,
                       const uint32_t *__restrict__ read_cnt, const unsigned long long *__restrict__ ordered_off,
                       uint32_t n_reads, bwb_hit *__restrict__ ordered, const unsigned long long *__restrict__ out_cursor,
                       unsigned long long out_cap) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    // K4 counted more hits than `unordered` / `ordered` hold: nothing was stored for the reads that did not fit,
    // the host regrows the buffers and runs the shard again -- do not touch memory on this attempt
    if (*out_cursor > out_cap) return;
    const uint32_t n = read_cnt[r];
    const uint4 *src = reinterpret_cast<const uint4 *>(unordered + read_off[r]);
    uint4 *dst = reinterpret_cast<uint4 *>(ordered + ordered_off[r]);
    for (uint32_t k = 0; k < n * 3; k++) dst[k] = src[k];
}

// ---------------------------------------------------------------------------------------------
// K0b: k-mer table for the top of the backward-search tree ("design levers").
// After every (re)start from the full range the next k steps of calculate_d depend only on the next
// k read bases.  For every k-mer X (k = KTAB, first applied base in the low digits) this kernel runs
// the ordinary multi-interval extension (same extend_step as everywhere else) and records
//   w[level s][X mod 4^s] = n_s << 33 | (n_s != 0) << 32 | (wrapped sum of widths after s steps)
// for every prefix (all k-mers sharing a prefix write the same value), and the interval list after
// k steps.  K3 then replaces up to k list steps by k look-ups -- same D arrays, bit for bit.
// ---------------------------------------------------------------------------------------------
constexpr int KTAB = 10;
__host__ __device__ __forceinline__ uint32_t ktab_level_off(int s) { return ((1u << (2 * s)) - 4u) / 3u; }   // sum_{t<s} 4^t

struct KtabArgs {
    IndexView ix;
    void *glists;                 // warp scratch [n_warps][2][list_cap]
    int list_cap;
    unsigned long long *w;        // all levels, level s at ktab_level_off(s)
    uint32_t *koff, *kcnt;        // list of every k-mer in `iv`
    void *iv;                     // Pair<T> pool
    unsigned long long iv_cap;
    unsigned long long *cursor;
    uint32_t *status;
};

template <class T>
__global__ void k_kmer_table(KtabArgs a) {
    typedef typename Pair<T>::type P;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ T sC[17];
    stage_C<T>(a.ix, sC);
    const int wpb = blockDim.x >> 5;
    const uint32_t gw = blockIdx.x * wpb + (threadIdx.x >> 5);
    const uint32_t nw = gridDim.x * wpb;
    unsigned char *wbase = smem + (size_t)(threadIdx.x >> 5) * LIST_SMEM_BYTES;
    ListStore<T> ls;
    warp_lists<T>(wbase, a.glists, gw, a.list_cap, ls);
    const uint32_t lane = lane_id();
    for (uint32_t X = gw; X < (1u << (2 * KTAB)); X += nw) {
        if (lane == 0) lset<T>(ls, 0, 0, (T)0, (T)(a.ix.length - 1));
        __syncwarp();
        int cur = 0, n = 1;
        for (int s = 1; s <= KTAB; s++) {
            const uint32_t c = (X >> (2 * (s - 1))) & 3u;
            uint32_t sumw = 0, nl = 0;
            n = extend_step<T>(a.ix, sC, ls, cur, n, c, sumw, nl);
            if (n < 0) { if (lane == 0) atomicExch(a.status, (uint32_t)(-BWB_ERR_CAPACITY)); n = 0; }
            cur ^= 1;
            if (lane == 0)
                a.w[ktab_level_off(s) + (X & ((1u << (2 * s)) - 1u))] =
                    n ? (((unsigned long long)n << 33) | (1ull << 32) | sumw) : 0ull;
            if (n == 0) break;
        }
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(a.cursor, (unsigned long long)n);
        base = shfl64(base, 0);
        if (base + n <= a.iv_cap)
            for (int k = lane; k < n; k += 32) reinterpret_cast<P *>(a.iv)[base + k] = lget<T>(ls, cur, k);
        if (lane == 0) { a.koff[X] = (uint32_t)base; a.kcnt[X] = (uint32_t)n; }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// K0c: the -P seed table (precalc_sa_intervals, align.c:200-224): exact_match() of every 12-mer from
// the full range.  Row X holds the 12-mer whose base-4 digits are X, most significant digit = first
// base (next_read, align.c:188-198); the search consumes the LAST base first, i.e. the low digits.
// Multi-genome: one warp per row (extend_step).  -S: one thread per row, 12 single-code steps
// (exact_match_1to1_bounded, exact_match.c:196-222).  Lists land in a pool in arrival order;
// off[X]/cnt[X] find them (the .pre writer walks X in order).
// ---------------------------------------------------------------------------------------------
constexpr int PRECALC_LEN = 12;                       // PRECALC_INTERVAL_LENGTH, align.h:31
constexpr uint32_t NUM_PRECALC = 1u << (2 * PRECALC_LEN);   // align.h:30

struct PrecalcArgs {
    IndexView ix;
    void *glists;                 // warp scratch [n_warps][2][list_cap]
    int list_cap;
    int is_multiref;
    uint32_t *off, *cnt;          // per row
    ulonglong2 *iv;               // (L,U) pool, 64-bit whatever the index width: K4's entry width
                                  // also depends on max_gapo
    unsigned long long iv_cap;
    unsigned long long *cursor;
    uint32_t *status;
};

template <class T>
__global__ void k_precalc(PrecalcArgs a) {
    typedef typename Pair<T>::type P;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ T sC[17];
    stage_C<T>(a.ix, sC);
    const uint32_t lane = lane_id();
    if (!a.is_multiref) {
        const uint32_t nt = gridDim.x * blockDim.x;
        for (uint32_t X0 = blockIdx.x * blockDim.x; X0 < NUM_PRECALC; X0 += nt) {     // warp-uniform trip count
            const uint32_t X = X0 + threadIdx.x;
            T L = 0, U = (T)(a.ix.length - 1);
            bool ok = true;
            for (int s = 0; s < PRECALC_LEN && ok; s++) {
                const uint32_t code = (0x173Fu >> (4u * ((X >> (2 * s)) & 3u))) & 15u;    // nt4_gray, io.h:94-110
                T oL, oU;
                occ_pair<T>(a.ix, sC, code, (T)(L - 1), U, oL, oU);
                L = (T)(sC[code] + oL + 1); U = (T)(sC[code] + oU);
                ok = L <= U;
            }
            const uint32_t V = __ballot_sync(FULL, ok);
            unsigned long long base = 0;
            if (lane == 0 && V) base = atomicAdd(a.cursor, (unsigned long long)__popc(V));
            base = shfl64(base, 0) + __popc(V & ((1u << lane) - 1u));
            if (ok && base < a.iv_cap) a.iv[base] = make_ulonglong2((unsigned long long)L, (unsigned long long)U);
            a.off[X] = (uint32_t)base; a.cnt[X] = ok ? 1u : 0u;
        }
        return;
    }
    const int wpb = blockDim.x >> 5;
    const uint32_t gw = blockIdx.x * wpb + (threadIdx.x >> 5);
    const uint32_t nw = gridDim.x * wpb;
    unsigned char *wbase = smem + (size_t)(threadIdx.x >> 5) * LIST_SMEM_BYTES;
    ListStore<T> ls;
    warp_lists<T>(wbase, a.glists, gw, a.list_cap, ls);
    for (uint32_t X = gw; X < NUM_PRECALC; X += nw) {
        if (lane == 0) lset<T>(ls, 0, 0, (T)0, (T)(a.ix.length - 1));
        __syncwarp();
        int cur = 0, n = 1;
        for (int s = 0; s < PRECALC_LEN && n; s++) {
            uint32_t sumw = 0, nl = 0;
            n = extend_step<T>(a.ix, sC, ls, cur, n, (X >> (2 * s)) & 3u, sumw, nl);
            if (n < 0) { if (lane == 0) atomicExch(a.status, (uint32_t)(-BWB_ERR_CAPACITY)); n = 0; }
            cur ^= 1;
        }
        unsigned long long base = 0;
        if (lane == 0 && n) base = atomicAdd(a.cursor, (unsigned long long)n);
        base = shfl64(base, 0);
        if (base + n <= a.iv_cap)
            for (int k = lane; k < n; k += 32) {
                const P v = lget<T>(ls, cur, k);
                a.iv[base + k] = make_ulonglong2((unsigned long long)v.x, (unsigned long long)v.y);
            }
        if (lane == 0) { a.off[X] = (uint32_t)base; a.cnt[X] = (uint32_t)n; }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// K6: SA locate of every read's first hit + the top1/top2 sums of eval_aln (align.c:760-812)
// ---------------------------------------------------------------------------------------------
// invPsi(i) = C[B(i)] + O(B(i), i), 0 for the sentinel row (bwt.c:311-317); O(0, i) does not count
// the sentinel row although it stores nibble 0 (bwt.c:362-370).
__device__ __forceinline__ uint64_t inv_psi(const IndexView &ix, uint64_t sa0, uint64_t i) {
    if (i == sa0) return 0;
    const uint4 *blk = ix.blocks + (i >> 7) * 8;
    const uint32_t p = (uint32_t)(i & 127u);
    const uint32_t *pw = reinterpret_cast<const uint32_t *>(blk) + 16 + (p >> 5);
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) c |= ((__ldg(pw + 4 * k) >> (p & 31u)) & 1u) << k;
    uint64_t o;
    if (i == ix.length - 1) {
        o = ix.C[c + 1] - ix.C[c];
    } else {
        const BlockBits b = load_block(blk, c);
        o = rank_in_block(b, p);
        if (c == 0 && (sa0 >> 7) == (i >> 7) && (sa0 & 127u) <= p) o -= 1;
    }
    return ix.C[c] + o;
}

__global__ void k_locate(IndexView ix, uint64_t sa0, const uint64_t *__restrict__ SA, const bwb_hit *__restrict__ hits,
                         const unsigned long long *__restrict__ off, const uint32_t *__restrict__ cnt, uint32_t n_reads,
                         bwb_loc *__restrict__ out, const unsigned long long *__restrict__ out_cursor,
                         unsigned long long out_cap) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    if (*out_cursor > out_cap) return;          // overflowed attempt (see k_emit): the shard is run again
    bwb_loc loc;
    loc.ref_pos = ~0ull; loc.top1 = 0; loc.top2 = 0;
    const uint32_t n = cnt[r];
    if (n) {
        const bwb_hit *h = hits + off[r];
        const int best = h[0].score;
        uint32_t t1 = 0, t2 = 0;
        for (uint32_t k = 0; k < n; k++) {
            const uint32_t w = (uint32_t)(h[k].U - h[k].L + 1);
            if (h[k].score > best) t2 += w; else t1 += w;
        }
        uint64_t i = h[0].L, j = 0;
        while (i & 31u) { i = inv_psi(ix, sa0, i); j++; }          // SA(), bwt.c:320-329
        loc.ref_pos = (SA[i >> 5] + j) % ix.length;
        loc.top1 = (int32_t)t1; loc.top2 = (int32_t)t2;
    }
    out[r] = loc;
}

}  // namespace bwb

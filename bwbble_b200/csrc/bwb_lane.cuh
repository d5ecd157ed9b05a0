// bwb_lane.cuh -- K4 "lane engine": one read per LANE, 32 independent reads per warp.
//
// Evidence for this mapping (profiles/): with a warp (k_align) or an 8-lane group (k_search_g)
// per read the kernel is instruction-issue bound (IPC 1.7/SM, DRAM < 2 % of peak) and ~95 % of the
// issued instructions are per-read *bookkeeping* (entry unpacking, pruning, BWA heuristics, bucket
// maintenance) that is uniform across the lanes serving the read: ~400 warp instructions per pop.
// Giving every lane its own read turns that bookkeeping into ordinary SIMT work over 32 reads, and
// the 2 x 15 rank computations of an expansion into a fully unrolled 15-code loop per lane
// (compile-time code => the plane XOR folds into LOP3, counters at immediate offsets).
//
// Per lane, in HBM (L1/L2 cached):
//   * one arena of 32-byte slots {L, U, z, w | next, r1, r2, r3}: heap entries, exact-tail interval
//     nodes and hit records are all singly linked slot chains (one sector per pop / push).  Slots
//     are bump-allocated and never recycled within a read -- allocation costs no memory access,
//     so the pushes of an expansion are pure stores; the cursor rewinds when the read is done.
//     A lane that outgrows its private range borrows 256-slot blocks from a sharded lock-free
//     pool and returns them at the end of the read;
//   * the bucket heap (priority_heap_t, inexact_match.h:16-34) as nb LIFO linked lists: push =
//     link in front of the bucket head, pop = unlink the head of the lowest non-empty bucket --
//     exactly "last entry of the lowest non-empty bucket" (inexact_match.c:594-610).  Bucket
//     heads live in shared memory ([bucket][lane], conflict-free), occupancy bits in registers.
//     Only the scores an entry can really have get a bucket: score = m*M + o*O + e*E with
//     m + o + e <= max_diff, o <= max_gapo, e <= max_gape and e > 0 only after an opening -- 19 of
//     the 68 buckets at the default parameters.  The host enumerates them (bucket_map, bwb_abi.cu) in
//     ascending score order, so "lowest non-empty bucket" is unchanged, the heads take 9.5 KB instead of
//     34 KB of shared memory per block (the L1 gets the difference), and the limit of 128 applies to
//     reachable scores, not to (n+1)*M + (o+1)*O + (e+1)*E.
// Each loop iteration of a lane is: [take a read] -> [pop + prune + classify] -> [one interval
// task: the 15-code rank loop, feeding either heap children or the next exact-tail list] ->
// [flush].  No warp-synchronous intrinsic is needed anywhere: lanes are independent.
#pragma once
#include "bwb_group.cuh"

namespace bwb {

constexpr uint32_t NIL = 0xffffffffu;
constexpr int LBLK = 256;                   // slots per borrowed block

struct LaneArgs {
    IndexView ix;
    const uint8_t *seq;
    const uint64_t *offsets;
    uint32_t n_reads;
    uint32_t read_id_base;
    int max_diff, max_gapo, max_gape, max_entries, mm_score, gapo_score, gape_score;
    int seed_len, max_diff_seed, max_best, no_indel_len;
    int nb;                                  // score buckets of the reference's heap (scores 0 .. nb-1)
    int nbc;                                 // buckets that can hold an entry with these parameters (<= 128), see bmap
    const uint8_t *bmap;                     // uint16 score_of[128] (compact bucket -> score), then uint8 bucket_of[nb]
                                             // (score -> compact bucket, 0xff = no entry can have this score)
    int is_multiref;                         // 0 = -S single-genome mode (codes A,G,C,T only)
    uint32_t *queue;
    const uint16_t *pk_main, *pk_seed;       // packed lower bounds from K3
    const uint16_t *n_count;                 // N bases per read, from K3
    const uint32_t *order;                   // read taken at queue position q (K3b: heavy reads first); null = q
    const uint32_t *n_queue_ptr;             // retry passes: queue length lives on the device (null: n_reads)
    uint32_t *retry_list, *retry_count;      // reads that ran out of arena are deferred to the next pass (null: fail)
    uint32_t lane_stride;                    // retry passes: only every lane_stride-th lane works, and owns the
                                             // private ranges of the idle lanes after it (power of two, 1 = all)
    const uint32_t *pre_off, *pre_cnt;       // -P seed table (K0c), rows of `pre_iv`; only read by k_search_l<.., true>
    const ulonglong2 *pre_iv;                // 64-bit (L,U) whatever T is
    uint4 *slots;                            // arena: 2 x uint4 per slot
    uint32_t slots_per_lane;                 // private range of lane-slot s: [s*spl, (s+1)*spl)
    uint32_t priv_total;                     // first slot of the shared region (multiple of LBLK)
    PoolState *pool;                         // shared region, in blocks of LBLK slots
    unsigned long long throttle_blocks;      // lanes do not start a new read while more shared blocks than this are lent out (0 = off)
    uint32_t *blk_link;                      // link word per block (pool free lists, borrowed lists)
    bwb_hit *out_hits;
    unsigned long long out_cap;
    unsigned long long *out_cursor;
    unsigned long long *read_off;
    uint32_t *read_cnt;
    uint32_t *status;
    unsigned long long *counters;
};

// ---- per-lane slot allocator: bump only ---------------------------------------------------------
struct LaneAlloc {
    uint32_t priv_lo, priv_hi, bump;
    uint32_t ov_cur, ov_end;                 // current borrowed block
    uint32_t borrowed;                       // list of borrowed blocks (through blk_link)
    uint32_t borrowed_last;
    uint32_t n_borrowed;                     // blocks on the borrowed list
};

// slow path: take a block of LBLK slots from the shared pool (own shard first); NIL if exhausted.
// Deliberately free of references to the caller's allocator state so that it stays in registers.
__device__ __noinline__ uint32_t pool_take_block(PoolState *pool, uint32_t *blk_link, uint32_t lane_slot) {
    for (int t = 0; t < POOL_SHARDS; t++) {
        const uint32_t sh = (lane_slot + (uint32_t)t) % POOL_SHARDS;
        unsigned long long *head = &pool->head[sh];
        unsigned long long old = atomicAdd(head, 0ull);
        for (;;) {
            const uint32_t id = (uint32_t)old;
            if (id == NIL) break;
            const uint32_t nx = *reinterpret_cast<volatile uint32_t *>(blk_link + id);
            const unsigned long long prev = atomicCAS(head, old, (((old >> 32) + 1ull) << 32) | nx);
            if (prev == old) { atomicAdd(&pool->n_borrowed, 1ull); return id; }
            old = prev;
        }
        if (*reinterpret_cast<volatile uint32_t *>(&pool->bump[sh]) < pool->limit[sh]) {
            const uint32_t o = atomicAdd(&pool->bump[sh], 1u);
            if (o < pool->limit[sh]) { atomicAdd(&pool->n_borrowed, 1ull); return o; }
        }
    }
    return NIL;
}

__device__ __forceinline__ uint32_t lane_alloc(LaneAlloc &al, const LaneArgs &a, uint32_t lane_slot) {
    if (al.bump < al.priv_hi) return al.bump++;
    if (al.ov_cur < al.ov_end) return al.ov_cur++;
    const uint32_t blk = pool_take_block(a.pool, a.blk_link, lane_slot);
    if (blk == NIL) return NIL;
    a.blk_link[blk] = al.borrowed;
    if (al.borrowed == NIL) al.borrowed_last = blk;
    al.borrowed = blk;
    al.n_borrowed++;
    al.ov_cur = blk * LBLK;
    al.ov_end = al.ov_cur + LBLK;
    return al.ov_cur++;
}
// n consecutive slots (n <= LBLK): the children of one expansion.  A range too short for them is left to the
// single-slot allocations (private range) or abandoned (borrowed block).
__device__ __forceinline__ uint32_t lane_alloc_n(LaneAlloc &al, const LaneArgs &a, uint32_t lane_slot, uint32_t n) {
    if (al.bump + n <= al.priv_hi) { const uint32_t s = al.bump; al.bump += n; return s; }
    if (al.ov_cur + n <= al.ov_end) { const uint32_t s = al.ov_cur; al.ov_cur += n; return s; }
    const uint32_t blk = pool_take_block(a.pool, a.blk_link, lane_slot);
    if (blk == NIL) return NIL;
    a.blk_link[blk] = al.borrowed;
    if (al.borrowed == NIL) al.borrowed_last = blk;
    al.borrowed = blk;
    al.n_borrowed++;
    al.ov_cur = blk * LBLK + n;
    al.ov_end = blk * LBLK + LBLK;
    return blk * LBLK;
}
// index of the k-th set bit (k = 0 is the lowest) of a 16-bit mask with more than k bits set
__device__ __forceinline__ int kth_bit16(uint32_t m, uint32_t k) {
    int pos = 0;
    uint32_t c = (uint32_t)__popc(m & 0xffu);
    if (k >= c) { k -= c; pos = 8; }
    c = (uint32_t)__popc((m >> pos) & 0xfu);
    if (k >= c) { k -= c; pos += 4; }
    c = (uint32_t)__popc((m >> pos) & 0x3u);
    if (k >= c) { k -= c; pos += 2; }
    c = (m >> pos) & 1u;
    if (k >= c) pos += 1;
    return pos;
}
// end of a read: every slot is dead; hand borrowed blocks back with one CAS
__device__ __forceinline__ void lane_alloc_reset(LaneAlloc &al, const LaneArgs &a, uint32_t lane_slot) {
    if (al.borrowed != NIL) {
        unsigned long long *head = &a.pool->head[lane_slot % POOL_SHARDS];
        unsigned long long old = atomicAdd(head, 0ull);
        for (;;) {
            *reinterpret_cast<volatile uint32_t *>(a.blk_link + al.borrowed_last) = (uint32_t)old;
            __threadfence();
            const unsigned long long prev = atomicCAS(head, old, (((old >> 32) + 1ull) << 32) | al.borrowed);
            if (prev == old) break;
            old = prev;
        }
        atomicAdd(&a.pool->n_borrowed, 0ull - (unsigned long long)al.n_borrowed);
    }
    al.n_borrowed = 0;
    al.bump = al.priv_lo;
    al.ov_cur = al.ov_end = 0;
    al.borrowed = NIL;
    al.borrowed_last = NIL;
}

// ---- slot payloads (2 x uint4) --------------------------------------------------------------------
//   first  = {L[31:0], U[31:0], z, w}
//   second = {next, r1, r2 | L[39:32]<<24, r3 | U[39:32]<<24}        (r1..r3 use 24 bits)
// heap entry: as is.  interval node: L, U, next.  hit: slot A = the entry, slot B.first = {score, alen}.
template <class T>
__device__ __forceinline__ void slot_write(uint4 *slots, uint32_t s, T L, T U, uint32_t z, uint32_t w, uint32_t nxt,
                                           uint32_t r1, uint32_t r2, uint32_t r3) {
    uint32_t hl = 0, hu = 0;
    if constexpr (sizeof(T) == 8) { hl = (uint32_t)(L >> 32) << 24; hu = (uint32_t)(U >> 32) << 24; }
    slots[2 * (size_t)s] = make_uint4((uint32_t)L, (uint32_t)U, z, w);
    slots[2 * (size_t)s + 1] = make_uint4(nxt, r1, (r2 & 0xffffffu) | hl, (r3 & 0xffffffu) | hu);
}
template <class T>
__device__ __forceinline__ uint32_t slot_read(const uint4 *slots, uint32_t s, PE<T> &e) {
    const uint4 a = slots[2 * (size_t)s], b = slots[2 * (size_t)s + 1];
    e.z = a.z; e.w = a.w; e.r1 = b.y; e.r2 = b.z & 0xffffffu; e.r3 = b.w & 0xffffffu;
    if constexpr (sizeof(T) == 8) {
        e.L = (uint64_t)a.x | ((uint64_t)(b.z >> 24) << 32);
        e.U = (uint64_t)a.y | ((uint64_t)(b.w >> 24) << 32);
    } else {
        e.L = a.x; e.U = a.y;
    }
    return b.x;
}
__device__ __forceinline__ void slot_set_next(uint4 *slots, uint32_t s, uint32_t nxt) {
    reinterpret_cast<uint32_t *>(slots + 2 * (size_t)s + 1)[0] = nxt;
}
__device__ __forceinline__ uint32_t slot_next(const uint4 *slots, uint32_t s) {
    return reinterpret_cast<const uint32_t *>(slots + 2 * (size_t)s + 1)[0];
}

// Bucket heap of one lane.  heads = shared memory, column of this lane ([bucket][lane]); buckets are
// COMPACT indices (reachable scores in ascending order; LaneArgs::bmap translates both ways);
// bm0..bm3 = occupancy bits of up to 128 buckets, so heads never need resetting between reads and
// "next non-empty bucket" is a find-first-set.
template <bool WIDE>
struct LaneHeap {
    typedef typename Coord<WIDE>::type T;
    uint32_t *heads;       // &sm_heads[0][lane]; bucket b at heads[b * 128]
    uint32_t bm0, bm1, bm2, bm3;
    int n;

    __device__ __forceinline__ void clear() { bm0 = bm1 = bm2 = bm3 = 0u; n = 0; }
    __device__ __forceinline__ bool occupied(int b) const {
        const uint32_t w = b < 64 ? (b < 32 ? bm0 : bm1) : (b < 96 ? bm2 : bm3);
        return (w >> (b & 31)) & 1u;
    }
    __device__ __forceinline__ void set_bit(int b) {
        const uint32_t bit = 1u << (b & 31);
        bm0 |= (b < 32) ? bit : 0u;
        bm1 |= (b >= 32 && b < 64) ? bit : 0u;
        bm2 |= (b >= 64 && b < 96) ? bit : 0u;
        bm3 |= (b >= 96) ? bit : 0u;
    }
    __device__ __forceinline__ void clear_bit(int b) {
        const uint32_t bit = 1u << (b & 31);
        bm0 &= ~((b < 32) ? bit : 0u);
        bm1 &= ~((b >= 32 && b < 64) ? bit : 0u);
        bm2 &= ~((b >= 64 && b < 96) ? bit : 0u);
        bm3 &= ~((b >= 96) ? bit : 0u);
    }
    __device__ __forceinline__ int best() const {        // lowest non-empty bucket (n > 0)
        return bm0 ? __ffs(bm0) - 1 : (bm1 ? 31 + __ffs(bm1) : (bm2 ? 63 + __ffs(bm2) : 95 + __ffs(bm3)));
    }

    // link a new entry in front of bucket sc; the caller sets the occupancy bit (mark) afterwards
    __device__ __forceinline__ bool push(LaneAlloc &al, const LaneArgs &a, uint32_t lane_slot, int sc, T L, T U,
                                         uint32_t z, uint32_t w, uint32_t r1, uint32_t r2, uint32_t r3) {
        const uint32_t s = lane_alloc(al, a, lane_slot);
        if (s == NIL) return false;
        slot_write<T>(a.slots, s, L, U, z, w, heads[sc * 128], r1, r2, r3);
        heads[sc * 128] = s;
        n++;
        return true;
    }
    __device__ __forceinline__ void mark(int b) { set_bit(b); }
    // heap_pop (inexact_match.c:594-610); returns the (compact) bucket
    __device__ __forceinline__ int pop(const LaneArgs &a, PE<T> &e, uint32_t &s) {
        const int b = best();
        s = heads[b * 128];
        const uint32_t rest = slot_read<T>(a.slots, s, e);
        heads[b * 128] = rest;
        if (rest == NIL) clear_bit(b);
        n--;
        return b;
    }
};

// ---------------------------------------------------------------------------------------------
// K3b: queue order for K4.  The cost of a read grows steeply with the whole-read lower bound
// D[len-1] that K3 just computed (every entry scoring up to best+mm_score is expanded, and the best
// score is at least D[len-1] mismatches away); reads whose bound already exceeds max_diff die at the
// root.  Handing the expensive reads out FIRST (longest-processing-time-first) lets the cheap ones
// fill the end of the launch, instead of a few lanes finishing a late heavy read while the rest of
// the GPU idles.  Class = bound (0 = trivial), counting sort by descending class; the order inside a
// class is arbitrary -- results are per read and K5 restores input order.
// ---------------------------------------------------------------------------------------------
constexpr int ORDER_CLASSES = 32;

__device__ __forceinline__ uint32_t order_class(const uint16_t *pk_main, const uint16_t *n_count, const uint64_t *offsets,
                                                uint32_t r, int max_diff) {
    const uint64_t off = offsets[r];
    const int len = (int)(offsets[r + 1] - off);
    if (len <= 0 || (int)n_count[r] > max_diff) return 0u;
    const int z = (int)(pk_main[off + r + (uint64_t)(len - 1)] & 0x1ffu);
    if (z > max_diff) return 0u;
    return (uint32_t)(1 + (z < ORDER_CLASSES - 2 ? z : ORDER_CLASSES - 2));
}

__global__ void k_order_hist(const uint16_t *__restrict__ pk_main, const uint16_t *__restrict__ n_count,
                             const uint64_t *__restrict__ offsets, uint32_t n_reads, int max_diff, uint32_t *hist) {
    __shared__ uint32_t sh[ORDER_CLASSES];
    if (threadIdx.x < ORDER_CLASSES) sh[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += gridDim.x * blockDim.x)
        atomicAdd(&sh[order_class(pk_main, n_count, offsets, r, max_diff)], 1u);
    __syncthreads();
    if (threadIdx.x < ORDER_CLASSES && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

// same grid as k_order_hist; cursor[] zeroed
__global__ void k_order_scatter(const uint16_t *__restrict__ pk_main, const uint16_t *__restrict__ n_count,
                                const uint64_t *__restrict__ offsets, uint32_t n_reads, int max_diff,
                                const uint32_t *__restrict__ hist, uint32_t *cursor, uint32_t *__restrict__ order) {
    __shared__ uint32_t start[ORDER_CLASSES], cnt[ORDER_CLASSES], base[ORDER_CLASSES];
    if (threadIdx.x < ORDER_CLASSES) {
        uint32_t s = 0;
        for (int c = ORDER_CLASSES - 1; c > (int)threadIdx.x; c--) s += hist[c];      // heavier classes come first
        start[threadIdx.x] = s;
    }
    for (uint32_t r0 = blockIdx.x * blockDim.x; r0 < n_reads; r0 += gridDim.x * blockDim.x) {     // block-uniform trip count
        if (threadIdx.x < ORDER_CLASSES) cnt[threadIdx.x] = 0;
        __syncthreads();
        const uint32_t r = r0 + threadIdx.x;
        uint32_t cls = 0, rank = 0;
        if (r < n_reads) {
            cls = order_class(pk_main, n_count, offsets, r, max_diff);
            rank = atomicAdd(&cnt[cls], 1u);
        }
        __syncthreads();
        if (threadIdx.x < ORDER_CLASSES && cnt[threadIdx.x])
            base[threadIdx.x] = start[threadIdx.x] + atomicAdd(&cursor[threadIdx.x], cnt[threadIdx.x]);
        __syncthreads();
        if (r < n_reads) order[base[cls] + rank] = r;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// K4 rank stage: (L, U, quirk mode) -> valid-child mask + (L_j, U_j) for j = 1..15, the straight-line
// 15-code loop over the two ends of the interval (2 x 15 x (2 LOP3 + 1 POPC) x 4 words).
//
// Tried in round 2 and dropped (profiles/r02_k4_exchange.md): re-dealing the rank tasks of a block BY TYPE
// through shared memory -- narrow same-block intervals (60 % of the expansions) as one cheap work item per
// row, the rest as full 32-lane batches of this function.  Bit-exact, but 2 block barriers per iteration with
// 12 warps per SM cost more issue slots (30 % busy instead of 43 %) than the cheaper rank path saved:
// 0.73-0.77 M reads/s against 1.00 M for every lane ranking its own task.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void low_bits128(int n, uint32_t &k0, uint32_t &k1, uint32_t &k2, uint32_t &k3) {
    k0 = n >= 32 ? ~0u : ((1u << n) - 1u);
    k1 = n >= 64 ? ~0u : (n <= 32 ? 0u : ((1u << (n - 32)) - 1u));
    k2 = n >= 96 ? ~0u : (n <= 64 ? 0u : ((1u << (n - 64)) - 1u));
    k3 = n >= 128 ? ~0u : (n <= 96 ? 0u : ((1u << (n - 96)) - 1u));
}

// The interval task [L, iU] of lane `o`: child intervals -> sLj[j][o], sUj[j][o] for every code j; returns the mask of
// the codes with L_j <= U_j.  trueq = "true counts for codes 5,9,11,13" (exact tails use O(), bwt.c:348-372, and -S
// never sees those codes; expansions in multi-genome mode use O_alphabet with quirk Q1, bwt.c:427-435,780).
// C[] comes straight from the kernel's constant-bank parameters (compile-time index: no load instruction).
// the upper end's index block, when it was requested before the task was known to exist (HOIST)
struct UpperBlock { Planes p; uint4 q0, q1, q2, q3; };
__device__ __forceinline__ void load_upper(UpperBlock &b, const uint4 *__restrict__ blk) {
    b.p = load_planes(blk);
    b.q0 = __ldg(blk); b.q1 = __ldg(blk + 1); b.q2 = __ldg(blk + 2); b.q3 = __ldg(blk + 3);
}

// (A/B, round 2: fetching the 32 index blocks of a warp's tasks with the WHOLE warp -- 8 lanes per block, 16 bytes each, so
// that one load instruction touches 4 cache lines instead of 32 -- and handing them to their owners through the warp's
// columns of sLj / sUj, XOR-swizzled, conflict-free: bit-exact and 17 % SLOWER, 0.85 against 1.02 M reads/s at 3 and at
// 4 blocks/SM.  The L1's tag stage is not what limits the rank stage; 16 shuffles, 16 shared-memory round trips and
// three more warp barriers per task are worse than 2 x 8 divergent 16-byte loads.  profiles/r02_ab_log.md)
template <class T, bool HOIST>
__device__ __forceinline__ uint32_t rank_general(const IndexView &ix, T (*sLj)[128], T (*sUj)[128], uint32_t o,
                                                 const T L, const T iU, const bool trueq, const T lastrow,
                                                 const bool have_ub, const UpperBlock &ub) {
    const T iL = (T)(L - 1);
    const bool negL = (L == 0), topU = (iU == lastrow);
    const T aL = negL ? (T)0 : iL, aU = topU ? (T)0 : iU;
    const uint4 *blkU = ix.blocks + (size_t)(aU >> 7) * 8;
    const uint4 *blkL = ix.blocks + (size_t)(aL >> 7) * 8;
    // The two ends one after the other: the upper end's 15 counts stay in registers while the lower end's block
    // (the same cache line for 83 % of the tasks) is loaded and matched -- half the live registers of doing both
    // ends side by side, which is what lets 4 blocks of 128 lanes share an SM.
    uint32_t vU[16];
    uint32_t firstU = 0;                          // bit j: row 0 of the upper block holds code j (Q1)
    // (A/B, round 2: prefetch.global.L1 of the lower end's four sectors here, and of the next pop's blocks when the
    // keeper child is chosen, made K4 17-22 % SLOWER on the 600 M-row index and 3-6 % slower on chr21 -- CCTL.PF1 is
    // not free and the L1 left beside 54 KB of shared memory per block is small; profiles/r02_ab_log.md)
    {
        UpperBlock mine;
        if (HOIST && have_ub) mine = ub; else load_upper(mine, blkU);
        const Planes pu = mine.p;
        const uint4 q0 = mine.q0, q1 = mine.q1, q2 = mine.q2, q3 = mine.q3;
        const uint32_t cU[16] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
        uint32_t k0, k1, k2, k3;
        low_bits128((int)(aU & 127u) + 1, k0, k1, k2, k3);
#pragma unroll
        for (int j = 1; j < 16; j++) {
            const uint32_t x0 = (j & 1) ? 0u : ~0u, x1 = (j & 2) ? 0u : ~0u, x2 = (j & 4) ? 0u : ~0u, x3 = (j & 8) ? 0u : ~0u;
            const uint32_t u0 = (pu.p0.x ^ x0) & (pu.p1.x ^ x1) & (pu.p2.x ^ x2) & (pu.p3.x ^ x3);
            const uint32_t u1 = (pu.p0.y ^ x0) & (pu.p1.y ^ x1) & (pu.p2.y ^ x2) & (pu.p3.y ^ x3);
            const uint32_t u2 = (pu.p0.z ^ x0) & (pu.p1.z ^ x1) & (pu.p2.z ^ x2) & (pu.p3.z ^ x3);
            const uint32_t u3 = (pu.p0.w ^ x0) & (pu.p1.w ^ x1) & (pu.p2.w ^ x2) & (pu.p3.w ^ x3);
            vU[j] = cU[j] + __popc(u0 & k0) + __popc(u1 & k1) + __popc(u2 & k2) + __popc(u3 & k3);
            if (j == 5 || j == 9 || j == 11 || j == 13) firstU |= (u0 & 1u) << j;
        }
    }
    const Planes pl = load_planes(blkL);
    const uint4 p0 = __ldg(blkL), p1 = __ldg(blkL + 1), p2 = __ldg(blkL + 2), p3 = __ldg(blkL + 3);
    const uint32_t cL[16] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w, p2.x, p2.y, p2.z, p2.w, p3.x, p3.y, p3.z, p3.w};
    uint32_t k0, k1, k2, k3;
    low_bits128((int)(aL & 127u) + 1, k0, k1, k2, k3);
    uint32_t okmask = 0;
#pragma unroll
    for (int j = 1; j < 16; j++) {
        const uint32_t x0 = (j & 1) ? 0u : ~0u, x1 = (j & 2) ? 0u : ~0u, x2 = (j & 4) ? 0u : ~0u, x3 = (j & 8) ? 0u : ~0u;
        const uint32_t l0 = (pl.p0.x ^ x0) & (pl.p1.x ^ x1) & (pl.p2.x ^ x2) & (pl.p3.x ^ x3);
        const uint32_t l1 = (pl.p0.y ^ x0) & (pl.p1.y ^ x1) & (pl.p2.y ^ x2) & (pl.p3.y ^ x3);
        const uint32_t l2 = (pl.p0.z ^ x0) & (pl.p1.z ^ x1) & (pl.p2.z ^ x2) & (pl.p3.z ^ x3);
        const uint32_t l3 = (pl.p0.w ^ x0) & (pl.p1.w ^ x1) & (pl.p2.w ^ x2) & (pl.p3.w ^ x3);
        const uint32_t vL = cL[j] + __popc(l0 & k0) + __popc(l1 & k1) + __popc(l2 & k2) + __popc(l3 & k3);
        // Q1: O_alphabet skips codes 5,9,11,13 except for the checkpoint-symbol decrement
        // (bwt.c:427-435,780); the exact search's O() counts them (bwt.c:348-372)
        const bool quirk = (j == 5 || j == 9 || j == 11 || j == 13);
        const T Cj = (T)ix.C[j], Cj1 = (T)ix.C[j + 1];
        T Lj, Uj;
        if (quirk) {
            const T qL = trueq ? (T)vL : (T)0 - (T)(l0 & 1u);
            const T qU = trueq ? (T)vU[j] : (T)0 - (T)((firstU >> j) & 1u);
            Lj = (T)(Cj + (negL ? (T)0 : qL) + 1);
            Uj = topU ? Cj1 : (T)(Cj + qU);
        } else {
            Lj = (T)(Cj + (negL ? (T)0 : (T)vL) + 1);
            Uj = topU ? Cj1 : (T)(Cj + (T)vU[j]);
        }
        sLj[j][o] = Lj;
        sUj[j][o] = Uj;
        okmask |= (Lj <= Uj) ? (1u << j) : 0u;
    }
    return okmask;
}

// Per warp iteration (the 32 reads of a warp advance in lock-step, one interval task each):
//   A  every lane advances its read: [take a read] -> [exact tail: level end / next interval] ->
//      [pop + prune + classify]; the outcome is at most one interval task;
//   B  rank stage: the 15-code loop over both ends of the lane's interval;
//   C  the result is consumed: next level of the exact tail (align.c:93-110), or the children of the
//      expansion (inexact_match.c:433-504) -- described per lane, then written by the whole warp, one
//      child per lane and pass; then [flush] of finished reads.
// 32-bit coordinates: 4 blocks of 128 lanes per SM at 128 registers (20-32 bytes of spills); 64-bit: 3 blocks at 168.
// (A/B on B200, chr21, 2 M reads per launch.  With 68 bucket heads per lane a block took 54 KB of shared memory, four
// of them left the SM almost no L1 and ran 5 % slower than three, 0.94 against 0.99 M reads/s; with the compact heads
// -- 30 KB per block -- four blocks are 2 % faster than three, and 4 % with the early request of the upper index
// block: 0.98 / 1.00 / 1.02 M reads/s, profiles/r02_ab_log.md.)
#ifndef BWB_FREE_RING
#define BWB_FREE_RING 8
#endif
constexpr int FREE_RING = BWB_FREE_RING;   // free slot ids kept per lane
static_assert(FREE_RING > 0 && FREE_RING < 256 && (FREE_RING & (FREE_RING - 1)) == 0, "FREE_RING: a power of two below 256");
#ifndef BWB_LANE_BLOCKS_NARROW
#define BWB_LANE_BLOCKS_NARROW 4
#endif
#ifndef BWB_HOIST_NARROW
#define BWB_HOIST_NARROW 1
#endif

#define BWB_LANE_MIN_BLOCKS(WIDE) ((WIDE) ? 3 : BWB_LANE_BLOCKS_NARROW)
template <bool WIDE, bool PRE, bool RECYCLE>
__global__ void __launch_bounds__(128, BWB_LANE_MIN_BLOCKS(WIDE)) k_search_l(const __grid_constant__ LaneArgs a) {
    typedef typename Coord<WIDE>::type T;
    __shared__ T sLj[16][128], sUj[16][128];      // child intervals of the lane's task, by code (row 0 unused)
    // Slots of popped entries are handed straight to the next children (a small per-lane stack of free slot ids):
    // the arena a read needs is its LIVE heap, not every push it ever made -- 3-4x less at genome scale, where
    // bump-only allocation (round 1) overflowed the private ranges of most lanes and sent the reads to the
    // retry passes; recently freed lines are also still in L2 when they are written again.
    // RECYCLE is chosen by the host: on for indexes beyond 2^28 rows, reads longer than 128 bases and the wide entry
    // format (there it is worth 1.8-2.4x); off at chr21 scale, where nothing overflows and the bookkeeping costs 4 %.
    __shared__ uint32_t sFree[RECYCLE ? FREE_RING : 1][128];
    __shared__ uint32_t sTe[5][128];              // z, w, r1, r2, r3 of the entry whose exact tail is in progress
                                                  // (touched when a tail starts / ends: kept out of the register file)
    extern __shared__ uint32_t sm_heads[];        // [nbc][128] bucket heads, then the two score <-> bucket tables
    uint16_t *const score_of = reinterpret_cast<uint16_t *>(sm_heads + (size_t)a.nbc * 128);   // [128]
    uint8_t *const bucket_of = reinterpret_cast<uint8_t *>(score_of + 128);                     // [nb]
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t lane_slot = blockIdx.x * blockDim.x + tid;
    const T lastrow = (T)(a.ix.length - 1);
    const bool multiref = a.is_multiref != 0;
    LaneAlloc al;
    al.priv_lo = lane_slot * a.slots_per_lane;
    al.priv_hi = al.priv_lo + a.slots_per_lane * a.lane_stride;
    {
        const uint32_t all = gridDim.x * blockDim.x * a.slots_per_lane;     // end of the private region
        if (al.priv_hi > all) al.priv_hi = all;
    }
    al.bump = al.priv_lo;
    al.ov_cur = al.ov_end = 0;
    al.borrowed = NIL;
    al.borrowed_last = NIL;
    al.n_borrowed = 0;
    LaneHeap<WIDE> h;
    h.heads = sm_heads + tid;
    h.clear();

    enum { NEED = 0, SEARCH = 1, TAIL = 2, FLUSH = 3, TADD = 4, DONE = 5 };
    int mode = (lane_slot & (a.lane_stride - 1u)) ? DONE : NEED;      // idle lanes of a retry pass never touch the queue
    const uint32_t n_queue = a.n_queue_ptr ? min(*a.n_queue_ptr, a.n_reads) : a.n_reads;
    uint32_t r = 0, read_id = 0;
    int len = 0, err = 0;
    uint32_t off = 0;                         // first base of the read (the host checks that a batch has < 2^32 - n bases)
    // pointers of the read's arrays, re-derived where they are used instead of being held in 6 registers
#define BWB_RSEQ (a.seq + off)
#define BWB_D (a.pk_main + ((size_t)off + r))
#define BWB_DS (a.pk_seed + (size_t)r * (uint32_t)(a.seed_len + 1))
    int best_score = 0, max_diff = 0, num_best = 0, n_hits = 0;
    uint32_t hit_head = NIL, hit_tail = NIL;
    // the interval task of this lane
    bool have_task = false, task_tail = false;
    PE<T> e;                                  // expansion: the popped entry; tail: its L,U hold the interval
    e.L = 0; e.U = 0; e.z = 0; e.w = 0; e.r1 = e.r2 = e.r3 = 0;
    int eb = 0;                               // bucket (= score) of e
    // The entry the next heap_pop would return is kept in registers whenever it is a child of the
    // current expansion: the parent came from the lowest bucket, so its last MATCH child (same
    // score) would be pushed on top of that bucket and popped straight back.
    bool have_next = false;
    PE<T> nx = e;
    int nx_bucket = 0;
    uint32_t t_flags = 0;                     // expansion flags, see below
    uint32_t cbase = 0;                       // read base of this step (rc[i-1])
    // exact tail in progress (its entry: sTe)
    int t_bucket = 0, t_r = 0;
    uint32_t cur_head = NIL, nx_head = NIL, nx_tail = NIL;
    int nx_n = 0;
    uint32_t nx_w = 0, old_tail = NIL;        // wrapped width sum of the level being built
    T nx_tailL = 0, nx_tailU = 0;

    // Entries pushed with a score beyond best_score + mm_score can never be expanded: the search breaks as soon
    // as one is popped (inexact_match.c:309) and best_score never grows.  They are only COUNTED (num_entries
    // feeds the max_entries check, :301), not written to the arena.
    int ghost = 0;
    uint32_t nfree = 0;                       // entries of this lane's free-slot stack
#ifdef BWB_PF_D
    uint32_t pf_d = 0, pf_s = 0, pf_b = 0;
    bool pf_ok = false;
#endif
    uint32_t c_pops = 0, c_push = 0, c_tails = 0, c_rank = 0;         // per read, added to the launch counters at its flush
    uint32_t c_maxheap = 0, c_maxlist = 0;
    score_of[tid] = reinterpret_cast<const uint16_t *>(a.bmap)[tid];
    for (int b = (int)tid; b < a.nb; b += 128) bucket_of[b] = a.bmap[256 + b];
    __syncthreads();

#define BWB_CODE_OF(t) (multiref ? (int)(t) : (int)((0x173Fu >> (4 * (t))) & 15u))
    for (;;) {
        if (__all_sync(FULL, mode == DONE)) break;
        // ================= A: take the next read =================
        // Admission control: while most of the shared arena blocks are lent out, finished lanes wait instead of
        // starting another read -- the reads in flight then find room for their heaps and end normally, instead of
        // running out, being thrown away and searched again in the 1/8-occupancy retry pass (genome-scale 150 bp
        // reads with gaps: 27 % of the reads went that way).
        bool take = mode == NEED;
        if (a.throttle_blocks && __any_sync(FULL, take)) {            // (uniform; a lane needs a read once in ~10^4 iterations)
            if (take && *reinterpret_cast<volatile unsigned long long *>(&a.pool->n_borrowed) > a.throttle_blocks) take = false;
            if (__all_sync(FULL, mode == DONE || (mode == NEED && !take))) __nanosleep(2000);      // a whole warp waiting
        }
        if (take) {
            r = atomicAdd(a.queue, 1u);
            if (r >= n_queue) { mode = DONE; take = false; }
        }
        if (take) {
            if (a.order) r = a.order[r];
            const uint64_t off64 = a.offsets[r];
            off = (uint32_t)off64;
            len = (int)(a.offsets[r + 1] - off64);
            read_id = a.read_id_base + r;
            const uint8_t *rseq = BWB_RSEQ;
            const int nN = (int)a.n_count[r];                           // counted by K3
            h.clear();
            ghost = 0;
            nfree = 0;
            n_hits = 0; hit_head = hit_tail = NIL; err = 0;
            best_score = a.nb; max_diff = a.max_diff; num_best = 0;
            for (int b = 0; b < a.nbc; b++) h.heads[b * 128] = NIL;
            have_next = false;
            mode = FLUSH;
            if (PRE) {
                // -P (inexact_match.c:50-57,269-279): the search starts from the exact-match intervals of
                // rc's last 12 bases (= the complement of the read's first 12, read2index align.c:174-186),
                // pushed in list order with 12 matches on their path; the last one pops first.
                uint32_t idx = 0, bad = len < PRECALC_LEN ? 1u : 0u;
                for (int j = 0; j < PRECALC_LEN && !bad; j++) {
                    const uint32_t c = rseq[j];
                    bad |= c > 3u ? 1u : 0u;
                    idx |= (3u - (c & 3u)) << (2 * j);
                }
                const uint32_t np = (bad || nN > a.max_diff) ? 0u : a.pre_cnt[idx];
                if (np) {
                    const ulonglong2 *src = a.pre_iv + a.pre_off[idx];
                    const uint32_t z0 = (uint32_t)(len - PRECALC_LEN);
                    bool ok = true;
                    for (uint32_t k = 0; k + 1 < np && ok; k++) {
                        const ulonglong2 v = src[k];
                        ok = h.push(al, a, lane_slot, 0, (T)v.x, (T)v.y, z0, 0u, 0u, 0u, 0u);
                    }
                    if (np > 1) h.mark(0);
                    const ulonglong2 v = src[np - 1];
                    have_next = true;
                    nx.L = (T)v.x; nx.U = (T)v.y; nx.z = z0; nx.w = 0; nx.r1 = nx.r2 = nx.r3 = 0;
                    nx_bucket = 0;
                    c_push += np;
                    mode = SEARCH;
                    if (!ok) { err = BWB_ERR_CAPACITY; mode = FLUSH; }
                }
            } else if (nN <= a.max_diff) {                            // N pre-check, inexact_match.c:259-266
                // root entry (inexact_match.c:281): straight into the next-pop registers
                have_next = true;
                nx.L = 0; nx.U = lastrow; nx.z = (uint32_t)len; nx.w = 0; nx.r1 = nx.r2 = nx.r3 = 0;
                nx_bucket = 0;
                c_push++;
                mode = SEARCH;
            }
        }

        // ================= A: exact tail -- level end, then the next interval of the level =================
        if (mode == TAIL) {
            if (cur_head == NIL) {                                      // level finished
                // (A/B, round 2: handing a level of ONE interval to the next level in registers -- node neither written
                // nor read back -- was 1 % SLOWER, 972 k against 982 k reads/s: profiles/r02_ab_log.md)
                if (nx_n) slot_write<T>(a.slots, nx_tail, nx_tailL, nx_tailU, 0u, 0u, NIL, 0u, 0u, 0u);
                if ((uint32_t)nx_n > c_maxlist) c_maxlist = (uint32_t)nx_n;
                if (nx_n == 0) {
                    mode = SEARCH;                                      // no match
                } else if (t_r > 0) {
                    t_r--;
                    cur_head = nx_head; nx_head = nx_tail = NIL; nx_n = 0; nx_w = 0;
                } else {
                    // tail matched: bookkeeping of a hit (inexact_match.c:347-362); the intervals are
                    // then added one per iteration (mode TADD)
                    const uint32_t z = sTe[0][tid];
                    const int used = (int)((z >> 8) & 0xffu) + (int)((z >> 24) & 15u) + (int)((z >> 16) & 0xffu);
                    bool stop = false;
                    if (n_hits == 0) {
                        best_score = t_bucket;
                        max_diff = (used + 1 > a.max_diff) ? a.max_diff : used + 1;
                    }
                    if (t_bucket == best_score) num_best = (int)((uint32_t)num_best + nx_w);
                    else if (num_best > a.max_best) stop = true;
                    old_tail = hit_tail;                                // dedupe only against earlier hits
                    cur_head = nx_head; nx_head = nx_tail = NIL; nx_n = 0; nx_w = 0;
                    mode = stop ? FLUSH : TADD;
                }
            }
            if (mode == TAIL) {                                         // next interval of the current level
                PE<T> nd;
                cur_head = slot_read<T>(a.slots, cur_head, nd);
                e.L = nd.L; e.U = nd.U;
                cbase = nt4_compl(BWB_RSEQ[len - 1 - t_r]);
                if (cbase > 3u) {                                       // N never matches (exact_match.c:84-87):
                    cur_head = nx_head = nx_tail = NIL; nx_n = 0;
                    mode = SEARCH;                                      // empty result
                } else {
                    have_task = true;
                    task_tail = true;
                }
            }
        } else if (mode == TADD) {                                      // add_alignment per interval (:366-370)
            if (cur_head == NIL) {
                mode = SEARCH;
            } else {
                const uint32_t q = cur_head;
                PE<T> nd;
                cur_head = slot_read<T>(a.slots, q, nd);
                const T L = nd.L, U = nd.U;
                const uint32_t z = sTe[0][tid], tw = sTe[1][tid];
                const int ei = (int)(z & 0xffu);
                const int go = (int)((z >> 24) & 15u);
                const uint32_t alen2 = ((uint32_t)(len - ei) + (tw & 0xffu) + (uint32_t)ei) & 0xffu;
                bool add = true;
                if (go && old_tail != NIL) {
                    for (uint32_t p = hit_head;;) {
                        PE<T> hh;
                        const uint32_t pb = slot_read<T>(a.slots, p, hh);
                        if (hh.L == L && hh.U == U) { add = false; break; }
                        if (pb == old_tail) break;
                        p = slot_next(a.slots, pb);
                    }
                }
                if (add) {                                              // the node becomes slot A of the hit
                    const uint32_t sb = lane_alloc(al, a, lane_slot);
                    if (sb == NIL) { err = BWB_ERR_CAPACITY; mode = FLUSH; }
                    else {
                        slot_write<T>(a.slots, q, L, U, z, tw, sb, sTe[2][tid], sTe[3][tid], sTe[4][tid]);
                        slot_write<T>(a.slots, sb, (T)t_bucket, (T)alen2, 0u, 0u, NIL, 0u, 0u, 0u);
                        if (hit_tail == NIL) hit_head = q; else slot_set_next(a.slots, hit_tail, q);
                        hit_tail = sb;
                        n_hits++;
                    }
                }
            }
        }

        // ================= A: pop + prune + classify (inexact_match.c:293-375) =================
        // 32-bit coordinates only (A/B on B200: +2.5 % at 4 blocks/SM, nothing at 3; the 64-bit kernel has no registers to spare)
        constexpr bool HOIST = !WIDE && BWB_HOIST_NARROW;
        UpperBlock ub;
        bool have_ub = false;
        if (HOIST) ub.p.p0 = ub.p.p1 = ub.p.p2 = ub.p.p3 = ub.q0 = ub.q1 = ub.q2 = ub.q3 = make_uint4(0u, 0u, 0u, 0u);
        if (mode == SEARCH) {
            const int nvirt = h.n + ghost + (have_next ? 1 : 0);      // heap->num_entries of the reference
            if ((uint32_t)nvirt > c_maxheap) c_maxheap = (uint32_t)nvirt;
            if (nvirt == 0 || nvirt > a.max_entries) {
                mode = FLUSH;
            } else if (!have_next && h.n == 0) {                      // only dead entries left: pop one, break (:309)
                c_pops++;
                mode = FLUSH;
            } else {
#ifdef BWB_PF_D
                const bool from_keeper = have_next && pf_ok;
#endif
                if (have_next) { e = nx; eb = nx_bucket; have_next = false; }
                else {
                    uint32_t freed;
                    eb = (int)score_of[h.pop(a, e, freed)];
                    if (RECYCLE && nfree < (uint32_t)FREE_RING) sFree[nfree++][tid] = freed;   // nothing reads a slot after it is unlinked
                }
                c_pops++;
                if (HOIST) {
                    // The popped entry's interval is the task of this iteration unless the entry is pruned or a hit: its
                    // upper block is requested NOW, so that the round trip overlaps the one of the lower-bound arrays
                    // below (heap slot -> bounds -> index block was three dependent latencies, now two).
                    load_upper(ub, a.ix.blocks + (size_t)((e.U == lastrow ? (T)0 : e.U) >> 7) * 8);
                    have_ub = true;
                }
                const uint32_t z = e.z;
                const int ei = (int)(z & 0xffu);
                const int go = (int)((z >> 24) & 15u), ge = (int)((z >> 16) & 0xffu);
                const int used = (int)((z >> 8) & 0xffu) + go + ge;
                const uint32_t state = (z >> 28) & 3u;
                const int dl = max_diff - used;
                const int dls = a.max_diff_seed - used;
                const int si = ei - (len - a.seed_len);
                // everything this entry may need from the read's arrays, fetched in one go (one latency)
                const uint16_t *D = BWB_D, *Ds = BWB_DS;
#ifdef BWB_PF_D
                // the keeper child (i - 1) needs D[i-2], D[i-3], D_seed[si-2], D_seed[si-3], seq[len-i+1]: two of them are
                // this entry's own, the other three are fetched now, a whole iteration before they are looked at
                uint32_t dA, dB, sA, sB, base_i1;
                if (from_keeper) {
                    dA = pf_d & 0xffffu; dB = pf_d >> 16; sA = pf_s & 0xffffu; sB = pf_s >> 16; base_i1 = pf_b;
                } else {
                    dA = D[ei > 0 ? ei - 1 : 0]; dB = D[ei > 1 ? ei - 2 : 0];
                    sA = Ds[si > 0 ? si - 1 : 0]; sB = Ds[si > 1 ? si - 2 : 0];
                    base_i1 = BWB_RSEQ[ei > 0 ? len - ei : 0];
                }
                {
                    const uint32_t dC = D[ei > 2 ? ei - 3 : 0], sCn = Ds[si > 2 ? si - 3 : 0];
                    pf_b = BWB_RSEQ[ei > 1 ? len - ei + 1 : 0];
                    pf_d = dB | (dC << 16);
                    pf_s = sB | (sCn << 16);
                    pf_ok = false;                                      // set when a keeper is chosen below
                }
#else
                const uint32_t dA = D[ei > 0 ? ei - 1 : 0];                 // D[i-1]
                const uint32_t dB = D[ei > 1 ? ei - 2 : 0];                 // D[i-2]
                const uint32_t sA = Ds[si > 0 ? si - 1 : 0];                // D_seed[si-1]
                const uint32_t sB = Ds[si > 1 ? si - 2 : 0];                // D_seed[si-2]
                const uint32_t base_i1 = BWB_RSEQ[ei > 0 ? len - ei : 0];   // seq[len-1-(i-1)]
#endif
                if (eb > best_score + a.mm_score) {
                    mode = FLUSH;                                       // inexact_match.c:309
                } else if (dl < 0 || (ei > 0 && dl < (int)(dA & 0x1ff)) || (si > 0 && dls < (int)(sA & 0x1ff))) {
                    // pruned
                } else if (ei == 0) {                                   // a hit (inexact_match.c:331-344)
                    bool add = true;
                    if (n_hits == 0) {
                        best_score = eb;
                        max_diff = (used + 1 > a.max_diff) ? a.max_diff : used + 1;
                    }
                    if (eb == best_score) num_best = (int)((uint32_t)num_best + (uint32_t)(e.U - e.L + 1));
                    else if (num_best > a.max_best) { mode = FLUSH; add = false; }
                    if (add && go) {                                    // align.c:273-280
                        for (uint32_t q = hit_head; q != NIL;) {
                            PE<T> hh;
                            const uint32_t qb = slot_read<T>(a.slots, q, hh);
                            if (hh.L == e.L && hh.U == e.U) { add = false; break; }
                            q = slot_next(a.slots, qb);
                        }
                    }
                    if (add) {
                        const uint32_t alen = ((uint32_t)(len - ei) + (e.w & 0xffu)) & 0xffu;
                        const uint32_t sa = lane_alloc(al, a, lane_slot), sb = lane_alloc(al, a, lane_slot);
                        if (sa == NIL || sb == NIL) { err = BWB_ERR_CAPACITY; mode = FLUSH; }
                        else {
                            slot_write<T>(a.slots, sa, e.L, e.U, z, e.w, sb, e.r1, e.r2, e.r3);
                            slot_write<T>(a.slots, sb, (T)eb, (T)alen, 0u, 0u, NIL, 0u, 0u, 0u);
                            if (hit_tail == NIL) hit_head = sa; else slot_set_next(a.slots, hit_tail, sa);
                            hit_tail = sb;
                            n_hits++;
                        }
                    }
                } else if (dl == 0) {                                   // exact tail (inexact_match.c:345-375)
                    // its first level is the entry's own interval: the task of this iteration
                    c_tails++;
                    sTe[0][tid] = e.z; sTe[1][tid] = e.w; sTe[2][tid] = e.r1; sTe[3][tid] = e.r2; sTe[4][tid] = e.r3;
                    t_bucket = eb; t_r = ei - 1;
                    cur_head = nx_head = nx_tail = NIL; nx_n = 0; nx_w = 0;
                    const uint32_t cb = nt4_compl(base_i1);             // rc[i-1]
                    if (cb <= 3u) {                                     // (N never matches: empty result, exact_match.c:84-87)
                        cbase = cb;
                        have_task = true;
                        task_tail = true;
                        mode = TAIL;
                    }
                } else {
                    // ---- expansion: BWA heuristics (inexact_match.c:391-430) -> flags for the task
                    bool allow_diff = true, allow_mm = true;
                    const int i1 = ei - 1;
                    if (i1 > 0) {
                        const uint32_t d1 = dA, d0 = dB;
                        if (dl - 1 < (int)(d0 & 0x1ff)) allow_diff = false;
                        else if ((int)(d1 & 0x1ff) == dl - 1 && (int)(d0 & 0x1ff) == dl - 1 && (d1 & 0x8000u)) allow_mm = false;
                    }
                    if (si - 1 > 0) {
                        const uint32_t s1 = sA, s0 = sB;
                        if (dls - 1 < (int)(s0 & 0x1ff)) allow_diff = false;
                        else if ((int)(s1 & 0x1ff) == dls - 1 && (int)(s0 & 0x1ff) == dls - 1 && (s1 & 0x8000u)) allow_mm = false;
                    }
                    const int gaps = go + ge;
                    const bool allow_indels = !(i1 < a.no_indel_len + gaps || len - i1 < a.no_indel_len + gaps) &&
                                              !(go >= a.max_gapo && ge >= a.max_gape);
                    const bool opening = (state == 0u);
                    const bool gap_allowed = allow_diff && allow_indels && (opening ? (go < a.max_gapo) : (ge < a.max_gape));
                    // bit0 full (mismatch children allowed), bit1 deletions allowed, bit2 insertion allowed
                    t_flags = ((allow_diff && allow_mm) ? 1u : 0u) | ((gap_allowed && state != 1u) ? 2u : 0u) |
                              ((gap_allowed && state != 2u) ? 4u : 0u);
                    cbase = nt4_compl(base_i1);                         // rc[i-1]
                    have_task = true;
                    task_tail = false;
                }
            }
        }

        // ================= B: rank stage =================
        __syncwarp();
        uint32_t okmask = 0;
        if (have_task) okmask = rank_general<T, HOIST>(a.ix, sLj, sUj, tid, e.L, e.U, task_tail || !multiref, lastrow, have_ub, ub);

        // ================= C: consume the result =================
        // children of this warp's expansions are described here (owner registers) and written below, one child
        // per lane and pass, whichever lane owns them
        uint32_t ch_n = 0, ch_base = NIL, ch_masks = 0, ch_compat = 0, ch_bk = 0, ch_free = 0;
        if (have_task) {
            have_task = false;
            c_rank += 2u;
            // Child index space t: multi-genome t = code (1..15, the reference's loop order);
            // single-genome (-S) t = 0..3 = A,G,C,T = codes 15,3,7,1 (O_actg_alphabet's order, bwt.c:440-463).
            // compat_set = indices that MATCH the read base: nucl_bases_table[c] (io.h:102-106, N excluded)
            // resp. the base itself (inexact_match.c:476).
            if (!multiref)
                okmask = ((okmask >> 15) & 1u) | (((okmask >> 3) & 1u) << 1) | (((okmask >> 7) & 1u) << 2) | (((okmask >> 1) & 1u) << 3);
            const uint32_t compat_set = !multiref ? (cbase < 4u ? (1u << cbase) : 0u)
                : (cbase == 0 ? 0xFB00u : (cbase == 1 ? 0x383Cu : (cbase == 2 ? 0x0BF0u : (cbase == 3 ? 0x6266u : 0u))));
            if (task_tail) {
                // ---- exact tail: ordered append with adjacent merge (align.c:93-110)
                bool ok_all = true;
                uint32_t m = okmask & compat_set;
                while (m) {
                    const int j = BWB_CODE_OF(__ffs(m) - 1);
                    m &= m - 1;
                    const T Lj = sLj[j][tid], Uj = sUj[j][tid];
                    nx_w += (uint32_t)(Uj - Lj + 1);
                    if (nx_n && Lj == (T)(nx_tailU + 1)) {
                        nx_tailU = Uj;
                    } else {
                        const uint32_t sl = lane_alloc(al, a, lane_slot);
                        if (sl == NIL) { ok_all = false; break; }
                        if (nx_n) slot_write<T>(a.slots, nx_tail, nx_tailL, nx_tailU, 0u, 0u, sl, 0u, 0u, 0u);
                        else nx_head = sl;
                        nx_tail = sl; nx_tailL = Lj; nx_tailU = Uj;
                        nx_n++;
                    }
                }
                if (!ok_all) { err = BWB_ERR_CAPACITY; mode = FLUSH; }
            } else {
                // ---- children in the reference's push order (:433-504): insertion, deletions by code, then
                // matches/mismatches by code.  Here: which children exist, which of them is the next pop, which are
                // dead; their slots (contiguous, in push order) and the bucket bookkeeping.
                const uint32_t z = e.z;
                const bool opening = ((z >> 28) & 3u) == 0u;
                const int b0 = eb, b1 = eb + a.mm_score, b2 = eb + (opening ? a.gapo_score : a.gape_score);
                const bool full = t_flags & 1u, del_ok = t_flags & 2u, ins_ok = t_flags & 4u;
                const uint32_t zm = (z - 1u) & ~(3u << 28);
                uint32_t md = del_ok ? okmask : 0u;
                uint32_t mmk = full ? okmask : (okmask & compat_set);
                c_push += __popc(md) + __popc(mmk) + (ins_ok ? 1u : 0u);
                // last match child = next pop: keep it in registers (see have_next);
                // (with mm_score == 0 mismatch children share the parent's bucket and count as well)
                const uint32_t cand = (b1 == b0) ? mmk : (mmk & compat_set);
                if (cand) {
                    const int jk = 31 - __clz(cand);
                    mmk &= ~(1u << jk);
                    have_next = true;
                    nx.L = sLj[BWB_CODE_OF(jk)][tid]; nx.U = sUj[BWB_CODE_OF(jk)][tid];
                    nx.z = zm + (((compat_set >> jk) & 1u) ? 0u : 0x100u);
                    nx.w = e.w; nx.r1 = e.r1; nx.r2 = e.r2; nx.r3 = e.r3;
                    nx_bucket = b0;
#ifdef BWB_PF_D
                    pf_ok = true;
#endif
                }
                bool ins_live = ins_ok;
#ifndef BWB_LANE_NO_GHOST
                {   // dead score classes: count, do not store (b0 = the parent's class is alive by construction)
                    const int dead_lim = best_score + a.mm_score;
                    if (b2 > dead_lim) { ghost += __popc(md) + (ins_ok ? 1 : 0); md = 0u; ins_live = false; }
                    if (b1 > dead_lim) { ghost += __popc(mmk & ~compat_set); mmk &= compat_set; }
                }
#endif
                const uint32_t n = (ins_live ? 1u : 0u) + (uint32_t)__popc(md) + (uint32_t)__popc(mmk);
                if (n) {
                    // the three score classes as compact buckets (equal scores <=> equal buckets)
                    const uint32_t k0 = bucket_of[b0], k1 = bucket_of[b1 < a.nb ? b1 : a.nb - 1], k2 = bucket_of[b2 < a.nb ? b2 : a.nb - 1];
                    const bool no_bucket = ((ins_live || md) && k2 == 0xffu) || ((mmk & ~compat_set) && k1 == 0xffu) ||
                                           ((mmk & compat_set) && k0 == 0xffu);
                    // recycled slots first (top of the free stack), the rest contiguous from the bump allocator
                    const uint32_t f = RECYCLE ? (n < nfree ? n : nfree) : 0u;
                    ch_base = (n > f) ? lane_alloc_n(al, a, lane_slot, n - f) : 0u;
                    if (no_bucket) { err = BWB_ERR_UNSUPPORTED; mode = FLUSH; ch_base = NIL; }     // bucket_map() missed a score: a bug, reported
                    else if (ch_base == NIL) { err = BWB_ERR_CAPACITY; mode = FLUSH; }
                    else {
                        ch_free = f | (nfree << 8);
                        nfree -= f;
                        // occupancy bits once per score class (of what is really pushed)
                        if (ins_live || md) h.mark((int)k2);
                        if (mmk & ~compat_set) h.mark((int)k1);
                        if (mmk & compat_set) h.mark((int)k0);
                        h.n += (int)n;
                        ch_n = n;
                        ch_masks = md | (mmk << 16);
                        ch_compat = (compat_set & 0xffffu) | (ins_live ? (1u << 16) : 0u);
                        ch_bk = k0 | (k1 << 8) | (k2 << 16) | ((uint32_t)len << 24);
                    }
                }
            }
        }

        // ---- the children of all 32 lanes, flattened: child q of owner o is item P(o) + q; each pass writes 32
        // of them.  (Round-1 pushed the k-th child of every lane together: max-over-lanes iterations with 5 of 32
        // lanes busy, 30 % of all issued instructions.)
        {
            uint32_t inc = ch_n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL, inc, o);
                if ((int)lane >= o) inc += t;
            }
            const uint32_t total = __shfl_sync(FULL, inc, 31);
            const uint32_t pex = inc - ch_n;
            for (uint32_t t0 = 0; t0 < total; t0 += 32u) {
                const uint32_t t = t0 + lane;
                const bool on = t < total;
                // owner = last lane whose exclusive prefix is <= t
                int o = 0;
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) {
                    const uint32_t v = __shfl_sync(FULL, pex, o + s);
                    if (v <= t) o += s;
                }
                const uint32_t q = t - __shfl_sync(FULL, pex, o);
                const uint32_t base = __shfl_sync(FULL, ch_base, o);
                const uint32_t masks = __shfl_sync(FULL, ch_masks, o);
                const uint32_t cmp = __shfl_sync(FULL, ch_compat, o);
                const uint32_t bk = __shfl_sync(FULL, ch_bk, o);
                const uint32_t fr = RECYCLE ? __shfl_sync(FULL, ch_free, o) : 0u;
                const uint32_t pz = __shfl_sync(FULL, e.z, o), pw = __shfl_sync(FULL, e.w, o);
                const T pL = shfl(e.L, o), pU = shfl(e.U, o);
                uint32_t pr1 = 0, pr2 = 0, pr3 = 0;
                if (WIDE) { pr1 = __shfl_sync(FULL, e.r1, o); pr2 = __shfl_sync(FULL, e.r2, o); pr3 = __shfl_sync(FULL, e.r3, o); }
                const uint32_t otid = (warp << 5) + (uint32_t)o;
                const uint32_t fcnt = fr & 0xffu, ftop = fr >> 8;
                // slot of the child at position `pos` of the owner's push order
#define BWB_SLOT_OF(pos) ((RECYCLE && (pos) < fcnt) ? sFree[(ftop - 1u - (pos)) & (uint32_t)(FREE_RING - 1)][otid] : base + ((pos) - fcnt))
                const uint32_t md = masks & 0xffffu, mmk = masks >> 16, compat = cmp & 0xffffu;
                const uint32_t n_ins = (cmp >> 16) & 1u, n_gap = n_ins + (uint32_t)__popc(md);
                const int b0 = (int)(bk & 0xffu), b1 = (int)((bk >> 8) & 0xffu), b2 = (int)((bk >> 16) & 0xffu);
                const bool is_gap = q < n_gap;
                // mm children that share a bucket with the gap children (score classes can coincide)
                const uint32_t mm_in_b2 = mmk & (((b0 == b2) ? compat : 0u) | ((b1 == b2) ? ~compat : 0u));
                T cL, cU;
                uint32_t cz, cw = pw, cr1 = pr1, cr2 = pr2, cr3 = pr3, nxt_slot = NIL;
                int X;
                bool need_old, last;
                if (is_gap) {
                    const bool is_ins = q < n_ins;
                    const bool opening = ((pz >> 28) & 3u) == 0u;
                    const int go = (int)((pz >> 24) & 15u);
                    const uint32_t zg = (pz & ~(3u << 28)) + (opening ? (1u << 24) : (1u << 16));
                    const uint32_t alen = ((uint32_t)((int)(bk >> 24) - (int)(pz & 0xffu)) + (pw & 0xffu)) & 0xffu;
                    const uint32_t newrun = alen | (1u << 8) | ((is_ins ? 1u : 2u) << 16);
                    if (!is_ins) cw = pw + 1u;
                    if (opening) {
                        if (go == 0) cw = (cw & 0xffu) | (newrun << 8);
                        else if (WIDE && go == 1) cr1 = newrun;
                        else if (WIDE && go == 2) cr2 = newrun;
                        else if (WIDE) cr3 = newrun;
                    } else {
                        if (go == 1) cw += 1u << 16;
                        else if (WIDE && go == 2) cr1 += 1u << 8;
                        else if (WIDE && go == 3) cr2 += 1u << 8;
                        else if (WIDE && go == 4) cr3 += 1u << 8;
                    }
                    X = b2;
                    need_old = (q == 0u);
                    if (!need_old) nxt_slot = BWB_SLOT_OF(q - 1u);
                    last = (q + 1u == n_gap) && (mm_in_b2 == 0u);
                    if (is_ins) {
                        cL = pL; cU = pU;
                        cz = (zg | (1u << 28)) - 1u;
                    } else {
                        const int j = BWB_CODE_OF(kth_bit16(md, q - n_ins));
                        cL = sLj[j][otid]; cU = sUj[j][otid];
                        cz = zg | (2u << 28);
                    }
                } else {
                    const int tt = on ? kth_bit16(mmk, q - n_gap) : 0;
                    const bool is_mm = !((compat >> tt) & 1u);
                    X = is_mm ? b1 : b0;
                    const uint32_t same = mmk & (((b0 == X) ? compat : 0u) | ((b1 == X) ? ~compat : 0u));
                    const uint32_t below = same & ((1u << tt) - 1u);
                    need_old = false;
                    if (below) {
                        const int pc = 31 - __clz(below);
                        const uint32_t pp = n_gap + (uint32_t)__popc(mmk & ((1u << pc) - 1u));
                        nxt_slot = BWB_SLOT_OF(pp);
                    } else if (b2 == X && n_gap > 0u) {
                        nxt_slot = BWB_SLOT_OF(n_gap - 1u);
                    } else {
                        need_old = true;
                    }
                    last = (same >> (tt + 1)) == 0u;
                    const int j = BWB_CODE_OF(tt);
                    cL = sLj[j][otid]; cU = sUj[j][otid];
                    cz = ((pz - 1u) & ~(3u << 28)) + (is_mm ? 0x100u : 0u);
                }
                uint32_t *hd = sm_heads + (uint32_t)X * 128u + otid;
                if (on && need_old) nxt_slot = *hd;
                __syncwarp();                                            // old heads are read before new ones are written
                if (on) {
                    const uint32_t mine = BWB_SLOT_OF(q);
                    slot_write<T>(a.slots, mine, cL, cU, cz, cw, nxt_slot, cr1, cr2, cr3);
                    if (last) *hd = mine;
                }
#undef BWB_SLOT_OF
            }
        }
#undef BWB_CODE_OF

        // ================= flush: hits -> output group (K5 restores input order) =================
        if (mode == FLUSH && !have_task) {
            if (err) {
                // out of arena: with every lane busy on a big heap the shared pool can run dry.  The read is
                // handed to the next pass, which runs only the deferred reads (so each finds far more room);
                // the last pass reports the overflow.
                if (a.retry_list && err == BWB_ERR_CAPACITY) a.retry_list[atomicAdd(a.retry_count, 1u)] = r;
                else if (atomicCAS(a.status, 0u, (uint32_t)(-err)) == 0u) a.status[1] = read_id;
                n_hits = 0;
            }
            const unsigned long long base = atomicAdd(a.out_cursor, (unsigned long long)n_hits);
            if (base + n_hits <= a.out_cap) {
                uint32_t q = hit_head;
                for (int k = 0; k < n_hits; k++) {
                    PE<T> hh, hb;
                    const uint32_t sb = slot_read<T>(a.slots, q, hh);
                    q = slot_read<T>(a.slots, sb, hb);
                    const uint32_t z = hh.z, go = (z >> 24) & 15u;
                    bwb_hit ht;
                    ht.L = (uint64_t)hh.L; ht.U = (uint64_t)hh.U; ht.score = (int32_t)hb.L;
                    ht.num_mm = (uint8_t)((z >> 8) & 0xffu); ht.num_gapo = (uint8_t)go; ht.num_gape = (uint8_t)((z >> 16) & 0xffu);
                    ht.aln_length = (uint8_t)hb.U; ht.n_runs = (uint8_t)go; ht.pad[0] = ht.pad[1] = ht.pad[2] = 0;
                    ht.read_id = read_id;
                    const uint32_t rr4[BWB_MAX_GAP_RUNS] = {hh.w >> 8, hh.r1, hh.r2, hh.r3};
#pragma unroll
                    for (int t = 0; t < BWB_MAX_GAP_RUNS; t++) {
                        const uint32_t v = (uint32_t)t < go ? rr4[t] : 0u;
                        ht.runs[t].start = (uint8_t)(v & 0xffu); ht.runs[t].len = (uint8_t)((v >> 8) & 0xffu);
                        ht.runs[t].state = (uint8_t)((v >> 16) & 0xffu); ht.runs[t].pad = 0;
                    }
                    a.out_hits[base + k] = ht;
                }
            }
            a.read_off[r] = base;
            a.read_cnt[r] = (uint32_t)n_hits;
            atomicAdd(a.counters + 0, (unsigned long long)c_pops);
            atomicAdd(a.counters + 1, (unsigned long long)c_push);
            atomicAdd(a.counters + 2, (unsigned long long)c_tails);
            atomicAdd(a.counters + 3, (unsigned long long)c_rank);
            c_pops = c_push = c_tails = c_rank = 0u;
            lane_alloc_reset(al, a, lane_slot);
            have_next = false;
            mode = NEED;
        }
    }

#undef BWB_RSEQ
#undef BWB_D
#undef BWB_DS
    atomicMax(a.counters + 4, (unsigned long long)c_maxheap);
    atomicMax(a.counters + 5, (unsigned long long)c_maxlist);
}

}  // namespace bwb

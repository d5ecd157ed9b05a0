// bwb_device.cuh -- device-side building blocks of the BWBBLE hot path on sm_100a.
//
//   * 128-byte index blocks (K0 layout) and the rank primitives O(c,i) / O_alphabet(i)
//     (reference: mg-aligner/bwt.c:348-372, :374-438, :575-600, :689-781)
//   * warp-cooperative multi-interval backward-extension step with ordered adjacent-merge
//     (reference: exact_match.c:88-113, inexact_match.c:219-232, align.c:93-110)
//
// Index block (128 B, 128-B aligned, one per 128 BWT rows):
//     uint32 cnt[16]      cnt[c] = #rows < 128*blk holding code c (sentinel row not counted)
//     uint32 plane[4][4]  bit p%32 of plane[k][p/32] = bit k of the code in row 128*blk+p
// so a rank query touches exactly one cache line: 1 x LDG.32 (its counter) + 4 x LDG.128 (planes),
// and the in-block count of code c up to row offset r is popc(AND_k (plane_k ^ ~c_k) & mask(r)).
// The reference stores inclusive checkpoints (row 128*blk counted, bwt.c:280-291) and subtracts the
// checkpoint symbol again (bwt.c:596-598); the exclusive counter here yields the same values.
//
// Everything is templated on the SA-coordinate type: uint32_t when the index has < 2^32-1 rows
// (chr21 scale), uint64_t otherwise (genome scale).  "-1" (L-1 of L==0) is the all-ones value of
// the type; it can never collide with a real row because length-1 < all-ones.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bwb {

constexpr unsigned FULL = 0xffffffffu;
constexpr int SL = 32;   // intervals of each list kept in shared memory; the rest spills to HBM

struct IndexView {
    const uint4 *blocks;   // num_blocks * 8 uint4
    uint64_t length;       // BWT rows (= text length + 1)
    uint64_t num_blocks;
    uint64_t C[17];
};

template <class T> struct Pair;
template <> struct Pair<uint32_t> { typedef uint2 type; };
template <> struct Pair<uint64_t> { typedef ulonglong2 type; };

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ uint32_t shfl(uint32_t v, int src) { return __shfl_sync(FULL, v, src); }
__device__ __forceinline__ uint64_t shfl(uint64_t v, int src) {
    uint32_t lo = __shfl_sync(FULL, (uint32_t)v, src);
    uint32_t hi = __shfl_sync(FULL, (uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint32_t shfl_xor(uint32_t v, int m) { return __shfl_xor_sync(FULL, v, m); }
__device__ __forceinline__ uint64_t shfl_xor(uint64_t v, int m) {
    uint32_t lo = __shfl_xor_sync(FULL, (uint32_t)v, m);
    uint32_t hi = __shfl_xor_sync(FULL, (uint32_t)(v >> 32), m);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) { return shfl(v, src); }

// per-block C[] staged in shared memory, in the coordinate type
template <class T>
__device__ __forceinline__ void stage_C(const IndexView &ix, T *sC) {
    if (threadIdx.x < 17) sC[threadIdx.x] = (T)ix.C[threadIdx.x];
    __syncthreads();
}

struct BlockBits {       // match mask of one code over the 128 rows of a block + its counter
    uint32_t cnt, m0, m1, m2, m3;
};

struct Planes { uint4 p0, p1, p2, p3; };   // the four bit planes of a block (64 B)

__device__ __forceinline__ Planes load_planes(const uint4 *__restrict__ blk) {
    Planes p;
    p.p0 = __ldg(blk + 4); p.p1 = __ldg(blk + 5); p.p2 = __ldg(blk + 6); p.p3 = __ldg(blk + 7);
    return p;
}

// counter of code c + the 128-bit mask of rows holding c
__device__ __forceinline__ BlockBits match_code(const Planes &p, const uint4 *__restrict__ blk, uint32_t c) {
    BlockBits b;
    b.cnt = __ldg(reinterpret_cast<const uint32_t *>(blk) + c);
    const uint32_t x0 = 0u - (~c & 1u), x1 = 0u - ((~c >> 1) & 1u);      // all ones where bit k of c is 0
    const uint32_t x2 = 0u - ((~c >> 2) & 1u), x3 = 0u - ((~c >> 3) & 1u);
    b.m0 = (p.p0.x ^ x0) & (p.p1.x ^ x1) & (p.p2.x ^ x2) & (p.p3.x ^ x3);
    b.m1 = (p.p0.y ^ x0) & (p.p1.y ^ x1) & (p.p2.y ^ x2) & (p.p3.y ^ x3);
    b.m2 = (p.p0.z ^ x0) & (p.p1.z ^ x1) & (p.p2.z ^ x2) & (p.p3.z ^ x3);
    b.m3 = (p.p0.w ^ x0) & (p.p1.w ^ x1) & (p.p2.w ^ x2) & (p.p3.w ^ x3);
    return b;
}

__device__ __forceinline__ BlockBits load_block(const uint4 *__restrict__ blk, uint32_t c) {
    return match_code(load_planes(blk), blk, c);
}

// cnt + #{rows p in [0, r] of the block holding the code}
__device__ __forceinline__ uint32_t rank_in_block(const BlockBits &b, uint32_t r) {
    // 128-bit mask of bits 0..r as four words, branch-free: word w gets clamp(r+1-32w, 0, 32) low bits
    const int n = (int)r + 1;
    const uint32_t k0 = n >= 32 ? ~0u : ((1u << n) - 1u);
    const int n1 = n - 32, n2 = n - 64, n3 = n - 96;
    const uint32_t k1 = n1 >= 32 ? ~0u : (n1 <= 0 ? 0u : ((1u << n1) - 1u));
    const uint32_t k2 = n2 >= 32 ? ~0u : (n2 <= 0 ? 0u : ((1u << n2) - 1u));
    const uint32_t k3 = n3 >= 32 ? ~0u : (n3 <= 0 ? 0u : ((1u << n3) - 1u));
    return b.cnt + __popc(b.m0 & k0) + __popc(b.m1 & k1) + __popc(b.m2 & k2) + __popc(b.m3 & k3);
}

// O(c, i), c in 1..15 (bwt.c:348-372).
template <class T>
__device__ __forceinline__ T occ1(const IndexView &ix, const T *sC, uint32_t c, T i) {
    const T last = (T)(ix.length - 1);
    const bool top = (i == last), neg = (i == (T)~(T)0);
    const T ii = (top || neg) ? (T)0 : i;
    const BlockBits b = load_block(ix.blocks + (size_t)(ii >> 7) * 8, c);
    const uint32_t v = rank_in_block(b, (uint32_t)(ii & 127u));
    return top ? (T)(sC[c + 1] - sC[c]) : (neg ? (T)0 : (T)v);
}

// O(c, iL) and O(c, iU) for the two ends of one interval.  When both rows fall into the same
// block (the common case once the interval is narrow) the block is loaded and matched once.
template <class T>
__device__ __forceinline__ void occ_pair(const IndexView &ix, const T *sC, uint32_t c, T iL, T iU, T &oL, T &oU) {
    const T last = (T)(ix.length - 1), none = (T)~(T)0;
    const bool topL = (iL == last), negL = (iL == none), topU = (iU == last), negU = (iU == none);
    const T aL = (topL || negL) ? (T)0 : iL, aU = (topU || negU) ? (T)0 : iU;
    const BlockBits bu = load_block(ix.blocks + (size_t)(aU >> 7) * 8, c);
    const uint32_t vU = rank_in_block(bu, (uint32_t)(aU & 127u));
    uint32_t vL;
    if ((aL >> 7) == (aU >> 7)) {
        vL = rank_in_block(bu, (uint32_t)(aL & 127u));
    } else {
        const BlockBits bl = load_block(ix.blocks + (size_t)(aL >> 7) * 8, c);
        vL = rank_in_block(bl, (uint32_t)(aL & 127u));
    }
    const T tot = (T)(sC[c + 1] - sC[c]);
    oL = topL ? tot : (negL ? (T)0 : (T)vL);
    oU = topU ? tot : (negU ? (T)0 : (T)vU);
}

// occ[j] of O_alphabet(i, inc) for one code j (bwt.c:374-438) including quirk Q1: codes 5,9,11,13
// get neither the in-block count nor the checkpoint, only the "checkpoint symbol" decrement
// (bwt.c:427-435,780), in wrapping arithmetic of the reference's u64 (low bits are what matter:
// the values are only compared L<=U and stored; for T=uint32_t see k_align's idx32 precondition).
template <class T>
__device__ __forceinline__ T occ_alpha(const IndexView &ix, const T *sC, uint32_t j, T i, uint32_t inc) {
    const T last = (T)(ix.length - 1);
    const bool top = (i == last), neg = (i == (T)~(T)0);
    const T ii = (top || neg) ? (T)0 : i;
    const BlockBits b = load_block(ix.blocks + (size_t)(ii >> 7) * 8, j);
    const uint32_t v = rank_in_block(b, (uint32_t)(ii & 127u));
    const bool quirk = (0x2A20u >> j) & 1u;
    const T mid = quirk ? (T)(sC[j] - (T)(b.m0 & 1u)) : (T)(sC[j] + (T)v);
    const T r = top ? sC[j + 1] : (neg ? sC[j] : mid);
    return (T)(r + inc);
}

// ---- interval lists: first SL entries in shared memory, the rest in a per-warp HBM scratch -------
template <class T>
struct ListStore {
    typedef typename Pair<T>::type P;
    P *s;   // [2][SL]
    P *g;   // [2][cap]
    int cap;
};
template <class T>
__device__ __forceinline__ typename Pair<T>::type lget(const ListStore<T> &ls, int which, int k) {
    return k < SL ? ls.s[which * SL + k] : ls.g[(size_t)which * ls.cap + k];
}
template <class T>
__device__ __forceinline__ void lset(const ListStore<T> &ls, int which, int k, T L, T U) {
    typename Pair<T>::type v;
    v.x = L;
    v.y = U;
    if (k < SL) ls.s[which * SL + k] = v;
    else ls.g[(size_t)which * ls.cap + k] = v;
}

// nucl_bases_table (io.h:102-106) packed one nibble per entry, rows in nt4 order A,G,C,T
__device__ __forceinline__ uint32_t compat_codes(uint32_t c) {
    const uint32_t lo = (c & 1u) ? 0xDCB5432u : 0xFEDCB98u;   // G : A
    const uint32_t hi = (c & 1u) ? 0xED96521u : 0xB987654u;   // T : C
    return (c & 2u) ? hi : lo;
}

// One backward-extension step of list `cur` (n_cur intervals) by read base c (0..3) into list cur^1.
// Work item = (interval, one of its 7 compatible codes); 4 intervals x 7 codes = 28 lanes per pass,
// each lane does the two rank queries of its item.  Items are kept in the reference's order
// (interval order x code ascending) and an item merges into its predecessor iff L == prev.U + 1
// (add_sa_interval, align.c:93-110): ballots find run heads, shuffles fetch each run's last U.
// Returns the new list length, or -1 if it exceeds ls.cap.  sumw = wrapped int sum of widths.
template <class T>
__device__ __forceinline__ int extend_step(const IndexView &ix, const T *sC, const ListStore<T> &ls, int cur,
                                           int n_cur, uint32_t c, uint32_t &sumw, uint32_t &nloads) {
    const uint32_t lane = lane_id();
    const uint32_t g = lane / 7u, k = lane - g * 7u;
    const uint32_t code = (compat_codes(c) >> (4u * k)) & 15u;
    const T Cc = sC[code];
    const int nxt = cur ^ 1;
    const uint32_t lt = (1u << lane) - 1u;
    int n_next = 0;
    bool tail_valid = false;
    T tailU = 0;
    uint32_t acc = 0;
    for (int base = 0; base < n_cur; base += 4) {
        const int s = base + (int)g;
        const bool active = (g < 4u) && (s < n_cur);
        typename Pair<T>::type iv;
        iv.x = 1; iv.y = 0;
        if (active) iv = lget(ls, cur, s);
        T oL, oU;
        occ_pair<T>(ix, sC, code, (T)(iv.x - 1), iv.y, oL, oU);
        const T nL = (T)(Cc + oL + 1), nU = (T)(Cc + oU);
        const bool valid = active && (nL <= nU);
        nloads += active ? 2u : 0u;

        const uint32_t V = __ballot_sync(FULL, valid);
        const uint32_t below = V & lt;
        const T prevU = shfl(nU, below ? (31 - __clz(below)) : 0);
        const bool cmp_ok = below ? true : tail_valid;
        const T cmpU = below ? prevU : tailU;
        const bool head = valid && !(cmp_ok && nL == (T)(cmpU + 1));
        const uint32_t H = __ballot_sync(FULL, head);
        // last valid lane of my run = highest valid lane below the next head
        const uint32_t above = H & ~((2u << lane) - 1u);
        const uint32_t lim = above ? ((1u << (__ffs(above) - 1)) - 1u) : FULL;
        const uint32_t runV = V & lim;
        const T endU = shfl(nU, runV ? (31 - __clz(runV)) : 0);
        // valid lanes before the first head extend the interval stored last
        const uint32_t leadV = V & (H ? ((1u << (__ffs(H) - 1)) - 1u) : FULL);
        if (leadV) {
            const T nu = shfl(nU, 31 - __clz(leadV));
            if (lane == 0) {
                typename Pair<T>::type t = lget(ls, nxt, n_next - 1);
                lset<T>(ls, nxt, n_next - 1, t.x, nu);
            }
        }
        const int nh = __popc(H);
        if (n_next + nh > ls.cap) return -1;
        if (head) lset<T>(ls, nxt, n_next + __popc(H & lt), nL, endU);
        n_next += nh;
        if (V) {
            tailU = shfl(nU, 31 - __clz(V));
            tail_valid = true;
        }
        acc += valid ? (uint32_t)(nU - nL + 1) : 0u;
        __syncwarp();
    }
    sumw = __reduce_add_sync(FULL, acc);
    return n_next;
}

}  // namespace bwb

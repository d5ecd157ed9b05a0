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

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
    uint32_t lo = __shfl_sync(FULL, (uint32_t)v, src);
    uint32_t hi = __shfl_sync(FULL, (uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl64_xor(uint64_t v, int m) {
    uint32_t lo = __shfl_xor_sync(FULL, (uint32_t)v, m);
    uint32_t hi = __shfl_xor_sync(FULL, (uint32_t)(v >> 32), m);
    return ((uint64_t)hi << 32) | lo;
}

// cnt[c] + #{rows p in [0, r] of this block holding code c};  first = [row 0 of the block holds c]
__device__ __forceinline__ uint32_t block_rank(const uint4 *__restrict__ blk, uint32_t c, uint32_t r,
                                               uint32_t &first) {
    const uint32_t cnt = __ldg(reinterpret_cast<const uint32_t *>(blk) + c);
    const uint4 p0 = __ldg(blk + 4), p1 = __ldg(blk + 5), p2 = __ldg(blk + 6), p3 = __ldg(blk + 7);
    const uint32_t x0 = (c & 1u) ? 0u : ~0u, x1 = (c & 2u) ? 0u : ~0u;
    const uint32_t x2 = (c & 4u) ? 0u : ~0u, x3 = (c & 8u) ? 0u : ~0u;
    const uint32_t m0 = (p0.x ^ x0) & (p1.x ^ x1) & (p2.x ^ x2) & (p3.x ^ x3);
    const uint32_t m1 = (p0.y ^ x0) & (p1.y ^ x1) & (p2.y ^ x2) & (p3.y ^ x3);
    const uint32_t m2 = (p0.z ^ x0) & (p1.z ^ x1) & (p2.z ^ x2) & (p3.z ^ x3);
    const uint32_t m3 = (p0.w ^ x0) & (p1.w ^ x1) & (p2.w ^ x2) & (p3.w ^ x3);
    const uint32_t w = r >> 5;
    const uint32_t last = (2u << (r & 31u)) - 1u;          // bits 0..r%32 (r%32==31 -> all ones)
    const uint32_t k0 = w > 0 ? ~0u : last;
    const uint32_t k1 = w > 1 ? ~0u : (w == 1 ? last : 0u);
    const uint32_t k2 = w > 2 ? ~0u : (w == 2 ? last : 0u);
    const uint32_t k3 = w == 3 ? last : 0u;
    first = m0 & 1u;
    return cnt + __popc(m0 & k0) + __popc(m1 & k1) + __popc(m2 & k2) + __popc(m3 & k3);
}

// O(c, i), c in 1..15 (bwt.c:348-372).  sC = C[] staged in shared memory.
__device__ __forceinline__ uint64_t occ1(const IndexView &ix, const uint64_t *sC, uint32_t c, uint64_t i) {
    const bool top = (i == ix.length - 1), neg = (i == ~0ull);
    const uint64_t ii = (top || neg) ? 0ull : i;
    uint32_t first;
    const uint32_t v = block_rank(ix.blocks + (ii >> 7) * 8, c, (uint32_t)(ii & 127u), first);
    return top ? (sC[c + 1] - sC[c]) : (neg ? 0ull : (uint64_t)v);
}

// occ[j] of O_alphabet(i, inc) for one code j (bwt.c:374-438) including quirk Q1: codes 5,9,11,13
// get neither the in-block count nor the checkpoint, only the "checkpoint symbol" decrement
// (bwt.c:427-435,780), in wrapping u64 arithmetic.
__device__ __forceinline__ uint64_t occ_alpha(const IndexView &ix, const uint64_t *sC, uint32_t j, uint64_t i,
                                              uint32_t inc) {
    const bool top = (i == ix.length - 1), neg = (i == ~0ull);
    const uint64_t ii = (top || neg) ? 0ull : i;
    uint32_t first;
    const uint32_t v = block_rank(ix.blocks + (ii >> 7) * 8, j, (uint32_t)(ii & 127u), first);
    const bool quirk = (0x2A20u >> j) & 1u;
    const uint64_t mid = quirk ? (sC[j] - (uint64_t)first) : (sC[j] + (uint64_t)v);
    const uint64_t r = top ? sC[j + 1] : (neg ? sC[j] : mid);
    return r + inc;
}

// ---- interval lists: first SL entries in shared memory, the rest in a per-warp HBM scratch -------
struct ListStore {
    ulonglong2 *s;   // [2][SL]
    ulonglong2 *g;   // [2][cap]
    int cap;
};
__device__ __forceinline__ ulonglong2 lget(const ListStore &ls, int which, int k) {
    return k < SL ? ls.s[which * SL + k] : ls.g[(size_t)which * ls.cap + k];
}
__device__ __forceinline__ void lset(const ListStore &ls, int which, int k, ulonglong2 v) {
    if (k < SL) ls.s[which * SL + k] = v;
    else ls.g[(size_t)which * ls.cap + k] = v;
}

// nucl_bases_table (io.h:102-106) packed one nibble per entry, rows in nt4 order A,G,C,T
__device__ __forceinline__ uint32_t compat_codes(uint32_t c) {
    return c == 0 ? 0xFEDCB98u : (c == 1 ? 0xDCB5432u : (c == 2 ? 0xB987654u : 0xED96521u));
}

// One backward-extension step of list `cur` (n_cur intervals) by read base c (0..3) into list cur^1.
// Work item = (interval, one of its 7 compatible codes); 4 intervals x 7 codes = 28 lanes per pass,
// each lane does the two rank queries of its item.  Items are kept in the reference's order
// (interval order x code ascending) and an item merges into its predecessor iff L == prev.U + 1
// (add_sa_interval, align.c:93-110): ballots find run heads, shuffles fetch each run's last U.
// Returns the new list length, or -1 if it exceeds ls.cap.  sumw = wrapped int sum of widths.
__device__ __forceinline__ int extend_step(const IndexView &ix, const uint64_t *sC, const ListStore &ls, int cur,
                                           int n_cur, uint32_t c, uint32_t &sumw, uint32_t &nloads) {
    const uint32_t lane = lane_id();
    const uint32_t g = lane / 7u, k = lane - g * 7u;
    const uint32_t code = (compat_codes(c) >> (4u * k)) & 15u;
    const uint64_t Cc = sC[code];
    const int nxt = cur ^ 1;
    const uint32_t lt = (1u << lane) - 1u;
    int n_next = 0;
    bool tail_valid = false;
    uint64_t tailU = 0;
    uint32_t acc = 0;
    for (int base = 0; base < n_cur; base += 4) {
        const int s = base + (int)g;
        const bool active = (g < 4u) && (s < n_cur);
        const ulonglong2 iv = active ? lget(ls, cur, s) : make_ulonglong2(1ull, 0ull);
        const uint64_t nL = Cc + occ1(ix, sC, code, iv.x - 1) + 1;
        const uint64_t nU = Cc + occ1(ix, sC, code, iv.y);
        const bool valid = active && (nL <= nU);
        nloads += active ? 2u : 0u;

        const uint32_t V = __ballot_sync(FULL, valid);
        const uint32_t below = V & lt;
        const uint64_t prevU = shfl64(nU, below ? (31 - __clz(below)) : 0);
        const bool cmp_ok = below ? true : tail_valid;
        const uint64_t cmpU = below ? prevU : tailU;
        const bool head = valid && !(cmp_ok && nL == cmpU + 1);
        const uint32_t H = __ballot_sync(FULL, head);
        // last valid lane of my run = highest valid lane below the next head
        const uint32_t above = H & ~((2u << lane) - 1u);
        const uint32_t lim = above ? ((1u << (__ffs(above) - 1)) - 1u) : FULL;
        const uint32_t runV = V & lim;
        const uint64_t endU = shfl64(nU, runV ? (31 - __clz(runV)) : 0);
        // valid lanes before the first head extend the interval stored last
        const uint32_t leadV = V & (H ? ((1u << (__ffs(H) - 1)) - 1u) : FULL);
        if (leadV) {
            const uint64_t nu = shfl64(nU, 31 - __clz(leadV));
            if (lane == 0) {
                ulonglong2 t = lget(ls, nxt, n_next - 1);
                t.y = nu;
                lset(ls, nxt, n_next - 1, t);
            }
        }
        const int nh = __popc(H);
        if (n_next + nh > ls.cap) return -1;
        if (head) lset(ls, nxt, n_next + __popc(H & lt), make_ulonglong2(nL, endU));
        n_next += nh;
        if (V) {
            tailU = shfl64(nU, 31 - __clz(V));
            tail_valid = true;
        }
        acc += valid ? (uint32_t)(nU - nL + 1) : 0u;
        __syncwarp();
    }
    sumw = __reduce_add_sync(FULL, acc);
    return n_next;
}

}  // namespace bwb

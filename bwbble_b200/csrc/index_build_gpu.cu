// index_build_gpu.cu -- K7: `bwbble index` with the suffix sort and everything after it on the device
// (SURVEY 8f #4b; reference: build_bwt / sais / compute_C / compute_O / SA samples, bwt.c:29-63,161-218).
//
//   suffix array   prefix doubling: ranks by the first h symbols -> sort (rank[i], rank[i+h]) -> ranks by
//                  2h symbols, until all n+1 ranks are distinct.  The first round packs 12 symbols (5 bits
//                  each) into the key, so h starts at 12.  The sorts and scans are cub::DeviceRadixSort /
//                  DeviceScan (library calls, like a cuBLAS GEMM elsewhere -- this is offline work, not the
//                  hot path); the key builders, rank relabelling and all index kernels below are ours.
//   BWT, sa0, SA samples, 4-bit packing, per-128-row histograms -> checkpoints O, C: one pass each.
//
// Text semantics are index_build.cpp's: codes 0..15 with '$' (0) after every record as an ordinary
// smallest symbol, plus one unique terminator below everything.  Output arrays are bit-identical to
// the host builder's (and so to the reference's .bwt); limited to < 2^31-16 rows (32-bit ranks, int counts).
#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include <cstdio>
#include <string>
#include <vector>

#include "bwbble_b200.h"
#include "host_common.h"

namespace {

constexpr int H0 = 12;   // symbols in the first-round key

__global__ void k_sa_first_keys(const uint8_t *__restrict__ text, uint32_t n, uint64_t *__restrict__ keys,
                                uint32_t *__restrict__ vals) {
    const uint32_t n1 = n + 1;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n1; i += gridDim.x * blockDim.x) {
        uint64_t k = 0;
#pragma unroll
        for (int t = 0; t < H0; t++) {
            const uint64_t p = (uint64_t)i + t;
            k = (k << 5) | (p < n ? (uint64_t)text[p] + 1u : 0u);      // terminator and beyond: 0
        }
        keys[i] = k;
        vals[i] = i;
    }
}

// head[j] = j+1 where a new key starts, else 0 (a max-scan turns it into the 1-based rank of j's group)
__global__ void k_sa_heads(const uint64_t *__restrict__ keys, uint32_t n1, uint32_t *__restrict__ head,
                           unsigned long long *n_groups) {
    __shared__ uint32_t cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    uint32_t mine = 0;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n1; j += gridDim.x * blockDim.x) {
        const bool h = j == 0 || keys[j] != keys[j - 1];
        head[j] = h ? j + 1 : 0u;
        mine += h;
    }
    atomicAdd(&cnt, mine);
    __syncthreads();
    if (threadIdx.x == 0 && cnt) atomicAdd(n_groups, (unsigned long long)cnt);
}

__global__ void k_sa_scatter_rank(const uint32_t *__restrict__ sa, const uint32_t *__restrict__ grp, uint32_t n1,
                                  uint32_t *__restrict__ rank) {
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n1; j += gridDim.x * blockDim.x) rank[sa[j]] = grp[j];
}

__global__ void k_sa_pair_keys(const uint32_t *__restrict__ rank, uint32_t n1, uint32_t h, uint64_t *__restrict__ keys,
                               uint32_t *__restrict__ vals) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n1; i += gridDim.x * blockDim.x) {
        const uint64_t p = (uint64_t)i + h;
        // a suffix shorter than h already holds the unique terminator in its first h symbols: its rank is final
        keys[i] = ((uint64_t)rank[i] << 32) | (p < n1 ? rank[p] : 0u);
        vals[i] = i;
    }
}

struct MaxU32 {
    __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

// BWT symbol per row, SA samples, the row of suffix 0
template <class S>
__global__ void k_bwt_rows(const uint8_t *__restrict__ text, const S *__restrict__ sa, uint64_t n1,
                           uint8_t *__restrict__ bw, uint64_t *__restrict__ sa_samples, unsigned long long *sa0_index) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n1; i += (uint64_t)gridDim.x * blockDim.x) {
        const S v = sa[i];
        if ((i & 31u) == 0) sa_samples[i >> 5] = v;
        if (v == 0) { *sa0_index = i; bw[i] = 0; }
        else bw[i] = text[v - 1];
    }
}

// one thread per 128-row block: nibble packing (8 rows per word, first row in the top nibble) + its histogram
__global__ void k_pack_and_count(const uint8_t *__restrict__ bw, uint64_t n1, const unsigned long long *sa0_index,
                                 uint32_t *__restrict__ words, uint64_t num_words, uint32_t *__restrict__ hist,
                                 uint32_t num_occ) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= num_occ) return;
    const uint64_t sa0 = *sa0_index;
    uint32_t cnt[16];
#pragma unroll
    for (int c = 0; c < 16; c++) cnt[c] = 0;
    for (int w = 0; w < 16; w++) {
        const uint64_t wi = (uint64_t)b * 16 + w;
        if (wi >= num_words) break;
        uint32_t word = 0;
        for (int j = 0; j < 8; j++) {
            const uint64_t row = wi * 8 + j;
            if (row < n1) {
                const uint32_t c = bw[row];
                word |= c << (28 - 4 * j);
                if (row != sa0) {
#pragma unroll
                    for (int q = 0; q < 16; q++) cnt[q] += (c == (uint32_t)q);
                }
            }
        }
        words[wi] = word;
    }
#pragma unroll
    for (int c = 0; c < 16; c++) hist[(size_t)c * num_occ + b] = cnt[c];
}

// O[blk][c] = #c in rows 0..128*blk inclusive (sa0 row excluded), bwt.c compute_O
__global__ void k_checkpoints(const uint32_t *__restrict__ hist, const uint64_t *__restrict__ pref, const uint8_t *__restrict__ bw,
                              const unsigned long long *sa0_index, uint32_t num_occ, uint64_t *__restrict__ O,
                              unsigned long long *totals) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= num_occ) return;
    const uint64_t row = (uint64_t)b * 128, sa0 = *sa0_index;
    const uint32_t first = bw[row];
    for (uint32_t c = 0; c < 16; c++) {
        const uint64_t before = pref[(size_t)c * num_occ + b];
        O[(size_t)b * 16 + c] = before + ((first == c && row != sa0) ? 1u : 0u);
        if (b == num_occ - 1) totals[c] = before + hist[(size_t)c * num_occ + b];
    }
}

struct Guard {   // frees everything on any exit path
    std::vector<void *> p;
    template <class T> cudaError_t alloc(T **q, size_t bytes) {
        cudaError_t e = cudaMalloc((void **)q, bytes ? bytes : 16);
        if (e == cudaSuccess) p.push_back((void *)*q);
        return e;
    }
    void release(void *q) {
        for (auto &x : p)
            if (x == q) { cudaFree(q); x = nullptr; }
    }
    ~Guard() { for (void *q : p) if (q) cudaFree(q); }
};

#define CUX(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            char m_[256];                                                                           \
            snprintf(m_, sizeof m_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return bwb_host::ctx_fail(ctx, BWB_ERR_CUDA, m_);                                       \
        }                                                                                           \
    } while (0)

int build_on_device(bwb_ctx *ctx, const std::vector<uint8_t> &text, bwb_host::HostIndex &ix, int *rounds_out) {
    int dev = 0;
    void *sv = nullptr;
    if (bwb_host::ctx_device(ctx, &dev, &sv)) return BWB_ERR_ARG;
    cudaStream_t st = (cudaStream_t)sv;
    const uint64_t n64 = text.size();
    if (n64 + 1 >= 0x7ffffff0ull)      // 32-bit ranks, int item counts in the cub calls (larger: build_on_device_wide)
        return bwb_host::ctx_fail(ctx, BWB_ERR_UNSUPPORTED, "the one-shot device sorter handles < 2^31-16 rows");
    const uint32_t n = (uint32_t)n64, n1 = n + 1;
    ix.length = n1;
    ix.num_words = ((uint64_t)n1 + 7) / 8;
    ix.num_occ = ((uint64_t)n1 + 127) / 128;
    ix.num_sa = ((uint64_t)n1 + 31) / 32;
    CUX(cudaSetDevice(dev));
    Guard g;
    uint8_t *d_text, *d_bw;
    uint64_t *k0, *k1, *d_sa_samples, *d_O, *d_pref;
    uint32_t *v0, *v1, *d_rank, *d_head, *d_words, *d_hist;
    unsigned long long *d_small;
    CUX(g.alloc(&d_text, (size_t)n + 16));
    CUX(g.alloc(&k0, (size_t)n1 * 8)); CUX(g.alloc(&k1, (size_t)n1 * 8));
    CUX(g.alloc(&v0, (size_t)n1 * 4)); CUX(g.alloc(&v1, (size_t)n1 * 4));
    CUX(g.alloc(&d_rank, (size_t)n1 * 4)); CUX(g.alloc(&d_head, (size_t)n1 * 4));
    CUX(g.alloc(&d_small, 64 * 8));
    CUX(cudaMemcpyAsync(d_text, text.data(), n, cudaMemcpyHostToDevice, st));

    size_t tmp_bytes = 0, t2 = 0;
    {
        cub::DoubleBuffer<uint64_t> kb(k0, k1);
        cub::DoubleBuffer<uint32_t> vb(v0, v1);
        CUX(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, vb, (int)n1, 0, 64, st));
        CUX(cub::DeviceScan::InclusiveScan(nullptr, t2, d_head, d_head, MaxU32(), (int)n1, st));
        if (t2 > tmp_bytes) tmp_bytes = t2;
        CUX(cub::DeviceScan::ExclusiveSum(nullptr, t2, (uint32_t *)nullptr, (uint64_t *)nullptr, (int)ix.num_occ, st));
        if (t2 > tmp_bytes) tmp_bytes = t2;
    }
    void *d_tmp;
    CUX(g.alloc(&d_tmp, tmp_bytes));

    const int threads = 256, blocks = 148 * 8;
    cub::DoubleBuffer<uint64_t> kb(k0, k1);
    cub::DoubleBuffer<uint32_t> vb(v0, v1);
    k_sa_first_keys<<<blocks, threads, 0, st>>>(d_text, n, kb.Current(), vb.Current());
    CUX(cudaGetLastError());
    int rounds = 0;
    int key_bits = 5 * H0;
    for (uint64_t h = H0;; h *= 2) {
        size_t tb = tmp_bytes;
        CUX(cub::DeviceRadixSort::SortPairs(d_tmp, tb, kb, vb, (int)n1, 0, key_bits, st));
        CUX(cudaMemsetAsync(d_small, 0, 8, st));
        k_sa_heads<<<blocks, threads, 0, st>>>(kb.Current(), n1, d_head, d_small);
        CUX(cudaGetLastError());
        unsigned long long groups = 0;
        CUX(cudaMemcpyAsync(&groups, d_small, 8, cudaMemcpyDeviceToHost, st));
        CUX(cudaStreamSynchronize(st));
        rounds++;
        if (groups == n1) break;
        if (h > n1) return bwb_host::ctx_fail(ctx, BWB_ERR_CUDA, "suffix sort did not converge");
        tb = tmp_bytes;
        CUX(cub::DeviceScan::InclusiveScan(d_tmp, tb, d_head, d_head, MaxU32(), (int)n1, st));
        k_sa_scatter_rank<<<blocks, threads, 0, st>>>(vb.Current(), d_head, n1, d_rank);
        CUX(cudaGetLastError());
        k_sa_pair_keys<<<blocks, threads, 0, st>>>(d_rank, n1, (uint32_t)(h > 0xffffffffull ? 0xffffffffu : h), kb.Current(), vb.Current());
        CUX(cudaGetLastError());
        key_bits = 64;
    }
    if (rounds_out) *rounds_out = rounds;
    const uint32_t *d_sa = vb.Current();

    // everything that follows the suffix array
    CUX(g.alloc(&d_bw, (size_t)n1 + 16));
    CUX(g.alloc(&d_sa_samples, ix.num_sa * 8));
    CUX(g.alloc(&d_words, ix.num_words * 4));
    CUX(g.alloc(&d_hist, ix.num_occ * 16 * 4));
    CUX(g.alloc(&d_pref, ix.num_occ * 16 * 8));
    CUX(g.alloc(&d_O, ix.num_occ * 16 * 8));
    k_bwt_rows<uint32_t><<<blocks, threads, 0, st>>>(d_text, d_sa, n1, d_bw, d_sa_samples, d_small + 1);
    CUX(cudaGetLastError());
    const unsigned ob = (unsigned)((ix.num_occ + 127) / 128);
    k_pack_and_count<<<ob, 128, 0, st>>>(d_bw, n1, d_small + 1, d_words, ix.num_words, d_hist, (uint32_t)ix.num_occ);
    CUX(cudaGetLastError());
    for (int c = 0; c < 16; c++) {
        size_t tb = tmp_bytes;
        CUX(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_hist + (size_t)c * ix.num_occ, d_pref + (size_t)c * ix.num_occ, (int)ix.num_occ, st));
    }
    k_checkpoints<<<ob, 128, 0, st>>>(d_hist, d_pref, d_bw, d_small + 1, (uint32_t)ix.num_occ, d_O, d_small + 8);
    CUX(cudaGetLastError());

    ix.bwt.resize(ix.num_words);
    ix.O.resize(ix.num_occ * 16);
    ix.SA.resize(ix.num_sa);
    unsigned long long small[64];
    CUX(cudaMemcpyAsync(ix.bwt.data(), d_words, ix.num_words * 4, cudaMemcpyDeviceToHost, st));
    CUX(cudaMemcpyAsync(ix.O.data(), d_O, ix.num_occ * 16 * 8, cudaMemcpyDeviceToHost, st));
    CUX(cudaMemcpyAsync(ix.SA.data(), d_sa_samples, ix.num_sa * 8, cudaMemcpyDeviceToHost, st));
    CUX(cudaMemcpyAsync(small, d_small, sizeof small, cudaMemcpyDeviceToHost, st));
    CUX(cudaStreamSynchronize(st));
    ix.sa0_index = small[1];
    ix.C[0] = 0;
    for (int c = 0; c < 16; c++) ix.C[c + 1] = ix.C[c] + small[8 + c];
    return BWB_OK;
}

// ---------------------------------------------------------------------------------------------
// K7w: the same construction for indexes of any size that fits the device (genome scale: 6.9 G rows,
// BASELINE configs[3]).  64-bit suffix positions and ranks, and never more than `chunk` suffixes in
// the sort buffers at a time (the plain K7 above sorts all n+1 (key, value) pairs in every round:
// 48 B per row).  Memory: text 1 B + SA 8 B + rank 8 B per row, + 32 B per chunk element.
//
//   round 0    suffixes are partitioned by their first 3 symbols (histogram over the text, 32768
//              buckets, consecutive buckets form a chunk); a chunk collects its positions, sorts them by
//              their first 12 symbols (60-bit key) and stores its part of SA; rank[i] = first SA position
//              of i's group;
//   round h    (h = 12, 24, 48, ...) SA windows of <= chunk rows, cut at group boundaries: members of
//              groups with more than one row are compacted, sorted by (group << B | rank[i + h] + 1),
//              written back into the rows of their group and re-ranked (Larsson & Sadakane 1999:
//              refining ranks in place inside a round is safe, a rank only ever gets finer).  Rows of
//              singleton groups are final and skipped.  Done when no group has two rows.
// ---------------------------------------------------------------------------------------------
constexpr int B3 = 15;                      // bits of the 3-symbol bucket key
constexpr uint32_t NB3 = 1u << B3;

__device__ __forceinline__ uint32_t bucket3(const uint8_t *__restrict__ text, uint64_t n, uint64_t i) {
    uint32_t k = 0;
#pragma unroll
    for (int t = 0; t < 3; t++) {
        const uint64_t p = i + t;
        k = (k << 5) | (p < n ? (uint32_t)text[p] + 1u : 0u);
    }
    return k;
}

// histogram of the 3-symbol buckets of all n+1 suffixes; runs of one symbol (N gaps) would serialise
// the atomics, so equal neighbours inside a warp are counted once
__global__ void k_w_hist3(const uint8_t *__restrict__ text, uint64_t n, unsigned long long *__restrict__ hist) {
    const uint64_t n1 = n + 1;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x; i0 < n1; i0 += stride) {       // warp-uniform trip count
        const uint64_t i = i0 + threadIdx.x;
        const bool on = i < n1;
        const uint32_t b = on ? bucket3(text, n, i) : 0xffffffffu;
        const uint32_t same = __match_any_sync(0xffffffffu, b);
        if (on && (threadIdx.x & 31u) == (uint32_t)(__ffs(same) - 1)) atomicAdd(&hist[b], (unsigned long long)__popc(same));
    }
}

// positions whose bucket lies in [blo, bhi) -> vals (any order), their 60-bit keys -> keys
__global__ void k_w_collect(const uint8_t *__restrict__ text, uint64_t n, uint32_t blo, uint32_t bhi,
                            uint64_t *__restrict__ keys, uint64_t *__restrict__ vals, unsigned long long *cursor) {
    const uint64_t n1 = n + 1;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x; i0 < n1; i0 += stride) {
        const uint64_t i = i0 + threadIdx.x;
        bool take = false;
        if (i < n1) {
            const uint32_t b = bucket3(text, n, i);
            take = b >= blo && b < bhi;
        }
        const uint32_t m = __ballot_sync(0xffffffffu, take);
        if (m) {
            const int leader = __ffs(m) - 1;
            unsigned long long base = 0;
            if ((int)(threadIdx.x & 31u) == leader) base = atomicAdd(cursor, (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (take) {
                const unsigned long long pos = base + (unsigned long long)__popc(m & ((1u << (threadIdx.x & 31u)) - 1u));
                uint64_t k = 0;
#pragma unroll
                for (int t = 0; t < H0; t++) {
                    const uint64_t p = i + t;
                    k = (k << 5) | (p < n ? (uint64_t)text[p] + 1u : 0u);
                }
                keys[pos] = k;
                vals[pos] = i;
            }
        }
    }
}

// head[t] = t + 1 where sorted key t starts a new (sub)group, else 0; hi_shift > 0 additionally marks the
// starts of the old groups (the key's bits above hi_shift) in head_hi.  A max-scan turns both into "first t".
__global__ void k_w_heads(const uint64_t *__restrict__ keys, uint32_t m, int hi_shift, uint32_t *__restrict__ head,
                          uint32_t *__restrict__ head_hi) {
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < m; t += gridDim.x * blockDim.x) {
        const uint64_t k = keys[t], kp = t ? keys[t - 1] : 0;
        head[t] = (t == 0 || k != kp) ? t + 1 : 0u;
        if (head_hi) head_hi[t] = (t == 0 || (k >> hi_shift) != (kp >> hi_shift)) ? t + 1 : 0u;
    }
}

// round 0: chunk rows [a, a + m) of SA and the ranks of their suffixes
__global__ void k_w_store0(const uint64_t *__restrict__ vals, const uint32_t *__restrict__ first, uint32_t m, uint64_t a,
                           uint64_t *__restrict__ sa, uint64_t *__restrict__ rank, unsigned long long *unsorted) {
    uint32_t mine = 0;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < m; t += gridDim.x * blockDim.x) {
        const uint64_t i = vals[t];
        sa[a + t] = i;
        rank[i] = a + (first[t] - 1u);
        const bool single = first[t] == t + 1 && (t + 1 == m || first[t + 1] == t + 2);
        mine += single ? 0u : 1u;
    }
    mine = __reduce_add_sync(0xffffffffu, mine);
    if ((threadIdx.x & 31u) == 0 && mine) atomicAdd(unsorted, (unsigned long long)mine);
}

// round h: rows [a, b) of SA; the members of groups with more than one row -> (key, suffix), any order.
// key = (group start - a) << rbits | (rank[i + h] + 1), 0 for suffixes shorter than h (unique already).
__global__ void k_w_pairs(const uint64_t *__restrict__ sa, const uint64_t *__restrict__ rank, uint64_t a, uint64_t b, uint64_t n1,
                          uint64_t h, int rbits, uint64_t *__restrict__ keys, uint64_t *__restrict__ vals, unsigned long long *cursor) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j0 = a + (uint64_t)blockIdx.x * blockDim.x; j0 < b; j0 += stride) {
        const uint64_t j = j0 + threadIdx.x;
        bool take = false;
        uint64_t i = 0, g = 0;
        if (j < b) {
            i = sa[j];
            g = rank[i];
            const bool single = g == j && (j + 1 >= n1 || rank[sa[j + 1]] == j + 1);
            take = !single;
        }
        const uint32_t m = __ballot_sync(0xffffffffu, take);
        if (m) {
            const int leader = __ffs(m) - 1;
            unsigned long long base = 0;
            if ((int)(threadIdx.x & 31u) == leader) base = atomicAdd(cursor, (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (take) {
                const unsigned long long pos = base + (unsigned long long)__popc(m & ((1u << (threadIdx.x & 31u)) - 1u));
                const uint64_t p = i + h;
                keys[pos] = ((g - a) << rbits) | (p < n1 ? rank[p] + 1u : 0u);
                vals[pos] = i;
            }
        }
    }
}

// sorted pairs back into the rows of their groups + finer ranks; counts the rows still sharing a group
__global__ void k_w_store(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ vals, const uint32_t *__restrict__ first,
                          const uint32_t *__restrict__ first_hi, uint32_t m, uint64_t a, int rbits, uint64_t *__restrict__ sa,
                          uint64_t *__restrict__ rank, unsigned long long *unsorted) {
    uint32_t mine = 0;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < m; t += gridDim.x * blockDim.x) {
        const uint64_t g = a + (keys[t] >> rbits);                 // first row of the old group
        const uint32_t fg = first_hi[t] - 1u, fk = first[t] - 1u;  // first sorted index of the old group / of the new one
        const uint64_t i = vals[t];
        sa[g + (t - fg)] = i;
        rank[i] = g + (fk - fg);
        const bool single = fk == t && (t + 1 == m || first[t + 1] == t + 2);
        mine += single ? 0u : 1u;
    }
    mine = __reduce_add_sync(0xffffffffu, mine);
    if ((threadIdx.x & 31u) == 0 && mine) atomicAdd(unsorted, (unsigned long long)mine);
}

__global__ void k_w_group_start(const uint64_t *__restrict__ sa, const uint64_t *__restrict__ rank, uint64_t j, unsigned long long *out) {
    *out = rank[sa[j]];
}

int build_on_device_wide(bwb_ctx *ctx, const std::vector<uint8_t> &text, bwb_host::HostIndex &ix, int *rounds_out,
                         uint64_t chunk) {
    int dev = 0;
    void *sv = nullptr;
    if (bwb_host::ctx_device(ctx, &dev, &sv)) return BWB_ERR_ARG;
    cudaStream_t st = (cudaStream_t)sv;
    const uint64_t n = text.size(), n1 = n + 1;
    if (n1 >= (1ull << 40)) return bwb_host::ctx_fail(ctx, BWB_ERR_UNSUPPORTED, "index longer than 2^40 rows");
    if (chunk == 0) chunk = 1ull << 29;
    if (chunk > (1ull << 30)) chunk = 1ull << 30;            // 32-bit item counts in the cub calls, 30-bit group offsets
    if (chunk < 1024) chunk = 1024;
    int rbits = 1;
    while ((1ull << rbits) < n1 + 2) rbits++;                // bits of rank + 1
    if (rbits + 31 > 64) return bwb_host::ctx_fail(ctx, BWB_ERR_UNSUPPORTED, "index too long for the 64-bit sort key");
    ix.length = n1;
    ix.num_words = (n1 + 7) / 8;
    ix.num_occ = (n1 + 127) / 128;
    ix.num_sa = (n1 + 31) / 32;
    CUX(cudaSetDevice(dev));
    Guard g;
    uint8_t *d_text;
    uint64_t *d_sa, *d_rank, *k0, *k1, *v0, *v1;
    uint32_t *d_first, *d_first_hi;
    unsigned long long *d_small, *d_hist3;
    CUX(g.alloc(&d_text, (size_t)n + 16));
    CUX(g.alloc(&d_sa, (size_t)n1 * 8));
    CUX(g.alloc(&d_rank, (size_t)n1 * 8));
    CUX(g.alloc(&k0, (size_t)chunk * 8)); CUX(g.alloc(&k1, (size_t)chunk * 8));
    CUX(g.alloc(&v0, (size_t)chunk * 8)); CUX(g.alloc(&v1, (size_t)chunk * 8));
    CUX(g.alloc(&d_first, (size_t)chunk * 4)); CUX(g.alloc(&d_first_hi, (size_t)chunk * 4));
    CUX(g.alloc(&d_small, 64 * 8));
    CUX(g.alloc(&d_hist3, (size_t)NB3 * 8));
    CUX(cudaMemcpyAsync(d_text, text.data(), n, cudaMemcpyHostToDevice, st));
    CUX(cudaMemsetAsync(d_hist3, 0, (size_t)NB3 * 8, st));

    size_t tmp_bytes = 0, t2 = 0;
    {
        cub::DoubleBuffer<uint64_t> kb(k0, k1), vb(v0, v1);
        CUX(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, vb, (int)chunk, 0, 64, st));
        CUX(cub::DeviceScan::InclusiveScan(nullptr, t2, d_first, d_first, MaxU32(), (int)chunk, st));
        if (t2 > tmp_bytes) tmp_bytes = t2;
        CUX(cub::DeviceScan::ExclusiveSum(nullptr, t2, (uint32_t *)nullptr, (uint64_t *)nullptr, (int)ix.num_occ, st));
        if (t2 > tmp_bytes) tmp_bytes = t2;
    }
    void *d_tmp;
    CUX(g.alloc(&d_tmp, tmp_bytes));
    const int threads = 256, blocks = 148 * 8;

    // ---- round 0
    k_w_hist3<<<blocks, threads, 0, st>>>(d_text, n, d_hist3);
    CUX(cudaGetLastError());
    std::vector<unsigned long long> hist3(NB3);
    CUX(cudaMemcpyAsync(hist3.data(), d_hist3, (size_t)NB3 * 8, cudaMemcpyDeviceToHost, st));
    CUX(cudaStreamSynchronize(st));
    CUX(cudaMemsetAsync(d_small, 0, 64 * 8, st));             // [0] cursor, [1] sa0, [2] unsorted rows, [3] scratch, [8..24) totals
    uint64_t row = 0;
    for (uint32_t blo = 0; blo < NB3;) {
        uint32_t bhi = blo;
        uint64_t cnt = 0;
        while (bhi < NB3 && cnt + hist3[bhi] <= chunk) cnt += hist3[bhi++];
        if (bhi == blo) {
            char msg[160];
            snprintf(msg, sizeof msg, "%llu suffixes share their first 3 symbols: more than the sort chunk (%llu); raise option index_chunk",
                     hist3[blo], (unsigned long long)chunk);
            return bwb_host::ctx_fail(ctx, BWB_ERR_CAPACITY, msg);
        }
        if (cnt) {
            cub::DoubleBuffer<uint64_t> kb(k0, k1), vb(v0, v1);
            CUX(cudaMemsetAsync(d_small, 0, 8, st));
            k_w_collect<<<blocks, threads, 0, st>>>(d_text, n, blo, bhi, kb.Current(), vb.Current(), d_small);
            CUX(cudaGetLastError());
            size_t tb = tmp_bytes;
            CUX(cub::DeviceRadixSort::SortPairs(d_tmp, tb, kb, vb, (int)cnt, 0, 5 * H0, st));
            k_w_heads<<<blocks, threads, 0, st>>>(kb.Current(), (uint32_t)cnt, 0, d_first, nullptr);
            CUX(cudaGetLastError());
            tb = tmp_bytes;
            CUX(cub::DeviceScan::InclusiveScan(d_tmp, tb, d_first, d_first, MaxU32(), (int)cnt, st));
            k_w_store0<<<blocks, threads, 0, st>>>(vb.Current(), d_first, (uint32_t)cnt, row, d_sa, d_rank, d_small + 2);
            CUX(cudaGetLastError());
            row += cnt;
        }
        blo = bhi;
    }
    if (row != n1) return bwb_host::ctx_fail(ctx, BWB_ERR_CUDA, "suffix sort: bucket histogram and collection disagree");
    unsigned long long unsorted = 0;
    CUX(cudaMemcpyAsync(&unsorted, d_small + 2, 8, cudaMemcpyDeviceToHost, st));
    CUX(cudaStreamSynchronize(st));
    int rounds = 1;

    // ---- rounds h = 12, 24, ...
    for (uint64_t h = H0; unsorted; h *= 2) {
        if (h > 2 * n1) return bwb_host::ctx_fail(ctx, BWB_ERR_CUDA, "suffix sort did not converge");
        CUX(cudaMemsetAsync(d_small + 2, 0, 8, st));
        for (uint64_t a = 0; a < n1;) {
            uint64_t b = n1;
            if (n1 - a > chunk) {                       // cut at the start of the group that holds row a + chunk
                unsigned long long gs = 0;
                k_w_group_start<<<1, 1, 0, st>>>(d_sa, d_rank, a + chunk, d_small + 3);
                CUX(cudaGetLastError());
                CUX(cudaMemcpyAsync(&gs, d_small + 3, 8, cudaMemcpyDeviceToHost, st));
                CUX(cudaStreamSynchronize(st));
                b = gs;
                if (b <= a) return bwb_host::ctx_fail(ctx, BWB_ERR_CAPACITY, "a group of equal suffix prefixes exceeds the sort chunk; raise option index_chunk");
            }
            cub::DoubleBuffer<uint64_t> kb(k0, k1), vb(v0, v1);
            CUX(cudaMemsetAsync(d_small, 0, 8, st));
            k_w_pairs<<<blocks, threads, 0, st>>>(d_sa, d_rank, a, b, n1, h, rbits, kb.Current(), vb.Current(), d_small);
            CUX(cudaGetLastError());
            unsigned long long m = 0;
            CUX(cudaMemcpyAsync(&m, d_small, 8, cudaMemcpyDeviceToHost, st));
            CUX(cudaStreamSynchronize(st));
            if (m) {
                int gbits = 1;
                while ((1ull << gbits) < b - a) gbits++;
                size_t tb = tmp_bytes;
                CUX(cub::DeviceRadixSort::SortPairs(d_tmp, tb, kb, vb, (int)m, 0, rbits + gbits, st));
                k_w_heads<<<blocks, threads, 0, st>>>(kb.Current(), (uint32_t)m, rbits, d_first, d_first_hi);
                CUX(cudaGetLastError());
                tb = tmp_bytes;
                CUX(cub::DeviceScan::InclusiveScan(d_tmp, tb, d_first, d_first, MaxU32(), (int)m, st));
                tb = tmp_bytes;
                CUX(cub::DeviceScan::InclusiveScan(d_tmp, tb, d_first_hi, d_first_hi, MaxU32(), (int)m, st));
                k_w_store<<<blocks, threads, 0, st>>>(kb.Current(), vb.Current(), d_first, d_first_hi, (uint32_t)m, a, rbits, d_sa, d_rank, d_small + 2);
                CUX(cudaGetLastError());
            }
            a = b;
        }
        CUX(cudaMemcpyAsync(&unsorted, d_small + 2, 8, cudaMemcpyDeviceToHost, st));
        CUX(cudaStreamSynchronize(st));
        rounds++;
    }
    if (rounds_out) *rounds_out = rounds;

    // ---- everything that follows the suffix array; the sort buffers and the ranks are no longer needed
    for (void *q : {(void *)d_rank, (void *)k0, (void *)k1, (void *)v0, (void *)v1, (void *)d_first, (void *)d_first_hi}) g.release(q);
    uint8_t *d_bw;
    uint64_t *d_sa_samples, *d_O, *d_pref;
    uint32_t *d_words, *d_hist;
    CUX(g.alloc(&d_bw, (size_t)n1 + 16));
    CUX(g.alloc(&d_sa_samples, ix.num_sa * 8));
    CUX(g.alloc(&d_words, ix.num_words * 4));
    CUX(g.alloc(&d_hist, ix.num_occ * 16 * 4));
    CUX(g.alloc(&d_pref, ix.num_occ * 16 * 8));
    CUX(g.alloc(&d_O, ix.num_occ * 16 * 8));
    k_bwt_rows<uint64_t><<<blocks, threads, 0, st>>>(d_text, d_sa, n1, d_bw, d_sa_samples, d_small + 1);
    CUX(cudaGetLastError());
    const unsigned ob = (unsigned)((ix.num_occ + 127) / 128);
    k_pack_and_count<<<ob, 128, 0, st>>>(d_bw, n1, d_small + 1, d_words, ix.num_words, d_hist, (uint32_t)ix.num_occ);
    CUX(cudaGetLastError());
    for (int c = 0; c < 16; c++) {
        size_t tb = tmp_bytes;
        CUX(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_hist + (size_t)c * ix.num_occ, d_pref + (size_t)c * ix.num_occ, (int)ix.num_occ, st));
    }
    k_checkpoints<<<ob, 128, 0, st>>>(d_hist, d_pref, d_bw, d_small + 1, (uint32_t)ix.num_occ, d_O, d_small + 8);
    CUX(cudaGetLastError());

    ix.bwt.resize(ix.num_words);
    ix.O.resize(ix.num_occ * 16);
    ix.SA.resize(ix.num_sa);
    unsigned long long small[64];
    CUX(cudaMemcpyAsync(ix.bwt.data(), d_words, ix.num_words * 4, cudaMemcpyDeviceToHost, st));
    CUX(cudaMemcpyAsync(ix.O.data(), d_O, ix.num_occ * 16 * 8, cudaMemcpyDeviceToHost, st));
    CUX(cudaMemcpyAsync(ix.SA.data(), d_sa_samples, ix.num_sa * 8, cudaMemcpyDeviceToHost, st));
    CUX(cudaMemcpyAsync(small, d_small, sizeof small, cudaMemcpyDeviceToHost, st));
    CUX(cudaStreamSynchronize(st));
    ix.sa0_index = small[1];
    ix.C[0] = 0;
    for (int c = 0; c < 16; c++) ix.C[c + 1] = ix.C[c] + small[8 + c];
    return BWB_OK;
}

}  // namespace

extern "C" int bwb_index_build_device(bwb_ctx *ctx, const char *fasta_path, int write_ref_file, int *sort_rounds) {
    if (!ctx || !fasta_path) return BWB_ERR_ARG;
    std::vector<uint8_t> text;
    int rc = bwb_host::prepare_index_text(fasta_path, write_ref_file, text);
    if (rc) return bwb_host::ctx_fail(ctx, rc, "cannot read the FASTA file / write .ann");
    bwb_host::HostIndex ix;
    // indexes of >= 2^31-16 rows (or option index_wide) take the chunked 64-bit sorter
    long long chunk = 0;
    const bool wide = bwb_host::ctx_index_options(ctx, &chunk) || (uint64_t)text.size() + 1 >= 0x7ffffff0ull;
    rc = wide ? build_on_device_wide(ctx, text, ix, sort_rounds, (uint64_t)(chunk > 0 ? chunk : 0))
              : build_on_device(ctx, text, ix, sort_rounds);
    if (rc) return rc;
    if (bwb_host::write_bwt_file(ix, (std::string(fasta_path) + ".bwt").c_str()))
        return bwb_host::ctx_fail(ctx, BWB_ERR_IO, "cannot write the .bwt file");
    return BWB_OK;
}

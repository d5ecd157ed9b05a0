// bwb_abi.cu -- host side of the C ABI declared in include/bwbble_b200.h.
//
// One bwb_ctx owns one or more CUDA devices.  The index is replicated per device; reads of one
// bwb_align call are sharded in contiguous ranges (static chunks, like the reference's OpenMP
// driver, inexact_match.c:115-116) and every shard runs K4 -> scan -> K5 on its device's stream.
// There is no CPU fallback: every entry point that needs the device fails with BWB_ERR_CUDA when
// CUDA is unusable.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "bwb_lane.cuh"
#include "bwb_group.cuh"
#include "bwb_kernels.cuh"
#include "bwbble_b200.h"
#include "host_common.h"

using namespace bwb;

namespace {

thread_local std::string g_last_error;

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
};

struct Device {
    int id = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    int sm_count = 0;
    // index
    uint4 *blocks = nullptr;
    uint64_t *sa = nullptr;   // sampled suffix array (optional: enables K6)
    // k-mer table of the top of the search tree (K0b)
    unsigned long long *ktab_w = nullptr;
    uint32_t *ktab_off = nullptr, *ktab_cnt = nullptr;
    void *ktab_iv = nullptr;
    // -P seed table (K0c): row offsets / sizes and the (L,U) pool
    uint32_t *pre_off = nullptr, *pre_cnt = nullptr;
    ulonglong2 *pre_iv = nullptr;
    // search scratch (sized for n_warps)
    int n_warps = 0, grid = 0, wpb = 0, grid3 = 0;
    size_t smem_bytes = 0;
    DevBuf glists, chunks, chunk_link, stage;
    uint32_t chunks_per_warp = 0, n_chunks = 0;
    // per-call buffers
    DevBuf seq, offsets, read_off, read_cnt, ordered_off, unordered, ordered, cub_tmp, small, d_main, d_seed, pool;
    DevBuf pk_main, pk_seed, n_count, nxt, blk_link, heads, order, retry, bmap, seed_src, seed_ext;      // lane engine
    DevBuf loc;                                                  // K6 output
    uint32_t slots_per_lane = 0, total_slots = 0, priv_total = 0;
    uint64_t auto_pool_bytes = 0;
    bool use_seed_src = false;       // this call: K3 takes the D_seed of short reads from donor reads (seed_donors)
    int engine = -1;          // engine the scratch was sized for
    // pinned staging for small D2H
    unsigned long long *h_small = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evm = nullptr;   // bracket K3 | K4 on the launch stream
};

}  // namespace

struct bwb_ctx {
    std::vector<Device> dev;
    std::string err;
    // index meta
    bool have_index = false;
    bool have_sa = false;
    uint64_t num_sa = 0;
    uint64_t length = 0, num_blocks = 0, sa0 = 0;
    uint64_t C[17] = {0};
    // options
    long long heap_pool_mb = 0;      // 0 = auto: 70 % of the free device memory, at most 128 GB
    int list_cap = 4096;
    int hits_per_read = 512;
    int warps_per_block = 8;
    int blocks_per_sm = 0;
    int engine = 0;           // 0 = read per lane (k_calc_d_g + k_search_l), 1 = warp per read (k_align),
                              // 2 = 8-lane groups (k_calc_d_g + k_search_g); 1 and 2 are A/B baselines
    int use_ktab = 1;         // k-mer table for calculate_d's top of tree (0 = off, for A/B and tests)
    int force_wide = 0;       // tests: run the 64-bit / 32-byte-entry kernels on a small index
    int throttle_forced = 0;         // set by bwb_set_option("throttle_pct"): apply it whatever the workload
    int throttle_pct = 60;           // K4 admission control: lanes wait while more than this share of the pool is lent out (0/100 = off)
    int recycle = 0;                 // K4 slot recycling: 0 = auto (index > 2^28 rows, reads > 128 bases, wide entries), 1 = on, 2 = off
    int arena_private_pct = 25;      // share of the K4 arena split into private per-lane ranges (rest: shared block pool)
    long long index_chunk = 0;       // K7w: suffixes per sort chunk (0 = 2^29)
    int index_wide = 0;              // tests: force the chunked 64-bit suffix sorter on a small genome
    long long hit_cap0 = -1;  // tests: initial capacity of the hit buffers (-1 = 2 per read + 65536), forces the regrow path
    int heavy_first = 1;      // K3b: K4 takes reads in descending order of K3's whole-read bound (0 = input order)
    // -P seed table, host copy in row order (what a .pre file holds)
    bool have_pre = false;
    int pre_multiref = 1;
    std::vector<uint32_t> pre_cnt_h, pre_off_h;
    std::vector<uint64_t> pre_lu_h;  // L,U pairs
    // every align call overwrites the per-device result buffers: un-fetched results of an older call are stale
    uint64_t launch_generation = 0;
    // K4's score <-> bucket tables for the parameters of the current call (bucket_map)
    int nbc = 0;
    std::vector<uint8_t> bmap;
    // SURVEY Q6 across calls (serial driver): the last read longer than the seed that the previous calls of a run saw;
    // short reads at the start of the next call inherit ITS D_seed (option "seed_carry", bwb_set_seed_carry)
    int seed_carry = 0;
    int carry_len = 0;
    uint8_t carry_seq[256] = {0};
};

struct bwb_reads {
    bwb_ctx *ctx = nullptr;
    uint64_t n_reads = 0;
    int max_len = 0;
    std::vector<uint64_t> shard_lo;      // per device, n_dev+1 entries
    std::vector<void *> d_seq, d_off;    // per device
    std::vector<uint64_t> shard_bases;   // per device
    // Q6 bookkeeping (seed_donors): only kept when the reads differ in length
    int min_len = 0;
    std::vector<uint8_t> len8;           // length of every read
    std::vector<uint8_t> skip_pre;       // 1 = -P skips the read: an N among its first 12 bases, or shorter than that
};

struct bwb_results {
    bwb_ctx *ctx = nullptr;
    uint64_t n_reads = 0;
    std::vector<uint32_t> counts;
    std::vector<bwb_hit> hits;
    std::vector<bwb_loc> loc;
    bool have_loc = false;
    uint64_t counters[8] = {0};
    float kernel_ms = 0.f;    // K4 duration (max over devices), CUDA events on the launch stream
    float k3_ms = 0.f;        // K3 duration (engines with a separate lower-bound kernel)
    // pending (device-resident) state
    bool fetched = false;
    uint64_t generation = 0;  // ctx->launch_generation of the call that produced the device-resident data
    std::vector<uint64_t> shard_lo;
    std::vector<uint64_t> shard_total;
    int status = 0;
};

namespace {

int fail(bwb_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    if (ctx) ctx->err = buf;
    return code;
}

#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return fail(ctx, BWB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),   \
                        __FILE__, __LINE__);                                                         \
    } while (0)

// device temporaries of one API call: released on every exit path (the CU() macro returns early)
struct DevTmp {
    std::vector<void *> p;
    template <class T>
    cudaError_t alloc(T **q, size_t bytes) {
        cudaError_t e = cudaMalloc((void **)q, bytes ? bytes : 16);
        if (e == cudaSuccess) p.push_back((void *)*q);
        return e;
    }
    void release(void *q) {
        for (auto &x : p)
            if (x == q) { cudaFree(q); x = nullptr; }
    }
    ~DevTmp() {
        for (void *q : p)
            if (q) cudaFree(q);
    }
};

int ensure(bwb_ctx *ctx, DevBuf &b, size_t bytes, bool slack = true) {
    if (b.bytes >= bytes && b.p) return BWB_OK;
    if (b.p) CU(cudaFree(b.p));
    b.p = nullptr;
    b.bytes = 0;
    size_t want = slack ? bytes + bytes / 4 + 256 : bytes + 256;
    CU(cudaMalloc(&b.p, want));
    b.bytes = want;
    return BWB_OK;
}

void release(DevBuf &b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.bytes = 0;
}

IndexView make_view(const bwb_ctx *ctx, const Device &d) {
    IndexView v;
    v.blocks = d.blocks;
    v.length = ctx->length;
    v.num_blocks = ctx->num_blocks;
    for (int i = 0; i < 17; i++) v.C[i] = ctx->C[i];
    return v;
}

int check_reads(bwb_ctx *ctx, const uint64_t *offsets, uint64_t n_reads, int &max_len, int *min_len = nullptr) {
    max_len = 0;
    int mn = n_reads ? 256 : 0;
    if (n_reads >= 0xffffffffull) return fail(ctx, BWB_ERR_ARG, "too many reads in one call");
    for (uint64_t r = 0; r < n_reads; r++) {
        if (offsets[r + 1] < offsets[r]) return fail(ctx, BWB_ERR_ARG, "offsets not monotone at read %llu", (unsigned long long)r);
        uint64_t l = offsets[r + 1] - offsets[r];
        if (l > 255) return fail(ctx, BWB_ERR_ARG, "read %llu is %llu bases; positions are 8-bit (align.h:104)", (unsigned long long)r, (unsigned long long)l);
        if ((int)l > max_len) max_len = (int)l;
        if ((int)l < mn) mn = (int)l;
    }
    if (min_len) *min_len = mn;
    // K4 addresses a read's bases and lower bounds with 32-bit offsets
    if (offsets[n_reads] - offsets[0] + n_reads >= 0xffffffffull)
        return fail(ctx, BWB_ERR_ARG, "more than 2^32 bases in one call: split the batch");
    return BWB_OK;
}

// 64-bit coordinates are needed from 2^32-16 rows on (all-ones is the "-1" row)
bool index_is_wide(const bwb_ctx *ctx) { return ctx->force_wide || ctx->length >= 0xfffffff0ull; }

// per-warp shared-memory layout of K4 (must match k_align)
struct SmemLayout {
    int per_warp, off_D, off_Ds, off_bk, off_seq;
};
SmemLayout k4_layout(int max_len, int seed_len, int nb) {
    SmemLayout L;
    int o = 2 * SL * (int)sizeof(ulonglong2);
    L.off_D = o;
    o += ((max_len + 1) * 8 + 15) & ~15;
    L.off_Ds = o;
    o += ((seed_len + 1) * 8 + 15) & ~15;
    L.off_bk = o;
    o += (nb * 12 + 15) & ~15;
    L.off_seq = o;
    o += (max_len + 15) & ~15;
    L.per_warp = o;
    return L;
}

// per-group shared-memory layout of k_search_g (must match the kernel)
SmemLayout g4_layout(int max_len, int seed_len, int nb) {
    SmemLayout L;
    int o = G_LIST_SMEM;
    L.off_D = o;
    o += ((max_len + 1) * 2 + 15) & ~15;
    L.off_Ds = o;
    o += ((seed_len + 1) * 2 + 15) & ~15;
    L.off_bk = o;
    o += (nb * 12 + 15) & ~15;
    L.off_seq = o;
    o += (max_len + 15) & ~15;
    L.per_warp = o;          // bytes per GROUP here
    return L;
}

#ifdef BWB_AB_ENGINES
// size the persistent grid and the per-group scratch for the group engine (K3 + K4)
int prepare_search_group(bwb_ctx *ctx, Device &d, const SmemLayout &L, bool wide) {
    CU(cudaSetDevice(d.id));
    const int tpb = 256, gpb = tpb / GL;
    const size_t smem = (size_t)gpb * L.per_warp;
    if (smem > 227 * 1024) return fail(ctx, BWB_ERR_ARG, "shared memory per block %zu exceeds 227 KB", smem);
    int bps = ctx->blocks_per_sm;
    if (wide) CU(cudaFuncSetAttribute(k_search_g<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else CU(cudaFuncSetAttribute(k_search_g<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (bps <= 0) {
        if (wide) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_search_g<true>, tpb, smem));
        else CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_search_g<false>, tpb, smem));
        if (bps <= 0) return fail(ctx, BWB_ERR_CUDA, "k_search_g does not fit on an SM");
    }
    const int grid = bps * d.sm_count;
    const int n_groups = grid * gpb;
    d.smem_bytes = smem;
    if (n_groups != d.n_warps || d.engine != 2) {
        release(d.glists); release(d.chunks); release(d.chunk_link); release(d.stage);
        d.n_warps = n_groups; d.engine = 2;
    }
    d.grid = grid; d.wpb = tpb / 32;
    int rc;
    if ((rc = ensure(ctx, d.glists, (size_t)n_groups * 2 * ctx->list_cap * sizeof(ulonglong2), false))) return rc;
    if ((rc = ensure(ctx, d.stage, (size_t)n_groups * ctx->hits_per_read * sizeof(bwb_hit), false))) return rc;
    {
        // chunk = 32 entries of 16 B (compact) or 32 B (wide); 3/4 of the pool is split into private
        // ranges (no atomics on the common path), 1/4 is shared between all groups
        uint64_t pool_bytes = (uint64_t)(ctx->heap_pool_mb > 0 ? ctx->heap_pool_mb : 8192) << 20;
        if (pool_bytes < (uint64_t)n_groups * 16 * 1024) pool_bytes = (uint64_t)n_groups * 16 * 1024;
        const size_t chunk_bytes = (size_t)CHUNK_ENTRIES * (wide ? 32 : 16);
        uint64_t n_chunks = pool_bytes / chunk_bytes;
        if (n_chunks > 0xfffffff0ull) n_chunks = 0xfffffff0ull;
        d.n_chunks = (uint32_t)n_chunks;
        d.chunks_per_warp = (uint32_t)((n_chunks - n_chunks / 4) / n_groups);      // 3/4 private, 1/4 shared
        if ((rc = ensure(ctx, d.chunks, pool_bytes, false))) return rc;
        if ((rc = ensure(ctx, d.chunk_link, (size_t)(pool_bytes / (CHUNK_ENTRIES * 16)) * 4, false))) return rc;
    }
    return BWB_OK;
}

#endif  // BWB_AB_ENGINES

// the eight instantiations of K4: coordinate width x seeding (-P) x slot recycling
typedef void (*K4Fn)(const LaneArgs);
K4Fn k4_fn(bool wide, bool pre, bool recycle) {
    static const K4Fn tab[8] = {k_search_l<false, false, false>, k_search_l<false, false, true>, k_search_l<false, true, false>,
                                k_search_l<false, true, true>,   k_search_l<true, false, false>, k_search_l<true, false, true>,
                                k_search_l<true, true, false>,   k_search_l<true, true, true>};
    return tab[(wide ? 4 : 0) | (pre ? 2 : 0) | (recycle ? 1 : 0)];
}

// size the persistent grid and the per-lane arena for the lane engine (K3 groups + K4 lanes)
int prepare_search_lane(bwb_ctx *ctx, Device &d, int nb, bool wide) {
    if (ctx->nbc <= 0 || ctx->nbc > 128 || ctx->bmap.size() != (size_t)(256 + nb))
        return fail(ctx, BWB_ERR_UNSUPPORTED, "%d reachable score buckets: the lane engine keeps the occupancy of at most 128 in registers", ctx->nbc);
    CU(cudaSetDevice(d.id));
    const int tpb = 128;
    // bucket heads [nbc][128], then score_of[128] (u16) and bucket_of[nb] (u8)
    const size_t smem = ((size_t)ctx->nbc * tpb * 4 + 256 + (size_t)nb + 15) & ~(size_t)15;
    // the -P instantiations differ only in how a read is seeded: same launch shape as the plain ones
    for (int v = 0; v < 8; v++) CU(cudaFuncSetAttribute(k4_fn(v & 4, v & 2, v & 1), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int bps = ctx->blocks_per_sm;
    if (bps <= 0) {
        int b0 = 1 << 30, b1 = 0;
        for (int v = 0; v < 4; v++) {          // the -P and recycling instantiations share the launch shape of the plain one
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, k4_fn(wide, v & 2, v & 1), tpb, smem));
            if (b1 < b0) b0 = b1;
        }
        b1 = b0;
        bps = b0 < b1 ? b0 : b1;
        if (bps <= 0) return fail(ctx, BWB_ERR_CUDA, "k_search_l does not fit on an SM");
    }
    {   // shared-memory carve-out: exactly what the resident blocks need (+1 KB per block the driver reserves), so that
        // what the compact bucket heads freed goes to the L1 (heap slots, lower-bound arrays and the second end of an
        // interval are L1 traffic)
        for (int v = 0; v < 4; v++) {
            cudaFuncAttributes fa;
            CU(cudaFuncGetAttributes(&fa, k4_fn(wide, v & 2, v & 1)));
            const size_t need = (size_t)bps * (fa.sharedSizeBytes + smem + 1024);
            int pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
            if (pct > 100) pct = 100;
            CU(cudaFuncSetAttribute(k4_fn(wide, v & 2, v & 1), cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        }
    }
    const int grid = bps * d.sm_count;
    const int n_lanes = grid * tpb;
    if (n_lanes != d.n_warps || d.engine != 0) {
        release(d.glists); release(d.chunks); release(d.chunk_link); release(d.stage);
        release(d.blk_link);
        d.n_warps = n_lanes; d.engine = 0;
    }
    d.grid = grid; d.wpb = tpb / 32; d.smem_bytes = smem;
    int rc;
    // K3 runs 8-lane groups in 256-thread blocks on its own persistent grid (occupancy query)
    {
        int bps3 = 0;
        const size_t smem3 = (size_t)(256 / GL) * (G_LIST_SMEM + 256);
        if (wide) {
            CU(cudaFuncSetAttribute(k_calc_d_g<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps3, k_calc_d_g<true>, 256, smem3));
        } else {
            CU(cudaFuncSetAttribute(k_calc_d_g<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps3, k_calc_d_g<false>, 256, smem3));
        }
        if (bps3 <= 0) bps3 = 1;
        d.grid3 = bps3 * d.sm_count;
    }
    const int n_groups3 = d.grid3 * (256 / GL);
    if ((rc = ensure(ctx, d.glists, (size_t)n_groups3 * 2 * ctx->list_cap * sizeof(ulonglong2), false))) return rc;
    {
        // arena of 32-byte slots: 3/4 split into private ranges, 1/4 shared in blocks of LBLK slots
        uint64_t pool_bytes = (uint64_t)ctx->heap_pool_mb << 20;
        if (ctx->heap_pool_mb <= 0) {
            if (d.chunks.p) pool_bytes = d.auto_pool_bytes;           // keep the size chosen at first use
            else {
                size_t fr = 0, tot = 0;
                CU(cudaMemGetInfo(&fr, &tot));
                pool_bytes = (uint64_t)fr / 10 * 7;
                if (pool_bytes > (128ull << 30)) pool_bytes = 128ull << 30;
                d.auto_pool_bytes = pool_bytes;
            }
        }
        uint64_t total = pool_bytes / 32;
        if (total < (uint64_t)n_lanes * 512) total = (uint64_t)n_lanes * 512;
        if (total > 0xfffff000ull) total = 0xfffff000ull;
        total &= ~(uint64_t)(LBLK - 1);
        // private : shared.  A quarter of the arena as private per-lane ranges is plenty for the common read now that
        // popped slots are recycled (chr21: < 100 KB live per read); the rest is the shared pool the few heavy reads
        // draw 8 KB blocks from -- at genome scale their heaps reach 2x10^5 live entries (6.6 MB), and with 3/4 of the
        // arena locked into private ranges 8 % of the reads ran out and waited for the 1/8-occupancy retry pass.
        const uint64_t priv_pct = ctx->arena_private_pct > 0 && ctx->arena_private_pct < 100 ? (uint64_t)ctx->arena_private_pct : 25;
        uint32_t spl = (uint32_t)(total * priv_pct / 100 / n_lanes);
        if (spl < 256) spl = 256;
        uint64_t priv = ((uint64_t)spl * n_lanes + LBLK - 1) & ~(uint64_t)(LBLK - 1);
        d.slots_per_lane = spl; d.total_slots = (uint32_t)total; d.priv_total = (uint32_t)priv;
        if ((rc = ensure(ctx, d.chunks, total * 32, false))) return rc;
        if ((rc = ensure(ctx, d.blk_link, (total / LBLK + 1) * 4, false))) return rc;
    }
    return BWB_OK;
}

#ifdef BWB_AB_ENGINES
// size the persistent grid and the per-warp scratch for K4
int prepare_search(bwb_ctx *ctx, Device &d, const SmemLayout &L, bool wide) {
    CU(cudaSetDevice(d.id));
    const int wpb = ctx->warps_per_block;
    const size_t smem = (size_t)wpb * L.per_warp;
    if (smem > 227 * 1024) return fail(ctx, BWB_ERR_ARG, "shared memory per block %zu exceeds 227 KB", smem);
    if (wide) CU(cudaFuncSetAttribute(k_align<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else CU(cudaFuncSetAttribute(k_align<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int bps = ctx->blocks_per_sm;
    if (bps <= 0) {
        if (wide) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_align<true>, wpb * 32, smem));
        else CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_align<false>, wpb * 32, smem));
        if (bps <= 0) return fail(ctx, BWB_ERR_CUDA, "k_align does not fit on an SM");
    }
    const int grid = bps * d.sm_count;
    const int n_warps = grid * wpb;
    d.smem_bytes = smem;
    if (n_warps != d.n_warps || wpb != d.wpb || d.engine != 1) {
        release(d.glists); release(d.chunks); release(d.chunk_link); release(d.stage);
        d.n_warps = n_warps; d.grid = grid; d.wpb = wpb; d.engine = 1;
    }
    d.grid = grid;
    int rc;
    if ((rc = ensure(ctx, d.glists, (size_t)n_warps * 2 * ctx->list_cap * sizeof(ulonglong2), false))) return rc;
    if ((rc = ensure(ctx, d.stage, (size_t)n_warps * ctx->hits_per_read * sizeof(bwb_hit), false))) return rc;
    if (!d.chunks.p) {
        const size_t chunk_bytes = (size_t)CHUNK_ENTRIES * 32;   // sized for the 32-byte entry format
        uint64_t n_chunks = ((uint64_t)(ctx->heap_pool_mb > 0 ? ctx->heap_pool_mb : 8192) << 20) / chunk_bytes;
        if (n_chunks < (uint64_t)n_warps * 8) n_chunks = (uint64_t)n_warps * 8;
        if (n_chunks > 0xfffffff0ull) n_chunks = 0xfffffff0ull;
        d.n_chunks = (uint32_t)n_chunks;
        d.chunks_per_warp = (uint32_t)((n_chunks - n_chunks / 4) / n_warps);   // 3/4 private, 1/4 shared overflow
        if ((rc = ensure(ctx, d.chunks, (size_t)n_chunks * chunk_bytes, false))) return rc;
        if ((rc = ensure(ctx, d.chunk_link, (size_t)n_chunks * 4, false))) return rc;
    }
    return BWB_OK;
}

#endif  // BWB_AB_ENGINES

// K0b: build the k-mer table on one device (after its blocks exist)
int build_ktab(bwb_ctx *ctx, Device &d) {
    CU(cudaSetDevice(d.id));
    if (d.ktab_w) { cudaFree(d.ktab_w); cudaFree(d.ktab_off); cudaFree(d.ktab_cnt); cudaFree(d.ktab_iv); }
    d.ktab_w = nullptr; d.ktab_off = d.ktab_cnt = nullptr; d.ktab_iv = nullptr;
    const bool wide = index_is_wide(ctx);
    const uint32_t nk = 1u << (2 * KTAB);
    const size_t nw_entries = ktab_level_off(KTAB + 1);
    const int wpb = 8, grid = d.sm_count * 2, n_warps = grid * wpb;
    void *gl = nullptr;
    unsigned char *sm = nullptr;
    CU(cudaMalloc(&d.ktab_w, nw_entries * 8));
    CU(cudaMalloc(&d.ktab_off, (size_t)nk * 4));
    CU(cudaMalloc(&d.ktab_cnt, (size_t)nk * 4));
    CU(cudaMalloc(&gl, (size_t)n_warps * 2 * ctx->list_cap * sizeof(ulonglong2)));
    CU(cudaMalloc(&sm, 64));
    unsigned long long cap = std::max<unsigned long long>(8ull << 20, ctx->length / 8);
    const size_t pair = wide ? 16 : 8;
    for (int attempt = 0; attempt < 6; attempt++) {
        CU(cudaMalloc(&d.ktab_iv, cap * pair));
        CU(cudaMemsetAsync(d.ktab_w, 0, nw_entries * 8, d.stream));
        CU(cudaMemsetAsync(sm, 0, 64, d.stream));
        KtabArgs a;
        memset(&a, 0, sizeof a);
        a.ix = make_view(ctx, d); a.glists = gl; a.list_cap = ctx->list_cap;
        a.w = d.ktab_w; a.koff = d.ktab_off; a.kcnt = d.ktab_cnt; a.iv = d.ktab_iv; a.iv_cap = cap;
        a.cursor = (unsigned long long *)sm; a.status = (uint32_t *)(sm + 8);
        const size_t smem = (size_t)wpb * LIST_SMEM_BYTES;
        if (wide) k_kmer_table<uint64_t><<<grid, wpb * 32, smem, d.stream>>>(a);
        else k_kmer_table<uint32_t><<<grid, wpb * 32, smem, d.stream>>>(a);
        CU(cudaGetLastError());
        unsigned long long used = 0;
        uint32_t st = 0;
        CU(cudaMemcpyAsync(&used, sm, 8, cudaMemcpyDeviceToHost, d.stream));
        CU(cudaMemcpyAsync(&st, sm + 8, 4, cudaMemcpyDeviceToHost, d.stream));
        CU(cudaStreamSynchronize(d.stream));
        if (st) {       // a top-of-tree list longer than list_cap: run without the table rather than fail
            cudaFree(d.ktab_iv); cudaFree(d.ktab_w); cudaFree(d.ktab_off); cudaFree(d.ktab_cnt);
            d.ktab_w = nullptr; d.ktab_off = d.ktab_cnt = nullptr; d.ktab_iv = nullptr;
            break;
        }
        if (used <= cap) break;
        CU(cudaFree(d.ktab_iv));
        d.ktab_iv = nullptr;
        cap = used + 1024;
    }
    cudaFree(gl); cudaFree(sm);
    return BWB_OK;
}

// put a row-ordered seed table (sizes + concatenated L,U pairs) on every device
int precalc_install(bwb_ctx *ctx, int is_multiref) {
    const uint64_t total = ctx->pre_lu_h.size() / 2;
    if (total > 0xfffffff0ull) return fail(ctx, BWB_ERR_UNSUPPORTED, "seed table with %llu intervals", (unsigned long long)total);
    std::vector<uint32_t> &off = ctx->pre_off_h;
    off.resize(NUM_PRECALC);
    uint64_t acc = 0;
    for (uint32_t x = 0; x < NUM_PRECALC; x++) { off[x] = (uint32_t)acc; acc += ctx->pre_cnt_h[x]; }
    if (acc != total) return fail(ctx, BWB_ERR_ARG, "seed table: sizes sum to %llu, %llu intervals given", (unsigned long long)acc, (unsigned long long)total);
    for (auto &d : ctx->dev) {
        CU(cudaSetDevice(d.id));
        if (d.pre_off) { cudaFree(d.pre_off); cudaFree(d.pre_cnt); cudaFree(d.pre_iv); d.pre_off = d.pre_cnt = nullptr; d.pre_iv = nullptr; }
        CU(cudaMalloc(&d.pre_off, (size_t)NUM_PRECALC * 4));
        CU(cudaMalloc(&d.pre_cnt, (size_t)NUM_PRECALC * 4));
        CU(cudaMalloc(&d.pre_iv, (total + 1) * sizeof(ulonglong2)));
        CU(cudaMemcpyAsync(d.pre_off, off.data(), (size_t)NUM_PRECALC * 4, cudaMemcpyHostToDevice, d.stream));
        CU(cudaMemcpyAsync(d.pre_cnt, ctx->pre_cnt_h.data(), (size_t)NUM_PRECALC * 4, cudaMemcpyHostToDevice, d.stream));
        if (total) CU(cudaMemcpyAsync(d.pre_iv, ctx->pre_lu_h.data(), total * 16, cudaMemcpyHostToDevice, d.stream));
        CU(cudaStreamSynchronize(d.stream));
    }
    ctx->have_pre = true;
    ctx->pre_multiref = is_multiref ? 1 : 0;
    return BWB_OK;
}

// K0c on device 0 -> row-ordered host copy
int precalc_compute(bwb_ctx *ctx, int is_multiref) {
    Device &d = ctx->dev[0];
    CU(cudaSetDevice(d.id));
    const bool wide = index_is_wide(ctx);
    const int wpb = 8, grid = d.sm_count * 2, n_warps = grid * wpb;
    uint32_t *off = nullptr, *cnt = nullptr;
    void *gl = nullptr;
    unsigned char *sm = nullptr;
    ulonglong2 *iv = nullptr;
    CU(cudaMalloc(&off, (size_t)NUM_PRECALC * 4));
    CU(cudaMalloc(&cnt, (size_t)NUM_PRECALC * 4));
    CU(cudaMalloc(&gl, (size_t)n_warps * 2 * ctx->list_cap * sizeof(ulonglong2)));
    CU(cudaMalloc(&sm, 64));
    unsigned long long cap = 32ull << 20, used = 0;
    uint32_t st = 0;
    int rc = BWB_OK;
    for (int attempt = 0; attempt < 3; attempt++) {
        CU(cudaMalloc(&iv, cap * sizeof(ulonglong2)));
        CU(cudaMemsetAsync(sm, 0, 64, d.stream));
        PrecalcArgs a;
        memset(&a, 0, sizeof a);
        a.ix = make_view(ctx, d); a.glists = gl; a.list_cap = ctx->list_cap; a.is_multiref = is_multiref;
        a.off = off; a.cnt = cnt; a.iv = iv; a.iv_cap = cap;
        a.cursor = (unsigned long long *)sm; a.status = (uint32_t *)(sm + 8);
        const size_t smem = (size_t)wpb * LIST_SMEM_BYTES;
        if (wide) k_precalc<uint64_t><<<grid, wpb * 32, smem, d.stream>>>(a);
        else k_precalc<uint32_t><<<grid, wpb * 32, smem, d.stream>>>(a);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(&used, sm, 8, cudaMemcpyDeviceToHost, d.stream));
        CU(cudaMemcpyAsync(&st, sm + 8, 4, cudaMemcpyDeviceToHost, d.stream));
        CU(cudaStreamSynchronize(d.stream));
        if (st || used <= cap) break;
        CU(cudaFree(iv));
        iv = nullptr;
        cap = used + 1024;
    }
    if (st) rc = fail(ctx, -(int)st, "seed table: interval list exceeded list_cap=%d (raise it with bwb_set_option)", ctx->list_cap);
    else if (used > 0xfffffff0ull) rc = fail(ctx, BWB_ERR_UNSUPPORTED, "seed table with %llu intervals", used);
    if (!rc) {
        std::vector<uint32_t> hoff(NUM_PRECALC);
        std::vector<ulonglong2> pool(used);
        ctx->pre_cnt_h.resize(NUM_PRECALC);
        CU(cudaMemcpy(hoff.data(), off, (size_t)NUM_PRECALC * 4, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(ctx->pre_cnt_h.data(), cnt, (size_t)NUM_PRECALC * 4, cudaMemcpyDeviceToHost));
        if (used) CU(cudaMemcpy(pool.data(), iv, used * sizeof(ulonglong2), cudaMemcpyDeviceToHost));
        ctx->pre_lu_h.resize(used * 2);
        uint64_t w = 0;
        for (uint32_t x = 0; x < NUM_PRECALC; x++)
            for (uint32_t k = 0; k < ctx->pre_cnt_h[x]; k++, w++) {
                ctx->pre_lu_h[2 * w] = pool[hoff[x] + k].x;
                ctx->pre_lu_h[2 * w + 1] = pool[hoff[x] + k].y;
            }
    }
    cudaFree(off); cudaFree(cnt); cudaFree(gl); cudaFree(sm);
    if (iv) cudaFree(iv);
    return rc;
}

void set_ktab(const bwb_ctx *ctx, const Device &d, CalcArgs &c) {
    const bool on = ctx->use_ktab && d.ktab_w && d.ktab_iv;
    c.ktab_w = on ? d.ktab_w : nullptr;
    c.ktab_off = d.ktab_off; c.ktab_cnt = d.ktab_cnt; c.ktab_iv = d.ktab_iv;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
extern "C" {

void bwb_default_params(bwb_params *p) {   // align.c:22-38
    memset(p, 0, sizeof *p);
    p->gape_score = 4; p->gapo_score = 11; p->mm_score = 3;
    p->max_diff = 0; p->max_gape = 6; p->max_gapo = 1;
    p->seed_length = 32; p->max_diff_seed = 2; p->max_entries = 3000000;
    p->use_precalc = 0; p->matched_Ncontig = 0; p->is_multiref = 1;
    p->max_best = 30; p->no_indel_length = 5; p->n_threads = 1;
}

const char *bwb_last_error(const bwb_ctx *ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

bwb_ctx *bwb_create(const int *devices, int ndev) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        fail(nullptr, BWB_ERR_CUDA, "no usable CUDA device: %s (this library has no CPU fallback)",
             e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return nullptr;
    }
    bwb_ctx *ctx = new bwb_ctx();
    std::vector<int> ids;
    if (!devices || ndev <= 0) ids.push_back(0);
    else ids.assign(devices, devices + ndev);
    for (int id : ids) {
        if (id < 0 || id >= count) {
            fail(nullptr, BWB_ERR_ARG, "device %d out of range (%d visible)", id, count);
            delete ctx;
            return nullptr;
        }
        Device d;
        d.id = id;
        cudaDeviceProp prop;
        if (cudaSetDevice(id) != cudaSuccess || cudaGetDeviceProperties(&prop, id) != cudaSuccess ||
            cudaStreamCreateWithFlags(&d.own_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaMallocHost((void **)&d.h_small, 64 * sizeof(unsigned long long)) != cudaSuccess ||
            cudaEventCreate(&d.ev0) != cudaSuccess || cudaEventCreate(&d.ev1) != cudaSuccess ||
            cudaEventCreate(&d.evm) != cudaSuccess) {
            fail(nullptr, BWB_ERR_CUDA, "cannot initialise device %d: %s", id, cudaGetErrorString(cudaGetLastError()));
            delete ctx;
            return nullptr;
        }
        d.stream = d.own_stream;
        d.sm_count = prop.multiProcessorCount;
        ctx->dev.push_back(d);
    }
    return ctx;
}

void bwb_destroy(bwb_ctx *ctx) {
    if (!ctx) return;
    for (auto &d : ctx->dev) {
        cudaSetDevice(d.id);
        cudaDeviceSynchronize();
        if (d.blocks) cudaFree(d.blocks);
        if (d.sa) cudaFree(d.sa);
        if (d.ktab_w) cudaFree(d.ktab_w);
        if (d.ktab_off) cudaFree(d.ktab_off);
        if (d.ktab_cnt) cudaFree(d.ktab_cnt);
        if (d.ktab_iv) cudaFree(d.ktab_iv);
        if (d.pre_off) cudaFree(d.pre_off);
        if (d.pre_cnt) cudaFree(d.pre_cnt);
        if (d.pre_iv) cudaFree(d.pre_iv);
        DevBuf *bufs[] = {&d.glists, &d.chunks, &d.chunk_link, &d.stage, &d.seq, &d.offsets, &d.read_off, &d.read_cnt,
                          &d.ordered_off, &d.unordered, &d.ordered, &d.cub_tmp, &d.small, &d.d_main, &d.d_seed, &d.pool, &d.loc, &d.pk_main, &d.pk_seed, &d.n_count, &d.nxt, &d.blk_link, &d.heads, &d.order, &d.retry, &d.bmap, &d.seed_src, &d.seed_ext};
        for (DevBuf *b : bufs) release(*b);
        if (d.h_small) cudaFreeHost(d.h_small);
        if (d.ev0) cudaEventDestroy(d.ev0);
        if (d.ev1) cudaEventDestroy(d.ev1);
        if (d.evm) cudaEventDestroy(d.evm);
        if (d.own_stream) cudaStreamDestroy(d.own_stream);
    }
    delete ctx;
}

int bwb_device_count(const bwb_ctx *ctx) { return ctx ? (int)ctx->dev.size() : 0; }

int bwb_set_option(bwb_ctx *ctx, const char *key, long long value) {
    if (!ctx || !key) return BWB_ERR_ARG;
    std::string k(key);
    if (value <= 0 && k != "blocks_per_sm" && k != "force_wide" && k != "engine" && k != "heap_pool_mb" && k != "kmer_table" && k != "heavy_first" && k != "hit_cap0" && k != "index_wide" && k != "index_chunk" && k != "arena_private_pct" && k != "recycle" && k != "throttle_pct" && k != "seed_carry") return fail(ctx, BWB_ERR_ARG, "option %s needs a positive value", key);
    if (k == "heap_pool_mb") ctx->heap_pool_mb = value;
    else if (k == "list_cap") ctx->list_cap = (int)(value < SL + 4 ? SL + 4 : value);
    else if (k == "hits_per_read") ctx->hits_per_read = (int)value;
    else if (k == "warps_per_block") {
        if (value > 8) return fail(ctx, BWB_ERR_ARG, "warps_per_block must be 1..8");
        ctx->warps_per_block = (int)value;
    } else if (k == "blocks_per_sm") ctx->blocks_per_sm = (int)value;
    else if (k == "force_wide") {
        const int nv = value > 1 ? 0 : 1;
        if (nv != ctx->force_wide && ctx->have_index)
            return fail(ctx, BWB_ERR_ARG, "force_wide must be set before the index is uploaded");
        ctx->force_wide = nv;
    }
    else if (k == "kmer_table") ctx->use_ktab = value > 1 ? 0 : 1;
    else if (k == "heavy_first") ctx->heavy_first = value > 1 ? 0 : 1;
    else if (k == "hit_cap0") ctx->hit_cap0 = value;
    else if (k == "arena_private_pct") ctx->arena_private_pct = (int)value;
    else if (k == "throttle_pct") { ctx->throttle_pct = (int)value; ctx->throttle_forced = 1; return BWB_OK; }
    else if (k == "seed_carry") { ctx->seed_carry = value == 1 ? 1 : 0; ctx->carry_len = 0; return BWB_OK; }     // (re)starts a run
    else if (k == "recycle") { ctx->recycle = (value == 1 || value == 2) ? (int)value : 0; return BWB_OK; }
    else if (k == "index_wide") { ctx->index_wide = value == 1 ? 1 : 0; return BWB_OK; }
    else if (k == "index_chunk") { ctx->index_chunk = value > 0 ? value : 0; return BWB_OK; }
    else if (k == "engine") {
#ifdef BWB_AB_ENGINES
        ctx->engine = (value == 1 || value == 2) ? (int)value : 0;
#else
        if (value == 1 || value == 2)
            return fail(ctx, BWB_ERR_UNSUPPORTED, "the round-1 A/B engines (1 = warp per read, 2 = 8 lanes per read) are only built with -DBWB_AB_ENGINES");
        ctx->engine = 0;
#endif
    }
    else return fail(ctx, BWB_ERR_ARG, "unknown option %s", key);
    for (auto &d : ctx->dev) {       // scratch is re-sized lazily
        cudaSetDevice(d.id);
        cudaDeviceSynchronize();
        release(d.glists); release(d.chunks); release(d.chunk_link); release(d.stage);
        d.n_warps = 0;
    }
    return BWB_OK;
}

int bwb_set_stream(bwb_ctx *ctx, int dev_slot, void *cuda_stream) {
    if (!ctx || dev_slot < 0 || dev_slot >= (int)ctx->dev.size()) return BWB_ERR_ARG;
    ctx->dev[dev_slot].stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->dev[dev_slot].own_stream;
    return BWB_OK;
}

// ---- index ------------------------------------------------------------------------------------
int bwb_index_upload(bwb_ctx *ctx, uint64_t length, uint64_t sa0_index, const uint64_t C[17], const uint32_t *bwt,
                     uint64_t num_words, const uint64_t *O, uint64_t num_occ) {
    if (!ctx || !C || !bwt || !O || length < 2) return fail(ctx, BWB_ERR_ARG, "bwb_index_upload: bad argument");
    if (num_words < (length + 7) / 8 || num_occ < (length + 127) / 128)
        return fail(ctx, BWB_ERR_ARG, "bwb_index_upload: arrays shorter than length implies");
    if (length >= (1ull << 40)) return fail(ctx, BWB_ERR_ARG, "index longer than 2^40 rows");
    ctx->length = length;
    ctx->have_sa = false;
    ctx->have_index = false;
    // a -P seed table belongs to the index it was computed on: drop it with the old index
    ctx->have_pre = false;
    ctx->pre_cnt_h.clear(); ctx->pre_off_h.clear(); ctx->pre_lu_h.clear();
    ctx->sa0 = sa0_index;
    ctx->num_blocks = (length + 127) / 128;
    memcpy(ctx->C, C, sizeof ctx->C);
    for (auto &d : ctx->dev) {
        CU(cudaSetDevice(d.id));
        if (d.pre_off) { cudaFree(d.pre_off); cudaFree(d.pre_cnt); cudaFree(d.pre_iv); d.pre_off = d.pre_cnt = nullptr; d.pre_iv = nullptr; }
        if (d.blocks) { CU(cudaFree(d.blocks)); d.blocks = nullptr; }
        DevTmp tmp;                                  // freed on every exit path
        uint32_t *d_bwt = nullptr, *d_err = nullptr;
        uint64_t *d_O = nullptr;
        CU(cudaMalloc(&d.blocks, ctx->num_blocks * 128));
        CU(tmp.alloc(&d_bwt, num_words * 4));
        CU(tmp.alloc(&d_O, num_occ * 16 * 8));
        CU(tmp.alloc(&d_err, 4));
        CU(cudaMemcpyAsync(d_bwt, bwt, num_words * 4, cudaMemcpyHostToDevice, d.stream));
        CU(cudaMemcpyAsync(d_O, O, num_occ * 16 * 8, cudaMemcpyHostToDevice, d.stream));
        CU(cudaMemsetAsync(d_err, 0, 4, d.stream));
        const unsigned tpb = 128;
        const unsigned grid = (unsigned)((ctx->num_blocks + tpb - 1) / tpb);
        k_relayout<<<grid, tpb, 0, d.stream>>>(d_bwt, num_words, d_O, num_occ, sa0_index,
                                              reinterpret_cast<uint32_t *>(d.blocks), ctx->num_blocks, d_err);
        CU(cudaGetLastError());
        uint32_t herr = 0;
        CU(cudaMemcpyAsync(&herr, d_err, 4, cudaMemcpyDeviceToHost, d.stream));
        CU(cudaStreamSynchronize(d.stream));
        if (herr) return fail(ctx, BWB_ERR_ARG, "a per-code rank counter exceeds 2^32 (index too large for u32 checkpoints)");
    }
    ctx->have_index = true;
    for (auto &d : ctx->dev) {
        int rc = build_ktab(ctx, d);
        if (rc) return rc;
    }
    return BWB_OK;
}

int bwb_index_load_file(bwb_ctx *ctx, const char *bwt_path) {
    if (!ctx || !bwt_path) return BWB_ERR_ARG;
    bwb_host::HostIndex ix;
    if (bwb_host::read_bwt_file(bwt_path, ix, false)) return fail(ctx, BWB_ERR_IO, "cannot read %s", bwt_path);
    return bwb_index_upload(ctx, ix.length, ix.sa0_index, ix.C, ix.bwt.data(), ix.num_words, ix.O.data(), ix.num_occ);
}

int bwb_sa_upload(bwb_ctx *ctx, const uint64_t *SA, uint64_t num_sa) {
    if (!ctx || !SA) return BWB_ERR_ARG;
    if (!ctx->have_index) return fail(ctx, BWB_ERR_NO_INDEX, "bwb_sa_upload before bwb_index_upload");
    if (num_sa < (ctx->length + 31) / 32) return fail(ctx, BWB_ERR_ARG, "sampled SA shorter than ceil(length/32)");
    for (auto &d : ctx->dev) {
        CU(cudaSetDevice(d.id));
        if (d.sa) { CU(cudaFree(d.sa)); d.sa = nullptr; }
        CU(cudaMalloc(&d.sa, num_sa * 8));
        CU(cudaMemcpyAsync(d.sa, SA, num_sa * 8, cudaMemcpyHostToDevice, d.stream));
        CU(cudaStreamSynchronize(d.stream));
    }
    ctx->num_sa = num_sa;
    ctx->have_sa = true;
    return BWB_OK;
}

int bwb_index_load_file_sa(bwb_ctx *ctx, const char *bwt_path) {
    if (!ctx || !bwt_path) return BWB_ERR_ARG;
    bwb_host::HostIndex ix;
    if (bwb_host::read_bwt_file(bwt_path, ix, true)) return fail(ctx, BWB_ERR_IO, "cannot read %s (with SA)", bwt_path);
    int rc = bwb_index_upload(ctx, ix.length, ix.sa0_index, ix.C, ix.bwt.data(), ix.num_words, ix.O.data(), ix.num_occ);
    if (rc) return rc;
    return bwb_sa_upload(ctx, ix.SA.data(), ix.num_sa);
}

// ---- -P seed table ------------------------------------------------------------------------
int bwb_precalc_build(bwb_ctx *ctx, int is_multiref) {
    if (!ctx) return BWB_ERR_ARG;
    if (!ctx->have_index) return fail(ctx, BWB_ERR_NO_INDEX, "bwb_precalc_build before bwb_index_upload");
    int rc = precalc_compute(ctx, is_multiref ? 1 : 0);
    if (rc) return rc;
    return precalc_install(ctx, is_multiref);
}

int bwb_precalc_upload(bwb_ctx *ctx, const int32_t *sizes, const uint64_t *intervals_LU, uint64_t n_intervals, int is_multiref) {
    if (!ctx || !sizes || (n_intervals && !intervals_LU)) return BWB_ERR_ARG;
    if (!ctx->have_index) return fail(ctx, BWB_ERR_NO_INDEX, "bwb_precalc_upload before bwb_index_upload");
    ctx->pre_cnt_h.resize(NUM_PRECALC);
    for (uint32_t x = 0; x < NUM_PRECALC; x++) {
        if (sizes[x] < 0) return fail(ctx, BWB_ERR_ARG, "seed table row %u has negative size", x);
        ctx->pre_cnt_h[x] = (uint32_t)sizes[x];
    }
    ctx->pre_lu_h.assign(intervals_LU, intervals_LU + 2 * n_intervals);
    return precalc_install(ctx, is_multiref);
}

int bwb_precalc_load_file(bwb_ctx *ctx, const char *pre_path, int is_multiref) {
    if (!ctx || !pre_path) return BWB_ERR_ARG;
    if (!ctx->have_index) return fail(ctx, BWB_ERR_NO_INDEX, "bwb_precalc_load_file before bwb_index_upload");
    if (bwb_host::read_pre_file(pre_path, ctx->pre_cnt_h, ctx->pre_lu_h)) return fail(ctx, BWB_ERR_IO, "cannot read %s", pre_path);
    return precalc_install(ctx, is_multiref);
}

int bwb_precalc_write(bwb_ctx *ctx, const char *pre_path) {
    if (!ctx || !pre_path) return BWB_ERR_ARG;
    if (!ctx->have_pre) return fail(ctx, BWB_ERR_ARG, "no seed table to write");
    if (bwb_host::write_pre_file(pre_path, ctx->pre_cnt_h, ctx->pre_lu_h)) return fail(ctx, BWB_ERR_IO, "cannot write %s", pre_path);
    return BWB_OK;
}

uint64_t bwb_precalc_num_intervals(const bwb_ctx *ctx) { return ctx && ctx->have_pre ? ctx->pre_lu_h.size() / 2 : 0; }

int bwb_precalc_row(const bwb_ctx *ctx, uint32_t row, uint64_t *intervals_LU, uint32_t cap, uint32_t *n) {
    if (!ctx || !n || row >= NUM_PRECALC || !ctx->have_pre) return BWB_ERR_ARG;
    *n = ctx->pre_cnt_h[row];
    const uint64_t base = ctx->pre_off_h[row];
    for (uint32_t k = 0; k < *n && k < cap && intervals_LU; k++) {
        intervals_LU[2 * k] = ctx->pre_lu_h[2 * (base + k)];
        intervals_LU[2 * k + 1] = ctx->pre_lu_h[2 * (base + k) + 1];
    }
    return BWB_OK;
}

uint64_t bwb_index_num_blocks(const bwb_ctx *ctx) { return ctx && ctx->have_index ? ctx->num_blocks : 0; }

int bwb_index_download_blocks(bwb_ctx *ctx, void *out) {
    if (!ctx || !out) return BWB_ERR_ARG;
    if (!ctx->have_index) return fail(ctx, BWB_ERR_NO_INDEX, "no index uploaded");
    Device &d = ctx->dev[0];
    CU(cudaSetDevice(d.id));
    CU(cudaMemcpy(out, d.blocks, ctx->num_blocks * 128, cudaMemcpyDeviceToHost));
    return BWB_OK;
}

// ---- K1 ---------------------------------------------------------------------------------------
int bwb_occ(bwb_ctx *ctx, const uint8_t *code, const uint64_t *pos, uint64_t n, uint64_t *out) {
    if (!ctx || !code || !pos || !out) return BWB_ERR_ARG;
    if (!ctx->have_index) return fail(ctx, BWB_ERR_NO_INDEX, "no index uploaded");
    if (n == 0) return BWB_OK;
    Device &d = ctx->dev[0];
    CU(cudaSetDevice(d.id));
    uint8_t *dc; uint64_t *dp, *dout;
    DevTmp tmp;
    CU(tmp.alloc(&dc, n)); CU(tmp.alloc(&dp, n * 8)); CU(tmp.alloc(&dout, n * 8));
    CU(cudaMemcpyAsync(dc, code, n, cudaMemcpyHostToDevice, d.stream));
    CU(cudaMemcpyAsync(dp, pos, n * 8, cudaMemcpyHostToDevice, d.stream));
    if (index_is_wide(ctx)) k_occ<uint64_t><<<(unsigned)((n + 255) / 256), 256, 0, d.stream>>>(make_view(ctx, d), dc, dp, n, dout);
    else k_occ<uint32_t><<<(unsigned)((n + 255) / 256), 256, 0, d.stream>>>(make_view(ctx, d), dc, dp, n, dout);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, dout, n * 8, cudaMemcpyDeviceToHost, d.stream));
    CU(cudaStreamSynchronize(d.stream));
    return BWB_OK;
}

int bwb_occ_alphabet(bwb_ctx *ctx, const uint64_t *pos, uint64_t n, int inc, uint64_t *out) {
    if (!ctx || !pos || !out) return BWB_ERR_ARG;
    if (!ctx->have_index) return fail(ctx, BWB_ERR_NO_INDEX, "no index uploaded");
    if (n == 0) return BWB_OK;
    Device &d = ctx->dev[0];
    CU(cudaSetDevice(d.id));
    uint64_t *dp, *dout;
    DevTmp tmp;
    CU(tmp.alloc(&dp, n * 8)); CU(tmp.alloc(&dout, n * 16 * 8));
    CU(cudaMemcpyAsync(dp, pos, n * 8, cudaMemcpyHostToDevice, d.stream));
    if (index_is_wide(ctx)) k_occ_alphabet<uint64_t><<<(unsigned)((n * 16 + 255) / 256), 256, 0, d.stream>>>(make_view(ctx, d), dp, n, (uint32_t)inc, dout);
    else k_occ_alphabet<uint32_t><<<(unsigned)((n * 16 + 255) / 256), 256, 0, d.stream>>>(make_view(ctx, d), dp, n, (uint32_t)inc, dout);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, dout, n * 16 * 8, cudaMemcpyDeviceToHost, d.stream));
    CU(cudaStreamSynchronize(d.stream));
    return BWB_OK;
}

int bwb_occ_bench(bwb_ctx *ctx, uint64_t n, uint64_t seed, int mode, int iters, float *ms_per_launch, uint64_t *checksum) {
    if (!ctx || n == 0 || iters <= 0) return BWB_ERR_ARG;
    if (!ctx->have_index) return fail(ctx, BWB_ERR_NO_INDEX, "no index uploaded");
    Device &d = ctx->dev[0];
    CU(cudaSetDevice(d.id));
    unsigned long long *dsink;
    DevTmp tmp;
    CU(tmp.alloc(&dsink, 8));
    CU(cudaMemsetAsync(dsink, 0, 8, d.stream));
    const int chain = mode >= 2 ? 8 : 1;          // modes 2/3 = modes 0/1 with 8 dependent queries per thread
    const int m = mode & 1;
    const uint64_t threads = m ? n * 16 : n;
    const unsigned grid = (unsigned)((threads + 255) / 256);
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    for (int it = -2; it < iters; it++) {          // two warm-up launches
        if (it == 0) CU(cudaEventRecord(e0, d.stream));
        if (m) k_occ_bench<1><<<grid, 256, 0, d.stream>>>(make_view(ctx, d), n, seed + it, chain, dsink);
        else k_occ_bench<0><<<grid, 256, 0, d.stream>>>(make_view(ctx, d), n, seed + it, chain, dsink);
    }
    CU(cudaEventRecord(e1, d.stream));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(d.stream));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    if (ms_per_launch) *ms_per_launch = ms / iters;
    unsigned long long h = 0;
    CU(cudaMemcpy(&h, dsink, 8, cudaMemcpyDeviceToHost));
    if (checksum) *checksum = h;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return BWB_OK;
}

// ---- K2 / K3 ----------------------------------------------------------------------------------
static int run_list_kernel(bwb_ctx *ctx, int which, const uint8_t *seq, const uint64_t *offsets, uint64_t n_reads,
                           int use_len, uint32_t *counts, uint64_t **intervals, uint64_t *n_intervals, int32_t *out_d) {
    if (!ctx || !seq || !offsets) return BWB_ERR_ARG;
    if (!ctx->have_index) return fail(ctx, BWB_ERR_NO_INDEX, "no index uploaded");
    int max_len, rc;
    if ((rc = check_reads(ctx, offsets, n_reads, max_len))) return rc;
    if (n_reads == 0) { if (n_intervals) *n_intervals = 0; if (intervals) *intervals = nullptr; return BWB_OK; }
    Device &d = ctx->dev[0];
    CU(cudaSetDevice(d.id));
    const int wpb = 4;
    const int grid = d.sm_count * 4;
    const int n_warps = grid * wpb;
    const uint64_t total = offsets[n_reads] - offsets[0];
    ListArgs a;
    memset(&a, 0, sizeof a);
    a.ix = make_view(ctx, d);
    a.n_reads = (uint32_t)n_reads;
    a.list_cap = ctx->list_cap;
    a.max_len = max_len;
    a.use_len = use_len;
    uint8_t *dseq; uint64_t *doff; ulonglong2 *gl; uint32_t *dstatus;
    const bool wide = index_is_wide(ctx);
    DevTmp tmp;
    CU(tmp.alloc(&dseq, total + 16)); CU(tmp.alloc(&doff, (n_reads + 1) * 8));
    CU(tmp.alloc(&gl, (size_t)n_warps * 2 * a.list_cap * sizeof(ulonglong2)));
    CU(tmp.alloc(&dstatus, 8));
    std::vector<uint64_t> rel(n_reads + 1);
    for (uint64_t r = 0; r <= n_reads; r++) rel[r] = offsets[r] - offsets[0];
    CU(cudaMemcpyAsync(dseq, seq + offsets[0], total, cudaMemcpyHostToDevice, d.stream));
    CU(cudaMemcpyAsync(doff, rel.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice, d.stream));
    CU(cudaMemsetAsync(dstatus, 0, 8, d.stream));
    a.seq = dseq; a.offsets = doff; a.glists = (void *)gl; a.status = dstatus;
    uint32_t hstatus = 0;
    if (which == 2) {
        unsigned long long *dcur, *droff; uint32_t *drcnt; ulonglong2 *dout;
        unsigned long long cap = n_reads * 8 + 4096;
        CU(tmp.alloc(&dcur, 8)); CU(tmp.alloc(&droff, n_reads * 8)); CU(tmp.alloc(&drcnt, n_reads * 4));
        for (int attempt = 0;; attempt++) {
            CU(tmp.alloc(&dout, cap * sizeof(ulonglong2)));
            CU(cudaMemsetAsync(dcur, 0, 8, d.stream));
            a.out_iv = dout; a.out_cap = cap; a.out_cursor = dcur; a.read_off = droff; a.read_cnt = drcnt;
            const size_t smem = (size_t)wpb * (2 * SL * sizeof(ulonglong2) + ((max_len + 15) & ~15));
            if (wide) k_exact<uint64_t><<<grid, wpb * 32, smem, d.stream>>>(a);
            else k_exact<uint32_t><<<grid, wpb * 32, smem, d.stream>>>(a);
            CU(cudaGetLastError());
            unsigned long long used = 0;
            CU(cudaMemcpyAsync(&used, dcur, 8, cudaMemcpyDeviceToHost, d.stream));
            CU(cudaMemcpyAsync(&hstatus, dstatus, 4, cudaMemcpyDeviceToHost, d.stream));
            CU(cudaStreamSynchronize(d.stream));
            if (used <= cap || attempt > 4) {
                if (used > cap) return fail(ctx, BWB_ERR_CAPACITY, "exact-match output does not fit");
                std::vector<unsigned long long> roff(n_reads);
                std::vector<ulonglong2> iv(used ? used : 1);
                CU(cudaMemcpy(roff.data(), droff, n_reads * 8, cudaMemcpyDeviceToHost));
                CU(cudaMemcpy(counts, drcnt, n_reads * 4, cudaMemcpyDeviceToHost));
                CU(cudaMemcpy(iv.data(), dout, used * sizeof(ulonglong2), cudaMemcpyDeviceToHost));
                uint64_t *res = (uint64_t *)malloc((used ? used : 1) * 16);
                uint64_t w = 0;
                for (uint64_t r = 0; r < n_reads; r++)
                    for (uint32_t k = 0; k < counts[r]; k++) { res[2 * w] = iv[roff[r] + k].x; res[2 * w + 1] = iv[roff[r] + k].y; w++; }
                *intervals = res;
                *n_intervals = w;
                break;
            }
            tmp.release(dout);
            cap = used + 4096;
        }
    } else {
        int32_t *dd;
        const size_t nd = 2 * (total + n_reads);
        CU(tmp.alloc(&dd, nd * 4 + 16));
        CU(cudaMemsetAsync(dd, 0, nd * 4, d.stream));
        a.out_d = dd;
        const size_t smem = (size_t)wpb * (2 * SL * sizeof(ulonglong2) + (((max_len + 1) * 8 + 15) & ~15) + ((max_len + 15) & ~15));
        if (wide) {
            CU(cudaFuncSetAttribute(k_calc_d<uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_calc_d<uint64_t><<<grid, wpb * 32, smem, d.stream>>>(a);
        } else {
            CU(cudaFuncSetAttribute(k_calc_d<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_calc_d<uint32_t><<<grid, wpb * 32, smem, d.stream>>>(a);
        }
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(out_d, dd, nd * 4, cudaMemcpyDeviceToHost, d.stream));
        CU(cudaMemcpyAsync(&hstatus, dstatus, 4, cudaMemcpyDeviceToHost, d.stream));
        CU(cudaStreamSynchronize(d.stream));
    }
    if (hstatus) return fail(ctx, -(int)hstatus, "interval list exceeded list_cap=%d (raise it with bwb_set_option)", ctx->list_cap);
    return BWB_OK;
}

int bwb_exact_match(bwb_ctx *ctx, const uint8_t *seq, const uint64_t *offsets, uint64_t n_reads, uint32_t *counts,
                    uint64_t **intervals_LU, uint64_t *n_intervals) {
    if (!counts || !intervals_LU || !n_intervals) return BWB_ERR_ARG;
    return run_list_kernel(ctx, 2, seq, offsets, n_reads, 0, counts, intervals_LU, n_intervals, nullptr);
}

int bwb_calculate_d(bwb_ctx *ctx, const uint8_t *seq, const uint64_t *offsets, uint64_t n_reads, int use_len, int32_t *out) {
    if (!out) return BWB_ERR_ARG;
    return run_list_kernel(ctx, 3, seq, offsets, n_reads, use_len, nullptr, nullptr, nullptr, out);
}

// K3 of the production engine: D over every whole read and D_seed over its first seed_len bases
int bwb_lower_bounds(bwb_ctx *ctx, const uint8_t *seq, const uint64_t *offsets, uint64_t n_reads, int seed_len,
                     int32_t *d_main, int32_t *d_seed) {
    if (!ctx || !seq || !offsets || !d_main || seed_len < 0 || seed_len > 255 || (seed_len && !d_seed)) return BWB_ERR_ARG;
    if (!ctx->have_index) return fail(ctx, BWB_ERR_NO_INDEX, "no index uploaded");
    int max_len, rc;
    if ((rc = check_reads(ctx, offsets, n_reads, max_len))) return rc;
    if (n_reads == 0) return BWB_OK;
    Device &d = ctx->dev[0];
    CU(cudaSetDevice(d.id));
    const bool wide = index_is_wide(ctx);
    const int grid = d.sm_count * 2, n_groups = grid * (256 / GL);
    const uint64_t total = offsets[n_reads] - offsets[0];
    uint8_t *dseq; uint64_t *doff; void *gl; unsigned char *sm; int2 *dm, *ds;
    DevTmp tmp;
    CU(tmp.alloc(&dseq, total + 16)); CU(tmp.alloc(&doff, (n_reads + 1) * 8));
    CU(tmp.alloc(&gl, (size_t)n_groups * 2 * ctx->list_cap * sizeof(ulonglong2)));
    CU(tmp.alloc(&sm, 256));
    CU(tmp.alloc(&dm, (total + n_reads + 1) * sizeof(int2)));
    CU(tmp.alloc(&ds, (n_reads * (size_t)(seed_len + 1) + 1) * sizeof(int2)));
    std::vector<uint64_t> rel(n_reads + 1);
    for (uint64_t r = 0; r <= n_reads; r++) rel[r] = offsets[r] - offsets[0];
    CU(cudaMemcpyAsync(dseq, seq + offsets[0], total, cudaMemcpyHostToDevice, d.stream));
    CU(cudaMemcpyAsync(doff, rel.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice, d.stream));
    CU(cudaMemsetAsync(sm, 0, 256, d.stream));
    CalcArgs c;
    memset(&c, 0, sizeof c);
    c.ix = make_view(ctx, d); c.seq = dseq; c.offsets = doff; c.n_reads = (uint32_t)n_reads;
    c.seed_len = seed_len; c.max_len = max_len; c.is_multiref = 1; c.queue = (uint32_t *)(sm + 4);
    c.glists = gl; c.list_cap = ctx->list_cap; c.d_main = dm; c.d_seed = ds;
    c.status = (uint32_t *)(sm + 8); c.counters = (unsigned long long *)(sm + 32);
    c.smem_per_group = G_LIST_SMEM + ((max_len + 15) & ~15);
    set_ktab(ctx, d, c);
    const size_t smem3 = (size_t)(256 / GL) * c.smem_per_group;
    if (wide) {
        CU(cudaFuncSetAttribute(k_calc_d_g<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
        k_calc_d_g<true><<<grid, 256, smem3, d.stream>>>(c);
    } else {
        CU(cudaFuncSetAttribute(k_calc_d_g<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
        k_calc_d_g<false><<<grid, 256, smem3, d.stream>>>(c);
    }
    CU(cudaGetLastError());
    uint32_t hstatus = 0;
    CU(cudaMemcpyAsync(d_main, dm, (total + n_reads) * sizeof(int2), cudaMemcpyDeviceToHost, d.stream));
    if (seed_len) CU(cudaMemcpyAsync(d_seed, ds, n_reads * (size_t)(seed_len + 1) * sizeof(int2), cudaMemcpyDeviceToHost, d.stream));
    CU(cudaMemcpyAsync(&hstatus, sm + 8, 4, cudaMemcpyDeviceToHost, d.stream));
    CU(cudaStreamSynchronize(d.stream));
    if (hstatus) return fail(ctx, -(int)hstatus, "interval list exceeded list_cap=%d (raise it with bwb_set_option)", ctx->list_cap);
    return BWB_OK;
}

// ---- K4 + K5 ----------------------------------------------------------------------------------
int bwb_reads_upload(bwb_ctx *ctx, const uint8_t *seq, const uint64_t *offsets, uint64_t n_reads, bwb_reads **out) {
    if (!ctx || !seq || !offsets || !out) return BWB_ERR_ARG;
    int max_len, min_len, rc;
    if ((rc = check_reads(ctx, offsets, n_reads, max_len, &min_len))) return rc;
    bwb_reads *R = new bwb_reads();
    R->ctx = ctx; R->n_reads = n_reads; R->max_len = max_len; R->min_len = min_len;
    if (min_len != max_len) {            // mixed lengths: what seed_donors() needs to know about every read
        R->len8.resize(n_reads);
        R->skip_pre.resize(n_reads);
        for (uint64_t r = 0; r < n_reads; r++) {
            const uint64_t o = offsets[r], l = offsets[r + 1] - o;
            R->len8[r] = (uint8_t)l;
            uint8_t bad = l < (uint64_t)PRECALC_LEN;
            for (uint64_t k = 0; k < (uint64_t)PRECALC_LEN && !bad; k++) bad = seq[o + k] > 3;
            R->skip_pre[r] = bad;
        }
    }
    const int G = (int)ctx->dev.size();
    R->shard_lo.resize(G + 1);
    for (int g = 0; g <= G; g++) R->shard_lo[g] = (uint64_t)g * n_reads / G;
    R->d_seq.assign(G, nullptr); R->d_off.assign(G, nullptr); R->shard_bases.assign(G, 0);
    for (int g = 0; g < G; g++) {
        Device &d = ctx->dev[g];
        const uint64_t lo = R->shard_lo[g], hi = R->shard_lo[g + 1], n = hi - lo;
        const uint64_t bytes = offsets[hi] - offsets[lo];
        R->shard_bases[g] = bytes;
        CU(cudaSetDevice(d.id));
        CU(cudaMalloc(&R->d_seq[g], bytes + 16));
        CU(cudaMalloc(&R->d_off[g], (n + 1) * 8));
        std::vector<uint64_t> rel(n + 1);
        for (uint64_t r = 0; r <= n; r++) rel[r] = offsets[lo + r] - offsets[lo];
        CU(cudaMemcpyAsync(R->d_seq[g], seq + offsets[lo], bytes, cudaMemcpyHostToDevice, d.stream));
        CU(cudaMemcpyAsync(R->d_off[g], rel.data(), (n + 1) * 8, cudaMemcpyHostToDevice, d.stream));
        CU(cudaStreamSynchronize(d.stream));
    }
    *out = R;
    return BWB_OK;
}

void bwb_reads_free(bwb_reads *r) {
    if (!r) return;
    for (size_t g = 0; g < r->d_seq.size(); g++) {
        cudaSetDevice(r->ctx->dev[g].id);
        if (r->d_seq[g]) cudaFree(r->d_seq[g]);
        if (r->d_off[g]) cudaFree(r->d_off[g]);
    }
    delete r;
}

// The scores a partial alignment can have: m*M + o*O + e*E for m mismatches, o gap openings and e gap extensions with
// m + o + e <= max_diff (children are only made of entries with differences to spare, inexact_match.c:377-389),
// o <= max_gapo, e <= max_gape (:413-416) and e > 0 only once a gap is open (an extension continues state I/D).  K4 gives
// only those scores a bucket, in ascending order (the order heap_pop looks at them, inexact_match.c:594-610).
// Layout of `map`: uint16 score_of[128], then uint8 bucket_of[nb] (0xff = unreachable).  Returns the bucket count.
static int bucket_map(const bwb_params *p, int nb, std::vector<uint8_t> &map) {
    std::vector<char> reach((size_t)nb, 0);
    for (int o = 0; o <= p->max_gapo && o <= p->max_diff; o++)
        for (int e = 0; e <= (o ? p->max_gape : 0) && o + e <= p->max_diff; e++)
            for (int m = 0; m + o + e <= p->max_diff; m++) {
                const long long sc = (long long)m * p->mm_score + (long long)o * p->gapo_score + (long long)e * p->gape_score;
                if (sc < nb) reach[(size_t)sc] = 1;
            }
    map.assign((size_t)256 + nb, 0xff);
    int nbc = 0;
    for (int sc = 0; sc < nb; sc++) {
        if (!reach[(size_t)sc]) continue;
        if (nbc < 128) {
            map[(size_t)256 + sc] = (uint8_t)nbc;
            const uint16_t v = (uint16_t)sc;
            memcpy(&map[2 * (size_t)nbc], &v, 2);
        }
        nbc++;
    }
    return nbc;
}

extern "C" int bwb_score_buckets(const bwb_params *p, uint8_t *bucket_of, int cap) {
    if (!p || p->max_diff < 0 || p->max_gapo < 0 || p->max_gape < 0 || p->mm_score < 0 || p->gapo_score < 0 || p->gape_score < 0 || p->max_diff > 200)
        return fail(nullptr, BWB_ERR_ARG, "negative or out-of-range alignment parameter");
    const int nb = (p->max_diff + 1) * p->mm_score + (p->max_gapo + 1) * p->gapo_score + (p->max_gape + 1) * p->gape_score;
    if (nb <= 0 || nb > 1024) return fail(nullptr, BWB_ERR_UNSUPPORTED, "%d score buckets (supported: 1..1024)", nb);
    std::vector<uint8_t> map;
    const int nbc = bucket_map(p, nb, map);
    for (int s = 0; bucket_of && s < cap && s < nb; s++) bucket_of[s] = map[(size_t)256 + s];
    return nbc;
}

static int check_params(bwb_ctx *ctx, const bwb_params *p, int max_len, int &nb) {
    if (p->use_precalc) {
        if (ctx->engine != 0) return fail(ctx, BWB_ERR_UNSUPPORTED, "-P (pre-calculated intervals) is only built in the lane engine");
        if (!ctx->have_pre) return fail(ctx, BWB_ERR_ARG, "use_precalc without a seed table (bwb_precalc_build / _load_file / _upload)");
        if (ctx->pre_multiref != (p->is_multiref ? 1 : 0))
            return fail(ctx, BWB_ERR_ARG, "the seed table was made for %s mode", ctx->pre_multiref ? "multi-genome" : "-S");
    }
    if (!p->is_multiref && ctx->engine != 0)
        return fail(ctx, BWB_ERR_UNSUPPORTED, "-S (single-genome mode) is only built in the lane engine");
    if (p->max_diff < 0 || p->max_gapo < 0 || p->max_gape < 0 || p->mm_score < 0 || p->gapo_score < 0 || p->gape_score < 0 ||
        p->seed_length < 0 || p->max_diff > 200)
        return fail(ctx, BWB_ERR_ARG, "negative or out-of-range alignment parameter");
    if (p->max_gapo > BWB_MAX_GAP_RUNS) return fail(ctx, BWB_ERR_UNSUPPORTED, "max_gapo > %d", BWB_MAX_GAP_RUNS);
    nb = (p->max_diff + 1) * p->mm_score + (p->max_gapo + 1) * p->gapo_score + (p->max_gape + 1) * p->gape_score;
    if (nb <= 0 || nb > 1024) return fail(ctx, BWB_ERR_UNSUPPORTED, "%d score buckets (supported: 1..1024)", nb);
    // the engine limits, all in one place and before anything is launched (INTEGRATION.md "Limits of the device path")
    if (ctx->engine == 0) {
        ctx->nbc = bucket_map(p, nb, ctx->bmap);
        if (ctx->nbc > 128)
            return fail(ctx, BWB_ERR_UNSUPPORTED, "%d distinct alignment scores m*M + o*O + e*E are reachable with -n %d -o %d -e %d -M %d -O %d -E %d; "
                        "the device search keeps the occupancy of at most 128 score buckets in registers (the reference allocates "
                        "(n+1)*M + (o+1)*O + (e+1)*E = %d of them, inexact_match.c:510-528)",
                        ctx->nbc, p->max_diff, p->max_gapo, p->max_gape, p->mm_score, p->gapo_score, p->gape_score, nb);
    }
    if (p->seed_length > 255) return fail(ctx, BWB_ERR_ARG, "seed_length > 255");
    int gaps = p->max_gapo + p->max_gape;
    if (gaps > p->max_diff) gaps = p->max_diff;
    if (max_len + gaps > 255) return fail(ctx, BWB_ERR_UNSUPPORTED, "read length %d + %d gap steps wraps the 8-bit path length (align.h:104)", max_len, gaps);
    return BWB_OK;
}

constexpr uint64_t READ_BATCH_REF = 0x40000;      // READ_BATCH_SIZE, align.h:14

// first `want` bases of read r of a resident batch -> host (Q6: donor reads that live on another device, the carry)
static int fetch_read_prefix(bwb_ctx *ctx, const bwb_reads *R, uint64_t r, int want, uint8_t *out, int *got) {
    const int G = (int)ctx->dev.size();
    int g = 0;
    while (g + 1 < G && r >= R->shard_lo[g + 1]) g++;
    uint64_t off = 0;
    int len = R->max_len;
    if (!R->len8.empty()) {
        for (uint64_t q = R->shard_lo[g]; q < r; q++) off += R->len8[q];
        len = R->len8[r];
    } else {
        off = (r - R->shard_lo[g]) * (uint64_t)R->max_len;
    }
    const int n = len < want ? len : want;
    Device &d = ctx->dev[g];
    CU(cudaSetDevice(d.id));
    CU(cudaMemcpyAsync(out, (const uint8_t *)R->d_seq[g] + off, (size_t)n, cudaMemcpyDeviceToHost, d.stream));
    CU(cudaStreamSynchronize(d.stream));
    *got = n;
    return BWB_OK;
}

// SURVEY Q6.  The reference computes D_seed only for reads longer than the seed (inexact_match.c:62-64,141-143); a
// shorter read consults whatever the array holds: the bounds of the last longer read its driver thread aligned before
// it, or the calloc'ed zeros.  "Before it" = in the whole run for the serial driver (one array per
// align_reads_inexact call, :36), inside the thread's static chunk of the 262144-read batch for the OpenMP driver
// (:115-121); with -P a read skipped for an N in its 12-mer leaves the array alone (:50-57).  One bwb_align call is one
// driver call: params->n_threads picks the driver, and with option "seed_carry" the serial chain continues across calls
// (streaming entry points).  The donor of every short read goes to K3 as an offset into the shard's bases
// (CalcArgs::seed_src); the one donor per shard that can lie outside it travels as its first seed_length bases.
// donor[r] for every read of one driver call: r itself (the read is longer than the seed), DONOR_NONE (zeros),
// DONOR_CARRY (the run's carry) or the index of an earlier read of the call.  len8 / skip may be null (all reads
// `uniform_len` long / no read skipped by -P).  Pure host arithmetic: bwb_seed_donor_plan exposes it to the CPU tests.
constexpr long long DONOR_NONE = -1, DONOR_CARRY = -2;
static void plan_seed_donors(const uint8_t *len8, int uniform_len, const uint8_t *skip, uint64_t n, int sl, int n_threads,
                             bool use_precalc, bool carry, long long *donor_of) {
    const bool serial = n_threads <= 1;
    const uint64_t nt = serial ? 1 : (uint64_t)n_threads;
    long long donor = (serial && carry) ? DONOR_CARRY : DONOR_NONE;
    for (uint64_t r = 0; r < n; r++) {
        if (!serial) {                            // a new thread chunk starts with a fresh (zero) array
            const uint64_t w0 = r / READ_BATCH_REF * READ_BATCH_REF, bs = std::min<uint64_t>(READ_BATCH_REF, n - w0), q = r - w0;
            const uint64_t t = ((q + 1) * nt - 1) / bs;                        // the chunk q falls in: [t*bs/nt, (t+1)*bs/nt)
            if (t * bs / nt == q) donor = DONOR_NONE;
        }
        const int len = len8 ? (int)len8[r] : uniform_len;
        const bool skipped = use_precalc && (skip ? skip[r] != 0 : len < PRECALC_LEN);
        if (skipped) { donor_of[r] = DONOR_NONE; continue; }                   // never searched: no array consulted
        if (len > sl) donor = (long long)r;
        donor_of[r] = donor;
    }
}

static int seed_donors(bwb_ctx *ctx, const bwb_params *p, const bwb_reads *R) {
    const int G = (int)ctx->dev.size();
    for (int g = 0; g < G; g++) ctx->dev[g].use_seed_src = false;
    const int sl = p->seed_length;
    if (sl <= 0 || R->n_reads == 0 || R->min_len > sl) return BWB_OK;         // every read has its own D_seed
    const bool carry = p->n_threads <= 1 && ctx->seed_carry && ctx->carry_len > sl;
    if (R->len8.empty() && !carry) return BWB_OK;                              // all reads short and equally long: zeros
    std::vector<long long> donor_of(R->n_reads);
    plan_seed_donors(R->len8.empty() ? nullptr : R->len8.data(), R->max_len, R->skip_pre.empty() ? nullptr : R->skip_pre.data(),
                     R->n_reads, sl, p->n_threads, p->use_precalc != 0, carry, donor_of.data());
    int rc;
    for (int g = 0; g < G; g++) {
        Device &d = ctx->dev[g];
        const uint64_t lo = R->shard_lo[g], hi = R->shard_lo[g + 1];
        std::vector<uint32_t> src(hi - lo, SEED_SRC_NONE), first_base(hi - lo);
        long long ext = DONOR_NONE;               // the one donor outside this shard, if it is used
        bool any = false;
        uint64_t off = 0;
        for (uint64_t r = lo; r < hi; r++) {
            first_base[r - lo] = (uint32_t)off;
            const int len = R->len8.empty() ? R->max_len : (int)R->len8[r];
            const long long dn = donor_of[r];
            if (len <= sl && dn != DONOR_NONE) {
                any = true;
                if (dn >= (long long)lo) src[r - lo] = first_base[(uint64_t)dn - lo];
                else {
                    if (ext != DONOR_NONE && ext != dn) return fail(ctx, BWB_ERR_ARG, "internal: two outside D_seed donors for one shard");
                    src[r - lo] = SEED_SRC_EXT;
                    ext = dn;
                }
            }
            off += (uint64_t)len;
        }
        if (!any) continue;
        CU(cudaSetDevice(d.id));
        if ((rc = ensure(ctx, d.seed_src, src.size() * 4))) return rc;
        if ((rc = ensure(ctx, d.seed_ext, 256))) return rc;
        uint8_t ext_seq[256] = {0};
        if (ext == DONOR_CARRY) memcpy(ext_seq, ctx->carry_seq, 256);
        else if (ext >= 0) {
            int got = 0;
            if ((rc = fetch_read_prefix(ctx, R, (uint64_t)ext, sl, ext_seq, &got))) return rc;
            CU(cudaSetDevice(d.id));
        }
        CU(cudaMemcpyAsync(d.seed_src.p, src.data(), src.size() * 4, cudaMemcpyHostToDevice, d.stream));
        CU(cudaMemcpyAsync(d.seed_ext.p, ext_seq, 256, cudaMemcpyHostToDevice, d.stream));
        CU(cudaStreamSynchronize(d.stream));           // src / ext_seq are locals
        d.use_seed_src = true;
    }
    return BWB_OK;
}

extern "C" int bwb_seed_donor_plan(const bwb_params *p, const uint8_t *seq, const uint64_t *offsets, uint64_t n_reads,
                                   int have_carry, int64_t *donor_of) {
    if (!p || !seq || !offsets || !donor_of) return fail(nullptr, BWB_ERR_ARG, "bwb_seed_donor_plan: null argument");
    std::vector<uint8_t> len8(n_reads), skip(n_reads);
    for (uint64_t r = 0; r < n_reads; r++) {
        const uint64_t o = offsets[r], l = offsets[r + 1] - o;
        if (offsets[r + 1] < o || l > 255) return fail(nullptr, BWB_ERR_ARG, "read %llu: bad offsets or longer than 255 bases", (unsigned long long)r);
        len8[r] = (uint8_t)l;
        uint8_t bad = l < (uint64_t)PRECALC_LEN;
        for (uint64_t k = 0; k < (uint64_t)PRECALC_LEN && !bad; k++) bad = seq[o + k] > 3;
        skip[r] = bad;
    }
    std::vector<long long> d(n_reads);
    plan_seed_donors(len8.data(), 0, skip.data(), n_reads, p->seed_length, p->n_threads, p->use_precalc != 0, have_carry != 0, d.data());
    for (uint64_t r = 0; r < n_reads; r++) donor_of[r] = (int64_t)d[r];
    return BWB_OK;
}

// after a call of a run with "seed_carry": remember the last read longer than the seed (serial driver only)
static int seed_carry_update(bwb_ctx *ctx, const bwb_params *p, const bwb_reads *R) {
    if (!ctx->seed_carry || p->n_threads > 1 || p->seed_length <= 0 || R->n_reads == 0 || R->max_len <= p->seed_length) return BWB_OK;
    for (uint64_t r = R->n_reads; r-- > 0;) {
        const int len = R->len8.empty() ? R->max_len : (int)R->len8[r];
        if (len <= p->seed_length) continue;
        if (p->use_precalc && !R->skip_pre.empty() && R->skip_pre[r]) continue;      // -P skipped it: D_seed untouched
        uint8_t tmp[256] = {0};
        int got = 0, rc;
        if ((rc = fetch_read_prefix(ctx, R, r, 255, tmp, &got))) return rc;
        if (p->use_precalc && R->skip_pre.empty()) {       // equally long reads: the skip condition was not recorded at upload
            bool bad = got < PRECALC_LEN;
            for (int k = 0; k < PRECALC_LEN && !bad; k++) bad = tmp[k] > 3;
            if (bad) continue;
        }
        memcpy(ctx->carry_seq, tmp, sizeof tmp);
        ctx->carry_len = len;
        return BWB_OK;
    }
    return BWB_OK;
}

// enqueue K4 + scan + K5 for one shard on its device stream
static int launch_shard(bwb_ctx *ctx, Device &d, const bwb_params *p, int nb, const SmemLayout &L, int max_len, bool wide,
                        const void *d_seq, const void *d_off, uint64_t n, uint64_t total_bases, uint64_t read_base, unsigned long long out_cap) {
    int rc;
    CU(cudaSetDevice(d.id));
    if ((rc = ensure(ctx, d.read_off, n * 8 + 8))) return rc;
    if ((rc = ensure(ctx, d.read_cnt, n * 4 + 4))) return rc;
    if ((rc = ensure(ctx, d.ordered_off, (n + 1) * 8))) return rc;
    if ((rc = ensure(ctx, d.unordered, out_cap * sizeof(bwb_hit)))) return rc;
    if ((rc = ensure(ctx, d.ordered, out_cap * sizeof(bwb_hit)))) return rc;
    if ((rc = ensure(ctx, d.small, 512))) return rc;
    // small block: [0] K4 queue u32 | [4] K3 queue u32 | [8] status 2xu32 | [16] out_cursor u64 |
    // [24] shared-pool bump cursor u32 | [32..96) counters 8xu64 | [96] shared-pool free-list head u64 |
    // [112],[116] queues of K4's retry passes | [120],[124] their lengths (deferred-read counts) | [256..512) K3b
    unsigned char *sm = (unsigned char *)d.small.p;
    CU(cudaMemsetAsync(sm, 0, 512, d.stream));
    const uint32_t ovf0 = (uint32_t)d.n_warps * d.chunks_per_warp;
    CU(cudaMemcpyAsync(sm + 24, &ovf0, 4, cudaMemcpyHostToDevice, d.stream));
    if (ctx->engine == 0) {        // lane engine: shared blocks of LBLK slots above the private ranges
        if ((rc = ensure(ctx, d.pool, sizeof(PoolState)))) return rc;
        static thread_local PoolState hp;
        const uint32_t blk0 = d.priv_total / LBLK, blk1 = d.total_slots / LBLK;
        const uint32_t per = (blk1 > blk0 ? blk1 - blk0 : 0u) / POOL_SHARDS;
        for (int s = 0; s < POOL_SHARDS; s++) {
            hp.head[s] = 0xffffffffull;
            hp.bump[s] = blk0 + (uint32_t)s * per;
            hp.limit[s] = blk0 + (uint32_t)(s + 1) * per;
        }
        hp.n_borrowed = 0;
        CU(cudaMemcpyAsync(d.pool.p, &hp, sizeof hp, cudaMemcpyHostToDevice, d.stream));
    }
    if (ctx->engine == 2) {        // shared chunk pool: POOL_SHARDS regions above the private ranges
        if ((rc = ensure(ctx, d.pool, sizeof(PoolState)))) return rc;
        static thread_local PoolState hp;
        const uint32_t shared = d.n_chunks > ovf0 ? d.n_chunks - ovf0 : 0u, per = shared / POOL_SHARDS;
        for (int s = 0; s < POOL_SHARDS; s++) {
            hp.head[s] = 0xffffffffull;                                      // tag 0, empty
            hp.bump[s] = ovf0 + (uint32_t)s * per;
            hp.limit[s] = ovf0 + (uint32_t)(s + 1) * per;
        }
        hp.n_borrowed = 0;
        CU(cudaMemcpyAsync(d.pool.p, &hp, sizeof hp, cudaMemcpyHostToDevice, d.stream));
    }

    AlignArgs a;
    memset(&a, 0, sizeof a);
    a.ix = make_view(ctx, d);
    a.seq = (const uint8_t *)d_seq; a.offsets = (const uint64_t *)d_off;
    a.n_reads = (uint32_t)n; a.read_id_base = (uint32_t)read_base;
    a.max_diff = p->max_diff; a.max_gapo = p->max_gapo; a.max_gape = p->max_gape; a.max_entries = p->max_entries;
    a.mm_score = p->mm_score; a.gapo_score = p->gapo_score; a.gape_score = p->gape_score;
    a.seed_len = p->seed_length; a.max_diff_seed = p->max_diff_seed; a.max_best = p->max_best;
    a.no_indel_len = p->no_indel_length;
    a.nb = nb; a.max_len = max_len;
    a.queue = (uint32_t *)sm;
    a.glists = d.glists.p; a.list_cap = ctx->list_cap;
    a.chunks = (uint4 *)d.chunks.p; a.chunk_link = (uint32_t *)d.chunk_link.p;
    a.chunks_per_warp = d.chunks_per_warp; a.n_chunks = d.n_chunks;
    a.overflow_cursor = (uint32_t *)(sm + 24);
    a.stage = (bwb_hit *)d.stage.p; a.hits_cap = ctx->hits_per_read;
    a.out_hits = (bwb_hit *)d.unordered.p; a.out_cap = out_cap;
    a.out_cursor = (unsigned long long *)(sm + 16);
    a.read_off = (unsigned long long *)d.read_off.p; a.read_cnt = (uint32_t *)d.read_cnt.p;
    a.status = (uint32_t *)(sm + 8);
    a.counters = (unsigned long long *)(sm + 32);
    a.smem_per_warp = L.per_warp; a.off_D = L.off_D; a.off_Ds = L.off_Ds; a.off_bk = L.off_bk; a.off_seq = L.off_seq;

    CU(cudaEventRecord(d.ev0, d.stream));
    CU(cudaEventRecord(d.evm, d.stream));
#ifdef BWB_AB_ENGINES
    if (n && ctx->engine == 1) {
        if (wide) k_align<true><<<d.grid, d.wpb * 32, d.smem_bytes, d.stream>>>(a);
        else k_align<false><<<d.grid, d.wpb * 32, d.smem_bytes, d.stream>>>(a);
        CU(cudaGetLastError());
    } else
#endif
    if (n && ctx->engine == 0) {
        // K3 (8-lane groups): packed lower-bound arrays of every read -> HBM
        if ((rc = ensure(ctx, d.pk_main, (total_bases + n + 1) * 2))) return rc;
        if ((rc = ensure(ctx, d.pk_seed, (n * (size_t)(p->seed_length + 1) + 1) * 2))) return rc;
        if ((rc = ensure(ctx, d.n_count, (n + 1) * 2))) return rc;
        CalcArgs c;
        memset(&c, 0, sizeof c);
        c.ix = a.ix; c.seq = a.seq; c.offsets = a.offsets; c.n_reads = a.n_reads;
        c.seed_len = p->seed_length; c.max_len = max_len; c.is_multiref = p->is_multiref;
        c.queue = (uint32_t *)(sm + 4);
        c.glists = d.glists.p; c.list_cap = ctx->list_cap;
        c.pk_main = (uint16_t *)d.pk_main.p; c.pk_seed = (uint16_t *)d.pk_seed.p; c.n_count = (uint16_t *)d.n_count.p;
        c.status = a.status; c.counters = a.counters;
        c.smem_per_group = G_LIST_SMEM + ((std::max(max_len, p->seed_length) + 15) & ~15);     // (a donor's seed is staged there too)
        if (d.use_seed_src) { c.seed_src = (const uint32_t *)d.seed_src.p; c.seed_ext = (const uint8_t *)d.seed_ext.p; }
        set_ktab(ctx, d, c);
        const size_t smem3 = (size_t)(256 / GL) * c.smem_per_group;
        const int grid3 = d.grid3;
        // K3's coordinate width follows the index alone (the k-mer table is typed by it); `wide` may also
        // be set by max_gapo > 1, which only concerns K4's entry format
        if (index_is_wide(ctx)) {
            CU(cudaFuncSetAttribute(k_calc_d_g<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
            k_calc_d_g<true><<<grid3, 256, smem3, d.stream>>>(c);
        } else {
            CU(cudaFuncSetAttribute(k_calc_d_g<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
            k_calc_d_g<false><<<grid3, 256, smem3, d.stream>>>(c);
        }
        CU(cudaGetLastError());
        // K3b: heavy reads first (counting sort on K3's whole-read bound); part of the K3 interval of the timing
        uint32_t *order = nullptr;
        if (ctx->heavy_first) {
            if ((rc = ensure(ctx, d.order, (n + 1) * 4))) return rc;
            order = (uint32_t *)d.order.p;
            uint32_t *hist = (uint32_t *)(sm + 256), *cursor = hist + ORDER_CLASSES;
            const int ogrid = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)d.sm_count * 8);
            k_order_hist<<<ogrid, 256, 0, d.stream>>>(c.pk_main, c.n_count, a.offsets, a.n_reads, p->max_diff, hist);
            CU(cudaGetLastError());
            k_order_scatter<<<ogrid, 256, 0, d.stream>>>(c.pk_main, c.n_count, a.offsets, a.n_reads, p->max_diff, hist, cursor, order);
            CU(cudaGetLastError());
        }
        CU(cudaEventRecord(d.evm, d.stream));
        // K4: one read per lane
        LaneArgs g;
        memset(&g, 0, sizeof g);
        g.order = order;
        g.ix = a.ix; g.seq = a.seq; g.offsets = a.offsets; g.n_reads = a.n_reads; g.read_id_base = a.read_id_base;
        g.max_diff = a.max_diff; g.max_gapo = a.max_gapo; g.max_gape = a.max_gape; g.max_entries = a.max_entries;
        g.mm_score = a.mm_score; g.gapo_score = a.gapo_score; g.gape_score = a.gape_score;
        g.seed_len = a.seed_len; g.max_diff_seed = a.max_diff_seed; g.max_best = a.max_best; g.no_indel_len = a.no_indel_len;
        g.nb = a.nb; g.queue = a.queue; g.is_multiref = p->is_multiref;
        if ((rc = ensure(ctx, d.bmap, ctx->bmap.size()))) return rc;
        CU(cudaMemcpyAsync(d.bmap.p, ctx->bmap.data(), ctx->bmap.size(), cudaMemcpyHostToDevice, d.stream));
        g.nbc = ctx->nbc; g.bmap = (const uint8_t *)d.bmap.p;
        g.pk_main = c.pk_main; g.pk_seed = c.pk_seed; g.n_count = c.n_count;
        g.slots = (uint4 *)d.chunks.p;
        g.slots_per_lane = d.slots_per_lane; g.priv_total = d.priv_total;
        g.pool = (PoolState *)d.pool.p; g.blk_link = (uint32_t *)d.blk_link.p;
        // popped slots are recycled where heaps get large (see k_search_l): big indexes, long reads, the wide entry format
        const bool recycle = ctx->recycle == 1 || (ctx->recycle != 2 && (wide || ctx->length > (1ull << 28) || max_len > 128));
        {   // admission control: no new reads while more than throttle_pct % of the shared blocks are lent out
            const uint64_t shared_blocks = d.total_slots / LBLK > d.priv_total / LBLK ? d.total_slots / LBLK - d.priv_total / LBLK : 0;
            // (with the heavy-heap configurations only, i.e. together with slot recycling, unless the option forces it)
            const bool on = ctx->throttle_pct > 0 && ctx->throttle_pct < 100 && (recycle || ctx->throttle_forced);
            g.throttle_blocks = on ? shared_blocks * (uint64_t)ctx->throttle_pct / 100 : 0;
        }
        g.out_hits = a.out_hits; g.out_cap = a.out_cap; g.out_cursor = a.out_cursor;
        g.read_off = a.read_off; g.read_cnt = a.read_cnt; g.status = a.status; g.counters = a.counters;
        g.pre_off = d.pre_off; g.pre_cnt = d.pre_cnt; g.pre_iv = d.pre_iv;
        // three passes: all reads; the reads pass 0 deferred for lack of arena; what pass 1 deferred (an
        // overflow there is reported).  Passes 1 and 2 read their queue length on the device: no host sync,
        // and an empty pass costs a few microseconds.
        if ((rc = ensure(ctx, d.retry, 2 * (n + 1) * 4))) return rc;
        uint32_t *rl0 = (uint32_t *)d.retry.p, *rl1 = rl0 + (n + 1);
        for (int pass = 0; pass < 3; pass++) {
            g.lane_stride = pass == 0 ? 1u : (pass == 1 ? 8u : 64u);      // 8x, then 64x the arena per read in flight
            if (pass == 0) { g.retry_list = rl0; g.retry_count = (uint32_t *)(sm + 120); }
            else if (pass == 1) {
                g.order = rl0; g.n_queue_ptr = (const uint32_t *)(sm + 120); g.queue = (uint32_t *)(sm + 112);
                g.retry_list = rl1; g.retry_count = (uint32_t *)(sm + 124);
            } else {
                g.order = rl1; g.n_queue_ptr = (const uint32_t *)(sm + 124); g.queue = (uint32_t *)(sm + 116);
                g.retry_list = nullptr; g.retry_count = nullptr;
            }
            k4_fn(wide, p->use_precalc != 0, recycle)<<<d.grid, 128, d.smem_bytes, d.stream>>>(g);
            CU(cudaGetLastError());
        }
        CU(cudaGetLastError());
    }
#ifdef BWB_AB_ENGINES
    else if (n) {
        // K3: lower-bound arrays of every read -> HBM
        if ((rc = ensure(ctx, d.d_main, (total_bases + n + 1) * sizeof(int2)))) return rc;
        if ((rc = ensure(ctx, d.d_seed, (n * (size_t)(p->seed_length + 1) + 1) * sizeof(int2)))) return rc;
        CalcArgs c;
        memset(&c, 0, sizeof c);
        c.ix = a.ix; c.seq = a.seq; c.offsets = a.offsets; c.n_reads = a.n_reads;
        c.seed_len = p->seed_length; c.max_len = max_len; c.is_multiref = 1;
        c.queue = (uint32_t *)(sm + 4);
        c.glists = d.glists.p; c.list_cap = ctx->list_cap;
        c.d_main = (int2 *)d.d_main.p; c.d_seed = (int2 *)d.d_seed.p;
        c.status = a.status; c.counters = a.counters;
        c.smem_per_group = G_LIST_SMEM + ((max_len + 15) & ~15);
        set_ktab(ctx, d, c);
        const size_t smem3 = (size_t)(256 / GL) * c.smem_per_group;
        if (index_is_wide(ctx)) {
            CU(cudaFuncSetAttribute(k_calc_d_g<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
            k_calc_d_g<true><<<d.grid, 256, smem3, d.stream>>>(c);
        } else {
            CU(cudaFuncSetAttribute(k_calc_d_g<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
            k_calc_d_g<false><<<d.grid, 256, smem3, d.stream>>>(c);
        }
        CU(cudaGetLastError());
        // K4: the search proper
        SearchArgs g;
        memset(&g, 0, sizeof g);
        g.ix = a.ix; g.seq = a.seq; g.offsets = a.offsets; g.n_reads = a.n_reads; g.read_id_base = a.read_id_base;
        g.max_diff = a.max_diff; g.max_gapo = a.max_gapo; g.max_gape = a.max_gape; g.max_entries = a.max_entries;
        g.mm_score = a.mm_score; g.gapo_score = a.gapo_score; g.gape_score = a.gape_score;
        g.seed_len = a.seed_len; g.max_diff_seed = a.max_diff_seed; g.max_best = a.max_best; g.no_indel_len = a.no_indel_len;
        g.nb = a.nb; g.max_len = a.max_len; g.queue = a.queue;
        g.d_main = c.d_main; g.d_seed = c.d_seed;
        g.glists = a.glists; g.list_cap = a.list_cap;
        g.chunks = a.chunks; g.chunk_link = a.chunk_link; g.chunks_per_group = a.chunks_per_warp; g.n_chunks = a.n_chunks;
        g.priv_total = ovf0; g.pool = (PoolState *)d.pool.p;
        g.stage = a.stage; g.hits_cap = a.hits_cap;
        g.out_hits = a.out_hits; g.out_cap = a.out_cap; g.out_cursor = a.out_cursor;
        g.read_off = a.read_off; g.read_cnt = a.read_cnt; g.status = a.status; g.counters = a.counters;
        g.smem_per_group = L.per_warp; g.off_D = L.off_D; g.off_Ds = L.off_Ds; g.off_bk = L.off_bk; g.off_seq = L.off_seq;
        if (wide) k_search_g<true><<<d.grid, 256, d.smem_bytes, d.stream>>>(g);
        else k_search_g<false><<<d.grid, 256, d.smem_bytes, d.stream>>>(g);
        CU(cudaGetLastError());
    }
#endif  // BWB_AB_ENGINES
    CU(cudaEventRecord(d.ev1, d.stream));
    // K5: exclusive scan of the per-read counts, then ordered copy
    const uint32_t *cnt = (const uint32_t *)d.read_cnt.p;
    unsigned long long *ooff = (unsigned long long *)d.ordered_off.p;
    if (n) {
        k_scan_counts<<<1, 1024, 0, d.stream>>>(cnt, (uint32_t)n, ooff);
        CU(cudaGetLastError());
        k_emit<<<(unsigned)((n + 127) / 128), 128, 0, d.stream>>>((const bwb_hit *)d.unordered.p, a.read_off, cnt, ooff,
                                                              (uint32_t)n, (bwb_hit *)d.ordered.p, a.out_cursor, out_cap);
        CU(cudaGetLastError());
        if (ctx->have_sa) {                          // K6: locate + top1/top2 (aln2sam's eval_aln)
            if ((rc = ensure(ctx, d.loc, n * sizeof(bwb_loc)))) return rc;
            k_locate<<<(unsigned)((n + 127) / 128), 128, 0, d.stream>>>(a.ix, ctx->sa0, d.sa, (const bwb_hit *)d.ordered.p, ooff,
                                                                    cnt, (uint32_t)n, (bwb_loc *)d.loc.p, a.out_cursor, out_cap);
            CU(cudaGetLastError());
        }
    }
    return BWB_OK;
}

static int align_impl(bwb_ctx *ctx, const bwb_params *p, const bwb_reads *R, bwb_results **out) {
    if (!ctx->have_index) return fail(ctx, BWB_ERR_NO_INDEX, "bwb_align before bwb_index_upload");
    int nb = 0, rc;
    if ((rc = check_params(ctx, p, R->max_len, nb))) return rc;
    const SmemLayout L = ctx->engine == 1 ? k4_layout(R->max_len > 0 ? R->max_len : 1, p->seed_length, nb)
                                          : g4_layout(R->max_len > 0 ? R->max_len : 1, p->seed_length, nb);   // engine 0: unused
    // 16-byte entries + 32-bit coordinates unless the index or the gap-run count needs the wide format
    const bool wide = index_is_wide(ctx) || p->max_gapo > 1;
    const int G = (int)ctx->dev.size();
    bwb_results *res = new bwb_results();
    res->ctx = ctx; res->n_reads = R->n_reads; res->shard_lo = R->shard_lo; res->shard_total.assign(G, 0);
    res->counts.assign(R->n_reads, 0);
    res->have_loc = ctx->have_sa;
    res->generation = ++ctx->launch_generation;

    if ((rc = seed_donors(ctx, p, R))) { delete res; return rc; }
    std::vector<unsigned long long> cap(G);
    for (int g = 0; g < G; g++) {
#ifdef BWB_AB_ENGINES
        rc = ctx->engine == 1 ? prepare_search(ctx, ctx->dev[g], L, wide)
           : ctx->engine == 2 ? prepare_search_group(ctx, ctx->dev[g], L, wide)
                              : prepare_search_lane(ctx, ctx->dev[g], nb, wide);
#else
        rc = prepare_search_lane(ctx, ctx->dev[g], nb, wide);
#endif
        if (rc) { delete res; return rc; }
        cap[g] = ctx->hit_cap0 >= 0 ? (unsigned long long)ctx->hit_cap0 : (R->shard_lo[g + 1] - R->shard_lo[g]) * 2 + 65536;
    }
    std::vector<char> done(G, 0);
    for (int attempt = 0; attempt < 6; attempt++) {
        for (int g = 0; g < G; g++) {
            if (done[g]) continue;
            const uint64_t lo = R->shard_lo[g], n = R->shard_lo[g + 1] - lo;
            if ((rc = launch_shard(ctx, ctx->dev[g], p, nb, L, R->max_len > 0 ? R->max_len : 1, wide, R->d_seq[g], R->d_off[g], n, R->shard_bases[g], lo, cap[g]))) {
                delete res;
                return rc;
            }
        }
        bool again = false;
        for (int g = 0; g < G; g++) {
            if (done[g]) continue;
            Device &d = ctx->dev[g];
            CU(cudaSetDevice(d.id));
            CU(cudaMemcpyAsync(d.h_small, d.small.p, 256, cudaMemcpyDeviceToHost, d.stream));
            CU(cudaStreamSynchronize(d.stream));
            const unsigned char *hs = (const unsigned char *)d.h_small;
            uint32_t st[2];
            memcpy(st, hs + 8, 8);
            unsigned long long used;
            memcpy(&used, hs + 16, 8);
            if (st[0]) {
                delete res;
                if ((int)st[0] == -BWB_ERR_UNSUPPORTED)
                    return fail(ctx, BWB_ERR_UNSUPPORTED, "internal: read %u produced an alignment score that has no bucket (bucket_map)", st[1]);
                return fail(ctx, -(int)st[0], "device pool overflow while aligning read %u (heap_pool_mb=%lld [0=auto] list_cap=%d hits_per_read=%d)",
                            st[1], ctx->heap_pool_mb, ctx->list_cap, ctx->hits_per_read);
            }
            if (used > cap[g]) { cap[g] = used + used / 8 + 1024; again = true; continue; }
            done[g] = 1;
            res->shard_total[g] = used;
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, d.evm, d.ev1));
            if (ms > res->kernel_ms) res->kernel_ms = ms;
            CU(cudaEventElapsedTime(&ms, d.ev0, d.evm));
            if (ms > res->k3_ms) res->k3_ms = ms;
            const unsigned long long *ctr = (const unsigned long long *)(hs + 32);
            for (int k = 0; k < 4; k++) res->counters[k] += ctr[k];

            for (int k = 4; k < 6; k++) if (ctr[k] > res->counters[k]) res->counters[k] = ctr[k];
            {   // reads K4 deferred for lack of arena: queue lengths of its retry passes (8x / 64x the arena per read)
                uint32_t dq[2];
                memcpy(dq, hs + 120, 8);
                res->counters[6] += dq[0];
                res->counters[7] += dq[1];
            }
        }
        if (!again) break;
    }
    for (int g = 0; g < G; g++)
        if (!done[g]) { delete res; return fail(ctx, BWB_ERR_CAPACITY, "hit output did not fit after retries"); }
    if ((rc = seed_carry_update(ctx, p, R))) { delete res; return rc; }
    *out = res;
    return BWB_OK;
}

int bwb_set_seed_carry(bwb_ctx *ctx, const uint8_t *read_seq, int len) {
    if (!ctx || len < 0 || len > 255 || (len && !read_seq)) return fail(ctx, BWB_ERR_ARG, "bwb_set_seed_carry: a read of 0..255 bases");
    ctx->carry_len = len;
    memset(ctx->carry_seq, 0, sizeof ctx->carry_seq);
    if (len) memcpy(ctx->carry_seq, read_seq, (size_t)len);
    ctx->seed_carry = 1;
    return BWB_OK;
}

int bwb_results_fetch(bwb_results *r) {
    if (!r) return BWB_ERR_ARG;
    if (r->fetched) return BWB_OK;
    bwb_ctx *ctx = r->ctx;
    if (r->generation != ctx->launch_generation)
        return fail(ctx, BWB_ERR_ARG, "bwb_results_fetch: a later bwb_align/bwb_align_resident on this context has overwritten "
                                      "the device buffers of these results (fetch before the next launch)");
    const int G = (int)ctx->dev.size();
    uint64_t total = 0;
    for (int g = 0; g < G; g++) total += r->shard_total[g];
    r->hits.resize(total);
    if (r->have_loc) r->loc.resize(r->n_reads);
    uint64_t w = 0;
    for (int g = 0; g < G; g++) {
        Device &d = ctx->dev[g];
        const uint64_t lo = r->shard_lo[g], n = r->shard_lo[g + 1] - lo;
        CU(cudaSetDevice(d.id));
        if (n) CU(cudaMemcpyAsync(r->counts.data() + lo, d.read_cnt.p, n * 4, cudaMemcpyDeviceToHost, d.stream));
        if (r->shard_total[g])
            CU(cudaMemcpyAsync(r->hits.data() + w, d.ordered.p, r->shard_total[g] * sizeof(bwb_hit), cudaMemcpyDeviceToHost, d.stream));
        if (r->have_loc && n) CU(cudaMemcpyAsync(r->loc.data() + lo, d.loc.p, n * sizeof(bwb_loc), cudaMemcpyDeviceToHost, d.stream));
        w += r->shard_total[g];
    }
    for (int g = 0; g < G; g++) {
        CU(cudaSetDevice(ctx->dev[g].id));
        CU(cudaStreamSynchronize(ctx->dev[g].stream));
    }
    r->fetched = true;
    return BWB_OK;
}

int bwb_align_resident(bwb_ctx *ctx, const bwb_params *params, const bwb_reads *reads, int fetch, bwb_results **out) {
    if (!ctx || !params || !reads || !out || reads->ctx != ctx) return BWB_ERR_ARG;
    int rc = align_impl(ctx, params, reads, out);
    if (rc) return rc;
    if (fetch) {
        rc = bwb_results_fetch(*out);
        if (rc) { delete *out; *out = nullptr; }
    }
    return rc;
}

int bwb_align(bwb_ctx *ctx, const bwb_params *params, const uint8_t *seq, const uint64_t *offsets, uint64_t n_reads,
              bwb_results **out) {
    if (!ctx || !params || !seq || !offsets || !out) return BWB_ERR_ARG;
    if (!ctx->have_index) return fail(ctx, BWB_ERR_NO_INDEX, "bwb_align before bwb_index_upload");
    bwb_reads *R = nullptr;
    int rc = bwb_reads_upload(ctx, seq, offsets, n_reads, &R);
    if (rc) return rc;
    rc = bwb_align_resident(ctx, params, R, 1, out);
    bwb_reads_free(R);
    return rc;
}

uint64_t bwb_results_num_reads(const bwb_results *r) { return r ? r->n_reads : 0; }
uint64_t bwb_results_num_hits(const bwb_results *r) {
    if (!r) return 0;
    uint64_t t = 0;
    for (uint64_t v : r->shard_total) t += v;
    return t;
}
const uint32_t *bwb_results_counts(const bwb_results *r) { return r && r->fetched ? r->counts.data() : nullptr; }
const bwb_hit *bwb_results_hits(const bwb_results *r) { return r && r->fetched ? r->hits.data() : nullptr; }
const bwb_loc *bwb_results_locations(const bwb_results *r) { return r && r->fetched && r->have_loc ? r->loc.data() : nullptr; }
int bwb_results_counters(const bwb_results *r, uint64_t out[8]) {
    if (!r || !out) return BWB_ERR_ARG;
    memcpy(out, r->counters, sizeof r->counters);
    return BWB_OK;
}
double bwb_results_kernel_ms(const bwb_results *r) { return r ? (double)r->kernel_ms : 0.0; }
double bwb_results_k3_ms(const bwb_results *r) { return r ? (double)r->k3_ms : 0.0; }
void bwb_results_free(bwb_results *r) { delete r; }
void bwb_free(void *p) { free(p); }

}  // extern "C"

// serialisation lives in aln_io.cpp; it needs the private layout of bwb_results
namespace bwb_host {
int ctx_device(const bwb_ctx *ctx, int *device_id, void **stream) {
    if (!ctx || ctx->dev.empty()) return BWB_ERR_ARG;
    *device_id = ctx->dev[0].id;
    *stream = (void *)ctx->dev[0].stream;
    return BWB_OK;
}
int ctx_fail(bwb_ctx *ctx, int code, const char *msg) { return fail(ctx, code, "%s", msg); }
bool ctx_index_options(const bwb_ctx *ctx, long long *chunk) {
    if (chunk) *chunk = ctx ? ctx->index_chunk : 0;
    return ctx && ctx->index_wide != 0;
}
const std::vector<uint32_t> &results_counts(const bwb_results *r) { return r->counts; }
const std::vector<bwb_hit> &results_hits(const bwb_results *r) { return r->hits; }
bool results_fetched(const bwb_results *r) { return r->fetched; }
const std::vector<bwb_loc> *results_loc(const bwb_results *r) { return r->have_loc ? &r->loc : nullptr; }
}  // namespace bwb_host

// host_common.h -- host-side containers shared by the index builder, the .bwt reader and the ABI.
#pragma once
#include <cstdint>
#include <vector>

struct bwb_ctx;

namespace bwb_host {

// In-memory image of a .bwt file (store_bwt layout, mg-aligner/bwt.c:66-82; fields of bwt_t,
// bwt.h:19-40).
struct HostIndex {
    uint64_t length = 0, num_words = 0, num_sa = 0, num_occ = 0, sa0_index = 0;
    uint64_t C[17] = {0};
    std::vector<uint32_t> bwt;
    std::vector<uint64_t> O;
    std::vector<uint64_t> SA;
};

int build_index_arrays(const uint8_t *text, uint64_t n, HostIndex &ix);
int write_bwt_file(const HostIndex &ix, const char *path);
int read_bwt_file(const char *path, HostIndex &ix, bool load_sa);
int prepare_index_text(const char *fasta_path, int write_ref_file, std::vector<uint8_t> &text);

// device 0 of a context and its launch stream (cudaStream_t as void*), for host modules outside bwb_abi.cu
int ctx_device(const struct ::bwb_ctx *ctx, int *device_id, void **stream);
int ctx_fail(struct ::bwb_ctx *ctx, int code, const char *msg);
// options of the device index builder: returns the "index_wide" flag, *chunk = "index_chunk" (0 = default)
bool ctx_index_options(const struct ::bwb_ctx *ctx, long long *chunk);

// .pre file of `bwbble align -P` (store_sa_interval_list / load_sa_interval_list, align.c:144-172,
// written row by row by precalc_sa_intervals, align.c:200-224): 4^12 records {int32 n; n x (u64 L, u64 U)}
int read_pre_file(const char *path, std::vector<uint32_t> &sizes, std::vector<uint64_t> &lu);
int write_pre_file(const char *path, const std::vector<uint32_t> &sizes, const std::vector<uint64_t> &lu);

}  // namespace bwb_host

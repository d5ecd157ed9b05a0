// precalc_io.cpp -- the .pre file of `bwbble align -P`: 4^12 records {int32 n; n x (u64 L, u64 U)}, one per
// 12-mer in next_read order (mg-aligner/align.c:144-172,188-224).  Host code, no device work.
#include <cstdio>
#include <cstring>

#include "host_common.h"

namespace bwb_host {

static const uint32_t kRows = 1u << 24;   // NUM_PRECALC, align.h:30

int read_pre_file(const char *path, std::vector<uint32_t> &sizes, std::vector<uint64_t> &lu) {
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    std::vector<char> iobuf(8u << 20);
    setvbuf(f, iobuf.data(), _IOFBF, iobuf.size());
    sizes.assign(kRows, 0);
    lu.clear();
    int rc = 0;
    for (uint32_t x = 0; x < kRows && !rc; x++) {
        int32_t n;
        if (fread(&n, 4, 1, f) != 1 || n < 0) { rc = -2; break; }
        sizes[x] = (uint32_t)n;
        if (n) {
            const size_t at = lu.size();
            lu.resize(at + 2 * (size_t)n);
            if (fread(lu.data() + at, 16, (size_t)n, f) != (size_t)n) rc = -2;
        }
    }
    fclose(f);
    return rc;
}

int write_pre_file(const char *path, const std::vector<uint32_t> &sizes, const std::vector<uint64_t> &lu) {
    if (sizes.size() != kRows) return -3;
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    std::vector<char> iobuf(8u << 20);
    setvbuf(f, iobuf.data(), _IOFBF, iobuf.size());
    size_t at = 0;
    int rc = 0;
    for (uint32_t x = 0; x < kRows && !rc; x++) {
        const int32_t n = (int32_t)sizes[x];
        if (fwrite(&n, 4, 1, f) != 1) rc = -2;
        if (n && !rc) {
            if (at + 2 * (size_t)n > lu.size() || fwrite(lu.data() + at, 16, (size_t)n, f) != (size_t)n) rc = -2;
            at += 2 * (size_t)n;
        }
    }
    if (fclose(f)) rc = -2;
    return rc;
}

}  // namespace bwb_host

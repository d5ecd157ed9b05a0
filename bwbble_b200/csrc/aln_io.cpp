// aln_io.cpp -- serialise device results exactly as the reference's alns2alnf_bin does
// (mg-aligner/align.c:345-382): per read  i32 n;  per hit  i32 score, u64 L, u64 U, i32 num_mm,
// i32 num_gapo, i32 num_gape, i32 aln_length, i32 n_pairs, n_pairs x i32 (state | run<<2), the edit
// path scanned from its LAST element to its first, run counter 16 bits wide (Q11).
// The device keeps only the gap runs of a path; the 256-byte path (align.h:118) is rebuilt here:
// zeros (STATE_M) everywhere except the runs, then cut to aln_length like the reference's memcpy.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "bwbble_b200.h"

struct bwb_results;
namespace bwb_host {
const std::vector<uint32_t> &results_counts(const bwb_results *r);
const std::vector<bwb_hit> &results_hits(const bwb_results *r);
bool results_fetched(const bwb_results *r);
}  // namespace bwb_host

namespace {

inline void put32(std::vector<uint8_t> &o, int32_t v) {
    const uint8_t *p = reinterpret_cast<const uint8_t *>(&v);
    o.insert(o.end(), p, p + 4);
}
inline void put64(std::vector<uint8_t> &o, uint64_t v) {
    const uint8_t *p = reinterpret_cast<const uint8_t *>(&v);
    o.insert(o.end(), p, p + 8);
}

void serialise(const bwb_results *r, std::vector<uint8_t> &o) {
    const auto &counts = bwb_host::results_counts(r);
    const auto &hits = bwb_host::results_hits(r);
    o.reserve(counts.size() * 4 + hits.size() * 52);
    size_t h = 0;
    uint8_t path[256];
    for (size_t rd = 0; rd < counts.size(); rd++) {
        put32(o, (int32_t)counts[rd]);
        for (uint32_t k = 0; k < counts[rd]; k++, h++) {
            const bwb_hit &t = hits[h];
            put32(o, t.score); put64(o, t.L); put64(o, t.U);
            put32(o, t.num_mm); put32(o, t.num_gapo); put32(o, t.num_gape); put32(o, t.aln_length);
            const int alen = t.aln_length;
            if (alen == 0) { put32(o, 0); continue; }
            memset(path, 0, sizeof path);
            for (int q = 0; q < t.n_runs && q < BWB_MAX_GAP_RUNS; q++)
                for (int s = 0; s < t.runs[q].len; s++) path[(uint8_t)(t.runs[q].start + s)] = t.runs[q].state;
            int32_t pairs[256];
            int np = 0;
            int state = path[alen - 1];
            uint16_t run = 1;
            for (int q = alen - 2; q >= 0; q--) {
                if (path[q] == state) run++;
                else { pairs[np++] = state | (run << 2); state = path[q]; run = 1; }
            }
            pairs[np++] = state | (run << 2);
            put32(o, np);
            for (int q = 0; q < np; q++) put32(o, pairs[q]);
        }
    }
}

}  // namespace

extern "C" int bwb_results_aln_bytes(const bwb_results *r, uint8_t **buf, uint64_t *len) {
    if (!r || !buf || !len) return BWB_ERR_ARG;
    if (!bwb_host::results_fetched(r)) return BWB_ERR_ARG;
    std::vector<uint8_t> o;
    serialise(r, o);
    uint8_t *p = (uint8_t *)malloc(o.size() ? o.size() : 1);
    if (!p) return BWB_ERR_IO;
    memcpy(p, o.data(), o.size());
    *buf = p;
    *len = o.size();
    return BWB_OK;
}

extern "C" int bwb_results_write_aln(const bwb_results *r, const char *path, int append) {
    if (!r || !path) return BWB_ERR_ARG;
    if (!bwb_host::results_fetched(r)) return BWB_ERR_ARG;
    std::vector<uint8_t> o;
    serialise(r, o);
    FILE *f = fopen(path, append ? "ab" : "wb");
    if (!f) return BWB_ERR_IO;
    const bool ok = fwrite(o.data(), 1, o.size(), f) == o.size();
    fclose(f);
    return ok ? BWB_OK : BWB_ERR_IO;
}

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
#include <cmath>
#include <string>
#include <vector>

#include "bwbble_b200.h"

struct bwb_results;
namespace bwb_host {
const std::vector<uint32_t> &results_counts(const bwb_results *r);
const std::vector<bwb_hit> &results_hits(const bwb_results *r);
bool results_fetched(const bwb_results *r);
const std::vector<bwb_loc> *results_loc(const bwb_results *r);
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


// ---------------------------------------------------------------------------------------------
// SAM output: what `bwbble aln2sam` does with the .aln records (alns2sam / eval_aln / mapq /
// print_aln2sam, mg-aligner/align.c:494-652,738-812), fed by K6's located positions.
// ---------------------------------------------------------------------------------------------
namespace {

struct Ann { std::string name; uint64_t start, end; };

// annf2ann, io.c:323-349: "<total>\t<n>\n" then "name\tstart\tend\n" (name = up to the first tab)
bool read_ann(const char *path, std::vector<Ann> &out) {
    FILE *f = fopen(path, "r");
    if (!f) return false;
    unsigned long long total;
    int n;
    if (fscanf(f, "%llu\t%d\n", &total, &n) != 2) { fclose(f); return false; }
    for (int i = 0; i < n; i++) {
        char name[1024];
        unsigned long long a, b;
        if (fscanf(f, "%1023[^\n\t]\t%llu\t%llu\n", name, &a, &b) < 3) { fclose(f); return false; }
        out.push_back(Ann{name, a, b});
    }
    fclose(f);
    return true;
}

// mapq(), align.c:738-746
int mapq(int top1, int top2, int num_mm, int max_mm) {
    if (top1 == 0) return 23;
    if (top1 > 1) return 0;
    if (num_mm == max_mm) return 25;
    if (top2 == 0) return 37;
    const int n = top2 >= 255 ? 255 : top2;
    const int q = (int)(4.343 * log((double)n) + 0.5);
    return 23 < q ? 0 : 23 - q;
}

// the edit path in SEARCH order (path[0] = first step), rebuilt from the gap runs
int search_path(const bwb_hit &t, uint8_t *path) {
    memset(path, 0, 256);
    for (int q = 0; q < t.n_runs && q < BWB_MAX_GAP_RUNS; q++)
        for (int s = 0; s < t.runs[q].len; s++) path[(uint8_t)(t.runs[q].start + s)] = t.runs[q].state;
    return t.aln_length;
}

}  // namespace

extern "C" int bwb_results_write_sam(const bwb_results *r, const char *ann_path, const char *const *names,
                                     const uint8_t *seq, const uint64_t *offsets, const char *const *quals,
                                     uint64_t index_length, int max_mm, const char *sam_path, int write_header,
                                     int append) {
    if (!r || !ann_path || !names || !seq || !offsets || !sam_path) return BWB_ERR_ARG;
    if (!bwb_host::results_fetched(r)) return BWB_ERR_ARG;
    const std::vector<bwb_loc> *loc = bwb_host::results_loc(r);
    if (!loc) return BWB_ERR_NO_INDEX;                       // no sampled SA was uploaded
    std::vector<Ann> ann;
    if (!read_ann(ann_path, ann)) return BWB_ERR_IO;
    FILE *f = fopen(sam_path, append ? "a" : "w");
    if (!f) return BWB_ERR_IO;
    if (write_header) {
        for (const Ann &a : ann) fprintf(f, "@SQ\tSN:%s\tLN:%d\n", a.name.c_str(), (int)(a.end - a.start + 1));
        fprintf(f, "@PG\tID:bwbble\tPN:bwbble\tVN:0.1-r01\n");
    }
    const auto &counts = bwb_host::results_counts(r);
    const auto &hits = bwb_host::results_hits(r);
    static const char nt[] = "AGCTN";
    static const uint8_t compl4[5] = {3, 2, 1, 0, 4};
    size_t h = 0;
    uint8_t path[256];
    std::string line;
    for (size_t rd = 0; rd < counts.size(); rd++) {
        const uint64_t o = offsets[rd];
        const int len = (int)(offsets[rd + 1] - o);
        const uint8_t *s = seq + o;
        if (counts[rd] == 0) {                               // unmapped (align.c:630-651)
            fprintf(f, "%s\t%d\t*\t0\t0\t*\t*\t0\t0\t", names[rd], 4);
            for (int i = 0; i < len; i++) fputc(nt[s[i] > 4 ? 4 : s[i]], f);
            fputc('\t', f);
            if (quals && quals[rd]) fputs(quals[rd], f); else fputc('*', f);
            fputc('\n', f);
            continue;
        }
        const bwb_hit &t = hits[h];
        h += counts[rd];
        const bwb_loc &lc = (*loc)[rd];
        // eval_aln (align.c:789-800): strand and position of hit 0
        const int alen = search_path(t, path);
        int n_ins = 0;
        for (int i = 0; i < alen; i++) n_ins += (path[i] == 1);
        int strand;
        uint64_t aln_pos;
        if (lc.ref_pos > (index_length - 1) / 2) {
            strand = 0;
            const uint64_t fwd_pos = (index_length - 1) - lc.ref_pos - 1;
            aln_pos = fwd_pos - (uint64_t)(alen - n_ins) + 1;
        } else {
            strand = 1;
            aln_pos = lc.ref_pos;
        }
        const int q = mapq(lc.top1, lc.top2, t.num_mm, max_mm);
        int seqid = -1;
        for (size_t i = 0; i < ann.size(); i++)
            if (aln_pos >= ann[i].start && aln_pos <= ann[i].end) { seqid = (int)i; break; }
        const char *rname = seqid >= 0 ? ann[seqid].name.c_str() : "*";
        const uint64_t start = seqid >= 0 ? ann[seqid].start : 0;
        fprintf(f, "%s\t%d\t%s\t%d\t%d\t", names[rd], strand ? 16 : 0, rname, (int)(aln_pos - start + 1), q);
        // CIGAR (align.c:578-607): the loaded path is the reverse of the search path; strand 1 reverses it
        // again; runs are then collected reading it back to front
        //   strand 0 -> search path front to back, strand 1 -> search path back to front
        {
            int i = strand ? alen - 1 : 0;
            const int step = strand ? -1 : 1;
            int cur = path[i], run = 0;
            for (int k = 0; k < alen; k++, i += step) {
                if (path[i] == cur) run++;
                else { fprintf(f, "%d%c", run, "MID"[cur]); cur = path[i]; run = 1; }
            }
            fprintf(f, "%d%c", run, "MID"[cur]);
        }
        fprintf(f, "\t*\t0\t0\t");
        if (strand) for (int i = len - 1; i >= 0; i--) fputc(nt[compl4[s[i] > 4 ? 4 : s[i]]], f);
        else for (int i = 0; i < len; i++) fputc(nt[s[i] > 4 ? 4 : s[i]], f);
        fputc('\t', f);
        if (quals && quals[rd]) {
            if (strand) { const size_t ql = strlen(quals[rd]); for (size_t i = ql; i > 0; i--) fputc(quals[rd][i - 1], f); }
            else fputs(quals[rd], f);
        } else fputc('*', f);
        fputc('\n', f);
    }
    fclose(f);
    return BWB_OK;
}

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
#include <thread>
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

// bytes one hit can take at most: 8 fixed fields (44 B) + n_pairs + up to 256 pairs
constexpr size_t HIT_MAX = 44 + 4 + 256 * 4;

inline uint8_t *put32(uint8_t *p, int32_t v) { memcpy(p, &v, 4); return p + 4; }
inline uint8_t *put64(uint8_t *p, uint64_t v) { memcpy(p, &v, 8); return p + 8; }

// reads [r0, r1) whose first hit is hits[h]; appends to o
void serialise_range(const std::vector<uint32_t> &counts, const std::vector<bwb_hit> &hits, size_t r0, size_t r1, size_t h,
                     std::vector<uint8_t> &o) {
    size_t nh = 0;
    for (size_t rd = r0; rd < r1; rd++) nh += counts[rd];
    o.resize((r1 - r0) * 4 + nh * 52 + HIT_MAX);         // 52 B = a hit without gaps; grown below when gaps make it longer
    uint8_t *p = o.data();
    uint8_t path[256];
    for (size_t rd = r0; rd < r1; rd++) {
        p = put32(p, (int32_t)counts[rd]);
        for (uint32_t k = 0; k < counts[rd]; k++, h++) {
            if ((size_t)(p - o.data()) + HIT_MAX + 4 > o.size()) {           // gapped hits carry more pairs
                const size_t used = (size_t)(p - o.data());
                o.resize(o.size() + o.size() / 4 + 4 * HIT_MAX);
                p = o.data() + used;
            }
            const bwb_hit &t = hits[h];
            p = put32(p, t.score); p = put64(p, t.L); p = put64(p, t.U);
            p = put32(p, t.num_mm); p = put32(p, t.num_gapo); p = put32(p, t.num_gape); p = put32(p, t.aln_length);
            const int alen = t.aln_length;
            if (alen == 0) { p = put32(p, 0); continue; }
            if (t.n_runs == 0) {                         // all STATE_M: one pair
                p = put32(p, 1);
                p = put32(p, (int32_t)((uint32_t)(uint16_t)alen << 2));
                continue;
            }
            memset(path, 0, sizeof path);
            for (int q = 0; q < t.n_runs && q < BWB_MAX_GAP_RUNS; q++)
                for (int s2 = 0; s2 < t.runs[q].len; s2++) path[(uint8_t)(t.runs[q].start + s2)] = t.runs[q].state;
            int32_t pairs[256];
            int np = 0;
            int state = path[alen - 1];
            uint16_t run = 1;
            for (int q = alen - 2; q >= 0; q--) {
                if (path[q] == state) run++;
                else { pairs[np++] = state | (run << 2); state = path[q]; run = 1; }
            }
            pairs[np++] = state | (run << 2);
            p = put32(p, np);
            for (int q = 0; q < np; q++) p = put32(p, pairs[q]);
        }
    }
    o.resize((size_t)(p - o.data()));
}

// the records of all reads, in input order, as a list of chunks (one per worker thread)
void serialise(const bwb_results *r, std::vector<std::vector<uint8_t>> &chunks) {
    const auto &counts = bwb_host::results_counts(r);
    const auto &hits = bwb_host::results_hits(r);
    const size_t n = counts.size();
    unsigned hw = std::thread::hardware_concurrency();
    size_t nt = n < (1u << 16) ? 1 : (hw ? (hw > 16 ? 16 : hw) : 4);
    chunks.assign(nt, std::vector<uint8_t>());
    std::vector<size_t> r0(nt + 1), h0(nt + 1, 0);
    for (size_t t = 0; t <= nt; t++) r0[t] = n * t / nt;
    {
        size_t h = 0, t = 1;
        for (size_t rd = 0; rd < n; rd++) {
            while (t <= nt && r0[t] == rd) h0[t++] = h;
            h += counts[rd];
        }
        while (t <= nt) h0[t++] = h;
    }
    if (nt == 1) { serialise_range(counts, hits, 0, n, 0, chunks[0]); return; }
    std::vector<std::thread> th;
    for (size_t t = 0; t < nt; t++)
        th.emplace_back([&, t] { serialise_range(counts, hits, r0[t], r0[t + 1], h0[t], chunks[t]); });
    for (auto &x : th) x.join();
}

}  // namespace

extern "C" int bwb_results_aln_bytes(const bwb_results *r, uint8_t **buf, uint64_t *len) {
    if (!r || !buf || !len) return BWB_ERR_ARG;
    if (!bwb_host::results_fetched(r)) return BWB_ERR_ARG;
    std::vector<std::vector<uint8_t>> chunks;
    serialise(r, chunks);
    size_t total = 0;
    for (auto &c : chunks) total += c.size();
    uint8_t *p = (uint8_t *)malloc(total ? total : 1);
    if (!p) return BWB_ERR_IO;
    size_t w = 0;
    for (auto &c : chunks) { memcpy(p + w, c.data(), c.size()); w += c.size(); }
    *buf = p;
    *len = total;
    return BWB_OK;
}

extern "C" int bwb_results_write_aln(const bwb_results *r, const char *path, int append) {
    if (!r || !path) return BWB_ERR_ARG;
    if (!bwb_host::results_fetched(r)) return BWB_ERR_ARG;
    std::vector<std::vector<uint8_t>> chunks;
    serialise(r, chunks);
    FILE *f = fopen(path, append ? "ab" : "wb");
    if (!f) return BWB_ERR_IO;
    bool ok = true;
    for (auto &c : chunks) ok = ok && fwrite(c.data(), 1, c.size(), f) == c.size();
    ok = (fclose(f) == 0) && ok;
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

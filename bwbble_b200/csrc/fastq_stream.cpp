// fastq_stream.cpp -- streaming FASTQ ingest for the device path.
//
// The reference loads ALL reads into memory before aligning (fastq2reads, io.c:410-515: one read_t of
// 408 bytes + three mallocs per read) -- ~0.9 kB per read, i.e. ~90 GB for the 100 M-read configs.
// Here the file is parsed in batches straight into the packed layout bwb_align consumes, every batch
// is aligned and its records are appended to the .aln (and optionally the SAM) file, so memory is
// bounded by the batch size (two batches: a reader thread parses batch k+1 while batch k is on the
// device and its records are written).  Parsing follows fastq2reads: records start at the next '@'; the name
// is the rest of that line (first 256 characters kept); the base line is mapped through nt4_table
// (io.h:113-130: A0 G1 C2 T3, anything else 4); then the '+' line; then the quality line, which
// must be as long as the base line.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "bwbble_b200.h"

namespace {

class Reader {
  public:
    explicit Reader(FILE *f) : f_(f), buf_(1 << 20), pos_(0), end_(0) {}
    int get() {
        if (pos_ == end_) {
            end_ = fread(buf_.data(), 1, buf_.size(), f_);
            pos_ = 0;
            if (end_ == 0) return EOF;
        }
        return (unsigned char)buf_[pos_++];
    }
  private:
    FILE *f_;
    std::vector<char> buf_;
    size_t pos_, end_;
};

inline uint8_t nt4(int c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'G': case 'g': return 1;
        case 'C': case 'c': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}

struct Batch {
    std::vector<uint8_t> seq;
    std::vector<uint64_t> off;
    std::vector<std::string> names, quals;
    void clear() { seq.clear(); off.assign(1, 0); names.clear(); quals.clear(); }
};

// returns 1 = read parsed, 0 = clean end of file, <0 = malformed
int next_read(Reader &in, Batch &b, bool keep_text) {
    int c;
    while ((c = in.get()) != EOF && c != '@') {}
    if (c == EOF) return 0;
    std::string name;
    while ((c = in.get()) != EOF && c != '\n') if (name.size() < 256) name.push_back((char)c);
    if (c == EOF) return BWB_ERR_IO;
    const size_t start = b.seq.size();
    while ((c = in.get()) != EOF && c != '\n') b.seq.push_back(nt4(c));
    if (c == EOF) return BWB_ERR_IO;
    const size_t len = b.seq.size() - start;
    while ((c = in.get()) != EOF && c != '+') {}
    if (c == EOF) return BWB_ERR_IO;
    while ((c = in.get()) != EOF && c != '\n') {}
    if (c == EOF) return BWB_ERR_IO;
    std::string qual;
    size_t qlen = 0;
    while ((c = in.get()) != EOF && c != '\n') { if (keep_text) qual.push_back((char)c); qlen++; }
    if (qlen != len) return BWB_ERR_ARG;          // "number of quality score symbols does not match" (io.c:497-500)
    b.off.push_back(b.seq.size());
    if (keep_text) { b.names.push_back(std::move(name)); b.quals.push_back(std::move(qual)); }
    return 1;
}

}  // namespace

extern "C" long long bwb_align_fastq(bwb_ctx *ctx, const bwb_params *params, const char *fastq_path, const char *aln_path,
                                     const char *sam_path, const char *ann_path, uint64_t index_length, int max_mm,
                                     uint64_t batch_reads) {
    if (!ctx || !params || !fastq_path || (!aln_path && !sam_path)) return BWB_ERR_ARG;
    if (sam_path && !ann_path) return BWB_ERR_ARG;
    if (batch_reads == 0) batch_reads = 1ull << 23;     // every launch ends in a ~0.35 s tail: amortise it
    FILE *f = fopen(fastq_path, "rb");
    if (!f) return BWB_ERR_IO;
    if (aln_path) remove(aln_path);                 // align.c:48
    if (aln_path) { FILE *t = fopen(aln_path, "wb"); if (!t) { fclose(f); return BWB_ERR_IO; } fclose(t); }
    // producer: parses up to batch_reads reads per batch, at most two parsed batches waiting
    struct Parsed { Batch b; int status = BWB_OK; bool eof = false; };
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::unique_ptr<Parsed>> ready;
    bool stop = false;
    const bool keep_text = sam_path != nullptr;
    std::thread producer([&] {
        Reader in(f);
        for (;;) {
            std::unique_ptr<Parsed> p(new Parsed());
            p->b.clear();
            while (p->b.off.size() - 1 < batch_reads) {
                const int st = next_read(in, p->b, keep_text);
                if (st == 0) { p->eof = true; break; }
                if (st < 0) { p->status = st; break; }
            }
            const bool last = p->eof || p->status != BWB_OK;
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return stop || ready.size() < 2; });
            if (stop) return;
            ready.push_back(std::move(p));
            cv.notify_all();
            if (last) return;
        }
    });
    long long total = 0;
    bool first = true;
    int rc = BWB_OK;
    for (;;) {
        std::unique_ptr<Parsed> p;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return !ready.empty(); });
            p = std::move(ready.front());
            ready.pop_front();
            cv.notify_all();
        }
        if (p->status != BWB_OK) { rc = p->status; break; }
        Batch &b = p->b;
        const uint64_t n = b.off.size() - 1;
        if (n == 0 && !first) break;
        bwb_results *res = nullptr;
        static const uint8_t none = 0;
        rc = bwb_align(ctx, params, n ? b.seq.data() : &none, b.off.data(), n, &res);
        if (rc != BWB_OK) break;
        if (aln_path) rc = bwb_results_write_aln(res, aln_path, 1);
        if (rc == BWB_OK && sam_path) {
            std::vector<const char *> nm(n ? n : 1), ql(n ? n : 1);
            for (uint64_t i = 0; i < n; i++) { nm[i] = b.names[i].c_str(); ql[i] = b.quals[i].c_str(); }
            rc = bwb_results_write_sam(res, ann_path, nm.data(), n ? b.seq.data() : &none, b.off.data(), ql.data(),
                                       index_length, max_mm, sam_path, first ? 1 : 0, first ? 0 : 1);
        }
        bwb_results_free(res);
        if (rc != BWB_OK) break;
        total += (long long)n;
        first = false;
        if (p->eof) break;
    }
    {
        std::lock_guard<std::mutex> lk(mu);
        stop = true;
        cv.notify_all();
    }
    producer.join();
    fclose(f);
    return rc == BWB_OK ? total : (long long)rc;
}

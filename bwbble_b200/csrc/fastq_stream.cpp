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

// Buffered input with two primitives the record grammar needs: "consume through the next occurrence of a
// character" and "the rest of the current line", both memchr over 4 MB blocks instead of a call per byte.
class Reader {
  public:
    explicit Reader(FILE *f) : f_(f), buf_(4u << 20), pos_(0), end_(0), eof_(false) {}
    // consume up to and including the next `c`; false at end of file
    bool skip_through(char c) {
        for (;;) {
            if (pos_ == end_ && !fill()) return false;
            const char *q = (const char *)memchr(buf_.data() + pos_, c, end_ - pos_);
            if (q) { pos_ = (size_t)(q - buf_.data()) + 1; return true; }
            pos_ = end_;
        }
    }
    // the rest of the current line (without its '\n', which is consumed); *p points into the buffer and stays valid
    // until the next call.  false if the file ends before a '\n'.
    bool rest_of_line(const char **p, size_t *n, bool eof_ends_line = false) {
        size_t from = pos_;
        for (;;) {
            const char *q = end_ > from ? (const char *)memchr(buf_.data() + from, '\n', end_ - from) : nullptr;
            if (q) {
                *p = buf_.data() + pos_;
                *n = (size_t)(q - (buf_.data() + pos_));
                pos_ = (size_t)(q - buf_.data()) + 1;
                return true;
            }
            // keep the partial line, read more behind it
            const size_t have = end_ - pos_;
            if (pos_ > 0) { memmove(buf_.data(), buf_.data() + pos_, have); pos_ = 0; end_ = have; }
            if (end_ == buf_.size()) buf_.resize(buf_.size() * 2);
            from = end_;
            size_t got = 0;
            if (!eof_) got = fread(buf_.data() + end_, 1, buf_.size() - end_, f_);
            if (got == 0) {
                eof_ = true;
                if (!eof_ends_line) return false;
                *p = buf_.data() + pos_;              // the file ends inside this line: it is the line (io.c:487-495)
                *n = end_ - pos_;
                pos_ = end_;
                return true;
            }
            end_ += got;
        }
    }
  private:
    bool fill() {
        if (eof_) return false;
        end_ = fread(buf_.data(), 1, buf_.size(), f_);
        pos_ = 0;
        if (end_ == 0) { eof_ = true; return false; }
        return true;
    }
    FILE *f_;
    std::vector<char> buf_;
    size_t pos_, end_;
    bool eof_;
};

struct Nt4Table {
    uint8_t t[256];
    Nt4Table() {
        memset(t, 4, sizeof t);
        t[(int)'A'] = t[(int)'a'] = 0; t[(int)'G'] = t[(int)'g'] = 1;
        t[(int)'C'] = t[(int)'c'] = 2; t[(int)'T'] = t[(int)'t'] = 3;
    }
};
const Nt4Table NT4;          // nt4_table, io.h:113-130

struct Batch {
    std::vector<uint8_t> seq;
    std::vector<uint64_t> off;
    std::vector<std::string> names, quals;
    void clear() { seq.clear(); off.assign(1, 0); names.clear(); quals.clear(); }
};

// returns 1 = read parsed, 0 = clean end of file, <0 = malformed
int next_read(Reader &in, Batch &b, bool keep_text) {
    const char *p;
    size_t n;
    if (!in.skip_through('@')) return 0;
    if (!in.rest_of_line(&p, &n)) return BWB_ERR_IO;
    if (keep_text) b.names.emplace_back(p, n < 256 ? n : 256);
    if (!in.rest_of_line(&p, &n)) return BWB_ERR_IO;
    const size_t start = b.seq.size(), len = n;
    b.seq.resize(start + len);
    uint8_t *dst = b.seq.data() + start;
    for (size_t i = 0; i < len; i++) dst[i] = NT4.t[(unsigned char)p[i]];
    if (!in.skip_through('+')) return BWB_ERR_IO;
    if (!in.rest_of_line(&p, &n)) return BWB_ERR_IO;
    in.rest_of_line(&p, &n, true);                // the quality line may end with the file
    if (n != len) return BWB_ERR_ARG;             // "number of quality score symbols does not match" (io.c:497-500)
    b.off.push_back(b.seq.size());
    if (keep_text) b.quals.emplace_back(p, n);
    return 1;
}

}  // namespace

// Host-only: the whole file through the same record grammar (no context, no device) -- what the Python mirror and
// the tests use, so that every entry point reads a FASTQ file the way fastq2reads does.  Arrays are malloc'ed
// (bwb_free); names / quals (optional) are the strings concatenated with a NUL after each.
extern "C" int bwb_fastq_parse(const char *fastq_path, uint8_t **seq, uint64_t **offsets, uint64_t *n_reads,
                               char **names, uint64_t *names_bytes, char **quals, uint64_t *quals_bytes) {
    if (!fastq_path || !seq || !offsets || !n_reads) return BWB_ERR_ARG;
    FILE *f = fopen(fastq_path, "rb");
    if (!f) return BWB_ERR_IO;
    Reader in(f);
    Batch b;
    b.clear();
    const bool keep = names != nullptr || quals != nullptr;
    int st;
    while ((st = next_read(in, b, keep)) == 1) {}
    fclose(f);
    if (st < 0) return st;
    const uint64_t n = b.off.size() - 1;
    *seq = (uint8_t *)malloc(b.seq.size() ? b.seq.size() : 1);
    *offsets = (uint64_t *)malloc(b.off.size() * 8);
    if (!*seq || !*offsets) return BWB_ERR_IO;
    memcpy(*seq, b.seq.data(), b.seq.size());
    memcpy(*offsets, b.off.data(), b.off.size() * 8);
    *n_reads = n;
    auto pack = [](const std::vector<std::string> &v, char **out, uint64_t *bytes) {
        size_t tot = 0;
        for (auto &x : v) tot += x.size() + 1;
        char *p = (char *)malloc(tot ? tot : 1);
        size_t w = 0;
        for (auto &x : v) { memcpy(p + w, x.data(), x.size()); w += x.size(); p[w++] = 0; }
        *out = p;
        if (bytes) *bytes = tot;
    };
    if (names) pack(b.names, names, names_bytes);
    if (quals) pack(b.quals, quals, quals_bytes);
    return BWB_OK;
}

extern "C" long long bwb_align_fastq(bwb_ctx *ctx, const bwb_params *params, const char *fastq_path, const char *aln_path,
                                     const char *sam_path, const char *ann_path, uint64_t index_length, int max_mm,
                                     uint64_t batch_reads) {
    if (!ctx || !params || !fastq_path || (!aln_path && !sam_path)) return BWB_ERR_ARG;
    if (sam_path && !ann_path) return BWB_ERR_ARG;
    if (batch_reads == 0) batch_reads = 1ull << 23;     // every launch ends in a ~0.35 s tail: amortise it
    // SURVEY Q6 (which stale D_seed a read no longer than the seed consults, see bwb_align): the serial driver's chain
    // runs through the whole file, so it is carried from launch to launch; the OpenMP driver restarts it per thread
    // chunk of every 262144-read batch of the FILE, so launches must hold whole batches
    if (params->n_threads > 1) batch_reads = (batch_reads + 0x3ffff) / 0x40000 * 0x40000;
    FILE *f = fopen(fastq_path, "rb");
    if (!f) return BWB_ERR_IO;
    bwb_set_option(ctx, "seed_carry", params->n_threads > 1 ? 0 : 1);
    if (aln_path) remove(aln_path);                 // align.c:48
    if (aln_path) { FILE *t = fopen(aln_path, "wb"); if (!t) { fclose(f); bwb_set_option(ctx, "seed_carry", 0); return BWB_ERR_IO; } fclose(t); }
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
    // consumer side: this thread aligns batch k while a writer thread serialises batch k-1 (results are host-side
    // once fetched, so they outlive the next launch) and the producer parses batch k+1
    struct Done { std::unique_ptr<Parsed> p; bwb_results *res = nullptr; bool first = false; };
    std::deque<Done> to_write;
    bool write_stop = false;
    int rc_w = BWB_OK;
    std::thread writer([&] {
        for (;;) {
            Done d;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return write_stop || !to_write.empty(); });
                if (to_write.empty()) return;
                d = std::move(to_write.front());
                to_write.pop_front();
                cv.notify_all();
            }
            int rcw = BWB_OK;
            Batch &b = d.p->b;
            const uint64_t n = b.off.size() - 1;
            static const uint8_t none = 0;
            if (rc_w == BWB_OK) {
                if (aln_path) rcw = bwb_results_write_aln(d.res, aln_path, 1);
                if (rcw == BWB_OK && sam_path) {
                    std::vector<const char *> nm(n ? n : 1), ql(n ? n : 1);
                    for (uint64_t i = 0; i < n; i++) { nm[i] = b.names[i].c_str(); ql[i] = b.quals[i].c_str(); }
                    rcw = bwb_results_write_sam(d.res, ann_path, nm.data(), n ? b.seq.data() : &none, b.off.data(), ql.data(),
                                                index_length, max_mm, sam_path, d.first ? 1 : 0, d.first ? 0 : 1);
                }
            }
            bwb_results_free(d.res);
            if (rcw != BWB_OK) { std::lock_guard<std::mutex> lk(mu); if (rc_w == BWB_OK) rc_w = rcw; cv.notify_all(); }
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
            if (rc_w != BWB_OK) { rc = rc_w; break; }
        }
        if (p->status != BWB_OK) { rc = p->status; break; }
        Batch &b = p->b;
        const uint64_t n = b.off.size() - 1;
        if (n == 0 && !first) break;
        bwb_results *res = nullptr;
        static const uint8_t none = 0;
        rc = bwb_align(ctx, params, n ? b.seq.data() : &none, b.off.data(), n, &res);
        if (rc != BWB_OK) break;
        const bool eof = p->eof;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return to_write.size() < 2; });
            Done d;
            d.p = std::move(p); d.res = res; d.first = first;
            to_write.push_back(std::move(d));
            cv.notify_all();
        }
        total += (long long)n;
        first = false;
        if (eof) break;
    }
    {
        std::lock_guard<std::mutex> lk(mu);
        write_stop = true;
        cv.notify_all();
    }
    writer.join();
    if (rc == BWB_OK && rc_w != BWB_OK) rc = rc_w;
    {
        std::lock_guard<std::mutex> lk(mu);
        stop = true;
        cv.notify_all();
    }
    producer.join();
    fclose(f);
    if (rc != BWB_OK) {
        // earlier batches are already on disk: do not leave a truncated file under the name of a complete one
        if (aln_path) rename(aln_path, (std::string(aln_path) + ".partial").c_str());
        if (sam_path) rename(sam_path, (std::string(sam_path) + ".partial").c_str());
    }
    bwb_set_option(ctx, "seed_carry", 0);
    return rc == BWB_OK ? total : (long long)rc;
}

// index_build.cpp -- host-side construction of the BWBBLE index files (<fasta>.bwt, <fasta>.ann).
//
// Out of the hot path but needed wherever the reference binary is not
// available (the GPU box has no /root/reference), so the bench and the tests can index the
// synthetic multi-genomes themselves.  The suffix array of a text is unique, so any correct
// builder reproduces the reference's files byte for byte; tests/test_index_build.py checks that
// against `bwbble index` run here and against tests/golden/.
//
// What is reproduced (file:line under /root/reference/mg-aligner):
//   io.c:190-321   fasta2ref : header = text after '>' up to 256 chars; letters upper-cased and
//                  mapped through nt16_table (unknown -> N=10); '$'(0) appended after EVERY
//                  record; .ann = "<fwd_len>\t<num_seq>\n" + "name\tstart\tend\n" per record;
//                  reverse complement (iupacCompl) appended.
//   is.c:214-243   is_bwt    : SA[0] = n (empty suffix first), BWT[i] = T[SA[i]-1], the row with
//                  SA[i]==0 stores code 0 and is remembered as sa0_index; SA sampled every 32 rows.
//   io.c:590-609   pack_word : 8 symbols per uint32, first symbol in the top nibble.
//   bwt.c:266-291  compute_C / compute_O (sentinel row excluded, checkpoint rows inclusive).
//   bwt.c:66-82    store_bwt : 5 x u64 header, C[17], bwt words, O rows, SA samples.
//
// The suffix sorter is an own implementation of induced sorting (SA-IS, Nong/Zhang/Chan 2009),
// written against the paper's description: L/S typing, LMS-substring naming, recursion on the
// reduced string, two induced-sort passes.  A unique smallest sentinel is appended to the text so
// that "shorter suffix sorts first" needs no special cases.

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "bwbble_b200.h"
#include "host_common.h"

namespace {

// ---------------------------------------------------------------------------------------------
// SA-IS over s[0..n) with s[n-1] the unique smallest symbol.  I is a signed index type.
// ---------------------------------------------------------------------------------------------
template <class S, class I>
class InducedSorter {
  public:
    InducedSorter(const S *s, I *sa, I n, I K) : s_(s), sa_(sa), n_(n), K_(K) {}

    void run() {
        std::vector<uint8_t> stype((size_t)n_);   // 1 = S-type, 0 = L-type
        stype[n_ - 1] = 1;
        for (I i = n_ - 2; i >= 0; --i)
            stype[i] = (s_[i] < s_[i + 1] || (s_[i] == s_[i + 1] && stype[i + 1])) ? 1 : 0;
        const uint8_t *t = stype.data();
        std::vector<I> bkt((size_t)K_);

        // pass 1: LMS suffixes at their bucket ends in text order, then induce -> LMS substrings sorted
        bucket_bounds(bkt.data(), true);
        for (I i = 0; i < n_; ++i) sa_[i] = -1;
        for (I i = 1; i < n_; ++i)
            if (is_lms(t, i)) sa_[--bkt[s_[i]]] = i;
        induce(t, bkt.data());

        // compact the sorted LMS substrings into sa_[0..m)
        I m = 0;
        for (I i = 0; i < n_; ++i)
            if (sa_[i] > 0 && is_lms(t, sa_[i])) sa_[m++] = sa_[i];
        // (position 0 is never LMS; the sentinel position n-1 always is)
        for (I i = m; i < n_; ++i) sa_[i] = -1;

        // name them: equal substrings get equal names
        I names = 0, prev = -1;
        for (I i = 0; i < m; ++i) {
            I pos = sa_[i];
            bool diff = false;
            if (prev < 0) diff = true;
            else {
                for (I d = 0;; ++d) {
                    if (pos + d >= n_ || prev + d >= n_ || s_[pos + d] != s_[prev + d] ||
                        t[pos + d] != t[prev + d]) { diff = true; break; }
                    if (d > 0 && (is_lms(t, pos + d) || is_lms(t, prev + d))) break;
                }
            }
            if (diff) { ++names; prev = pos; }
            sa_[m + pos / 2] = names - 1;
        }
        // reduced string into sa_[n-m .. n)
        I *s1 = sa_ + (n_ - m);
        for (I i = n_ - 1, j = n_ - 1; i >= m; --i)
            if (sa_[i] >= 0) sa_[j--] = sa_[i];

        // sort the reduced problem
        I *sa1 = sa_;
        if (names < m) {
            InducedSorter<I, I>(s1, sa1, m, names).run();
        } else {
            for (I i = 0; i < m; ++i) sa1[s1[i]] = i;
        }

        // map reduced suffixes back to text positions
        for (I i = 1, j = 0; i < n_; ++i)
            if (is_lms(t, i)) s1[j++] = i;
        for (I i = 0; i < m; ++i) sa1[i] = s1[sa1[i]];
        for (I i = m; i < n_; ++i) sa_[i] = -1;

        // pass 2: sorted LMS suffixes to their bucket ends (right to left), induce the rest
        bucket_bounds(bkt.data(), true);
        for (I i = m - 1; i >= 0; --i) {
            I j = sa_[i];
            sa_[i] = -1;
            sa_[--bkt[s_[j]]] = j;
        }
        induce(t, bkt.data());
    }

  private:
    static bool is_lms(const uint8_t *t, I i) { return i > 0 && t[i] && !t[i - 1]; }

    void bucket_bounds(I *bkt, bool ends) {
        for (I c = 0; c < K_; ++c) bkt[c] = 0;
        for (I i = 0; i < n_; ++i) ++bkt[s_[i]];
        I sum = 0;
        for (I c = 0; c < K_; ++c) {
            sum += bkt[c];
            bkt[c] = ends ? sum : sum - bkt[c];
        }
    }

    void induce(const uint8_t *t, I *bkt) {
        bucket_bounds(bkt, false);                 // L-types, left to right from bucket starts
        for (I i = 0; i < n_; ++i) {
            I j = sa_[i] - 1;
            if (sa_[i] > 0 && !t[j]) sa_[bkt[s_[j]]++] = j;
        }
        bucket_bounds(bkt, true);                  // S-types, right to left from bucket ends
        for (I i = n_ - 1; i >= 0; --i) {
            I j = sa_[i] - 1;
            if (sa_[i] > 0 && t[j]) sa_[--bkt[s_[j]]] = j;
        }
    }

    const S *s_;
    I *sa_;
    I n_, K_;
};

template <class I>
bool suffix_array(const uint8_t *text, uint64_t n, std::vector<I> &sa) {
    // shifted copy with the unique sentinel 0 at the end; codes 0..15 -> 1..16
    std::vector<uint8_t> s(n + 1);
    for (uint64_t i = 0; i < n; ++i) s[i] = (uint8_t)(text[i] + 1);
    s[n] = 0;
    sa.assign(n + 1, 0);
    InducedSorter<uint8_t, I>(s.data(), sa.data(), (I)(n + 1), (I)17).run();
    return true;
}

const uint8_t kNt16[128] = {
    // nt16_table, io.h:132-149 (ASCII half; >=128 maps to N)
    10,10,10,10,10,10,10,10,10,10,10,10,10,10,10,10, 10,10,10,10,10,10,10,10,10,10,10,10,10,10,10,10,
    10,10,10,10, 0,10,10,10,10,10,10,10,10,10,10,10, 10,10,10,10,10,10,10,10,10,10,10,10,10,10,10,10,
    10,15, 5, 7,13,10,10, 3, 9,10,10, 2,10, 8,10,10, 10,10,12, 4, 1,10,11,14,10, 6,10,10,10,10,10,10,
    10,15, 5, 7,13,10,10, 3, 9,10,10, 2,10, 8,10,10, 10,10,12, 4, 1,10,11,14,10, 6,10,10,10,10,10,10};
const uint8_t kCompl[16] = {0, 15, 8, 7, 4, 11, 12, 3, 2, 13, 10, 5, 6, 9, 14, 1};   // io.h:32

struct AnnRecord { std::string name; uint64_t start, end; };

// fasta2ref semantics on an in-memory file image
int parse_fasta(const std::vector<uint8_t> &file, std::vector<uint8_t> &text, std::vector<AnnRecord> &ann) {
    size_t p = 0, N = file.size();
    if (N == 0 || file[0] != '>') return BWB_ERR_IO;
    p = 1;
    text.clear();
    text.reserve(2 * N + 16);
    while (p <= N) {
        AnnRecord rec;
        // header: up to 256 chars, rest of the line skipped
        while (p < N && file[p] != '\n' && rec.name.size() < 256) rec.name.push_back((char)file[p++]);
        while (p < N && file[p] != '\n') ++p;
        if (p >= N) return BWB_ERR_IO;            // header without sequence
        uint64_t start = text.size();
        // sequence until '>' or EOF; the '\n' that ended the header is skipped by the same loop
        while (p < N && file[p] != '>') {
            uint8_t c = file[p++];
            if (c == '\n') continue;
            if (c >= 'a' && c <= 'z') c = (uint8_t)(c - 'a' + 'A');
            text.push_back(c < 128 ? kNt16[c] : 10);
        }
        text.push_back(0);                        // '$' after every record
        rec.start = start;
        rec.end = text.size() - 1;
        ann.push_back(rec);
        if (p >= N) break;
        ++p;                                      // consume '>'
    }
    return BWB_OK;
}

}  // namespace

namespace bwb_host {

int build_index_arrays(const uint8_t *text, uint64_t n, HostIndex &ix) {
    // text: fwd+rc codes, n symbols; BWT has n+1 rows
    const uint64_t length = n + 1;
    ix.length = length;
    ix.num_words = (length + 7) / 8;
    ix.num_occ = (length + 127) / 128;
    ix.num_sa = (length + 31) / 32;
    std::vector<uint8_t> bw(length);
    ix.SA.assign(ix.num_sa, 0);
    auto emit = [&](auto &sa) {
        for (uint64_t i = 0; i < length; ++i) {
            uint64_t v = (uint64_t)sa[i];
            if ((i & 31) == 0) ix.SA[i >> 5] = v;
            if (v == 0) { ix.sa0_index = i; bw[i] = 0; }
            else bw[i] = text[v - 1];
        }
    };
    if (length < (1ull << 31) - 2) {
        std::vector<int32_t> sa;
        suffix_array<int32_t>(text, n, sa);
        emit(sa);
    } else {
        std::vector<int64_t> sa;
        suffix_array<int64_t>(text, n, sa);
        emit(sa);
    }
    // pack
    ix.bwt.assign(ix.num_words, 0);
    for (uint64_t i = 0; i < length; ++i) ix.bwt[i >> 3] |= (uint32_t)bw[i] << (28 - 4 * (i & 7));
    // C
    for (int c = 0; c < 17; ++c) ix.C[c] = 0;
    for (uint64_t i = 0; i < length; ++i)
        if (i != ix.sa0_index) ix.C[bw[i] + 1]++;
    for (int c = 1; c < 17; ++c) ix.C[c] += ix.C[c - 1];
    // O
    ix.O.assign(ix.num_occ * 16, 0);
    uint64_t occ[16] = {0};
    for (uint64_t i = 0; i < length; ++i) {
        if (i != ix.sa0_index) occ[bw[i]]++;
        if ((i & 127) == 0) memcpy(&ix.O[(i >> 7) * 16], occ, sizeof occ);
    }
    return BWB_OK;
}

int write_bwt_file(const HostIndex &ix, const char *path) {
    FILE *f = fopen(path, "wb");
    if (!f) return BWB_ERR_IO;
    uint64_t hdr[5] = {ix.length, ix.num_words, ix.num_sa, ix.num_occ, ix.sa0_index};
    bool ok = fwrite(hdr, 8, 5, f) == 5 && fwrite(ix.C, 8, 17, f) == 17 &&
              fwrite(ix.bwt.data(), 4, ix.num_words, f) == ix.num_words &&
              fwrite(ix.O.data(), 8, ix.num_occ * 16, f) == ix.num_occ * 16 &&
              fwrite(ix.SA.data(), 8, ix.num_sa, f) == ix.num_sa;
    fclose(f);
    return ok ? BWB_OK : BWB_ERR_IO;
}

int read_bwt_file(const char *path, HostIndex &ix, bool load_sa) {
    FILE *f = fopen(path, "rb");
    if (!f) return BWB_ERR_IO;
    uint64_t hdr[5];
    bool ok = fread(hdr, 8, 5, f) == 5 && fread(ix.C, 8, 17, f) == 17;
    if (ok) {
        ix.length = hdr[0]; ix.num_words = hdr[1]; ix.num_sa = hdr[2]; ix.num_occ = hdr[3]; ix.sa0_index = hdr[4];
        ix.bwt.resize(ix.num_words);
        ix.O.resize(ix.num_occ * 16);
        ok = fread(ix.bwt.data(), 4, ix.num_words, f) == ix.num_words &&
             fread(ix.O.data(), 8, ix.num_occ * 16, f) == ix.num_occ * 16;
        if (ok && load_sa) {
            ix.SA.resize(ix.num_sa);
            ok = fread(ix.SA.data(), 8, ix.num_sa, f) == ix.num_sa;
        }
    }
    fclose(f);
    return ok ? BWB_OK : BWB_ERR_IO;
}

}  // namespace bwb_host

namespace bwb_host {

// Everything of `bwbble index` before the suffix sort (fasta2ref, io.c; bwt.c:29-63): parse the FASTA, write
// <fasta>.ann (and .ref), return the text = forward codes + '$' per record, followed by its reverse complement.
int prepare_index_text(const char *fasta_path, int write_ref_file, std::vector<uint8_t> &text) {
    if (!fasta_path) return BWB_ERR_ARG;
    FILE *f = fopen(fasta_path, "rb");
    if (!f) return BWB_ERR_IO;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> file((size_t)sz);
    if (sz > 0 && fread(file.data(), 1, (size_t)sz, f) != (size_t)sz) { fclose(f); return BWB_ERR_IO; }
    fclose(f);

    std::vector<AnnRecord> ann;
    int rc = parse_fasta(file, text, ann);
    if (rc) return rc;
    file.clear(); file.shrink_to_fit();
    const uint64_t fwd = text.size();

    std::string base(fasta_path);
    {   // .ann
        FILE *a = fopen((base + ".ann").c_str(), "wb");
        if (!a) return BWB_ERR_IO;
        fprintf(a, "%llu\t%d\n", (unsigned long long)fwd, (int)ann.size());
        for (auto &r : ann)
            fprintf(a, "%s\t%llu\t%llu\n", r.name.c_str(), (unsigned long long)r.start, (unsigned long long)r.end);
        fclose(a);
    }
    text.resize(2 * fwd);
    for (uint64_t i = 0; i < fwd; ++i) text[2 * fwd - 1 - i] = kCompl[text[i]];
    if (write_ref_file) {
        FILE *r = fopen((base + ".ref").c_str(), "wb");
        if (!r) return BWB_ERR_IO;
        fwrite(text.data(), 1, text.size(), r);
        fclose(r);
    }
    return BWB_OK;
}

}  // namespace bwb_host

extern "C" int bwb_index_build(const char *fasta_path, int write_ref_file) {
    std::vector<uint8_t> text;
    int rc = bwb_host::prepare_index_text(fasta_path, write_ref_file, text);
    if (rc) return rc;
    std::string base(fasta_path);
    bwb_host::HostIndex ix;
    rc = bwb_host::build_index_arrays(text.data(), text.size(), ix);
    if (rc) return rc;
    return bwb_host::write_bwt_file(ix, (base + ".bwt").c_str());
}

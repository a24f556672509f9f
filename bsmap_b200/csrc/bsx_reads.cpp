// bsx_reads.cpp -- host ingest: ReadClass::CheckFile / LoadBatchReads (reads.cpp:13-119) for FASTA / FASTQ.
//
// The reference pulls reads through ifstream `>>` / getline one token at a time under a mutex.  Here the
// file is memory-mapped and a batch is cut in two steps: one memchr pass finds the line starts, then
// `threads` host threads validate and slice their share of records straight into the batch buffers the
// GPU upload reads from.  A record is taken by the line cutter only when the token reader would load
// exactly the same thing (header char in column 0, one token per line, nothing but blanks after it);
// anything else -- blank lines, indented headers, wrapped sequences, trailing junk -- is handed to a
// token reader that restates the reference's stream semantics, record by record, until the input is
// regular again.  No copies are made of names / bases / qualities: the batch holds views into the map.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <chrono>
#include <atomic>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#include "bsx_internal.h"

namespace {

inline bool ws(int c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }

// token reader with ifstream `>>` / getline semantics over the mapped bytes
struct Tok {
    const char *p; size_t n, pos;
    int get() { return pos < n ? (unsigned char)p[pos++] : -1; }
    int nonws() { int c; while ((c = get()) >= 0 && ws(c)) {} return c; }
    bool token(std::string &s) {
        s.clear();
        int c = nonws();
        if (c < 0) return false;
        do { s.push_back((char)c); c = get(); } while (c >= 0 && !ws(c));
        if (c >= 0) pos--;
        return true;
    }
    void skipline() { const void *q = pos < n ? memchr(p + pos, '\n', n - pos) : nullptr; pos = q ? (size_t)((const char *)q - p) + 1 : n; }
};

// one record by the reference's rules (reads.cpp:93-113); false at end of input
bool slow_record(bsx_reads *r, std::string &nm, std::string &sq, std::string &ql) {
    Tok t{r->p, r->n, r->pos};
    std::string tok;
    const int c = t.nonws();
    if (c < 0) { r->pos = t.pos; return false; }
    if (!t.token(nm)) { r->pos = t.pos; return false; }
    t.skipline();
    if (!t.token(sq)) sq.clear();
    if (r->kind == 0) { t.token(tok); t.skipline(); if (!t.token(ql)) ql.clear(); }
    else ql.assign(sq.size(), (char)(r->zero_qual + 40));   // zero_qual + default_qual (reads.cpp:108)
    if ((int)sq.size() > r->max_readlen) { sq.erase(r->max_readlen); if ((int)ql.size() > r->max_readlen) ql.erase(r->max_readlen); }
    r->pos = t.pos;
    return true;
}

// [b, e) holds exactly one token starting in column 0, followed by blanks only
inline bool one_token(const char *b, const char *e, uint32_t *len) {
    if (b == e || ws((unsigned char)*b)) return false;
    const char *q = b + 1;
    while (q < e && !ws((unsigned char)*q)) q++;
    *len = (uint32_t)(q - b);
    while (q < e) { if (!ws((unsigned char)*q)) return false; q++; }
    return true;
}

inline void put_seq(char *seqs, uint16_t *lens, uint32_t stride, size_t slot, const char *s, uint32_t l) {
    if (!seqs) return;
    const uint32_t m = l < stride ? l : stride;
    char *d = seqs + slot * (size_t)stride;
    memcpy(d, s, m);
    memset(d + m, 0, stride - m);
    lens[slot] = (uint16_t)m;
}

// Line starts of a window of the file, found by `threads` memchr scanners.  The starts stay where the scanners put them --
// one array per scanner, r->scan_parts -- and are addressed as one sequence: entry 0 is the window's first byte, then the
// parts in order, then (when the file's last line has no line feed) one sentinel n + 1; line i is [at(i), at(i + 1) - 1).
// (Merging them into one array was a 4 MB serial copy per batch: a third of the cutter's time on sixteen threads.)
struct LineIndex {
    const std::vector<std::vector<uint64_t>> *part; std::vector<size_t> first;   // first[t] = global index of part t's entry 0
    uint64_t head, tail; bool has_tail; size_t total;
    uint64_t at(size_t g) const {
        if (g == 0) return head;
        if (has_tail && g == total - 1) return tail;
        size_t t = (size_t)(std::upper_bound(first.begin(), first.end(), g) - first.begin()) - 1;
        return (*part)[t][g - first[t]];
    }
    // entries g .. g + k of the sequence, k <= 4 (a record's line starts and the start of the next record)
    void get(size_t g, int k, uint64_t *L) const {
        if (g == 0 || (has_tail && g + (size_t)k >= total - 1)) { for (int i = 0; i <= k; i++) L[i] = at(g + (size_t)i); return; }
        size_t t = (size_t)(std::upper_bound(first.begin(), first.end(), g) - first.begin()) - 1, o = g - first[t];
        for (int i = 0; i <= k; i++) {
            while (o >= (*part)[t].size()) { t++; o = 0; }
            L[i] = (*part)[t][o++];
        }
    }
};

LineIndex scan_lines(bsx_reads *r, size_t wend, int threads) {
    const char *p = r->p;
    const size_t pos = r->pos, span = wend - pos;
    if ((size_t)threads > span / 65536 + 1) threads = (int)(span / 65536 + 1);
    std::vector<std::vector<uint64_t>> &part = r->scan_parts;
    if (part.size() < (size_t)threads) part.resize((size_t)threads);
    for (auto &v : part) v.clear();
    bsx_parallel(threads, span, [&, p, pos](int t, size_t b, size_t e) {
        std::vector<uint64_t> &v = part[t];
        v.reserve((e - b) / 24 + 16);
        for (size_t q = pos + b; q < pos + e;) {
            const void *h = memchr(p + q, '\n', pos + e - q);
            if (!h) break;
            q = (size_t)((const char *)h - p) + 1;
            v.push_back(q);
        }
    });
    LineIndex li;
    li.part = &part; li.head = pos; li.first.resize(part.size());
    size_t tot = 1; uint64_t last = pos;
    for (size_t t = 0; t < part.size(); t++) { li.first[t] = tot; tot += part[t].size(); if (!part[t].empty()) last = part[t].back(); }
    li.has_tail = wend == r->n && last != r->n;          // last line without a line feed
    li.tail = r->n + 1;
    li.total = tot + (li.has_tail ? 1 : 0);
    return li;
}

// Cut up to `want` regular records starting at r->pos (which must be a line start); returns how many
// were taken.  *bad = stopped in front of a record the token reader has to look at (or at the end of
// the input); otherwise the scan window was merely short and the caller comes back for more.
uint32_t fast_batch(bsx_reads *r, uint32_t want, uint32_t stride, char *seqs, uint16_t *lens, size_t base, int threads, bool *bad) {
    const int lpr = r->kind == 0 ? 4 : 2;
    const char hdr = r->kind == 0 ? '@' : '>';
    *bad = true;
    if (r->pos >= r->n) return 0;
    if (r->rec_bytes <= 0) {   // size the first window from the first record
        size_t q = r->pos; int k = 0;
        while (k < lpr && q < r->n) { const void *h = memchr(r->p + q, '\n', r->n - q); q = h ? (size_t)((const char *)h - r->p) + 1 : r->n; k++; }
        r->rec_bytes = (double)(q - r->pos) + 1;
    }
    const double wbytes = (double)want * r->rec_bytes * 1.03 + 4096;
    const size_t wend = wbytes >= (double)(r->n - r->pos) ? r->n : r->pos + (size_t)wbytes;
    if (wend == r->n && !r->stream_eof) { r->rec_bytes *= 2; *bad = false; return 0; }   // streamed input: the caller widens the window first
#ifdef MADV_POPULATE_READ
    // map the window's pages with one call: faulting them in one by one from all cutter threads serialises on the
    // address-space lock (measured: 8 threads no faster than 1)
    if (r->mapped) { const size_t a0 = r->pos & ~(size_t)4095; madvise((void *)(r->p + a0), wend - a0, MADV_POPULATE_READ); }
#endif
    const LineIndex li = scan_lines(r, wend, threads);
    const size_t n_lines = li.total - 1;
    const bool short_window = wend < r->n && n_lines / lpr < want;
    const size_t nrec = std::min<size_t>(n_lines / lpr, want);
    if (nrec == 0) { *bad = !short_window; if (short_window) r->rec_bytes *= 2; return 0; }
    std::atomic<size_t> first_bad(nrec);
    const char *p = r->p;
    const int mrl = r->max_readlen;
    bsx_view *vn = r->name.data() + base, *vs = r->seq.data() + base, *vq = r->qual.data() + base;
    const char *qfill = r->qual_fill.data();
    bsx_parallel(threads, nrec, [&, p, hdr, lpr, mrl, stride, seqs, lens, base, vn, vs, vq, qfill](int, size_t b, size_t e) {
        // this thread's records: a cursor over the scanners' arrays (the few records that touch the sequence's ends go through at())
        for (size_t i = b; i < e; i++) {
            if (i >= first_bad.load(std::memory_order_relaxed)) return;
            uint64_t L[5];
            li.get(i * (size_t)lpr, lpr, L);
            const char *h0 = p + L[0], *h1 = p + L[1] - 1;          // header line without its '\n'
            bool ok = h1 - h0 >= 2 && h0[0] == hdr && !ws((unsigned char)h0[1]);
            uint32_t nl = 0, sl = 0, ql = 0;
            if (ok) { const char *t = h0 + 1; while (t < h1 && !ws((unsigned char)*t)) t++; nl = (uint32_t)(t - h0 - 1); }
            ok = ok && one_token(p + L[1], p + L[2] - 1, &sl);
            if (ok && lpr == 4) {
                const char *c0 = p + L[2], *c1 = p + L[3] - 1;
                ok = c1 > c0 && !ws((unsigned char)*c0) && one_token(p + L[3], p + L[4] - 1, &ql);
            }
            if (!ok) { size_t cur = first_bad.load(); while (i < cur && !first_bad.compare_exchange_weak(cur, i)) {} return; }
            if ((int)sl > mrl) { sl = (uint32_t)mrl; if (lpr == 4 && (int)ql > mrl) ql = (uint32_t)mrl; }
            vn[i] = bsx_view{h0 + 1, nl};
            vs[i] = bsx_view{p + L[1], sl};
            vq[i] = lpr == 4 ? bsx_view{p + L[3], ql} : bsx_view{qfill, sl};
            put_seq(seqs, lens, stride, base + i, p + L[1], sl);
        }
    });
    const size_t took = first_bad.load();
    const uint64_t end_at = li.at(took * (size_t)lpr);
    r->pos = (size_t)std::min<uint64_t>(end_at, r->n);
    r->n_fast += took;
    if (took) r->rec_bytes = (double)(end_at - li.head) / (double)took;
    *bad = took < nrec || !short_window;
    return (uint32_t)took;
}

// after token-reader records: may the line cutter resume?  (only blanks up to the end of the line)
bool resume_at_line_start(bsx_reads *r) {
    size_t q = r->pos;
    while (q < r->n && r->p[q] != '\n') { if (!ws((unsigned char)r->p[q])) return false; q++; }
    r->pos = q < r->n ? q + 1 : r->n;
    return true;
}

}  // namespace

// gzip'ed input (an extension over the reference): inflated into memory, then handled like a map
static bool is_gzip(const char *p, size_t n) { return n >= 2 && (unsigned char)p[0] == 0x1f && (unsigned char)p[1] == 0x8b; }
static int gunzip_file(const char *path, std::vector<char> &out) {
    gzFile g = gzopen(path, "rb");
    if (!g) return BSX_ERR_IO;
    gzbuffer(g, 1 << 20);
    out.clear();
    std::vector<char> buf((size_t)1 << 22);
    int got;
    while ((got = gzread(g, buf.data(), (unsigned)buf.size())) > 0) out.insert(out.end(), buf.begin(), buf.begin() + got);
    const bool bad = got < 0;
    gzclose(g);
    return bad ? BSX_ERR_IO : BSX_OK;
}

int bsx_inflate_file(const char *path, std::vector<char> &out) { return gunzip_file(path, out); }

// Streamed input: make the window hold at least `need` unread bytes, or everything up to the end of the stream.  The unread
// tail moves to the front of a fresh window; the old one stays alive through r->keep for the views of the current batch.
static void stream_ensure(bsx_reads *r, size_t need) {
    if (r->stream_eof || r->n - r->pos >= need) return;
    const size_t have = r->n - r->pos;
    need += (size_t)8 << 20;                                 // refills stay rare when the caller advances in small steps
    auto nw = std::make_shared<std::vector<char>>(need);
    if (have) memcpy(nw->data(), r->p + r->pos, have);
    r->win_starts_line = r->pos == 0 ? r->win_starts_line : r->p[r->pos - 1] == '\n';
    size_t fill = have;
    while (fill < need) {
        const int got = gzread((gzFile)r->gz, nw->data() + fill, (unsigned)std::min<size_t>(need - fill, (size_t)1 << 30));
        if (got <= 0) {
            // the end of the stream, or an error: zlib reports a truncated member as Z_BUF_ERROR at this point, damaged data as Z_DATA_ERROR
            int zerr = 0;
            const char *msg = gzerror((gzFile)r->gz, &zerr);
            if (got < 0 || (zerr != Z_OK && zerr != Z_STREAM_END)) { r->io_error = true; bsx_set_error("read input: %s", msg && *msg ? msg : "gzip stream ended unexpectedly"); }
            r->stream_eof = true;
            break;
        }
        fill += (size_t)got;
    }
    nw->resize(fill);
    r->win = nw; r->keep.push_back(nw);
    r->p = nw->data(); r->n = fill; r->pos = 0;
}
static inline bool at_line_start(const bsx_reads *r) { return r->pos == 0 ? r->win_starts_line : r->p[r->pos - 1] == '\n'; }

int bsx_host_threads(int requested) {
    if (requested > 0) return requested > 64 ? 64 : requested;
    if (const char *e = getenv("BSX_THREADS")) { int v = atoi(e); if (v > 0) return v > 64 ? 64 : v; }
    unsigned hc = std::thread::hardware_concurrency();
    if (hc == 0) hc = 4;
    return hc > 32 ? 32 : (int)hc;
}

extern "C" int bsx_reads_open(const char *path, int zero_qual, int max_readlen, bsx_reads **out) {
    if (!path || !out) { bsx_set_error("bsx_reads_open: bad argument"); return BSX_ERR_ARG; }
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { bsx_set_error("failed to open read file: %s", path); return BSX_ERR_IO; }
    bsx_reads *r = new bsx_reads();
    r->fd = fd; r->zero_qual = zero_qual; r->max_readlen = max_readlen;
    struct stat st;
    const bool regular = fstat(fd, &st) == 0 && S_ISREG(st.st_mode);
    if (regular) {
        r->n = (size_t)st.st_size;
        if (r->n) {
            void *m = mmap(nullptr, r->n, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m == MAP_FAILED) { close(fd); delete r; bsx_set_error("mmap failed: %s", path); return BSX_ERR_IO; }
            madvise(m, r->n, MADV_SEQUENTIAL); madvise(m, r->n, MADV_WILLNEED);
            r->p = (const char *)m; r->mapped = true;
        }
    }
    if (!regular || is_gzip(r->p, r->n)) {
        // gzip'ed files and pipes are streamed through zlib (which passes plain data through): a window at a time
        if (r->mapped) { munmap((void *)r->p, r->n); r->mapped = false; }
        r->p = nullptr; r->n = 0;
        r->gz = regular ? (void *)gzopen(path, "rb") : (void *)gzdopen(fd, "rb");
        if (!r->gz) { close(fd); delete r; bsx_set_error("failed to open read file: %s", path); return BSX_ERR_IO; }
        if (!regular) r->fd = -1;                              // gzclose closes the descriptor it was given
        gzbuffer((gzFile)r->gz, 1 << 20);
        r->stream_eof = false;
        stream_ensure(r, (size_t)1 << 20);
    }
    // CheckFile (reads.cpp:19-50): the first non-blank character decides; anything else is tried as BAM
    size_t q = 0; while (q < r->n && ws((unsigned char)r->p[q])) q++;
    const int c = q < r->n ? r->p[q] : -1;
    if (r->n >= 12 && memcmp(r->p, "BAM\1", 4) == 0) {
        // BAM (BGZF members are gzip members, streamed like any gzip input): skip the header text and the reference dictionary
        r->kind = 3;
        auto have = [&](size_t upto) { if (upto > r->n) stream_ensure(r, upto); return upto <= r->n; };   // pos is 0: offsets stay valid
        auto i32 = [&](size_t at) { int32_t v; memcpy(&v, r->p + at, 4); return v; };
        size_t at = 4;
        const int32_t l_text = i32(at); at += 4 + (size_t)std::max(l_text, 0);
        bool ok = have(at + 4);
        if (ok) {
            const int32_t n_ref = i32(at); at += 4;
            for (int32_t k = 0; ok && k < n_ref; k++) { ok = have(at + 4); if (ok) { const int32_t l_name = i32(at); at += 4 + (size_t)std::max(l_name, 0) + 4; ok = have(at); } }
        }
        if (!ok) { bsx_reads_close(r); bsx_set_error("truncated BAM header: %s", path); return BSX_ERR_IO; }
        r->pos = at;
    } else if (c == '>') r->kind = 1; else if (c == '@') r->kind = 0;
    else { bsx_reads_close(r); bsx_set_error("fatal error: unrecognizable format of reads file."); return BSX_ERR_ARG; }
    r->qual_fill.assign((size_t)std::max(max_readlen, 1), (char)(zero_qual + 40));
    *out = r;
    return BSX_OK;
}

extern "C" void bsx_reads_close(bsx_reads *r) {
    if (!r) return;
    if (r->mapped) munmap((void *)r->p, r->n);
    if (r->gz) gzclose((gzFile)r->gz);
    if (r->fd >= 0) close(r->fd);
    delete r;
}

extern "C" int bsx_reads_kind(const bsx_reads *r) { return r ? r->kind : -1; }
extern "C" int bsx_reads_failed(const bsx_reads *r) { return r && r->io_error ? 1 : 0; }
extern "C" void bsx_reads_force_token_reader(bsx_reads *r, int on) { if (r) r->force_slow = on != 0; }

extern "C" void bsx_reads_skip(bsx_reads *r, uint64_t n_reads) {
    if (!r) return;
    if (r->kind == 3) return;   // reference quirk: CheckFile's -B skip covers _file_format 0..2 only, BAM is 3 (reads.cpp:54-75)
    const uint64_t nl = n_reads * (r->kind == 0 ? 4u : 2u);
    for (uint64_t i = 0; i < nl; i++) {
        stream_ensure(r, (size_t)4 << 20);
        r->keep.clear();                                     // no views point into the windows left behind
        if (r->pos >= r->n) break;
        for (;;) {   // a line may run past the window of a streamed input
            const void *q = memchr(r->p + r->pos, '\n', r->n - r->pos);
            if (q) { r->pos = (size_t)((const char *)q - r->p) + 1; break; }
            if (r->stream_eof) { r->pos = r->n; break; }
            stream_ensure(r, (r->n - r->pos) * 2 + ((size_t)4 << 20));
        }
    }
}

// BAM records (reads.cpp:120-143): name = qname, bases through bam_nt16_rev_table, qualities + 33, truncated to
// max_readlen.  readset 1 (file a of a pair) takes a record and skips its mate, readset 2 skips one and takes the next.
static uint32_t bam_batch(bsx_reads *r, uint32_t want, uint32_t stride, char *seqs, uint16_t *lens) {
    static const char nt16[] = "=ACMGRSVTWYHKDBN";
    auto record = [&](size_t &pos, const unsigned char *&body, uint32_t &bytes) {   // false at the end of the file; pos is r->pos
        if (!r->stream_eof && r->n - pos < 4 + 65536) stream_ensure(r, (size_t)16 << 20);   // streamed: the records are copied out, no window is kept
        if (pos + 4 > r->n) return false;
        int32_t bs; memcpy(&bs, r->p + pos, 4);
        if (bs >= 32 && pos + 4 + (size_t)bs > r->n) stream_ensure(r, 4 + (size_t)bs);
        if (bs < 32 || pos + 4 + (size_t)bs > r->n) return false;
        body = (const unsigned char *)r->p + pos + 4; bytes = (uint32_t)bs; pos += 4 + (size_t)bs;
        return true;
    };
    std::vector<std::string> &st = r->slow_store;
    st.clear(); st.reserve((size_t)want * 3);
    uint32_t got = 0;
    const unsigned char *b; uint32_t nb;
    while (got < want) {
        if (r->readset == 2 && !record(r->pos, b, nb)) break;
        if (!record(r->pos, b, nb)) break;
        const uint32_t l_name = b[8], n_cig = (uint32_t)b[12] | ((uint32_t)b[13] << 8);
        int32_t l_seq; memcpy(&l_seq, b + 16, 4);
        const size_t off_seq = 32 + (size_t)l_name + 4 * (size_t)n_cig, off_qual = off_seq + ((size_t)std::max(l_seq, 0) + 1) / 2;
        if (l_seq < 0 || off_qual + (size_t)l_seq > nb) break;
        const uint32_t l = (uint32_t)std::min<int64_t>(l_seq, r->max_readlen);
        std::string nm((const char *)b + 32), sq(l, 'N'), ql(l, '!');
        for (uint32_t i = 0; i < l; i++) { sq[i] = nt16[(b[off_seq + (i >> 1)] >> ((~i & 1) << 2)) & 15]; ql[i] = (char)(b[off_qual + i] + 33); }
        st.push_back(std::move(nm)); st.push_back(std::move(sq)); st.push_back(std::move(ql));
        if (r->readset == 1 && !record(r->pos, b, nb)) { st.resize(st.size() - 3); break; }   // the reference drops a last read without a mate
        got++;
    }
    for (uint32_t i = 0; i < got; i++) {
        const std::string &a = st[3 * i], &s = st[3 * i + 1], &q = st[3 * i + 2];
        r->name[i] = bsx_view{a.data(), (uint32_t)a.size()};
        r->seq[i] = bsx_view{s.data(), (uint32_t)s.size()};
        r->qual[i] = bsx_view{q.data(), (uint32_t)q.size()};
        put_seq(seqs, lens, stride, i, s.data(), (uint32_t)s.size());
    }
    r->n_slow += got;
    r->name.resize(got); r->seq.resize(got); r->qual.resize(got);
    return got;
}

extern "C" void bsx_reads_set_readset(bsx_reads *r, int readset) { if (r) r->readset = readset; }

extern "C" uint32_t bsx_reads_next(bsx_reads *r, uint32_t want, uint32_t stride, char *seqs, uint16_t *lens, int threads) {
    if (!r || want == 0) return 0;
    threads = bsx_host_threads(threads);
    r->name.resize(want); r->seq.resize(want); r->qual.resize(want);
    r->slow_store.clear();
    r->keep.clear();
    if (r->win) r->keep.push_back(r->win);
    uint32_t got = 0;
    if (r->kind == 3) return bam_batch(r, want, stride, seqs, lens);
    bool line_start = at_line_start(r);
    std::vector<size_t> slow_slots;
    std::string nm, sq, ql;
    while (got < want) {
        // streamed input: the records still wanted plus a margin no single record outgrows, or the rest of the stream
        if (!r->stream_eof) stream_ensure(r, (size_t)((double)(want - got) * std::max(r->rec_bytes, 64.0) * 1.25) + ((size_t)4 << 20));
        if (!r->force_slow && !line_start) line_start = resume_at_line_start(r);
        if (!r->force_slow && line_start) {
            bool bad = true;
            got += fast_batch(r, want - got, stride, seqs, lens, got, threads, &bad);
            if (got == want) break;
            if (!bad) continue;   // the scan window was short: cut some more
        }
        // the record at pos is irregular (or the input is exhausted): token reader, a few records at a time
        uint32_t k = 0;
        const uint32_t burst = r->force_slow ? want - got : std::min<uint32_t>(want - got, 16);
        for (; k < burst; k++) {
            if (!slow_record(r, nm, sq, ql)) break;
            r->slow_store.push_back(nm); r->slow_store.push_back(sq); r->slow_store.push_back(ql);
            slow_slots.push_back(got + k);
            r->n_slow++;
        }
        got += k;
        if (k < burst) break;   // end of input
        line_start = false;
    }
    for (size_t j = 0; j < slow_slots.size(); j++) {   // strings no longer move: take the views now
        const size_t i = slow_slots[j];
        const std::string &a = r->slow_store[3 * j], &b = r->slow_store[3 * j + 1], &c = r->slow_store[3 * j + 2];
        r->name[i] = bsx_view{a.data(), (uint32_t)a.size()};
        r->seq[i] = bsx_view{b.data(), (uint32_t)b.size()};
        r->qual[i] = bsx_view{c.data(), (uint32_t)c.size()};
        put_seq(seqs, lens, stride, i, b.data(), (uint32_t)b.size());
    }
    r->name.resize(got); r->seq.resize(got); r->qual.resize(got);
    return got;
}

extern "C" int bsx_reads_get(const bsx_reads *r, uint32_t i, const char **name, uint32_t *name_len,
                             const char **seq, uint32_t *seq_len, const char **qual, uint32_t *qual_len) {
    if (!r || i >= r->name.size()) { bsx_set_error("bsx_reads_get: index out of range"); return BSX_ERR_ARG; }
    if (name) *name = r->name[i].p; if (name_len) *name_len = r->name[i].n;
    if (seq) *seq = r->seq[i].p; if (seq_len) *seq_len = r->seq[i].n;
    if (qual) *qual = r->qual[i].p; if (qual_len) *qual_len = r->qual[i].n;
    return BSX_OK;
}

// Reference FASTA (RefSeq::LoadNextSeq, dbseq.cpp:18-54).  A '>' outside a header line opens a record;
// its name is the first blank-delimited token of that line; the sequence is every non-blank byte up
// to the next '>'.  One memchr pass finds the records, then the bodies are compacted in parallel.
// ---- packed read slots (include/bsmap_b200.h): what ConvertBinaySeq (align.cpp:90-162) derives from the text, on the host
extern "C" size_t bsx_packed_stride(uint32_t stride) { return (((size_t)stride + 3) / 4 + ((size_t)stride + 7) / 8 + 3) & ~(size_t)3; }

extern "C" int bsx_pack_reads(uint32_t n, const char *seqs, uint32_t stride, const uint16_t *lens, uint8_t *packed,
                              uint64_t *n_lowercase, int threads) {
    if ((n && (!seqs || !lens || !packed)) || stride < 16) { bsx_set_error("bsx_pack_reads: bad argument"); return BSX_ERR_ARG; }
    const size_t ps = bsx_packed_stride(stride), moff = ((size_t)stride + 3) / 4;
    // per ASCII byte: 2-bit code | valid << 2 | lower-case base << 3
    uint8_t lut[256];
    memset(lut, 0, sizeof lut);
    lut['A'] = 4; lut['C'] = 5; lut['G'] = 6; lut['T'] = 7; lut['a'] = 12; lut['c'] = 13; lut['g'] = 14; lut['t'] = 15;
    threads = bsx_host_threads(threads);
    std::vector<uint64_t> low((size_t)threads, 0);
    bsx_parallel(threads, n, [&](int tid, size_t b, size_t e) {
        uint64_t lc = 0;
        for (size_t r = b; r < e; r++) {
            const uint8_t *sq = (const uint8_t *)seqs + r * stride;
            uint8_t *o = packed + r * ps;
            memset(o, 0, ps);
            const uint32_t len = std::min<uint32_t>(lens[r], stride);
            for (uint32_t i = 0; i < len; i++) {
                const uint8_t c = lut[sq[i]];
                o[i >> 2] |= (uint8_t)((c & 3u) << (6 - 2 * (i & 3u)));
                o[moff + (i >> 3)] |= (uint8_t)(((c >> 2) & 1u) << (7 - (i & 7u)));
                lc += c >> 3;
            }
        }
        low[tid] = lc;
    });
    if (n_lowercase) { uint64_t t = 0; for (uint64_t x : low) t += x; *n_lowercase = t; }
    return BSX_OK;
}

int bsx_load_fasta(const char *path, std::vector<std::string> &names, std::vector<std::string> &seqs) {
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { bsx_set_error("fatal error: failed to open ref file %s", path); return BSX_ERR_IO; }
    struct stat st;
    const char *p = nullptr; size_t n = 0; bool mapped = false; std::vector<char> owned;
    if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
        n = (size_t)st.st_size;
        void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { close(fd); bsx_set_error("mmap failed: %s", path); return BSX_ERR_IO; }
        madvise(m, n, MADV_SEQUENTIAL); madvise(m, n, MADV_WILLNEED);
        p = (const char *)m; mapped = true;
    } else {
        char buf[1 << 16]; ssize_t g;
        while ((g = read(fd, buf, sizeof buf)) > 0) owned.insert(owned.end(), buf, buf + g);
        p = owned.data(); n = owned.size();
    }
    if (is_gzip(p, n)) {
        if (mapped) { munmap((void *)p, n); mapped = false; }
        if (gunzip_file(path, owned) != BSX_OK) { close(fd); bsx_set_error("failed to inflate gzip reference file: %s", path); return BSX_ERR_IO; }
        p = owned.data(); n = owned.size();
    }
    struct Body { size_t b, e; };
    std::vector<Body> body;
    names.clear(); seqs.clear();
    const bool timing = getenv("BSX_CLI_TIMING") != nullptr;
    auto clk = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = clk();
    const int threads = bsx_host_threads(0);
    // every '>' of the file, found on all threads; one that lies inside a header line does not open a record
    std::vector<std::vector<size_t>> gts((size_t)threads);
    bsx_parallel(threads, (size_t)threads, [&](int t, size_t, size_t) {
        const size_t b = n * (size_t)t / (size_t)threads, e = n * (size_t)(t + 1) / (size_t)threads;
        for (size_t a = b; a < e;) {
            const void *g = memchr(p + a, '>', e - a);
            if (!g) break;
            a = (size_t)((const char *)g - p);
            gts[t].push_back(a++);
        }
    });
    size_t q = 0;
    for (const std::vector<size_t> &v : gts)
        for (const size_t gt : v) {
            if (gt < q) continue;                            // inside the previous record's header line
            if (!body.empty()) body.back().e = gt;
            const void *nl = memchr(p + gt, '\n', n - gt);
            const size_t he = nl ? (size_t)((const char *)nl - p) : n;
            size_t a = gt + 1;
            while (a < he && (p[a] == ' ' || p[a] == '\t' || p[a] == '\r')) a++;
            size_t b = a;
            while (b < he && !(p[b] == ' ' || p[b] == '\t' || p[b] == '\r')) b++;
            names.emplace_back(p + a, b - a);
            q = he < n ? he + 1 : n;
            body.push_back(Body{q, n});
        }
    if (body.empty()) { if (mapped) munmap((void *)p, n); close(fd); bsx_set_error("no sequences in %s", path); return BSX_ERR_IO; }
    // pieces of at most 4 MB: count the bases, then copy them to their final offsets
    struct Piece { size_t seq, b, e, off, cnt; bool plain; };   // plain: the only blanks of the piece are line feeds
    std::vector<Piece> piece;
    const size_t PIECE = (size_t)4 << 20;
    for (size_t k = 0; k < body.size(); k++) {
        size_t b = body[k].b;
        do { piece.push_back(Piece{k, b, std::min(b + PIECE, body[k].e), 0, 0, false}); b += PIECE; } while (b < body[k].e);
    }
    const double t1 = clk();
    std::atomic<size_t> next_piece(0);
    bsx_parallel(threads, (size_t)threads, [&](int, size_t, size_t) {
        for (size_t i; (i = next_piece.fetch_add(1)) < piece.size();) {
            size_t c = 0, nl = 0;
            for (size_t a = piece[i].b; a < piece[i].e; a++) { c += !ws((unsigned char)p[a]); nl += p[a] == '\n'; }
            piece[i].cnt = c; piece[i].plain = c + nl == piece[i].e - piece[i].b;
        }
    });
    const double t2 = clk();
    seqs.resize(body.size());
    std::vector<size_t> total(body.size(), 0);
    for (Piece &pc : piece) { pc.off = total[pc.seq]; total[pc.seq] += pc.cnt; }
    {   // resize() zero-fills, i.e. first-touches every page: one sequence per thread at a time (it was 60 % of the load when serial)
        std::atomic<size_t> next_seq(0);
        bsx_parallel(threads, (size_t)threads, [&](int, size_t, size_t) { for (size_t k; (k = next_seq.fetch_add(1)) < seqs.size();) seqs[k].resize(total[k]); });
    }
    const double t3 = clk();
    next_piece = 0;
    bsx_parallel(threads, (size_t)threads, [&](int, size_t, size_t) {
        for (size_t i; (i = next_piece.fetch_add(1)) < piece.size();) {
            char *d = &seqs[piece[i].seq][0] + piece[i].off;
            if (piece[i].plain) {   // whole lines at a time
                for (size_t a = piece[i].b; a < piece[i].e;) {
                    const void *nl = memchr(p + a, '\n', piece[i].e - a);
                    const size_t e = nl ? (size_t)((const char *)nl - p) : piece[i].e;
                    memcpy(d, p + a, e - a); d += e - a; a = e + 1;
                }
                continue;
            }
            for (size_t a = piece[i].b; a < piece[i].e; a++) { const char c = p[a]; if (!ws((unsigned char)c)) *d++ = c; }
        }
    });
    if (timing) fprintf(stderr, "[bsx timing] FASTA: records %.3f s, count %.3f s, allocate %.3f s, copy %.3f s\n", t1 - t0, t2 - t1, t3 - t2, clk() - t3);
    if (mapped) munmap((void *)p, n);
    close(fd);
    return BSX_OK;
}

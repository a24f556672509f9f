// bsx_bam.cpp -- `-o out.bam`: SAM text -> coordinate-sorted BAM + BAI index, in process (output half).
//
// The reference writes SAM text into the .bam path and shells out to `sam2bam.sh` (main.cpp:466-473), i.e. the vendored
// samtools 0.1.7: `view -bS` (bam_import.c sam_read1), `sort` (bam_sort.c: stable merge sort on (tid, pos)), `index`
// (bam_index.c).  This file restates those three steps for BSMAP's own SAM output:
//   * records are encoded exactly as sam_read1 does (bin from reg2bin / calend, 4-bit bases through bam_nt16_table,
//     qualities - 33 or 0xff for '*', integer tags in the smallest type, mapped-without-CIGAR -> unmapped),
//   * sorted stably by ((uint32)tid << 32 | pos + 1) so that unmapped reads (tid -1) go last,
//   * written as BGZF with bgzf.c's blocking (64 KiB of input per block, header and records flowing through without a
//     flush, an empty block at the end), the blocks deflated on all host threads,
//   * indexed with bam_index_core's rules (bin chunks closed at every change of bin, 16 kb linear index that records
//     windows after the first one a read overlaps -- a 0.1.7 trait -- and chunks merged when they share a block).
// Decompressed, the BAM equals samtools' output byte for byte; the .bai holds the same bins, chunks and linear index
// as `samtools index` computes for the file (tests/test_bam_output_cpu.py checks both against oracle/_ref/samtools).
//
// Memory is bounded the way `samtools sort` bounds it: the SAM text is taken in chunks (BSX_BAM_CHUNK_MB of text, 1 GiB
// by default); a file that fits one chunk is sorted in memory, a larger one leaves one sorted run per chunk on disk
// (<out>.runNNNN.tmp) and the runs are merged -- ties go to the earlier chunk, i.e. input order, exactly bam_sort.c's
// heap rule.  Either way the sorted records stream through a window of 1 024 BGZF blocks that is deflated on all host
// threads and written as soon as it is full, and the index is built behind it; nothing holds the whole BAM.
#include <algorithm>
#include <memory>
#include <future>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <queue>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#include "bsx_internal.h"

namespace {

struct Nt16 { unsigned char t[256]; Nt16() { memset(t, 15, 256); const char *s = "=ACMGRSVTWYHKDBN"; for (int i = 0; s[i]; i++) { t[(unsigned char)s[i]] = (unsigned char)i; t[(unsigned char)tolower(s[i])] = (unsigned char)i; }
                                              t['0'] = 1; t['1'] = 2; t['2'] = 4; t['3'] = 8; } };   // bam_nt16_table (bam_import.c:24-41), incl. its colour-space digits
const Nt16 kNt16;

inline int reg2bin(uint32_t beg, uint32_t end) {   // bam.h:648-657
    --end;
    if (beg >> 14 == end >> 14) return 4681 + (beg >> 14);
    if (beg >> 17 == end >> 17) return 585 + (beg >> 17);
    if (beg >> 20 == end >> 20) return 73 + (beg >> 20);
    if (beg >> 23 == end >> 23) return 9 + (beg >> 23);
    if (beg >> 26 == end >> 26) return 1 + (beg >> 26);
    return 0;
}

struct Field { const char *p; size_t n; };
inline bool is_digit(char c) { return c >= '0' && c <= '9'; }
inline long to_long(const char *p, size_t n) { char b[32]; const size_t k = std::min<size_t>(n, 31); memcpy(b, p, k); b[k] = 0; return atol(b); }

template <class T> inline void put(std::string &o, T v) { o.append(reinterpret_cast<const char *>(&v), sizeof v); }

struct Rec { uint64_t key, off; uint32_t len; int32_t tid, pos, end; uint16_t bin; };   // end = calend (for the linear index)

// one SAM line -> one BAM record appended to `out` (block_size first); false for lines that are not alignments
bool encode(const char *b, const char *e, const std::unordered_map<std::string_view, int32_t> &tids, std::string &out, Rec &r) {
    if (e > b && e[-1] == '\r') e--;
    Field f[11]; int nf = 0;
    const char *s = b;
    while (nf < 11) { const char *t = (const char *)memchr(s, '\t', (size_t)(e - s)); if (!t) { f[nf++] = Field{s, (size_t)(e - s)}; s = e; break; } f[nf++] = Field{s, (size_t)(t - s)}; s = t + 1; }
    if (nf < 11) return false;
    const char *aux = s;                                             // rest of the line (may be empty)
    auto tid_of = [&](Field x) -> int32_t { auto it = tids.find(std::string_view(x.p, x.n)); return it == tids.end() ? -1 : it->second; };
    uint32_t flag;
    { char tmp[32]; const size_t k = std::min<size_t>(f[1].n, 31); memcpy(tmp, f[1].p, k); tmp[k] = 0; char *endp; long v = strtol(tmp, &endp, 0); flag = *endp ? 0u : (uint32_t)v; }
    const int32_t tid = tid_of(f[2]);
    const int32_t pos = f[3].n && is_digit(f[3].p[0]) ? (int32_t)to_long(f[3].p, f[3].n) - 1 : -1;
    const uint32_t mapq = f[4].n && is_digit(f[4].p[0]) ? (uint32_t)to_long(f[4].p, f[4].n) : 0;
    uint32_t cigar_small[8]; std::vector<uint32_t> cigar_big;     // BSMAP writes one operation; anything longer spills to the heap
    size_t n_cigar = 0;
    uint32_t endpos = (uint32_t)pos;
    int bin;
    if (f[5].n && f[5].p[0] != '*') {
        const char *c = f[5].p, *ce = f[5].p + f[5].n;
        while (c < ce) {
            long x = 0; while (c < ce && is_digit(*c)) x = x * 10 + (*c++ - '0');
            if (c >= ce) break;
            const char opc = (char)toupper((unsigned char)*c++);
            int op; switch (opc) { case 'M': case '=': case 'X': op = 0; break; case 'I': op = 1; break; case 'D': op = 2; break; case 'N': op = 3; break;
                                   case 'S': op = 4; break; case 'H': op = 5; break; case 'P': op = 6; break; default: return false; }
            { const uint32_t cv = (uint32_t)x << 4 | (uint32_t)op;
              if (n_cigar < 8) cigar_small[n_cigar] = cv; else { if (n_cigar == 8) cigar_big.assign(cigar_small, cigar_small + 8); cigar_big.push_back(cv); }
              n_cigar++; }
            if (op == 0 || op == 2 || op == 3) endpos += (uint32_t)x;
        }
        bin = reg2bin((uint32_t)pos, endpos);
    } else {
        flag |= 0x4;                                                 // mapped sequence without CIGAR (bam_import.c:297-300)
        bin = reg2bin((uint32_t)pos, (uint32_t)pos + 1);
    }
    const int32_t mtid = (f[6].n == 1 && f[6].p[0] == '=') ? tid : tid_of(f[6]);
    const int32_t mpos = f[7].n && is_digit(f[7].p[0]) ? (int32_t)to_long(f[7].p, f[7].n) - 1 : -1;
    const int32_t isize = f[8].n && (f[8].p[0] == '-' || is_digit(f[8].p[0])) ? (int32_t)to_long(f[8].p, f[8].n) : 0;
    const bool has_seq = !(f[9].n == 1 && f[9].p[0] == '*');
    const int32_t l_seq = has_seq ? (int32_t)f[9].n : 0;
    const size_t start = out.size();
    put<int32_t>(out, 0);                                            // block_size, patched below
    put<int32_t>(out, tid); put<int32_t>(out, pos);
    put<uint32_t>(out, (uint32_t)bin << 16 | (mapq & 0xff) << 8 | (uint32_t)((f[0].n + 1) & 0xff));
    put<uint32_t>(out, (flag & 0xffff) << 16 | (uint32_t)(n_cigar & 0xffff));
    put<int32_t>(out, l_seq); put<int32_t>(out, mtid); put<int32_t>(out, mpos); put<int32_t>(out, isize);
    out.append(f[0].p, f[0].n); out.push_back('\0');
    { const uint32_t *cg = n_cigar > 8 ? cigar_big.data() : cigar_small; out.append(reinterpret_cast<const char *>(cg), n_cigar * 4); }
    if (has_seq) {
        const size_t at = out.size(), nb = ((size_t)l_seq + 1) / 2;
        out.resize(at + nb + (size_t)l_seq);                         // packed bases, then qualities: written through a raw cursor
        unsigned char *w = reinterpret_cast<unsigned char *>(&out[at]);
        const unsigned char *sq = reinterpret_cast<const unsigned char *>(f[9].p);
        for (int32_t i = 0; i + 1 < l_seq; i += 2) *w++ = (unsigned char)(kNt16.t[sq[i]] << 4 | kNt16.t[sq[i + 1]]);
        if (l_seq & 1) *w++ = (unsigned char)(kNt16.t[sq[l_seq - 1]] << 4);
        if (f[10].n == 1 && f[10].p[0] == '*') memset(w, 0xff, (size_t)l_seq);
        else { const int32_t nq = (int32_t)std::min<size_t>(f[10].n, (size_t)l_seq);
               for (int32_t i = 0; i < nq; i++) w[i] = (unsigned char)(f[10].p[i] - 33);
               for (int32_t i = nq; i < l_seq; i++) w[i] = (unsigned char)('!' - 33); }
    }
    // auxiliary fields (bam_import.c:336-395)
    for (const char *a = aux; a < e;) {
        const char *t = (const char *)memchr(a, '\t', (size_t)(e - a)); if (!t) t = e;
        const size_t n = (size_t)(t - a);
        if (n >= 5 && a[2] == ':' && a[4] == ':') {
            out.push_back(a[0]); out.push_back(a[1]);
            const char type = a[3]; const char *v = a + 5; const size_t vn = n - 5;
            if (type == 'A' || type == 'a' || type == 'c' || type == 'C') { out.push_back('A'); out.push_back(vn ? v[0] : '\0'); }
            else if (type == 'I' || type == 'i') {
                char tmp[32]; const size_t k = std::min<size_t>(vn, 31); memcpy(tmp, v, k); tmp[k] = 0;
                const long long x = atoll(tmp);
                if (x < 0) {
                    if (x >= -127) { out.push_back('c'); put<int8_t>(out, (int8_t)x); }
                    else if (x >= -32767) { out.push_back('s'); put<int16_t>(out, (int16_t)x); }
                    else { out.push_back('i'); put<int32_t>(out, (int32_t)x); }
                } else {
                    if (x <= 255) { out.push_back('C'); put<uint8_t>(out, (uint8_t)x); }
                    else if (x <= 65535) { out.push_back('S'); put<uint16_t>(out, (uint16_t)x); }
                    else { out.push_back('I'); put<uint32_t>(out, (uint32_t)x); }
                }
            } else if (type == 'f') { char tmp[64]; const size_t k = std::min<size_t>(vn, 63); memcpy(tmp, v, k); tmp[k] = 0; out.push_back('f'); put<float>(out, (float)atof(tmp)); }
            else if (type == 'd') { char tmp[64]; const size_t k = std::min<size_t>(vn, 63); memcpy(tmp, v, k); tmp[k] = 0; out.push_back('d'); put<double>(out, atof(tmp)); }
            else if (type == 'Z' || type == 'H') { out.push_back(type); out.append(v, vn); out.push_back('\0'); }
            else { out.resize(out.size() - 2); }
        }
        a = t < e ? t + 1 : e;
    }
    const int32_t block_size = (int32_t)(out.size() - start - 4);
    memcpy(&out[start], &block_size, 4);
    r.key = (uint64_t)(uint32_t)tid << 32 | (uint32_t)(pos + 1);
    r.off = start; r.len = (uint32_t)(out.size() - start);
    r.tid = tid; r.pos = pos; r.end = (int32_t)endpos; r.bin = (uint16_t)bin;
    return true;
}

// one BGZF block (bgzf.c deflate_block): header, raw deflate at the default level, CRC32, ISIZE
bool bgzf_block(const unsigned char *in, size_t n, std::string &out) {
    unsigned char buf[65536];
    z_stream zs; memset(&zs, 0, sizeof zs);
    zs.next_in = const_cast<unsigned char *>(in); zs.avail_in = (uInt)n;
    zs.next_out = buf + 18; zs.avail_out = sizeof buf - 18 - 8;
    if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    const int st = deflate(&zs, Z_FINISH);
    deflateEnd(&zs);
    if (st != Z_STREAM_END) return false;                            // would not fit: the caller falls back to smaller input
    const size_t clen = zs.total_out + 26;
    static const unsigned char head[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
    memcpy(buf, head, 16);
    buf[16] = (unsigned char)((clen - 1) & 0xff); buf[17] = (unsigned char)((clen - 1) >> 8);
    const uint32_t crc = (uint32_t)crc32(crc32(0L, nullptr, 0), in, (uInt)n), isz = (uint32_t)n;
    memcpy(buf + clen - 8, &crc, 4); memcpy(buf + clen - 4, &isz, 4);
    out.assign(reinterpret_cast<char *>(buf), clen);
    return true;
}

// what the index needs to know about a record, and where its bytes are
struct Meta { uint64_t key; uint32_t len; int32_t tid, pos, end; uint16_t bin; uint16_t pad; };

// a sorted run on disk: Meta, then `len` bytes of BAM record, repeated
struct RunReader {
    FILE *f = nullptr; std::vector<char> rec; Meta m{}; bool live = false;
    bool open(const std::string &path) { f = fopen(path.c_str(), "rb"); if (f) setvbuf(f, nullptr, _IOFBF, 4 << 20); return f != nullptr; }
    bool advance() {
        live = fread(&m, sizeof m, 1, f) == 1;
        if (live) { rec.resize(m.len); live = fread(rec.data(), 1, m.len, f) == m.len; }
        return live;
    }
    void close() { if (f) fclose(f); f = nullptr; }
};

// BGZF writer + bam_index_core behind it.  Records arrive in sorted order; raw bytes collect in a window that is cut
// into 64 KiB blocks (bgzf.c's blocking: records flow through block boundaries), deflated in parallel and written.
static double bam_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
struct BamSink {
    static const size_t BLK = 65536;
    size_t WINDOW = 1024;                             // blocks per flush (BSX_BAM_WINDOW_BLOCKS: tests shrink it)
    int threads; FILE *fo = nullptr;
    std::string win;                                  // raw bytes not yet written, starting at raw offset `flushed`
    uint64_t flushed = 0, total = 0, file_pos = 0;    // raw bytes written / appended so far; compressed bytes written
    std::vector<uint64_t> blk_in_start, blk_file_start;   // per written block (sentinel at the end)
    bool fail = false;
    double t_deflate = 0, t_write = 0, t_index = 0;   // BSX_CLI_TIMING
    // index state (bam_index.c:133-190)
    struct Chunk { uint64_t u, v; };
    std::vector<std::map<uint32_t, std::vector<Chunk>>> bins;
    std::vector<std::vector<uint64_t>> lidx; std::vector<int32_t> lidx_n;
    uint32_t last_bin = 0xffffffffu, save_bin = 0xffffffffu; int32_t last_tid = -2 /* 0xffffffff in the source */, save_tid = -2;
    bool first = true, stopped = false, have_off = false;
    uint64_t save_off = 0, last_off = 0;
    struct Pending { Meta m; uint64_t at, next; };
    std::deque<Pending> pend;                         // records whose end lies in a block that is not written yet

    BamSink(int t, size_t nref) : threads(t), bins(nref), lidx(nref), lidx_n(nref, 0) {
        if (const char *e = getenv("BSX_BAM_WINDOW_BLOCKS")) { const long w = atol(e); if (w > 0) WINDOW = (size_t)w; }
        blk_in_start.push_back(0); blk_file_start.push_back(0); win.reserve((WINDOW + 2) * BLK);
    }
    void append(const char *p, size_t n) { win.append(p, n); total += n; }
    void add_record(const Meta &m, const char *bytes) {
        pend.push_back(Pending{m, total, total + m.len});
        append(bytes, m.len);
        if (win.size() >= WINDOW * BLK) flush(false);
    }
    // reader-side tell() of raw offset x: the offset of a record start as the *reader* sees it (start of the next block
    // when the previous one is exhausted), which is what bam_index_core records
    uint64_t rtell(uint64_t x) const {
        const size_t nblk = blk_in_start.size() - 1;
        const size_t k = (size_t)(std::upper_bound(blk_in_start.begin(), blk_in_start.end(), x) - blk_in_start.begin()) - 1;
        if (k >= nblk) return blk_file_start[nblk] << 16;
        return blk_file_start[k] << 16 | (x - blk_in_start[k]);
    }
    void index_ready(bool final) {
        while (!pend.empty() && !stopped) {
            const Pending &q = pend.front();
            if (!final && q.next >= flushed) break;                   // its end is in a block still in the window
            if (!have_off) { save_off = last_off = rtell(q.at); have_off = true; }
            const Meta &r = q.m;
            if (first || last_tid != r.tid) { last_tid = r.tid; last_bin = 0xffffffffu; first = false; }
            if (r.tid >= 0 && r.bin < 4681) {                          // insert_offset2 (bam_index.c:89-104)
                const int beg = r.pos >> 14, end = (int)(((uint32_t)r.end - 1) >> 14);
                std::vector<uint64_t> &lx = lidx[r.tid];
                if ((int)lx.size() < end + 1) lx.resize((size_t)end + 1, 0);
                for (int i = beg + 1; i <= end; i++) if (lx[i] == 0) lx[i] = last_off;
                lidx_n[r.tid] = end + 1;
            }
            if (r.bin != last_bin) {
                if (save_bin != 0xffffffffu) bins[save_tid][save_bin].push_back(Chunk{save_off, last_off});
                save_off = last_off; save_bin = last_bin = r.bin; save_tid = r.tid;
                if (save_tid < 0) { stopped = true; break; }
            }
            last_off = rtell(q.next);
            pend.pop_front();
        }
        if (stopped) pend.clear();
    }
    // Deflate and write every full block of the window (everything when final).  The blocks of a window are deflated on all
    // threads by a background task while the caller goes on filling the next window (the producer -- an in-order walk over
    // the sorted records or the merge of the runs -- is one thread; the deflate, at zlib's default level because the bytes
    // must equal samtools', is the bulk of the work).  The index catches up with a window when its task is joined.
    std::string job;                                  // the bytes the background task is working on
    std::future<size_t> task;                         // -> bytes of `job` it consumed (all of them unless a block did not fit)
    size_t deflate_job(bool final) {
        const size_t nblk = (job.size() + BLK - 1) / BLK;
        std::vector<std::string> comp(nblk); std::vector<char> okv(nblk, 1);
        const double td0 = bam_now();
        bsx_parallel(threads, nblk, [&](int, size_t b, size_t e) {
            for (size_t k = b; k < e; k++) okv[k] = bgzf_block((const unsigned char *)job.data() + k * BLK, std::min(BLK, job.size() - k * BLK), comp[k]) ? 1 : 0;
        });
        const double td1 = bam_now(); t_deflate += td1 - td0;
        size_t good = 0; while (good < nblk && okv[good]) good++;
        size_t used = 0;
        for (size_t k = 0; k < good; k++) {
            const size_t len = std::min(BLK, job.size() - k * BLK);
            if (fwrite(comp[k].data(), 1, comp[k].size(), fo) != comp[k].size()) fail = true;
            file_pos += comp[k].size(); used += len;
            blk_in_start.push_back(flushed + used); blk_file_start.push_back(file_pos);
        }
        if (good < nblk) {
            // a block that does not fit is redone the way bgzf.c does it (1 KiB less input at a time), which shifts
            // every later boundary -- never seen on BAM data.  What is left of the job goes back to the window.
            while (used < job.size() && !fail) {
                size_t len = std::min(BLK, job.size() - used); std::string c;
                if (!final && len < BLK) break;
                while (!bgzf_block((const unsigned char *)job.data() + used, len, c)) { if (len <= 1024) { fail = true; break; } len -= 1024; }
                if (fail) break;
                if (fwrite(c.data(), 1, c.size(), fo) != c.size()) fail = true;
                file_pos += c.size(); used += len;
                blk_in_start.push_back(flushed + used); blk_file_start.push_back(file_pos);
            }
        }
        flushed += used;
        t_write += bam_now() - td1;
        return used;
    }
    void join_task() {
        if (!task.valid()) return;
        const size_t used = task.get();
        if (used < job.size()) win.insert(0, job, used, std::string::npos);   // the rare shifted-boundary case
        job.clear();
        const double ti0 = bam_now();
        index_ready(false);
        t_index += bam_now() - ti0;
    }
    void flush(bool final) {
        join_task();
        const size_t take = final ? win.size() : win.size() / BLK * BLK;
        if (take && !fail) {
            job.assign(win, 0, take); win.erase(0, take);
            if (final) { deflate_job(true); if (!win.empty() && !fail) { job.swap(win); win.clear(); deflate_job(true); } job.clear(); }
            else task = std::async(std::launch::async, [this] { return deflate_job(false); });
        }
        if (final) { const double ti0 = bam_now(); index_ready(true); t_index += bam_now() - ti0; }
    }
};

}  // namespace

extern "C" int bsx_sam_to_sorted_bam(const char *sam_path, const char *bam_path, int threads) {
    if (!sam_path || !bam_path) { bsx_set_error("bsx_sam_to_sorted_bam: bad argument"); return BSX_ERR_ARG; }
    threads = bsx_host_threads(threads);
    const int fd = open(sam_path, O_RDONLY);
    struct stat st;
    if (fd < 0 || fstat(fd, &st) != 0) { bsx_set_error("cannot open %s", sam_path); return BSX_ERR_IO; }
    const size_t n = (size_t)st.st_size;
    const char *p = n ? (const char *)mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0) : "";
    if (n && p == MAP_FAILED) { close(fd); bsx_set_error("mmap failed: %s", sam_path); return BSX_ERR_IO; }
    // header: every leading '@' line verbatim (sam_header_read); @SQ lines define the reference dictionary
    size_t q = 0;
    std::vector<std::string> ref_names; std::vector<int32_t> ref_lens;
    while (q < n && p[q] == '@') {
        const void *nl = memchr(p + q, '\n', n - q);
        const size_t le = nl ? (size_t)((const char *)nl - p) : n;
        if (le - q >= 3 && memcmp(p + q, "@SQ", 3) == 0) {
            std::string name; long len = 0;
            for (size_t a = q; a < le;) {
                const void *t = memchr(p + a, '\t', le - a); const size_t te = t ? (size_t)((const char *)t - p) : le;
                if (te - a > 3 && memcmp(p + a, "SN:", 3) == 0) name.assign(p + a + 3, te - a - 3);
                if (te - a > 3 && memcmp(p + a, "LN:", 3) == 0) len = to_long(p + a + 3, te - a - 3);
                a = te + 1;
            }
            ref_names.push_back(name); ref_lens.push_back((int32_t)len);
        }
        q = nl ? le + 1 : n;
    }
    const size_t header_end = q;
    std::unordered_map<std::string_view, int32_t> tids;               // keys view ref_names (complete by now: no reallocation)
    for (size_t k = 0; k < ref_names.size(); k++) tids.emplace(std::string_view(ref_names[k]), (int32_t)k);   // first definition wins, like the hash in bam_aux.c

    const bool timing = getenv("BSX_CLI_TIMING") != nullptr;
    const double t_begin = bam_now();
    double t_encode = 0, t_sort = 0;
    // ---- sorted runs: the text in chunks cut at line starts
    size_t chunk_bytes = (size_t)1 << 30;
    if (const char *e = getenv("BSX_BAM_CHUNK_MB")) { const double mb = atof(e); if (mb > 0) chunk_bytes = (size_t)(mb * 1048576.0); }
    if (chunk_bytes < 4096) chunk_bytes = 4096;
    std::vector<std::string> enc((size_t)threads); std::vector<std::vector<Rec>> recs((size_t)threads);
    struct Ord { uint64_t key; uint32_t t, i; };
    std::vector<Ord> ord;
    std::vector<std::string> run_paths;
    bool io_fail = false;
    auto cleanup = [&]() { for (const std::string &r : run_paths) remove(r.c_str()); if (n) munmap((void *)p, n); close(fd); };
    size_t cb = header_end;
    bool in_memory = false;
    while (cb < n || (cb == header_end && !in_memory && run_paths.empty())) {
        size_t ce = std::min(n, cb + chunk_bytes);
        if (ce < n) { const void *x = memchr(p + ce - 1, '\n', n - (ce - 1)); ce = x ? (size_t)((const char *)x - p) + 1 : n; }
        // records: byte ranges cut at line starts, one per thread
        const int tn = (ce - cb) < ((size_t)1 << 20) ? 1 : threads;
        const double t_chunk0 = bam_now();
        for (auto &v : enc) v.clear();
        for (auto &v : recs) v.clear();
        bsx_parallel(tn, (size_t)tn, [&](int t, size_t, size_t) {
            size_t b = cb + (ce - cb) * (size_t)t / tn, e = cb + (ce - cb) * (size_t)(t + 1) / tn;
            if (t > 0) { const void *x = memchr(p + b - 1, '\n', ce - (b - 1)); b = x ? (size_t)((const char *)x - p) + 1 : ce; }
            if (t + 1 < tn) { const void *x = memchr(p + e - 1, '\n', ce - (e - 1)); e = x ? (size_t)((const char *)x - p) + 1 : ce; }
            enc[t].reserve((e > b ? e - b : 0) + 1024);
            while (b < e) {
                const void *x = memchr(p + b, '\n', ce - b);
                const size_t le = x ? (size_t)((const char *)x - p) : ce;
                Rec r;
                if (le > b && encode(p + b, p + le, tids, enc[t], r)) recs[t].push_back(r);
                b = le + 1;
            }
        });
        t_encode += bam_now() - t_chunk0;
        const double t_sort0 = bam_now();
        // stable sort by (tid, pos + 1).  (key, thread, index) is a total order that equals input order on equal keys, so
        // the threads sort slices with a plain sort and the slices are merged pairwise, level by level, in parallel.
        auto before = [](const Ord &a, const Ord &b) { return a.key != b.key ? a.key < b.key : (a.t != b.t ? a.t < b.t : a.i < b.i); };
        ord.clear();
        { size_t tot = 0; for (auto &v : recs) tot += v.size(); ord.reserve(tot); }
        for (int t = 0; t < threads; t++) for (uint32_t i = 0; i < recs[t].size(); i++) ord.push_back(Ord{recs[t][i].key, (uint32_t)t, i});
        {
            const size_t N = ord.size();
            const int S = N < ((size_t)1 << 16) ? 1 : threads;
            std::vector<size_t> cut((size_t)S + 1);
            for (int k = 0; k <= S; k++) cut[k] = N * (size_t)k / S;
            bsx_parallel(S, (size_t)S, [&](int k, size_t, size_t) { std::sort(ord.begin() + cut[k], ord.begin() + cut[k + 1], before); });
            for (int w = 1; w < S; w *= 2) {
                const int pairs = (S + 2 * w - 1) / (2 * w);
                bsx_parallel(pairs, (size_t)pairs, [&](int q, size_t, size_t) {
                    const int lo = q * 2 * w, mid = std::min(lo + w, S), hi = std::min(lo + 2 * w, S);
                    if (mid < hi) std::inplace_merge(ord.begin() + cut[lo], ord.begin() + cut[mid], ord.begin() + cut[hi], before);
                });
            }
        }
        t_sort += bam_now() - t_sort0;
        if (cb == header_end && ce >= n) { in_memory = true; break; }    // the whole file is one chunk: no run files
        char name[32]; snprintf(name, sizeof name, ".run%04zu.tmp", run_paths.size());
        run_paths.push_back(std::string(bam_path) + name);
        FILE *fr = fopen(run_paths.back().c_str(), "wb");
        if (!fr) { cleanup(); bsx_set_error("cannot write %s", run_paths.back().c_str()); return BSX_ERR_IO; }
        {
            // the run = (Meta, record bytes) in sorted order.  Walking the sorted order touches the encoded records at random:
            // all threads gather them into slabs of ~256 MB, each written with one call.
            const size_t N = ord.size();
            std::vector<uint32_t> sz(N);
            bsx_parallel(threads, N, [&](int, size_t b, size_t e) { for (size_t j = b; j < e; j++) sz[j] = (uint32_t)sizeof(Meta) + recs[ord[j].t][ord[j].i].len; });
            const size_t SLAB = (size_t)256 << 20;
            std::unique_ptr<char[]> slab(new char[SLAB + (1 << 20)]);
            std::vector<size_t> at;
            for (size_t j0 = 0; j0 < N && !io_fail;) {
                at.clear();
                size_t bytes = 0, j1 = j0;
                while (j1 < N && bytes + sz[j1] <= SLAB + (1 << 20) && (bytes < SLAB || j1 == j0)) { at.push_back(bytes); bytes += sz[j1++]; }
                bsx_parallel(threads, j1 - j0, [&](int, size_t b, size_t e) {
                    for (size_t j = j0 + b; j < j0 + e; j++) {
                        const Rec &r = recs[ord[j].t][ord[j].i];
                        const Meta m{r.key, r.len, r.tid, r.pos, r.end, r.bin, 0};
                        char *w = slab.get() + at[j - j0];
                        memcpy(w, &m, sizeof m); memcpy(w + sizeof m, enc[ord[j].t].data() + r.off, r.len);
                    }
                });
                if (fwrite(slab.get(), 1, bytes, fr) != bytes) io_fail = true;
                j0 = j1;
            }
        }
        if (fclose(fr) != 0) io_fail = true;
        if (io_fail) { cleanup(); bsx_set_error("write to %s failed", run_paths.back().c_str()); return BSX_ERR_IO; }
        cb = ce;
    }
    if (!in_memory) {                                                     // the runs are on disk: give the chunk's memory back
        for (auto &v : enc) std::string().swap(v);
        for (auto &v : recs) std::vector<Rec>().swap(v);
        std::vector<Ord>().swap(ord);
    }

    const double t_phase1 = bam_now();
    // ---- the BAM: header, then the sorted records through the BGZF window
    BamSink sink(threads, ref_names.size());
    sink.fo = fopen(bam_path, "wb");
    if (!sink.fo) { cleanup(); bsx_set_error("cannot write %s", bam_path); return BSX_ERR_IO; }
    {
        std::string hd;
        hd.append("BAM\1", 4);
        put<int32_t>(hd, (int32_t)header_end); hd.append(p, header_end);
        put<int32_t>(hd, (int32_t)ref_names.size());
        for (size_t k = 0; k < ref_names.size(); k++) { put<int32_t>(hd, (int32_t)ref_names[k].size() + 1); hd.append(ref_names[k]); hd.push_back('\0'); put<int32_t>(hd, ref_lens[k]); }
        sink.append(hd.data(), hd.size());
    }
    if (in_memory) {
        for (const Ord &o : ord) {
            const Rec &r = recs[o.t][o.i];
            sink.add_record(Meta{r.key, r.len, r.tid, r.pos, r.end, r.bin, 0}, enc[o.t].data() + r.off);
        }
    } else {
        std::vector<RunReader> rd(run_paths.size());
        typedef std::pair<uint64_t, size_t> HeapKey;                      // (key, run): ties go to the earlier chunk = input order
        std::priority_queue<HeapKey, std::vector<HeapKey>, std::greater<HeapKey>> heap;
        for (size_t k = 0; k < rd.size(); k++) {
            if (!rd[k].open(run_paths[k])) { io_fail = true; break; }
            if (rd[k].advance()) heap.push(HeapKey(rd[k].m.key, k));
        }
        while (!io_fail && !heap.empty()) {
            const size_t k = heap.top().second; heap.pop();
            sink.add_record(rd[k].m, rd[k].rec.data());
            if (rd[k].advance()) heap.push(HeapKey(rd[k].m.key, k));
        }
        for (RunReader &r : rd) r.close();
    }
    sink.flush(true);
    if (timing) fprintf(stderr, "[bsx timing] BAM: chunks %.3f s (encode %.3f s, sort %.3f s, the rest sorted runs), merge + BGZF %.3f s (deflate %.3f s in the background, write %.3f s, index %.3f s), total %.3f s (%d threads)\n",
                        t_phase1 - t_begin, t_encode, t_sort, bam_now() - t_phase1, sink.t_deflate, sink.t_write, sink.t_index, bam_now() - t_begin, threads);
    std::string eof_block; bgzf_block((const unsigned char *)"", 0, eof_block);
    if (fwrite(eof_block.data(), 1, eof_block.size(), sink.fo) != eof_block.size()) sink.fail = true;
    if (fclose(sink.fo) != 0) sink.fail = true;
    cleanup();
    if (io_fail || sink.fail) { bsx_set_error("write to %s failed", bam_path); return BSX_ERR_IO; }
    // the index loop ran into the end of the file: the reader has swallowed the empty EOF block too (bgzf_read), so its
    // tell() is the file size
    if (sink.save_tid >= 0 && sink.save_bin != 0xffffffffu && !sink.stopped)
        sink.bins[sink.save_tid][sink.save_bin].push_back(BamSink::Chunk{sink.save_off, (sink.file_pos + eof_block.size()) << 16});

    const size_t nref = ref_names.size();
    std::vector<std::map<uint32_t, std::vector<BamSink::Chunk>>> &bins = sink.bins;
    std::vector<std::vector<uint64_t>> &lidx = sink.lidx; std::vector<int32_t> &lidx_n = sink.lidx_n;
    typedef BamSink::Chunk Chunk;
    const std::string bai_path = std::string(bam_path) + ".bai";
    FILE *fi = fopen(bai_path.c_str(), "wb");
    if (!fi) { bsx_set_error("cannot write %s", bai_path.c_str()); return BSX_ERR_IO; }
    fwrite("BAI\1", 1, 4, fi);
    { const int32_t x = (int32_t)nref; fwrite(&x, 4, 1, fi); }
    for (size_t t = 0; t < nref; t++) {
        const int32_t nb = (int32_t)bins[t].size(); fwrite(&nb, 4, 1, fi);
        for (auto &kv : bins[t]) {
            std::vector<Chunk> &l = kv.second;                        // merge_chunks, BAM_VIRTUAL_OFFSET16: same block -> one chunk
            size_t m = 0;
            for (size_t i = 1; i < l.size(); i++) { if (l[m].v >> 16 == l[i].u >> 16) l[m].v = l[i].v; else l[++m] = l[i]; }
            l.resize(l.empty() ? 0 : m + 1);
            const uint32_t bin = kv.first; const int32_t nc = (int32_t)l.size();
            fwrite(&bin, 4, 1, fi); fwrite(&nc, 4, 1, fi);
            for (const Chunk &c : l) { fwrite(&c.u, 8, 1, fi); fwrite(&c.v, 8, 1, fi); }
        }
        const int32_t nl = lidx_n[t]; fwrite(&nl, 4, 1, fi);
        for (int32_t i = 0; i < nl; i++) { const uint64_t v = i < (int32_t)lidx[t].size() ? lidx[t][i] : 0; fwrite(&v, 8, 1, fi); }
    }
    if (fclose(fi) != 0) { bsx_set_error("write to %s failed", bai_path.c_str()); return BSX_ERR_IO; }
    return BSX_OK;
}

// bsx_format.cpp -- host text layer: s_OutHit (align.cpp:631-765), s_OutHitPair / s_OutHitUnpair
// (pairs.cpp:288-498), FixPairReadName (pairs.cpp:535-555), the SAM header (main.cpp:405-413).
// Pure host code: it turns the device's fixed-size records into the bytes the reference writes.
#include <algorithm>
#include <cctype>
#include <unistd.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "bsx_internal.h"

namespace {

const char kNt[] = "ACGTacgt";   // useful_nt after SetAlign('T','C') (param.cpp:220-221)

// text sink: a caller-sized buffer (the C ABI's two-call protocol: bytes beyond cap are counted, not written) or a growing
// string (chunked emit).  The string is grown ahead of the cursor -- ensure() once per record with an upper bound of the
// record's size -- so that the twenty-odd pieces of a record are plain stores, not twenty capacity checks.
struct Out {
    char *p; size_t cap, n; std::string *dyn;
    void ensure(size_t bound) {
        if (!dyn || n + bound <= cap) return;
        dyn->resize(std::max(cap * 2, n + bound + (size_t)65536));
        p = &(*dyn)[0]; cap = dyn->size();
    }
    void finish() { if (dyn) { dyn->resize(n); cap = n; } }
    void put(const char *s, size_t len) {
        if (p && n + len <= cap) memcpy(p + n, s, len);
        n += len;
    }
    void put(const std::string &s) { put(s.data(), s.size()); }
    template <size_t N> void puts(const char (&s)[N]) { put(s, N - 1); }     // string literals: length known at compile time
    void putc(char c) { if (p && n < cap) p[n] = c; n++; }
    void putu(unsigned long long v) { char b[24]; char *e = b + 24, *q = e; do { *--q = (char)('0' + v % 10); v /= 10; } while (v); put(q, (size_t)(e - q)); }
    void puti(long long v) { if (v < 0) { putc('-'); putu(0ull - (unsigned long long)v); } else putu((unsigned long long)v); }
};
// upper bound of what one read can print (name, bases, qualities, reference name, XR / BSP reference window, BSP counts, fixed text)
inline size_t record_bound(size_t max_ref_name, const bsx_view &name) { return (size_t)name.n + 4u * (BSX_MAX_READLEN + 16u) + max_ref_name + 384u; }
inline size_t longest_name(const bsx_index *ix) { size_t m = 0; for (const std::string &n : ix->names) m = std::max(m, n.size()); return m; }

char comp(char c) {   // rev_char[] (param.cpp:166-177)
    switch (c) {
        case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
        case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
        default: return 'N';
    }
}
// bases / qualities of one read, clipped to the mapped length (<= BSX_MAX_READLEN)
struct Str {
    char c[BSX_MAX_READLEN + 16]; int n;
    size_t size() const { return (size_t)n; }
    void assign(const char *s, size_t l) { n = (int)std::min<size_t>(l, sizeof c); memcpy(c, s, (size_t)n); }
    void erase(size_t at) { if ((int)at < n) n = (int)at; }
};
void revcomp(Str &s) { for (int i = 0, j = s.n - 1; i <= j; i++, j--) { const char a = comp(s.c[i]), b = comp(s.c[j]); s.c[i] = b; s.c[j] = a; } }
void reverse(Str &s) { for (int i = 0, j = s.n - 1; i < j; i++, j--) std::swap(s.c[i], s.c[j]); }

const std::vector<uint32_t> &watson(const bsx_index *ix) {
    bsx_index *m = const_cast<bsx_index *>(ix);
    if (m->h_refcat.empty()) {
        m->h_refcat.resize(ix->n_words);
        if (bsx_index_download(ix, 0, m->h_refcat.data(), ix->n_words * 4) != BSX_OK) m->h_refcat.assign(ix->n_words, 0);
    }
    return m->h_refcat;
}

// reference bases around a hit: 2 upstream (lower case) + len + 2 downstream (lower case)
std::string mapseq(const bsx_index *ix, uint32_t chr, uint32_t loc, int len) {
    const std::vector<uint32_t> &ref = watson(ix);
    const uint32_t *m = ref.data() + (ix->anchor[chr >> 1] >> 4);
    std::string s;
    for (uint32_t ii = 2; ii > 0; ii--) {
        if (loc < ii) continue;                      // App. B Q12
        uint32_t q = loc - ii;
        s.push_back((char)(kNt[(m[q >> 4] >> (30 - 2 * (q & 15))) & 3] + 32));
    }
    for (int ii = 0; ii < len + 2; ii++) {
        uint32_t q = loc + (uint32_t)ii;
        s.push_back(kNt[(m[q >> 4] >> (30 - 2 * (q & 15))) & 3]);
    }
    s[s.size() - 1] += 32; s[s.size() - 2] += 32;
    return s;
}

// RefSeq::CCGG_seglen (dbseq.cpp:541-567)
void seglen(const bsx_index *ix, uint32_t chr, uint32_t pos, int readlen, uint32_t *first, int *second) {
    const std::vector<uint32_t> &st = ix->sites[chr >> 1];
    const int n = (int)st.size();
    if (n == 0) { *first = 0; *second = 0; return; }
    int left = 0, right = n - 1;
    while (left < right - 1) {
        int mid = (left + right) / 2;
        if (st[mid] == pos) { left = mid; right = mid + 1; break; }
        else if (st[mid] < pos) left = mid; else right = mid;
    }
    const uint32_t add = (uint32_t)(strlen(ix->par.digest_site) - 2 * ix->par.digest_pos);
    uint32_t seg_end;
    for (;;) {
        int rr = right < n ? right : n - 1;          // App. B Q20: clamp
        seg_end = st[rr] + add;
        if (seg_end < pos + (uint32_t)readlen && right < n) right++; else break;
    }
    *first = st[left] + 1; *second = (int)(seg_end - st[left]);
}

struct Read { bsx_view name; Str seq, qual; int raw; };

int clip(const bsx_params *p, size_t l) { int v = (int)l; if (v > p->max_readlen) v = p->max_readlen; if (v > BSX_MAX_READLEN) v = BSX_MAX_READLEN; return v; }
int rmsn(const bsx_params *p, int len, int raw) { return raw > 0 ? (int)((size_t)(p->max_snp_num + 1) * (size_t)(len - 1) / (size_t)raw) : 0; }

inline void make_read(Read &r, const bsx_params *p, bsx_view name, bsx_view seq, bsx_view qual, int len) {
    r.name = name;
    r.raw = clip(p, seq.n);
    r.seq.assign(seq.p, std::min<size_t>(seq.n, (size_t)len));
    r.qual.assign(qual.p, std::min<size_t>(qual.n, (size_t)len));
}

// s_OutHit.  n: -1 filtered (QC), 0 no hit (NM), >0 hits.
void out_hit(const bsx_index *ix, const bsx_params *p, Out &o, Read &rd, int readset, int chain, int n, int nsnps,
             uint32_t chr, uint32_t loc, int insert_size, const uint16_t *counts, int max_snp, uint32_t *n_aligned) {
    const int len = (int)rd.seq.size();
    const bool rev = (chain ^ (int)(chr & 1)) != 0;
    if (p->out_sam) {
        int flag = 0x40 * readset;
        if (n <= 0 || (n > 1 && p->report_repeat_hits == 0)) {
            if (!p->out_unmap) return;
            flag |= (n < 0) ? 0x204 : (n == 0 ? 0x4 : 0x104);
            o.put(rd.name.p, rd.name.n); o.putc('\t'); o.puti(flag); o.puts("\t*\t0\t0\t*\t*\t0\t0\t"); o.put(rd.seq.c, rd.seq.size()); o.putc('\t'); o.put(rd.qual.c, rd.qual.size()); o.putc('\n');
            return;
        }
        ++*n_aligned;
        if (n > 1) flag |= 0x100;
        if (rev) { flag |= 0x10; revcomp(rd.seq); reverse(rd.qual); }
        o.put(rd.name.p, rd.name.n); o.putc('\t'); o.puti(flag); o.putc('\t'); o.put(ix->names[chr >> 1]); o.putc('\t'); o.putu(loc + 1u);
        o.puts("\t255\t"); o.puti(len); o.puts("M\t*\t0\t0\t"); o.put(rd.seq.c, rd.seq.size()); o.putc('\t'); o.put(rd.qual.c, rd.qual.size()); o.puts("\tNM:i:"); o.puti(nsnps);
        if (p->out_ref) { o.puts("\tXR:Z:"); o.put(mapseq(ix, chr, loc, len)); }
        if (p->rrbs) { uint32_t f; int sl; seglen(ix, chr, loc, len, &f, &sl); o.puts("\tZP:i:"); o.puti((int)f); o.puts("\tZL:i:"); o.puti(sl); }
        o.puts("\tZS:Z:"); o.putc("+-"[chr & 1]); o.putc("+-"[chain]); o.putc('\n');
        return;
    }
    // BSP
    if (!p->out_unmap && (n <= 0 || (n > 1 && p->report_repeat_hits == 0))) return;
    o.put(rd.name.p, rd.name.n); o.putc('\t');
    if (rev && n) { revcomp(rd.seq); reverse(rd.qual); }
    o.put(rd.seq.c, rd.seq.size()); o.putc('\t'); o.put(rd.qual.c, rd.qual.size()); o.putc('\t');
    o.put(n < 0 ? "QC" : n == 0 ? "NM" : n == 1 ? "UM" : n >= p->max_num_hits ? "OF" : "MA", 2);
    if ((n > 0 && p->report_repeat_hits == 1) || (n == 1 && p->report_repeat_hits == 0)) {
        ++*n_aligned;
        o.putc('\t'); o.put(ix->names[chr >> 1]); o.putc('\t'); o.putu(loc + 1u); o.putc('\t'); o.putc("+-"[chr & 1]); o.putc("+-"[chain]);
        o.putc('\t'); o.puti(insert_size); o.putc('\t'); o.put(mapseq(ix, chr, loc, len)); o.putc('\t'); o.puti(nsnps); o.putc('\t');
        for (int ii = 0; ii < max_snp; ii++) { o.puti(counts ? counts[ii] : 0); o.putc(':'); }
        o.puti(counts ? counts[max_snp] : 0);
    }
    o.putc('\n');
}

// s_OutHitUnpair, SAM branch
void out_unpair_sam(const bsx_index *ix, const bsx_params *p, Out &o, Read &rd, int readset, int chain_a, int chain_b,
                    int ma, int na, uint32_t a_chr, uint32_t a_loc, int mb, uint32_t b_chr, uint32_t b_loc, uint32_t *n_al) {
    int flag = 1 | (0x40 * readset);
    const bool mate_un = (mb <= 0 || (mb > 1 && p->report_repeat_hits == 0));
    if (ma <= 0 || (ma > 1 && p->report_repeat_hits == 0)) {
        if (!p->out_unmap) return;
        if (ma < 0) flag |= 0x204;
        if (ma == 0) flag |= 0x004;
        if (ma > 1) flag |= 0x104;
        if (mate_un) {
            flag |= 0x008;
            o.put(rd.name.p, rd.name.n); o.putc('\t'); o.puti(flag); o.puts("\t*\t0\t0\t*\t*\t0\t0\t");
        } else {
            if (chain_b ^ (int)(b_chr & 1)) flag |= 0x020;
            o.put(rd.name.p, rd.name.n); o.putc('\t'); o.puti(flag); o.puts("\t*\t0\t0\t*\t"); o.put(ix->names[b_chr >> 1]); o.putc('\t'); o.putu(b_loc + 1u); o.puts("\t0\t");
        }
        o.put(rd.seq.c, rd.seq.size()); o.putc('\t'); o.put(rd.qual.c, rd.qual.size()); o.putc('\n');
        return;
    }
    ++*n_al;
    if (ma > 1) flag |= 0x100;
    if (chain_a ^ (int)(a_chr & 1)) { flag |= 0x010; revcomp(rd.seq); reverse(rd.qual); }
    const int len = (int)rd.seq.size();
    if (mate_un) flag |= 0x008; else if (chain_b ^ (int)(b_chr & 1)) flag |= 0x020;
    o.put(rd.name.p, rd.name.n); o.putc('\t'); o.puti(flag); o.putc('\t'); o.put(ix->names[a_chr >> 1]); o.putc('\t'); o.putu(a_loc + 1u);
    o.puts("\t255\t"); o.puti(len); o.puts("M\t");
    if (mate_un) o.puts("*\t0\t0\t");
    else { o.put(ix->names[b_chr >> 1]); o.putc('\t'); o.putu(b_loc + 1u); o.puts("\t0\t"); }
    o.put(rd.seq.c, rd.seq.size()); o.putc('\t'); o.put(rd.qual.c, rd.qual.size()); o.puts("\tNM:i:"); o.puti(na);
    if (p->out_ref) { o.puts("\tXR:Z:"); o.put(mapseq(ix, a_chr, a_loc, len)); }
    if (p->rrbs) { uint32_t f; int sl; seglen(ix, a_chr, a_loc, len, &f, &sl); o.puts("\tZP:i:"); o.puti((int)f); o.puts("\tZL:i:"); o.puti(sl); }
    o.puts("\tZS:Z:"); o.putc("+-"[a_chr & 1]); o.putc("+-"[chain_a]); o.putc('\n');
}

}  // namespace

extern "C" size_t bsx_format_header(const bsx_index *ix, char *out, size_t cap) {
    Out o{out, cap, 0, nullptr};
    o.puts("@HD\tVN:1.0\n");
    for (uint32_t k = 0; k < ix->n_seq; k++) { o.puts("@SQ\tSN:"); o.put(ix->names[k]); o.puts("\tLN:"); o.putu(ix->size[k]); o.putc('\n'); }
    o.puts("@PG\tID:BSMAP_2.6\n");
    return o.n;
}

namespace {

// reads [b, e) of a single-end batch
void se_range(const bsx_index *ix, const bsx_params *p, uint32_t b, uint32_t e, const bsx_view *names, const bsx_view *seqs,
              const bsx_view *quals, int readset, const bsx_rec *recs, const uint16_t *counts, Out &o, uint32_t *n_aligned, size_t max_ref_name) {
    uint32_t na = 0;
    Read rd;
    for (uint32_t t = b; t < e; t++) {
        const bsx_rec &rc = recs[t];
        o.ensure(record_bound(max_ref_name, names[t]));
        make_read(rd, p, names[t], seqs[t], quals[t], rc.len);
        if (rc.status == 1) {   // Do_Batch (align.cpp:598-600): filtered reads are printed only when -r != 0
            if (p->report_repeat_hits) out_hit(ix, p, o, rd, readset, 0, -1, 0, 0, 0, 0, nullptr, 0, &na);
            continue;
        }
        out_hit(ix, p, o, rd, readset, rc.chain, (int)rc.nhits, rc.nm, rc.chr, rc.loc, 0, counts ? counts + (size_t)t * 16 : nullptr,
                rmsn(p, rc.len, rd.raw), &na);
    }
    o.finish();
    *n_aligned = na;
}

// pairs [b, e) of a paired-end batch; st = pairs, single a, single b
void pe_range(const bsx_index *ix, const bsx_params *p, uint32_t b, uint32_t e,
              const bsx_view *names_a, const bsx_view *seqs_a, const bsx_view *quals_a,
              const bsx_view *names_b, const bsx_view *seqs_b, const bsx_view *quals_b,
              const bsx_pair_rec *pr, const bsx_rec *ra, const bsx_rec *rb, const uint16_t *counts_a, const uint16_t *counts_b,
              Out &o, Out &ou, uint32_t *st, size_t max_ref_name) {
    uint32_t n_pairs = 0, n_a = 0, n_b = 0, dummy = 0;
    Read A, B;
    for (uint32_t t = b; t < e; t++) {
        { const size_t bound = record_bound(max_ref_name, names_a[t]) + record_bound(max_ref_name, names_b[t]); o.ensure(bound); ou.ensure(bound); }
        make_read(A, p, names_a[t], seqs_a[t], quals_a[t], ra[t].len);
        make_read(B, p, names_b[t], seqs_b[t], quals_b[t], rb[t].len);
        if (p->out_sam && !(A.name.n == B.name.n && memcmp(A.name.p, B.name.p, A.name.n) == 0)) {
            // FixPairReadName: cut both names after the last digit of their common prefix
            size_t i = 0, i0 = std::min(A.name.n, B.name.n); long d = -1;
            for (; i < i0; i++) { if (A.name.p[i] != B.name.p[i]) break; else if (isdigit((unsigned char)A.name.p[i])) d = (long)i; }
            if (i > 0) { if (d < 0) d = (long)i - 1; if (A.name.n > (size_t)d + 1) A.name.n = (uint32_t)d + 1; if (B.name.n > (size_t)d + 1) B.name.n = (uint32_t)d + 1; }
        }
        const uint16_t *ca = counts_a ? counts_a + (size_t)t * 16 : nullptr, *cb = counts_b ? counts_b + (size_t)t * 16 : nullptr;
        const bsx_pair_rec &pp = pr[t];
        if (pp.paired) {
            n_pairs++;
            uint32_t a_loc = pp.a_loc, b_loc = pp.b_loc;
            const int ins = pp.insert, chain = pp.chain;
            const int lena0 = (int)A.seq.size(), lenb0 = (int)B.seq.size();
            // fragment shorter than the read: drop the adapter part (pairs.cpp:296-306)
            if (ins < lena0) { if (chain ^ (int)(pp.a_chr & 1)) a_loc += (uint32_t)(lena0 - ins); A.seq.erase((size_t)ins); if (A.qual.size() > (size_t)ins) A.qual.erase((size_t)ins); }
            if (ins < lenb0) { if ((!chain) ^ (int)(pp.b_chr & 1)) b_loc += (uint32_t)(lenb0 - ins); B.seq.erase((size_t)ins); if (B.qual.size() > (size_t)ins) B.qual.erase((size_t)ins); }
            const int np = (int)pp.npairs;
            if (p->out_sam) {
                for (int mate = 0; mate < 2; mate++) {
                    Read &rd = mate ? B : A;
                    const uint32_t chr = mate ? pp.b_chr : pp.a_chr, loc = mate ? b_loc : a_loc, mloc = mate ? a_loc : b_loc;
                    const int ch = mate ? !chain : chain, len = (int)rd.seq.size();
                    int flag = 0x3, tlen; uint32_t seg_start;
                    if (np > 1) flag |= 0x100;
                    if (ch ^ (int)(chr & 1)) { flag |= 0x10; seg_start = mloc + 1; tlen = -ins; revcomp(rd.seq); reverse(rd.qual); }
                    else { flag |= 0x20; seg_start = loc + 1; tlen = ins; }
                    flag |= 0x40 * (mate ? 2 : 1);
                    o.put(rd.name.p, rd.name.n); o.putc('\t'); o.puti(flag); o.putc('\t'); o.put(ix->names[chr >> 1]); o.putc('\t'); o.putu(loc + 1u);
                    o.puts("\t255\t"); o.puti(len); o.puts("M\t=\t"); o.putu(mloc + 1u); o.putc('\t'); o.puti(tlen); o.putc('\t');
                    o.put(rd.seq.c, rd.seq.size()); o.putc('\t'); o.put(rd.qual.c, rd.qual.size()); o.puts("\tNM:i:"); o.puti(mate ? pp.nb : pp.na);
                    if (p->out_ref) { o.puts("\tXR:Z:"); o.put(mapseq(ix, chr, loc, len)); }
                    if (p->rrbs) { o.puts("\tZP:i:"); o.puti((int)seg_start); o.puts("\tZL:i:"); o.puti(ins); }
                    o.puts("\tZS:Z:"); o.putc("+-"[chr & 1]); o.putc("+-"[ch]); o.putc('\n');
                }
            } else {
                out_hit(ix, p, o, A, 1, chain, np, pp.na, pp.a_chr, a_loc, ins, ca, rmsn(p, ra[t].len, A.raw), &dummy);
                out_hit(ix, p, o, B, 2, !chain, np, pp.nb, pp.b_chr, b_loc, ins, cb, rmsn(p, rb[t].len, B.raw), &dummy);
            }
            continue;
        }
        // StringAlignUnpair
        const int ma = ra[t].status ? -1 : (int)ra[t].nhits, mb = rb[t].status ? -1 : (int)rb[t].nhits;
        if (p->out_sam) {
            out_unpair_sam(ix, p, o, A, 1, ra[t].chain, rb[t].chain, ma, ra[t].nm, ra[t].chr, ra[t].loc, mb, rb[t].chr, rb[t].loc, &n_a);
            out_unpair_sam(ix, p, o, B, 2, rb[t].chain, ra[t].chain, mb, rb[t].nm, rb[t].chr, rb[t].loc, ma, ra[t].chr, ra[t].loc, &n_b);
        } else {
            out_hit(ix, p, ou, A, 1, ra[t].chain, ma, ra[t].nm, ra[t].chr, ra[t].loc, 0, ca, ra[t].status ? 0 : rmsn(p, ra[t].len, A.raw), &dummy);
            out_hit(ix, p, ou, B, 2, rb[t].chain, mb, rb[t].nm, rb[t].chr, rb[t].loc, 0, cb, rb[t].status ? 0 : rmsn(p, rb[t].len, B.raw), &dummy);
        }
    }
    o.finish(); ou.finish();
    st[0] = n_pairs; st[1] = n_a; st[2] = n_b;
}

std::vector<bsx_view> views_of(const char *const *s, uint32_t n) {
    std::vector<bsx_view> v(n);
    for (uint32_t t = 0; t < n; t++) v[t] = bsx_view{s[t], (uint32_t)strlen(s[t])};
    return v;
}

// a growing sink over a string that may come back from an earlier batch with its buffer: the text starts at offset 0
Out sink_of(std::string &s, size_t estimate) {
    if (s.size() < estimate) s.resize(estimate);
    return Out{s.empty() ? nullptr : &s[0], s.size(), 0, &s};
}

bool needs_watson(const bsx_params *p) { return p->out_ref || !p->out_sam; }

size_t write_all(int fd, const std::vector<std::string> &chunks) {
    size_t tot = 0;
    for (const std::string &c : chunks) {
        size_t off = 0;
        while (off < c.size()) { ssize_t w = write(fd, c.data() + off, c.size() - off); if (w <= 0) return tot; off += (size_t)w; tot += (size_t)w; }
    }
    return tot;
}

}  // namespace

extern "C" size_t bsx_format_se(const bsx_index *ix, const bsx_params *p, uint32_t n, const char *const *names,
                                const char *const *seqs, const char *const *quals, int readset,
                                const bsx_rec *recs, const uint16_t *counts, char *out, size_t cap, uint32_t *n_aligned) {
    Out o{out, cap, 0, nullptr};
    uint32_t na = 0;
    const std::vector<bsx_view> vn = views_of(names, n), vs = views_of(seqs, n), vq = views_of(quals, n);
    se_range(ix, p, 0, n, vn.data(), vs.data(), vq.data(), readset, recs, counts, o, &na, 0);
    if (n_aligned) *n_aligned = na;
    return o.n;
}

extern "C" size_t bsx_format_pe(const bsx_index *ix, const bsx_params *p, uint32_t n,
                                const char *const *names_a, const char *const *seqs_a, const char *const *quals_a,
                                const char *const *names_b, const char *const *seqs_b, const char *const *quals_b,
                                const bsx_pair_rec *pr, const bsx_rec *ra, const bsx_rec *rb,
                                const uint16_t *counts_a, const uint16_t *counts_b,
                                char *out, size_t cap, char *out_unpair, size_t cap_unpair, size_t *n_unpair, uint32_t *n_stats) {
    Out o{out, cap, 0, nullptr}, ou{out_unpair, cap_unpair, 0, nullptr};
    uint32_t st[3];
    const std::vector<bsx_view> na = views_of(names_a, n), sa = views_of(seqs_a, n), qa = views_of(quals_a, n);
    const std::vector<bsx_view> nb = views_of(names_b, n), sb = views_of(seqs_b, n), qb = views_of(quals_b, n);
    pe_range(ix, p, 0, n, na.data(), sa.data(), qa.data(), nb.data(), sb.data(), qb.data(), pr, ra, rb, counts_a, counts_b, o, ou, st, 0);
    if (n_unpair) *n_unpair = ou.n;
    if (n_stats) { n_stats[0] = st[0]; n_stats[1] = st[1]; n_stats[2] = st[2]; }
    return o.n;
}

// ---- chunked, multi-threaded emit: contiguous ranges, one string each, concatenation = input order ----

void bsx_format_se_chunks(const bsx_index *ix, const bsx_params *p, uint32_t n, const bsx_view *names, const bsx_view *seqs,
                          const bsx_view *quals, int readset, const bsx_rec *recs, const uint16_t *counts, int threads,
                          std::vector<std::string> &chunks, uint32_t *n_aligned) {
    threads = std::max(1, std::min<int>(threads, (int)std::max<uint32_t>(n, 1)));
    if (needs_watson(p)) watson(ix);   // lazily downloaded once, before the workers read it
    chunks.resize((size_t)threads);                 // strings a caller hands back keep their buffers: no fresh pages, no zero fill
    std::vector<uint32_t> na((size_t)threads, 0);
    const size_t max_ref_name = longest_name(ix);
    bsx_parallel(threads, n, [&](int t, size_t b, size_t e) {
        Out o = sink_of(chunks[t], (e - b) * 320);
        se_range(ix, p, (uint32_t)b, (uint32_t)e, names, seqs, quals, readset, recs, counts, o, &na[t], max_ref_name);
    });
    if (n_aligned) { uint32_t s = 0; for (uint32_t v : na) s += v; *n_aligned = s; }
}

void bsx_format_pe_chunks(const bsx_index *ix, const bsx_params *p, uint32_t n, const bsx_view *names_a, const bsx_view *seqs_a,
                          const bsx_view *quals_a, const bsx_view *names_b, const bsx_view *seqs_b, const bsx_view *quals_b,
                          const bsx_pair_rec *pr, const bsx_rec *ra, const bsx_rec *rb, const uint16_t *counts_a,
                          const uint16_t *counts_b, int threads, std::vector<std::string> &chunks,
                          std::vector<std::string> &chunks_unpair, uint32_t *n_stats) {
    threads = std::max(1, std::min<int>(threads, (int)std::max<uint32_t>(n, 1)));
    if (needs_watson(p)) watson(ix);
    chunks.resize((size_t)threads); chunks_unpair.resize((size_t)threads);
    std::vector<uint32_t> st((size_t)threads * 3, 0);
    const size_t max_ref_name = longest_name(ix);
    bsx_parallel(threads, n, [&](int t, size_t b, size_t e) {
        Out o = sink_of(chunks[t], (e - b) * 640), ou = sink_of(chunks_unpair[t], 0);
        pe_range(ix, p, (uint32_t)b, (uint32_t)e, names_a, seqs_a, quals_a, names_b, seqs_b, quals_b, pr, ra, rb, counts_a, counts_b, o, ou, &st[3 * t], max_ref_name);
    });
    if (n_stats) { n_stats[0] = n_stats[1] = n_stats[2] = 0; for (int t = 0; t < threads; t++) for (int k = 0; k < 3; k++) n_stats[k] += st[3 * t + k]; }
}

extern "C" size_t bsx_emit_se(const bsx_index *ix, const bsx_params *p, const bsx_reads *a, uint32_t n, int readset,
                              const bsx_rec *recs, const uint16_t *counts, int threads, int fd, uint32_t *n_aligned) {
    if (!ix || !p || !a || !recs || n > a->name.size()) { bsx_set_error("bsx_emit_se: bad argument"); return 0; }
    static thread_local std::vector<std::string> chunks;   // buffers reused from call to call
    bsx_format_se_chunks(ix, p, n, a->name.data(), a->seq.data(), a->qual.data(), readset, recs, counts, bsx_host_threads(threads), chunks, n_aligned);
    return write_all(fd, chunks);
}

extern "C" size_t bsx_emit_pe(const bsx_index *ix, const bsx_params *p, const bsx_reads *a, const bsx_reads *b, uint32_t n,
                              const bsx_pair_rec *pr, const bsx_rec *ra, const bsx_rec *rb,
                              const uint16_t *counts_a, const uint16_t *counts_b, int threads, int fd, int fd_unpair, uint32_t *n_stats) {
    if (!ix || !p || !a || !b || !pr || !ra || !rb || n > a->name.size() || n > b->name.size()) { bsx_set_error("bsx_emit_pe: bad argument"); return 0; }
    static thread_local std::vector<std::string> chunks, chunks2;
    bsx_format_pe_chunks(ix, p, n, a->name.data(), a->seq.data(), a->qual.data(), b->name.data(), b->seq.data(), b->qual.data(),
                         pr, ra, rb, counts_a, counts_b, bsx_host_threads(threads), chunks, chunks2, n_stats);
    const size_t w = write_all(fd, chunks);
    if (fd_unpair >= 0) write_all(fd_unpair, chunks2);
    return w;
}

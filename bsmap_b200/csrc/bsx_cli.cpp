// bsx_cli.cpp -- the `bsmap` command line (main.cpp:234-476) over the C ABI.
//
// Same option grammar as mGetOptions (`-x val` and `-x=val`), same banner, same output files
// (SAM when -o ends in .sam, BSP otherwise; -2 for unpaired BSP hits).  Reads are parsed with the
// reference's token semantics (reads.cpp:83-146), mapped in large batches on the GPU, formatted on
// the host and written in input order (= the reference with -p 1).
// Out of scope (errors out): BAM/SAM input, .bam output, -q quality trimming, -M other than TC.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>
#include "bsx_internal.h"

namespace {

struct Opts {
    bsx_params p;
    std::string a, b, d, o, o2;
    unsigned read_start = 1, read_end = ~0u;
    int num_procs = 8, zero_qual = '!', qual_threshold = 0;
    unsigned batch = 1u << 20;
};

// token reader with ifstream `>>` / getline semantics
struct Reader {
    FILE *f = nullptr; std::vector<char> buf; size_t pos = 0, len = 0; bool eof = false;
    bool open(const char *path) { f = fopen(path, "rb"); buf.resize(1 << 22); return f != nullptr; }
    void close() { if (f) fclose(f); f = nullptr; }
    int get() { if (pos == len) { if (eof) return -1; len = fread(buf.data(), 1, buf.size(), f); pos = 0; if (len == 0) { eof = true; return -1; } } return (unsigned char)buf[pos++]; }
    void unget() { if (pos > 0) pos--; }
    static bool ws(int c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }
    bool token(std::string &s) { s.clear(); int c; while ((c = get()) >= 0 && ws(c)) {} if (c < 0) return false; do { s.push_back((char)c); c = get(); } while (c >= 0 && !ws(c)); if (c >= 0) unget(); return true; }
    int nonws() { int c; while ((c = get()) >= 0 && ws(c)) {} return c; }
    void skipline() { int c; while ((c = get()) >= 0 && c != '\n') {} }
};

struct Batch {
    std::vector<std::string> name, seq, qual;
    void clear() { name.clear(); seq.clear(); qual.clear(); }
};

// ReadClass::LoadBatchReads (reads.cpp:83-119) for FASTA / FASTQ
unsigned load_batch(Reader &r, int fmt, const Opts &o, unsigned &index, unsigned want, Batch &b) {
    b.clear();
    std::string tok;
    while (b.name.size() < want && index < o.read_end) {
        int c = r.nonws();
        if (c < 0) break;
        std::string nm, sq, ql;
        if (!r.token(nm)) break;
        r.skipline();
        if (!r.token(sq)) sq.clear();
        if (fmt == 0) { r.token(tok); r.skipline(); r.token(ql); }
        else ql.assign(sq.size(), (char)(o.zero_qual + 40));          // zero_qual + default_qual (reads.cpp:108)
        if ((int)sq.size() > o.p.max_readlen) { sq.erase(o.p.max_readlen); if ((int)ql.size() > o.p.max_readlen) ql.erase(o.p.max_readlen); }
        b.name.push_back(nm); b.seq.push_back(sq); b.qual.push_back(ql);
        index++;
    }
    return (unsigned)b.name.size();
}

int sniff(const char *path) {   // CheckFile (reads.cpp:13-54): 1 = FASTA, 0 = FASTQ, -1 = unsupported
    FILE *f = fopen(path, "rb"); if (!f) return -2;
    int c; while ((c = fgetc(f)) >= 0 && Reader::ws(c)) {}
    fclose(f);
    return c == '>' ? 1 : c == '@' ? 0 : -1;
}

void skip_reads(Reader &r, int fmt, unsigned n) {   // -B (reads.cpp:56-66): 4 (fq) or 2 (fa) lines per read
    for (unsigned long long i = 0; i < (unsigned long long)n * (fmt == 0 ? 4 : 2); i++) { if (r.eof) break; r.skipline(); }
}

void usage() {
    printf("Usage:\tbsmap [options]\n"
           "       -a  <str>   query a file, FASTA/FASTQ format\n"
           "       -d  <str>   reference sequences file, FASTA format\n"
           "       -o  <str>   output alignment file, BSP/SAM format\n"
           "\n  Options for alignment:\n"
           "       -s  <int>   seed size, default=16(WGBS mode), 12(RRBS mode). min=8, max=16.\n"
           "       -v  <int>   maximum number of mismatches allowed on a read, <=15. default=2.\n"
           "       -w  <int>   maximum number of equal best hits to count, <=1000\n"
           "       -B  <int>   start from the Nth read or read pair, default: 1\n"
           "       -E  <int>   end at the Nth read or read pair, default: 4,294,967,295\n"
           "       -I  <int>   index interval, default=4\n"
           "       -p  <int>   accepted for compatibility (the GPU path ignores it)\n"
           "       -D  <str>   activating RRBS mapping mode and set restriction enzyme digestion sites, example: -D C-CGG\n"
           "       -S  <int>   seed for random number generation used in selecting multiple hits\n"
           "       -n  [0,1]   set mapping strand information. default: -n 0\n"
           "\n  Options for trimming:\n"
           "       -f  <int>   filter low-quality reads containing >n Ns, default=5\n"
           "       -A  <str>   3-end adapter sequence, default: none (no trim)\n"
           "       -L  <int>   map the first N nucleotides of the read, default:144 (map the whole read).\n"
           "\n  Options for reporting:\n"
           "       -r  [0,1]   how to report repeat hits, 0=none(unique hit/pair only); 1=random one, default:1.\n"
           "       -R          print corresponding reference sequences in SAM output, default=off\n"
           "       -u          report unmapped reads, default=off\n"
           "\n  Options for pair-end alignment:\n"
           "       -b  <str>   query b file\n"
           "       -m  <int>   minimal insert size allowed, default=28\n"
           "       -x  <int>   maximal insert size allowed, default=500\n"
           "       -2  <str>   output file of unpaired alignment hits\n"
           "       -h          help\n\n");
    exit(1);
}

void set_digestion(Opts &o, const char *a) {   // Param::SetDigestionSite (param.cpp:95-106)
    std::string s = a;
    size_t pos = s.find('-');
    if (pos == std::string::npos) { printf("Digestion position not marked, use '-' to mark. example: 'C-CGG'\n"); exit(1); }
    s.erase(pos, 1);
    memset(o.p.digest_site, 0, sizeof o.p.digest_site);
    strncpy(o.p.digest_site, s.c_str(), sizeof o.p.digest_site - 1);
    o.p.digest_pos = (int)pos; o.p.rrbs = 1; o.p.index_interval = 1; o.p.seed_size = 12;
}

// mGetOptions (main.cpp:234-289); returns the argv index of an unknown option, 0 when fine
int get_options(int argc, char **argv, Opts &o) {
    for (int i = 1; i < argc; i++) {
        if (argv[i][0] != '-') return i;
        const char c = argv[i][1];
        const char *val = nullptr;
        const bool flag = (c == 'R' || c == 'u' || c == 'h');
        if (c == 0 || !strchr("abdo2smxnrIvwqfzpARuBEDMLSh", c)) return i;   // default: return i (main.cpp:283)
        if (!flag) {
            if (argv[i][2] == 0) { if (i + 1 >= argc) return i; val = argv[++i]; }
            else if (argv[i][2] == '=') val = argv[i] + 3;
            else return i;
        } else if (argv[i][2] != 0) return i;
        switch (c) {
            case 'a': o.a = val; break;
            case 'b': o.b = val; o.p.pairend = 1; break;
            case 'd': o.d = val; break;
            case 'o': o.o = val; break;
            case '2': o.o2 = val; break;
            case 's': o.p.seed_size = atoi(val); if (o.p.rrbs) o.p.seed_size = 12; break;
            case 'm': o.p.min_insert = atoi(val); break;
            case 'x': o.p.max_insert = atoi(val); break;
            case 'n': o.p.chains = atoi(val) != 0; break;
            case 'r': o.p.report_repeat_hits = atoi(val); break;
            case 'I': o.p.index_interval = atoi(val); if (o.p.rrbs) o.p.index_interval = 1;
                      if (o.p.index_interval > 16) { fprintf(stderr, "index interval exceeds max value:16\n"); exit(1); } break;
            case 'v': o.p.max_snp_num = atoi(val); if (o.p.max_snp_num > BSX_MAXSNPS) { fprintf(stderr, "number of mismatches exceeds max value:%d\n", BSX_MAXSNPS); exit(1); } break;
            case 'w': o.p.max_num_hits = atoi(val); if (o.p.max_num_hits > BSX_MAXHITS) { fprintf(stderr, "number of multi-hits exceeds max value:%d\n", BSX_MAXHITS); exit(1); } break;
            case 'q': o.qual_threshold = atoi(val); break;
            case 'f': o.p.max_ns = atoi(val); break;
            case 'z': o.zero_qual = atoi(val); break;
            case 'p': o.num_procs = atoi(val); break;
            case 'A': if (o.p.n_adapter < BSX_MAX_ADAPTERS) { strncpy(o.p.adapter[o.p.n_adapter], val, 63); o.p.n_adapter++; } break;
            case 'R': o.p.out_ref = 1; break;
            case 'u': o.p.out_unmap = 1; break;
            case 'B': { int v = atoi(val); o.read_start = v > 1 ? (unsigned)v : 1u; } break;
            case 'E': o.read_end = (unsigned)atoi(val); break;
            case 'D': set_digestion(o, val); break;
            case 'M': if (!((val[0] == 'T' || val[0] == 't') && (val[1] == 'C' || val[1] == 'c'))) { fprintf(stderr, "-M %s: only the TC transition is supported by the GPU path\n", val); exit(1); } break;
            case 'L': o.p.max_readlen = atoi(val); break;
            case 'S': o.p.randseed = atoi(val); break;
            case 'h': usage(); break;
            default: return i;
        }
    }
    return 0;
}

struct Cbuf {   // arrays of C strings for the ABI
    std::vector<const char *> v;
    const char *const *set(const std::vector<std::string> &s) { v.resize(s.size()); for (size_t i = 0; i < s.size(); i++) v[i] = s[i].c_str(); return v.data(); }
};

void pack(const Batch &b, unsigned stride, std::vector<char> &buf, std::vector<uint16_t> &lens) {
    const size_t n = b.seq.size();
    buf.assign(n * stride, 0); lens.resize(n);
    for (size_t i = 0; i < n; i++) {
        size_t l = b.seq[i].size(); if (l > stride) l = stride;
        memcpy(&buf[i * stride], b.seq[i].data(), l);
        lens[i] = (uint16_t)l;
    }
}

}  // namespace

extern "C" int bsx_cli_main(int argc, char **argv) {
    const time_t t0 = time(nullptr);
    printf("\nBSMAP v2.6 (bsmap_b200: B200-native hot path)\n");
    if (argc == 1) usage();
    { time_t t = time(nullptr); printf("Start at:  %s\n", ctime(&t)); }
    Opts o; bsx_params_default(&o.p);
    if (int bad = get_options(argc, argv, o)) { printf("unknown option: %s\n", argv[bad]); exit(bad); }
    if (o.qual_threshold != 0) { fprintf(stderr, "-q quality trimming is not supported by the GPU path\n"); return 1; }
    if (o.o.size() > 4) {
        if (o.o.compare(o.o.size() - 4, 4, ".sam") == 0) o.p.out_sam = 1;
        else if (o.o.compare(o.o.size() - 4, 4, ".bam") == 0) { fprintf(stderr, ".bam output is not supported; write .sam and convert\n"); return 1; }
    }
    bsx_index *ix = nullptr;
    if (bsx_index_create_from_fasta(&o.p, o.d.c_str(), 0, &ix) != BSX_OK) { fprintf(stderr, "%s\n", bsx_last_error()); return 1; }
    bsx_index_info info; bsx_index_get_info(ix, &info);
    unsigned long long sum_len = 0; for (uint32_t k = 0; k < info.n_seq; k++) sum_len += bsx_index_seq_size(ix, k);
    printf("Load in %u db seqs, total size %llu bp. %ld secs passed\n", info.n_seq, sum_len, (long)(time(nullptr) - t0));
    printf("total_kmers: %llu\n", (unsigned long long)info.n_keys);
    printf("Create seed table. %ld secs passed\n", (long)(time(nullptr) - t0));

    const bsx_params &p = o.p;
    printf("max mismatches: %d\tmax multi-hits: %d\tmax Ns: %d\tseed size: %d\tindex interval: %d\n", p.max_snp_num, p.max_num_hits, p.max_ns, p.seed_size, p.index_interval);
    printf("quality cutoff: %d\tbase quality char: '%c'\n", o.qual_threshold, o.zero_qual);
    printf("min fragment size:%d\tmax fragemt size:%d\n", p.min_insert, p.max_insert);
    printf("start from read #%u\tend at read #%u\n", o.read_start, o.read_end);
    printf("additional alignment: T in reads => C in reference\n");
    printf("param.chains:%d\n", p.chains);
    if (p.pairend) { printf("mapping strand (read_1): ++,-+%s\n", p.chains ? ",+-,--" : ""); printf("mapping strand (read_2): +-,--%s\n", p.chains ? ",++,-+" : ""); }
    else printf("mapping strand: ++,-+%s\n", p.chains ? ",+-,--" : "");
    for (int i = 0; i < p.n_adapter; i++) printf("adapter sequence%d: %s\n", i + 1, p.adapter[i]);
    if (p.rrbs) { std::string s = p.digest_site; printf("RRBS mode. digestion site: %s-%s\n", s.substr(0, p.digest_pos).c_str(), s.substr(p.digest_pos).c_str()); }

    const bool pe = !o.a.empty() && !o.b.empty();
    if (o.a.empty()) { fprintf(stderr, "missing query file(s)\n"); return 1; }
    Reader ra, rb;
    const int fa = sniff(o.a.c_str()), fb = pe ? sniff(o.b.c_str()) : 0;
    if (fa == -2 || !ra.open(o.a.c_str())) { fprintf(stderr, "failed to open read file%s (check -a option): %s\n", pe ? " #1" : "", o.a.c_str()); return 1; }
    if (pe && (fb == -2 || !rb.open(o.b.c_str()))) { fprintf(stderr, "failed to open read file #2 (check -b option): %s\n", o.b.c_str()); return 1; }
    if (fa < 0 || fb < 0) { fprintf(stderr, "fatal error: unrecognizable format of reads file (FASTA/FASTQ only; SAM/BAM input is not supported).\n"); return 1; }
    if (pe) printf("Pair-end alignment(GPU)\nQuery: %s  %s  Reference: %s  Output: %s  %s\n", o.a.c_str(), o.b.c_str(), o.d.c_str(), o.o.c_str(), o.o2.c_str());
    else printf("Single read alignment(GPU)\nQuery: %s  Reference: %s  Output: %s\n", o.a.c_str(), o.d.c_str(), o.o.c_str());
    FILE *fout = fopen(o.o.c_str(), "wb");
    if (!fout) { fprintf(stderr, "failed to open output file (check -o option): %s\n", o.o.c_str()); return 1; }
    FILE *fun = nullptr;
    if (pe && !p.out_sam) { fun = fopen(o.o2.c_str(), "wb"); if (!fun) { fprintf(stderr, "failed to open output file for unpaired hits (check -2 option): %s\n", o.o2.c_str()); return 1; } }
    std::vector<char> text, text2;
    if (p.out_sam) { size_t n = bsx_format_header(ix, nullptr, 0); text.resize(n + 1); bsx_format_header(ix, text.data(), n + 1); fwrite(text.data(), 1, n, fout); }

    const unsigned stride = 160;
    bsx_mapper *mp = nullptr;
    if (bsx_mapper_create(ix, &p, o.batch, stride, &mp) != BSX_OK) { fprintf(stderr, "%s\n", bsx_last_error()); return 1; }
    skip_reads(ra, fa, o.read_start - 1); if (pe) skip_reads(rb, fb, o.read_start - 1);
    unsigned index_a = o.read_start - 1, index_b = o.read_start - 1;
    Batch ba, bb; Cbuf ca1, ca2, ca3, cb1, cb2, cb3;
    std::vector<char> sa, sb; std::vector<uint16_t> la, lb;
    std::vector<bsx_rec> reca, recb; std::vector<bsx_pair_rec> recp; std::vector<uint16_t> cnta, cntb;
    unsigned long long n_aligned = 0, n_pairs = 0, n_a = 0, n_b = 0;
    for (;;) {
        const unsigned first = index_a;
        const unsigned n1 = load_batch(ra, fa, o, index_a, o.batch, ba);
        const unsigned n2 = pe ? load_batch(rb, fb, o, index_b, o.batch, bb) : n1;
        if (!n1 || n1 != n2) break;
        pack(ba, stride, sa, la);
        reca.resize(n1); cnta.resize((size_t)n1 * 16);
        size_t need, need2 = 0;
        if (!pe) {
            if (bsx_map_se(mp, n1, sa.data(), la.data(), first, 0, reca.data(), cnta.data()) != BSX_OK) { fprintf(stderr, "%s\n", bsx_last_error()); return 1; }
            uint32_t na = 0;
            need = bsx_format_se(ix, &p, n1, ca1.set(ba.name), ca2.set(ba.seq), ca3.set(ba.qual), 0, reca.data(), cnta.data(), nullptr, 0, &na);
            text.resize(need + 1);
            bsx_format_se(ix, &p, n1, ca1.v.data(), ca2.v.data(), ca3.v.data(), 0, reca.data(), cnta.data(), text.data(), need + 1, &na);
            n_aligned += na;
        } else {
            pack(bb, stride, sb, lb);
            recb.resize(n1); recp.resize(n1); cntb.resize((size_t)n1 * 16);
            if (bsx_map_pe(mp, n1, sa.data(), la.data(), sb.data(), lb.data(), first, recp.data(), reca.data(), recb.data(), cnta.data(), cntb.data()) != BSX_OK) {
                fprintf(stderr, "%s\n", bsx_last_error()); return 1; }
            uint32_t st[3] = {0, 0, 0};
            need = bsx_format_pe(ix, &p, n1, ca1.set(ba.name), ca2.set(ba.seq), ca3.set(ba.qual), cb1.set(bb.name), cb2.set(bb.seq), cb3.set(bb.qual),
                                 recp.data(), reca.data(), recb.data(), cnta.data(), cntb.data(), nullptr, 0, nullptr, 0, &need2, st);
            text.resize(need + 1); text2.resize(need2 + 1);
            bsx_format_pe(ix, &p, n1, ca1.v.data(), ca2.v.data(), ca3.v.data(), cb1.v.data(), cb2.v.data(), cb3.v.data(),
                          recp.data(), reca.data(), recb.data(), cnta.data(), cntb.data(), text.data(), need + 1, text2.data(), need2 + 1, &need2, st);
            n_pairs += st[0]; n_a += st[1]; n_b += st[2];
        }
        fwrite(text.data(), 1, need, fout);
        if (fun && need2) fwrite(text2.data(), 1, need2, fun);
        printf("%u reads finished. %ld secs passed\n", index_a - o.read_start + 1, (long)(time(nullptr) - t0));
    }
    fclose(fout); if (fun) fclose(fun);
    ra.close(); rb.close();
    const double tot = (double)(index_a - o.read_start + 1);
    if (pe) printf("Total number of aligned reads: \npairs:       %llu (%.2g%%)\nsingle a:    %llu (%.2g%%)\nsingle b:    %llu (%.2g%%)\n",
                   n_pairs, 100.0 * n_pairs / tot, n_a, 100.0 * n_a / tot, n_b, 100.0 * n_b / tot);
    else printf("Total number of aligned reads: %llu (%.2g%%)\n", n_aligned, 100.0 * n_aligned / tot);
    printf("Done.\n");
    { time_t t = time(nullptr); printf("Finished at %s", ctime(&t)); }
    printf("Total time consumed:  %ld secs\n", (long)(time(nullptr) - t0));
    bsx_mapper_destroy(mp); bsx_index_destroy(ix);
    return 0;
}

// oracle/glue/bsmap_glue.cpp -- TEST INFRASTRUCTURE: the reference-side binding of INTEGRATION.md, as a real
// translation unit.
//
// `make -C oracle glued` links the UNMODIFIED reference program (main.cpp with its option parser, reader threads and
// output loop, reads.cpp, param.cpp, dbseq.cpp, utilities.cpp -- compiled where they lie under /root/reference) with
// this file instead of the bodies of SingleAlign::Do_Batch (align.cpp:591) and PairAlign::Do_Batch (pairs.cpp:192):
// the reference's align.cpp / pairs.cpp are compiled with -DDo_Batch=Do_Batch_reference, so that the name the worker
// threads call (main.cpp:62, 101) resolves here.  Every batch then goes through the C ABI of include/bsmap_b200.h:
//
//     mreads (ImportBatchReads, align.cpp:42)  ->  bsx_map_se / bsx_map_pe  ->  bsx_format_se / bsx_format_pe  ->  _str_align
//
// and the reference writes _str_align to its own output stream.  tests/test_glue_gpu.py runs the result
// (oracle/_ref/bsmap_glued) against the goldens of the unmodified binary: the seam is link-checked and byte-checked.
// The reference still builds its own seed table in main() (Do_Formatdb, main.cpp:174-178); the glue ignores it and keeps
// its device index next to it.
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>
#include <pthread.h>

#include "align.h"
#include "pairs.h"
#include "bsmap_b200.h"

extern Param param;          // main.cpp:23
extern string ref_file;      // main.cpp:26 (-d)

namespace {

// Param (param.h:54-121) -> bsx_params, field for field
bsx_params to_bsx(const Param &p) {
    bsx_params q; bsx_params_default(&q);
    q.seed_size = p.seed_size;           q.index_interval = p.index_interval;
    q.max_snp_num = p.max_snp_num;       q.max_num_hits = p.max_num_hits;
    q.report_repeat_hits = p.report_repeat_hits;
    q.min_insert = p.min_insert;         q.max_insert = p.max_insert;
    q.chains = p.chains;                 q.pairend = p.pairend;       q.rrbs = p.RRBS_flag;
    q.randseed = p.randseed;             q.max_ns = p.max_ns;         q.max_readlen = p.max_readlen;
    q.out_sam = p.out_sam;               q.out_unmap = p.out_unmap;   q.out_ref = p.out_ref;
    q.digest_pos = p.digest_pos;         strncpy(q.digest_site, p.digest_site.c_str(), sizeof q.digest_site - 1);
    q.n_adapter = p.n_adapter;
    for (int i = 0; i < p.n_adapter; i++) strncpy(q.adapter[i], p.adapter[i].c_str(), sizeof q.adapter[i] - 1);
    return q;
}

void die() { std::cerr << "bsmap_b200: " << bsx_last_error() << std::endl; exit(1); }

const uint32_t kStride = 160;     // read slot: a multiple of 8, at least the longest read (READ_144)
const uint32_t kBatch = 50000;    // BatchNum (reads.h:13)
pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
bsx_index *g_index = 0;           // RefSeq::Run_ConvertBinseq + CreateIndex (dbseq.cpp:215, 516), once per process

bsx_index *the_index(const bsx_params &q) {
    pthread_mutex_lock(&g_mu);
    if (!g_index && bsx_index_create_from_fasta(&q, ref_file.c_str(), /*device*/ 0, &g_index) != BSX_OK) die();
    pthread_mutex_unlock(&g_mu);
    return g_index;
}

// one worker thread = one mapper (like each worker's SingleAlign / PairAlign object)
bsx_mapper *the_mapper(const bsx_params &q) {
    static __thread bsx_mapper *mp = 0;
    if (!mp && bsx_mapper_create(the_index(q), &q, kBatch, kStride, &mp) != BSX_OK) die();
    return mp;
}

struct Batch {
    std::vector<char> seq; std::vector<uint16_t> len; std::vector<const char *> nm, sq, ql;
    void fill(vector<ReadInf> &r, bit32_t n) {
        seq.assign((size_t)n * kStride, 0); len.resize(n); nm.resize(n); sq.resize(n); ql.resize(n);
        for (bit32_t i = 0; i < n; i++) {
            memcpy(&seq[(size_t)i * kStride], r[i].seq.data(), std::min<size_t>(kStride, r[i].seq.size()));
            len[i] = (uint16_t)r[i].seq.size();
            nm[i] = r[i].name.c_str(); sq[i] = r[i].seq.c_str(); ql[i] = r[i].qual.c_str();
        }
    }
};

}  // namespace

// SingleAlign::Do_Batch (align.cpp:591-608): mreads / num_reads were filled by ImportBatchReads
void SingleAlign::Do_Batch(RefSeq &) {
    _str_align.clear();
    if (!num_reads) return;
    const bsx_params q = to_bsx(param);
    bsx_mapper *mp = the_mapper(q);
    Batch b; b.fill(mreads, num_reads);
    std::vector<bsx_rec> rec(num_reads); std::vector<uint16_t> cnt((size_t)num_reads * 16);
    if (bsx_map_se(mp, num_reads, b.seq.data(), b.len.data(), mreads[0].index, (int)mreads[0].readset, rec.data(), cnt.data()) != BSX_OK) die();
    uint32_t aligned = 0;
    const size_t need = bsx_format_se(g_index, &q, num_reads, b.nm.data(), b.sq.data(), b.ql.data(), (int)mreads[0].readset, rec.data(), cnt.data(), 0, 0, &aligned);
    _str_align.resize(need + 1);
    bsx_format_se(g_index, &q, num_reads, b.nm.data(), b.sq.data(), b.ql.data(), (int)mreads[0].readset, rec.data(), cnt.data(), &_str_align[0], need + 1, &aligned);
    _str_align.resize(need);
    n_aligned += aligned;
}

// PairAlign::Do_Batch (pairs.cpp:192-220): _sa.mreads / _sb.mreads hold the mates
void PairAlign::Do_Batch(RefSeq &) {
    _str_align.clear(); _str_align_unpair.clear();
    if (!num_reads) return;
    const bsx_params q = to_bsx(param);
    bsx_mapper *mp = the_mapper(q);
    Batch a, b; a.fill(_sa.mreads, num_reads); b.fill(_sb.mreads, num_reads);
    std::vector<bsx_pair_rec> pr(num_reads); std::vector<bsx_rec> ra(num_reads), rb(num_reads);
    std::vector<uint16_t> ca((size_t)num_reads * 16), cb((size_t)num_reads * 16);
    if (bsx_map_pe(mp, num_reads, a.seq.data(), a.len.data(), b.seq.data(), b.len.data(), _sa.mreads[0].index,
                   pr.data(), ra.data(), rb.data(), ca.data(), cb.data()) != BSX_OK) die();
    uint32_t st[3] = {0, 0, 0}; size_t need_un = 0;
    const size_t need = bsx_format_pe(g_index, &q, num_reads, a.nm.data(), a.sq.data(), a.ql.data(), b.nm.data(), b.sq.data(), b.ql.data(),
                                      pr.data(), ra.data(), rb.data(), ca.data(), cb.data(), 0, 0, 0, 0, &need_un, st);
    _str_align.resize(need + 1); _str_align_unpair.resize(need_un + 1);
    bsx_format_pe(g_index, &q, num_reads, a.nm.data(), a.sq.data(), a.ql.data(), b.nm.data(), b.sq.data(), b.ql.data(),
                  pr.data(), ra.data(), rb.data(), ca.data(), cb.data(), &_str_align[0], need + 1, &_str_align_unpair[0], need_un + 1, &need_un, st);
    _str_align.resize(need); _str_align_unpair.resize(need_un);
    n_aligned_pairs += st[0]; n_aligned_a += st[1]; n_aligned_b += st[2];
}

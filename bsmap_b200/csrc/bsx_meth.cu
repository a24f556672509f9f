// bsx_meth.cu -- methratio.py on the device: per-position methylation counters piled up from
// the mappings by a warp-per-alignment kernel with global atomics, CpG combining, and the host report writer.
//
// Reference: methratio.py.  get_alignment (30-65): filters, fill-in trimming, mate-overlap removal;
// pile-up (95-118): on a '+' (Watson) hit every reference C under the read counts T as depth and C as
// meth + depth, on a '-' (Crick) hit every reference G counts A / G; -g (122-131); report (133-154).
// The reference holds two u32 arrays per chromosome in host memory and walks each read with str.find;
// here both arrays cover the packed Watson strand of the index in HBM (8 bytes per reference position) and the
// reference base comes from the 2-bit refcat (non-ACGT packs to A, which is neither C nor G -- same outcome).
// -r (methratio.py:52-55: of the alignments that pass the filters, the first in file order per (chromosome, fragment
// end, direction) counts) is a two-kernel min-reduction: every alignment of a batch claims its key with
// atomicMin(first[key], order + 1), then the pile-up kernel lets the alignment through whose claim stands and marks
// the key 0 = taken for good, so later batches lose against earlier ones.  8 bytes per reference position, allocated
// on first use.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>
#include "bsx_internal.h"
#include "bsx_common.cuh"

struct bsx_meth {
    const bsx_index *ix = nullptr;
    uint32_t *d_meth = nullptr, *d_depth = nullptr;   // n_words * 16 counters each, indexed like refcat bases
    unsigned long long *d_valid = nullptr;
    uint32_t *d_first = nullptr;                      // -r: per (position, direction) the order + 1 of the claiming alignment; 0 = taken
    cudaEvent_t dup_order = nullptr;                  // -r, in-process: the pile-up of batch k+1 waits for that of batch k
    bool combined = false;
    // staging for bsx_meth_add
    size_t cap = 0; uint32_t stride = 0;
    char *d_seq = nullptr; uint16_t *d_len = nullptr; uint32_t *d_chr = nullptr, *d_pos = nullptr;
    uint8_t *d_strand = nullptr, *d_flags = nullptr; int32_t *d_ins = nullptr, *d_mate = nullptr;
};

namespace {

__device__ __forceinline__ uint32_t ref_code(const uint32_t *refcat, uint32_t q) { return (refcat[q >> 4] >> (30 - 2 * (q & 15))) & 3u; }

// One alignment, executed by a warp: get_alignment's trimming (methratio.py:56-64), the bounds test (105) and the
// pile-up (107-118).  getc(i) = i-th character of SEQ as printed.  Filters (-u / -p) were applied by the caller.
template <class GetC>
__device__ __forceinline__ void pile_alignment(const uint32_t *__restrict__ refcat, const uint32_t *__restrict__ seqinfo, uint32_t n_seq,
                                               uint32_t *meth, uint32_t *depth, unsigned long long *n_valid, const bsx_meth_opts &o,
                                               uint32_t k, long long pos, long long len, int st, int ins, long long mate_pos, bool sam, int lane, GetC getc) {
    const uint32_t *anchor = seqinfo, *size = seqinfo + n_seq + 1;
    long long start = 0;
    const int N = o.trim_fillin;
    const bool first_minus = st & 1, second_minus = (st >> 1) & 1;
    if (N > 0) {                                             // trim fill-in nucleotides (methratio.py:56-63)
        if (!first_minus && second_minus) len = len - N > 0 ? len - N : 0;                     // '+-': seq[:-N]
        else if (first_minus && second_minus) { start = N < len ? N : len; len -= start; pos += N; }   // '--': seq[N:], pos + N
        else if (ins != 0 && len > (long long)abs(ins) - N) {
            const long long trim = len - ((long long)abs(ins) - N);
            if (!first_minus) len = len - trim > 0 ? len - trim : 0;                           // '++': seq[:-trim]
            else { start = trim < len ? trim : len; len -= start; pos += trim; }                // '-+': seq[trim:], pos + trim
        }
    }
    if (sam && ins > 0) {                                    // remove the region overlapped by the mate (methratio.py:64)
        const long long e = mate_pos - pos;                  // seq[:e] with Python slice semantics
        if (e < 0) len = len + e > 0 ? len + e : 0; else if (e < len) len = e;
    }
    if (pos + len > (long long)size[k]) return;             // methratio.py:105
    if (lane == 0) atomicAdd(n_valid, 1ull);
    const uint32_t base = anchor[k] + (uint32_t)pos;
    const uint32_t match = first_minus ? 2u : 1u;           // '+': C (converted reads show T), '-': G (A)
    const char cm = first_minus ? 'G' : 'C', cc = first_minus ? 'A' : 'T';
    for (int i = lane; i < (int)len; i += 32) {
        if (ref_code(refcat, base + (uint32_t)i) != match) continue;
        const char c = getc((int)start + i);
        if (c == cc) atomicAdd(depth + base + i, 1u);
        else if (c == cm) { atomicAdd(meth + base + i, 1u); atomicAdd(depth + base + i, 1u); }
    }
}

// what get_alignment hands on (before trimming): sequence index, 0-based position, printed length, strand bits, ...
struct AlnView { uint32_t k; long long pos, len; int st, ins; long long mate_pos; bool sam; };

// -r key of an alignment (methratio.py:52-53): '+-' / '-+' hits are identified by where they end (direction 2), the
// others by where they start (direction 1).  A fragment end outside [0, size) makes the script raise IndexError (or
// wrap to the chromosome's last entry for -1); BSMAP never prints such a record, and here it skips the duplicate test.
__device__ __forceinline__ bool dup_key(const uint32_t *__restrict__ seqinfo, uint32_t n_seq, const AlnView &v, uint64_t &key) {
    const uint32_t *anchor = seqinfo, *size = seqinfo + n_seq + 1;
    const bool dir2 = ((v.st ^ (v.st >> 1)) & 1) != 0;
    const long long fe = dir2 ? v.pos + v.len : v.pos;
    if (fe < 0 || fe >= (long long)size[v.k]) return false;
    key = ((uint64_t)anchor[v.k] + (uint64_t)fe) * 2u + (dir2 ? 1u : 0u);
    return true;
}
// resolve (warp-uniform): does alignment `order` hold its key?  The holder marks it taken for every later batch.  A
// loser reads either the holder's claim or the 0 the holder has just written: both differ from its own claim.
__device__ __forceinline__ bool dup_holds(uint32_t *first, const uint32_t *__restrict__ seqinfo, uint32_t n_seq, const AlnView &v, uint32_t order, int lane) {
    uint64_t key;
    if (!dup_key(seqinfo, n_seq, v, key)) return true;
    uint32_t w = 0;
    if (lane == 0) { w = first[key]; if (w == order + 1u) first[key] = 0u; }
    w = __shfl_sync(0xffffffffu, w, 0);
    return w == order + 1u;
}

// alignment a of a batch parsed from SAM / BSP text on the host, after the -u / -p filters (methratio.py:35-36, 48-49)
__device__ __forceinline__ bool parsed_alignment(const bsx_meth_opts &o, uint32_t n_seq, uint32_t a, const uint16_t *__restrict__ lens,
                                                 const uint32_t *__restrict__ chr, const uint32_t *__restrict__ pos0, const uint8_t *__restrict__ strand,
                                                 const int32_t *__restrict__ insert, const int32_t *__restrict__ mate_pos, const uint8_t *__restrict__ flags, AlnView &v) {
    const uint32_t fl = flags[a];
    if (o.unique && (fl & BSX_METH_SECONDARY)) return false;
    if (o.pair && !(fl & BSX_METH_PROPER)) return false;
    v.k = chr[a];
    if (v.k >= n_seq) return false;
    v.pos = (long long)pos0[a]; v.len = (long long)lens[a]; v.st = strand[a]; v.ins = insert[a]; v.mate_pos = (long long)mate_pos[a];
    v.sam = (fl & BSX_METH_SAM) != 0;
    return true;
}

// -r, first kernel: one thread per alignment claims its key
__global__ void __launch_bounds__(256) meth_claim_kernel(const uint32_t *__restrict__ seqinfo, uint32_t n_seq, uint32_t *first, bsx_meth_opts o, uint32_t n,
                                                         const uint16_t *__restrict__ lens, const uint32_t *__restrict__ chr, const uint32_t *__restrict__ pos0,
                                                         const uint8_t *__restrict__ strand, const int32_t *__restrict__ insert,
                                                         const int32_t *__restrict__ mate_pos, const uint8_t *__restrict__ flags) {
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    AlnView v; uint64_t key;
    if (a < n && parsed_alignment(o, n_seq, a, lens, chr, pos0, strand, insert, mate_pos, flags, v) && dup_key(seqinfo, n_seq, v, key)) atomicMin(first + key, a + 1u);
}

// alignments parsed from SAM / BSP text on the host: one warp per alignment
__global__ void __launch_bounds__(256) meth_pileup_kernel(const uint32_t *__restrict__ refcat, const uint32_t *__restrict__ seqinfo, uint32_t n_seq,
                                                          uint32_t *meth, uint32_t *depth, unsigned long long *n_valid, uint32_t *first, bsx_meth_opts o, uint32_t n,
                                                          const char *__restrict__ seqs, uint32_t stride, const uint16_t *__restrict__ lens,
                                                          const uint32_t *__restrict__ chr, const uint32_t *__restrict__ pos0,
                                                          const uint8_t *__restrict__ strand, const int32_t *__restrict__ insert,
                                                          const int32_t *__restrict__ mate_pos, const uint8_t *__restrict__ flags) {
    const uint32_t a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (a >= n) return;
    AlnView v;
    if (!parsed_alignment(o, n_seq, a, lens, chr, pos0, strand, insert, mate_pos, flags, v)) return;
    if (first && !dup_holds(first, seqinfo, n_seq, v, a, lane)) return;
    const char *sq = seqs + (size_t)a * stride;
    pile_alignment(refcat, seqinfo, n_seq, meth, depth, n_valid, o, v.k, v.pos, v.len, v.st, v.ins, v.mate_pos, v.sam, lane, [sq](int i) { return sq[i]; });
}

__device__ __forceinline__ char comp_char(char c) {   // rev_char[] (param.cpp:166-177)
    switch (c) {
        case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
        case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
        default: return 'N';
    }
}

// Unit u of the batch a mapper has just mapped (one read; PE: one mate), straight from its device buffers.  Which reads
// are printed as mapped, with which SEQ orientation / POS / TLEN / PNEXT, restates s_OutHit (align.cpp:631-765),
// s_OutHitPair (pairs.cpp:288-424) and s_OutHitUnpair (pairs.cpp:426-498); then the -u / -p filters.
__device__ __forceinline__ bool mapped_alignment(const bsx_meth_opts &o, int sam, int report_repeat_hits, uint32_t u, int mates,
                                                 const bsx_rec *__restrict__ out_a, const bsx_rec *__restrict__ out_b,
                                                 const bsx_pair_rec *__restrict__ out_pair, AlnView &v, bool &rev) {
    const uint32_t r = mates == 2 ? u >> 1 : u;
    const int mate = mates == 2 ? (int)(u & 1u) : 0;
    const bsx_rec rc = (mate ? out_b : out_a)[r];
    uint32_t chr, loc; int chain, lp = rc.len, ins = 0; long long mate_pos = -1; bool secondary, proper = false;
    if (mates == 2 && out_pair[r].paired) {
        const bsx_pair_rec pp = out_pair[r];
        const int la = out_a[r].len, lb = out_b[r].len;
        uint32_t a_loc = pp.a_loc, b_loc = pp.b_loc;
        // fragment shorter than the read: the adapter part is dropped (pairs.cpp:296-306)
        if (pp.insert < la && ((int)pp.chain ^ (int)(pp.a_chr & 1u))) a_loc += (uint32_t)(la - pp.insert);
        if (pp.insert < lb && ((!pp.chain) ^ (int)(pp.b_chr & 1u))) b_loc += (uint32_t)(lb - pp.insert);
        chr = mate ? pp.b_chr : pp.a_chr; loc = mate ? b_loc : a_loc; chain = mate ? !pp.chain : pp.chain;
        if (pp.insert < lp) lp = pp.insert;
        const bool rv = (chain ^ (int)(chr & 1u)) != 0;
        ins = sam ? (rv ? -pp.insert : pp.insert) : pp.insert;       // TLEN (pairs.cpp:330-340) / BSP insert column
        mate_pos = mate ? a_loc : b_loc;                             // PNEXT - 1
        secondary = pp.npairs > 1; proper = true;
    } else {
        const int nh = rc.status ? -1 : (int)rc.nhits;
        if (nh <= 0 || (nh > 1 && report_repeat_hits == 0)) return false;   // printed as unmapped ('u') or not at all
        chr = rc.chr; loc = rc.loc; chain = rc.chain; secondary = nh > 1;
    }
    if (o.unique && secondary) return false;
    if (o.pair && !proper) return false;
    rev = (chain ^ (int)(chr & 1u)) != 0;
    v.k = chr >> 1; v.pos = (long long)loc; v.len = (long long)lp; v.st = (int)(chr & 1u) | (chain << 1); v.ins = ins; v.mate_pos = mate_pos; v.sam = sam != 0;
    return true;
}

// -r, first kernel of the in-process form: one thread per unit claims its key (the order of the SAM lines: read by read, mate a before mate b)
__global__ void __launch_bounds__(256) meth_claim_mapped_kernel(const uint32_t *__restrict__ seqinfo, uint32_t n_seq, uint32_t *first, bsx_meth_opts o,
                                                                int sam, int report_repeat_hits, uint32_t n, int mates,
                                                                const bsx_rec *__restrict__ out_a, const bsx_rec *__restrict__ out_b,
                                                                const bsx_pair_rec *__restrict__ out_pair) {
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    AlnView v; bool rev; uint64_t key;
    if (u < n * (uint32_t)mates && mapped_alignment(o, sam, report_repeat_hits, u, mates, out_a, out_b, out_pair, v, rev) && dup_key(seqinfo, n_seq, v, key))
        atomicMin(first + key, u + 1u);
}

// pile-up of the mapped batch: one warp per read (PE: per mate)
__global__ void __launch_bounds__(256) meth_pileup_mapped_kernel(const uint32_t *__restrict__ refcat, const uint32_t *__restrict__ seqinfo, uint32_t n_seq,
                                                                 uint32_t *meth, uint32_t *depth, unsigned long long *n_valid, uint32_t *first, bsx_meth_opts o,
                                                                 int sam, int report_repeat_hits, uint32_t n, int mates, uint32_t stride,
                                                                 const uint8_t *__restrict__ seq_a, const uint8_t *__restrict__ seq_b,
                                                                 const bsx_rec *__restrict__ out_a, const bsx_rec *__restrict__ out_b,
                                                                 const bsx_pair_rec *__restrict__ out_pair) {
    const uint32_t u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (u >= n * (uint32_t)mates) return;
    AlnView v; bool rev;
    if (!mapped_alignment(o, sam, report_repeat_hits, u, mates, out_a, out_b, out_pair, v, rev)) return;
    if (first && !dup_holds(first, seqinfo, n_seq, v, u, lane)) return;
    const uint32_t r = mates == 2 ? u >> 1 : u;
    const uint8_t *rd = ((mates == 2 && (u & 1u)) ? seq_b : seq_a) + (size_t)r * stride;
    const int lp = (int)v.len;
    pile_alignment(refcat, seqinfo, n_seq, meth, depth, n_valid, o, v.k, v.pos, v.len, v.st, v.ins, v.mate_pos, v.sam, lane,
                   [rd, rev, lp](int i) { return rev ? comp_char((char)rd[lp - 1 - i]) : (char)rd[i]; });
}

// -g: every reference "CG": both counters of the C take the G's, the G's become 0 (methratio.py:122-131)
__global__ void meth_combine_kernel(const uint32_t *__restrict__ refcat, uint64_t n_pos, uint32_t *meth, uint32_t *depth) {
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q + 1 >= n_pos) return;
    if (ref_code(refcat, (uint32_t)q) == 1u && ref_code(refcat, (uint32_t)q + 1u) == 2u) {
        depth[q] += depth[q + 1]; meth[q] += meth[q + 1];
        depth[q + 1] = 0; meth[q + 1] = 0;
    }
}

int ensure_staging(bsx_meth *m, size_t n, uint32_t stride) {
    if (n <= m->cap && stride <= m->stride) return BSX_OK;
    cudaFree(m->d_seq); cudaFree(m->d_len); cudaFree(m->d_chr); cudaFree(m->d_pos); cudaFree(m->d_strand); cudaFree(m->d_flags); cudaFree(m->d_ins); cudaFree(m->d_mate);
    m->cap = std::max(n, m->cap); m->stride = std::max(stride, m->stride);
    BSX_CUDA_CHECK(cudaMalloc(&m->d_seq, m->cap * m->stride));
    BSX_CUDA_CHECK(cudaMalloc(&m->d_len, m->cap * 2));
    BSX_CUDA_CHECK(cudaMalloc(&m->d_chr, m->cap * 4));
    BSX_CUDA_CHECK(cudaMalloc(&m->d_pos, m->cap * 4));
    BSX_CUDA_CHECK(cudaMalloc(&m->d_strand, m->cap));
    BSX_CUDA_CHECK(cudaMalloc(&m->d_flags, m->cap));
    BSX_CUDA_CHECK(cudaMalloc(&m->d_ins, m->cap * 4));
    BSX_CUDA_CHECK(cudaMalloc(&m->d_mate, m->cap * 4));
    return BSX_OK;
}

// -r: the claim table, all keys free
int ensure_dup_table(bsx_meth *m) {
    if (m->d_first) return BSX_OK;
    const size_t bytes = m->ix->n_words * 16 * 2 * sizeof(uint32_t);
    if (cudaMalloc(&m->d_first, bytes) != cudaSuccess) { cudaGetLastError(); bsx_set_error("-r: out of device memory (%zu bytes for the duplicate table)", bytes); return BSX_ERR_CUDA; }
    BSX_CUDA_CHECK(cudaMemset(m->d_first, 0xff, bytes));
    BSX_CUDA_CHECK(cudaEventCreateWithFlags(&m->dup_order, cudaEventDisableTiming));
    BSX_CUDA_CHECK(cudaDeviceSynchronize());
    return BSX_OK;
}

int combine_once(bsx_meth *m, const bsx_meth_opts *o) {
    if (!o->combine_cpg || m->combined) return BSX_OK;
    const uint64_t n_pos = m->ix->n_words * 16;
    meth_combine_kernel<<<(unsigned)((n_pos + 255) / 256), 256>>>(m->ix->d_refcat, n_pos, m->d_meth, m->d_depth);
    BSX_CUDA_CHECK(cudaGetLastError());
    BSX_CUDA_CHECK(cudaDeviceSynchronize());
    m->combined = true;
    return BSX_OK;
}

struct Sink {
    std::string *s;
    void put(const char *p, size_t n) { s->append(p, n); }
    void putc(char c) { s->push_back(c); }
    void putu(unsigned long long v) { char b[24]; char *e = b + 24, *q = e; do { *--q = (char)('0' + v % 10); v /= 10; } while (v); put(q, (size_t)(e - q)); }
    void putf3(double v) { char b[48]; const int l = snprintf(b, sizeof b, "%.3f", v); put(b, (size_t)l); }
};

}  // namespace

extern "C" void bsx_meth_opts_default(bsx_meth_opts *o) {
    if (!o) return;
    memset(o, 0, sizeof *o);
    o->trim_fillin = 2; o->min_depth = 1;
}

extern "C" int bsx_meth_create(const bsx_index *ix, bsx_meth **out) {
    if (!ix || !out) { bsx_set_error("bsx_meth_create: bad argument"); return BSX_ERR_ARG; }
    if (ix->device < 0) { bsx_set_error("text-only index has no device arrays: methratio needs the index on a CUDA device"); return BSX_ERR_CUDA; }
    BSX_CUDA_CHECK(cudaSetDevice(ix->device));
    bsx_meth *m = new bsx_meth();
    m->ix = ix;
    const size_t bytes = ix->n_words * 16 * sizeof(uint32_t);
    if (cudaMalloc(&m->d_meth, bytes) != cudaSuccess || cudaMalloc(&m->d_depth, bytes) != cudaSuccess || cudaMalloc(&m->d_valid, 8) != cudaSuccess) {
        cudaGetLastError(); bsx_meth_destroy(m); bsx_set_error("bsx_meth_create: out of device memory (%zu bytes per counter array)", bytes); return BSX_ERR_CUDA; }
    BSX_CUDA_CHECK(cudaMemset(m->d_meth, 0, bytes));
    BSX_CUDA_CHECK(cudaMemset(m->d_depth, 0, bytes));
    BSX_CUDA_CHECK(cudaMemset(m->d_valid, 0, 8));
    *out = m;
    return BSX_OK;
}

extern "C" int bsx_meth_destroy(bsx_meth *m) {
    if (!m) return BSX_OK;
    cudaFree(m->d_meth); cudaFree(m->d_depth); cudaFree(m->d_valid); cudaFree(m->d_first);
    if (m->dup_order) cudaEventDestroy(m->dup_order);
    cudaFree(m->d_seq); cudaFree(m->d_len); cudaFree(m->d_chr); cudaFree(m->d_pos); cudaFree(m->d_strand); cudaFree(m->d_flags); cudaFree(m->d_ins); cudaFree(m->d_mate);
    delete m;
    return BSX_OK;
}

extern "C" int bsx_meth_add(bsx_meth *m, const bsx_meth_opts *o, uint32_t n, const char *seqs, uint32_t stride,
                            const uint16_t *lens, const uint32_t *chr, const uint32_t *pos, const uint8_t *strand,
                            const int32_t *insert, const int32_t *mate_pos, const uint8_t *flags, uint64_t *n_valid) {
    if (!m || !o || (n && (!seqs || !lens || !chr || !pos || !strand || !insert || !mate_pos || !flags))) { bsx_set_error("bsx_meth_add: bad argument"); return BSX_ERR_ARG; }
    if (m->combined) { bsx_set_error("bsx_meth_add: counters were already combined (-g); create a new bsx_meth"); return BSX_ERR_ARG; }
    BSX_CUDA_CHECK(cudaSetDevice(m->ix->device));
    if (n) {
        int rc = ensure_staging(m, n, stride); if (rc) return rc;
        BSX_CUDA_CHECK(cudaMemcpy2D(m->d_seq, m->stride, seqs, stride, stride, n, cudaMemcpyHostToDevice));
        BSX_CUDA_CHECK(cudaMemcpy(m->d_len, lens, (size_t)n * 2, cudaMemcpyHostToDevice));
        BSX_CUDA_CHECK(cudaMemcpy(m->d_chr, chr, (size_t)n * 4, cudaMemcpyHostToDevice));
        BSX_CUDA_CHECK(cudaMemcpy(m->d_pos, pos, (size_t)n * 4, cudaMemcpyHostToDevice));
        BSX_CUDA_CHECK(cudaMemcpy(m->d_strand, strand, n, cudaMemcpyHostToDevice));
        BSX_CUDA_CHECK(cudaMemcpy(m->d_flags, flags, n, cudaMemcpyHostToDevice));
        BSX_CUDA_CHECK(cudaMemcpy(m->d_ins, insert, (size_t)n * 4, cudaMemcpyHostToDevice));
        BSX_CUDA_CHECK(cudaMemcpy(m->d_mate, mate_pos, (size_t)n * 4, cudaMemcpyHostToDevice));
        if (o->rm_dup) {
            rc = ensure_dup_table(m); if (rc) return rc;
            meth_claim_kernel<<<(n + 255) / 256, 256>>>(m->ix->d_seqinfo, m->ix->n_seq, m->d_first, *o, n, m->d_len, m->d_chr, m->d_pos, m->d_strand,
                                                       m->d_ins, m->d_mate, m->d_flags);
            BSX_CUDA_CHECK(cudaGetLastError());
        }
        const unsigned blocks = (unsigned)(((uint64_t)n * 32 + 255) / 256);
        meth_pileup_kernel<<<blocks, 256>>>(m->ix->d_refcat, m->ix->d_seqinfo, m->ix->n_seq, m->d_meth, m->d_depth, m->d_valid,
                                            o->rm_dup ? m->d_first : nullptr, *o, n, m->d_seq, m->stride, m->d_len, m->d_chr, m->d_pos, m->d_strand, m->d_ins, m->d_mate, m->d_flags);
        BSX_CUDA_CHECK(cudaGetLastError());
    }
    if (n_valid) {
        unsigned long long v = 0;
        BSX_CUDA_CHECK(cudaMemcpy(&v, m->d_valid, 8, cudaMemcpyDeviceToHost));
        *n_valid = v;
    }
    return BSX_OK;
}

// pile up the batch a mapper has just mapped (called by bsx_api.cu::run_slot on the batch's stream)
int bsx_meth_pile_mapped(bsx_meth *m, const bsx_meth_opts *o, int sam, int report_repeat_hits, uint32_t n, int mates, uint32_t stride,
                         const uint8_t *seq_a, const uint8_t *seq_b, const bsx_rec *out_a, const bsx_rec *out_b, const bsx_pair_rec *out_pair,
                         cudaStream_t st) {
    if (!m || !o || n == 0) return BSX_OK;
    if (m->combined) { bsx_set_error("bsx_meth: counters were already combined (-g); create a new bsx_meth"); return BSX_ERR_ARG; }
    if (o->rm_dup) {
        // file order = batch order: this batch's claims start when the previous batch has resolved its own (the batches
        // alternate between two streams).  Paired-end BSP output goes to two files, whose concatenation is the script's
        // order: that case is refused by bsx_mapper_attach_meth.
        int rc = ensure_dup_table(m); if (rc) return rc;
        BSX_CUDA_CHECK(cudaStreamWaitEvent(st, m->dup_order, 0));
        meth_claim_mapped_kernel<<<(unsigned)(((uint64_t)n * (uint64_t)mates + 255) / 256), 256, 0, st>>>(m->ix->d_seqinfo, m->ix->n_seq, m->d_first, *o, sam,
                                                                                                      report_repeat_hits, n, mates, out_a, out_b, out_pair);
        BSX_CUDA_CHECK(cudaGetLastError());
    }
    const unsigned blocks = (unsigned)(((uint64_t)n * (uint64_t)mates * 32 + 255) / 256);
    meth_pileup_mapped_kernel<<<blocks, 256, 0, st>>>(m->ix->d_refcat, m->ix->d_seqinfo, m->ix->n_seq, m->d_meth, m->d_depth, m->d_valid,
                                                       o->rm_dup ? m->d_first : nullptr, *o, sam, report_repeat_hits, n, mates, stride, seq_a, seq_b,
                                                       out_a, out_b, out_pair);
    BSX_CUDA_CHECK(cudaGetLastError());
    if (o->rm_dup) BSX_CUDA_CHECK(cudaEventRecord(m->dup_order, st));
    return BSX_OK;
}

extern "C" int bsx_meth_valid_count(bsx_meth *m, uint64_t *n_valid) {
    if (!m || !n_valid) return BSX_ERR_ARG;
    BSX_CUDA_CHECK(cudaSetDevice(m->ix->device));
    BSX_CUDA_CHECK(cudaDeviceSynchronize());
    unsigned long long v = 0;
    BSX_CUDA_CHECK(cudaMemcpy(&v, m->d_valid, 8, cudaMemcpyDeviceToHost));
    *n_valid = v;
    return BSX_OK;
}

extern "C" int bsx_meth_download(bsx_meth *m, const bsx_meth_opts *o, uint32_t k, uint32_t *meth, uint32_t *depth) {
    if (!m || !o || k >= m->ix->n_seq) { bsx_set_error("bsx_meth_download: bad argument"); return BSX_ERR_ARG; }
    BSX_CUDA_CHECK(cudaSetDevice(m->ix->device));
    int rc = combine_once(m, o); if (rc) return rc;
    const size_t off = m->ix->anchor[k], cnt = m->ix->size[k];
    if (meth) BSX_CUDA_CHECK(cudaMemcpy(meth, m->d_meth + off, cnt * 4, cudaMemcpyDeviceToHost));
    if (depth) BSX_CUDA_CHECK(cudaMemcpy(depth, m->d_depth + off, cnt * 4, cudaMemcpyDeviceToHost));
    return BSX_OK;
}

extern "C" size_t bsx_meth_write(bsx_meth *m, const bsx_meth_opts *o, const char *const *seqs, const uint32_t *lens,
                                 const uint8_t *chroms, int threads, int fd, uint64_t *stats) {
    if (!m || !o || !seqs || !lens) { bsx_set_error("bsx_meth_write: bad argument"); return 0; }
    const bsx_index *ix = m->ix;
    threads = bsx_host_threads(threads);
    size_t written = 0;
    auto out = [&](const std::string &s) { size_t off = 0; while (off < s.size()) { ssize_t w = write(fd, s.data() + off, s.size() - off); if (w <= 0) return; off += (size_t)w; written += (size_t)w; } };
    out("chr\tpos\tstrand\tcontext\tratio\ttotal_C\tmethy_C\tCI_lower\tCI_upper\n");
    std::vector<uint32_t> order;
    for (uint32_t k = 0; k < ix->n_seq; k++) if (!chroms || chroms[k]) order.push_back(k);
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return ix->names[a] < ix->names[b]; });   // sorted(depth.keys())
    uint64_t nc = 0, nd = 0;
    const double z95 = 1.96, z95sq = 1.96 * 1.96;
    std::vector<uint32_t> hm, hd;
    for (uint32_t k : order) {
        const uint32_t L = ix->size[k];
        if (lens[k] != L) { bsx_set_error("bsx_meth_write: sequence %u has %u bases, the index %u", k, lens[k], L); return written; }
        hm.resize(L); hd.resize(L);
        if (bsx_meth_download(m, o, k, hm.data(), hd.data()) != BSX_OK) return written;
        const char *rs = seqs[k];
        const std::string &name = ix->names[k];
        std::vector<std::string> chunk((size_t)threads);
        std::vector<uint64_t> cnc((size_t)threads, 0), cnd((size_t)threads, 0);
        bsx_parallel(threads, L, [&](int t, size_t b, size_t e) {
            Sink s{&chunk[t]};
            uint64_t lc = 0, ld = 0;
            for (size_t i = b; i < e; i++) {
                const uint32_t d = hd[i];
                if ((long long)d < (long long)o->min_depth || d == 0) continue;
                lc++; ld += d;
                const uint32_t mm = hm[i];
                if (mm == 0 && !o->meth0) continue;
                const double ratio = (double)mm / d;
                s.put(name.data(), name.size()); s.putc('\t'); s.putu(i + 1); s.putc('\t');
                const char rc = (char)(rs[i] >= 'a' && rs[i] <= 'z' ? rs[i] - 32 : rs[i]);
                s.putc(rc == 'C' ? '+' : '-'); s.putc('\t');
                if (i >= 2) for (size_t j = i - 2; j < i + 3 && j < L; j++) s.putc((char)(rs[j] >= 'a' && rs[j] <= 'z' ? rs[j] - 32 : rs[j]));   // refcr[i-2:i+3] ('' when i < 2)
                s.putc('\t'); s.putf3(ratio); s.putc('\t'); s.putu(d); s.putc('\t'); s.putu(mm); s.putc('\t');
                const double pmid = ratio + z95sq / (2 * (double)d);
                const double sd = z95 * pow(ratio * (1 - ratio) / d + z95sq / (4 * (double)d * d), 0.5);
                const double nm = 1 + z95sq / d;
                s.putf3((pmid - sd) / nm); s.putc('\t'); s.putf3((pmid + sd) / nm); s.putc('\n');
            }
            cnc[t] = lc; cnd[t] = ld;
        });
        for (int t = 0; t < threads; t++) { out(chunk[t]); nc += cnc[t]; nd += cnd[t]; }
    }
    if (stats) { stats[0] = nc; stats[1] = nd; }
    return written;
}

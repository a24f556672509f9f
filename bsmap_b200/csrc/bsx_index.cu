// bsx_index.cu -- device-resident 2-bit reference and seed table (RefSeq, dbseq.cpp).
//
//   K0  pack_strands_kernel     BinSeq / cBinSeq (dbseq.cpp:58-111): one thread packs one u32 of the
//                               forward strand and one of the reverse-complement strand.
//   K1a seed_items_kernel       t_CalKmerFreq_ab (dbseq.cpp:349-359): one thread per indexed position
//                               computes its 3-letter key (s_MakeSeed_1 + XT), emits (key, position)
//                               and bumps the (key, strand) histogram.
//   K1b exclusive scan          AllocIndex (dbseq.cpp:365-388) becomes CSR offsets: tab[2k] = list
//                               start, tab[2k+1] = start of the rc half, tab[2k+2] = end.
//   K1c stable LSD radix sort   t_CreateIndex_ab (dbseq.cpp:441-480): the reference appends positions
//                               in enumeration order (all forward blocks ascending, then all rc blocks
//                               ascending); a STABLE sort of the enumeration by key reproduces every
//                               list exactly.  Warp-private sub-tiles keep the scatter stable without
//                               any block-level synchronisation.
//
// All kernels are HBM-bound integer work (no tensor cores).  The build is one-time per run
// (the reference spends 250-300 s here for a 3.1 Gb genome).
#include <algorithm>
#include <cstdio>
#include <cstring>
#include "bsx_common.cuh"
#include "bsx_internal.h"

// ------------------------------------------------------------------------------------------ K0
__global__ void pack_strands_kernel(const uint8_t *__restrict__ seq, uint32_t len, uint32_t nwords,
                                    uint32_t *__restrict__ fwd, uint32_t *__restrict__ rc) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nwords) return;
    const uint32_t T = nwords * BSX_SEGLEN;
    uint32_t wf = 0, wc = 0;
#pragma unroll
    for (uint32_t j = 0; j < BSX_SEGLEN; j++) {
        uint32_t pf = i * BSX_SEGLEN + j;          // forward: pad with 'N' -> code 0
        uint32_t pc = T - 1 - pf;                  // rc strand reads the PADDED buffer backwards
        wf = (wf << 2) | (pf < len ? bsx_code_fwd(seq[pf]) : 0u);
        wc = (wc << 2) | (pc < len ? bsx_code_rev(seq[pc]) : 3u);
    }
    fwd[i] = wf;
    rc[i] = wc;
}

// The reverse-complement strand from the packed forward strand (used when the reference comes from a packed cache):
// cBinSeq reads the PADDED forward buffer backwards through rev_alphabet, and both N and padding pack to code 0
// forward / code 3 reverse, so word i of the rc strand is the 2-bit-wise reversed complement of word nwords-1-i.
__global__ void rc_from_packed_kernel(const uint32_t *__restrict__ fwd, uint32_t nwords, uint32_t *__restrict__ rc) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nwords) return;
    uint32_t x = __brev(fwd[nwords - 1 - i]);                       // fields reversed, bits inside each field swapped
    x = ((x & 0x55555555u) << 1) | ((x >> 1) & 0x55555555u);        // swap them back
    rc[i] = ~x;                                                     // complement: code -> 3 - code
}

// ------------------------------------------------------------------------------------------ K1a
struct bsx_dev_block {       // one UnmaskRegion block, ready for enumeration
    uint32_t word_base;      // anchor/16 of its sequence
    uint32_t anchor;         // ref_anchor of its sequence
    uint32_t i0;             // (begin / I) * I
    uint32_t strand;         // 0 forward, 1 rc
};

__device__ __forceinline__ uint32_t seed_key_at(const uint32_t *__restrict__ m, uint32_t p, int s, uint32_t seed_bits) {
    // RefSeq::s_MakeSeed_1 (dbseq.cpp:286-291)
    const uint32_t *w = m + (p >> 4);
    int a = 64 - 2 * s - 2 * (int)(p & 15u);
    uint64_t v = (((uint64_t)w[0] << 32) | w[1]) >> a;
    return bsx_xt((uint32_t)v & seed_bits, s);
}

__global__ void seed_items_kernel(const uint32_t *__restrict__ refcat, const uint32_t *__restrict__ crefcat,
                                  const bsx_dev_block *__restrict__ blocks, const uint64_t *__restrict__ prefix,
                                  uint32_t n_blocks, uint64_t n_items, int s, int I, uint32_t seed_bits,
                                  uint64_t *__restrict__ items, uint32_t *__restrict__ hist) {
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_items) return;
    uint32_t lo = 0, hi = n_blocks;            // last block with prefix[b] <= e
    while (lo + 1 < hi) { uint32_t mid = (lo + hi) >> 1; if (prefix[mid] <= e) lo = mid; else hi = mid; }
    const bsx_dev_block b = blocks[lo];
    uint32_t p = b.i0 + (uint32_t)(e - prefix[lo]) * (uint32_t)I;
    const uint32_t *m = (b.strand ? crefcat : refcat) + b.word_base;
    uint32_t key = seed_key_at(m, p, s, seed_bits);
    // hit2int (dbseq.cpp:570); the strand rides along in bit 63, above every radix digit
    items[e] = ((uint64_t)b.strand << 63) | ((uint64_t)key << 32) | (uint64_t)(b.anchor + p);
    atomicAdd(&hist[2 * key + b.strand], 1u);
}

// RRBS: entries are enumerated on the host; the device computes their keys.  items = (key << 32) | entry index.
// The reference appends to a key's list segment by segment, sequence by sequence, plain entries before mirrored ones
// (dbseq.cpp:418-438), and SnpAlign walks the whole list skipping every entry whose (segment, mirror) tag is not the
// one the mode wants (align.cpp:187, 229).  Here the host enumerates group-major -- (segment, mirror), then sequence,
// then site -- so the stable sort by key leaves every list partitioned into its groups, each in the reference's
// order, and the table is a CSR over (key, group): a mode reads exactly the entries the reference would not skip.
__global__ void rrbs_items_kernel(const uint32_t *__restrict__ refcat, const uint32_t *__restrict__ crefcat,
                                  const uint32_t *__restrict__ seqinfo, const uint32_t *__restrict__ loc,
                                  const uint32_t *__restrict__ tag, uint64_t n_items, int s, uint32_t seed_bits,
                                  uint32_t groups, uint64_t *__restrict__ items, uint32_t *__restrict__ hist) {
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_items) return;
    const uint32_t t = tag[e], chr = t & 0xffffu;
    const uint32_t *m = ((chr & 1u) ? crefcat : refcat) + (seqinfo[chr >> 1] >> 4);
    uint32_t key = seed_key_at(m, loc[e], s, seed_bits);
    items[e] = ((uint64_t)key << 32) | e;
    atomicAdd(&hist[(uint64_t)key * groups + (2u * ((t >> 16) & 0xffu) + (t >> 24))], 1u);
}

// sorted order -> pos / tag, plus the inline context of every entry (the 16 bases before and the 16 after the seed on
// the entry's strand), exactly as the WGBS table carries it
__global__ void rrbs_gather_kernel(const uint32_t *__restrict__ order, const uint32_t *__restrict__ loc,
                                   const uint32_t *__restrict__ tag, uint64_t n, const uint32_t *__restrict__ refcat,
                                   const uint32_t *__restrict__ crefcat, const uint32_t *__restrict__ seqinfo, int seed_size,
                                   uint32_t *__restrict__ out_loc, uint32_t *__restrict__ out_tag, uint2 *__restrict__ out_ctx) {
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    uint32_t src = order[e];
    const uint32_t l = loc[src], t = tag[src], chr = t & 0xffffu;
    out_loc[e] = l;
    out_tag[e] = t;
    const uint32_t *m = (chr & 1u) ? crefcat : refcat;
    const uint32_t c = seqinfo[chr >> 1] + l, bb = c - 16u, aa = c + (uint32_t)seed_size;
    out_ctx[e] = make_uint2(__funnelshift_l(m[(bb >> 4) + 1], m[bb >> 4], (bb & 15u) * 2u),
                            __funnelshift_l(m[(aa >> 4) + 1], m[aa >> 4], (aa & 15u) * 2u));
}

// ------------------------------------------------------------------------------------------ scan
// exclusive prefix sum over u32 (n up to 2^32), three kernels: tile sums, scan of sums, rescan.
#define SCAN_THREADS 256
#define SCAN_ITEMS 16
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total, uint32_t *sh /*[SCAN_THREADS/32]*/) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(BSX_FULL, x, d); if (lane >= d) x += y; }
    if (lane == 31) sh[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = lane < SCAN_THREADS / 32 ? sh[lane] : 0, t = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(BSX_FULL, t, d); if (lane >= d) t += y; }
        if (lane < SCAN_THREADS / 32) sh[lane] = t - s;          // exclusive warp offsets
        if (lane == SCAN_THREADS / 32 - 1) sh[SCAN_THREADS / 32] = t;   // block total
    }
    __syncthreads();
    uint32_t r = x - v + sh[wid];
    *total = sh[SCAN_THREADS / 32];
    __syncthreads();
    return r;
}

__global__ void scan_tile_sums_kernel(const uint32_t *__restrict__ in, uint64_t n, uint32_t *__restrict__ sums) {
    __shared__ uint32_t sh[SCAN_THREADS / 32 + 1];
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) if (base + k < n) s += in[base + k];
    uint32_t total;
    block_exclusive_scan(s, &total, sh);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void scan_sums_kernel(uint32_t *sums, uint32_t n_tiles) {   // single block
    __shared__ uint32_t sh[SCAN_THREADS / 32 + 1];
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n_tiles; base += SCAN_THREADS) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < n_tiles ? sums[i] : 0, total;
        uint32_t ex = block_exclusive_scan(v, &total, sh);
        if (i < n_tiles) sums[i] = ex + carry;
        carry += total;
    }
}

__global__ void scan_apply_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint64_t n,
                                  const uint32_t *__restrict__ sums) {
    __shared__ uint32_t sh[SCAN_THREADS / 32 + 1];
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = (base + k < n) ? in[base + k] : 0; s += v[k]; }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(s, &total, sh) + sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) out[base + k] = ex; ex += v[k]; }
}

static int exclusive_scan_u32(const uint32_t *d_in, uint32_t *d_out, uint64_t n, cudaStream_t st) {
    if (n == 0) return BSX_OK;
    uint64_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    uint32_t *d_sums = nullptr;
    BSX_CUDA_CHECK(cudaMalloc(&d_sums, n_tiles * sizeof(uint32_t)));
    scan_tile_sums_kernel<<<(unsigned)n_tiles, SCAN_THREADS, 0, st>>>(d_in, n, d_sums);
    scan_sums_kernel<<<1, SCAN_THREADS, 0, st>>>(d_sums, (uint32_t)n_tiles);
    scan_apply_kernel<<<(unsigned)n_tiles, SCAN_THREADS, 0, st>>>(d_in, d_out, n, d_sums);
    BSX_CUDA_CHECK(cudaGetLastError());
    BSX_CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFree(d_sums);
    return BSX_OK;
}

// ------------------------------------------------------------------------------------------ K1c
// Stable LSD radix sort of 64-bit items on bits [32+shift, 32+shift+bits).  Every warp owns one
// contiguous sub-tile of SUB_ITEMS items and walks it in order, 32 items at a time, so the scatter
// is stable by construction: rank within the 32 via match_any, running per-digit cursors private to
// the warp in shared memory.
#define RS_MAX_BITS 9
#define RS_WARPS 4
#define RS_SUB_ITEMS 8192u

__global__ void __launch_bounds__(RS_WARPS * 32)
radix_hist_kernel(const uint64_t *__restrict__ in, uint64_t n, int shift, int bits, uint32_t n_sub,
                  uint32_t *__restrict__ hist /* [digit][sub] */) {
    __shared__ uint32_t sh[RS_WARPS][1 << RS_MAX_BITS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t sub = blockIdx.x * RS_WARPS + wid;
    const uint32_t nd = 1u << bits;
    for (uint32_t d = lane; d < nd; d += 32) sh[wid][d] = 0;
    __syncwarp();
    if (sub < n_sub) {
        uint64_t b = (uint64_t)sub * RS_SUB_ITEMS, e = b + RS_SUB_ITEMS < n ? b + RS_SUB_ITEMS : n;
        for (uint64_t i0 = b; i0 < e; i0 += 32) {
            const uint64_t i = i0 + lane;
            const bool act = i < e;
            uint32_t d = act ? ((uint32_t)(in[i] >> (32 + shift)) & (nd - 1)) : 0xffffffffu;
            unsigned peers = __match_any_sync(BSX_FULL, d);
            if (act && lane == __ffs(peers) - 1) sh[wid][d] += __popc(peers);
            __syncwarp();
        }
        __syncwarp();
        for (uint32_t d = lane; d < nd; d += 32) hist[(uint64_t)d * n_sub + sub] = sh[wid][d];
    }
}

template <typename OutT>
__global__ void __launch_bounds__(RS_WARPS * 32)
radix_scatter_kernel(const uint64_t *__restrict__ in, uint64_t n, int shift, int bits, uint32_t n_sub,
                     const uint32_t *__restrict__ offs /* exclusive scan of hist */, OutT *__restrict__ out) {
    __shared__ uint32_t sh[RS_WARPS][1 << RS_MAX_BITS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t sub = blockIdx.x * RS_WARPS + wid;
    const uint32_t nd = 1u << bits;
    if (sub >= n_sub) return;
    for (uint32_t d = lane; d < nd; d += 32) sh[wid][d] = offs[(uint64_t)d * n_sub + sub];
    __syncwarp();
    uint64_t b = (uint64_t)sub * RS_SUB_ITEMS, e = b + RS_SUB_ITEMS < n ? b + RS_SUB_ITEMS : n;
    for (uint64_t i0 = b; i0 < e; i0 += 32) {
        uint64_t i = i0 + lane;
        const bool act = i < e;
        uint64_t it = act ? in[i] : 0;
        uint32_t d = act ? ((uint32_t)(it >> (32 + shift)) & (nd - 1)) : 0xffffffffu;   // inactive lanes never match a digit
        unsigned peers = __match_any_sync(BSX_FULL, d);
        uint32_t rank = __popc(peers & ((1u << lane) - 1));
        uint32_t dst = 0;
        if (act) dst = sh[wid][d] + rank;
        __syncwarp();
        if (act && rank == 0) sh[wid][d] += __popc(peers);
        __syncwarp();
        if (act) out[dst] = (OutT)it;      // OutT = uint32_t keeps the low word (the position) on the last pass
    }
}

// Last pass for the WGBS table: besides the position it stores, next to every entry, the 16 reference
// bases that precede the seed and the 16 that follow it on the entry's strand ("inline context").
// The mapping kernel rejects >98 % of candidates from these 8 bytes, which arrive with the coalesced
// list stream, instead of paying one random HBM access per candidate.
__global__ void __launch_bounds__(RS_WARPS * 32)
radix_scatter_ctx_kernel(const uint64_t *__restrict__ in, uint64_t n, int shift, int bits, uint32_t n_sub,
                         const uint32_t *__restrict__ offs, const uint32_t *__restrict__ refcat,
                         const uint32_t *__restrict__ crefcat, int seed_size, uint32_t *__restrict__ out_pos,
                         uint2 *__restrict__ out_ctx, uint4 *__restrict__ out_ctx16) {
    __shared__ uint32_t sh[RS_WARPS][1 << RS_MAX_BITS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t sub = blockIdx.x * RS_WARPS + wid;
    const uint32_t nd = 1u << bits;
    if (sub >= n_sub) return;
    for (uint32_t d = lane; d < nd; d += 32) sh[wid][d] = offs[(uint64_t)d * n_sub + sub];
    __syncwarp();
    uint64_t b = (uint64_t)sub * RS_SUB_ITEMS, e = b + RS_SUB_ITEMS < n ? b + RS_SUB_ITEMS : n;
    for (uint64_t i0 = b; i0 < e; i0 += 32) {
        uint64_t i = i0 + lane;
        const bool act = i < e;
        uint64_t it = act ? in[i] : 0;
        uint32_t d = act ? ((uint32_t)(it >> (32 + shift)) & (nd - 1)) : 0xffffffffu;
        unsigned peers = __match_any_sync(BSX_FULL, d);
        uint32_t rank = __popc(peers & ((1u << lane) - 1));
        uint32_t dst = 0;
        if (act) dst = sh[wid][d] + rank;
        __syncwarp();
        if (act && rank == 0) sh[wid][d] += __popc(peers);
        __syncwarp();
        if (act) {
            const uint32_t c = (uint32_t)it;
            const uint32_t *m = (it >> 63) ? crefcat : refcat;
            const uint32_t bb = c - 16u, aa = c + (uint32_t)seed_size;
            const uint32_t before = __funnelshift_l(m[(bb >> 4) + 1], m[bb >> 4], (bb & 15u) * 2u);
            const uint32_t after = __funnelshift_l(m[(aa >> 4) + 1], m[aa >> 4], (aa & 15u) * 2u);
            out_pos[dst] = c;
            if (!out_ctx16) out_ctx[dst] = make_uint2(before, after);
            else {            // indexes built for -v >= 8: the next 16 bases outwards as well, [c-32, c-16) and [c+s+16, c+s+32)
                const uint32_t b2 = c - 32u, a2 = aa + 16u;
                out_ctx16[dst] = make_uint4(__funnelshift_l(m[(b2 >> 4) + 1], m[b2 >> 4], (b2 & 15u) * 2u), before, after,
                                            __funnelshift_l(m[(a2 >> 4) + 1], m[a2 >> 4], (a2 & 15u) * 2u));
            }
        }
    }
}

struct bsx_ctx_out { const uint32_t *refcat, *crefcat; int seed_size; uint2 *ctx; uint4 *ctx16; };

static int radix_sort_items(uint64_t *d_a, uint64_t *d_b, uint64_t n, int key_bits, uint32_t *d_out32, cudaStream_t st,
                            const bsx_ctx_out *cx = nullptr) {
    // sorts by the key held in the high word; final pass writes the low word to d_out32
    int passes = (key_bits + RS_MAX_BITS - 1) / RS_MAX_BITS;
    if (passes < 1) passes = 1;
    int bits = (key_bits + passes - 1) / passes;
    if (bits < 1) bits = 1;
    uint32_t n_sub = (uint32_t)((n + RS_SUB_ITEMS - 1) / RS_SUB_ITEMS);
    uint64_t n_hist = (uint64_t)n_sub << bits;
    uint32_t *d_hist = nullptr;
    BSX_CUDA_CHECK(cudaMalloc(&d_hist, n_hist * sizeof(uint32_t)));
    unsigned grid = (n_sub + RS_WARPS - 1) / RS_WARPS;
    uint64_t *src = d_a, *dst = d_b;
    for (int p = 0; p < passes; p++) {
        int shift = p * bits;
        radix_hist_kernel<<<grid, RS_WARPS * 32, 0, st>>>(src, n, shift, bits, n_sub, d_hist);
        BSX_CUDA_CHECK(cudaGetLastError());
        int rc = exclusive_scan_u32(d_hist, d_hist, n_hist, st);
        if (rc) { cudaFree(d_hist); return rc; }
        if (p == passes - 1 && cx)
            radix_scatter_ctx_kernel<<<grid, RS_WARPS * 32, 0, st>>>(src, n, shift, bits, n_sub, d_hist, cx->refcat, cx->crefcat,
                                                                      cx->seed_size, d_out32, cx->ctx, cx->ctx16);
        else if (p == passes - 1)
            radix_scatter_kernel<uint32_t><<<grid, RS_WARPS * 32, 0, st>>>(src, n, shift, bits, n_sub, d_hist, d_out32);
        else
            radix_scatter_kernel<uint64_t><<<grid, RS_WARPS * 32, 0, st>>>(src, n, shift, bits, n_sub, d_hist, dst);
        BSX_CUDA_CHECK(cudaGetLastError());
        std::swap(src, dst);
    }
    BSX_CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFree(d_hist);
    return BSX_OK;
}

// ------------------------------------------------------------------------------------------ host
static const uint8_t *class_table() {
    // 1 = ACGTacgt (useful_nt), 2 = NXnx (nx_nt), 0 = anything else
    static uint8_t t[256]; static bool init = false;
    if (!init) {
        memset(t, 0, sizeof t);
        for (const char *c = "ACGTacgt"; *c; c++) t[(uint8_t)*c] = 1;
        for (const char *c = "NXnx"; *c; c++) t[(uint8_t)*c] = 2;
        init = true;
    }
    return t;
}

// UnmaskRegion (dbseq.cpp:114-142): maximal runs that start at the first ACGTacgt, end at the first
// NXnx (or the sequence end), kept when >= 30 nt; never merged (the merge test is dead code, it
// compares against the rc block pushed just before).
// The scan is a two-state machine driven only by the class-1 (start) and class-2 (end) characters, so a sequence is cut
// into pieces scanned on all host threads -- each keeps the positions where the class changes -- and the pieces' short
// event lists are stitched in order.
static void unmask_region(const char *seq, uint32_t len, uint32_t id, uint32_t T, std::vector<bsx_block> &out, int threads) {
    const uint8_t *cls = class_table();
    const uint8_t *s = (const uint8_t *)seq;
    struct Ev { uint32_t pos; uint8_t c; };
    const size_t pieces = std::max<size_t>(1, std::min<size_t>((size_t)threads * 4, len / (1u << 20)));
    std::vector<std::vector<Ev>> ev(pieces);
    bsx_parallel(threads, pieces, [&](int, size_t pb, size_t pe) {
        for (size_t pc = pb; pc < pe; pc++) {
            const uint32_t b = (uint32_t)((uint64_t)len * pc / pieces), e = (uint32_t)((uint64_t)len * (pc + 1) / pieces);
            uint8_t last = 0;
            for (uint32_t q = b; q < e; q++) { const uint8_t c = cls[s[q]]; if (c != 0 && c != last) { ev[pc].push_back({q, c}); last = c; } }
        }
    });
    bool inside = false; uint32_t b = 0;
    auto close = [&](uint32_t e) { if (e - b >= 30) { out.push_back({id, b, e}); out.push_back({id + 1, T - e, T - b}); } };
    for (const std::vector<Ev> &v : ev)
        for (const Ev &x : v) {
            if (x.c == 1 && !inside) { inside = true; b = x.pos; }
            else if (x.c == 2 && inside) { inside = false; close(x.pos); }
        }
    if (inside) close(len);
}

void bsx_index_free_device(bsx_index *ix) {
    if (ix->device >= 0) cudaSetDevice(ix->device);
    cudaFree(ix->d_refcat); cudaFree(ix->d_crefcat); cudaFree(ix->d_tab); cudaFree(ix->d_pos);
    cudaFree(ix->d_tag); cudaFree(ix->d_seqinfo); cudaFree(ix->d_sites); cudaFree(ix->d_site_off); cudaFree(ix->d_ctx);
    ix->d_refcat = ix->d_crefcat = ix->d_tab = ix->d_pos = ix->d_tag = ix->d_seqinfo = ix->d_sites = ix->d_site_off = nullptr;
    ix->d_ctx = nullptr;
}

static int upload_seqinfo(bsx_index *ix) {
    std::vector<uint32_t> h(3 * (size_t)ix->n_seq + 1);
    for (uint32_t k = 0; k <= ix->n_seq; k++) h[k] = ix->anchor[k];
    for (uint32_t k = 0; k < ix->n_seq; k++) { h[ix->n_seq + 1 + k] = ix->size[k]; h[2 * ix->n_seq + 1 + k] = ix->rc_offset[k]; }
    BSX_CUDA_CHECK(cudaMalloc(&ix->d_seqinfo, h.size() * 4));
    BSX_CUDA_CHECK(cudaMemcpy(ix->d_seqinfo, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    if (ix->par.rrbs) {
        std::vector<uint32_t> off(ix->n_seq + 1, 0), all;
        for (uint32_t k = 0; k < ix->n_seq; k++) { off[k + 1] = off[k] + (uint32_t)ix->sites[k].size(); all.insert(all.end(), ix->sites[k].begin(), ix->sites[k].end()); }
        BSX_CUDA_CHECK(cudaMalloc(&ix->d_site_off, off.size() * 4));
        BSX_CUDA_CHECK(cudaMemcpy(ix->d_site_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
        BSX_CUDA_CHECK(cudaMalloc(&ix->d_sites, (all.size() + 1) * 4));
        if (!all.empty()) BSX_CUDA_CHECK(cudaMemcpy(ix->d_sites, all.data(), all.size() * 4, cudaMemcpyHostToDevice));
    }
    return BSX_OK;
}

// allocate the device arrays of a replica whose metadata is already filled in
int bsx_index_alloc_device(bsx_index *ix) {
    BSX_CUDA_CHECK(cudaSetDevice(ix->device));
    BSX_CUDA_CHECK(cudaMalloc(&ix->d_refcat, ix->n_words * 4));
    BSX_CUDA_CHECK(cudaMalloc(&ix->d_crefcat, ix->n_words * 4));
    BSX_CUDA_CHECK(cudaMalloc(&ix->d_tab, bsx_tab_len(ix) * 4));
    BSX_CUDA_CHECK(cudaMalloc(&ix->d_pos, (ix->n_entries + 64) * 4));
    ix->ctx_wide = !ix->par.rrbs && ix->par.max_snp_num >= BSX_WIDE_CTX_V;
    BSX_CUDA_CHECK(cudaMalloc(&ix->d_ctx, (ix->n_entries + 64) * bsx_ctx_entry_bytes(ix)));
    if (ix->par.rrbs) BSX_CUDA_CHECK(cudaMalloc(&ix->d_tag, (ix->n_entries + 64) * 4));
    return upload_seqinfo(ix);
}

// seqs: the sequences as ASCII, or NULL when `packed` (host copy of the whole packed forward strand, n_words words
// with margins) and ix->blocks are given instead (bsx_index_create_from_packed; WGBS only)
int bsx_index_build_device(bsx_index *ix, const char *const *seqs, const uint32_t *packed) {
    const bsx_params &p = ix->par;
    const int s = p.seed_size, I = p.index_interval;
    BSX_CUDA_CHECK(cudaSetDevice(ix->device));
    cudaStream_t st;
    BSX_CUDA_CHECK(cudaStreamCreate(&st));
    cudaEvent_t ev0, ev1;
    cudaEventCreate(&ev0); cudaEventCreate(&ev1);

    // --- geometry (Run_ConvertBinseq, dbseq.cpp:215-282)
    uint64_t tot = 0; uint32_t max_len = 0;
    ix->nwords.resize(ix->n_seq); ix->rc_offset.resize(ix->n_seq); ix->anchor.resize(ix->n_seq + 1);
    for (uint32_t k = 0; k < ix->n_seq; k++) {
        ix->nwords[k] = (ix->size[k] + BSX_SEGLEN - 1) / BSX_SEGLEN + 2;
        ix->rc_offset[k] = ix->nwords[k] * BSX_SEGLEN;
        ix->anchor[k] = (uint32_t)((tot + BSX_REF_MARGIN) * BSX_SEGLEN);
        tot += ix->nwords[k];
        max_len = std::max(max_len, ix->size[k]);
    }
    if ((tot + 2 * BSX_REF_MARGIN) * BSX_SEGLEN >= (1ull << 32)) {
        bsx_set_error("reference too large for 32-bit coordinates (ref_loc_t, param.h:36)");
        return BSX_ERR_ARG;
    }
    ix->anchor[ix->n_seq] = (uint32_t)((tot + BSX_REF_MARGIN) * BSX_SEGLEN);
    ix->n_words = tot + 2 * BSX_REF_MARGIN;
    ix->n_keys = 1; for (int i = 0; i < s; i++) ix->n_keys *= 3;
    const uint32_t seed_bits = (s == 16) ? 0xffffffffu : ((1u << (2 * s)) - 1);

    BSX_CUDA_CHECK(cudaMalloc(&ix->d_refcat, ix->n_words * 4));
    BSX_CUDA_CHECK(cudaMalloc(&ix->d_crefcat, ix->n_words * 4));
    BSX_CUDA_CHECK(cudaMemsetAsync(ix->d_refcat, 0, ix->n_words * 4, st));    // margins defined as zero (App. B Q5)
    BSX_CUDA_CHECK(cudaMemsetAsync(ix->d_crefcat, 0, ix->n_words * 4, st));

    // --- blocks (UnmaskRegion): a host scan of the ASCII on all threads, while the sequences go to the device below
    std::vector<bsx_block> &blocks = ix->blocks;
    std::thread block_scan;
    if (!p.rrbs && !packed && !ix->ref_only) {
        blocks.clear();
        block_scan = std::thread([ix, seqs, &blocks] {
            const int threads = bsx_host_threads(0);
            for (uint32_t k = 0; k < ix->n_seq; k++) unmask_region(seqs[k], ix->size[k], 2 * k, ix->rc_offset[k], blocks, threads);
            std::stable_sort(blocks.begin(), blocks.end(), [](const bsx_block &a, const bsx_block &b) {
                return a.id < b.id || (a.id == b.id && a.begin < b.begin); });
        });
    }
    struct Joiner { std::thread &t; ~Joiner() { if (t.joinable()) t.join(); } } block_scan_joiner{block_scan};   // error returns below

    // --- K0: pack every sequence (ASCII staged through one device buffer), or take the packed forward strand as is
    cudaEventRecord(ev0, st);
    float ms_total = 0;
    if (packed) {
        if (p.rrbs) { bsx_set_error("a packed reference cache cannot seed an RRBS index (digestion sites need the text)"); return BSX_ERR_ARG; }
        BSX_CUDA_CHECK(cudaMemcpyAsync(ix->d_refcat, packed, ix->n_words * 4, cudaMemcpyHostToDevice, st));
        for (uint32_t k = 0; k < ix->n_seq; k++) {
            const uint32_t nw = ix->nwords[k];
            rc_from_packed_kernel<<<(nw + 255) / 256, 256, 0, st>>>(ix->d_refcat + (ix->anchor[k] >> 4), nw, ix->d_crefcat + (ix->anchor[k] >> 4));
        }
        BSX_CUDA_CHECK(cudaGetLastError());
    } else {
        uint8_t *d_seq = nullptr;
        BSX_CUDA_CHECK(cudaMalloc(&d_seq, (size_t)max_len + 64));
        for (uint32_t k = 0; k < ix->n_seq; k++) {
            BSX_CUDA_CHECK(cudaMemcpyAsync(d_seq, seqs[k], ix->size[k], cudaMemcpyHostToDevice, st));
            uint32_t nw = ix->nwords[k];
            pack_strands_kernel<<<(nw + 255) / 256, 256, 0, st>>>(d_seq, ix->size[k], nw, ix->d_refcat + (ix->anchor[k] >> 4),
                                                                 ix->d_crefcat + (ix->anchor[k] >> 4));
            BSX_CUDA_CHECK(cudaGetLastError());
            BSX_CUDA_CHECK(cudaStreamSynchronize(st));   // d_seq is reused
        }
        cudaFree(d_seq);
    }
    if (ix->ref_only) {   // packed reference only (bsx_index_create_packed): no seed table
        BSX_CUDA_CHECK(cudaStreamSynchronize(st));
        cudaEventDestroy(ev0); cudaEventDestroy(ev1);
        cudaStreamDestroy(st);
        return upload_seqinfo(ix);
    }

    // --- RRBS sites (find_CCGG)
    std::vector<uint32_t> rr_loc, rr_tag;
    if (block_scan.joinable()) block_scan.join();
    if (p.rrbs) {
        const int max_seg = (BSX_FIXWORDS - 1) * 16 / s;      // dbseq.cpp:217
        const int sl = (int)strlen(p.digest_site), dp = p.digest_pos;
        const bool mirror = p.pairend || p.chains;
        ix->sites.assign(ix->n_seq, {});
        std::vector<std::vector<uint32_t>> cidx((size_t)max_seg * 2 * ix->n_seq);
        for (uint32_t k = 0; k < ix->n_seq; k++) {
            const uint8_t *sq = (const uint8_t *)seqs[k]; const uint32_t len = ix->size[k];
            std::vector<uint32_t> &stv = ix->sites[k];
            for (uint32_t q = 0; q + sl <= len; q++) {
                bool ok = true;
                for (int t = 0; t < sl; t++) { uint8_t c = sq[q + t]; if (c >= 'a' && c <= 'z') c -= 32; if (c != (uint8_t)p.digest_site[t]) { ok = false; break; } }
                if (ok) stv.push_back(q + dp);
            }
            const uint32_t tmp_offset = ix->rc_offset[k] - s, tmp_max = len - s;
            for (size_t q = 0; q + 1 < stv.size(); q++)
                if (stv[q + 1] - stv[q] <= (uint32_t)p.max_insert) {
                    uint32_t seedloc = stv[q];
                    for (int i = 0; i < max_seg && seedloc <= tmp_max; i++, seedloc += s) cidx[(size_t)i * 2 * ix->n_seq + 2 * k].push_back(seedloc);
                }
            for (size_t q = 1; q < stv.size(); q++)
                if (stv[q] - stv[q - 1] <= (uint32_t)p.max_insert) {
                    int seedloc = (int)(stv[q] + sl - 2 * dp - s);
                    for (int i = 0; i < max_seg && seedloc >= 0; i++, seedloc -= s) cidx[(size_t)i * 2 * ix->n_seq + 2 * k + 1].push_back(tmp_offset - (uint32_t)seedloc);
                }
        }
        // group-major enumeration (see rrbs_items_kernel): the reference's order is j, chr, {plain, mirrored}
        for (int j = 0; j < max_seg; j++)
            for (int flag = 0; flag < (mirror ? 2 : 1); flag++)
                for (uint32_t chr = 0; chr < 2 * ix->n_seq; chr++) {
                    if (!flag) {
                        for (uint32_t v : cidx[(size_t)j * 2 * ix->n_seq + chr]) { rr_loc.push_back(v); rr_tag.push_back(chr | ((uint32_t)j << 16)); }
                    } else {
                        const uint32_t tmp_offset = ix->rc_offset[chr >> 1] - s;
                        for (uint32_t v : cidx[(size_t)j * 2 * ix->n_seq + (chr ^ 1)]) { rr_loc.push_back(tmp_offset - v); rr_tag.push_back(chr | ((uint32_t)j << 16) | 0x1000000u); }
                    }
                }
    }
    if (upload_seqinfo(ix)) return BSX_ERR_CUDA;

    // --- enumeration: forward blocks first, then rc blocks (t_CreateIndex_ab)
    std::vector<bsx_dev_block> hb; std::vector<uint64_t> prefix;
    uint64_t n_items = 0;
    if (!p.rrbs) {
        for (int pass = 0; pass < 2; pass++)
            for (const bsx_block &b : blocks) {
                if ((int)(b.id & 1) != pass) continue;
                uint32_t i0 = (b.begin / I) * I, i2 = ((b.end - s) / I) * I;
                if (i2 < i0) continue;
                hb.push_back({ix->anchor[b.id >> 1] >> 4, ix->anchor[b.id >> 1], i0, b.id & 1});
                prefix.push_back(n_items);
                n_items += (uint64_t)(i2 - i0) / I + 1;
            }
    } else n_items = rr_loc.size();
    if (n_items >= (1ull << 32)) { bsx_set_error("seed table exceeds 2^32 entries"); return BSX_ERR_ARG; }
    ix->n_entries = n_items;

    BSX_CUDA_CHECK(cudaMalloc(&ix->d_tab, bsx_tab_len(ix) * 4));
    BSX_CUDA_CHECK(cudaMemsetAsync(ix->d_tab, 0, bsx_tab_len(ix) * 4, st));
    BSX_CUDA_CHECK(cudaMalloc(&ix->d_pos, (n_items + 64) * 4));
    BSX_CUDA_CHECK(cudaMemsetAsync(ix->d_pos, 0, (n_items + 64) * 4, st));

    if (n_items) {
        uint64_t *d_a = nullptr, *d_b = nullptr;
        BSX_CUDA_CHECK(cudaMalloc(&d_a, n_items * 8));
        BSX_CUDA_CHECK(cudaMalloc(&d_b, n_items * 8));
        const unsigned grid = (unsigned)((n_items + 255) / 256);
        uint32_t *d_rl = nullptr, *d_rt = nullptr;
        if (!p.rrbs) {
            bsx_dev_block *d_blocks = nullptr; uint64_t *d_prefix = nullptr;
            BSX_CUDA_CHECK(cudaMalloc(&d_blocks, hb.size() * sizeof(bsx_dev_block)));
            BSX_CUDA_CHECK(cudaMalloc(&d_prefix, prefix.size() * 8));
            BSX_CUDA_CHECK(cudaMemcpyAsync(d_blocks, hb.data(), hb.size() * sizeof(bsx_dev_block), cudaMemcpyHostToDevice, st));
            BSX_CUDA_CHECK(cudaMemcpyAsync(d_prefix, prefix.data(), prefix.size() * 8, cudaMemcpyHostToDevice, st));
            seed_items_kernel<<<grid, 256, 0, st>>>(ix->d_refcat, ix->d_crefcat, d_blocks, d_prefix, (uint32_t)hb.size(), n_items,
                                                    s, I, seed_bits, d_a, ix->d_tab);
            BSX_CUDA_CHECK(cudaGetLastError());
            BSX_CUDA_CHECK(cudaStreamSynchronize(st));
            cudaFree(d_blocks); cudaFree(d_prefix);
        } else {
            BSX_CUDA_CHECK(cudaMalloc(&d_rl, n_items * 4));
            BSX_CUDA_CHECK(cudaMalloc(&d_rt, n_items * 4));
            BSX_CUDA_CHECK(cudaMemcpyAsync(d_rl, rr_loc.data(), n_items * 4, cudaMemcpyHostToDevice, st));
            BSX_CUDA_CHECK(cudaMemcpyAsync(d_rt, rr_tag.data(), n_items * 4, cudaMemcpyHostToDevice, st));
            rrbs_items_kernel<<<grid, 256, 0, st>>>(ix->d_refcat, ix->d_crefcat, ix->d_seqinfo, d_rl, d_rt, n_items, s, seed_bits,
                                                    bsx_rrbs_groups(s), d_a, ix->d_tab);
            BSX_CUDA_CHECK(cudaGetLastError());
        }
        // histogram -> CSR offsets (in place; the last slot becomes the total)
        int rc = exclusive_scan_u32(ix->d_tab, ix->d_tab, bsx_tab_len(ix), st);
        if (rc) return rc;
        int key_bits = 1; while ((1ull << key_bits) < ix->n_keys) key_bits++;
        if (!p.rrbs) {
            ix->ctx_wide = p.max_snp_num >= BSX_WIDE_CTX_V;
            BSX_CUDA_CHECK(cudaMalloc(&ix->d_ctx, (n_items + 64) * bsx_ctx_entry_bytes(ix)));
            BSX_CUDA_CHECK(cudaMemsetAsync(ix->d_ctx, 0, (n_items + 64) * bsx_ctx_entry_bytes(ix), st));
            bsx_ctx_out cx = {ix->d_refcat, ix->d_crefcat, s, ix->ctx_wide ? nullptr : (uint2 *)ix->d_ctx, ix->ctx_wide ? (uint4 *)ix->d_ctx : nullptr};
            rc = radix_sort_items(d_a, d_b, n_items, key_bits, ix->d_pos, st, &cx);
            if (rc) return rc;
        } else {
            uint32_t *d_order = nullptr;
            BSX_CUDA_CHECK(cudaMalloc(&d_order, n_items * 4));
            rc = radix_sort_items(d_a, d_b, n_items, key_bits, d_order, st);
            if (rc) return rc;
            BSX_CUDA_CHECK(cudaMalloc(&ix->d_tag, (n_items + 64) * 4));
            BSX_CUDA_CHECK(cudaMemsetAsync(ix->d_tag, 0, (n_items + 64) * 4, st));
            BSX_CUDA_CHECK(cudaMalloc(&ix->d_ctx, (n_items + 64) * sizeof(uint2)));
            BSX_CUDA_CHECK(cudaMemsetAsync(ix->d_ctx, 0, (n_items + 64) * sizeof(uint2), st));
            rrbs_gather_kernel<<<grid, 256, 0, st>>>(d_order, d_rl, d_rt, n_items, ix->d_refcat, ix->d_crefcat, ix->d_seqinfo, s,
                                                     ix->d_pos, ix->d_tag, (uint2 *)ix->d_ctx);
            BSX_CUDA_CHECK(cudaGetLastError());
            BSX_CUDA_CHECK(cudaStreamSynchronize(st));
            cudaFree(d_order); cudaFree(d_rl); cudaFree(d_rt);
        }
        cudaFree(d_a); cudaFree(d_b);
    } else {
        if (p.rrbs) { BSX_CUDA_CHECK(cudaMalloc(&ix->d_tag, 64 * 4)); BSX_CUDA_CHECK(cudaMemsetAsync(ix->d_tag, 0, 64 * 4, st)); }
        ix->ctx_wide = !p.rrbs && p.max_snp_num >= BSX_WIDE_CTX_V;
        BSX_CUDA_CHECK(cudaMalloc(&ix->d_ctx, 64 * 16));
        BSX_CUDA_CHECK(cudaMemsetAsync(ix->d_ctx, 0, 64 * 16, st));
    }
    cudaEventRecord(ev1, st);
    BSX_CUDA_CHECK(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&ms_total, ev0, ev1);
    ix->build_seconds = ms_total * 1e-3;
    cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    cudaStreamDestroy(st);
    return BSX_OK;
}

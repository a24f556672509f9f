// bsx_map_impl.cuh -- SingleAlign / PairAlign on the device (align.cpp, align.h, pairs.cpp).
// Included by bsx_map_se.cu (everything inlined: fastest for the single-end kernel) and by bsx_map_pe.cu
// (big device functions are real calls: the paired-end kernel shrinks from 321 KB to 134 KB of SASS and gains
// 30 %, it was instruction-cache bound).
//
// One warp owns one read (SE) or one read pair (PE) from ASCII to result record; warps are
// persistent and fetch work with an atomic counter, so heavy-tailed candidate lists balance.
//
//   K2  load / trim / filter / pack   TrimAdapter, FilterReads, ConvertBinaySeq (align.cpp:371-425,
//                                     579-589, 90-162): ASCII staged in shared memory, 2-bit words
//                                     and the N mask built by lanes 0..9, all seed keys by XT.
//   K3  seed selection                ReorderSeed / AdjustSeedStartArray / CountSeeds
//                                     (align.cpp:454-577): every DISTINCT read offset that can carry
//                                     a seed is probed once (coalesced across lanes into 8-byte table
//                                     loads), then lane 0 replays the reference's argmin / sort logic
//                                     from shared memory.
//   K4  probe + extend + commit       SnpAlign + CountMismatch (align.cpp:253-346, align.h:167-200):
//                                     32 list entries per step, coalesced 128-B list loads; each lane
//                                     extends one candidate.  Extension is two-phase: first ONE aligned
//                                     16-byte gather (3 read words = 48 bases away from the seed),
//                                     XOR + asymmetric C/T mask + popcount; only survivors load the rest
//                                     of the window.  The partial count is a lower bound, so accept /
//                                     reject decisions are identical to the reference.  Survivors are
//                                     committed in lane order (ballot + serial loop) so dedupe, bucket
//                                     counts, -w threshold lowering and the -r 0 exits fire at exactly
//                                     the candidate the sequential reference would stop at.
//   K5  selection                     StringAlign (align.cpp:610-627) + myrand.
//   K6  pairing                       PairAlign::RunAlign / GetPairs (pairs.cpp:34-190).
//
// Integer, HBM-latency/bandwidth bound: no tensor cores.
#include <cstdio>
#include "bsx_map.cuh"

#ifndef BSX_READ_BLOCK
#define BSX_READ_BLOCK 4        // consecutive reads a warp takes per work-counter atomic
#endif
// RRBS mode (-D) as a compile-time constant where a translation unit fixes it (dead code leaves the binary)
#ifndef BSX_RRBS
#define BSX_RRBS(A) ((A).rrbs)
#endif
#ifndef BSX_SE_KERNEL
#define BSX_SE_KERNEL bsx_map_se_kernel
#define BSX_SE_OCC bsx_map_occupancy_se
#define BSX_SE_LAUNCH bsx_launch_map_se
#endif
#ifndef BSX_CALLS
#define BSX_CALLS 1             // 1: big device functions are real calls (small binary), 0: everything inlined
#endif
#if BSX_CALLS
#define BSX_FN __noinline__
#else
#define BSX_FN __forceinline__
#endif
#ifndef BSX_PIPE
#define BSX_PIPE 1              // software-pipeline the inline-context loads one step ahead
#endif
#ifndef BSX_SE_MIN_CTAS
#define BSX_SE_MIN_CTAS 5
#endif

namespace {

// work counters live in shared memory (SelSm::ctr, bsx_stats order); lane 0 updates them
enum { CT_CAND = 0, CT_PROBE, CT_OVER, CT_FULL, CT_COMMIT, CT_MAPPED, CT_LIST, CT_GATHER };
typedef uint32_t Ctr;
// warp-uniform state in shared memory is written by lane 0 only, fenced on both sides
#define WSET(lhs, rhs) do { const auto v_ = (rhs); __syncwarp(); if (lane == 0) (lhs) = v_; __syncwarp(); } while (0)
#define CTR_ADD(C, k, v) do { const uint32_t v_ = (uint32_t)(v); if (lane == 0) (C)[k] += v_; } while (0)   // v may hold warp collectives

__device__ __forceinline__ uint4 *plan_of(ReadSm *R, int chain, const MapArgs &A) {
    return reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(R) + sizeof(ReadSm)) + chain * A.chain_stride;
}
__device__ __forceinline__ uint4 *flank_of(ReadSm *R, int chain, const MapArgs &A) {
    return plan_of(R, chain, A) + A.flank_off;
}

__device__ __forceinline__ const CtaSm *cta_tables() {   // the CTA's tables sit at the start of dynamic shared memory
    extern __shared__ __align__(16) uint8_t bsx_dyn_smem_[];
    return reinterpret_cast<const CtaSm *>(bsx_dyn_smem_);
}

// per-CTA tables: seed profile and p -> (segment, remainder), so per-read code never divides
__device__ __forceinline__ void init_cta_tables(const MapArgs &A, CtaSm *K) {
    #pragma unroll 1
    for (int t = threadIdx.x; t < 256; t += blockDim.x) K->profA[t] = (uint8_t)bsx_profile_a(A.s, A.I, t >> 4, t & 15);
    #pragma unroll 1
    for (int t = threadIdx.x; t < 160; t += blockDim.x) { K->segof[t] = (uint8_t)(t / A.s); K->remof[t] = (uint8_t)(t % A.s); }
    const int per = BSX_RRBS(A) ? 1 : A.I;
    #pragma unroll 1
    for (int t = threadIdx.x; t < 256; t += blockDim.x) { K->divI[t] = (uint8_t)(t / per); K->modI[t] = (uint8_t)(t % per); }
    // chr_lut[g] = largest k with anchor[k] <= g << 24 (0 when none): int2hit then searches [chr_lut[g], chr_lut[g+1]]
    #pragma unroll 1
    for (int t = threadIdx.x; t < 258; t += blockDim.x) {
        int left = 0;
        if (t < 256 && A.n_seq < 65536u) {
            const uint32_t x = (uint32_t)t << 24;
            int right = (int)A.n_seq;
            while (left < right - 1) { const int mid = (left + right) / 2; if (x >= A.seqinfo[mid]) left = mid; else right = mid; }
        } else if (A.n_seq < 65536u) left = (int)A.n_seq - 1;
        K->chr_lut[t] = (uint16_t)left;
    }
    __syncthreads();
}

// ------------------------------------------------------------------ K2: load, trim, filter, pack
__device__ __forceinline__ void load_read(const MapArgs &A, ReadSm *R, const uint8_t *seqs,
                                          const uint16_t *lens, uint32_t r, int readset, int lane) {
    int len = lens[r];
    if (len > A.max_readlen) len = A.max_readlen;          // reads.cpp:115-117
    if (len > BSX_MAX_READLEN) len = BSX_MAX_READLEN;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(seqs + (size_t)r * A.stride);
    uint32_t *dst = reinterpret_cast<uint32_t *>(R->ascii);
    #pragma unroll 1
    for (int t = lane; t < 40; t += 32) dst[t] = (t * 4 < (int)A.stride) ? __ldg(src + t) : 0u;
    __syncwarp();
    __syncwarp();
    if (lane == 0) { R->len = len; R->raw = len; R->readset = readset; R->index = A.first_index + r; }
    __syncwarp();
}

// TrimAdapter (align.cpp:371-425): adapters in -A order, positions ascending, first success wins
__device__ BSX_FN void trim_adapter(const MapArgs &A, ReadSm *R, int lane) {
    WSET(R->raw, R->len);
    const int len = R->len, s = A.s;
    const uint8_t *sq = R->ascii;
    const int tail = BSX_RRBS(A) ? 5 : 4;
    #pragma unroll 1
    for (int a = 0; a < A.n_adapter; a++) {
        const int al = A.adapter_len[a];
        #pragma unroll 1
        for (int pos0 = s; pos0 < len - tail; pos0 += 32) {
            const int pos = pos0 + lane;
            bool ok = false;
            if (pos < len - tail) {
                int m0 = 0, k = 0;
                #pragma unroll 1
                for (; k < al && k < 15 && pos + k < len; k++) {
                    m0 += (A.adapter[a][k] != (char)sq[pos + k]);
                    if (m0 > 4) break;
                }
                if (!BSX_RRBS(A)) ok = (k >= m0 * 5 && k > 3);
                else if (k >= m0 * 5) {
                    // digestion-site remnant just before the adapter (align.cpp:383-404)
                    const int sl = A.site_len, dp = A.digest_pos;
                    int m = m0, m2 = m0;
                    #pragma unroll 1
                    for (int t = 0; t < sl - dp; t++) {
                        char x = A.digest_site[t], y = (char)sq[pos - sl + dp + t];
                        m += (x != y) && (x != 'C' || y != 'T');
                        m2 += (x != y) && (x != 'G' || y != 'A');
                    }
                    ok = (k >= m * 5) || (A.pairend && k >= m2 * 5);
                }
            }
            unsigned b = __ballot_sync(BSX_FULL, ok);
            if (b) { WSET(R->len, pos0 + __ffs(b) - 1); return; }
        }
    }
}

// FilterReads (align.cpp:579-589); returns 1 when the read is rejected.  The chains the read will be
// aligned with are packed here (ConvertBinaySeq), because the valid-base mask also gives CountNs.
__device__ __forceinline__ void pack_chain(const MapArgs &A, ReadSm *R, int chain, int lane);
__device__ __forceinline__ int filter_read(const MapArgs &A, ReadSm *R, int lane) {
    trim_adapter(A, R,  lane);
    {   // flag_chain / cflag_chain (align.cpp:93-94)
        const int fc = A.chains || (R->readset < 2), cc = A.chains || (R->readset == 2);
        __syncwarp();
        if (lane == 0) { R->fc = fc; R->cc = cc; }
        __syncwarp();
    }
    if (R->len < A.s) return 1;
    if (R->fc) pack_chain(A, R,  0, lane);
    if (R->cc) pack_chain(A, R,  1, lane);
    int nv = lane < BSX_FIXWORDS ? __popc(R->m5[R->fc ? 0 : 1][lane]) : 0;
#pragma unroll
    for (int d = 8; d; d >>= 1) nv += __shfl_xor_sync(BSX_FULL, nv, d);
    nv = __shfl_sync(BSX_FULL, nv, 0);
    if (R->len - nv > A.max_ns) return 1;            // CountNs (align.cpp:48-55)
    // read_max_snp_num = (v+1)*(len-1)/raw_readlen (align.cpp:586); equals v for an untrimmed read longer than v
    WSET(R->rmsn, (R->len == R->raw && A.v + 1 <= R->len) ? A.v : (int)((unsigned)(A.v + 1) * (unsigned)(R->len - 1) / (unsigned)R->raw));
    return 0;
}

// four ASCII bases in a u32 (first base in the low byte) -> their 2-bit codes and validity, byte-wise
__device__ __forceinline__ void codes4(uint32_t w, int rev, uint32_t &code, uint32_t &valid) {
    const uint32_t v = w | 0x20202020u;
    valid = __vcmpeq4(v, 0x61616161u) | __vcmpeq4(v, 0x63636363u) | __vcmpeq4(v, 0x67676767u) | __vcmpeq4(v, 0x74747474u);
    uint32_t c = (w >> 1) & 0x03030303u;          // A 0, C 1, G 3, T 2
    c ^= (c >> 1) & 0x01010101u;                   // A 0, C 1, G 2, T 3
    c &= valid;                                    // everything else -> 0 (alphabet[], param.cpp:210)
    if (rev) c = (~c) & 0x03030303u;               // complement; everything else -> 3 (rev_alphabet[], param.cpp:215)
    code = c;
    valid &= 0x01010101u;
}
// byte-wise 2-bit fields (first base in the low byte) -> 8 bits, first base most significant
__device__ __forceinline__ uint32_t squeeze4(uint32_t c) {
    return ((c & 0x3u) << 6) | ((c >> 4) & 0x30u) | ((c >> 14) & 0xCu) | (c >> 24);
}

// ConvertBinaySeq (align.cpp:90-162) for one chain: packed words + valid-base mask.
// Lane t < 20 converts bases [8t, 8t+8) with byte-SIMD ops; lane pairs are merged by shuffle.
__device__ __forceinline__ void pack_chain(const MapArgs &A, ReadSm *R, int chain, int lane) {
    const int len = R->len;
    uint32_t half = 0, mhalf = 0;
    if (lane < 2 * BSX_FIXWORDS) {
        uint32_t w0, w1;
        if (!chain) {
            const uint32_t *a32 = reinterpret_cast<const uint32_t *>(R->ascii);
            w0 = a32[2 * lane]; w1 = a32[2 * lane + 1];
        } else {                                    // reversed read: base i is ascii[len-1-i]
            w0 = 0; w1 = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int i0 = len - 1 - (8 * lane + k), i1 = i0 - 4;
                w0 |= (i0 >= 0 ? (uint32_t)R->ascii[i0] : 0u) << (8 * k);
                w1 |= (i1 >= 0 ? (uint32_t)R->ascii[i1] : 0u) << (8 * k);
            }
        }
        // bases beyond the read count as invalid code 0
        const int rem = len - 8 * lane;             // valid bases in this lane's 8
        uint32_t keep0 = rem >= 4 ? 0xffffffffu : (rem <= 0 ? 0u : (0xffffffffu >> (8 * (4 - rem))));
        uint32_t keep1 = rem >= 8 ? 0xffffffffu : (rem <= 4 ? 0u : (0xffffffffu >> (8 * (8 - rem))));
        uint32_t c0, v0, c1, v1;
        codes4(w0, chain, c0, v0); codes4(w1, chain, c1, v1);
        c0 &= keep0; v0 &= keep0; c1 &= keep1; v1 &= keep1;
        half = (squeeze4(c0) << 8) | squeeze4(c1);
        mhalf = (squeeze4(v0) << 8) | squeeze4(v1);
    }
    const uint32_t ohalf = __shfl_down_sync(BSX_FULL, half, 1), omhalf = __shfl_down_sync(BSX_FULL, mhalf, 1);
    if (lane < 2 * BSX_FIXWORDS && !(lane & 1)) {
        R->rw[chain][lane >> 1] = (half << 16) | ohalf;
        R->m5[chain][lane >> 1] = (mhalf << 16) | omhalf;
    }
    __syncwarp();
}

// seed_array[p] (align.cpp:101-105): 3-letter key of the seed starting at read offset p
__device__ __forceinline__ uint32_t seed_key(const MapArgs &A, const ReadSm *R, int chain, int p) {
    const int j = p >> 4, sh = (p & 15) * 2;
    const uint32_t hi = R->rw[chain][j], lo = (j + 1 < BSX_FIXWORDS) ? R->rw[chain][j + 1] : 0u;
    const uint32_t v = __funnelshift_l(lo, hi, sh) >> (32 - 2 * A.s);
    return bsx_xt(v & A.seed_bits, A.s);
}

// ------------------------------------------------------------------ K3: probe + choose seeds
__device__ BSX_FN void select_seeds(const MapArgs &A, const CtaSm *K, ReadSm *R, SelSm *X, int chain, int lane, Ctr *C) {
    const int s = A.s, I = A.I, len = R->len, seg = R->seedseg;
    const int mo = (BSX_RRBS(A) || len - I + 1 < 0) ? 0 : (int)K->remof[len - I + 1];   // max_offset = (len-I+1) % s
    const int cso = (BSX_RRBS(A) && chain) ? (int)K->remof[len] : 0;    // cseed_offset (RRBS rc chain)
    const int lim = I - 1 + mo;
    // 1. every read offset that can carry a seed: segment n owns [n*s, n*s + I-1 + max_offset]
    //    (profile.a - i lies in [n*s, n*s+I-1]); its list header is read ONCE, coalesced across lanes.
    //    The union of those ranges is enumerated directly: w offsets per segment, the last one takes the tail
    //    (when the ranges overlap, w = s and the union is one interval).  idx / w by a float reciprocal (idx < 4096).
    int np = 0;
    {
        const int w = min(lim + 1, s);
        const int total = BSX_RRBS(A) ? seg : (seg > 0 ? seg * w + (lim + 1 - w) : 0);
        const float rw = __frcp_rn((float)w);
        #pragma unroll 1
        for (int idx = lane; idx < total; idx += 32) {
            int p;
            if (BSX_RRBS(A)) p = cso + idx * s;
            else { const int n = min((int)(((float)idx + 0.5f) * rw), seg - 1); p = n * s + (idx - n * w); }
            if (p + s <= len) {
                const uint32_t key = seed_key(A, R, chain, p);
                const uint2 a = __ldg(reinterpret_cast<const uint2 *>(A.tab) + key);
                const uint32_t e = __ldg(A.tab + 2 * (size_t)key + 2);
                const uint32_t n = e - a.x;
                X->st[p] = a.x; X->md[p] = a.y;
                X->sz[p] = BSX_RRBS(A) ? n : (n ? n + 2 : 0u);    // index2[key][0] = n + 2 (App. B Q7); RRBS: n1
                np++;
            }
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) np += __shfl_xor_sync(BSX_FULL, np, d);
    CTR_ADD(C, CT_PROBE, np);
    __syncwarp();
    // 2. T[n][o] = CountSeeds(n, o) (align.cpp:549-556) for every segment and start offset o <= max_offset, in parallel
    if (!BSX_RRBS(A)) {
        const int mo1 = mo + 1;
        const float rm = __frcp_rn((float)mo1);
        #pragma unroll 1
        for (int idx = lane; idx < seg * mo1; idx += 32) {
            const int n = (int)(((float)idx + 0.5f) * rm), o = idx - n * mo1;
            uint32_t tt = 0;
            #pragma unroll 1
            for (int k = 0; k < I; k++) tt += X->sz[(int)K->profA[n * 16 + k] + o - k];
            X->T[n * 16 + o] = tt;
        }
        __syncwarp();
    }
    // 3. ReorderSeed (align.cpp:454-468): global offset = FIRST minimum of GetTotalSeedLoc over [0, max_offset)
    int og = 0;                                  // App. B Q4: defined as 0 when the loop is empty
    if (!BSX_RRBS(A) && mo > 1) {                 // with a single candidate (max_offset == 1) the answer is 0
        unsigned long long best = ~0ull;
        if (lane < mo) {
            uint32_t tt = 0;
            #pragma unroll 1
            for (int n = 0; n < seg; n++) tt += X->T[n * 16 + lane];
            best = ((unsigned long long)tt << 8) | (unsigned)lane;
        }
#pragma unroll
        for (int d = 16; d; d >>= 1) { unsigned long long o2 = __shfl_xor_sync(BSX_FULL, best, d); best = o2 < best ? o2 : best; }
        og = (int)(best & 0xff);
    }
    uint4 *plan = plan_of(R, chain, A), *flank = flank_of(R, chain, A);
    if (lane == 0) {
        // AdjustSeedStartArray (align.cpp:506-528)
        #pragma unroll 1
        for (int n = 0; n < seg; n++) X->arr[n] = og;
        if (!BSX_RRBS(A)) {
            #pragma unroll 1
            for (int i = 0; i < seg; i++) {
                const int ptr = (i & 1) == 0 ? i / 2 : seg - 1 - i / 2;
                uint32_t total = 0xffffffffu;
                const int start = (ptr == 0) ? 0 : X->arr[ptr - 1];
                const int end = (ptr == seg - 1) ? mo : X->arr[ptr + 1];
                int bi = start;
                #pragma unroll 1
                for (int ii = start; ii <= end; ii++) {
                    const uint32_t tt = X->T[ptr * 16 + ii];
                    if (tt < total) { total = tt; bi = ii; }
                }
                X->arr[ptr] = bi;
            }
        }
    }
    __syncwarp();
    // seedindex: (sum of list sizes, segment) ascending (align.cpp:474-485): rank sort, one segment per lane
    int mine = 0;
    if (lane < seg) {
        mine = (int)(BSX_RRBS(A) ? X->sz[lane * s + cso] : X->T[lane * 16 + X->arr[lane]]);
        X->sidx[lane][0] = mine;
    }
    __syncwarp();
    int rank = 0;
    if (lane < seg) {
        #pragma unroll 1
        for (int m = 0; m < seg; m++) {
            const int other = X->sidx[m][0];
            rank += (other < mine) || (other == mine && m < lane);
        }
    }
    __syncwarp();
    if (lane < seg) X->sidx[rank][1] = lane;
    __syncwarp();
    // plan[mode][k]: list bounds and read offset of sub-seed k of the segment processed in that mode
    const int per = BSX_RRBS(A) ? 1 : I;
    #pragma unroll 1
    for (int t = lane; t < seg * per; t += 32) {
        const int m = K->divI[t], k = K->modI[t];
        const int sg = X->sidx[m][1];
        const int p = BSX_RRBS(A) ? (sg * s + cso) : ((int)K->profA[sg * 16 + k] + X->arr[sg] - k);
        const uint32_t st0 = X->st[p], sz0 = X->sz[p];
        const uint32_t en0 = st0 + (BSX_RRBS(A) ? sz0 : (sz0 ? sz0 - 2u : 0u));
        plan[t] = make_uint4(st0, X->md[p], en0, (uint32_t)p | ((uint32_t)sg << 16));
        if (!BSX_RRBS(A)) {
            // read bases / valid mask facing an entry's inline context: [p-16, p) and [p+s, p+s+16)
            const int xb = p - 16, xa = p + s;
            uint32_t rb = 0, mb = 0;
            if (xb >= 0) {
                const int j = xb >> 4, sh = (xb & 15) * 2;
                rb = __funnelshift_l(R->rw[chain][j + 1], R->rw[chain][j], sh);      // j + 1 <= 9 because p <= 144
                mb = __funnelshift_l(R->m5[chain][j + 1], R->m5[chain][j], sh);
            } else if (xb > -16) {                                                   // fewer than 16 bases before the seed
                rb = R->rw[chain][0] >> (2 * (-xb)); mb = R->m5[chain][0] >> (2 * (-xb));
            }
            const int j = xa >> 4, sh = (xa & 15) * 2;
            const uint32_t r1 = (j + 1 < BSX_FIXWORDS) ? R->rw[chain][j + 1] : 0u, m1 = (j + 1 < BSX_FIXWORDS) ? R->m5[chain][j + 1] : 0u;
            const uint32_t r0 = (j < BSX_FIXWORDS) ? R->rw[chain][j] : 0u, m0 = (j < BSX_FIXWORDS) ? R->m5[chain][j] : 0u;
            flank[t] = make_uint4(rb, mb, __funnelshift_l(r1, r0, sh), __funnelshift_l(m1, m0, sh));
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------ K4: extension
// which 16-byte chunk of the window to gather first: for every residue o = (word index & 3) pick
// delta in {0,1,2} maximising the number of valid read bases outside the seed that the chunk's
// three fully covered read words hold.  Returns delta for o = 0..3 packed 2 bits each.
__device__ __forceinline__ uint32_t chunk_table(const ReadSm *R, int chain, int nw, int zlo, int zhi, int lane) {
    const int o = (lane / 3) & 3, dl = lane % 3;
    int score = 0;
    if (lane < 12) {
        const int jlo = 4 * dl - o;
#pragma unroll
        for (int t = 0; t < 3; t++) {
            const int j = jlo + t;
            if (j >= 0 && j < nw) {
                int lo = max(zlo, 16 * j) - 16 * j, hi = min(zhi, 16 * j + 16) - 16 * j;
                uint32_t seedm = 0;
                if (lo < hi) {
                    uint32_t a = 0xffffffffu >> (2 * lo);
                    uint32_t b = (hi >= 16) ? 0u : (0xffffffffu >> (2 * hi));
                    seedm = a & ~b;
                }
                score += __popc(R->m5[chain][j] & ~seedm);
            }
        }
    }
    const int oo = lane & 3;
    const int a0 = __shfl_sync(BSX_FULL, score, 3 * oo), a1 = __shfl_sync(BSX_FULL, score, 3 * oo + 1),
              a2 = __shfl_sync(BSX_FULL, score, 3 * oo + 2);
    int best = 0, bs = a0;
    if (a1 > bs) { best = 1; bs = a1; }
    if (a2 > bs) { best = 2; }
    uint32_t tb = (uint32_t)best << (2 * oo);
    return __shfl_sync(BSX_FULL, tb, 0) | __shfl_sync(BSX_FULL, tb, 1) | __shfl_sync(BSX_FULL, tb, 2) | __shfl_sync(BSX_FULL, tb, 3);
}

__device__ __forceinline__ uint32_t partial_mismatch(const ReadSm *R, int chain, int nw, const uint32_t *__restrict__ refbase,
                                                     uint32_t loc, uint32_t tbl) {
    const uint32_t wi = loc >> 4, sh2 = (loc & 15u) * 2u, o = wi & 3u;
    const uint32_t dl = (tbl >> (2 * o)) & 3u;
    const int jlo = 4 * (int)dl - (int)o;
    // ld.global.nc with a 64-byte L2 fetch: the default promotes every miss to a full 128-byte line
    // (measured: 121 B of HBM traffic per 16-byte gather vs 62 B with .L2::64B, profiles/ubench)
    uint4 c;
    {
        const uint4 *gp = reinterpret_cast<const uint4 *>(refbase) + (wi >> 2) + dl;
        asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w) : "l"(gp));
    }
    uint32_t w = 0;
    int j = jlo;
    if (j >= 0 && j < nw) w += __popc(bsx_mm_word_bits(R->rw[chain][j], R->m5[chain][j], __funnelshift_l(c.y, c.x, sh2)));
    j++;
    if (j >= 0 && j < nw) w += __popc(bsx_mm_word_bits(R->rw[chain][j], R->m5[chain][j], __funnelshift_l(c.z, c.y, sh2)));
    j++;
    if (j >= 0 && j < nw) w += __popc(bsx_mm_word_bits(R->rw[chain][j], R->m5[chain][j], __funnelshift_l(c.w, c.z, sh2)));
    return w;
}

// CountMismatch (align.h:167-200) over the whole read; stops early once above the threshold
// (measured: requesting all window words up front is 10 % slower on config 2 -- most phase-1 survivors
// are rejected within the first two words)
__device__ __forceinline__ uint32_t full_mismatch(const ReadSm *R, int chain, int nw, const uint32_t *__restrict__ refbase,
                                                  uint32_t loc, uint32_t thres) {
    const uint32_t *rp = refbase + (loc >> 4);
    const uint32_t sh2 = (loc & 15u) * 2u;
    uint32_t w = 0, prev = __ldg(rp);
    #pragma unroll 1
    for (int j = 0; j < nw; j++) {
        const uint32_t next = __ldg(rp + j + 1);
        w += __popc(bsx_mm_word_bits(R->rw[chain][j], R->m5[chain][j], __funnelshift_l(next, prev, sh2)));
        prev = next;
        if (w > thres) break;
    }
    return w;
}

// RefSeq::CCGG_seglen (dbseq.cpp:541-567), clamped at the last site (App. B Q20)
__device__ int ccgg_seglen(const MapArgs &A, uint32_t chr, uint32_t pos, int readlen) {
    const uint32_t *st = A.sites + A.site_off[chr >> 1];
    const int n = (int)(A.site_off[(chr >> 1) + 1] - A.site_off[chr >> 1]);
    int left = 0, right = n - 1;
    while (left < right - 1) {
        int mid = (left + right) / 2;
        uint32_t mv = st[mid];
        if (mv == pos) { left = mid; right = mid + 1; break; }
        else if (mv < pos) left = mid; else right = mid;
    }
    const uint32_t seg_start = st[left], add = (uint32_t)(A.site_len - 2 * A.digest_pos);
    uint32_t seg_end;
    for (;;) {
        int rr = right < n ? right : n - 1;
        seg_end = st[rr] + add;
        if (seg_end < pos + (uint32_t)readlen && right < n) right++; else break;
    }
    return (int)(seg_end - seg_start);
}

// One accepted candidate, executed warp-uniformly: int2hit, bounds, dedupe, bucket append, exits
// (align.cpp:270-278 and its three twins).  Returns 1 when SnpAlign must return.
__device__ int commit_hit(const MapArgs &A, ReadSm *R, uint2 *hits, uint32_t *dd, int store_all, int chain,
                          uint32_t chr, uint32_t loc, uint32_t w, int mode, int frag_filter, int lane, Ctr *C) {
    const uint32_t *anchor = A.seqinfo, *size = A.seqinfo + A.n_seq + 1, *rcoff = A.seqinfo + 2 * A.n_seq + 1;
    const uint32_t k = chr >> 1;
    if (chr & 1u) loc = rcoff[k] - (uint32_t)R->len - loc;
    if (loc + (uint32_t)R->len > size[k]) return 0;                      // overflow the end of refseq
    const uint32_t key = anchor[k] + loc;                               // == (chr>>1, loc), see DESIGN.md
    bool found = false;
    const uint32_t dn = R->dn;
    #pragma unroll 1
    for (uint32_t t = lane; t < dn; t += 32) found |= (dd[t] == key);
    if (__any_sync(BSX_FULL, found)) return 0;                          // hit already exists
    if (dn < A.dd_stride) {                                              // capacity guard (RRBS fragment-filtered hits are uncounted)
        __syncwarp();
        if (lane == 0) { dd[dn] = key; R->dn = dn + 1; }
        __syncwarp();
    }
    if (frag_filter) {
        const int sl = ccgg_seglen(A, chr, loc, R->len);
        if (sl > A.max_insert || sl < A.min_insert) { __syncwarp(); return 0; }
    }
    const uint32_t cnt = chain ? R->nc[w] : R->nh[w];
    if ((int)w < R->best) WSET(R->best, (int)w);       // lowest level that holds a hit
    if (lane == 0) {
        if (store_all) hits[((size_t)w * 2 + chain) * (A.W + 1) + cnt] = make_uint2(chr, loc);
        else if ((int)w == R->best) hits[(size_t)chain * (A.W + 1) + cnt] = make_uint2(chr, loc);
        if (chain) R->nc[w] = (uint16_t)(cnt + 1); else R->nh[w] = (uint16_t)(cnt + 1);
    }
    __syncwarp();
    CTR_ADD(C, CT_COMMIT, 1);
    const int tot = (int)R->nh[w] + (int)R->nc[w];
    if ((int)w == mode && !A.pairend && A.r == 0 && tot > 1) return 1;
    if (tot >= A.W) { if (w == 0) return 1; WSET(R->thres, w - 1); }
    return 0;
}

// 32 candidates that survived phase 0 (table entries idx0 + lane of one list): phase 1 (one aligned 16-byte
// gather), phase 2 (whole window, exact CountMismatch) and the ordered commit.  Everything a survivor
// needs is re-derived here from its table index, so the filter loop carries no state for it.
// Returns 1 when SnpAlign must return; `last` = exiting lane.
__device__ BSX_FN int extend_and_commit(const MapArgs &A, ReadSm *R, uint2 *hits, uint32_t *dd,
                                              int store_all, int chain, int mode, bool pass, uint32_t idx0, uint32_t md,
                                              uint32_t p, uint32_t tbl, int use_p1, int lane, Ctr *C) {
    const uint32_t *anchor = A.seqinfo;
    uint32_t strand = 0, loc = anchor[0], chr = 0;
    if (pass) {
        const uint32_t idx = idx0 + lane;
        const uint32_t entry = __ldg(A.pos + idx);
        if (!BSX_RRBS(A)) { strand = idx >= md; loc = entry - p; }           // h = -profile.a + i - seed_start_array
        else { chr = __ldg(A.tag + idx) & 0xffffu; strand = chr & 1u; loc = entry - p + anchor[chr >> 1]; }
    }
    const uint32_t *refbase = strand ? A.crefcat : A.refcat;
    uint32_t w = 0xffffu;
    if (use_p1) {                                                        // many survivors: one 16-byte gather filters first
        CTR_ADD(C, CT_GATHER, __popc(__ballot_sync(BSX_FULL, pass)));
        if (pass) {
            w = partial_mismatch(R, chain, R->nw, refbase, loc, tbl);
            pass = w <= R->thres;
        }
    }
    const unsigned pm1 = __ballot_sync(BSX_FULL, pass);
    unsigned pm = 0;
    if (pm1) {
        if (pass) {
            w = full_mismatch(R, chain, R->nw, refbase, loc, R->thres);
            pass = w <= R->thres;
        }
        pm = __ballot_sync(BSX_FULL, pass);
        CTR_ADD(C, CT_FULL, __popc(pm1));
    }
    int ret = 0, last = 31;
    while (pm) {
        const int src = __ffs(pm) - 1;
        pm &= pm - 1;
        const uint32_t w_s = __shfl_sync(BSX_FULL, w, src);
        if (w_s > R->thres) continue;                                     // threshold lowered by an earlier commit
        uint32_t loc_s = __shfl_sync(BSX_FULL, loc, src);
        const uint32_t strand_s = __shfl_sync(BSX_FULL, strand, src);
        uint32_t chr_s;
        if (!BSX_RRBS(A)) {
            // RefSeq::int2hit (dbseq.cpp:585-595); the per-CTA table narrows the search to the sequences
            // that overlap the position's 16 Mb granule (almost always one)
            int left = 0, right = (int)A.n_seq;
            if (A.n_seq < 65536u) { const CtaSm *K = cta_tables(); left = K->chr_lut[loc_s >> 24]; right = K->chr_lut[(loc_s >> 24) + 1] + 1; }
            while (left < right - 1) { int mid = (left + right) / 2; if (loc_s >= anchor[mid]) left = mid; else right = mid; }
            chr_s = (uint32_t)left * 2u + strand_s;
            loc_s -= anchor[left];
        } else {
            chr_s = __shfl_sync(BSX_FULL, chr, src);
            loc_s -= anchor[chr_s >> 1];
        }
        ret = commit_hit(A, R,  hits, dd, store_all, chain, chr_s, loc_s, w_s, mode,
                         BSX_RRBS(A) && chain == 0 && !A.pairend, lane, C);
        if (ret) { last = src; break; }
    }
    return ret | (last << 8);        // bit 0: SnpAlign returns; bits 8..: exiting lane
}

// SnpAlign (align.cpp:168-347) for one mode; returns 1 if it `return`ed early.
// The I position lists of the mode are walked in the reference's order (sub-seed 0's forward entries, its
// rc entries, sub-seed 1's, ...), 64 table entries per step (two per lane).  Everything that depends on
// the list (its bounds, the read bases that face the inline context) is warp-uniform, so the per-candidate
// work is one 8-byte load, two masked XOR/popcount words and a compare.
__device__ BSX_FN int snp_align(const MapArgs &A, ReadSm *R, SelSm *X, uint2 *hits, uint32_t *dd, int store_all, int mode, int lane, Ctr *C) {
    const int per = BSX_RRBS(A) ? 1 : A.I;
    for (int chain = 0; chain < 2; chain++) {
        if (chain == 0 ? !R->fc : !R->cc) continue;
        const uint4 *plan = plan_of(R, chain, A) + mode * per, *flank = flank_of(R, chain, A) + mode * per;
        uint32_t tbl = 0; bool have_tbl = false;
        uint32_t visited = 0, counted = 0;                               // list entries loaded / reference-visible candidates
        int ret = 0;
        #pragma unroll 1
        for (int i = 0; i < per && !ret; i++) {
            const uint4 e = plan[i];                                     // {list start, rc start, list end, p | segment << 16}
            if (e.x == e.z) continue;                                    // index2[_seed] == NULL
            const uint32_t p = e.w & 0xffffu;
            uint32_t rb = 0, mb = 0, ra = 0, ma = 0, want = 0;
            if (!BSX_RRBS(A)) {
                const uint4 f = flank[i];                                // prepared with the plan (select_seeds)
                rb = f.x; mb = f.y; ra = f.z; ma = f.w;
            } else {
                const int sg = (int)(e.w >> 16);
                want = chain ? (uint32_t)(R->len / A.s - 1 - sg) : (uint32_t)sg;     // RRBS segment tag
            }
            if (!BSX_RRBS(A)) {
                // ---- WGBS: phase 0 = mismatches among the <= 32 read bases that face the entry's inline context
                // (8 bytes that arrive with the list stream).  It is a lower bound of CountMismatch, so
                // `> snp_thres` rejects exactly like the reference; pos[] and the reference are only touched by
                // survivors.  Every entry of the list is a candidate, so the counters need no per-step work.
                uint32_t thres = R->thres, c0 = e.x, exit_pos = 0;
                for (; c0 < e.z; c0 += 64) {
                    const uint32_t i0 = c0 + lane, i1 = i0 + 32;
                    bool pass0 = false, pass1 = false;
                    uint2 cx0 = make_uint2(0, 0), cx1 = make_uint2(0, 0);
                    if (i0 < e.z) cx0 = __ldg(A.ctx + i0);
                    if (i1 < e.z) cx1 = __ldg(A.ctx + i1);
                    if (i0 < e.z) pass0 = __popc(bsx_mm_word_bits(rb, mb, cx0.x)) + __popc(bsx_mm_word_bits(ra, ma, cx0.y)) <= thres;
                    if (i1 < e.z) pass1 = __popc(bsx_mm_word_bits(rb, mb, cx1.x)) + __popc(bsx_mm_word_bits(ra, ma, cx1.y)) <= thres;
                    if (!__any_sync(BSX_FULL, pass0 || pass1)) continue;
                    const unsigned pm0 = __ballot_sync(BSX_FULL, pass0), pm1 = __ballot_sync(BSX_FULL, pass1);
                    // phase 1 (one aligned 16-byte gather per survivor) only pays when phase 0 lets many through
                    // (high -v); otherwise survivors go straight to the exact count
                    const int use_p1 = __popc(pm0) + __popc(pm1) > 2;
                    if (use_p1 && !have_tbl) {
                        // phase-1 chunk choice: keep away from the seed zone of this mode (all sub-seeds)
                        int zlo = 1000, zhi = -1;
                        if (lane < per) { zlo = zhi = (int)(plan[lane].w & 0xffffu); }
#pragma unroll
                        for (int d = 8; d; d >>= 1) { zlo = min(zlo, __shfl_xor_sync(BSX_FULL, zlo, d)); zhi = max(zhi, __shfl_xor_sync(BSX_FULL, zhi, d)); }
                        zlo = __shfl_sync(BSX_FULL, zlo, 0); zhi = __shfl_sync(BSX_FULL, zhi, 0) + A.s;
                        tbl = chunk_table(R, chain, R->nw, zlo, zhi, lane);
                        have_tbl = true;
                    }
#pragma unroll 1
                    for (int h = 0; h < 2; h++) {                        // one call site: the slow path exists once in the binary
                        if (h ? pm1 : pm0) {
                            const int rc = extend_and_commit(A, R, hits, dd, store_all, chain, mode, h ? pass1 : pass0, c0 + 32u * h, e.y, p, tbl, use_p1, lane, C);
                            if (rc & 1) { ret = 1; exit_pos = (c0 - e.x) + 32u * h + (uint32_t)(rc >> 8) + 1u; break; }
                        }
                    }
                    if (ret) break;
                    thres = R->thres;
                }
                if (!ret) { visited += e.z - e.x; counted += e.z - e.x; }
                else { visited += min(c0 + 64u, e.z) - e.x; counted += exit_pos; }
            } else {
                // ---- RRBS: tagged Hit{chr, loc}: segment/strand filter, then underflow test (align.cpp:187-194,
                // 229-236); survivors of the filter are the reference's candidates
                for (uint32_t c0 = e.x; c0 < e.z; c0 += 64) {
                    const uint32_t i0 = c0 + lane, i1 = i0 + 32;
                    bool pass0 = false, pass1 = false;
                    unsigned vm0 = 0, vm1 = 0;
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const uint32_t idx = h ? i1 : i0;
                        bool valid = idx < e.z;
                        if (valid) {
                            const uint32_t tag = __ldg(A.tag + idx);
                            if (((chain ? (tag ^ 0x1000000u) : tag) >> 16) != want) valid = false;
                            if (__ldg(A.pos + idx) < p) valid = false;
                        }
                        if (h == 0) { pass0 = valid; vm0 = __ballot_sync(BSX_FULL, valid); }
                        else { pass1 = valid; vm1 = __ballot_sync(BSX_FULL, valid); }
                    }
                    visited += min(64u, e.z - c0);
                    if ((vm0 | vm1) == 0) continue;
                    if (!have_tbl) {
                        int zlo = 1000, zhi = -1;
                        if (lane < per) { zlo = zhi = (int)(plan[lane].w & 0xffffu); }
#pragma unroll
                        for (int d = 8; d; d >>= 1) { zlo = min(zlo, __shfl_xor_sync(BSX_FULL, zlo, d)); zhi = max(zhi, __shfl_xor_sync(BSX_FULL, zhi, d)); }
                        zlo = __shfl_sync(BSX_FULL, zlo, 0); zhi = __shfl_sync(BSX_FULL, zhi, 0) + A.s;
                        tbl = chunk_table(R, chain, R->nw, zlo, zhi, lane);
                        have_tbl = true;
                    }
#pragma unroll 1
                    for (int h = 0; h < 2; h++) {
                        const unsigned vmh = h ? vm1 : vm0;
                        if (vmh) {
                            const int rc = extend_and_commit(A, R, hits, dd, store_all, chain, mode, h ? pass1 : pass0, c0 + 32u * h, e.y, p, tbl, 1, lane, C);
                            if (rc & 1) { ret = 1; counted += __popc(vmh & ((2u << (rc >> 8)) - 1u)); break; }
                        }
                        counted += __popc(vmh);
                    }
                    if (ret) break;
                }
            }
        }
        // C = candidates the sequential reference visits (RRBS: tag-filtered entries are not counted);
        // entries evaluated past an exit point count as over-fetch
        if (lane == 0) {
            C[CT_LIST] += visited;
            C[CT_CAND] += counted;
            if (ret) C[CT_OVER] += (BSX_RRBS(A) ? 0u : visited - counted);
        }
        if (ret) return 1;
    }
    return 0;
}

// everything RunAlign does before the mode loop (align.cpp:435-444)
__device__ BSX_FN void prepare_read(const MapArgs &A, const CtaSm *K, ReadSm *R, SelSm *X, int lane, Ctr *C, uint32_t *dbg) {
    {
        const int q = R->len - A.I + 1;
        int seg = q > 0 ? min((int)K->segof[q], R->rmsn + 1) : 0;
        const int rmsn = R->rmsn, nw = (R->len + 15) >> 4;
        __syncwarp();
        if (lane == 0) { R->seedseg = seg; R->thres = (uint32_t)rmsn; R->nw = nw; R->dn = 0; R->best = 99; }
        __syncwarp();
    }
    if (lane < 16) { R->nh[lane] = 0; R->nc[lane] = 0; }
    __syncwarp();
    for (int chain = 0; chain < 2; chain++) {
        if (chain == 0 ? !R->fc : !R->cc) continue;
        select_seeds(A, K, R, X,  chain, lane, C);
        if (dbg && lane == 0) {
            dbg[chain * 20 + 0] = (uint32_t)R->seedseg;
            for (int n = 0; n < R->seedseg && n < 9; n++) { dbg[chain * 20 + 1 + n] = (uint32_t)X->arr[n]; dbg[chain * 20 + 10 + n] = (uint32_t)X->sidx[n][1]; }
        }
        __syncwarp();
    }
}

// SingleAlign::RunAlign (align.cpp:435-452)
__device__ BSX_FN void run_align(const MapArgs &A, const CtaSm *K, ReadSm *R, SelSm *X, uint2 *hits, uint32_t *dd, int store_all, int lane, Ctr *C, uint32_t *dbg) {
    prepare_read(A, K, R, X,  lane, C, dbg);
    #pragma unroll 1
    for (int m = 0; m < R->seedseg; m++) {
        snp_align(A, R, X,  hits, dd, store_all, m, lane, C);
        if (!BSX_RRBS(A) && R->best <= m) return;       // a bucket <= m is non-empty (align.cpp:448)
    }
}

// StringAlign (align.cpp:610-627) -> record
__device__ void write_record(const MapArgs &A, const ReadSm *R, const uint2 *hits, int store_all,
                             bsx_rec *out, uint16_t *cnt, int lane) {
    if (cnt && lane < 16) cnt[lane] = (!R->filtered && lane <= R->rmsn) ? (uint16_t)(R->nh[lane] + R->nc[lane]) : (uint16_t)0;
    if (lane != 0) return;
    bsx_rec o;
    o.loc = 0; o.chr = 0; o.nhits = 0; o.nm = 0; o.chain = 0; o.status = (uint8_t)R->filtered; o.len = (uint8_t)R->len;
    if (!R->filtered) {
        const int ii = R->best <= R->rmsn ? R->best : R->rmsn + 1;     // lowest non-empty bucket
        const int sum = ii <= R->rmsn ? R->nh[ii] + R->nc[ii] : 0;
        o.nm = (uint8_t)ii;
        if (sum > 0) {
            const int j = (int)(bsx_myrand(R->index, A.randseed) % (uint32_t)sum);
            const int nh = R->nh[ii];
            const int chain = j >= nh;
            const size_t lvl = store_all ? (size_t)ii * 2 : 0;
            const uint2 h = hits[(lvl + chain) * (A.W + 1) + (chain ? j - nh : j)];
            o.chr = h.x; o.loc = h.y; o.nhits = (uint32_t)sum; o.chain = (uint8_t)chain;
        }
    }
    *out = o;
}

__device__ __forceinline__ void flush_counters(const MapArgs &A, Ctr *C, int lane) {
    __syncwarp();
    if (lane < 8) { atomicAdd(A.stats + lane, (unsigned long long)C[lane]); C[lane] = 0; }
    __syncwarp();
}

#ifdef BSX_BUILD_SE
// ------------------------------------------------------------------ SE kernel
__global__ void __launch_bounds__(BSX_WARPS_PER_CTA * 32, BSX_SE_MIN_CTAS)
BSX_SE_KERNEL(const __grid_constant__ MapArgs A) {
    extern __shared__ __align__(16) uint8_t smem[];
#ifdef BSX_OPAQUE_LANE
    int lane, wid;   // opaque to the optimiser: held in registers instead of re-read from %tid at every use
    asm volatile("{ .reg .u32 t; mov.u32 t, %%tid.x; and.b32 %0, t, 31; shr.u32 %1, t, 5; }" : "=r"(lane), "=r"(wid));
#else
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#endif
    CtaSm *K = reinterpret_cast<CtaSm *>(smem);
    init_cta_tables(A, K);
    uint8_t *base = smem + sizeof(CtaSm) + A.warp_smem_se * wid;
    ReadSm *R = reinterpret_cast<ReadSm *>(base);
    SelSm *X = reinterpret_cast<SelSm *>(base + A.read_smem);
    const uint32_t gw = blockIdx.x * BSX_WARPS_PER_CTA + wid;
    uint2 *hits = A.hit_scratch + (size_t)gw * A.hit_stride;
    uint32_t *dd = A.dd_scratch + (size_t)gw * A.dd_stride;
    Ctr *C = X->ctr;
    if (lane < 8) C[lane] = 0;
    __syncwarp();
    uint32_t r = 0, r_end = 0;
    for (;;) {
        if (r == r_end) {                       // one atomic hands this warp BSX_READ_BLOCK consecutive reads
            if (lane == 0) r = atomicAdd(A.work_counter, (uint32_t)BSX_READ_BLOCK);
            r = __shfl_sync(BSX_FULL, r, 0);
            if (r >= A.n) break;
            r_end = min(r + (uint32_t)BSX_READ_BLOCK, A.n);
        }
        if ((C[CT_CAND] | C[CT_LIST]) & 0x80000000u) flush_counters(A, C, lane);
        __syncwarp();
        if (lane == 0) { R->rmsn = 0; R->seedseg = 0; R->nw = 0; R->thres = 0; R->fc = R->cc = 0; R->dn = 0; R->best = 99; }
        __syncwarp();
        load_read(A, R,  A.seq_a, A.len_a, r, A.readset, lane);
        WSET(R->filtered, filter_read(A, R, lane));
        uint32_t *dbg = A.debug ? A.debug + (size_t)r * 40 : nullptr;
        if (!R->filtered) run_align(A, K, R, X,  hits, dd, 0, lane, C, dbg);
        else if (lane < 16) { R->nh[lane] = 0; R->nc[lane] = 0; }
        __syncwarp();
        write_record(A, R,  hits, 0, A.out_a + r, A.cnt_a ? A.cnt_a + (size_t)r * 16 : nullptr, lane);
        if (!R->filtered && R->best <= R->rmsn) CTR_ADD(C, CT_MAPPED, 1);
        __syncwarp();
        r++;
    }
    flush_counters(A, C, lane);
}

#endif  // BSX_BUILD_SE

#ifdef BSX_BUILD_PE
// ------------------------------------------------------------------ PE kernel (pairs.cpp)
// pair buckets: pairhits[na+nb][..] as uint4 {a.chr, a.loc, b.chr, b.loc}; insert / chain / na / nb
// in a parallel uint4.  One warp per pair; GetPairs runs on lane 0 (its loops are short and strictly
// ordered), the two mates' SnpAlign calls are the warp-parallel part.
struct PairHitDev { uint32_t a_chr, a_loc, b_chr, b_loc; int32_t insert; uint32_t meta; /* chain | na<<8 | nb<<16 */ uint32_t pad0, pad1; };

__device__ void sort_hits(uint2 *h, int n, int lane) {
    // SortHits4PE (align.cpp:363-368) with HitComp (utilities.cpp:53): ascending (chr, loc).  Bucket
    // elements are distinct (dedupe), so any correct sort agrees with std::sort.  Warp-parallel
    // odd-even transposition in place; buckets are almost always 0-2 entries.
    if (n < 2) return;
    bool dirty = true;
    while (dirty) {
        bool sw = false;
        #pragma unroll 1
        for (int phase = 0; phase < 2; phase++) {
            #pragma unroll 1
            for (int i = phase + 2 * lane; i + 1 < n; i += 64) {
                const uint2 a = h[i], b = h[i + 1];
                if (a.x > b.x || (a.x == b.x && a.y > b.y)) { h[i] = b; h[i + 1] = a; sw = true; }
            }
            __syncwarp();
        }
        dirty = __any_sync(BSX_FULL, sw);
    }
}

// GetPairs (pairs.cpp:34-135), lane 0 only
__device__ int get_pairs(const MapArgs &A, const ReadSm *Ra, const ReadSm *Rb,
                         const uint2 *ha_all, const uint2 *hb_all, PairHitDev *pairs, uint16_t *npairs, int na, int nb) {
    if (na > Ra->rmsn || nb > Rb->rmsn) return 0;
    const size_t W1 = (size_t)A.W + 1;
    uint16_t &cnt = npairs[na + nb];
    PairHitDev *bucket = pairs + (size_t)(na + nb) * W1;
    for (int dir = 0; dir < 2; dir++) {
        const uint2 *ha = ha_all + ((size_t)na * 2 + dir) * W1;           // dir 0: a.hits  x b.chits
        const uint2 *hb = hb_all + ((size_t)nb * 2 + (1 - dir)) * W1;     // dir 1: a.chits x b.hits
        const int cnt_a = dir ? Ra->nc[na] : Ra->nh[na];
        const int cnt_b = dir ? Rb->nh[nb] : Rb->nc[nb];
        uint32_t chra = 0xffffffffu; int bstart = 0, bend = 0;
        #pragma unroll 1
        for (int i = 0; i < cnt_a; i++) {
            const uint2 x = ha[i];
            if (chra != x.x) {
                chra = x.x;
                for (bstart = bend; bstart < cnt_b; bstart++) if (hb[bstart].x >= chra) break;
                for (bend = bstart; bend < cnt_b; bend++) if (hb[bend].x > chra) break;
            }
            #pragma unroll 1
            for (int j = bstart; j < bend; j++) {
                const uint2 y = hb[j];
                const bool a_first = dir ? ((chra & 1u) != 0) : ((chra & 1u) == 0);
                uint32_t seg_start, seg_end;
                if (!a_first) { seg_start = y.y; seg_end = x.y + (uint32_t)Ra->len; }
                else { seg_start = x.y; seg_end = y.y + (uint32_t)Rb->len; }
                const int ins = (int)(seg_end - seg_start);
                if (ins >= A.min_insert && ins <= A.max_insert) {
                    PairHitDev ph;
                    ph.a_chr = x.x; ph.a_loc = x.y; ph.b_chr = y.x; ph.b_loc = y.y; ph.insert = ins;
                    ph.meta = (uint32_t)dir | ((uint32_t)na << 8) | ((uint32_t)nb << 16); ph.pad0 = ph.pad1 = 0;
                    bucket[cnt++] = ph;
                    if ((int)cnt >= A.W) return 1;
                }
            }
        }
    }
    return cnt > 0;
}

// the selection half of StringAlignUnpair (pairs.cpp:244-286) for one mate -> record
__device__ void write_unpaired(const MapArgs &A, const ReadSm *R, const uint2 *hits, bsx_rec *out, uint16_t *cnt, int lane) {
    if (cnt && lane < 16) cnt[lane] = (!R->filtered && lane <= R->rmsn) ? (uint16_t)(R->nh[lane] + R->nc[lane]) : (uint16_t)0;
    if (lane != 0) return;
    bsx_rec o;
    o.loc = 0; o.chr = 0; o.nhits = 0; o.nm = 0; o.chain = 0; o.status = (uint8_t)R->filtered; o.len = (uint8_t)R->len;
    if (!R->filtered) {
        int na, ma = 0, ra = 0;
        for (na = 0; na <= R->rmsn; na++) if ((ma = R->nh[na] + R->nc[na]) > 0) break;
        uint2 h = make_uint2(0, 0);
        if (ma) {
            if (ma > 1) ra = (int)(bsx_myrand(R->index, A.randseed) % (uint32_t)ma);
            const int nh = R->nh[na];
            h = (ra < nh) ? hits[((size_t)na * 2) * (A.W + 1) + ra] : hits[((size_t)na * 2 + 1) * (A.W + 1) + (ra - nh)];
        }
        na %= (R->rmsn + 1);
        o.chr = h.x; o.loc = h.y; o.nhits = (uint32_t)ma; o.nm = (uint8_t)na;
        o.chain = (uint8_t)(ra >= (int)R->nh[na]);
    }
    *out = o;
}

// Fix_Unpaired_Short_Fragment (align.cpp:768-791), lane 0
__device__ void fix_unpaired_short(const MapArgs &A, ReadSm *R, uint2 *hits) {
    if (R->len >= A.min_insert) return;
    const size_t W1 = (size_t)A.W + 1;
    #pragma unroll 1
    for (int ii = 0; ii <= R->rmsn; ii++) {
        for (int pass = 0; pass < 2; pass++) {
            uint2 *h = hits + ((size_t)ii * 2 + pass) * W1;
            int cnt = pass ? R->nc[ii] : R->nh[ii];
            #pragma unroll 1
            for (int j = 0; j < cnt; j++) {
                const int sl = ccgg_seglen(A, h[j].x, h[j].y, R->len);
                if (sl < A.min_insert || sl > A.max_insert) { cnt--; for (int k = j; k < cnt; k++) h[k] = h[k + 1]; j--; }
            }
            if (pass) R->nc[ii] = (uint16_t)cnt; else R->nh[ii] = (uint16_t)cnt;
        }
        if (R->nh[ii] + R->nc[ii] > 0) break;
    }
}

__global__ void __launch_bounds__(BSX_WARPS_PER_CTA * 32, 4)
bsx_map_pe_kernel(const __grid_constant__ MapArgs A) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const size_t read_sm = bsx_read_smem_bytes(A.plan_cap, A.nslot);
    const size_t per_warp = 2 * read_sm + sizeof(SelSm);
    CtaSm *K = reinterpret_cast<CtaSm *>(smem);
    init_cta_tables(A, K);
    uint8_t *base = smem + sizeof(CtaSm) + per_warp * wid;
    ReadSm *Ra = reinterpret_cast<ReadSm *>(base);
    ReadSm *Rb = reinterpret_cast<ReadSm *>(base + read_sm);
    SelSm *X = reinterpret_cast<SelSm *>(base + 2 * read_sm);
    const uint32_t gw = blockIdx.x * BSX_WARPS_PER_CTA + wid;
    uint2 *hits_a = A.hit_scratch + (size_t)gw * 2 * A.hit_stride, *hits_b = hits_a + A.hit_stride;
    uint32_t *dd_a = A.dd_scratch + (size_t)gw * 2 * A.dd_stride, *dd_b = dd_a + A.dd_stride;
    PairHitDev *pairs = reinterpret_cast<PairHitDev *>(A.pair_scratch + (size_t)gw * A.pair_stride);
    uint16_t *npairs = X->npairs;
    Ctr *C = X->ctr;
    if (lane < 8) C[lane] = 0;
    __syncwarp();
    const size_t W1 = (size_t)A.W + 1;
    for (;;) {
        uint32_t r = 0;
        if (lane == 0) r = atomicAdd(A.work_counter, 1u);
        r = __shfl_sync(BSX_FULL, r, 0);
        if (r >= A.n) break;
        if ((C[CT_CAND] | C[CT_LIST]) & 0x80000000u) flush_counters(A, C, lane);
        __syncwarp();
        if (lane == 0) {
            Ra->rmsn = Rb->rmsn = 0; Ra->seedseg = Rb->seedseg = 0; Ra->dn = Rb->dn = 0; Ra->best = Rb->best = 99;
            Ra->nw = Rb->nw = 0; Ra->thres = Rb->thres = 0; Ra->fc = Ra->cc = Rb->fc = Rb->cc = 0;
        }
        __syncwarp();
        load_read(A, Ra,  A.seq_a, A.len_a, r, 1, lane);
        load_read(A, Rb,  A.seq_b, A.len_b, r, 2, lane);
        WSET(Ra->filtered, filter_read(A, Ra, lane));
        WSET(Rb->filtered, filter_read(A, Rb, lane));
        if (lane < 16) { Ra->nh[lane] = Ra->nc[lane] = 0; Rb->nh[lane] = Rb->nc[lane] = 0; }
        __syncwarp();
        int paired = 0;
        bsx_pair_rec po;
        po.a_loc = po.a_chr = po.b_loc = po.b_chr = 0; po.insert = 0; po.npairs = 0; po.na = po.nb = po.chain = po.paired = 0;
        if (!Ra->filtered && !Rb->filtered) {
            // PairAlign::RunAlign (pairs.cpp:137-190)
            prepare_read(A, K, Ra, X,  lane, C, nullptr);
            prepare_read(A, K, Rb, X,  lane, C, nullptr);
            if (lane < 31) npairs[lane] = 0;
            __syncwarp();
            const int maxi = max(Ra->rmsn, Rb->rmsn);
            #pragma unroll 1
            for (int i = 0; i <= maxi && !paired; i++) {
                if (i < Ra->seedseg) snp_align(A, Ra, X,  hits_a, dd_a, 1, i, lane, C);
                if (i < Rb->seedseg) snp_align(A, Rb, X,  hits_b, dd_b, 1, i, lane, C);
                if (i <= Ra->rmsn) { sort_hits(hits_a + ((size_t)i * 2) * W1, Ra->nh[i], lane); sort_hits(hits_a + ((size_t)i * 2 + 1) * W1, Ra->nc[i], lane); }
                if (i <= Rb->rmsn) { sort_hits(hits_b + ((size_t)i * 2) * W1, Rb->nh[i], lane); sort_hits(hits_b + ((size_t)i * 2 + 1) * W1, Rb->nc[i], lane); }
                __syncwarp();
                int n = 0;
                if (lane == 0) {
                    n = get_pairs(A, Ra, Rb, hits_a, hits_b, pairs, npairs, i, i);
                    for (int j = 0; j < i; j++) { n += get_pairs(A, Ra, Rb, hits_a, hits_b, pairs, npairs, i, j);
                                                  n += get_pairs(A, Ra, Rb, hits_a, hits_b, pairs, npairs, j, i); }
                }
                n = __shfl_sync(BSX_FULL, n, 0);
                if (n > 0) paired = i + 1;
            }
            __syncwarp();
            if (paired && lane == 0) {
                // StringAlignPair (pairs.cpp:222-242)
                #pragma unroll 1
                for (int i = 0; i <= A.v * 2; i++) {
                    const int np = npairs[i];
                    if (!np) continue;
                    int j = -1;
                    if (np == 1) j = 0;
                    else if (A.r == 1) j = (int)(bsx_myrand(Ra->index, A.randseed) % (uint32_t)np);
                    if (j >= 0) {
                        const PairHitDev ph = pairs[(size_t)i * W1 + j];
                        po.a_chr = ph.a_chr; po.a_loc = ph.a_loc; po.b_chr = ph.b_chr; po.b_loc = ph.b_loc; po.insert = ph.insert;
                        po.npairs = (uint32_t)np; po.chain = (uint8_t)(ph.meta & 0xff); po.na = (uint8_t)((ph.meta >> 8) & 0xff);
                        po.nb = (uint8_t)((ph.meta >> 16) & 0xff); po.paired = 1;
                    }
                    break;
                }
            }
        } else {
            if (!Ra->filtered) run_align(A, K, Ra, X,  hits_a, dd_a, 1, lane, C, nullptr);
            if (!Rb->filtered) run_align(A, K, Rb, X,  hits_b, dd_b, 1, lane, C, nullptr);
        }
        const int out_paired = __shfl_sync(BSX_FULL, (int)po.paired, 0);
        if (!out_paired && BSX_RRBS(A)) {
            if (lane == 0) { if (!Ra->filtered) fix_unpaired_short(A, Ra,  hits_a); if (!Rb->filtered) fix_unpaired_short(A, Rb,  hits_b); }
            __syncwarp();
        }
        if (lane == 0) A.out_pair[r] = po;
        write_unpaired(A, Ra,  hits_a, A.out_a + r, A.cnt_a ? A.cnt_a + (size_t)r * 16 : nullptr, lane);
        write_unpaired(A, Rb,  hits_b, A.out_b + r, A.cnt_b ? A.cnt_b + (size_t)r * 16 : nullptr, lane);
        if (out_paired) CTR_ADD(C, CT_MAPPED, 1);
        __syncwarp();
    }
    flush_counters(A, C, lane);
}

#endif  // BSX_BUILD_PE

}  // namespace

// resident CTAs per SM for the persistent grid (0 when the kernel cannot launch with `smem`), and the launchers
#ifdef BSX_BUILD_SE
int BSX_SE_OCC(size_t smem) {
    int occ = 0;
    cudaError_t e = cudaFuncSetAttribute(BSX_SE_KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, BSX_SE_KERNEL, BSX_WARPS_PER_CTA * 32, smem);
    if (e != cudaSuccess) { bsx_set_error("occupancy query failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return 0; }
    return occ;
}
int BSX_SE_LAUNCH(const MapArgs &a, int n_ctas, cudaStream_t st) {
    const size_t smem = bsx_cta_smem_bytes(1, a.plan_cap, a.nslot);
    static size_t configured = 0;
    if (smem > configured) {
        BSX_CUDA_CHECK(cudaFuncSetAttribute(BSX_SE_KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    BSX_SE_KERNEL<<<n_ctas, BSX_WARPS_PER_CTA * 32, smem, st>>>(a);
    BSX_CUDA_CHECK(cudaGetLastError());
    return BSX_OK;
}
#endif

#ifdef BSX_BUILD_PE
int bsx_map_occupancy_pe(size_t smem) {
    int occ = 0;
    cudaError_t e = cudaFuncSetAttribute(bsx_map_pe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bsx_map_pe_kernel, BSX_WARPS_PER_CTA * 32, smem);
    if (e != cudaSuccess) { bsx_set_error("occupancy query failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return 0; }
    return occ;
}
int bsx_launch_map_pe(const MapArgs &a, int n_ctas, cudaStream_t st) {
    const size_t smem = bsx_cta_smem_bytes(2, a.plan_cap, a.nslot);
    static size_t configured = 0;
    if (smem > configured) {
        BSX_CUDA_CHECK(cudaFuncSetAttribute(bsx_map_pe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    bsx_map_pe_kernel<<<n_ctas, BSX_WARPS_PER_CTA * 32, smem, st>>>(a);
    BSX_CUDA_CHECK(cudaGetLastError());
    return BSX_OK;
}
#endif

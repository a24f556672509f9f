// bsx_common.cuh -- primitives shared by the index-build and mapping kernels (sm_100a).
//
// Bit conventions follow the reference so that results are bit-exact:
//   * bases are 2-bit codes A=0 C=1 G=2 T=3 (param.cpp:187-231, default -M TC), 16 bases per u32,
//     first base in bits 31:30 (dbseq.cpp:71-80);
//   * a seed key is the base-3 reading of the T->C collapsed seed (Param::XT, param.h:123);
//   * a mismatch is read!=ref except read T vs ref C (Param::XC64/XM64, param.h:126,139-147).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define BSX_FULL 0xffffffffu
#define BSX_SEGLEN 16
#define BSX_REF_MARGIN 400u      // dbseq.h:15 REF_MARGIN (u32 words each side of refcat/crefcat)
#define BSX_FIXWORDS 10          // FIXELEMENT for READ_144 (param.h:23-25)

#if defined(__CUDACC__)
#define BSX_HD __host__ __device__ __forceinline__
#else
#define BSX_HD inline
#endif

// forward-strand code of an ASCII base: alphabet[] (param.cpp:210-213): c/C g/G t/T, all else 0
BSX_HD uint32_t bsx_code_fwd(uint8_t c) {
    c |= 0x20;
    return c == 'c' ? 1u : c == 'g' ? 2u : c == 't' ? 3u : 0u;
}
// reverse-strand code: rev_alphabet[] (param.cpp:215-218): c/C->G, g/G->C, t/T->A, all else T(3)
BSX_HD uint32_t bsx_code_rev(uint8_t c) {
    c |= 0x20;
    return c == 'c' ? 2u : c == 'g' ? 1u : c == 't' ? 0u : 3u;
}
// reg_alphabet[] != 0 (param.cpp:153-163): exactly ACGTacgt
BSX_HD uint32_t bsx_is_acgt(uint8_t c) {
    c |= 0x20;
    return (c == 'a' || c == 'c' || c == 'g' || c == 't') ? 1u : 0u;
}

// Param::XT: collapse T(11)->C(01), then read the 2-bit fields of v (right-aligned, leading fields zero)
// as base-3 digits, first base most significant.  Pairwise reduction: 2-bit digits -> 4-bit (x3) ->
// 8-bit (x9) -> 16-bit (x81) -> 32-bit (x6561), i.e. the reference's two 8-base table lookups
// (_T[lo16] + 6561*_T[hi16], param.h:123) without the table.
BSX_HD uint32_t bsx_xt(uint32_t v, int /*nbases*/) {
    v &= ~((v & (v << 1)) & 0xAAAAAAAAu);   // clear the high bit where both bits are set
    v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u) * 3u;
    v = (v & 0x0F0F0F0Fu) + ((v >> 4) & 0x0F0F0F0Fu) * 9u;
    v = (v & 0x00FF00FFu) + ((v >> 8) & 0x00FF00FFu) * 81u;
    return (v & 0xFFFFu) + (v >> 16) * 6561u;
}

// mismatches of one 16-base word: q = read word, m5 = valid-base mask (01 per valid base),
// s = reference word aligned to the read.  Equivalent to XM(((q & XC(s)) ^ s) & r).
BSX_HD uint32_t bsx_mm_word_bits(uint32_t q, uint32_t m5, uint32_t s) {
    uint32_t xc = ((~s) << 1) | s | 0x55555555u;
    uint32_t t = (q & xc) ^ s;
    return (t | (t >> 1)) & m5;
}

// Param::InitMapping (param.cpp:85-93): profile[n][i].a = roundup(n*s + i, I)  (bit8_t)
BSX_HD int bsx_profile_a(int s, int I, int n, int i) {
    return (int)(uint8_t)(((n * s + i + I - 1) / I) * I);
}

// myrand (utilities.cpp:40-50) for -S != 0: stateless mix of the read index and the seed;
// `randseed*1000000` is 32-bit int arithmetic in the reference (App. B Q16)
BSX_HD uint32_t bsx_myrand(uint32_t index, int32_t randseed) {
    int32_t k = (int32_t)((uint32_t)randseed * 1000000u);
    uint64_t v = ((uint64_t)(int64_t)(int32_t)index + (uint64_t)(int64_t)k) * 3935559000370003845ULL + 2691343689449507681ULL;
    v ^= v >> 21; v ^= v << 37; v ^= v >> 4;
    v *= 4768777513237032717ULL;
    v ^= v << 20; v ^= v >> 41; v ^= v << 5;
    return (uint32_t)(v & 0xffffffffULL);
}

#define BSX_CUDA_CHECK(expr)                                                                   \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            bsx_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return BSX_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

void bsx_set_error(const char *fmt, ...);

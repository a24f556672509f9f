// bsx_internal.h -- host-side structures behind the opaque C-ABI handles.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/bsmap_b200.h"

struct bsx_block { uint32_t id, begin, end; };   // Block (dbseq.h:31-36)

// RefSeq (dbseq.h:59-114) as resident on one device
struct bsx_index {
    int device = -1;
    bsx_params par{};
    uint32_t n_seq = 0;
    std::vector<std::string> names;
    std::vector<uint32_t> size;        // title[2k].size
    std::vector<uint32_t> rc_offset;   // title[2k].rc_offset = 16 * nwords   (dbseq.cpp:225)
    std::vector<uint32_t> nwords;      // bfa[2k].n = ceil(len/16) + 2          (dbseq.cpp:60)
    std::vector<uint32_t> anchor;      // ref_anchor, n_seq + 1                  (dbseq.cpp:253-256)
    uint64_t n_words = 0, n_keys = 0, n_entries = 0;
    double build_seconds = 0;
    // device arrays
    uint32_t *d_refcat = nullptr, *d_crefcat = nullptr;   // 2-bit packed strands, margins zeroed
    uint32_t *d_tab = nullptr;      // 2*n_keys+1: [2k] list start, [2k+1] start of rc part, [2k+2] end
    uint32_t *d_pos = nullptr;      // n_entries positions (ref_anchor + p), lists fwd-ascending then rc-ascending
    uint2 *d_ctx = nullptr;         // WGBS: per entry the 16 reference bases before the seed (.x) and the 16 after it (.y)
    uint32_t *d_tag = nullptr;      // RRBS: Hit.chr tag per entry
    uint32_t *d_seqinfo = nullptr;  // anchor[n_seq+1] | size[n_seq] | rc_offset[n_seq]
    uint32_t *d_sites = nullptr;    // RRBS: all digestion sites, concatenated
    uint32_t *d_site_off = nullptr; // RRBS: n_seq+1 offsets into d_sites
    // host copies used by the text formatter
    std::vector<uint32_t> h_refcat;                 // Watson strands (XR:Z / BSP refseq column)
    std::vector<std::vector<uint32_t>> sites;       // RRBS CCGG_sites (dbseq.cpp:158-163)
};

struct bsx_mapper;

int bsx_index_build_device(bsx_index *ix, const char *const *seqs);   // bsx_index.cu
int bsx_index_alloc_device(bsx_index *ix);                            // shell: allocate device arrays
void bsx_index_free_device(bsx_index *ix);
void bsx_set_error(const char *fmt, ...);

/* bsmap_oracle.h -- TEST INFRASTRUCTURE ONLY (see bsmap_oracle.c header).
 *
 * Plain-C CPU restatement of BSMAP 2.6's hot path: reference packing + seed index
 * (dbseq.cpp), read filter/pack/seed selection/probe/extension/selection (align.cpp, align.h),
 * paired-end pairing (pairs.cpp) and SAM/BSP record text (align.cpp:631-765, pairs.cpp:288-498).
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may use this library.
 */
#ifndef BSMAP_ORACLE_H
#define BSMAP_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSO_MAXSNPS 15
#define BSO_MAXHITS 1000
#define BSO_MAX_ADAPTERS 10

typedef struct {
    int32_t seed_size;          /* -s  (param.h:88)  */
    int32_t index_interval;     /* -I  (param.h:117) */
    int32_t max_snp_num;        /* -v  */
    int32_t max_num_hits;       /* -w  */
    int32_t report_repeat_hits; /* -r  */
    int32_t min_insert;         /* -m  */
    int32_t max_insert;         /* -x  */
    int32_t chains;             /* -n  */
    int32_t pairend;            /* set by -b */
    int32_t rrbs;               /* set by -D */
    int32_t randseed;           /* -S  */
    int32_t max_ns;             /* -f  */
    int32_t max_readlen;        /* -L  */
    int32_t out_sam;            /* 1 if -o ends in .sam */
    int32_t out_unmap;          /* -u  */
    int32_t out_ref;            /* -R  */
    int32_t digest_pos;         /* position of '-' in -D */
    int32_t n_adapter;
    char    digest_site[32];    /* -D with the '-' removed */
    char    adapter[BSO_MAX_ADAPTERS][64];
} bso_params;

/* one record per read: what StringAlign (align.cpp:610-627) decides, before text formatting */
typedef struct {
    uint32_t loc;     /* Hit.loc: 0-based start on the Watson strand */
    uint32_t chr;     /* Hit.chr: 2*k + strand */
    uint32_t nhits;   /* hits in the lowest non-empty mismatch bucket (0 = none) */
    uint8_t  nm;      /* that bucket's mismatch count */
    uint8_t  chain;   /* 0: hits[] (read as is), 1: chits[] (reverse-complemented read) */
    uint8_t  status;  /* 0 = aligned attempt, 1 = filtered by FilterReads (QC) */
    uint8_t  len;     /* read length after trimming */
} bso_rec;

typedef struct {
    uint32_t a_loc, a_chr, b_loc, b_chr;
    int32_t  insert;
    uint32_t npairs;  /* pairs in the lowest non-empty total-mismatch bucket */
    uint8_t  na, nb, chain;
    uint8_t  paired;  /* 1: s_OutHitPair path taken; 0: fall back to the two unpaired records */
} bso_pair_rec;

typedef struct bso_ref bso_ref;

bso_ref *bso_ref_create(const bso_params *p, int n_seq, const char *const *names,
                        const char *const *seqs, const uint32_t *lens);
/* CPU-baseline helper: adopt an already built index (arrays are borrowed, not copied; the caller keeps
 * them alive).  Used by bench.py so the CPU leg times MAPPING on the full-size workload without the
 * reference's 250-300 s single-threaded table build; tests prove the imported arrays are bit-identical
 * to what bso_ref_create builds (tests/test_gpu_parity.py::test_index_matches_oracle). */
bso_ref *bso_ref_import(const bso_params *p, int n_seq, const char *const *names, const uint32_t *lens,
                        const uint32_t *refcat, const uint32_t *crefcat, const uint32_t *tab, const uint32_t *pos,
                        uint64_t n_entries);
void bso_ref_destroy(bso_ref *r);

/* introspection for index parity tests */
uint64_t bso_ref_n_words(const bso_ref *r);          /* words in refcat incl. margins */
uint64_t bso_ref_n_keys(const bso_ref *r);
uint64_t bso_ref_n_entries(const bso_ref *r);
const uint32_t *bso_ref_refcat(const bso_ref *r);
const uint32_t *bso_ref_crefcat(const bso_ref *r);
const uint32_t *bso_ref_anchor(const bso_ref *r);    /* n_seq+1 */
const uint32_t *bso_ref_tab(const bso_ref *r);       /* 2*n_keys+1: [2k]=list start, [2k+1]=rc start */
const uint32_t *bso_ref_pos(const bso_ref *r);       /* WGBS entries */
const uint32_t *bso_ref_pos_tag(const bso_ref *r);   /* RRBS: Hit.chr tag per entry (else NULL) */

/* work counters: [0]=CountMismatch calls (C), [1]=distinct list headers (P),
 * [2]=reference's own header probes, [3]=reads mapped */
int bso_map_se(const bso_ref *r, const bso_params *p, uint32_t n, const char *seqs, uint32_t stride,
               const uint16_t *lens, uint32_t first_index, int readset,
               bso_rec *out, uint16_t *counts /* n*16 or NULL */, uint64_t *stats /* 4 or NULL */);

int bso_map_pe(const bso_ref *r, const bso_params *p, uint32_t n,
               const char *seqs_a, const char *seqs_b, uint32_t stride,
               const uint16_t *lens_a, const uint16_t *lens_b, uint32_t first_index,
               bso_pair_rec *out, bso_rec *out_a, bso_rec *out_b,
               uint16_t *counts_a, uint16_t *counts_b, uint64_t *stats);

/* text: returns bytes written (excluding NUL), or the needed size if > cap */
size_t bso_format_header(const bso_ref *r, char *out, size_t cap);
size_t bso_format_se(const bso_ref *r, const bso_params *p, uint32_t n, const char *const *names,
                     const char *const *seqs, const char *const *quals, int readset,
                     const bso_rec *recs, const uint16_t *counts, char *out, size_t cap,
                     uint32_t *n_aligned);
size_t bso_format_pe(const bso_ref *r, const bso_params *p, uint32_t n,
                     const char *const *names_a, const char *const *seqs_a, const char *const *quals_a,
                     const char *const *names_b, const char *const *seqs_b, const char *const *quals_b,
                     const bso_pair_rec *pr, const bso_rec *ra, const bso_rec *rb,
                     const uint16_t *counts_a, const uint16_t *counts_b,
                     char *out, size_t cap, char *out_unpair, size_t cap_unpair, size_t *n_unpair,
                     uint32_t *n_stats /* pairs, a, b */);

/* unit-level helpers (KATs) */
uint32_t bso_xt(uint32_t packed);                       /* Param::XT, param.h:123 */
uint32_t bso_pack16(const char *s);                     /* 16 chars -> word, dbseq.cpp:73-76 */
uint32_t bso_mismatch_cell(uint32_t readbase, uint32_t refbase);
uint32_t bso_myrand(int32_t index, int32_t randseed);   /* utilities.cpp:40-50 */
int      bso_profile_a(int seed_size, int index_interval, int n, int i); /* param.cpp:85-93 */

#ifdef __cplusplus
}
#endif
#endif

/* bsmap_oracle.c -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * A plain-C, single-threaded CPU restatement of the BSMAP 2.6 hot path, written from the
 * behaviour of the reference (file:line citations are into the reference tree).  It exists so
 * the CUDA path has a bit-exact checker that travels to the GPU box, and so bench.py has a CPU
 * baseline ("port").  Parity status: PINNED -- tests/test_oracle_vs_reference.py compares its SAM
 * and BSP text byte-for-byte with files produced by the unmodified reference binary
 * (oracle/_ref/bsmap, built by oracle/Makefile; fixtures in tests/golden/), and
 * tests/test_oracle_kats.py checks the known-answer vectors taken from the reference's object
 * code.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (bsmap_b200/csrc) never links or calls it.
 *
 * Deliberate, documented deviations from undefined behaviour in the reference:
 *   Q4  seed_start_offset when (L-I+1)%s==0: defined as 0.
 *   Q5  refcat/crefcat margins: zero-filled (what fresh pages give the reference in practice).
 *   Q20 CCGG_seglen past-the-end read: clamped to the last site.
 */
#include "bsmap_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

#define SEGLEN 16
#define FIXELEMENT 10            /* READ_144, param.h:23-25 */
#define FIXSIZE (SEGLEN * FIXELEMENT)
#define REF_MARGIN 400           /* dbseq.h:15 */
#define MAXSNPS BSO_MAXSNPS
#define MAXHITS BSO_MAXHITS

typedef struct { uint32_t id, begin, end; } block_t;
typedef struct { uint32_t chr, loc; } hit_t;

struct bso_ref {
    int n_seq;
    char **name;
    uint32_t *size;        /* title[2k].size */
    uint32_t *rc_offset;   /* title[2k].rc_offset = n*16 */
    uint32_t *nwords;      /* bfa[2k].n */
    uint32_t *anchor;      /* ref_anchor, n_seq+1 */
    uint64_t n_words;
    uint32_t *refcat, *crefcat;
    block_t *blocks; size_t n_blocks;
    uint64_t n_keys, n_entries;
    uint32_t *tab;         /* 2*n_keys+1 */
    uint32_t *pos;
    uint32_t *tag;         /* RRBS only */
    /* RRBS */
    uint32_t **sites; uint32_t *n_sites;
    int rrbs; int seed_size; int site_len; int digest_pos;
    int borrowed;          /* arrays belong to the caller (bso_ref_import) */
};

/* ---------------------------------------------------------------- tables (param.cpp:139-231) */
static uint8_t T_alpha[256], T_rev[256], T_reg[256]; static char T_revchar[256];
static int tables_ready = 0;
static void init_tables(void) {
    if (tables_ready) return;
    memset(T_alpha, 0, 256); memset(T_rev, 3, 256); memset(T_reg, 0, 256);
    memset(T_revchar, 'N', 256);
    T_alpha['c'] = T_alpha['C'] = 1; T_alpha['g'] = T_alpha['G'] = 2; T_alpha['t'] = T_alpha['T'] = 3;
    T_rev['c'] = T_rev['C'] = 2; T_rev['g'] = T_rev['G'] = 1; T_rev['t'] = T_rev['T'] = 0;
    const char *u = "ACGTacgt", *v = "TGCAtgca";
    for (int i = 0; i < 8; i++) { T_reg[(uint8_t)u[i]] = 3; T_revchar[(uint8_t)u[i]] = v[i]; }
    tables_ready = 1;
}
static const char USEFUL_NT[] = "ACGTacgt";

/* Param::XT (param.h:123, param.cpp:122-137): T->C, then read the 2-bit fields as base-3 digits */
uint32_t bso_xt(uint32_t tt) {
    uint32_t key = 0;
    for (int j = 15; j >= 0; j--) {
        uint32_t d = (tt >> (2 * j)) & 3u;
        if (d == 3) d = 1;
        key = key * 3 + d;
    }
    return key;
}
uint32_t bso_pack16(const char *s) {
    init_tables();
    uint32_t w = 0;
    for (int j = 0; j < 16; j++) w = (w << 2) | T_alpha[(uint8_t)s[j]];
    return w;
}
/* XC64 / XM64 (param.h:126,139-147) on one 64-bit word */
static inline uint32_t mm64(uint64_t q, uint64_t r, uint64_t s) {
    uint64_t xc = ((~s) << 1) | s | 0x5555555555555555ULL;
    uint64_t t = ((q & xc) ^ s) & r;
    t = (t | (t >> 1)) & 0x5555555555555555ULL;
    return (uint32_t)__builtin_popcountll(t);
}
uint32_t bso_mismatch_cell(uint32_t q, uint32_t s) { return mm64(q & 3, 3, s & 3); }

uint32_t bso_myrand(int32_t i, int32_t randseed) {
    /* utilities.cpp:40-50; `param.randseed*1000000` is int arithmetic (wraps, App. B Q16) */
    int32_t k = (int32_t)((uint32_t)randseed * 1000000u);
    uint64_t v = ((uint64_t)(int64_t)i + (uint64_t)(int64_t)k) * 3935559000370003845ULL + 2691343689449507681ULL;
    v ^= v >> 21; v ^= v << 37; v ^= v >> 4;
    v *= 4768777513237032717ULL;
    v ^= v << 20; v ^= v >> 41; v ^= v << 5;
    return (uint32_t)(v & 0xffffffffULL);
}
int bso_profile_a(int s, int I, int n, int i) { return (uint8_t)(((n * s + i + I - 1) / I) * I); }

/* ---------------------------------------------------------------- reference (dbseq.cpp) */
static inline uint32_t make_seed(const uint32_t *m, uint32_t p, int s) {
    /* RefSeq::s_MakeSeed_1 (dbseq.cpp:286-291) at base offset p of strand array m */
    const uint32_t *w = m + p / SEGLEN;
    int a = 64 - 2 * s - 2 * (int)(p % SEGLEN);
    uint64_t v = (((uint64_t)w[0] << 32) | w[1]) >> a;
    uint32_t bits = (s == 16) ? 0xffffffffu : ((1u << (2 * s)) - 1);
    return bso_xt((uint32_t)v & bits);
}

static int block_cmp(const void *a, const void *b) {
    const block_t *x = a, *y = b;
    if (x->id != y->id) return x->id < y->id ? -1 : 1;
    if (x->begin != y->begin) return x->begin < y->begin ? -1 : 1;
    return 0;
}

typedef struct { uint32_t *v; size_t n, cap; } vec32;
static void vpush(vec32 *a, uint32_t x) {
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 64; a->v = realloc(a->v, a->cap * 4); }
    a->v[a->n++] = x;
}

bso_ref *bso_ref_create(const bso_params *p, int n_seq, const char *const *names,
                        const char *const *seqs, const uint32_t *lens) {
    init_tables();
    bso_ref *r = calloc(1, sizeof *r);
    const int s = p->seed_size, I = p->index_interval;
    r->n_seq = n_seq; r->rrbs = p->rrbs; r->seed_size = s;
    r->site_len = (int)strlen(p->digest_site); r->digest_pos = p->digest_pos;
    r->name = calloc(n_seq, sizeof(char *));
    r->size = calloc(n_seq, 4); r->rc_offset = calloc(n_seq, 4); r->nwords = calloc(n_seq, 4);
    r->anchor = calloc(n_seq + 1, 4);
    uint64_t tot = 0;
    for (int k = 0; k < n_seq; k++) {
        r->name[k] = strdup(names[k]);
        r->size[k] = lens[k];
        r->nwords[k] = (lens[k] + SEGLEN - 1) / SEGLEN + 2;       /* dbseq.cpp:60 */
        r->rc_offset[k] = r->nwords[k] * SEGLEN;                 /* dbseq.cpp:225 */
        r->anchor[k] = (uint32_t)((tot + REF_MARGIN) * SEGLEN);  /* dbseq.cpp:253-256 */
        tot += r->nwords[k];
    }
    r->anchor[n_seq] = (uint32_t)((tot + REF_MARGIN) * SEGLEN);
    r->n_words = tot + 2 * REF_MARGIN;
    r->refcat = calloc(r->n_words, 4); r->crefcat = calloc(r->n_words, 4);

    size_t bcap = 16; r->blocks = malloc(bcap * sizeof(block_t)); r->n_blocks = 0;
    const int max_seedseg_num = (FIXELEMENT - 1) * 16 / s;       /* dbseq.cpp:217 */
    vec32 *cidx = NULL;  /* CCGG_index[j][chr] flattened: j*(2*n_seq)+chr */
    if (p->rrbs) {
        cidx = calloc((size_t)max_seedseg_num * 2 * n_seq, sizeof(vec32));
        r->sites = calloc(n_seq, sizeof(uint32_t *)); r->n_sites = calloc(n_seq, 4);
    }

    for (int k = 0; k < n_seq; k++) {
        const uint32_t len = lens[k], n = r->nwords[k], T = n * SEGLEN;
        const uint8_t *sq = (const uint8_t *)seqs[k];
        uint32_t *f = r->refcat + r->anchor[k] / SEGLEN, *c = r->crefcat + r->anchor[k] / SEGLEN;
        /* BinSeq (dbseq.cpp:58-83): pad with 'N' (code 0) to n*16; cBinSeq (85-111): reverse
           complement of the PADDED buffer, non-CGT -> 3 */
        for (uint32_t i = 0; i < n; i++) {
            uint32_t wf = 0, wc = 0;
            for (uint32_t j = 0; j < SEGLEN; j++) {
                uint32_t pf = i * SEGLEN + j, pc = T - 1 - i * SEGLEN - j;
                wf = (wf << 2) | (pf < len ? T_alpha[sq[pf]] : 0);
                wc = (wc << 2) | (pc < len ? T_rev[sq[pc]] : 3);
            }
            f[i] = wf; c[i] = wc;
        }
        /* UnmaskRegion (dbseq.cpp:114-142): maximal runs terminated by NXnx (or the end), >=30 nt,
           starting at the first ACGTacgt; never merged (App. B Q6) */
        uint32_t e = 0;
        while (e < len) {
            uint32_t b = e;
            while (b < len && !T_reg[sq[b]]) b++;
            if (b >= len) break;
            e = b;
            while (e < len && !(sq[e] == 'N' || sq[e] == 'X' || sq[e] == 'n' || sq[e] == 'x')) e++;
            if (e - b < 30) continue;
            if (r->n_blocks + 2 > bcap) { bcap *= 2; r->blocks = realloc(r->blocks, bcap * sizeof(block_t)); }
            r->blocks[r->n_blocks++] = (block_t){ (uint32_t)(2 * k), b, e };
            r->blocks[r->n_blocks++] = (block_t){ (uint32_t)(2 * k + 1), T - e, T - b };
        }
        if (p->rrbs) {
            /* find_CCGG (dbseq.cpp:144-211) on the upper-cased sequence */
            const int sl = r->site_len, dp = p->digest_pos;
            vec32 st = {0};
            for (uint32_t q = 0; q + sl <= len; q++) {
                int ok = 1;
                for (int t = 0; t < sl; t++) if (toupper(sq[q + t]) != p->digest_site[t]) { ok = 0; break; }
                if (ok) vpush(&st, q + dp);
            }
            r->sites[k] = st.v; r->n_sites[k] = (uint32_t)st.n;
            uint32_t tmp_offset = r->rc_offset[k] - s, tmp_max = len - s;
            for (size_t q = 0; q + 1 < st.n; q++)
                if (st.v[q + 1] - st.v[q] <= (uint32_t)p->max_insert) {
                    uint32_t seedloc = st.v[q];
                    for (int i = 0; i < max_seedseg_num && seedloc <= tmp_max; i++, seedloc += s)
                        vpush(&cidx[(size_t)i * 2 * n_seq + 2 * k], seedloc);
                }
            for (size_t q = 1; q < st.n; q++)
                if (st.v[q] - st.v[q - 1] <= (uint32_t)p->max_insert) {
                    int seedloc = (int)(st.v[q] + sl - 2 * dp - s);
                    for (int i = 0; i < max_seedseg_num && seedloc >= 0; i++, seedloc -= s)
                        vpush(&cidx[(size_t)i * 2 * n_seq + 2 * k + 1], tmp_offset - (uint32_t)seedloc);
                }
        }
    }
    qsort(r->blocks, r->n_blocks, sizeof(block_t), block_cmp);   /* dbseq.cpp:249 */

    uint64_t nk = 1; for (int i = 0; i < s; i++) nk *= 3;          /* dbseq.cpp:314 */
    r->n_keys = nk;
    r->tab = calloc(2 * nk + 1, 4);
    uint32_t *cnt = calloc(nk + 1, 4), *cntf = NULL;

    if (!p->rrbs) {
        /* t_CalKmerFreq_ab (dbseq.cpp:349-359) */
        cntf = calloc(nk + 1, 4);
        for (size_t b = 0; b < r->n_blocks; b++) {
            const block_t *q = &r->blocks[b];
            const uint32_t *m = ((q->id & 1) ? r->crefcat : r->refcat) + r->anchor[q->id / 2] / SEGLEN;
            uint32_t i2 = ((q->end - s) / I) * I;
            for (uint32_t i = (q->begin / I) * I; i <= i2; i += I) {
                uint32_t key = make_seed(m, i, s);
                cnt[key]++; if (!(q->id & 1)) cntf[key]++;
            }
        }
        uint64_t acc = 0;
        for (uint64_t k = 0; k < nk; k++) { r->tab[2 * k] = (uint32_t)acc; r->tab[2 * k + 1] = (uint32_t)(acc + cntf[k]); acc += cnt[k]; }
        r->tab[2 * nk] = (uint32_t)acc; r->n_entries = acc;
        r->pos = malloc((acc ? acc : 1) * 4);
        /* t_CreateIndex_ab (dbseq.cpp:441-480): all forward-strand blocks first, then all rc blocks */
        uint32_t *cur = cnt; /* reuse as cursors */
        for (uint64_t k = 0; k < nk; k++) { cur[k] = r->tab[2 * k]; cntf[k] = r->tab[2 * k + 1]; }
        for (int pass = 0; pass < 2; pass++)
            for (size_t b = 0; b < r->n_blocks; b++) {
                const block_t *q = &r->blocks[b];
                if ((int)(q->id & 1) != pass) continue;
                const uint32_t *m = (pass ? r->crefcat : r->refcat) + r->anchor[q->id / 2] / SEGLEN;
                uint32_t i2 = ((q->end - s) / I) * I;
                for (uint32_t i = (q->begin / I) * I; i <= i2; i += I) {
                    uint32_t key = make_seed(m, i, s);
                    uint32_t *c = pass ? &cntf[key] : &cur[key];
                    r->pos[(*c)++] = r->anchor[q->id / 2] + i;       /* hit2int, dbseq.cpp:570 */
                }
            }
        free(cntf);
    } else {
        /* RRBS (dbseq.cpp:332-347, 418-438): j major, then chr, own sites then mirrored sites */
        const int mirror = p->pairend || p->chains;
        for (int fill = 0; fill < 2; fill++) {
            if (fill) {
                uint64_t acc = 0;
                for (uint64_t k = 0; k < nk; k++) { r->tab[2 * k] = (uint32_t)acc; acc += cnt[k]; cnt[k] = r->tab[2 * k]; }
                r->tab[2 * nk] = (uint32_t)acc; r->n_entries = acc;
                for (uint64_t k = 0; k < nk; k++) r->tab[2 * k + 1] = r->tab[2 * k + 2];
                r->pos = malloc((acc ? acc : 1) * 4); r->tag = malloc((acc ? acc : 1) * 4);
            }
            for (int j = 0; j < max_seedseg_num; j++)
                for (int chr = 0; chr < 2 * n_seq; chr++) {
                    const uint32_t *m = ((chr & 1) ? r->crefcat : r->refcat) + r->anchor[chr / 2] / SEGLEN;
                    const vec32 *v = &cidx[(size_t)j * 2 * n_seq + chr];
                    for (size_t t = 0; t < v->n; t++) {
                        uint32_t key = make_seed(m, v->v[t], s);
                        if (!fill) cnt[key]++;
                        else { r->pos[cnt[key]] = v->v[t]; r->tag[cnt[key]++] = (uint32_t)chr | ((uint32_t)j << 16); }
                    }
                    if (mirror) {
                        const vec32 *v1 = &cidx[(size_t)j * 2 * n_seq + (chr ^ 1)];
                        uint32_t tmp_offset = r->rc_offset[chr / 2] - s;
                        for (size_t t = 0; t < v1->n; t++) {
                            uint32_t loc = tmp_offset - v1->v[t];
                            uint32_t key = make_seed(m, loc, s);
                            if (!fill) cnt[key]++;
                            else { r->pos[cnt[key]] = loc; r->tag[cnt[key]++] = (uint32_t)chr | ((uint32_t)j << 16) | 0x1000000u; }
                        }
                    }
                }
        }
        for (size_t i = 0; i < (size_t)max_seedseg_num * 2 * n_seq; i++) free(cidx[i].v);
        free(cidx);
    }
    free(cnt);
    return r;
}

bso_ref *bso_ref_import(const bso_params *p, int n_seq, const char *const *names, const uint32_t *lens,
                        const uint32_t *refcat, const uint32_t *crefcat, const uint32_t *tab, const uint32_t *pos,
                        uint64_t n_entries) {
    init_tables();
    if (p->rrbs) return NULL;
    bso_ref *r = calloc(1, sizeof *r);
    r->n_seq = n_seq; r->rrbs = 0; r->seed_size = p->seed_size; r->borrowed = 1;
    r->name = calloc(n_seq, sizeof(char *));
    r->size = calloc(n_seq, 4); r->rc_offset = calloc(n_seq, 4); r->nwords = calloc(n_seq, 4);
    r->anchor = calloc(n_seq + 1, 4);
    uint64_t tot = 0;
    for (int k = 0; k < n_seq; k++) {
        r->name[k] = strdup(names[k]); r->size[k] = lens[k];
        r->nwords[k] = (lens[k] + SEGLEN - 1) / SEGLEN + 2; r->rc_offset[k] = r->nwords[k] * SEGLEN;
        r->anchor[k] = (uint32_t)((tot + REF_MARGIN) * SEGLEN); tot += r->nwords[k];
    }
    r->anchor[n_seq] = (uint32_t)((tot + REF_MARGIN) * SEGLEN);
    r->n_words = tot + 2 * REF_MARGIN;
    r->refcat = (uint32_t *)refcat; r->crefcat = (uint32_t *)crefcat; r->tab = (uint32_t *)tab; r->pos = (uint32_t *)pos;
    r->n_keys = 1; for (int i = 0; i < p->seed_size; i++) r->n_keys *= 3;
    r->n_entries = n_entries;
    return r;
}

void bso_ref_destroy(bso_ref *r) {
    if (!r) return;
    if (r->borrowed) { r->refcat = r->crefcat = r->tab = r->pos = NULL; }
    for (int k = 0; k < r->n_seq; k++) { free(r->name[k]); if (r->sites) free(r->sites[k]); }
    free(r->name); free(r->size); free(r->rc_offset); free(r->nwords); free(r->anchor);
    free(r->refcat); free(r->crefcat); free(r->blocks); free(r->tab); free(r->pos); free(r->tag);
    free(r->sites); free(r->n_sites); free(r);
}
uint64_t bso_ref_n_words(const bso_ref *r) { return r->n_words; }
uint64_t bso_ref_n_keys(const bso_ref *r) { return r->n_keys; }
uint64_t bso_ref_n_entries(const bso_ref *r) { return r->n_entries; }
const uint32_t *bso_ref_refcat(const bso_ref *r) { return r->refcat; }
const uint32_t *bso_ref_crefcat(const bso_ref *r) { return r->crefcat; }
const uint32_t *bso_ref_anchor(const bso_ref *r) { return r->anchor; }
const uint32_t *bso_ref_tab(const bso_ref *r) { return r->tab; }
const uint32_t *bso_ref_pos(const bso_ref *r) { return r->pos; }
const uint32_t *bso_ref_pos_tag(const bso_ref *r) { return r->tag; }

/* RefSeq::int2hit (dbseq.cpp:585-595) */
static inline hit_t int2hit(const bso_ref *r, uint32_t p, int c) {
    int left = 0, right = r->n_seq;
    while (left < right - 1) { int mid = (left + right) / 2; if (p >= r->anchor[mid]) left = mid; else right = mid; }
    hit_t h = { (uint32_t)(left * 2 + c), p - r->anchor[left] };
    return h;
}

/* RefSeq::CCGG_seglen (dbseq.cpp:541-567); past-the-end read clamped (App. B Q20) */
static void ccgg_seglen(const bso_ref *r, uint32_t chr, uint32_t pos, int readlen, uint32_t *first, int *second) {
    const uint32_t *st = r->sites[chr / 2]; int n = (int)r->n_sites[chr / 2];
    int left = 0, right = n - 1, mid;
    while (left < right - 1) {
        mid = (left + right) / 2;
        uint32_t mv = st[mid];
        if (mv == pos) { left = mid; right = mid + 1; break; }
        else if (mv < pos) left = mid; else right = mid;
    }
    uint32_t seg_start = st[left], seg_end;
    const uint32_t add = (uint32_t)(r->site_len - 2 * r->digest_pos);
    for (;;) {
        int rr = right < n ? right : n - 1;
        seg_end = st[rr] + add;
        if (seg_end < pos + (uint32_t)readlen && right < n) right++; else break;
    }
    *first = seg_start + 1; *second = (int)(seg_end - seg_start);
}

/* ---------------------------------------------------------------- per-read state (align.h) */
typedef struct {
    const bso_ref *ref; const bso_params *par;
    /* read */
    char seq[FIXSIZE + 16]; int len; int raw_readlen; int readset; uint32_t index;
    int read_max_snp_num; int seedseg_num; uint32_t snp_thres; uint32_t cseed_offset;
    int flag_chain, cflag_chain;
    uint32_t bseq[SEGLEN][FIXELEMENT + 2], reg[SEGLEN][FIXELEMENT + 2];
    uint32_t cbseq[SEGLEN][FIXELEMENT + 2], creg[SEGLEN][FIXELEMENT + 2];
    uint32_t seed_array[FIXSIZE], cseed_array[FIXSIZE];
    uint32_t seeds[MAXSNPS + 1][16], cseeds[MAXSNPS + 1][16];
    int seed_start_offset, cseed_start_offset;
    int seed_start_array[MAXSNPS + 1], cseed_start_array[MAXSNPS + 1];
    int seedindex[MAXSNPS + 1][2], cseedindex[MAXSNPS + 1][2];   /* (sum, segment) sorted */
    int n_hit[MAXSNPS + 1], n_chit[MAXSNPS + 1];
    hit_t (*hits)[MAXHITS + 1], (*chits)[MAXHITS + 1];
    /* dedupe set on (chr>>1, loc) == Watson concatenated coordinate (align.cpp:274 ...) */
    uint32_t *dset; uint32_t dcap, dn; uint32_t *dlist;
    uint64_t n_cand, n_probe_ref;
    uint8_t probed[2][FIXSIZE];
    uint64_t n_probe_distinct;
} sa_t;

static sa_t *sa_new(const bso_ref *r, const bso_params *p) {
    sa_t *a = calloc(1, sizeof *a);
    a->ref = r; a->par = p;
    a->hits = malloc(sizeof(hit_t) * (MAXSNPS + 1) * (MAXHITS + 1));
    a->chits = malloc(sizeof(hit_t) * (MAXSNPS + 1) * (MAXHITS + 1));
    a->dcap = 1u << 16; a->dset = malloc(a->dcap * 4); memset(a->dset, 0xff, a->dcap * 4);
    a->dlist = malloc(((MAXSNPS + 1) * (MAXHITS + 1) + 8) * 4); a->dn = 0;
    return a;
}
static void sa_free(sa_t *a) { free(a->hits); free(a->chits); free(a->dset); free(a->dlist); free(a); }

/* hitset[chr>>1].insert(loc).second  -- key = anchor[chr>>1] + loc (u32, wrapping), which is
   collision-free for every (chr>>1, loc) the reference can produce */
static int dset_insert(sa_t *a, uint32_t key) {
    uint32_t h = (key * 2654435761u) >> 16;
    for (;;) {
        uint32_t v = a->dset[h];
        if (v == 0xffffffffu) break;
        if (v == key) return 0;
        h = (h + 1) & (a->dcap - 1);
    }
    a->dset[h] = key; a->dlist[a->dn++] = h;
    return 1;
}
static void clear_hits(sa_t *a) {                           /* align.cpp:428-433 */
    for (int i = 0; i <= MAXSNPS; i++) a->n_hit[i] = a->n_chit[i] = 0;
    for (uint32_t i = 0; i < a->dn; i++) a->dset[a->dlist[i]] = 0xffffffffu;
    a->dn = 0;
}

static inline uint32_t list_size(const bso_ref *r, uint32_t key) {
    /* index2[key]==NULL ? 0 : index2[key][0]  (= n+2, App. B Q7); RRBS: index[key].n1 */
    uint32_t n = r->tab[2 * key + 2] - r->tab[2 * key];
    if (r->rrbs) return n;
    return n ? n + 2 : 0;
}

/* TrimAdapter (align.cpp:371-425) */
static int trim_adapter(sa_t *a) {
    const bso_params *p = a->par; const char *sq = a->seq;
    a->raw_readlen = a->len;
    const int s = p->seed_size, len = a->len;
    if (p->rrbs) {
        const int sl = (int)strlen(p->digest_site), dp = p->digest_pos;
        for (int i = 0; i < p->n_adapter; i++) {
            const char *ad = p->adapter[i]; const int al = (int)strlen(ad);
            for (int pos = s; pos < len - 5; pos++) {
                int m0 = 0, k;
                for (k = 0; k < al && k < 15 && pos + k < len; k++)
                    if ((m0 += (ad[k] != sq[pos + k])) > 4) break;
                if (k < m0 * 5) continue;
                int m = m0;
                for (int t = 0; t < sl - dp; t++) {
                    char x = p->digest_site[t], y = sq[pos - sl + dp + t];
                    m += (x != y) && (x != 'C' || y != 'T');
                }
                if (k >= m * 5) { a->len = pos; a->seq[pos] = 0; return 1; }
                if (p->pairend) {
                    m = m0;
                    for (int t = 0; t < sl - dp; t++) {
                        char x = p->digest_site[t], y = sq[pos - sl + dp + t];
                        m += (x != y) && (x != 'G' || y != 'A');
                    }
                    if (k >= m * 5) { a->len = pos; a->seq[pos] = 0; return 1; }
                }
            }
        }
    } else {
        for (int i = 0; i < p->n_adapter; i++) {
            const char *ad = p->adapter[i]; const int al = (int)strlen(ad);
            for (int pos = s; pos < len - 4; pos++) {
                int m0 = 0, k;
                for (k = 0; k < al && k < 15 && pos + k < len; k++)
                    if ((m0 += (ad[k] != sq[pos + k])) > 4) break;
                if (k >= m0 * 5 && k > 3) { a->len = pos; a->seq[pos] = 0; return 1; }
            }
        }
    }
    return 0;
}

/* FilterReads (align.cpp:579-589); -q trimming not supported (qual_threshold == 0) */
static int filter_read(sa_t *a) {
    trim_adapter(a);
    if (a->len < a->par->seed_size) return 1;
    int n = 0;
    for (int i = 0; i < a->len; i++) if (!T_reg[(uint8_t)a->seq[i]]) n++;
    if (n > a->par->max_ns) return 1;
    a->read_max_snp_num = (int)((size_t)(a->par->max_snp_num + 1) * (size_t)(a->len - 1) / (size_t)a->raw_readlen);
    return 0;
}

static void right_shift(const uint32_t *o, uint32_t *n) {    /* align.cpp:82-87 */
    n[0] = o[0] >> 2;
    for (int i = 1; i < FIXELEMENT; i++) n[i] = (o[i] >> 2) | (o[i - 1] << 30);
}

/* ConvertBinaySeq (align.cpp:90-162) */
static void convert_binary(sa_t *a) {
    const bso_params *p = a->par; const int s = p->seed_size, len = a->len;
    const uint32_t bits = (s == 16) ? 0xffffffffu : ((1u << (2 * s)) - 1);
    a->flag_chain = p->chains || (a->readset < 2);
    a->cflag_chain = p->chains || (a->readset == 2);
    for (int chain = 0; chain < 2; chain++) {
        if (chain == 0 ? !a->flag_chain : !a->cflag_chain) continue;
        uint32_t (*bs)[FIXELEMENT + 2] = chain ? a->cbseq : a->bseq;
        uint32_t (*rg)[FIXELEMENT + 2] = chain ? a->creg : a->reg;
        uint32_t *sa = chain ? a->cseed_array : a->seed_array;
        memset(bs, 0, sizeof a->bseq); memset(rg, 0, sizeof a->reg);
        uint64_t roll = 0;
        for (int i = 0; i < FIXSIZE; i++) {
            uint32_t code = 0, m = 0;
            if (i < len) {
                uint8_t ch = (uint8_t)(chain ? a->seq[len - 1 - i] : a->seq[i]);
                code = chain ? T_rev[ch] : T_alpha[ch]; m = T_reg[ch];
                roll = (roll << 2) | code;
                if (i + 1 >= s) sa[i + 1 - s] = bso_xt((uint32_t)roll & bits);
            }
            bs[0][i / SEGLEN] |= code << (30 - 2 * (i % SEGLEN));
            rg[0][i / SEGLEN] |= m << (30 - 2 * (i % SEGLEN));
        }
        for (int i = 1; i < SEGLEN; i++) { right_shift(bs[i - 1], bs[i]); right_shift(rg[i - 1], rg[i]); }
    }
}

/* CountMismatch (align.h:167-200), READ_144: 5 x 64-bit words */
static inline uint32_t count_mismatch(sa_t *a, const uint32_t *q, const uint32_t *r, const uint32_t *s) {
    a->n_cand++;
    uint32_t w = 0;
    for (int i = 0; i < 5; i++) {
        uint64_t Q = ((uint64_t)q[2 * i + 1] << 32) | q[2 * i];
        uint64_t R = ((uint64_t)r[2 * i + 1] << 32) | r[2 * i];
        uint64_t S = ((uint64_t)s[2 * i + 1] << 32) | s[2 * i];
        w += mm64(Q, R, S);
        if (i < 2 && w > a->snp_thres) return w;
    }
    return w;
}

static inline int prof_a(const bso_params *p, int n, int i) {
    return bso_profile_a(p->seed_size, p->index_interval, n, i);
}

static inline uint32_t probe(sa_t *a, int chain, int off) {
    const uint32_t key = (chain ? a->cseed_array : a->seed_array)[off];
    a->n_probe_ref++;
    if (!a->probed[chain][off]) { a->probed[chain][off] = 1; a->n_probe_distinct++; }
    return list_size(a->ref, key);
}

/* CountSeeds / CountCSeeds (align.cpp:549-565) */
static int count_seeds(sa_t *a, int chain, int n, int start) {
    int total = 0;
    for (int i = 0; i < a->par->index_interval; i++) total += (int)probe(a, chain, prof_a(a->par, n, i) + start - i);
    return total;
}
static uint32_t total_seed_loc(sa_t *a, int chain, int start) {   /* align.cpp:567-577 */
    int total = 0;
    for (int i = 0; i < a->seedseg_num; i++) total += count_seeds(a, chain, i, start);
    return (uint32_t)total;
}
static void adjust_start_array(sa_t *a, int chain) {               /* align.cpp:506-547 */
    int *arr = chain ? a->cseed_start_array : a->seed_start_array;
    const int off = chain ? a->cseed_start_offset : a->seed_start_offset;
    const bso_params *p = a->par;
    for (int i = 0; i < a->seedseg_num; i++) arr[i] = off;
    if (p->rrbs) return;
    const int max_offset = (a->len - p->index_interval + 1) % p->seed_size;
    for (int i = 0; i < a->seedseg_num; i++) {
        int ptr = (i % 2 == 0) ? i / 2 : a->seedseg_num - 1 - i / 2;
        uint32_t total = 0xffffffffu;
        int start = (ptr == 0) ? 0 : arr[ptr - 1];
        int end = (ptr == a->seedseg_num - 1) ? max_offset : arr[ptr + 1];
        arr[ptr] = start;
        for (uint32_t ii = (uint32_t)start; ii <= (uint32_t)end; ii++) {
            uint32_t tt = (uint32_t)count_seeds(a, chain, ptr, (int)ii);
            if (tt < total) { total = tt; arr[ptr] = (int)ii; }
        }
    }
}
static int pair_cmp(const void *x, const void *y) {
    const int *a = x, *b = y;
    if (a[0] != b[0]) return a[0] < b[0] ? -1 : 1;
    return (a[1] > b[1]) - (a[1] < b[1]);
}
/* ReorderSeed (align.cpp:454-504) + GenerateSeeds (align.h:138-164) */
static void reorder_seed(sa_t *a) {
    const bso_params *p = a->par; const int I = p->index_interval;
    uint32_t total = 0xffffffffu, ctotal = 0xffffffffu;
    a->seed_start_offset = a->cseed_start_offset = 0;   /* App. B Q4: defined as 0 */
    if (!p->rrbs) {
        uint32_t ii = (uint32_t)((a->len - I + 1) % p->seed_size);
        for (uint32_t i = 0; i < ii; i++) {
            if (a->flag_chain) { uint32_t tt = total_seed_loc(a, 0, (int)i); if (tt < total) { total = tt; a->seed_start_offset = (int)i; } }
            if (a->cflag_chain) { uint32_t tt = total_seed_loc(a, 1, (int)i); if (tt < ctotal) { ctotal = tt; a->cseed_start_offset = (int)i; } }
        }
    }
    for (int chain = 0; chain < 2; chain++) {
        if (chain == 0 ? !a->flag_chain : !a->cflag_chain) continue;
        adjust_start_array(a, chain);
        int (*sidx)[2] = chain ? a->cseedindex : a->seedindex;
        uint32_t (*sd)[16] = chain ? a->cseeds : a->seeds;
        const uint32_t *sarr = chain ? a->cseed_array : a->seed_array;
        const int *arr = chain ? a->cseed_start_array : a->seed_start_array;
        for (int n = 0; n < a->seedseg_num; n++) {
            uint32_t sum = 0;
            if (p->rrbs) {
                int off = prof_a(p, n, 0) + arr[n] + (chain ? (int)a->cseed_offset : 0);
                sd[n][0] = sarr[off];
                sum += probe(a, chain, off);
            } else {
                for (int i = 0; i < I; i++) {
                    int off = prof_a(p, n, i) + arr[n] - i;
                    sd[n][i] = sarr[off];
                    sum += probe(a, chain, off);
                }
            }
            sidx[n][0] = (int)sum; sidx[n][1] = n;
        }
        qsort(sidx, a->seedseg_num, sizeof sidx[0], pair_cmp);
    }
}

/* one candidate after CountMismatch passed: bounds, dedupe, bucket append, exits.
   returns 1 if SnpAlign must return (align.cpp:271-278 and twins) */
static inline int commit(sa_t *a, int chain, hit_t h, uint32_t w, int mode, int rrbs_frag_filter) {
    const bso_ref *r = a->ref; const bso_params *p = a->par;
    if (h.chr & 1) h.loc = r->rc_offset[h.chr >> 1] - (uint32_t)a->len - h.loc;
    if (h.loc + (uint32_t)a->len > r->size[h.chr >> 1]) return 0;
    if (!dset_insert(a, r->anchor[h.chr >> 1] + h.loc)) return 0;
    if (rrbs_frag_filter) {
        uint32_t f; int sl; ccgg_seglen(r, h.chr, h.loc, a->len, &f, &sl);
        if (sl > p->max_insert) return 0;
        if (sl < p->min_insert) return 0;
    }
    if (chain) a->chits[w][a->n_chit[w]++] = h; else a->hits[w][a->n_hit[w]++] = h;
    if ((int)w == mode && !p->pairend && p->report_repeat_hits == 0)
        if (a->n_hit[w] + a->n_chit[w] > 1) return 1;
    if (a->n_hit[w] + a->n_chit[w] >= p->max_num_hits) {
        if (w == 0) return 1; else a->snp_thres = w - 1;
    }
    return 0;
}

/* SnpAlign (align.cpp:168-347) */
static void snp_align(sa_t *a, int mode) {
    const bso_ref *r = a->ref; const bso_params *p = a->par;
    for (int chain = 0; chain < 2; chain++) {
        if (chain == 0 ? !a->flag_chain : !a->cflag_chain) continue;
        uint32_t (*bs)[FIXELEMENT + 2] = chain ? a->cbseq : a->bseq;
        uint32_t (*rg)[FIXELEMENT + 2] = chain ? a->creg : a->reg;
        const int modeindex = (chain ? a->cseedindex : a->seedindex)[mode][1];
        if (p->rrbs) {
            const uint32_t key = (chain ? a->cseeds : a->seeds)[modeindex][0];
            const uint32_t b = r->tab[2 * key], e = r->tab[2 * key + 2];
            const uint32_t want = chain ? (uint32_t)(a->len / p->seed_size - 1 - modeindex) : (uint32_t)modeindex;
            const uint32_t h = (uint32_t)prof_a(p, modeindex, 0) + (chain ? a->cseed_offset : 0);
            for (uint32_t j = b; j < e; j++) {
                uint32_t tag = r->tag[j], loc = r->pos[j];
                if (((chain ? (tag ^ 0x1000000u) : tag) >> 16) != want) continue;
                uint32_t chr = tag & 0xffff;
                /* n_cand counts the entries that pass the tag test: those the reference goes on to examine.  The few
                   whose window would start before the sequence are dropped here without a CountMismatch call. */
                if (loc < h) { a->n_cand++; continue; }
                loc -= h;
                const uint32_t *m = ((chr & 1) ? r->crefcat : r->refcat) + r->anchor[chr >> 1] / SEGLEN;
                uint32_t w = count_mismatch(a, bs[loc % SEGLEN], rg[loc % SEGLEN], m + loc / SEGLEN);
                if (w > a->snp_thres) continue;
                hit_t hh = { chr, loc };
                if (commit(a, chain, hh, w, mode, chain == 0 && !p->pairend)) return;
            }
        } else {
            const int *starts = chain ? a->cseed_start_array : a->seed_start_array;
            for (int i = 0; i < p->index_interval; i++) {
                const uint32_t key = (chain ? a->cseeds : a->seeds)[modeindex][i];
                const uint32_t b = r->tab[2 * key], mc = r->tab[2 * key + 1], e = r->tab[2 * key + 2];
                if (b == e) continue;
                const uint32_t h = (uint32_t)(-prof_a(p, modeindex, i) + i - starts[modeindex]);
                for (uint32_t j = b; j < e; j++) {
                    const int strand = j >= mc;
                    uint32_t loc = r->pos[j] + h;
                    const uint32_t *m = strand ? r->crefcat : r->refcat;
                    uint32_t w = count_mismatch(a, bs[loc % SEGLEN], rg[loc % SEGLEN], m + loc / SEGLEN);
                    if (w > a->snp_thres) continue;
                    hit_t hh = int2hit(r, loc, strand);
                    if (commit(a, chain, hh, w, mode, 0)) return;
                }
            }
        }
    }
}

/* SingleAlign::RunAlign (align.cpp:435-452) */
static int run_align(sa_t *a) {
    const bso_params *p = a->par;
    clear_hits(a);
    memset(a->probed, 0, sizeof a->probed);
    a->seedseg_num = (a->len - p->index_interval + 1) / p->seed_size;
    if (a->seedseg_num > a->read_max_snp_num + 1) a->seedseg_num = a->read_max_snp_num + 1;
    convert_binary(a);
    a->snp_thres = (uint32_t)a->read_max_snp_num;
    a->cseed_offset = (uint32_t)(a->len % p->seed_size);
    reorder_seed(a);
    for (int i = 0; i < a->seedseg_num; i++) {
        snp_align(a, i);
        if (!p->rrbs) for (int ii = 0; ii <= i; ii++) if (a->n_hit[ii] || a->n_chit[ii]) return 1;
    }
    for (int i = 0; i <= a->read_max_snp_num; i++) if (a->n_hit[i] || a->n_chit[i]) return 1;
    return 0;
}

static void load_read(sa_t *a, const char *s, int len, int readset, uint32_t index) {
    if (len > a->par->max_readlen) len = a->par->max_readlen;   /* reads.cpp:115-117 */
    if (len > FIXSIZE - 16) len = FIXSIZE - 16;
    memcpy(a->seq, s, len); a->seq[len] = 0; a->len = len; a->readset = readset; a->index = index;
}

/* StringAlign (align.cpp:610-627) -> record */
static void select_hit(sa_t *a, bso_rec *o, uint16_t *counts) {
    memset(o, 0, sizeof *o);
    o->len = (uint8_t)a->len;
    int ii, sum = 0;
    for (ii = 0; ii <= a->read_max_snp_num; ii++) if ((sum = a->n_hit[ii] + a->n_chit[ii]) > 0) break;
    if (counts) for (int i = 0; i <= MAXSNPS; i++) counts[i] = (uint16_t)(i <= a->read_max_snp_num ? a->n_hit[i] + a->n_chit[i] : 0);
    if (sum == 0) { o->nm = (uint8_t)ii; return; }
    int j = (int)(bso_myrand((int32_t)a->index, a->par->randseed) % (uint32_t)sum);
    hit_t h;
    if (j < a->n_hit[ii]) { h = a->hits[ii][j]; o->chain = 0; } else { h = a->chits[ii][j - a->n_hit[ii]]; o->chain = 1; }
    o->loc = h.loc; o->chr = h.chr; o->nhits = (uint32_t)sum; o->nm = (uint8_t)ii;
}

int bso_map_se(const bso_ref *r, const bso_params *p, uint32_t n, const char *seqs, uint32_t stride,
               const uint16_t *lens, uint32_t first_index, int readset,
               bso_rec *out, uint16_t *counts, uint64_t *stats) {
    init_tables();
    sa_t *a = sa_new(r, p);
    uint64_t mapped = 0;
    for (uint32_t t = 0; t < n; t++) {
        load_read(a, seqs + (size_t)t * stride, lens[t], readset, first_index + t);
        if (filter_read(a)) {
            memset(&out[t], 0, sizeof out[t]); out[t].status = 1; out[t].len = (uint8_t)a->len;
            if (counts) memset(counts + (size_t)t * 16, 0, 32);
            continue;
        }
        run_align(a);
        select_hit(a, &out[t], counts ? counts + (size_t)t * 16 : NULL);
        if (out[t].nhits) mapped++;
    }
    if (stats) { stats[0] += a->n_cand; stats[1] += a->n_probe_distinct; stats[2] += a->n_probe_ref; stats[3] += mapped; }
    sa_free(a);
    return 0;
}

/* ---------------------------------------------------------------- paired end (pairs.cpp) */
typedef struct { uint16_t chain; uint8_t na, nb; int insert; hit_t a, b; } pairhit_t;
typedef struct {
    sa_t *sa, *sb; const bso_params *par;
    uint32_t n_pairs[2 * MAXSNPS + 1];
    pairhit_t (*pairhits)[MAXHITS + 1];
} pa_t;

static int hit_cmp(const void *x, const void *y) {               /* HitComp, utilities.cpp:53 */
    const hit_t *a = x, *b = y;
    if (a->chr != b->chr) return a->chr < b->chr ? -1 : 1;
    if (a->loc != b->loc) return a->loc < b->loc ? -1 : 1;
    return 0;
}
static void sort_hits_pe(sa_t *a, int n) {                        /* align.cpp:363-368 */
    qsort(a->hits[n], a->n_hit[n], sizeof(hit_t), hit_cmp);
    qsort(a->chits[n], a->n_chit[n], sizeof(hit_t), hit_cmp);
}

/* GetPairs (pairs.cpp:34-135) */
static int get_pairs(pa_t *P, int na, int nb) {
    sa_t *sa = P->sa, *sb = P->sb; const bso_params *p = P->par;
    if (na > sa->read_max_snp_num || nb > sb->read_max_snp_num) return 0;
    pairhit_t pp; memset(&pp, 0, sizeof pp); pp.na = (uint8_t)na; pp.nb = (uint8_t)nb;
    for (int dir = 0; dir < 2; dir++) {
        const hit_t *ha = dir ? sa->chits[na] : sa->hits[na]; const int cnt_a = dir ? sa->n_chit[na] : sa->n_hit[na];
        const hit_t *hb = dir ? sb->hits[nb] : sb->chits[nb]; const int cnt_b = dir ? sb->n_hit[nb] : sb->n_chit[nb];
        pp.chain = (uint16_t)dir;
        uint32_t chra = ~0u; int bstart = 0, bend = 0;
        for (int i = 0; i < cnt_a; i++) {
            if (chra != ha[i].chr) {
                chra = ha[i].chr;
                for (bstart = bend; bstart < cnt_b; bstart++) if (hb[bstart].chr >= chra) break;
                for (bend = bstart; bend < cnt_b; bend++) if (hb[bend].chr > chra) break;
            }
            for (int j = bstart; j < bend; j++) {
                uint32_t seg_start, seg_end;
                const int a_first = dir ? ((chra & 1) != 0) : ((chra & 1) == 0);
                if (!a_first) { seg_start = hb[j].loc; seg_end = ha[i].loc + (uint32_t)sa->len; }
                else { seg_start = ha[i].loc; seg_end = hb[j].loc + (uint32_t)sb->len; }
                int insert_size = (int)(seg_end - seg_start);
                if (insert_size >= p->min_insert && insert_size <= p->max_insert) {
                    pp.a = ha[i]; pp.b = hb[j]; pp.insert = insert_size;
                    P->pairhits[na + nb][P->n_pairs[na + nb]++] = pp;
                    if ((int)P->n_pairs[na + nb] >= p->max_num_hits) return 1;
                }
            }
        }
    }
    return P->n_pairs[na + nb] > 0;
}

static void pe_prepare(sa_t *a) {
    const bso_params *p = a->par;
    clear_hits(a); memset(a->probed, 0, sizeof a->probed);
    a->seedseg_num = (a->len - p->index_interval + 1) / p->seed_size;
    if (a->seedseg_num > a->read_max_snp_num + 1) a->seedseg_num = a->read_max_snp_num + 1;
}
/* PairAlign::RunAlign (pairs.cpp:137-190) */
static int pe_run_align(pa_t *P) {
    sa_t *sa = P->sa, *sb = P->sb; const bso_params *p = P->par;
    for (int i = 0; i <= p->max_snp_num * 2; i++) P->n_pairs[i] = 0;
    pe_prepare(sa); pe_prepare(sb);
    convert_binary(sa); convert_binary(sb);
    sa->snp_thres = (uint32_t)sa->read_max_snp_num; sb->snp_thres = (uint32_t)sb->read_max_snp_num;
    sa->cseed_offset = (uint32_t)(sa->len % p->seed_size); sb->cseed_offset = (uint32_t)(sb->len % p->seed_size);
    reorder_seed(sa); reorder_seed(sb);
    const int maxi = sa->read_max_snp_num > sb->read_max_snp_num ? sa->read_max_snp_num : sb->read_max_snp_num;
    for (int i = 0; i <= maxi; i++) {
        if (i < sa->seedseg_num) snp_align(sa, i);
        if (i < sb->seedseg_num) snp_align(sb, i);
        if (i <= sa->read_max_snp_num) sort_hits_pe(sa, i);
        if (i <= sb->read_max_snp_num) sort_hits_pe(sb, i);
        int n = get_pairs(P, i, i);
        for (int j = 0; j < i; j++) { n += get_pairs(P, i, j); n += get_pairs(P, j, i); }
        if (n > 0) return i + 1;
    }
    return 0;
}

/* Fix_Unpaired_Short_Fragment (align.cpp:768-791) */
static void fix_unpaired_short(sa_t *a) {
    const bso_params *p = a->par;
    if (a->len >= p->min_insert) return;
    for (int ii = 0; ii <= a->read_max_snp_num; ii++) {
        for (int pass = 0; pass < 2; pass++) {
            hit_t *h = pass ? a->chits[ii] : a->hits[ii]; int *cnt = pass ? &a->n_chit[ii] : &a->n_hit[ii];
            for (int j = 0; j < *cnt; j++) {
                uint32_t f; int sl; ccgg_seglen(a->ref, h[j].chr, h[j].loc, a->len, &f, &sl);
                if (sl < p->min_insert || sl > p->max_insert) {
                    (*cnt)--;
                    for (int k = j; k < *cnt; k++) h[k] = h[k + 1];
                    j--;
                }
            }
        }
        if (a->n_hit[ii] + a->n_chit[ii] > 0) break;
    }
}

/* the selection half of StringAlignUnpair (pairs.cpp:244-286) for one mate */
static void select_unpaired(sa_t *a, int filtered, bso_rec *o, uint16_t *counts) {
    memset(o, 0, sizeof *o); o->len = (uint8_t)a->len;
    if (counts) memset(counts, 0, 32);
    if (filtered) { o->status = 1; return; }
    int na, ma = 0, ra = 0;
    for (na = 0; na <= a->read_max_snp_num; na++) if ((ma = a->n_hit[na] + a->n_chit[na]) > 0) break;
    if (counts) for (int i = 0; i <= a->read_max_snp_num; i++) counts[i] = (uint16_t)(a->n_hit[i] + a->n_chit[i]);
    hit_t h = {0, 0};
    if (ma) {
        if (ma > 1) ra = (int)(bso_myrand((int32_t)a->index, a->par->randseed) % (uint32_t)ma);
        h = (ra < a->n_hit[na]) ? a->hits[na][ra] : a->chits[na][ra - a->n_hit[na]];
    }
    na %= (a->read_max_snp_num + 1);
    o->loc = h.loc; o->chr = h.chr; o->nhits = (uint32_t)ma; o->nm = (uint8_t)na;
    o->chain = (uint8_t)(ra >= a->n_hit[na]);
}

int bso_map_pe(const bso_ref *r, const bso_params *p, uint32_t n,
               const char *seqs_a, const char *seqs_b, uint32_t stride,
               const uint16_t *lens_a, const uint16_t *lens_b, uint32_t first_index,
               bso_pair_rec *out, bso_rec *out_a, bso_rec *out_b,
               uint16_t *counts_a, uint16_t *counts_b, uint64_t *stats) {
    init_tables();
    pa_t P; P.par = p; P.sa = sa_new(r, p); P.sb = sa_new(r, p);
    P.pairhits = malloc(sizeof(pairhit_t) * (2 * MAXSNPS + 1) * (MAXHITS + 1));
    for (uint32_t t = 0; t < n; t++) {
        load_read(P.sa, seqs_a + (size_t)t * stride, lens_a[t], 1, first_index + t);
        load_read(P.sb, seqs_b + (size_t)t * stride, lens_b[t], 2, first_index + t);
        int f1 = filter_read(P.sa), f2 = filter_read(P.sb), paired = 0;
        if (!f1 && !f2) paired = pe_run_align(&P);
        else { if (!f1) run_align(P.sa); if (!f2) run_align(P.sb); }
        bso_pair_rec *o = &out[t]; memset(o, 0, sizeof *o);
        int need_unpair = 1;
        if (paired) {
            /* StringAlignPair (pairs.cpp:222-242) */
            for (int i = 0; i <= p->max_snp_num * 2; i++) {
                if (!P.n_pairs[i]) continue;
                int j = -1;
                if (P.n_pairs[i] == 1) j = 0;
                else if (p->report_repeat_hits == 1) j = (int)(bso_myrand((int32_t)P.sa->index, p->randseed) % P.n_pairs[i]);
                if (j >= 0) {
                    pairhit_t *pp = &P.pairhits[i][j];
                    o->a_loc = pp->a.loc; o->a_chr = pp->a.chr; o->b_loc = pp->b.loc; o->b_chr = pp->b.chr;
                    o->insert = pp->insert; o->npairs = P.n_pairs[i]; o->na = pp->na; o->nb = pp->nb;
                    o->chain = (uint8_t)pp->chain; o->paired = 1; need_unpair = 0;
                }
                break;
            }
        }
        if (need_unpair && p->rrbs) { if (!f1) fix_unpaired_short(P.sa); if (!f2) fix_unpaired_short(P.sb); }
        /* records for the unpaired path are always produced (the formatter uses them only when
           paired == 0); when paired they still carry the trimmed length */
        select_unpaired(P.sa, f1, &out_a[t], counts_a ? counts_a + (size_t)t * 16 : NULL);
        select_unpaired(P.sb, f2, &out_b[t], counts_b ? counts_b + (size_t)t * 16 : NULL);
    }
    if (stats) { stats[0] += P.sa->n_cand + P.sb->n_cand; stats[1] += P.sa->n_probe_distinct + P.sb->n_probe_distinct;
                 stats[2] += P.sa->n_probe_ref + P.sb->n_probe_ref; }
    free(P.pairhits); sa_free(P.sa); sa_free(P.sb);
    return 0;
}

/* ---------------------------------------------------------------- text (align.cpp:631-765, pairs.cpp:288-498) */
typedef struct { char *p; size_t cap, n; } obuf;
static void oput(obuf *o, const char *s, size_t len) {
    if (o->n + len < o->cap) memcpy(o->p + o->n, s, len);
    o->n += len;
}
#define OPRINTF(o, ...) do { char _b[2048]; int _l = snprintf(_b, sizeof _b, __VA_ARGS__); oput((o), _b, (size_t)_l); } while (0)

static void revcomp(char *s, int n) {
    for (int i = 0, j = n - 1; i < j; i++, j--) { char t = s[i]; s[i] = s[j]; s[j] = t; }
    for (int i = 0; i < n; i++) s[i] = T_revchar[(uint8_t)s[i]];
}
static void reverse(char *s, int n) { for (int i = 0, j = n - 1; i < j; i++, j--) { char t = s[i]; s[i] = s[j]; s[j] = t; } }

/* the XR / BSP refseq payload (align.cpp:670-682): 2 upstream + len + 2 downstream Watson bases */
static void mapseq(const bso_ref *r, uint32_t chr, uint32_t loc, int len, char *out) {
    const uint32_t *m = r->refcat + r->anchor[chr >> 1] / SEGLEN;
    int ptr = 0;
    for (uint32_t ii = 2; ii > 0; ii--) {
        if (loc < ii) continue;   /* App. B Q12: reference leaves stale chars; we skip them */
        uint32_t q = loc - ii;
        out[ptr++] = (char)(USEFUL_NT[(m[q / SEGLEN] >> (30 - 2 * (q % SEGLEN))) & 3] + 32);
    }
    for (int ii = 0; ii < len + 2; ii++) {
        uint32_t q = loc + (uint32_t)ii;
        out[ptr++] = USEFUL_NT[(m[q / SEGLEN] >> (30 - 2 * (q % SEGLEN))) & 3];
    }
    out[ptr] = 0; out[ptr - 1] += 32; out[ptr - 2] += 32;
}

size_t bso_format_header(const bso_ref *r, char *out, size_t cap) {   /* main.cpp:405-413 */
    obuf o = { out, cap, 0 };
    OPRINTF(&o, "@HD\tVN:1.0\n");
    for (int i = 0; i < r->n_seq; i++) OPRINTF(&o, "@SQ\tSN:%s\tLN:%u\n", r->name[i], r->size[i]);
    OPRINTF(&o, "@PG\tID:BSMAP_2.6\n");
    if (o.n < cap) out[o.n] = 0;
    return o.n;
}

/* s_OutHit (align.cpp:631-765).  n: -1 QC, 0 NM, >0 hits. */
static void out_hit(const bso_ref *r, const bso_params *p, obuf *o, const char *name, char *seq, char *qual,
                    int len, int readset, int chain, int n, int nsnps, uint32_t chr, uint32_t loc,
                    int insert_size, const uint16_t *counts, int read_max_snp_num, uint32_t *n_aligned) {
    char ms[256];
    if (p->out_sam) {
        int flag = 0x40 * readset;
        if (n < 0 || n == 0 || (n > 1 && p->report_repeat_hits == 0)) {
            if (!p->out_unmap) return;
            flag |= (n < 0) ? 0x204 : (n == 0 ? 0x4 : 0x104);
            OPRINTF(o, "%s\t%d\t*\t0\t0\t*\t*\t0\t0\t%s\t%s\n", name, flag, seq, qual);
            return;
        }
        (*n_aligned)++;
        if (n > 1) flag |= 0x100;
        if (chain ^ (int)(chr % 2)) { flag |= 0x10; revcomp(seq, len); reverse(qual, len); }
        OPRINTF(o, "%s\t%d\t%s\t%u\t255\t%dM\t*\t0\t0\t%s\t%s\tNM:i:%d", name, flag, r->name[chr >> 1], loc + 1, len, seq, qual, nsnps);
        if (p->out_ref) { mapseq(r, chr, loc, len, ms); OPRINTF(o, "\tXR:Z:%s", ms); }
        if (p->rrbs) { uint32_t f; int sl; ccgg_seglen(r, chr, loc, len, &f, &sl); OPRINTF(o, "\tZP:i:%d\tZL:i:%d", (int)f, sl); }
        OPRINTF(o, "\tZS:Z:%c%c\n", "+-"[chr % 2], "+-"[chain]);
    } else {
        if (!p->out_unmap && (n <= 0 || (n > 1 && p->report_repeat_hits == 0))) return;
        OPRINTF(o, "%s\t", name);
        if ((chain ^ (int)(chr % 2)) && n) { revcomp(seq, len); reverse(qual, len); }
        OPRINTF(o, "%s\t%s\t", seq, qual);
        if (n < 0) oput(o, "QC", 2); else if (n == 0) oput(o, "NM", 2); else if (n == 1) oput(o, "UM", 2);
        else if (n >= p->max_num_hits) oput(o, "OF", 2); else oput(o, "MA", 2);
        if ((n > 0 && p->report_repeat_hits == 1) || (n == 1 && p->report_repeat_hits == 0)) {
            (*n_aligned)++;
            mapseq(r, chr, loc, len, ms);
            OPRINTF(o, "\t%s\t%u\t%c%c\t%d\t%s\t%d\t", r->name[chr >> 1], loc + 1, "+-"[chr % 2], "+-"[chain], insert_size, ms, nsnps);
            int ii;
            for (ii = 0; ii < read_max_snp_num; ii++) OPRINTF(o, "%d:", counts ? counts[ii] : 0);
            OPRINTF(o, "%d", counts ? counts[ii] : 0);
        }
        oput(o, "\n", 1);
    }
}

static int rmsn(const bso_params *p, int len, int raw) { return (int)((size_t)(p->max_snp_num + 1) * (size_t)(len - 1) / (size_t)raw); }
static int clip_len(const bso_params *p, const char *s) { int l = (int)strlen(s); if (l > p->max_readlen) l = p->max_readlen; if (l > FIXSIZE - 16) l = FIXSIZE - 16; return l; }

size_t bso_format_se(const bso_ref *r, const bso_params *p, uint32_t n, const char *const *names,
                     const char *const *seqs, const char *const *quals, int readset,
                     const bso_rec *recs, const uint16_t *counts, char *out, size_t cap, uint32_t *n_aligned) {
    init_tables();
    obuf o = { out, cap, 0 }; uint32_t na = 0;
    char sq[FIXSIZE + 16], ql[FIXSIZE + 16];
    for (uint32_t t = 0; t < n; t++) {
        const bso_rec *rc = &recs[t];
        int raw = clip_len(p, seqs[t]), len = rc->len;
        memcpy(sq, seqs[t], len); sq[len] = 0;
        int qlen = (int)strlen(quals[t]); if (qlen > len) qlen = len;   /* erase(pos) on seq and qual */
        memcpy(ql, quals[t], qlen); ql[qlen] = 0;
        if (rc->status == 1) {                                         /* align.cpp:598-600 */
            if (p->report_repeat_hits) out_hit(r, p, &o, names[t], sq, ql, len, readset, 0, -1, 0, 0, 0, 0, NULL, 0, &na);
            continue;
        }
        out_hit(r, p, &o, names[t], sq, ql, len, readset, rc->chain, (int)rc->nhits, rc->nm, rc->chr, rc->loc, 0,
                counts ? counts + (size_t)t * 16 : NULL, rmsn(p, len, raw), &na);
    }
    if (o.n < cap) out[o.n] = 0;
    if (n_aligned) *n_aligned = na;
    return o.n;
}

/* FixPairReadName (pairs.cpp:535-555); returns common length or -1 for "names do not match" */
static int fix_pair_name(const bso_params *p, const char *a, const char *b, int *la, int *lb) {
    *la = (int)strlen(a); *lb = (int)strlen(b);
    if (!p->out_sam) return 0;
    if (strcmp(a, b) == 0) return 0;
    int i, d = -1, i0 = *la < *lb ? *la : *lb;
    for (i = 0; i < i0; i++) { if (a[i] != b[i]) break; else if (isdigit((uint8_t)a[i])) d = i; }
    if (i > 0) { if (d < 0) d = i - 1; if (*la > d + 1) *la = d + 1; if (*lb > d + 1) *lb = d + 1; return 0; }
    return -1;
}

/* s_OutHitUnpair (pairs.cpp:426-498), SAM branch */
static void out_unpair_sam(const bso_ref *r, const bso_params *p, obuf *o, const char *name, char *seq, char *qual, int len,
                           int readset, int chain_a, int chain_b, int ma, int na, uint32_t a_chr, uint32_t a_loc,
                           int mb, uint32_t b_chr, uint32_t b_loc, uint32_t *n_al) {
    char ms[256];
    int flag = 1 | (0x40 * readset);
    const int mate_un = (mb <= 0 || (mb > 1 && p->report_repeat_hits == 0));
    if (ma <= 0 || (ma > 1 && p->report_repeat_hits == 0)) {
        if (!p->out_unmap) return;
        if (ma < 0) flag |= 0x204; if (ma == 0) flag |= 0x004; if (ma > 1) flag |= 0x104;
        if (mate_un) { flag |= 0x008; OPRINTF(o, "%s\t%d\t*\t0\t0\t*\t*\t0\t0\t%s\t%s\n", name, flag, seq, qual); }
        else { if (chain_b ^ (int)(b_chr % 2)) flag |= 0x020;
               OPRINTF(o, "%s\t%d\t*\t0\t0\t*\t%s\t%u\t0\t%s\t%s\n", name, flag, r->name[b_chr >> 1], b_loc + 1, seq, qual); }
        return;
    }
    (*n_al)++;
    if (ma > 1) flag |= 0x100;
    if (chain_a ^ (int)(a_chr % 2)) { flag |= 0x010; revcomp(seq, len); reverse(qual, (int)strlen(qual)); }
    if (mate_un) { flag |= 0x008;
        OPRINTF(o, "%s\t%d\t%s\t%u\t255\t%dM\t*\t0\t0\t%s\t%s\tNM:i:%d", name, flag, r->name[a_chr >> 1], a_loc + 1, len, seq, qual, na); }
    else { if (chain_b ^ (int)(b_chr % 2)) flag |= 0x020;
        OPRINTF(o, "%s\t%d\t%s\t%u\t255\t%dM\t%s\t%u\t0\t%s\t%s\tNM:i:%d", name, flag, r->name[a_chr >> 1], a_loc + 1, len, r->name[b_chr >> 1], b_loc + 1, seq, qual, na); }
    if (p->out_ref) { mapseq(r, a_chr, a_loc, len, ms); OPRINTF(o, "\tXR:Z:%s", ms); }
    if (p->rrbs) { uint32_t f; int sl; ccgg_seglen(r, a_chr, a_loc, len, &f, &sl); OPRINTF(o, "\tZP:i:%d\tZL:i:%d", (int)f, sl); }
    OPRINTF(o, "\tZS:Z:%c%c\n", "+-"[a_chr % 2], "+-"[chain_a]);
}

size_t bso_format_pe(const bso_ref *r, const bso_params *p, uint32_t n,
                     const char *const *names_a, const char *const *seqs_a, const char *const *quals_a,
                     const char *const *names_b, const char *const *seqs_b, const char *const *quals_b,
                     const bso_pair_rec *pr, const bso_rec *ra, const bso_rec *rb,
                     const uint16_t *counts_a, const uint16_t *counts_b,
                     char *out, size_t cap, char *out_unpair, size_t cap_unpair, size_t *n_unpair, uint32_t *n_stats) {
    init_tables();
    obuf o = { out, cap, 0 }, ou = { out_unpair, cap_unpair, 0 };
    uint32_t n_pairs = 0, n_a = 0, n_b = 0;
    char sa[FIXSIZE + 16], qa[FIXSIZE + 16], sb[FIXSIZE + 16], qb[FIXSIZE + 16], nma[1024], nmb[1024], ms[256];
    for (uint32_t t = 0; t < n; t++) {
        int rawa = clip_len(p, seqs_a[t]), rawb = clip_len(p, seqs_b[t]);
        int lena = ra[t].len, lenb = rb[t].len;
        memcpy(sa, seqs_a[t], lena); sa[lena] = 0; memcpy(sb, seqs_b[t], lenb); sb[lenb] = 0;
        int q = (int)strlen(quals_a[t]); if (q > lena) q = lena; memcpy(qa, quals_a[t], q); qa[q] = 0;
        q = (int)strlen(quals_b[t]); if (q > lenb) q = lenb; memcpy(qb, quals_b[t], q); qb[q] = 0;
        int la, lb; fix_pair_name(p, names_a[t], names_b[t], &la, &lb);
        memcpy(nma, names_a[t], la); nma[la] = 0; memcpy(nmb, names_b[t], lb); nmb[lb] = 0;
        const bso_pair_rec *pp = &pr[t];
        if (pp->paired) {
            /* s_OutHitPair (pairs.cpp:288-424) */
            n_pairs++;
            uint32_t a_loc = pp->a_loc, b_loc = pp->b_loc; const int ins = pp->insert, chain = pp->chain;
            if (ins < lena) { if (chain ^ (int)(pp->a_chr % 2)) a_loc += (uint32_t)(lena - ins); lena = ins; sa[lena] = 0; if ((int)strlen(qa) > ins) qa[ins] = 0; }
            if (ins < lenb) { if ((!chain) ^ (int)(pp->b_chr % 2)) b_loc += (uint32_t)(lenb - ins); lenb = ins; sb[lenb] = 0; if ((int)strlen(qb) > ins) qb[ins] = 0; }
            const int n_p = (int)pp->npairs;
            if (p->out_sam) {
                for (int mate = 0; mate < 2; mate++) {
                    char *sq = mate ? sb : sa, *ql = mate ? qb : qa; const char *nm = mate ? nmb : nma;
                    const int len = mate ? lenb : lena, readset = mate ? 2 : 1;
                    const uint32_t chr = mate ? pp->b_chr : pp->a_chr, loc = mate ? b_loc : a_loc, mloc = mate ? a_loc : b_loc;
                    const int ch = mate ? !chain : chain;
                    int flag = 0x3, pp_insert; uint32_t seg_start;
                    if (n_p > 1) flag |= 0x100;
                    if (ch ^ (int)(chr % 2)) { flag |= 0x10; seg_start = mloc + 1; pp_insert = -ins; revcomp(sq, len); reverse(ql, (int)strlen(ql)); }
                    else { flag |= 0x20; seg_start = loc + 1; pp_insert = ins; }
                    flag |= 0x40 * readset;
                    OPRINTF(&o, "%s\t%d\t%s\t%u\t255\t%dM\t=\t%u\t%d\t%s\t%s\tNM:i:%d", nm, flag, r->name[chr >> 1], loc + 1, len, mloc + 1, pp_insert, sq, ql, mate ? pp->nb : pp->na);
                    if (p->out_ref) { mapseq(r, chr, loc, len, ms); OPRINTF(&o, "\tXR:Z:%s", ms); }
                    if (p->rrbs) OPRINTF(&o, "\tZP:i:%d\tZL:i:%d", (int)seg_start, ins);
                    OPRINTF(&o, "\tZS:Z:%c%c\n", "+-"[chr % 2], "+-"[ch]);
                }
            } else {
                uint32_t dummy = 0;
                out_hit(r, p, &o, nma, sa, qa, lena, 1, chain, n_p, pp->na, pp->a_chr, a_loc, ins, counts_a ? counts_a + (size_t)t * 16 : NULL, rmsn(p, ra[t].len, rawa), &dummy);
                out_hit(r, p, &o, nmb, sb, qb, lenb, 2, !chain, n_p, pp->nb, pp->b_chr, b_loc, ins, counts_b ? counts_b + (size_t)t * 16 : NULL, rmsn(p, rb[t].len, rawb), &dummy);
            }
            continue;
        }
        /* StringAlignUnpair + s_OutHitUnpair */
        obuf *dst = p->out_sam ? &o : &ou;
        const int ma = ra[t].status ? -1 : (int)ra[t].nhits, mb = rb[t].status ? -1 : (int)rb[t].nhits;
        if (p->out_sam) {
            out_unpair_sam(r, p, dst, nma, sa, qa, lena, 1, ra[t].chain, rb[t].chain, ma, ra[t].nm, ra[t].chr, ra[t].loc, mb, rb[t].chr, rb[t].loc, &n_a);
            out_unpair_sam(r, p, dst, nmb, sb, qb, lenb, 2, rb[t].chain, ra[t].chain, mb, rb[t].nm, rb[t].chr, rb[t].loc, ma, ra[t].chr, ra[t].loc, &n_b);
        } else {
            uint32_t dummy = 0;   /* BSP: SingleAlign::n_aligned is bumped, not PairAlign's counters */
            out_hit(r, p, dst, nma, sa, qa, lena, 1, ra[t].chain, ma, ra[t].nm, ra[t].chr, ra[t].loc, 0, counts_a ? counts_a + (size_t)t * 16 : NULL, ra[t].status ? 0 : rmsn(p, lena, rawa), &dummy);
            out_hit(r, p, dst, nmb, sb, qb, lenb, 2, rb[t].chain, mb, rb[t].nm, rb[t].chr, rb[t].loc, 0, counts_b ? counts_b + (size_t)t * 16 : NULL, rb[t].status ? 0 : rmsn(p, lenb, rawb), &dummy);
        }
    }
    if (o.n < cap) out[o.n] = 0;
    if (out_unpair && ou.n < cap_unpair) out_unpair[ou.n] = 0;
    if (n_unpair) *n_unpair = ou.n;
    if (n_stats) { n_stats[0] = n_pairs; n_stats[1] = n_a; n_stats[2] = n_b; }
    return o.n;
}

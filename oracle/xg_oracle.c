/* xg_oracle.c -- CPU restatement of the reference's two counting paths.
 *
 * ORACLE / TEST INFRASTRUCTURE ONLY: used by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py as the checker and the CPU baseline.
 * The product path (xcltk_b200/) never links, imports or calls anything in this directory.
 *
 * It follows the reference's own control flow (feature loop -> fetch -> per-read Python
 * logic), not the GPU formulation:
 *   basefc: fc_features / fc_fet1   xcltk/rdr/fc/core.py:69-178
 *           check_read              xcltk/rdr/fc/core.py:46-62
 *           __get_include_frac/_len xcltk/rdr/fc/core.py:32-43 (per-position list, as written)
 *           MCount/SCount           xcltk/rdr/fc/mcount.py:34-54,102-149
 *   baf:    fc_features / fc_fet1   xcltk/baf/fc/core.py:42-194
 *           plp_snp                 xcltk/baf/fc/core.py:198-247
 *           UCount/SCount/MCount    xcltk/baf/fc/mcount.py:39-60,109-150,206-256
 *           get_query_bases         xcltk/utils/sam.py:4-40
 *           SNP.get_region_allele_index  xcltk/baf/fc/gfeature.py:38-39
 *   pysam/htslib semantics (fetch overlap, positions, 4-bit bases): SURVEY.md A.3.
 * Input is the decoded record arrays of include/xcltk_b200.h (the decoder is checked against
 * the pure-Python BAM reader of oracle/shim separately); strings are compared through their
 * lossless 64-bit keys.
 *
 * Pinned against the unmodified reference: tests/test_oracle.py runs it on every case of
 * tests/golden (outputs of /root/reference produced by oracle/make_golden.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/xcltk_b200.h"

typedef struct {
    int64_t nnz, cap;
    int32_t *row, *col, *val;
} orc_coo;

typedef struct {
    double min_mapq;          /* conf.min_mapq (int or float in Python)        */
    int32_t min_len;
    double min_include;       /* conf.min_include                              */
    int32_t min_include_is_int; /* Python type of min_include (int vs float)   */
    uint32_t incl_flag, excl_flag;
    int32_t no_orphan, use_cell_tag, need_umi_tag;
} orc_params;

static int coo_push(orc_coo *m, int32_t r, int32_t c, int32_t v) {
    if (m->nnz == m->cap) {
        int64_t nc = m->cap ? m->cap * 2 : 1024;
        int32_t *a = (int32_t *)realloc(m->row, (size_t)nc * 4);
        int32_t *b = (int32_t *)realloc(m->col, (size_t)nc * 4);
        int32_t *d = (int32_t *)realloc(m->val, (size_t)nc * 4);
        if (a) m->row = a;
        if (b) m->col = b;
        if (d) m->val = d;
        if (!a || !b || !d) return -1;
        m->cap = nc;
    }
    m->row[m->nnz] = r;
    m->col[m->nnz] = c;
    m->val[m->nnz] = v;
    m->nnz++;
    return 0;
}

void orc_coo_free(orc_coo *m) {
    free(m->row);
    free(m->col);
    free(m->val);
    memset(m, 0, sizeof(*m));
}

/* ---- barcode dict: `smp in self.cell_cnt` (rdr/fc/mcount.py:123) ----------------------- */
typedef struct {
    uint64_t *keys;
    int32_t *cols;
    uint32_t mask;
} bc_dict;

static uint64_t hash64(uint64_t x) {
    x ^= x >> 31;
    x *= 0x7fb5d329728ea185ULL;
    x ^= x >> 27;
    x *= 0x81dadef4bc2dd44dULL;
    x ^= x >> 33;
    return x;
}

static int bc_build(bc_dict *d, int32_t n, const uint64_t *keys) {
    uint32_t cap = 16;
    while (cap < (uint32_t)n * 2u + 2u) cap <<= 1;
    d->keys = (uint64_t *)malloc((size_t)cap * 8);
    d->cols = (int32_t *)malloc((size_t)cap * 4);
    if (!d->keys || !d->cols) return -1;
    d->mask = cap - 1;
    for (uint32_t i = 0; i < cap; i++) d->keys[i] = XG_KEY_NONE;
    for (int32_t i = 0; i < n; i++) {
        uint32_t s = (uint32_t)hash64(keys[i]) & d->mask;
        while (d->keys[s] != XG_KEY_NONE) s = (s + 1) & d->mask;
        d->keys[s] = keys[i];
        d->cols[s] = i;
    }
    return 0;
}

static int32_t bc_get(const bc_dict *d, uint64_t key) {
    uint32_t s = (uint32_t)hash64(key) & d->mask;
    while (d->keys[s] != XG_KEY_NONE) {
        if (d->keys[s] == key) return d->cols[s];
        s = (s + 1) & d->mask;
    }
    return -1;
}

/* ---- a (cell, key) -> value map, cleared per feature / SNP ------------------------------ */
typedef struct {
    uint64_t *key;      /* UMI / query-name key */
    int32_t *cell;      /* -1 = empty           */
    int32_t *val;
    uint32_t cap, n;
} ckmap;

static int ck_init(ckmap *m, uint32_t cap) {
    uint32_t c = 64;
    while (c < cap) c <<= 1;
    m->key = (uint64_t *)malloc((size_t)c * 8);
    m->cell = (int32_t *)malloc((size_t)c * 4);
    m->val = (int32_t *)malloc((size_t)c * 4);
    if (!m->key || !m->cell || !m->val) return -1;
    m->cap = c;
    m->n = 0;
    for (uint32_t i = 0; i < c; i++) m->cell[i] = -1;
    return 0;
}
static void ck_free(ckmap *m) {
    free(m->key);
    free(m->cell);
    free(m->val);
}
static int ck_grow(ckmap *m);
/* returns slot; *isnew = 1 when inserted */
static uint32_t ck_find_or_insert(ckmap *m, int32_t cell, uint64_t key, int *isnew) {
    if ((uint64_t)m->n * 2 >= m->cap) ck_grow(m);
    uint32_t s = (uint32_t)hash64(key ^ ((uint64_t)(uint32_t)cell << 40)) & (m->cap - 1);
    while (m->cell[s] >= 0) {
        if (m->cell[s] == cell && m->key[s] == key) {
            *isnew = 0;
            return s;
        }
        s = (s + 1) & (m->cap - 1);
    }
    m->cell[s] = cell;
    m->key[s] = key;
    m->val[s] = 0;
    m->n++;
    *isnew = 1;
    return s;
}
static int ck_grow(ckmap *m) {
    ckmap o = *m;
    if (ck_init(m, o.cap * 2)) return -1;
    for (uint32_t i = 0; i < o.cap; i++)
        if (o.cell[i] >= 0) {
            int nw;
            uint32_t s = ck_find_or_insert(m, o.cell[i], o.key[i], &nw);
            m->val[s] = o.val[i];
        }
    ck_free(&o);
    return 0;
}
static void ck_clear(ckmap *m) {
    if (m->n == 0) return;
    for (uint32_t i = 0; i < m->cap; i++) m->cell[i] = -1;
    m->n = 0;
}

/* ---- the slice of pysam.AlignedSegment the paths use -------------------------------------- */
static uint32_t rd_flag(const xg_reads *r, int64_t i) { return r->fmq[i] & 0xffffu; }
static uint32_t rd_mapq(const xg_reads *r, int64_t i) { return (r->fmq[i] >> 16) & 0xffu; }

/* read.cigartuples: returns the op words (one synthesized M op for "simple" records) */
static uint32_t rd_cigar(const xg_reads *r, int64_t i, const uint32_t **words, uint32_t *one) {
    uint32_t ncw = r->fmq[i] >> 24;
    if (ncw == 0) {
        *one = ((uint32_t)(r->pos_end[2 * i + 1] - r->pos_end[2 * i]) << 4) | 0u;
        *words = one;
        return 1;
    }
    *words = r->cigar + r->cig_off[i];
    if (ncw == 255) return (*words)[-1];
    return ncw;
}

/* read.positions (pysam get_reference_positions): reference positions of M/=/X bases.
 * Fills `buf` (realloc'd) and returns the count. */
static int32_t rd_positions(const xg_reads *r, int64_t i, int32_t **buf, int32_t *cap) {
    const uint32_t *w;
    uint32_t one, n = rd_cigar(r, i, &w, &one);
    int32_t p = r->pos_end[2 * i], k = 0;
    for (uint32_t q = 0; q < n; q++) {
        uint32_t op = w[q] & 15u, l = w[q] >> 4;
        if (op == 0 || op == 7 || op == 8) {
            if (k + (int32_t)l > *cap) {
                *cap = (k + (int32_t)l) * 2 + 64;
                *buf = (int32_t *)realloc(*buf, (size_t)*cap * 4);
            }
            for (uint32_t t = 0; t < l; t++) (*buf)[k++] = p++;
        } else if (op == 2 || op == 3) {
            p += (int32_t)l;
        }
    }
    return k;
}

/* check_read() without the length test (rdr/fc/core.py:46-59) */
static int check_read_head(const xg_reads *r, int64_t i, const orc_params *c) {
    uint32_t flag = rd_flag(r, i);
    if ((double)rd_mapq(r, i) < c->min_mapq) return -2;
    if (c->excl_flag && (flag & c->excl_flag)) return -3;
    if (c->incl_flag && !(flag & c->incl_flag)) return -4;
    if (c->no_orphan && (flag & 1u) && !(flag & 2u)) return -5;
    if (c->use_cell_tag && r->keys[2 * i] == XG_KEY_NONE) return -11;
    if (c->need_umi_tag && r->keys[2 * i + 1] == XG_KEY_NONE) return -12;
    return 0;
}

/* sam.fetch(contig, beg0, end0) over one run: records with pos < end0 and bam_endpos > beg0 in
 * file order.  The scan starts max_span before beg0 (stand-in for the BAI linear index).     */
static int64_t fetch_begin(const xg_reads *r, const xg_run *run, int32_t beg0) {
    int64_t lo = run->rec_beg, hi = run->rec_end;
    int64_t want = (int64_t)beg0 - (int64_t)r->max_span;
    while (lo < hi) {
        int64_t mid = (lo + hi) / 2;
        if ((int64_t)r->pos_end[2 * mid] < want) lo = mid + 1; else hi = mid;
    }
    return lo;
}

/* ============================== basefc ==================================================== */
int orc_basefc(const xg_reads *r, int32_t n_feat, const int32_t *gid, const int32_t *beg,
               const int32_t *end, int32_t n_bc, const uint64_t *bc_keys, int32_t n_samples,
               const orc_params *c, int32_t n_threads, orc_coo *out) {
    memset(out, 0, sizeof(*out));
    bc_dict bc;
    memset(&bc, 0, sizeof(bc));
    if (c->use_cell_tag && bc_build(&bc, n_bc, bc_keys)) return -1;
    orc_coo *parts = (orc_coo *)calloc((size_t)n_feat, sizeof(orc_coo));
    if (!parts && n_feat) return -1;
    int err = 0;
    (void)n_threads;
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads > 0 ? n_threads : 1)
#endif
    {
        ckmap set;                      /* SCount.umi_set of every cell of one feature */
        int32_t *cnt = (int32_t *)calloc((size_t)n_samples, 4);
        int32_t *posbuf = NULL, poscap = 0;
        int bad = ck_init(&set, 1024) || !cnt;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
        for (int32_t f = 0; f < n_feat; f++) {           /* for reg in reg_list (core.py:96) */
            if (bad) continue;
            if (gid[f] < 0 || beg[f] < 0 || end[f] <= beg[f]) continue;   /* fetch fails: no reads */
            ck_clear(&set);
            memset(cnt, 0, (size_t)n_samples * 4);
            for (int32_t ri = 0; ri < r->n_runs; ri++) {  /* for sam in sam_list (core.py:153) */
                const xg_run *run = &r->runs[ri];
                if (run->gid != gid[f]) continue;
                for (int64_t i = fetch_begin(r, run, beg[f]); i < run->rec_end; i++) {
                    int32_t pos = r->pos_end[2 * i], rend = r->pos_end[2 * i + 1];
                    if (pos >= end[f]) break;
                    if (rend <= beg[f]) continue;
                    if (check_read_head(r, i, c) < 0) continue;
                    int32_t n = rd_positions(r, i, &posbuf, &poscap);
                    if (n < c->min_len) continue;                     /* len(read.positions) */
                    /* include test: positions x with s <= x <= e, s = start-1, e = end-2 */
                    int32_t s = beg[f], e = end[f] - 1, m = 0;
                    for (int32_t k = 0; k < n; k++)
                        if (s <= posbuf[k] && posbuf[k] <= e) m++;
                    if (0 < c->min_include && c->min_include < 1) {
                        if (n <= 0) continue;       /* frac None: TypeError in py3, unreachable */
                        if ((double)m / (double)n < c->min_include) continue;
                    } else {
                        if ((double)m < c->min_include) continue;
                    }
                    int32_t col;
                    if (c->use_cell_tag) {
                        col = bc_get(&bc, r->keys[2 * i]);
                        if (col < 0) continue;                        /* push_read -> -2 */
                    } else {
                        col = run->bam_idx;
                    }
                    uint64_t umi = r->keys[2 * i + 1];
                    if (umi == XG_KEY_EMPTY || umi == XG_KEY_NONE) continue;   /* `if umi:` */
                    int isnew;
                    ck_find_or_insert(&set, col, umi, &isnew);
                    if (isnew) cnt[col]++;
                }
            }
            for (int32_t col = 0; col < n_samples; col++)      /* emit loop, core.py:109-117 */
                if (cnt[col] > 0 && coo_push(&parts[f], f, col, cnt[col])) bad = 1;
        }
        if (bad) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
            err = 1;
        }
        ck_free(&set);
        free(cnt);
        free(posbuf);
    }
    for (int32_t f = 0; f < n_feat && !err; f++) {
        for (int64_t k = 0; k < parts[f].nnz; k++)
            if (coo_push(out, parts[f].row[k], parts[f].col[k], parts[f].val[k])) err = 1;
    }
    for (int32_t f = 0; f < n_feat; f++) orc_coo_free(&parts[f]);
    free(parts);
    free(bc.keys);
    free(bc.cols);
    return err ? -1 : 0;
}

/* ============================== baf ======================================================= */
typedef struct {
    int32_t n;
    const int32_t *gid, *pos;        /* 0-based pos                                     */
    const char *ref, *alt;           /* base letters (A C G T N)                        */
    const int8_t *ref_idx, *alt_idx; /* haplotype index of the REF / ALT allele (0 / 1) */
} orc_snps;

static const char SEQ_NT16[] = "=ACMGRSVTWYHKDBN";

/* UCount.push_read (baf/fc/mcount.py:39-60): allele letter at the SNP or 0 for None */
static char read_allele(const xg_reads *r, int64_t i, int32_t snp_pos0, int32_t **posbuf, int32_t *poscap) {
    int32_t n = rd_positions(r, i, posbuf, poscap), idx = -1;
    for (int32_t k = 0; k < n; k++)
        if ((*posbuf)[k] == snp_pos0) {          /* positions.index(snp.pos - 1) */
            idx = k;
            break;
        }
    if (idx < 0) return 0;
    if (r->seq_off[i] == 0xFFFFFFFFu) return 0;  /* no sequence stored (reference would raise) */
    /* get_query_bases(read)[idx]: idx-th base among the M/=/X-aligned query bases */
    const uint32_t *w;
    uint32_t one, nops = rd_cigar(r, i, &w, &one);
    int32_t q = 0, seen = 0;
    const uint8_t *seq = (const uint8_t *)(r->seq + r->seq_off[i]);
    for (uint32_t t = 0; t < nops; t++) {
        uint32_t op = w[t] & 15u, l = w[t] >> 4;
        if (op == 4 || op == 1) {
            q += (int32_t)l;
        } else if (op == 0 || op == 7 || op == 8) {
            if (idx < seen + (int32_t)l) {
                int32_t qi = q + (idx - seen);
                uint8_t b = seq[qi >> 1];
                return SEQ_NT16[(qi & 1) ? (b & 15) : (b >> 4)];
            }
            seen += (int32_t)l;
            q += (int32_t)l;
        }
    }
    return 0;
}

static int base_idx(char b) {    /* MCount.base_idx; anything else -> N (mcount.py:145-149) */
    switch (b) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
        default: return 4;
    }
}

int orc_baf(const xg_reads *r, const orc_snps *snps, int32_t n_regions, const int64_t *reg_ptr,
            const int32_t *reg_snp, int32_t n_bc, const uint64_t *bc_keys, int32_t n_samples,
            const orc_params *c, double min_count, double min_maf, int32_t no_dup_hap,
            int32_t n_threads, orc_coo *ad, orc_coo *dp, orc_coo *oth) {
    memset(ad, 0, sizeof(*ad));
    memset(dp, 0, sizeof(*dp));
    memset(oth, 0, sizeof(*oth));
    bc_dict bc;
    memset(&bc, 0, sizeof(bc));
    if (c->use_cell_tag && bc_build(&bc, n_bc, bc_keys)) return -1;
    orc_coo *parts = (orc_coo *)calloc((size_t)n_regions * 3, sizeof(orc_coo));
    if (!parts && n_regions) return -1;
    int err = 0;
    (void)n_threads;
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads > 0 ? n_threads : 1)
#endif
    {
        ckmap umi_cnt;   /* per SNP: (cell, UMI) -> allele letter of the first read (0 = None) */
        ckmap reg;       /* per region: (cell, UMI) -> bit0 ref-hap, bit1 alt-hap, bit2 other   */
        int32_t *posbuf = NULL, poscap = 0;
        int32_t *cnt = (int32_t *)calloc((size_t)n_samples * 4, 4);
        int bad = ck_init(&umi_cnt, 1024) || ck_init(&reg, 1024) || !cnt;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
        for (int32_t g = 0; g < n_regions; g++) {        /* for reg in reg_list (core.py:70) */
            if (bad) continue;
            ck_clear(&reg);
            for (int64_t k = reg_ptr[g]; k < reg_ptr[g + 1]; k++) {   /* for snp in reg.snp_list */
                int32_t si = reg_snp[k];
                int32_t spos = snps->pos[si];
                /* ---- plp_snp (core.py:198-247) */
                ck_clear(&umi_cnt);
                if (snps->gid[si] >= 0 && spos >= 0) {
                    for (int32_t ri = 0; ri < r->n_runs; ri++) {
                        const xg_run *run = &r->runs[ri];
                        if (run->gid != snps->gid[si]) continue;
                        for (int64_t i = fetch_begin(r, run, spos); i < run->rec_end; i++) {
                            int32_t pos = r->pos_end[2 * i], rend = r->pos_end[2 * i + 1];
                            if (pos >= spos + 1) break;
                            if (rend <= spos) continue;
                            if (check_read_head(r, i, c) < 0) continue;
                            if (rd_positions(r, i, &posbuf, &poscap) < c->min_len) continue;
                            int32_t col;
                            if (c->use_cell_tag) {
                                col = bc_get(&bc, r->keys[2 * i]);
                                if (col < 0) continue;
                            } else {
                                col = run->bam_idx;
                            }
                            uint64_t umi = r->keys[2 * i + 1];
                            if (umi == XG_KEY_EMPTY || umi == XG_KEY_NONE) continue;
                            int isnew;
                            uint32_t s = ck_find_or_insert(&umi_cnt, col, umi, &isnew);
                            if (isnew)                       /* first read wins, even if None */
                                umi_cnt.val[s] = (int32_t)read_allele(r, i, spos, &posbuf, &poscap);
                        }
                    }
                }
                /* mcnt.stat(): totals over all cells (mcount.py:140-150, 250-256) */
                int64_t tc[5] = {0, 0, 0, 0, 0};
                for (uint32_t s = 0; s < umi_cnt.cap; s++)
                    if (umi_cnt.cell[s] >= 0 && umi_cnt.val[s]) tc[base_idx((char)umi_cnt.val[s])]++;
                int64_t snp_cnt = tc[0] + tc[1] + tc[2] + tc[3] + tc[4];
                if ((double)snp_cnt < min_count) continue;
                int64_t rc = tc[base_idx(snps->ref[si])], ac = tc[base_idx(snps->alt[si])];
                int64_t minor = rc < ac ? rc : ac;
                if ((double)minor < (double)snp_cnt * min_maf) continue;
                /* region UMI sets (core.py:156-166) */
                for (uint32_t s = 0; s < umi_cnt.cap; s++) {
                    if (umi_cnt.cell[s] < 0 || !umi_cnt.val[s]) continue;
                    char a = (char)umi_cnt.val[s];
                    /* snp.gt = {ref: ref_idx, alt: alt_idx}: the ALT entry wins if ref == alt */
                    int ale = -1;
                    if (a == snps->ref[si]) ale = snps->ref_idx[si];
                    if (a == snps->alt[si]) ale = snps->alt_idx[si];
                    int isnew;
                    uint32_t t = ck_find_or_insert(&reg, umi_cnt.cell[s], umi_cnt.key[s], &isnew);
                    reg.val[t] |= (ale == 0) ? 1 : (ale == 1) ? 2 : 4;
                }
            }
            /* set algebra per cell (core.py:173-192) */
            memset(cnt, 0, (size_t)n_samples * 16);
            for (uint32_t s = 0; s < reg.cap; s++) {
                if (reg.cell[s] < 0) continue;
                int32_t *q = cnt + (size_t)reg.cell[s] * 4, m = reg.val[s];
                if (m & 1) q[0]++;                          /* ref UMIs          */
                if (m & 2) q[1]++;                          /* alt UMIs          */
                if (m & 3) q[2]++;                          /* |ref union alt|   */
                if ((m & 4) && !(m & 3)) q[3]++;            /* oth minus dp      */
            }
            for (int32_t col = 0; col < n_samples; col++) {     /* emit, core.py:84-101 */
                int32_t ref = cnt[col * 4], alt = cnt[col * 4 + 1], dpu = cnt[col * 4 + 2];
                int32_t o = cnt[col * 4 + 3], dpv = dpu;
                if (ref + alt != dpu) {
                    if (no_dup_hap) {
                        int32_t share = ref + alt - dpu;
                        ref -= share;
                        alt -= share;
                    }
                    dpv = ref + alt;
                }
                if (dpv + o <= 0) continue;
                if (alt > 0 && coo_push(&parts[g * 3], g, col, alt)) bad = 1;
                if (dpv > 0 && coo_push(&parts[g * 3 + 1], g, col, dpv)) bad = 1;
                if (o > 0 && coo_push(&parts[g * 3 + 2], g, col, o)) bad = 1;
            }
        }
        if (bad) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
            err = 1;
        }
        ck_free(&umi_cnt);
        ck_free(&reg);
        free(cnt);
        free(posbuf);
    }
    orc_coo *outs[3] = {ad, dp, oth};
    for (int32_t g = 0; g < n_regions && !err; g++)
        for (int w = 0; w < 3; w++) {
            orc_coo *p = &parts[g * 3 + w];
            for (int64_t k = 0; k < p->nnz; k++)
                if (coo_push(outs[w], p->row[k], p->col[k], p->val[k])) err = 1;
        }
    for (int32_t g = 0; g < n_regions * 3; g++) orc_coo_free(&parts[g]);
    free(parts);
    free(bc.keys);
    free(bc.cols);
    return err ? -1 : 0;
}

// baf.cu -- feature-level allele counting at phased heterozygous SNPs (AD / DP / OTH matrices).
//
// Reference being replaced (xcltk v0.5.2):
//   plp_snp                  xcltk/baf/fc/core.py:198-247   pileup of one SNP over all BAMs
//   SCount.push_read         xcltk/baf/fc/mcount.py:109-127 first read per (SNP, cell, UMI) wins
//   UCount.push_read         xcltk/baf/fc/mcount.py:39-60   + get_query_bases utils/sam.py:4-40
//   SCount.stat/MCount.stat  xcltk/baf/fc/mcount.py:140-150,250-256  A/C/G/T/N totals per SNP
//   fc_fet1                  xcltk/baf/fc/core.py:143-194   region-level UMI set algebra
//
// Phase 1 (xg_baf_pileup): stream the reads once; a tile whose window holds no SNP is skipped
// without touching its records; a read overlapping SNPs emits (SNP, cell, UMI, ordinal, base)
// pairs.  First-read-wins = atomicMin of the record ordinal per (SNP, cell, UMI) key in a
// 128-bit-CAS hash table -- decided BEFORE looking at the base, so a read that skips the SNP
// (N / D) still claims the UMI.  Winners with a base add to the SNP totals.
// The caller applies the min_count / min_maf filter in Python (float exactness, B8).
// Phase 2 (xg_baf_count): winners of kept SNPs OR a 3-bit haplotype mask into
// (region, cell, UMI); the masks reduce to ref / alt / shared / other UMI counts per
// (region, cell) and then to AD / DP / OTH (B9, B10).
#include <algorithm>
#include <cstring>

#include "compact.cuh"

struct xg_baf_state {
    int64_t n_pairs = 0;
    int32_t n_snps = 0, n_cols = 0;
    uint32_t *pr_snp = nullptr;      // SNP index (caller's order)
    uint32_t *pr_colal = nullptr;    // col | base code << 24 | winner << 31
    uint64_t *pr_umi = nullptr;
};

namespace {

enum { CODE_NONE = 6, CODE_OTHER = 5 };

struct BafScanDev {
    const int2 *pos_end;
    const uint32_t *fmq, *cig_off, *cigar, *seq_off, *seq;
    const ulonglong2 *keys;
    const xg_run *runs;
    const xg_tile *tiles;
    int32_t n_gid;
    const int32_t *snp_goff;    // per gid range in the (gid, pos)-sorted SNP arrays
    const int32_t *snp_pos;
    const int32_t *snp_idx;     // caller's SNP index
    BarcodeTable bc;
    FilterParams fp;
    // pair output
    uint32_t *pr_snp, *pr_colal;
    uint64_t *pr_umi, *pr_ord;
    unsigned long long *n_pairs;
    unsigned long long cap_pairs;
};

// base code at reference position sp: 0..3 ACGT, 4 N, 5 other (=, IUPAC), 6 not covered (None)
__device__ __forceinline__ uint32_t base_at(const BafScanDev &P, int64_t i, int32_t pos, int32_t end,
                                            uint32_t n_ops, const uint32_t *cig, int32_t sp) {
    int32_t q = -1;
    if (n_ops == 0) {
        q = sp - pos;          // single M block covering [pos, end)
    } else {
        int32_t p = pos, qi = 0;
        for (uint32_t k = 0; k < n_ops; k++) {
            uint32_t w = __ldg(&cig[k]), op = w & 15u;
            int32_t l = (int32_t)(w >> 4);
            if (cig_aligned(op)) {
                if (sp >= p && sp < p + l) {
                    q = qi + (sp - p);
                    break;
                }
                p += l;
                qi += l;
            } else if (cig_skips_ref(op)) {
                p += l;
            } else if (op == 1 || op == 4) {
                qi += l;
            }
        }
    }
    if (q < 0) return CODE_NONE;
    uint32_t so = P.seq_off[i];
    if (so == 0xFFFFFFFFu) return CODE_NONE;   // record without sequence
    uint32_t byte = (uint32_t)q >> 1;
    uint32_t w = __ldg(&P.seq[so + (byte >> 2)]);
    uint32_t nib = (w >> (8 * (byte & 3u) + ((q & 1) ? 0u : 4u))) & 15u;
    switch (nib) {
        case 1: return 0;
        case 2: return 1;
        case 4: return 2;
        case 8: return 3;
        case 15: return 4;
        default: return CODE_OTHER;
    }
}

__global__ void __launch_bounds__(256) k_baf_scan(const __grid_constant__ BafScanDev P) {
    const xg_tile tile = P.tiles[blockIdx.x];
    const xg_run run = P.runs[tile.run];
    const int32_t gid = run.gid;
    if (gid < 0 || gid >= P.n_gid) return;
    // SNPs inside the tile window [first_pos, max_end)
    int32_t sa, sb;
    {
        int32_t g0 = P.snp_goff[gid], g1 = P.snp_goff[gid + 1];
        int32_t lo = g0, hi = g1;
        while (lo < hi) {
            int32_t mid = (lo + hi) >> 1;
            if (__ldg(&P.snp_pos[mid]) < tile.first_pos) lo = mid + 1; else hi = mid;
        }
        sa = lo;
        hi = g1;
        while (lo < hi) {
            int32_t mid = (lo + hi) >> 1;
            if (__ldg(&P.snp_pos[mid]) < tile.max_end) lo = mid + 1; else hi = mid;
        }
        sb = lo;
    }
    if (sa == sb) return;     // no SNP under this tile: its records are never read

    for (int32_t k = threadIdx.x; k < tile.n_rec; k += blockDim.x) {
        const int64_t i = tile.rec_beg + k;
        const int2 pe = P.pos_end[i];
        // first SNP with pos >= read.pos (fetch(chrom, pos-1, pos): pos0 in [read.pos, read.end))
        int32_t lo = sa, hi = sb;
        while (lo < hi) {
            int32_t mid = (lo + hi) >> 1;
            if (__ldg(&P.snp_pos[mid]) < pe.x) lo = mid + 1; else hi = mid;
        }
        if (lo >= sb || __ldg(&P.snp_pos[lo]) >= pe.y) continue;
        const uint32_t fmq = P.fmq[i];
        if (!read_passes_flags(P.fp, fmq)) continue;
        const ulonglong2 ky = P.keys[i];
        const uint64_t umi = ky.y;
        if (umi == XG_KEY_NONE || umi == XG_KEY_EMPTY) continue;
        uint32_t col;
        if (P.fp.use_cell_tag) {
            if (ky.x == XG_KEY_NONE) continue;
            int32_t c = barcode_lookup(P.bc, ky.x);
            if (c < 0) continue;
            col = (uint32_t)c;
        } else {
            col = (uint32_t)run.bam_idx;
        }
        uint32_t n_ops = fmq >> 24;
        const uint32_t *cig = nullptr;
        int32_t aln;
        if (n_ops == 0) {
            aln = pe.y - pe.x;
        } else {
            cig = P.cigar + P.cig_off[i];
            if (n_ops == 255) n_ops = __ldg(cig - 1);
            aln = 0;
            for (uint32_t q = 0; q < n_ops; q++) {
                uint32_t w = __ldg(&cig[q]);
                if (cig_aligned(w & 15u)) aln += (int32_t)(w >> 4);
            }
        }
        if (aln < P.fp.min_len) continue;
        for (int32_t s = lo; s < sb; s++) {
            const int32_t sp = __ldg(&P.snp_pos[s]);
            if (sp >= pe.y) break;
            const uint32_t code = base_at(P, i, pe.x, pe.y, n_ops, cig, sp);
            unsigned long long o = atomicAdd(P.n_pairs, 1ull);
            if (o < P.cap_pairs) {
                P.pr_snp[o] = (uint32_t)__ldg(&P.snp_idx[s]);
                P.pr_colal[o] = col | (code << 24);
                P.pr_umi[o] = umi;
                P.pr_ord[o] = (uint64_t)i;
            }
        }
    }
}

// find-or-insert of a 128-bit key; returns the slot
__device__ __forceinline__ uint32_t table_slot(xg_e128 *tbl, uint32_t cap, xg_e128 want) {
    uint32_t s = hash_to_range(mix64(want.a ^ (want.b * 0x9E3779B97F4A7C15ULL)), cap);
    while (true) {
        xg_e128 cur = ld128_relaxed(&tbl[s]);
        if (cur.b == 0) {
            xg_e128 empty;
            empty.a = 0;
            empty.b = 0;
            cur = cas128(&tbl[s], empty, want);
            if (cur.b == 0) return s;
        }
        if (cur.a == want.a && cur.b == want.b) return s;
        s = (s + 1 == cap) ? 0 : s + 1;
    }
}

__global__ void k_baf_first(int64_t n, const uint32_t *pr_snp, const uint32_t *pr_colal,
                            const uint64_t *pr_umi, const uint64_t *pr_ord, xg_e128 *tbl, uint32_t cap,
                            unsigned long long *min_ord, uint32_t *pr_slot) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    xg_e128 want;
    want.a = pr_umi[p];
    want.b = ((unsigned long long)(pr_colal[p] & 0xFFFFFFu) << 32) | ((unsigned long long)pr_snp[p] + 1ull);
    uint32_t s = table_slot(tbl, cap, want);
    pr_slot[p] = s;
    atomicMin(&min_ord[s], (unsigned long long)pr_ord[p]);
}

__global__ void k_baf_resolve(int64_t n, const uint32_t *pr_snp, uint32_t *pr_colal,
                              const uint64_t *pr_ord, const unsigned long long *min_ord,
                              const uint32_t *pr_slot, unsigned long long *totals) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    if (min_ord[pr_slot[p]] != pr_ord[p]) return;
    uint32_t ca = pr_colal[p];
    pr_colal[p] = ca | 0x80000000u;              // winner: the first read of this (SNP, cell, UMI)
    uint32_t code = (ca >> 24) & 0x7u;
    if (code == CODE_NONE) return;               // allele None: claims the UMI, counts nowhere
    uint32_t bucket = code < 4 ? code : 4;       // non-ACGT -> "N" bucket (mcount.py:145-149)
    atomicAdd(&totals[(size_t)pr_snp[p] * 5 + bucket], 1ull);
}

struct BafRegDev {
    int64_t n_pairs;
    const uint32_t *pr_snp, *pr_colal;
    const uint64_t *pr_umi;
    const int64_t *snp_reg_ptr;
    const int32_t *snp_reg;
    const uint8_t *hap_of, *keep;
    xg_e128 *tbl;
    uint32_t cap;
    uint32_t *mask;
};

__global__ void k_baf_count_combos(BafRegDev P, unsigned long long *total) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long c = 0;
    if (p < P.n_pairs) {
        uint32_t ca = P.pr_colal[p], snp = P.pr_snp[p];
        if ((ca & 0x80000000u) && ((ca >> 24) & 7u) != CODE_NONE && P.keep[snp])
            c = (unsigned long long)(P.snp_reg_ptr[snp + 1] - P.snp_reg_ptr[snp]);
    }
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(total, c);
}

__global__ void k_baf_region_masks(BafRegDev P) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_pairs) return;
    uint32_t ca = P.pr_colal[p], snp = P.pr_snp[p];
    uint32_t code = (ca >> 24) & 7u;
    if (!(ca & 0x80000000u) || code == CODE_NONE || !P.keep[snp]) return;
    uint32_t bit = 1u << P.hap_of[(size_t)snp * 8 + code];     // 1 ref-hap, 2 alt-hap, 4 other
    uint32_t col = ca & 0xFFFFFFu;
    uint64_t umi = P.pr_umi[p];
    for (int64_t k = P.snp_reg_ptr[snp]; k < P.snp_reg_ptr[snp + 1]; k++) {
        xg_e128 want;
        want.a = umi;
        want.b = ((unsigned long long)col << 32) | ((unsigned long long)P.snp_reg[k] + 1ull);
        uint32_t s = table_slot(P.tbl, P.cap, want);
        atomicOr(&P.mask[s], bit);
    }
}

// masks of (region, cell, UMI) -> ref / alt / share / oth counters of (region, cell), rows [r0, r1)
__global__ void k_baf_accumulate(const xg_e128 *tbl, const uint32_t *mask, uint32_t cap, int32_t r0,
                                 int32_t r1, int32_t n_cols, uint32_t *cnt /* [4][rows][cols] */) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= cap) return;
    xg_e128 e = tbl[s];
    if (e.b == 0) return;
    int32_t r = (int32_t)((e.b & 0xFFFFFFFFull) - 1ull);
    if (r < r0 || r >= r1) return;
    uint32_t col = (uint32_t)(e.b >> 32), m = mask[s];
    size_t plane = (size_t)(r1 - r0) * (size_t)n_cols, o = (size_t)(r - r0) * (size_t)n_cols + col;
    if (m & 1u) atomicAdd(&cnt[o], 1u);
    if (m & 2u) atomicAdd(&cnt[plane + o], 1u);
    if ((m & 3u) == 3u) atomicAdd(&cnt[2 * plane + o], 1u);
    if ((m & 3u) == 0u && (m & 4u)) atomicAdd(&cnt[3 * plane + o], 1u);
}

struct BafVal {
    const uint32_t *cnt;
    size_t plane;
    int32_t n_cols, which, no_dup;
    __device__ int operator()(int row, int col) const {
        size_t o = (size_t)row * (size_t)n_cols + col;
        int ref = (int)cnt[o], alt = (int)cnt[plane + o], share = (int)cnt[2 * plane + o];
        if (which == 2) return (int)cnt[3 * plane + o];
        if (no_dup) {
            ref -= share;
            alt -= share;
        }
        return which == 0 ? alt : ref + alt;
    }
};

// host-side concatenation of per-chunk results (rows offset by the chunk start)
struct CooAccum {
    std::vector<int32_t> row, col, val;
    std::vector<int64_t> row_ptr{0};
    void append(const xg_coo *m, int32_t r0) {
        for (int64_t k = 0; k < m->nnz; k++) {
            row.push_back(m->row[k] + r0);
            col.push_back(m->col[k]);
            val.push_back(m->val[k]);
        }
        int64_t base = row_ptr.back();
        for (int32_t r = 0; r < m->n_rows; r++) row_ptr.push_back(base + m->row_ptr[r + 1]);
    }
    int finish(xg_ctx *ctx, int32_t n_rows, int32_t n_cols, xg_coo **out) {
        while ((int32_t)row_ptr.size() < n_rows + 1) row_ptr.push_back(row_ptr.back());
        xg_coo_owner *o = new xg_coo_owner();
        memset(&o->m, 0, sizeof(o->m));
        size_t nnz = row.size();
        void *p[4] = {nullptr, nullptr, nullptr, nullptr};
        size_t sz[4] = {(nnz + 1) * 4, (nnz + 1) * 4, (nnz + 1) * 4, row_ptr.size() * 8};
        for (int k = 0; k < 4; k++)
            if (!(p[k] = ctx->pinned_get(sz[k]))) {
                for (int q = 0; q < k; q++) ctx->pinned_put(p[q]);
                delete o;
                return ctx->fail(XG_E_NOMEM, "out of pinned host memory");
            }
        o->ctx = ctx;
        if (nnz) {
            memcpy(p[0], row.data(), nnz * 4);
            memcpy(p[1], col.data(), nnz * 4);
            memcpy(p[2], val.data(), nnz * 4);
        }
        memcpy(p[3], row_ptr.data(), row_ptr.size() * 8);
        o->bufs = {p[0], p[1], p[2], p[3]};
        o->m.nnz = (int64_t)nnz;
        o->m.n_rows = n_rows;
        o->m.n_cols = n_cols;
        o->m.row = (const int32_t *)p[0];
        o->m.col = (const int32_t *)p[1];
        o->m.val = (const int32_t *)p[2];
        o->m.row_ptr = (const int64_t *)p[3];
        *out = &o->m;
        return XG_OK;
    }
};

template <class T>
int upload_arr(xg_ctx *ctx, const T *src, size_t n, const char *name, const T **out) {
    T *d = (T *)ctx->get(name, sizeof(T) * (n + 1));
    if (!d) return XG_E_CUDA;
    if (n) XG_CUDA(cudaMemcpyAsync(d, src, sizeof(T) * n, cudaMemcpyHostToDevice, ctx->stream));
    *out = d;
    return XG_OK;
}

}  // namespace

extern "C" void xg_baf_state_free(xg_ctx *ctx, xg_baf_state *st) {
    if (!st) return;
    if (ctx) cudaSetDevice(ctx->device);
    for (void *p : {(void *)st->pr_snp, (void *)st->pr_colal, (void *)st->pr_umi})
        if (p) {
            if (ctx) ctx->dev_put(p); else cudaFree(p);
        }
    delete st;
}

extern "C" int xg_baf_pileup(xg_ctx *ctx, const xg_dreads *rd, const xg_snps *snps,
                             const xg_barcodes *cells, const xg_params *par, int64_t *totals,
                             xg_baf_state **state) {
    if (!ctx || !ctx->stream) return ctx ? ctx->fail(XG_E_CUDA, "context has no device") : XG_E_ARG;
    if (!rd || !snps || !cells || !par || !totals || !state)
        return ctx->fail(XG_E_ARG, "xg_baf_pileup: null argument");
    if (!rd->seq || !rd->seq_off) return ctx->fail(XG_E_ARG, "xg_baf_pileup: reads were decoded without sequences");
    if (cells->n_samples <= 0 || cells->n_samples >= (1 << 24))
        return ctx->fail(XG_E_ARG, "xg_baf_pileup: number of cells must be in [1, 2^24)");
    if (par->use_cell_tag && cells->n != cells->n_samples)
        return ctx->fail(XG_E_ARG, "xg_baf_pileup: barcode mode needs one key per column");
    XG_CUDA(cudaSetDevice(ctx->device));
    for (double &t : ctx->timing) t = 0;
    int launches = 0;

    int32_t n_gid = 0;
    for (auto &r : rd->h_runs) n_gid = std::max(n_gid, r.gid + 1);
    if (!par->use_cell_tag)
        for (auto &r : rd->h_runs)
            if (r.bam_idx >= cells->n_samples) return ctx->fail(XG_E_ARG, "more BAMs than sample columns");
    // SNPs sorted by (gid, pos); pos < 0 (1-based pos <= 0) can never be fetched
    std::vector<int32_t> ord;
    for (int32_t i = 0; i < snps->n; i++)
        if (snps->gid[i] >= 0 && snps->gid[i] < n_gid && snps->pos[i] >= 0) ord.push_back(i);
    std::sort(ord.begin(), ord.end(), [&](int32_t a, int32_t b) {
        if (snps->gid[a] != snps->gid[b]) return snps->gid[a] < snps->gid[b];
        if (snps->pos[a] != snps->pos[b]) return snps->pos[a] < snps->pos[b];
        return a < b;
    });
    std::vector<int32_t> goff((size_t)n_gid + 1, 0), spos(ord.size()), sidx(ord.size());
    for (size_t k = 0; k < ord.size(); k++) {
        goff[(size_t)snps->gid[ord[k]] + 1]++;
        spos[k] = snps->pos[ord[k]];
        sidx[k] = ord[k];
    }
    for (int32_t g = 0; g < n_gid; g++) goff[(size_t)g + 1] += goff[(size_t)g];

    BafScanDev P;
    memset(&P, 0, sizeof(P));
    P.pos_end = rd->pos_end;
    P.fmq = rd->fmq;
    P.cig_off = rd->cig_off;
    P.cigar = rd->cigar;
    P.seq_off = rd->seq_off;
    P.seq = rd->seq;
    P.keys = rd->keys;
    P.runs = rd->runs;
    P.tiles = rd->tiles;
    P.n_gid = n_gid;
    int rc;
    if ((rc = upload_arr(ctx, goff.data(), goff.size(), "bf_goff", &P.snp_goff))) return rc;
    if ((rc = upload_arr(ctx, spos.data(), spos.size(), "bf_spos", &P.snp_pos))) return rc;
    if ((rc = upload_arr(ctx, sidx.data(), sidx.size(), "bf_sidx", &P.snp_idx))) return rc;
    P.fp.min_mapq = par->min_mapq;
    P.fp.min_len = par->min_len;
    P.fp.incl_flag = par->incl_flag;
    P.fp.excl_flag = par->excl_flag;
    P.fp.no_orphan = par->no_orphan;
    P.fp.use_cell_tag = par->use_cell_tag;
    P.fp.need_umi_tag = par->need_umi_tag;
    if (par->use_cell_tag && (rc = xg_build_barcode_table(ctx, cells, &P.bc))) return rc;
    XG_GET(d_npairs, unsigned long long, "bf_npairs", 2);
    XG_GET(d_totals, unsigned long long, "bf_totals", (size_t)snps->n * 5 + 1);
    XG_CUDA(cudaStreamSynchronize(ctx->stream));

    cudaEventRecord(ctx->ev[0], ctx->stream);
    // scan; the pair buffer grows and the scan is repeated in the (rare) overflow case
    unsigned long long cap_pairs = std::max<unsigned long long>(1ull << 16, (unsigned long long)rd->n_reads / 4);
    unsigned long long n_pairs = 0;
    float t_scan = 0;
    for (int attempt = 0; attempt < 8; attempt++) {
        XG_GET(pr_snp, uint32_t, "bf_pr_snp", cap_pairs);
        XG_GET(pr_colal, uint32_t, "bf_pr_colal", cap_pairs);
        XG_GET(pr_umi, uint64_t, "bf_pr_umi", cap_pairs);
        XG_GET(pr_ord, uint64_t, "bf_pr_ord", cap_pairs);
        P.pr_snp = pr_snp;
        P.pr_colal = pr_colal;
        P.pr_umi = pr_umi;
        P.pr_ord = pr_ord;
        P.n_pairs = d_npairs;
        P.cap_pairs = cap_pairs;
        XG_CUDA(cudaMemsetAsync(d_npairs, 0, 16, ctx->stream));
        cudaEventRecord(ctx->ev[1], ctx->stream);
        if (rd->n_tiles > 0 && !ord.empty()) {
            k_baf_scan<<<rd->n_tiles, 256, 0, ctx->stream>>>(P);
            launches++;
            XG_CUDA(cudaGetLastError());
        }
        cudaEventRecord(ctx->ev[2], ctx->stream);
        XG_CUDA(cudaMemcpyAsync(&n_pairs, d_npairs, 8, cudaMemcpyDeviceToHost, ctx->stream));
        XG_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaEventElapsedTime(&t_scan, ctx->ev[1], ctx->ev[2]);
        if (n_pairs <= cap_pairs) break;
        cap_pairs = n_pairs + n_pairs / 8 + 1024;
    }
    if (n_pairs > cap_pairs) return ctx->fail(XG_E_LIMIT, "pair buffer overflow");
    if (n_pairs >= (1ull << 31)) return ctx->fail(XG_E_LIMIT, "more than 2^31 (read, SNP) pairs in one batch");

    XG_CUDA(cudaMemsetAsync(d_totals, 0, sizeof(unsigned long long) * ((size_t)snps->n * 5 + 1), ctx->stream));
    xg_baf_state *st = new xg_baf_state();
    st->n_pairs = (int64_t)n_pairs;
    st->n_snps = snps->n;
    st->n_cols = cells->n_samples;
    if (n_pairs > 0) {
        const uint32_t cap = (uint32_t)(2 * n_pairs + 16);
        XG_GET(tbl, xg_e128, "bf_tbl1", cap);
        XG_GET(min_ord, unsigned long long, "bf_minord", cap);
        XG_GET(pr_slot, uint32_t, "bf_pr_slot", n_pairs);
        XG_CUDA(cudaMemsetAsync(tbl, 0, sizeof(xg_e128) * cap, ctx->stream));
        XG_CUDA(cudaMemsetAsync(min_ord, 0xFF, sizeof(unsigned long long) * cap, ctx->stream));
        unsigned grid = (unsigned)((n_pairs + 255) / 256);
        k_baf_first<<<grid, 256, 0, ctx->stream>>>((int64_t)n_pairs, P.pr_snp, P.pr_colal, P.pr_umi,
                                                   P.pr_ord, tbl, cap, min_ord, pr_slot);
        k_baf_resolve<<<grid, 256, 0, ctx->stream>>>((int64_t)n_pairs, P.pr_snp, P.pr_colal, P.pr_ord,
                                                     min_ord, pr_slot, d_totals);
        launches += 4;
        XG_CUDA(cudaGetLastError());
        // the state keeps its own copy of the pairs (scratch buffers are reused by later calls)
        st->pr_snp = (uint32_t *)ctx->dev_get(n_pairs * 4);
        st->pr_colal = (uint32_t *)ctx->dev_get(n_pairs * 4);
        st->pr_umi = (uint64_t *)ctx->dev_get(n_pairs * 8);
        if (!st->pr_snp || !st->pr_colal || !st->pr_umi) {
            xg_baf_state_free(ctx, st);
            return ctx->fail(XG_E_CUDA, "out of device memory for the pileup state");
        }
        cudaMemcpyAsync(st->pr_snp, P.pr_snp, n_pairs * 4, cudaMemcpyDeviceToDevice, ctx->stream);
        cudaMemcpyAsync(st->pr_colal, P.pr_colal, n_pairs * 4, cudaMemcpyDeviceToDevice, ctx->stream);
        cudaMemcpyAsync(st->pr_umi, P.pr_umi, n_pairs * 8, cudaMemcpyDeviceToDevice, ctx->stream);
    }
    cudaEventRecord(ctx->ev[3], ctx->stream);
    std::vector<unsigned long long> h_tot((size_t)snps->n * 5 + 1);
    cudaEventRecord(ctx->ev[4], ctx->stream);
    cudaMemcpyAsync(h_tot.data(), d_totals, sizeof(unsigned long long) * (size_t)snps->n * 5,
                    cudaMemcpyDeviceToHost, ctx->stream);
    cudaEventRecord(ctx->ev[5], ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        xg_baf_state_free(ctx, st);
        return ctx->fail(XG_E_CUDA, std::string("baf pileup: ") + cudaGetErrorString(e));
    }
    for (size_t k = 0; k < (size_t)snps->n * 5; k++) totals[k] = (int64_t)h_tot[k];
    float t_all = 0, t_d2h = 0;
    cudaEventElapsedTime(&t_all, ctx->ev[0], ctx->ev[3]);
    cudaEventElapsedTime(&t_d2h, ctx->ev[4], ctx->ev[5]);
    ctx->timing[0] = t_all;
    ctx->timing[1] = t_scan;
    ctx->timing[2] = launches;
    ctx->timing[4] = t_d2h;
    ctx->timing[6] = (double)n_pairs;
    *state = st;
    return XG_OK;
}

extern "C" int xg_baf_count(xg_ctx *ctx, xg_baf_state *st, int32_t n_regions, const int64_t *reg_ptr,
                            const int32_t *reg_snp, const uint8_t *hap_of, const uint8_t *keep,
                            int32_t no_dup_hap, xg_coo **ad, xg_coo **dp, xg_coo **oth) {
    if (!ctx || !ctx->stream) return ctx ? ctx->fail(XG_E_CUDA, "context has no device") : XG_E_ARG;
    if (!st || n_regions < 0 || !reg_ptr || !hap_of || !keep || !ad || !dp || !oth)
        return ctx->fail(XG_E_ARG, "xg_baf_count: bad argument");
    XG_CUDA(cudaSetDevice(ctx->device));
    for (double &t : ctx->timing) t = 0;
    int launches = 0;
    const int32_t n_cols = st->n_cols, n_snps = st->n_snps;
    // invert region -> SNP lists
    std::vector<int64_t> sr_ptr((size_t)n_snps + 1, 0);
    int64_t n_mem = reg_ptr[n_regions];
    for (int64_t k = 0; k < n_mem; k++) {
        if (reg_snp[k] < 0 || reg_snp[k] >= n_snps) return ctx->fail(XG_E_ARG, "reg_snp out of range");
        sr_ptr[(size_t)reg_snp[k] + 1]++;
    }
    for (int32_t s = 0; s < n_snps; s++) sr_ptr[(size_t)s + 1] += sr_ptr[(size_t)s];
    std::vector<int32_t> sr((size_t)n_mem);
    {
        std::vector<int64_t> cur(sr_ptr.begin(), sr_ptr.end() - 1);
        for (int32_t r = 0; r < n_regions; r++)
            for (int64_t k = reg_ptr[r]; k < reg_ptr[r + 1]; k++) sr[(size_t)cur[(size_t)reg_snp[k]]++] = r;
    }
    for (size_t k = 0; k < (size_t)n_snps * 8; k++)
        if (hap_of[k] > 2) return ctx->fail(XG_E_ARG, "hap_of entries must be 0, 1 or 2");

    BafRegDev R;
    memset(&R, 0, sizeof(R));
    R.n_pairs = st->n_pairs;
    R.pr_snp = st->pr_snp;
    R.pr_colal = st->pr_colal;
    R.pr_umi = st->pr_umi;
    int rc;
    if ((rc = upload_arr(ctx, sr_ptr.data(), sr_ptr.size(), "bf_sr_ptr", &R.snp_reg_ptr))) return rc;
    if ((rc = upload_arr(ctx, sr.data(), sr.size(), "bf_sr", &R.snp_reg))) return rc;
    if ((rc = upload_arr(ctx, hap_of, (size_t)n_snps * 8, "bf_hap_of", &R.hap_of))) return rc;
    if ((rc = upload_arr(ctx, keep, (size_t)n_snps, "bf_keep", &R.keep))) return rc;
    XG_GET(d_total, unsigned long long, "bf_combo_total", 2);
    XG_CUDA(cudaStreamSynchronize(ctx->stream));

    cudaEventRecord(ctx->ev[0], ctx->stream);
    unsigned long long combos = 0;
    unsigned grid = (unsigned)((st->n_pairs + 255) / 256);
    XG_CUDA(cudaMemsetAsync(d_total, 0, 16, ctx->stream));
    if (st->n_pairs > 0) {
        k_baf_count_combos<<<grid, 256, 0, ctx->stream>>>(R, d_total);
        launches++;
    }
    XG_CUDA(cudaMemcpyAsync(&combos, d_total, 8, cudaMemcpyDeviceToHost, ctx->stream));
    XG_CUDA(cudaStreamSynchronize(ctx->stream));
    if (combos >= (1ull << 31)) return ctx->fail(XG_E_LIMIT, "more than 2^31 (region, cell, UMI) candidates");
    const uint32_t cap = (uint32_t)(2 * combos + 16);
    XG_GET(tbl, xg_e128, "bf_tbl2", cap);
    XG_GET(mask, uint32_t, "bf_mask", cap);
    XG_CUDA(cudaMemsetAsync(tbl, 0, sizeof(xg_e128) * cap, ctx->stream));
    XG_CUDA(cudaMemsetAsync(mask, 0, sizeof(uint32_t) * cap, ctx->stream));
    R.tbl = tbl;
    R.cap = cap;
    R.mask = mask;
    if (combos > 0) {
        k_baf_region_masks<<<grid, 256, 0, ctx->stream>>>(R);
        launches++;
        XG_CUDA(cudaGetLastError());
    }
    // dense (region, cell) counters in row chunks of bounded size
    const size_t budget = (size_t)1 << 31;    // bytes for the 4 counter planes
    int32_t rows_per_chunk = (int32_t)std::max<size_t>(1, budget / ((size_t)n_cols * 16));
    rows_per_chunk = std::min(rows_per_chunk, std::max(n_regions, 1));
    XG_GET(cnt, uint32_t, "bf_cnt", (size_t)rows_per_chunk * (size_t)n_cols * 4);
    CooAccum acc[3];
    for (int32_t r0 = 0; r0 < n_regions; r0 += rows_per_chunk) {
        int32_t r1 = std::min(n_regions, r0 + rows_per_chunk), nr = r1 - r0;
        size_t plane = (size_t)nr * (size_t)n_cols;
        XG_CUDA(cudaMemsetAsync(cnt, 0, plane * 16, ctx->stream));
        if (combos > 0) {
            k_baf_accumulate<<<(cap + 255) / 256, 256, 0, ctx->stream>>>(tbl, mask, cap, r0, r1, n_cols, cnt);
            launches++;
        }
        for (int which = 0; which < 3; which++) {
            BafVal v{cnt, plane, n_cols, which, no_dup_hap};
            xg_coo *part = nullptr;
            rc = xg_dense_to_coo(ctx, v, nr, n_cols, "bf", &part, &launches);
            if (rc) return rc;
            acc[which].append(part, r0);
            xg_coo_free(part);
        }
    }
    cudaEventRecord(ctx->ev[3], ctx->stream);
    XG_CUDA(cudaStreamSynchronize(ctx->stream));
    float t_all = 0;
    cudaEventElapsedTime(&t_all, ctx->ev[0], ctx->ev[3]);
    ctx->timing[0] = t_all - ctx->timing[4];
    ctx->timing[2] = launches;
    xg_coo **outs[3] = {ad, dp, oth};
    for (int which = 0; which < 3; which++)
        if ((rc = acc[which].finish(ctx, n_regions, n_cols, outs[which]))) return rc;
    return XG_OK;
}

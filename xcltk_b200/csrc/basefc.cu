// basefc.cu -- per-feature, per-cell distinct-UMI counting (the RDR total-depth matrix).
//
// Reference being replaced (xcltk v0.5.2):
//   fc_features / fc_fet1      xcltk/rdr/fc/core.py:69-178   for feature: fetch reads, filter, count
//   check_read                 xcltk/rdr/fc/core.py:46-62
//   __get_include_frac/_len    xcltk/rdr/fc/core.py:32-43, used :160-165
//   MCount/SCount.push_read    xcltk/rdr/fc/mcount.py:34-43,102-132 (cell lookup, UMI set)
//   sam_fetch                  xcltk/utils/sam.py:85-118  (reads overlapping the feature)
//
// The reference walks features and re-fetches the reads of each one.  Here the reads are
// streamed ONCE in file order (coalesced, one CTA per tile of <= 1024 records); each read finds
// the features it overlaps through a per-contig interval index (sorted boundaries + per-segment
// stabbing lists + start-sorted features), evaluates the include test arithmetically on its CIGAR
// and inserts (cell, UMI) into the feature's open-addressing set in HBM with a 128-bit CAS.
// A new element bumps the dense (feature, cell) counter; the counters are then compacted into
// (row, col)-sorted COO.  Features are counted independently (a read overlapping k features is
// evaluated k times), exactly as the reference does.
#include <algorithm>
#include <cstring>
#include <numeric>

#include "compact.cuh"

namespace {

struct FeatIndexHost {
    int32_t n_gid = 0;
    std::vector<int32_t> sf_goff, sf_beg, sf_end, sf_row;
    std::vector<int32_t> bnd_goff, bnd, stab_off, stab;
};

// Interval index over the valid features of every contig.
int build_feat_index(xg_ctx *ctx, const xg_features *f, int32_t n_gid, FeatIndexHost &ix) {
    ix.n_gid = n_gid;
    std::vector<int32_t> order;
    order.reserve((size_t)f->n);
    for (int32_t i = 0; i < f->n; i++) {
        // never fetched: unknown contig, start <= 0 (fetch raises), empty interval
        if (f->gid[i] < 0 || f->gid[i] >= n_gid || f->beg[i] < 0 || f->end[i] <= f->beg[i]) continue;
        order.push_back(i);
    }
    std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
        if (f->gid[a] != f->gid[b]) return f->gid[a] < f->gid[b];
        if (f->beg[a] != f->beg[b]) return f->beg[a] < f->beg[b];
        if (f->end[a] != f->end[b]) return f->end[a] < f->end[b];
        return a < b;
    });
    size_t m = order.size();
    ix.sf_beg.resize(m);
    ix.sf_end.resize(m);
    ix.sf_row.resize(m);
    ix.sf_goff.assign((size_t)n_gid + 1, 0);
    for (size_t j = 0; j < m; j++) {
        int32_t i = order[j];
        ix.sf_beg[j] = f->beg[i];
        ix.sf_end[j] = f->end[i];
        ix.sf_row[j] = i;
        ix.sf_goff[(size_t)f->gid[i] + 1]++;
    }
    for (int32_t g = 0; g < n_gid; g++) ix.sf_goff[g + 1] += ix.sf_goff[g];
    ix.bnd_goff.assign((size_t)n_gid + 1, 0);
    ix.stab_off.clear();
    for (int32_t g = 0; g < n_gid; g++) {
        int32_t j0 = ix.sf_goff[g], j1 = ix.sf_goff[g + 1];
        std::vector<int32_t> b;
        b.reserve(2 * (size_t)(j1 - j0));
        for (int32_t j = j0; j < j1; j++) {
            b.push_back(ix.sf_beg[j]);
            b.push_back(ix.sf_end[j]);
        }
        std::sort(b.begin(), b.end());
        b.erase(std::unique(b.begin(), b.end()), b.end());
        size_t nb = b.size();
        // stabbing list of segment k = [b[k], b[k+1]): features with beg <= b[k] and end >= b[k+1]
        std::vector<int32_t> cnt(nb + 1, 0);
        std::vector<std::pair<int32_t, int32_t>> span((size_t)(j1 - j0));
        int64_t total = 0;
        for (int32_t j = j0; j < j1; j++) {
            int32_t lo = (int32_t)(std::lower_bound(b.begin(), b.end(), ix.sf_beg[j]) - b.begin());
            int32_t hi = (int32_t)(std::lower_bound(b.begin(), b.end(), ix.sf_end[j]) - b.begin());
            span[(size_t)(j - j0)] = {lo, hi};
            for (int32_t k = lo; k < hi; k++) cnt[(size_t)k]++;
            total += hi - lo;
        }
        if ((int64_t)ix.stab.size() + total > (1LL << 30))
            return ctx->fail(XG_E_LIMIT, "feature overlap structure too large (stabbing lists > 2^30)");
        size_t base_seg = ix.bnd.size();
        size_t base_stab = ix.stab.size();
        ix.bnd.insert(ix.bnd.end(), b.begin(), b.end());
        std::vector<int32_t> off(nb + 1, 0);
        for (size_t k = 0; k < nb; k++) off[k + 1] = off[k] + cnt[k];
        ix.stab.resize(base_stab + (size_t)total);
        std::vector<int32_t> cur(off.begin(), off.end() - 1);
        for (int32_t j = j0; j < j1; j++) {
            auto sp = span[(size_t)(j - j0)];
            for (int32_t k = sp.first; k < sp.second; k++) ix.stab[base_stab + (size_t)cur[(size_t)k]++] = j;
        }
        for (size_t k = 0; k < nb; k++) ix.stab_off.push_back((int32_t)(base_stab + (size_t)off[k]));
        (void)base_seg;
        ix.bnd_goff[(size_t)g + 1] = (int32_t)ix.bnd.size();
    }
    ix.stab_off.push_back((int32_t)ix.stab.size());
    return XG_OK;
}

// Upper bound on the reads that can overlap each (sorted) feature, from the tile index:
// tiles of the feature's contig with prefix-max(end) > beg and first_pos < end.
void feature_windows(const xg_dreads *rd, const FeatIndexHost &ix, std::vector<int64_t> &cand) {
    size_t m = ix.sf_beg.size();
    cand.assign(m, 0);
    // tiles are laid out run by run
    size_t nt = rd->h_tiles.size();
    std::vector<int32_t> pmax(nt);
    std::vector<size_t> run_t0((size_t)rd->n_runs + 1, nt);
    for (size_t t = 0; t < nt; t++) {
        int32_t r = rd->h_tiles[t].run;
        bool first = (t == 0) || rd->h_tiles[t - 1].run != r;
        if (first) run_t0[(size_t)r] = t;
        pmax[t] = first ? rd->h_tiles[t].max_end : std::max(pmax[t - 1], rd->h_tiles[t].max_end);
    }
    for (int32_t r = 0; r < rd->n_runs; r++) {
        const xg_run &run = rd->h_runs[(size_t)r];
        if (run.gid < 0 || run.gid >= ix.n_gid || run.rec_end == run.rec_beg) continue;
        size_t t0 = run_t0[(size_t)r];
        size_t t1 = t0 + (size_t)((run.rec_end - run.rec_beg + XG_TILE - 1) / XG_TILE);
        for (int32_t j = ix.sf_goff[run.gid]; j < ix.sf_goff[run.gid + 1]; j++) {
            int32_t beg = ix.sf_beg[(size_t)j], end = ix.sf_end[(size_t)j];
            // lo: first tile with pmax > beg ; hi: first tile with first_pos >= end
            size_t lo = t0, hi = t1, a = t0, b = t1;
            while (a < b) {
                size_t mid = (a + b) / 2;
                if (pmax[mid] > beg) b = mid; else a = mid + 1;
            }
            lo = a;
            a = t0, b = t1;
            while (a < b) {
                size_t mid = (a + b) / 2;
                if (rd->h_tiles[mid].first_pos >= end) b = mid; else a = mid + 1;
            }
            hi = a;
            if (hi > lo) {
                int64_t e = (hi == t1) ? run.rec_end : rd->h_tiles[hi].rec_beg;
                cand[(size_t)j] += e - rd->h_tiles[lo].rec_beg;
            }
        }
    }
}

struct BasefcDev {
    const int2 *pos_end;
    const uint32_t *fmq, *cig_off, *cigar;
    const ulonglong2 *keys;
    const xg_run *runs;
    const xg_tile *tiles;
    const int32_t *sf_goff, *sf_beg, *sf_end, *sf_row, *bnd_goff, *bnd, *stab_off, *stab;
    int32_t n_gid;
    xg_e128 *tbl;
    const uint64_t *tbl_base;
    const uint32_t *tbl_cap;
    uint32_t *counts;
    int32_t n_cols;
    BarcodeTable bc;
    FilterParams fp;
    const int32_t *incl_tab;
    int32_t incl_tab_len, incl_len;
};

// (cell, UMI) -> the feature's set; returns true when the element is new.
__device__ __forceinline__ bool set_insert(xg_e128 *tbl, uint32_t cap, uint64_t umi, uint32_t col) {
    xg_e128 want;
    want.a = umi;
    want.b = (unsigned long long)col + 1ull;       // b == 0 marks an empty slot
    uint32_t s = hash_to_range(mix64(umi ^ ((uint64_t)col * 0x9E3779B97F4A7C15ULL)), cap);
    for (uint32_t probe = 0; probe < cap; probe++) {
        xg_e128 cur = ld128_relaxed(&tbl[s]);
        if (cur.b == 0) {
            xg_e128 empty;
            empty.a = 0;
            empty.b = 0;
            cur = cas128(&tbl[s], empty, want);
            if (cur.b == 0) return true;
        }
        if (cur.a == want.a && cur.b == want.b) return false;
        s = (s + 1 == cap) ? 0 : s + 1;
    }
    return false;   // table full: cannot happen, cap > number of candidate reads
}

// m = number of aligned (M/=/X) reference positions p of the read with s0 <= p < e0
// (== len([x for x in read.positions if s <= x <= e]), rdr/fc/core.py:40-43)
__device__ __forceinline__ int32_t included_len(const uint32_t *cig, uint32_t n_ops, int32_t pos,
                                                 int32_t s0, int32_t e0) {
    int32_t m = 0, p = pos;
    for (uint32_t k = 0; k < n_ops; k++) {
        uint32_t w = __ldg(&cig[k]), op = w & 15u;
        int32_t l = (int32_t)(w >> 4);
        if (cig_aligned(op)) {
            int32_t a = max(p, s0), b = min(p + l, e0);
            if (b > a) m += b - a;
            p += l;
        } else if (cig_skips_ref(op)) {
            p += l;
        }
    }
    return m;
}

__device__ __forceinline__ void count_pair(const BasefcDev &P, int32_t j, int32_t pos, int32_t end,
                                           const uint32_t *cig, uint32_t n_ops, int32_t need,
                                           uint64_t umi, uint32_t col) {
    int32_t s0 = __ldg(&P.sf_beg[j]), e0 = __ldg(&P.sf_end[j]);
    int32_t m;
    if (n_ops == 0) {
        int32_t a = max(pos, s0), b = min(end, e0);
        m = b > a ? b - a : 0;
    } else {
        m = included_len(cig, n_ops, pos, s0, e0);
    }
    if (m < need) return;
    uint32_t cap = __ldg(&P.tbl_cap[j]);
    if (cap == 0) return;
    if (set_insert(P.tbl + __ldg(&P.tbl_base[j]), cap, umi, col)) {
        int32_t row = __ldg(&P.sf_row[j]);
        atomicAdd(&P.counts[(size_t)row * (size_t)P.n_cols + col], 1u);
    }
}

__global__ void __launch_bounds__(256) k_basefc_count(const __grid_constant__ BasefcDev P) {
    const xg_tile tile = P.tiles[blockIdx.x];
    const xg_run run = P.runs[tile.run];
    const int32_t gid = run.gid;
    if (gid < 0 || gid >= P.n_gid) return;
    const int32_t f0 = P.sf_goff[gid], f1 = P.sf_goff[gid + 1];
    if (f0 == f1) return;
    const int32_t b0 = P.bnd_goff[gid], b1 = P.bnd_goff[gid + 1];

    for (int32_t k = threadIdx.x; k < tile.n_rec; k += blockDim.x) {
        const int64_t i = tile.rec_beg + k;
        const int2 pe = P.pos_end[i];
        const uint32_t fmq = P.fmq[i];
        if (!read_passes_flags(P.fp, fmq)) continue;
        const ulonglong2 ky = P.keys[i];
        const uint64_t umi = ky.y;
        if (umi == XG_KEY_NONE || umi == XG_KEY_EMPTY) continue;   // has_tag / `if umi:`
        uint32_t col;
        if (P.fp.use_cell_tag) {
            if (ky.x == XG_KEY_NONE) continue;
            int32_t c = barcode_lookup(P.bc, ky.x);
            if (c < 0) continue;
            col = (uint32_t)c;
        } else {
            col = (uint32_t)run.bam_idx;
        }
        // aligned length = len(read.positions)
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
        const int32_t need = P.incl_tab ? __ldg(&P.incl_tab[min(aln, P.incl_tab_len - 1)]) : P.incl_len;

        // (1) features covering `pos`: stabbing list of the segment that contains it
        {
            int32_t lo = b0, hi = b1;          // upper_bound(bnd, pos)
            while (lo < hi) {
                int32_t mid = (lo + hi) >> 1;
                if (__ldg(&P.bnd[mid]) <= pe.x) lo = mid + 1; else hi = mid;
            }
            int32_t seg = lo - 1;
            if (seg >= b0) {
                int32_t s1 = __ldg(&P.stab_off[seg + 1]);
                for (int32_t s = __ldg(&P.stab_off[seg]); s < s1; s++)
                    count_pair(P, __ldg(&P.stab[s]), pe.x, pe.y, cig, n_ops, need, umi, col);
            }
        }
        // (2) features starting inside (pos, end)
        {
            int32_t lo = f0, hi = f1;          // upper_bound(sf_beg, pos)
            while (lo < hi) {
                int32_t mid = (lo + hi) >> 1;
                if (__ldg(&P.sf_beg[mid]) <= pe.x) lo = mid + 1; else hi = mid;
            }
            for (int32_t j = lo; j < f1 && __ldg(&P.sf_beg[j]) < pe.y; j++)
                count_pair(P, j, pe.x, pe.y, cig, n_ops, need, umi, col);
        }
    }
}

struct DenseCounts {
    const uint32_t *counts;
    int32_t n_cols;
    __device__ int operator()(int row, int col) const {
        return (int)counts[(size_t)row * (size_t)n_cols + col];
    }
};

}  // namespace

template <class T>
static int upload_vec(xg_ctx *ctx, const std::vector<T> &v, const char *name, const T **out) {
    T *d = (T *)ctx->get(name, sizeof(T) * (v.size() + 1));
    if (!d) return XG_E_CUDA;
    if (!v.empty())
        XG_CUDA(cudaMemcpyAsync(d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice, ctx->stream));
    *out = d;
    return XG_OK;
}

extern "C" int xg_basefc(xg_ctx *ctx, const xg_dreads *rd, const xg_features *feats,
                         const xg_barcodes *cells, const xg_params *par, xg_coo **out) {
    if (!ctx || !ctx->stream) return ctx ? ctx->fail(XG_E_CUDA, "context has no device") : XG_E_ARG;
    if (!rd || !feats || !cells || !par || !out) return ctx->fail(XG_E_ARG, "xg_basefc: null argument");
    if (feats->n < 0 || cells->n_samples <= 0) return ctx->fail(XG_E_ARG, "xg_basefc: empty sample list");
    if (par->use_cell_tag && cells->n != cells->n_samples)
        return ctx->fail(XG_E_ARG, "xg_basefc: barcode mode needs one key per column");
    XG_CUDA(cudaSetDevice(ctx->device));
    for (double &t : ctx->timing) t = 0;
    int launches = 0;
    const int32_t n_rows = feats->n, n_cols = cells->n_samples;

    int32_t n_gid = 0;
    for (auto &r : rd->h_runs) n_gid = std::max(n_gid, r.gid + 1);
    if (!par->use_cell_tag)
        for (auto &r : rd->h_runs)
            if (r.bam_idx >= n_cols) return ctx->fail(XG_E_ARG, "more BAMs than sample columns");

    FeatIndexHost ix;
    int rc = build_feat_index(ctx, feats, n_gid, ix);
    if (rc) return rc;
    std::vector<int64_t> cand;
    feature_windows(rd, ix, cand);
    size_t m = ix.sf_beg.size();
    std::vector<uint64_t> tbl_base(m + 1, 0);
    std::vector<uint32_t> tbl_cap(m, 0);
    for (size_t j = 0; j < m; j++) {
        int64_t c = cand[j];
        int64_t cap = c > 0 ? c + c / 4 + 8 : 0;
        if (cap >= (1LL << 32)) return ctx->fail(XG_E_LIMIT, "feature window exceeds 2^32 reads");
        tbl_cap[j] = (uint32_t)cap;
        tbl_base[j + 1] = tbl_base[j] + (uint64_t)cap;
    }
    const uint64_t tbl_total = tbl_base[m];

    BasefcDev P;
    memset(&P, 0, sizeof(P));
    P.pos_end = rd->pos_end;
    P.fmq = rd->fmq;
    P.cig_off = rd->cig_off;
    P.cigar = rd->cigar;
    P.keys = rd->keys;
    P.runs = rd->runs;
    P.tiles = rd->tiles;
    P.n_gid = n_gid;
    P.n_cols = n_cols;
    if ((rc = upload_vec(ctx, ix.sf_goff, "fx_sf_goff", &P.sf_goff))) return rc;
    if ((rc = upload_vec(ctx, ix.sf_beg, "fx_sf_beg", &P.sf_beg))) return rc;
    if ((rc = upload_vec(ctx, ix.sf_end, "fx_sf_end", &P.sf_end))) return rc;
    if ((rc = upload_vec(ctx, ix.sf_row, "fx_sf_row", &P.sf_row))) return rc;
    if ((rc = upload_vec(ctx, ix.bnd_goff, "fx_bnd_goff", &P.bnd_goff))) return rc;
    if ((rc = upload_vec(ctx, ix.bnd, "fx_bnd", &P.bnd))) return rc;
    if ((rc = upload_vec(ctx, ix.stab_off, "fx_stab_off", &P.stab_off))) return rc;
    if ((rc = upload_vec(ctx, ix.stab, "fx_stab", &P.stab))) return rc;
    if ((rc = upload_vec(ctx, tbl_base, "fx_tbl_base", &P.tbl_base))) return rc;
    if ((rc = upload_vec(ctx, tbl_cap, "fx_tbl_cap", &P.tbl_cap))) return rc;
    if (par->min_incl_tab) {
        if (par->min_incl_tab_len <= rd->max_aln_len)
            return ctx->fail(XG_E_ARG, "min_incl_tab shorter than max aligned length + 1");
        std::vector<int32_t> t(par->min_incl_tab, par->min_incl_tab + par->min_incl_tab_len);
        if ((rc = upload_vec(ctx, t, "fx_incl_tab", &P.incl_tab))) return rc;
        P.incl_tab_len = par->min_incl_tab_len;
        XG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    P.incl_len = par->min_incl_len;
    P.fp.min_mapq = par->min_mapq;
    P.fp.min_len = par->min_len;
    P.fp.incl_flag = par->incl_flag;
    P.fp.excl_flag = par->excl_flag;
    P.fp.no_orphan = par->no_orphan;
    P.fp.use_cell_tag = par->use_cell_tag;
    P.fp.need_umi_tag = par->need_umi_tag;
    if (par->use_cell_tag) {
        if ((rc = xg_build_barcode_table(ctx, cells, &P.bc))) return rc;
    }
    XG_GET(tbl, xg_e128, "fx_tbl", tbl_total + 1);
    XG_GET(counts, uint32_t, "fx_counts", (size_t)n_rows * (size_t)n_cols + 1);
    P.tbl = tbl;
    P.counts = counts;
    XG_CUDA(cudaStreamSynchronize(ctx->stream));   // host vectors above are about to die

    cudaEventRecord(ctx->ev[0], ctx->stream);
    XG_CUDA(cudaMemsetAsync(tbl, 0, sizeof(xg_e128) * (size_t)tbl_total, ctx->stream));
    XG_CUDA(cudaMemsetAsync(counts, 0, sizeof(uint32_t) * (size_t)n_rows * (size_t)n_cols, ctx->stream));
    launches += 2;
    cudaEventRecord(ctx->ev[1], ctx->stream);
    if (rd->n_tiles > 0 && m > 0) {
        k_basefc_count<<<rd->n_tiles, 256, 0, ctx->stream>>>(P);
        launches++;
        XG_CUDA(cudaGetLastError());
    }
    cudaEventRecord(ctx->ev[2], ctx->stream);
    DenseCounts dc{counts, n_cols};
    rc = xg_dense_to_coo(ctx, dc, n_rows, n_cols, "fx", out, &launches);
    if (rc) return rc;
    cudaEventRecord(ctx->ev[3], ctx->stream);
    XG_CUDA(cudaStreamSynchronize(ctx->stream));
    float t_all = 0, t_cnt = 0, t_zero = 0;
    cudaEventElapsedTime(&t_all, ctx->ev[0], ctx->ev[3]);
    cudaEventElapsedTime(&t_cnt, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&t_zero, ctx->ev[0], ctx->ev[1]);
    ctx->timing[0] = t_all - ctx->timing[4];   // kernels (incl. zeroing + compaction), excl. result D2H
    ctx->timing[1] = t_cnt;
    ctx->timing[2] = launches;
    ctx->timing[5] = t_zero;
    ctx->timing[6] = (double)tbl_total * sizeof(xg_e128);
    return XG_OK;
}

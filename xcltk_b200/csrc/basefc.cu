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
// and inserts (cell, UMI) into the feature's open-addressing set with a 128-bit CAS.  When the
// last read that can touch a feature has been streamed, one CTA scans the feature's set into a
// shared-memory histogram over cells and writes the row's non-zeros in column order.  Features
// are counted independently (a read overlapping k features is evaluated k times), exactly as
// the reference does.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <iterator>
#include <map>

#include "compact.cuh"

namespace {

struct FeatIndexHost {
    int32_t n_gid = 0;
    std::vector<int32_t> sf_goff, sf_beg, sf_end, sf_row;
    std::vector<int32_t> bnd_goff, bnd, stab_off, stab;
    std::vector<int32_t> fb;   // per boundary: first sorted feature with beg >= bnd[k] (+ terminator)
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
    // features beginning exactly at boundary k are the sorted features [fb[k], fb[k+1])
    ix.fb.resize(ix.bnd.size() + 1);
    for (int32_t g = 0; g < n_gid; g++) {
        int32_t j = ix.sf_goff[g];
        for (int32_t k = ix.bnd_goff[g]; k < ix.bnd_goff[(size_t)g + 1]; k++) {
            while (j < ix.sf_goff[(size_t)g + 1] && ix.sf_beg[(size_t)j] < ix.bnd[(size_t)k]) j++;
            ix.fb[(size_t)k] = j;
        }
    }
    ix.fb[ix.bnd.size()] = (int32_t)m;
    return XG_OK;
}

// Per (feature, run of its contig): the tiles that can hold an overlapping read -- tiles with
// prefix-max(end) > beg and first_pos < end (two binary searches over the tile index) --
// tightened to records: from the first record of the first tile whose end > beg to the first
// record of the last tile whose pos >= end.  One warp per (feature, run).  Outputs per feature:
// candidate-read count (capacity of its set), first / last tile (its lifetime).
struct WinJob {
    int32_t j0, j1;          // sorted features [j0, j1) of the run's contig
    int32_t t0, t1;          // tiles [t0, t1) of the run
    int64_t warp0;           // first warp of the job (prefix of j1 - j0)
};
__global__ void __launch_bounds__(256) k_feature_windows(const WinJob *jobs, int32_t n_jobs, int64_t n_warps,
                                                         const xg_tile *tiles, const int32_t *pmax,
                                                         const int2 *pos_end, int32_t refine,
                                                         const int32_t *sf_beg, const int32_t *sf_end,
                                                         unsigned long long *cand, int32_t *tlo, int32_t *thi) {
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= n_warps) return;
    int a = 0, b = n_jobs;                  // last job with warp0 <= w
    while (b - a > 1) {
        int mid = (a + b) >> 1;
        if (jobs[mid].warp0 <= w) a = mid; else b = mid;
    }
    const WinJob job = jobs[a];
    const int32_t j = job.j0 + (int32_t)(w - job.warp0);
    const int32_t beg = sf_beg[j], end = sf_end[j];
    int32_t lo = job.t0, hi = job.t1;
    while (lo < hi) {                       // first tile with pmax > beg
        int32_t mid = (lo + hi) >> 1;
        if (pmax[mid] > beg) hi = mid; else lo = mid + 1;
    }
    const int32_t t_lo = lo;
    hi = job.t1;
    while (lo < hi) {                       // first tile with first_pos >= end
        int32_t mid = (lo + hi) >> 1;
        if (tiles[mid].first_pos >= end) hi = mid; else lo = mid + 1;
    }
    const int32_t t_hi = lo;
    if (t_hi <= t_lo) return;
    const xg_tile L = tiles[t_lo], H = tiles[t_hi - 1];
    int64_t first = L.rec_beg, last = H.rec_beg + H.n_rec;
    if (refine) {
        first = L.rec_beg + L.n_rec;
        for (int base = 0; base < L.n_rec; base += 32) {
            int k = base + lane;
            unsigned msk = __ballot_sync(0xffffffffu, k < L.n_rec && pos_end[L.rec_beg + k].y > beg);
            if (msk) {
                first = L.rec_beg + base + (__ffs(msk) - 1);
                break;
            }
        }
        for (int base = 0; base < H.n_rec; base += 32) {
            int k = base + lane;
            unsigned msk = __ballot_sync(0xffffffffu, k < H.n_rec && pos_end[H.rec_beg + k].x >= end);
            if (msk) {
                last = H.rec_beg + base + (__ffs(msk) - 1);
                break;
            }
        }
    }
    if (lane == 0) {
        if (last > first) atomicAdd(&cand[j], (unsigned long long)(last - first));
        atomicMin(&tlo[j], t_lo);
        atomicMax(&thi[j], t_hi);
    }
}

struct FeatCache {
    bool valid = false;
    uint64_t hash = 0;
    FeatIndexHost ix;
};

// ---- epoch plan ------------------------------------------------------------------------
// The read stream is cut into epochs of `epoch_tiles` tiles.  A feature owns a block of the
// pool (its (cell, UMI) set) from the first epoch that can hold one of its reads to the last;
// the block is zeroed just before the first, reduced to the row's non-zeros just after the
// last, and then reused by later features.  The pool therefore stays about as large as the
// state of the features under the current genomic window, so that it can live in the 126 MB
// L2 instead of streaming through HBM.
static inline uint64_t plan_blk_bytes(uint32_t cap, uint32_t log_cap) {
    return (uint64_t)cap * 16 + 16 + ((((uint64_t)log_cap * 4) + 15) & ~15ull);
}

struct EpochPlan {
    int32_t n_epochs = 0, epoch_tiles = 0;
    std::vector<uint64_t> blk_off;      // per sorted feature: byte offset of its set
    std::vector<uint32_t> tbl_cap;      // slots of its set (0 = feature never active)
    std::vector<uint32_t> log_cap;      // entries of its new-element log (= candidate reads)
    uint64_t pool_bytes = 0;
    std::vector<int32_t> zero_ptr, fin_ptr;           // per epoch ranges
    std::vector<uint64_t> zseg_off, zseg_pre;
    std::vector<int32_t> fin_feat;
    int64_t staging_cap = 0;
};

int make_plan(xg_ctx *ctx, const std::vector<unsigned long long> &cand, const std::vector<int32_t> &tlo,
              const std::vector<int32_t> &thi, int32_t n_tiles, int32_t n_cols, int32_t epoch_tiles,
              EpochPlan &pl) {
    size_t m = cand.size();
    pl.epoch_tiles = epoch_tiles;
    pl.n_epochs = std::max(1, (n_tiles + epoch_tiles - 1) / epoch_tiles);
    pl.blk_off.assign(m, 0);
    pl.tbl_cap.assign(m, 0);
    pl.log_cap.assign(m, 0);
    std::vector<std::vector<int32_t>> starts((size_t)pl.n_epochs), ends((size_t)pl.n_epochs);
    for (size_t j = 0; j < m; j++) {
        if (cand[j] == 0) continue;
        unsigned long long cap = cand[j] + cand[j] / 4 + 8;
        if (cap >= (1ull << 32)) return ctx->fail(XG_E_LIMIT, "feature window exceeds 2^32 reads");
        pl.tbl_cap[j] = (uint32_t)cap;
        pl.log_cap[j] = (uint32_t)cand[j];
        pl.staging_cap += (int64_t)std::min<unsigned long long>(cand[j], (unsigned long long)n_cols);
        starts[(size_t)(tlo[j] / epoch_tiles)].push_back((int32_t)j);
        ends[(size_t)((thi[j] - 1) / epoch_tiles)].push_back((int32_t)j);
    }
    std::map<uint64_t, uint64_t> free_blocks;   // first-fit free list keyed by offset
    auto release = [&](uint64_t off, uint64_t len) {
        auto it = free_blocks.emplace(off, len).first;
        auto nx = std::next(it);
        if (nx != free_blocks.end() && it->first + it->second == nx->first) {
            it->second += nx->second;
            free_blocks.erase(nx);
        }
        if (it != free_blocks.begin()) {
            auto pv = std::prev(it);
            if (pv->first + pv->second == it->first) {
                pv->second += it->second;
                free_blocks.erase(it);
            }
        }
    };
    pl.zero_ptr.assign((size_t)pl.n_epochs + 1, 0);
    pl.fin_ptr.assign((size_t)pl.n_epochs + 1, 0);
    for (int32_t e = 0; e < pl.n_epochs; e++) {
        // a block is reused two epochs after its feature ended, so that zeroing the blocks of
        // epoch e never races with the (overlapped) counting of epoch e-1
        if (e > 1)
            for (int32_t j : ends[(size_t)e - 2])
                release(pl.blk_off[(size_t)j], plan_blk_bytes(pl.tbl_cap[(size_t)j], pl.log_cap[(size_t)j]));
        uint64_t pre = 0;
        for (int32_t j : starts[(size_t)e]) {
            uint64_t need = plan_blk_bytes(pl.tbl_cap[(size_t)j], pl.log_cap[(size_t)j]), off = UINT64_MAX;
            const uint64_t zero_len = (uint64_t)pl.tbl_cap[(size_t)j] * 16 + 16;   // set + cursor
            for (auto it = free_blocks.begin(); it != free_blocks.end(); ++it)
                if (it->second >= need) {
                    off = it->first;
                    uint64_t rest = it->second - need;
                    free_blocks.erase(it);
                    if (rest) free_blocks.emplace(off + need, rest);
                    break;
                }
            if (off == UINT64_MAX) {
                off = pl.pool_bytes;
                if (!free_blocks.empty()) {      // extend a free block that touches the pool end
                    auto last = std::prev(free_blocks.end());
                    if (last->first + last->second == pl.pool_bytes) {
                        off = last->first;
                        free_blocks.erase(last);
                    }
                }
                pl.pool_bytes = off + need;
            }
            pl.blk_off[(size_t)j] = off;
            pl.zseg_off.push_back(off);
            pl.zseg_pre.push_back(pre);
            pre += zero_len;
        }
        pl.zseg_pre.push_back(pre);      // terminator of the epoch: total bytes
        pl.zseg_off.push_back(0);
        pl.zero_ptr[(size_t)e + 1] = (int32_t)pl.zseg_off.size();
        for (int32_t j : ends[(size_t)e]) pl.fin_feat.push_back(j);
        pl.fin_ptr[(size_t)e + 1] = (int32_t)pl.fin_feat.size();
    }
    return XG_OK;
}

#define SB_MAX 512       // boundaries staged in shared memory per tile
#define PAIR_CAP 1024    // (feature, cell, UMI) triples staged per tile
#define RPT 4            // records per thread (XG_TILE / 256)
#define PB 2             // staged pairs a thread keeps in flight in the insert phase

// per sorted feature: where its block lives.  Block layout:
//   [ set: cap x 16 B ][ cursor: 16 B ][ log: log_cap x 4 B ]
// The log receives the cell of every NEW (cell, UMI) element, so that the row can be reduced
// from n_new x 4 B instead of a scan of the whole (mostly empty or duplicate-free) set.
struct __align__(16) FeatDesc {
    unsigned long long blk_off;
    uint32_t cap, log_cap;
};
__host__ __device__ __forceinline__ uint64_t blk_zero_bytes(uint32_t cap) { return (uint64_t)cap * 16 + 16; }
__host__ __device__ __forceinline__ uint64_t blk_bytes(uint32_t cap, uint32_t log_cap) {
    return blk_zero_bytes(cap) + ((((uint64_t)log_cap * 4) + 15) & ~15ull);
}

struct BasefcDev {
    const int2 *pos_end;
    const uint32_t *fmq, *cig_off, *cigar;
    const ulonglong2 *keys;
    const xg_run *runs;
    const xg_tile *tiles;
    const int2 *tile_bnd;         // per tile: boundaries [x, y) under its window
    const int32_t *sf_goff, *sf_end, *bnd_goff, *bnd, *stab_off, *fb;
    const int4 *stab4;            // stabbing lists: {sorted feature, beg, end, 0}
    int32_t n_gid;
    uint8_t *pool;
    const FeatDesc *fdesc;
    int32_t tile0;                // first tile of this launch (epoch)
    int32_t ablate;               // profiling only: 1 skip inserts, 2 stop after cell lookup, 3 after loads
    BarcodeTable bc;
    FilterParams fp;
    const int32_t *incl_tab;
    int32_t incl_tab_len, incl_len;
};

// per tile: the range of boundaries that its window [first_pos, max_end) can touch
__global__ void k_tile_bounds(const xg_tile *tiles, const xg_run *runs, int32_t n_tiles, int32_t n_gid,
                              const int32_t *bnd_goff, const int32_t *bnd, const int32_t *stab_off,
                              const int32_t *fb, int2 *out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const xg_tile tl = tiles[t];
    const int32_t gid = runs[tl.run].gid;
    if (gid < 0 || gid >= n_gid || bnd_goff[gid] == bnd_goff[gid + 1]) {
        out[t] = make_int2(-1, -1);
        return;
    }
    const int32_t b0 = bnd_goff[gid], b1 = bnd_goff[gid + 1];
    int32_t lo = b0, hi = b1;
    while (lo < hi) {              // upper_bound(first_pos)
        int32_t mid = (lo + hi) >> 1;
        if (bnd[mid] <= tl.first_pos) lo = mid + 1; else hi = mid;
    }
    int32_t x = max(b0, lo - 1);
    hi = b1;
    while (lo < hi) {              // lower_bound(max_end)
        int32_t mid = (lo + hi) >> 1;
        if (bnd[mid] < tl.max_end) lo = mid + 1; else hi = mid;
    }
    const int32_t y = max(x, lo);
    // features that a record of this tile can overlap: those stabbing a segment in [x, y] or
    // beginning at a boundary in (x, y).  None (e.g. another GPU's genomic chunk): skip the tile.
    const int32_t y_seg = min(y + 1, b1);
    const bool any = stab_off[y_seg] > stab_off[x] || fb[y] > fb[min(x + 1, y)];
    out[t] = any ? make_int2(x, y) : make_int2(-1, -1);
}

__device__ __forceinline__ uint32_t set_home(uint64_t umi, uint32_t col, uint32_t cap) {
    return hash_to_range(mix64(umi ^ ((uint64_t)col * 0x9E3779B97F4A7C15ULL)), cap);
}

// (cell, UMI) -> the feature's set, starting at slot s whose content `cur` was already loaded.
// Empty slot: b == 0.
// Returns true when the element is new.
__device__ __forceinline__ bool set_insert_from(xg_e128 *tbl, uint32_t cap, uint32_t s, xg_e128 cur,
                                                xg_e128 want) {
    for (uint32_t probe = 0; probe < cap; probe++) {
        if (cur.b == 0) {
            xg_e128 empty;
            empty.a = 0;
            empty.b = 0;
            cur = cas128(&tbl[s], empty, want);
            if (cur.b == 0) return true;
        }
        if (cur.a == want.a && cur.b == want.b) return false;
        s = (s + 1 == cap) ? 0 : s + 1;
        cur = ld128_relaxed(&tbl[s]);
    }
    return false;
}

// Append the cell of a new element to the feature's log.  Lanes of the warp that append to
// the same feature are grouped with match_any: one cursor atomic per group.
__device__ __forceinline__ void log_append(xg_e128 *tbl, uint32_t cap, bool is_new, uint32_t col) {
    const unsigned active = __ballot_sync(0xffffffffu, is_new);
    if (!is_new) return;
    const unsigned peers = __match_any_sync(active, (unsigned long long)tbl);
    const int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
    uint32_t *cursor = (uint32_t *)(tbl + cap);
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(cursor, (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    uint32_t *log = cursor + 4;
    log[base + __popc(peers & ((1u << lane) - 1u))] = col;
}

#define CIG_CAP 1024     // CIGAR words of the tile staged in shared memory
#define STAB_CAP 256     // stabbing-list entries of the tile's boundaries staged in shared memory

struct PairStage {
    unsigned long long umi[PAIR_CAP];
    uint32_t j[PAIR_CAP], col[PAIR_CAP];
    int4 stab4[STAB_CAP];
    uint32_t cigar[CIG_CAP];
    int32_t bnd[SB_MAX], stab_off[SB_MAX + 1];
    int n_pairs;
};

// m = number of aligned (M/=/X) reference positions p of the read with s0 <= p < e0
// (== len([x for x in read.positions if s <= x <= e]), rdr/fc/core.py:40-43)
__device__ __forceinline__ int32_t included_len(const uint32_t *cig, uint32_t n_ops, int32_t pos,
                                                 int32_t s0, int32_t e0) {
    int32_t m = 0, p = pos;
    for (uint32_t k = 0; k < n_ops; k++) {
        uint32_t w = cig[k], op = w & 15u;
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

// include test of one (read, feature) pair; a passing pair is staged for the insert phase
__device__ __forceinline__ void emit_pair(const BasefcDev &P, PairStage &S, int32_t j, int32_t s0, int32_t e0,
                                          int32_t pos, int32_t end, const uint32_t *cig, uint32_t n_ops,
                                          int32_t aln, int32_t need, uint64_t umi, uint32_t col) {
    int32_t m;
    if (n_ops == 0) {
        int32_t a = max(pos, s0), b = min(end, e0);
        m = b > a ? b - a : 0;
    } else if (s0 <= pos && end <= e0) {
        m = aln;                  // the read lies inside the feature: every aligned position counts
    } else {
        m = included_len(cig, n_ops, pos, s0, e0);
    }
    if (m < need) return;
    int slot = atomicAdd(&S.n_pairs, 1);
    if (slot < PAIR_CAP) {
        S.umi[slot] = umi;
        S.j[slot] = (uint32_t)j;
        S.col[slot] = col;
    } else {                       // stage full (very deep feature overlap): insert right away
        const FeatDesc fd = P.fdesc[j];
        if (fd.cap) {
            xg_e128 want;
            want.a = umi;
            want.b = (unsigned long long)col + 1ull;
            xg_e128 *tbl = (xg_e128 *)(P.pool + fd.blk_off);
            uint32_t s = set_home(umi, col, fd.cap);
            if (set_insert_from(tbl, fd.cap, s, ld128_relaxed(&tbl[s]), want)) {
                uint32_t *cursor = (uint32_t *)(tbl + fd.cap);
                cursor[4 + atomicAdd(cursor, 1u)] = col;
            }
        }
    }
}

// Insert phase: all lanes insert staged pairs; descriptor and home-slot loads of a batch are
// issued together before any of them is consumed.  Ends with the stage empty.
// Insert phase: all lanes insert the staged (feature, cell, UMI) triples into the features'
// sets; the descriptor and home-slot loads of a batch are issued together before any of them
// is consumed.  Ends with the stage empty.  (A tile-local dedup table in shared memory and a
// CAS-first probe were tried and measured slower: the phase is issue-bound, not L2-bound.)
__device__ __forceinline__ void flush_pairs(const BasefcDev &P, PairStage &S) {
    const int np = P.ablate == 1 ? 0 : min(S.n_pairs, PAIR_CAP);
    for (int p0 = 0; p0 < np; p0 += 256 * PB) {
        xg_e128 *tbl[PB];
        uint32_t cap[PB], home[PB];
        xg_e128 cur[PB];
#pragma unroll
        for (int r = 0; r < PB; r++) {
            const int p = p0 + r * 256 + threadIdx.x;
            cap[r] = 0;
            if (p < np) {
                const FeatDesc fd = P.fdesc[S.j[p]];
                cap[r] = fd.cap;
                tbl[r] = (xg_e128 *)(P.pool + fd.blk_off);
            }
        }
#pragma unroll
        for (int r = 0; r < PB; r++) {
            const int p = p0 + r * 256 + threadIdx.x;
            if (cap[r]) {
                home[r] = set_home(S.umi[p], S.col[p], cap[r]);
                cur[r] = ld128_relaxed(tbl[r] + home[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < PB; r++) {
            const int p = p0 + r * 256 + threadIdx.x;
            bool is_new = false;
            uint32_t col = 0;
            if (cap[r]) {
                xg_e128 want;
                col = S.col[p];
                want.a = S.umi[p];
                want.b = (unsigned long long)col + 1ull;
                is_new = set_insert_from(tbl[r], cap[r], home[r], cur[r], want);
            }
            log_append(tbl[r], cap[r], is_new, col);     // whole warp: uses match_any
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) S.n_pairs = 0;
    __syncthreads();
}

__global__ void __launch_bounds__(256, 6) k_basefc_count(const __grid_constant__ BasefcDev P) {
    __shared__ PairStage S;
    const int t = P.tile0 + blockIdx.x;
    const xg_tile tile = P.tiles[t];
    const xg_run run = P.runs[tile.run];
    const int32_t gid = run.gid;
    if (gid < 0 || gid >= P.n_gid) return;
    if (P.sf_goff[gid] == P.sf_goff[gid + 1]) return;
    const int32_t b0 = P.bnd_goff[gid], b1 = P.bnd_goff[gid + 1];
    const int2 tb = P.tile_bnd[t];
    if (tb.x < 0) return;          // no feature under this tile's window: its records are never read
    const int32_t nb = tb.y - tb.x;

    // ---- the thread's records: independent, coalesced loads issued up front
    int2 pe[RPT];
    uint32_t fq[RPT], co[RPT];
    ulonglong2 ky[RPT];
#pragma unroll
    for (int r = 0; r < RPT; r++) {
        const int32_t k = threadIdx.x + r * 256;
        const bool live = k < tile.n_rec;
        const int64_t i = tile.rec_beg + (live ? k : 0);
        pe[r] = __ldcs(&P.pos_end[i]);         // streamed once: evict-first, keep L2 for the pool
        fq[r] = __ldcs(&P.fmq[i]);
        co[r] = __ldcs(&P.cig_off[i]);
        ky[r] = __ldcs(&P.keys[i]);
        if (!live) ky[r].y = XG_KEY_NONE;          // an absent UMI drops the record
    }
    // ---- stage the slice of the interval index under the tile window and the tile's CIGAR words
    const uint32_t c_first = __ldg(&P.cig_off[tile.rec_beg]);
    const uint32_t c_lo = c_first ? c_first - 1 : 0;      // one word back: a >=255-op count word
    const uint32_t c_hi = __ldg(&P.cig_off[tile.rec_beg + tile.n_rec]);
    const bool cig_staged = c_hi - c_lo <= CIG_CAP;
    const bool staged = nb <= SB_MAX;
    int32_t st_lo = 0, st_n = 0;
    if (staged) {
        st_lo = __ldg(&P.stab_off[tb.x]);
        st_n = __ldg(&P.stab_off[tb.y]) - st_lo;
        for (int k = threadIdx.x; k < nb; k += blockDim.x) S.bnd[k] = __ldg(&P.bnd[tb.x + k]);
        for (int k = threadIdx.x; k <= nb; k += blockDim.x) S.stab_off[k] = __ldg(&P.stab_off[tb.x + k]);
        if (st_n <= STAB_CAP)
            for (int k = threadIdx.x; k < st_n; k += blockDim.x) S.stab4[k] = __ldg(&P.stab4[st_lo + k]);
    }
    const bool stab_staged = staged && st_n <= STAB_CAP;
    if (cig_staged)
        for (uint32_t k = threadIdx.x; k < c_hi - c_lo; k += blockDim.x) S.cigar[k] = __ldg(&P.cigar[c_lo + k]);
    if (threadIdx.x == 0) S.n_pairs = 0;

    // ---- cell lookup: the home slots of the records are probed together
    int32_t colv[RPT];
    if (P.fp.use_cell_tag) {
        ulonglong2 slot[RPT];
        uint32_t hs[RPT];
#pragma unroll
        for (int r = 0; r < RPT; r++) {
            hs[r] = (uint32_t)mix64(ky[r].x) & P.bc.mask;
            slot[r] = __ldg(&P.bc.slots[hs[r]]);
        }
#pragma unroll
        for (int r = 0; r < RPT; r++) {
            int32_t c = -1;
            if (ky[r].x != XG_KEY_NONE) {
                ulonglong2 e = slot[r];
                uint32_t s = hs[r];
                while (true) {
                    if (e.x == ky[r].x) {
                        c = (int32_t)e.y;
                        break;
                    }
                    if (e.x == XG_KEY_NONE) break;
                    s = (s + 1) & P.bc.mask;
                    e = __ldg(&P.bc.slots[s]);
                }
            }
            colv[r] = c;
        }
    } else {
#pragma unroll
        for (int r = 0; r < RPT; r++) colv[r] = run.bam_idx;
    }
    __syncthreads();

    // ---- phase 1: filters, overlapping features, include test; passing pairs are staged
#pragma unroll
    for (int r = 0; r < RPT; r++) {
        do {
        const uint32_t fmq = fq[r];
        const uint64_t umi = ky[r].y;
        if (P.ablate == 3) {
            if (umi == 12345 && fmq == 77 && pe[r].x == 3) S.n_pairs = 1;
            break;
        }
        if (umi == XG_KEY_NONE || umi == XG_KEY_EMPTY) break;   // has_tag / `if umi:`
        if (!read_passes_flags(P.fp, fmq)) break;
        if (colv[r] < 0) break;                                   // cell tag absent or not listed
        const uint32_t col = (uint32_t)colv[r];
        if (P.ablate == 2) {
            if (col == 0x7fffffff) S.n_pairs = 1;
            break;
        }
        const int32_t pos = pe[r].x, end = pe[r].y;
        uint32_t n_ops = fmq >> 24;               // aligned length = len(read.positions)
        const uint32_t *cig = nullptr;
        int32_t aln;
        if (n_ops == 0) {
            aln = end - pos;
        } else {
            cig = cig_staged ? &S.cigar[co[r] - c_lo] : P.cigar + co[r];
            if (n_ops == 255) n_ops = cig[-1];
            aln = 0;
            for (uint32_t q = 0; q < n_ops; q++) {
                uint32_t w = cig[q];
                if (cig_aligned(w & 15u)) aln += (int32_t)(w >> 4);
            }
        }
        if (aln < P.fp.min_len) break;
        const int32_t need = P.incl_tab ? __ldg(&P.incl_tab[min(aln, P.incl_tab_len - 1)]) : P.incl_len;

        // first boundary > pos (global index)
        int32_t ub;
        if (staged) {
            int32_t lo = 0;
            if (nb <= 8) {                     // few boundaries under the tile: branch-free count
                for (int k = 0; k < nb; k++) lo += S.bnd[k] <= pos;
            } else {
                int32_t hi = nb;
                while (lo < hi) {
                    int32_t mid = (lo + hi) >> 1;
                    if (S.bnd[mid] <= pos) lo = mid + 1; else hi = mid;
                }
            }
            ub = tb.x + lo;
        } else {
            int32_t lo = b0, hi = b1;
            while (lo < hi) {
                int32_t mid = (lo + hi) >> 1;
                if (__ldg(&P.bnd[mid]) <= pos) lo = mid + 1; else hi = mid;
            }
            ub = lo;
        }
        // (1) features covering `pos`: stabbing list of the segment [bnd[ub-1], bnd[ub])
        if (ub > b0) {
            int32_t s0i, s1i;
            if (staged && ub > tb.x) {
                s0i = S.stab_off[ub - 1 - tb.x];
                s1i = S.stab_off[ub - tb.x];
            } else {
                s0i = __ldg(&P.stab_off[ub - 1]);
                s1i = __ldg(&P.stab_off[ub]);
            }
            for (int32_t s = s0i; s < s1i; s++) {
                const int4 f = (stab_staged && s >= st_lo) ? S.stab4[s - st_lo] : __ldg(&P.stab4[s]);
                emit_pair(P, S, f.x, f.y, f.z, pos, end, cig, n_ops, aln, need, umi, col);
            }
        }
        // (2) features beginning at a boundary inside (pos, end); every boundary from tb.y on is
        // >= the tile's max end, so a staged tile never looks past its staged range
        const int32_t kb_end = staged ? tb.y : b1;
        for (int32_t kb = ub; kb < kb_end; kb++) {
            const int32_t bv = staged ? S.bnd[kb - tb.x] : __ldg(&P.bnd[kb]);
            if (bv >= end) break;
            const int32_t j1 = __ldg(&P.fb[kb + 1]);
            for (int32_t j = __ldg(&P.fb[kb]); j < j1; j++)
                emit_pair(P, S, j, bv, __ldg(&P.sf_end[j]), pos, end, cig, n_ops, aln, need, umi, col);
        }
        } while (false);
        // deep feature overlap: drain the stage between rounds so that the next 256 records
        // find room (the direct-insert path of emit_pair stays the last resort)
        if (r + 1 < RPT) {
            __syncthreads();
            if (S.n_pairs > PAIR_CAP / 2) flush_pairs(P, S);
        }
    }
    __syncthreads();

    flush_pairs(P, S);
}

// "narrow" result entries: column | count << 16; a count that does not fit goes to the side list
__global__ void k_pack_rows(const int32_t *col, const int32_t *val, uint32_t *packed, long long from, long long to,
                            long long *over_idx, int32_t *over_val, unsigned int *over_n, unsigned int over_cap) {
    const long long i = from + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= to) return;
    const uint32_t c = (uint32_t)col[i], v = (uint32_t)val[i];
    uint32_t v16 = v;
    if (v >= 0xffffu) {
        v16 = 0xffffu;
        const unsigned int k = atomicAdd(over_n, 1u);
        if (k < over_cap) {
            over_idx[k] = i;
            over_val[k] = (int32_t)v;
        }
    }
    packed[i] = c | (v16 << 16);
}

__global__ void k_snapshot_cursor(const unsigned long long *cursor, unsigned long long *host_slot) {
    *host_slot = *cursor;
    __threadfence_system();
}

// Zero the sets of the features that become active in this epoch.  The segments are laid
// end to end in a virtual byte space (pre[] = exclusive prefix, pre[n_seg] = total).
#define ZERO_CHUNK 16384
__global__ void __launch_bounds__(256) k_zero_segments(uint8_t *pool, const uint64_t *off,
                                                       const uint64_t *pre, int32_t n_seg) {
    const uint64_t total = pre[n_seg];
    const uint64_t c0 = (uint64_t)blockIdx.x * ZERO_CHUNK;
    int32_t lo = 0, hi = n_seg;          // last segment with pre <= c0
    while (hi - lo > 1) {
        int32_t mid = (lo + hi) >> 1;
        if (pre[mid] <= c0) lo = mid; else hi = mid;
    }
    int32_t s = lo;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (uint64_t p = c0 + (uint64_t)threadIdx.x * 16; p < c0 + ZERO_CHUNK && p < total;
         p += (uint64_t)blockDim.x * 16) {
        while (p >= pre[s + 1]) s++;
        *(uint4 *)(pool + off[s] + (p - pre[s])) = z;
    }
}

// Reduce a feature whose last epoch just finished to its row of the matrix: histogram of its
// new-element log (one cell index per distinct (cell, UMI)) in shared memory plus a bitmap of
// the touched cells; the set bits enumerated in order give the non-zeros in column order (the
// reference's emit loop, rdr/fc/core.py:109-117).  The row goes to a staging area at an
// atomically reserved offset; k_gather_rows puts the rows in input order.  Persistent CTAs
// take features from a work counter; histogram and bitmap are cleared while they are read, so
// the next feature starts clean.  When the cells do not fit the histogram (n_cols > hist_cols)
// the log is read once per column range, first to count, then to write.
__global__ void __launch_bounds__(256) k_basefc_finalize(const uint8_t *pool, const FeatDesc *fdesc,
                                                         const int32_t *sf_row, const int32_t *fin_feat,
                                                         int32_t n_fin, int32_t n_cols, int32_t hist_cols,
                                                         unsigned int *work, unsigned long long *cursor,
                                                         int64_t *seg_base, int32_t *seg_nnz, int32_t *st_col,
                                                         int32_t *st_val) {
    extern __shared__ uint32_t smem[];
    uint32_t *hist = smem;                          // hist_cols
    uint32_t *bitmap = smem + hist_cols;            // (hist_cols + 31) / 32
    __shared__ int warp_tot[8];
    __shared__ long long base_s;
    __shared__ int f_s;
    const int n_words_max = (hist_cols + 31) >> 5;
    for (int c = threadIdx.x; c < hist_cols; c += blockDim.x) hist[c] = 0;
    for (int c = threadIdx.x; c < n_words_max; c += blockDim.x) bitmap[c] = 0;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n_pass = (n_cols + hist_cols - 1) / hist_cols;

    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) f_s = (int)atomicAdd(work, 1u);
        __syncthreads();
        const int f = f_s;
        if (f >= n_fin) break;
        const int32_t j = fin_feat[f];
        const FeatDesc fd = fdesc[j];
        const uint32_t *cur_p = (const uint32_t *)(pool + fd.blk_off + (size_t)fd.cap * 16);
        const uint32_t n_new = cur_p[0];
        const uint32_t *log = cur_p + 4;
        long long base = 0;
        // stage 0 (only when n_pass > 1): count; stage 1: write
        for (int stage = (n_pass > 1 ? 0 : 1); stage < 2; stage++) {
            int total_nz = 0;
            for (int pass = 0; pass < n_pass; pass++) {
                const uint32_t c_lo = (uint32_t)pass * (uint32_t)hist_cols;
                const int nc = min(hist_cols, n_cols - (int)c_lo);
                const int nw = (nc + 31) >> 5;
                for (uint32_t s = threadIdx.x; s < n_new; s += blockDim.x) {
                    const uint32_t col = log[s] - c_lo;
                    if (col < (uint32_t)nc && atomicAdd(&hist[col], 1u) == 0)
                        atomicOr(&bitmap[col >> 5], 1u << (col & 31));
                }
                __syncthreads();
                const bool writing = (stage == 1);
                if (!writing || n_pass == 1) {            // non-zero cells of this range
                    int nz = 0;
                    for (int k = threadIdx.x; k < nw; k += blockDim.x) nz += __popc(bitmap[k]);
                    for (int d = 16; d > 0; d >>= 1) nz += __shfl_xor_sync(0xffffffffu, nz, d);
                    if (lane == 0) warp_tot[w] = nz;
                    __syncthreads();
                    for (int k = 0; k < 8; k++) total_nz += warp_tot[k];
                    __syncthreads();
                }
                if (writing && n_pass == 1) {             // single range: reserve now
                    if (threadIdx.x == 0) {
                        const int32_t row = sf_row[j];
                        long long b = total_nz ? (long long)atomicAdd(cursor, (unsigned long long)total_nz) : 0;
                        seg_base[row] = b;
                        seg_nnz[row] = total_nz;
                        base_s = b;
                    }
                    __syncthreads();
                    base = base_s;
                }
                // ordered walk over the bitmap words; clears histogram and bitmap as it goes
                for (int k0 = 0; k0 < nw; k0 += blockDim.x) {
                    const int k = k0 + threadIdx.x;
                    uint32_t bits = k < nw ? bitmap[k] : 0u;
                    int cnt = __popc(bits), incl = cnt;
                    for (int d = 1; d < 32; d <<= 1) {
                        int y = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += y;
                    }
                    if (lane == 31) warp_tot[w] = incl;
                    __syncthreads();
                    int before = 0, tot = 0;
                    for (int q = 0; q < 8; q++) {
                        const int x = warp_tot[q];
                        if (q < w) before += x;
                        tot += x;
                    }
                    long long o = base + before + (incl - cnt);
                    while (bits) {
                        const int bpos = __ffs(bits) - 1;
                        bits &= bits - 1;
                        const int c = (k << 5) + bpos;
                        if (writing) {
                            st_col[o] = (int32_t)(c_lo + (uint32_t)c);
                            st_val[o] = (int32_t)hist[c];
                            o++;
                        }
                        hist[c] = 0;
                    }
                    if (k < nw) bitmap[k] = 0;
                    base += tot;
                    __syncthreads();
                }
            }
            if (stage == 0) {                             // multi-range: reserve after counting
                if (threadIdx.x == 0) {
                    const int32_t row = sf_row[j];
                    long long b = total_nz ? (long long)atomicAdd(cursor, (unsigned long long)total_nz) : 0;
                    seg_base[row] = b;
                    seg_nnz[row] = total_nz;
                    base_s = b;
                }
                __syncthreads();
                base = base_s;
            }
        }
    }
}

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

// `src` != nullptr: the records of `rd` are not in HBM yet -- its device arrays are allocated
// but empty, and every epoch's slice is copied from the pinned host batch `src` on a copy
// stream just ahead of the epoch that counts it (H2D overlaps the kernels).
static int basefc_run(xg_ctx *ctx, const xg_dreads *rd, const xg_reads *src, const xg_features *feats,
                      const xg_barcodes *cells, const xg_params *par, xg_coo **out) {
    if (!ctx || !ctx->stream) return ctx ? ctx->fail(XG_E_CUDA, "context has no device") : XG_E_ARG;
    if (!rd || !feats || !cells || !par || !out) return ctx->fail(XG_E_ARG, "xg_basefc: null argument");
    if (feats->n < 0 || cells->n_samples <= 0) return ctx->fail(XG_E_ARG, "xg_basefc: empty sample list");
    if (par->use_cell_tag && cells->n != cells->n_samples)
        return ctx->fail(XG_E_ARG, "xg_basefc: barcode mode needs one key per column");
    if (rd->mapped) return ctx->fail(XG_E_ARG, "xg_basefc: needs an uploaded batch (xg_upload_reads), not xg_map_reads");
    XG_CUDA(cudaSetDevice(ctx->device));
    for (double &t : ctx->timing) t = 0;
    int launches = 0;
    const int32_t n_rows = feats->n, n_cols = cells->n_samples;

    int32_t n_gid = 0;
    for (auto &r : rd->h_runs) n_gid = std::max(n_gid, r.gid + 1);
    if (!par->use_cell_tag)
        for (auto &r : rd->h_runs)
            if (r.bam_idx >= n_cols) return ctx->fail(XG_E_ARG, "more BAMs than sample columns");

    // ---- interval index and per-feature tile windows (host), exact candidate counts (device)
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [](std::chrono::steady_clock::time_point a) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count();
    };
    const auto t_call = now();
    auto t_ph = now();
    // the interval index depends on the features only: reuse it (host copy and the uploaded
    // device arrays, which live in named scratch buffers) while the caller passes the same set
    uint64_t fh = 1469598103934665603ull;
    auto mixh = [&](const void *p, size_t n) {
        const uint8_t *b = (const uint8_t *)p;
        size_t k = 0;
        for (; k + 8 <= n; k += 8) {               // word-wise FNV-style mix
            uint64_t wv;
            memcpy(&wv, b + k, 8);
            fh = (fh ^ wv) * 1099511628211ull;
            fh ^= fh >> 29;
        }
        for (; k < n; k++) fh = (fh ^ b[k]) * 1099511628211ull;
    };
    mixh(&n_gid, sizeof n_gid);
    mixh(&feats->n, sizeof feats->n);
    mixh(feats->gid, sizeof(int32_t) * (size_t)feats->n);
    mixh(feats->beg, sizeof(int32_t) * (size_t)feats->n);
    mixh(feats->end, sizeof(int32_t) * (size_t)feats->n);
    FeatCache *fc = (FeatCache *)ctx->fx_cache;
    if (!fc) {
        fc = new FeatCache();
        ctx->fx_cache = fc;
        ctx->fx_cache_free = [](void *p) { delete (FeatCache *)p; };
    }
    const bool index_cached = fc->valid && fc->hash == fh;
    int rc = XG_OK;
    if (!index_cached) {
        fc->valid = false;
        fc->ix = FeatIndexHost();
        if ((rc = build_feat_index(ctx, feats, n_gid, fc->ix))) return rc;
        fc->hash = fh;
    }
    const FeatIndexHost &ix = fc->ix;
    const size_t m = ix.sf_beg.size();
    const double ms_index = ms_since(t_ph);
    t_ph = now();
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
    const int32_t *d_sf_row = nullptr;
    const int32_t *d_sf_beg = nullptr;
    if (!index_cached) {
        std::vector<int4> stab4(ix.stab.size());
        for (size_t k = 0; k < ix.stab.size(); k++) {
            const int32_t j = ix.stab[k];
            stab4[k] = make_int4(j, ix.sf_beg[(size_t)j], ix.sf_end[(size_t)j], 0);
        }
        const int4 *d_stab4 = nullptr;
        const int32_t *d = nullptr;
        if ((rc = upload_vec(ctx, ix.sf_goff, "fx_sf_goff", &d))) return rc;
        if ((rc = upload_vec(ctx, ix.sf_beg, "fx_sf_beg", &d))) return rc;
        if ((rc = upload_vec(ctx, ix.sf_end, "fx_sf_end", &d))) return rc;
        if ((rc = upload_vec(ctx, ix.sf_row, "fx_sf_row", &d))) return rc;
        if ((rc = upload_vec(ctx, ix.bnd_goff, "fx_bnd_goff", &d))) return rc;
        if ((rc = upload_vec(ctx, ix.bnd, "fx_bnd", &d))) return rc;
        if ((rc = upload_vec(ctx, ix.stab_off, "fx_stab_off", &d))) return rc;
        if ((rc = upload_vec(ctx, stab4, "fx_stab4", &d_stab4))) return rc;
        if ((rc = upload_vec(ctx, ix.fb, "fx_fb", &d))) return rc;
        XG_CUDA(cudaStreamSynchronize(ctx->stream));
        fc->valid = true;
    }
    // the named scratch buffers keep their addresses between calls
    P.sf_goff = (const int32_t *)ctx->scratch["fx_sf_goff"].p;
    d_sf_beg = (const int32_t *)ctx->scratch["fx_sf_beg"].p;
    P.sf_end = (const int32_t *)ctx->scratch["fx_sf_end"].p;
    d_sf_row = (const int32_t *)ctx->scratch["fx_sf_row"].p;
    P.bnd_goff = (const int32_t *)ctx->scratch["fx_bnd_goff"].p;
    P.bnd = (const int32_t *)ctx->scratch["fx_bnd"].p;
    P.stab_off = (const int32_t *)ctx->scratch["fx_stab_off"].p;
    P.stab4 = (const int4 *)ctx->scratch["fx_stab4"].p;
    P.fb = (const int32_t *)ctx->scratch["fx_fb"].p;
    // feature windows on the device: one job per run whose contig has features
    std::vector<WinJob> jobs;
    int64_t n_warps = 0;
    {
        size_t t = 0;
        const size_t nt = rd->h_tiles.size();
        while (t < nt) {
            const int32_t r = rd->h_tiles[t].run;
            size_t e = t;
            while (e < nt && rd->h_tiles[e].run == r) e++;
            const int32_t g = rd->h_runs[(size_t)r].gid;
            if (g >= 0 && g < n_gid && ix.sf_goff[(size_t)g + 1] > ix.sf_goff[(size_t)g]) {
                jobs.push_back(WinJob{ix.sf_goff[(size_t)g], ix.sf_goff[(size_t)g + 1], (int32_t)t, (int32_t)e, n_warps});
                n_warps += ix.sf_goff[(size_t)g + 1] - ix.sf_goff[(size_t)g];
            }
            t = e;
        }
    }
    const WinJob *d_jobs = nullptr;
    if ((rc = upload_vec(ctx, jobs, "fx_jobs", &d_jobs))) return rc;
    XG_GET(d_cand, unsigned long long, "fx_cand", m + 1);
    XG_GET(d_tlo, int32_t, "fx_tlo", m + 1);
    XG_GET(d_thi, int32_t, "fx_thi", m + 1);
    XG_GET(d_tile_bnd, int2, "fx_tile_bnd", rd->n_tiles + 1);
    P.tile_bnd = d_tile_bnd;
    cudaEventRecord(ctx->ev[0], ctx->stream);
    XG_CUDA(cudaMemsetAsync(d_cand, 0, sizeof(unsigned long long) * (m + 1), ctx->stream));
    XG_CUDA(cudaMemsetAsync(d_tlo, 0x7f, sizeof(int32_t) * (m + 1), ctx->stream));
    XG_CUDA(cudaMemsetAsync(d_thi, 0xff, sizeof(int32_t) * (m + 1), ctx->stream));
    std::vector<unsigned long long> cand(m, 0);
    std::vector<int32_t> tlo(m, INT32_MAX), thi(m, -1);
    if (n_warps > 0) {
        k_feature_windows<<<(unsigned)((n_warps + 7) / 8), 256, 0, ctx->stream>>>(
            d_jobs, (int32_t)jobs.size(), n_warps, rd->tiles, rd->tile_pmax, rd->pos_end, src ? 0 : 1, d_sf_beg,
            P.sf_end, d_cand, d_tlo, d_thi);
        launches++;
    }
    if (rd->n_tiles > 0) {
        k_tile_bounds<<<(rd->n_tiles + 255) / 256, 256, 0, ctx->stream>>>(rd->tiles, rd->runs, rd->n_tiles, n_gid,
                                                                       P.bnd_goff, P.bnd, P.stab_off, P.fb,
                                                                       d_tile_bnd);
        launches++;
    }
    if (m) {
        XG_CUDA(cudaMemcpyAsync(cand.data(), d_cand, sizeof(unsigned long long) * m, cudaMemcpyDeviceToHost, ctx->stream));
        XG_CUDA(cudaMemcpyAsync(tlo.data(), d_tlo, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, ctx->stream));
        XG_CUDA(cudaMemcpyAsync(thi.data(), d_thi, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, ctx->stream));
    }
    XG_CUDA(cudaStreamSynchronize(ctx->stream));
    const double ms_windows = ms_since(t_ph);

    // ---- pool layout over epochs
    int32_t epoch_tiles = src ? 8192 : 65536;     // streaming: finer epochs = finer H2D / kernel overlap
    if (const char *e = getenv(src ? "XG_EPOCH_TILES_HOST" : "XG_EPOCH_TILES")) epoch_tiles = std::max(1, atoi(e));
    EpochPlan pl;
    t_ph = now();
    if ((rc = make_plan(ctx, cand, tlo, thi, rd->n_tiles, n_cols, epoch_tiles, pl))) return rc;
    const double ms_plan = ms_since(t_ph);
    t_ph = now();
    const int32_t *d_fin_feat = nullptr;
    const uint64_t *d_zoff = nullptr, *d_zpre = nullptr;
    std::vector<FeatDesc> fdesc(m);
    for (size_t j = 0; j < m; j++) fdesc[j] = FeatDesc{pl.blk_off[j], pl.tbl_cap[j], pl.log_cap[j]};
    if ((rc = upload_vec(ctx, fdesc, "fx_fdesc", &P.fdesc))) return rc;
    if ((rc = upload_vec(ctx, pl.fin_feat, "fx_fin_feat", &d_fin_feat))) return rc;
    if ((rc = upload_vec(ctx, pl.zseg_off, "fx_zseg_off", &d_zoff))) return rc;
    if ((rc = upload_vec(ctx, pl.zseg_pre, "fx_zseg_pre", &d_zpre))) return rc;
    if (par->min_incl_tab) {
        if (par->min_incl_tab_len <= rd->max_aln_len)
            return ctx->fail(XG_E_ARG, "min_incl_tab shorter than max aligned length + 1");
        std::vector<int32_t> t(par->min_incl_tab, par->min_incl_tab + par->min_incl_tab_len);
        if ((rc = upload_vec(ctx, t, "fx_incl_tab", &P.incl_tab))) return rc;
        P.incl_tab_len = par->min_incl_tab_len;
        XG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    P.incl_len = par->min_incl_len;
    if (const char *e = getenv("XG_ABLATE")) P.ablate = atoi(e);
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
    XG_GET(pool, uint8_t, "fx_pool", pl.pool_bytes + 16);
    XG_GET(seg_base, int64_t, "fx_seg_base", n_rows + 1);
    XG_GET(seg_nnz, int32_t, "fx_seg_nnz", n_rows + 1);
    XG_GET(st_col, int32_t, "fx_st_col", pl.staging_cap + 1);
    XG_GET(st_val, int32_t, "fx_st_val", pl.staging_cap + 1);
    XG_GET(cursor, unsigned long long, "fx_cursor", 2);
    XG_GET(fin_work, unsigned int, "fx_fin_work", pl.n_epochs + 1);
    P.pool = pool;
    XG_CUDA(cudaStreamSynchronize(ctx->stream));   // host vectors above are about to die
    const double ms_upload = ms_since(t_ph);

    // shared-memory histogram (+ bitmap) of the finalize kernel: all cells if they fit
    const int32_t hist_cols = std::min(n_cols, 40 * 1024);
    const size_t hist_bytes = (size_t)hist_cols * 4 + (size_t)((hist_cols + 31) / 32) * 4;
    XG_CUDA(cudaFuncSetAttribute(k_basefc_finalize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_bytes));
    const int fin_ctas_per_sm = std::max(1, std::min(8, (int)(200 * 1024 / (hist_bytes + 1024))));

    // ---- device: epochs.  zero(e) -> count(e) -> finalize(e) per epoch; with overlap the three
    // kinds run on their own streams and count(e+1) fills the SMs while count(e) drains:
    //   zero(e)     waits finalize(e-2)           (its blocks were released by then)
    //   count(e)    waits zero(e)                 (alternating between two streams)
    //   finalize(e) waits count(e), count(e-1)
    bool overlap = pl.n_epochs > 1;
    if (const char *e = getenv("XG_OVERLAP")) overlap = overlap && atoi(e) != 0;
    if ((overlap || src) && !ctx->aux[0])
        for (auto &st : ctx->aux) XG_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    if (src && !ctx->copy_stream) XG_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    while ((int32_t)ctx->ev_pool.size() < 5 * pl.n_epochs + 1) {
        cudaEvent_t ev;
        XG_CUDA(cudaEventCreate(&ev));
        ctx->ev_pool.push_back(ev);
    }
    auto EV = [&](int kind, int32_t e) { return ctx->ev_pool[(size_t)5 * e + kind]; };   // 0 Z, 1 S, 2 C, 3 F, 4 H2D
    cudaEvent_t ev_init = ctx->ev_pool[(size_t)5 * pl.n_epochs];
    XG_CUDA(cudaMemsetAsync(seg_nnz, 0, sizeof(int32_t) * (size_t)(n_rows + 1), ctx->stream));
    XG_CUDA(cudaMemsetAsync(seg_base, 0, sizeof(int64_t) * (size_t)(n_rows + 1), ctx->stream));
    XG_CUDA(cudaMemsetAsync(cursor, 0, 16, ctx->stream));
    XG_CUDA(cudaMemsetAsync(fin_work, 0, sizeof(unsigned int) * (size_t)(pl.n_epochs + 1), ctx->stream));
    launches += 5;
    cudaEventRecord(ev_init, ctx->stream);
    cudaStream_t st_z = overlap ? ctx->aux[0] : ctx->stream, st_f = overlap ? ctx->aux[1] : ctx->stream;
    if (overlap) {
        cudaStreamWaitEvent(st_z, ev_init, 0);
        cudaStreamWaitEvent(st_f, ev_init, 0);
        cudaStreamWaitEvent(ctx->aux[2], ev_init, 0);
    }
    if (src) cudaStreamWaitEvent(ctx->copy_stream, ev_init, 0);
    // "row_order" 0: the rows stay in the order they were completed; after every epoch the
    // staging cursor is snapshotted so that the host can copy that epoch's rows out while the
    // later epochs are still being counted
    const bool staged_out = !ctx->row_order;
    unsigned long long *h_cur = nullptr, *d_cur = nullptr;
    if (staged_out) {
        if (!ctx->d2h_stream) XG_CUDA(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
        h_cur = (unsigned long long *)ctx->pinned_get(sizeof(unsigned long long) * (size_t)(pl.n_epochs + 1));
        if (!h_cur) return ctx->fail(XG_E_NOMEM, "out of pinned host memory");
        if (cudaHostGetDevicePointer((void **)&d_cur, h_cur, 0) != cudaSuccess) {
            cudaGetLastError();
            ctx->pinned_put(h_cur);
            return ctx->fail(XG_E_CUDA, "pinned host memory is not mapped for the device");
        }
    }
    int64_t h2d_bytes = 0;
    for (int32_t e = 0; e < pl.n_epochs; e++) {
        cudaStream_t st_c = overlap ? ((e & 1) ? ctx->aux[2] : ctx->stream) : ctx->stream;
        if (src) {      // this epoch's records: host -> HBM on the copy stream
            const int32_t ta = e * pl.epoch_tiles, tb_ = std::min(rd->n_tiles, ta + pl.epoch_tiles);
            if (tb_ > ta) {
                const int64_t ra = rd->h_tiles[(size_t)ta].rec_beg;
                const int64_t rb = rd->h_tiles[(size_t)tb_ - 1].rec_beg + rd->h_tiles[(size_t)tb_ - 1].n_rec;
                const size_t nr = (size_t)(rb - ra);
                const uint32_t ca = src->cig_off[ra], cb = src->cig_off[rb];
                cudaStream_t cs = ctx->copy_stream;
                cudaMemcpyAsync(rd->pos_end + ra, src->pos_end + 2 * ra, nr * 8, cudaMemcpyHostToDevice, cs);
                cudaMemcpyAsync(rd->fmq + ra, src->fmq + ra, nr * 4, cudaMemcpyHostToDevice, cs);
                cudaMemcpyAsync(rd->cig_off + ra, src->cig_off + ra, (nr + 1) * 4, cudaMemcpyHostToDevice, cs);
                cudaMemcpyAsync(rd->keys + ra, src->keys + 2 * ra, nr * 16, cudaMemcpyHostToDevice, cs);
                const uint32_t ca1 = ca ? ca - 1 : 0;      // one word back: a >=255-op count word
                if (cb > ca1)
                    cudaMemcpyAsync(rd->cigar + ca1, src->cigar + ca1, (size_t)(cb - ca1) * 4, cudaMemcpyHostToDevice, cs);
                h2d_bytes += (int64_t)(nr * 32 + 4 + (size_t)(cb - ca1) * 4);
            }
            cudaEventRecord(EV(4, e), ctx->copy_stream);
            cudaStreamWaitEvent(st_c, EV(4, e), 0);
        }
        const int32_t z0 = pl.zero_ptr[(size_t)e], z1 = pl.zero_ptr[(size_t)e + 1];
        const int32_t n_seg = z1 - z0 - 1;
        if (overlap && e >= 2) cudaStreamWaitEvent(st_z, EV(3, e - 2), 0);
        if (n_seg > 0) {
            const uint64_t total = pl.zseg_pre[(size_t)z1 - 1];
            k_zero_segments<<<(unsigned)((total + ZERO_CHUNK - 1) / ZERO_CHUNK), 256, 0, st_z>>>(
                pool, d_zoff + z0, d_zpre + z0, n_seg);
            launches++;
        }
        cudaEventRecord(EV(0, e), st_z);
        if (overlap) cudaStreamWaitEvent(st_c, EV(0, e), 0);
        const int32_t t0 = e * pl.epoch_tiles, t1 = std::min(rd->n_tiles, t0 + pl.epoch_tiles);
        cudaEventRecord(EV(1, e), st_c);
        if (t1 > t0 && m > 0) {
            P.tile0 = t0;
            k_basefc_count<<<t1 - t0, 256, 0, st_c>>>(P);
            launches++;
        }
        cudaEventRecord(EV(2, e), st_c);
        const int32_t n_fin = pl.fin_ptr[(size_t)e + 1] - pl.fin_ptr[(size_t)e];
        if (overlap) {
            cudaStreamWaitEvent(st_f, EV(2, e), 0);
            if (e > 0) cudaStreamWaitEvent(st_f, EV(2, e - 1), 0);
        }
        if (n_fin > 0) {
            const int grid = std::min(n_fin, 148 * fin_ctas_per_sm);
            k_basefc_finalize<<<grid, 256, hist_bytes, st_f>>>(
                pool, P.fdesc, d_sf_row, d_fin_feat + pl.fin_ptr[(size_t)e], n_fin, n_cols, hist_cols,
                fin_work + e, cursor, seg_base, seg_nnz, st_col, st_val);
            launches++;
        }
        // the snapshot is a store into mapped host memory, not a copy: a D2H of 8 bytes would queue
        // behind the result copies on the copy engine and stall the finalize stream with them
        if (h_cur) k_snapshot_cursor<<<1, 1, 0, st_f>>>(cursor, d_cur + e);
        cudaEventRecord(EV(3, e), st_f);
    }
    if (overlap) {
        cudaStreamWaitEvent(ctx->stream, EV(3, pl.n_epochs - 1), 0);
        cudaStreamWaitEvent(ctx->stream, EV(2, pl.n_epochs - 1), 0);
        if (pl.n_epochs > 1) cudaStreamWaitEvent(ctx->stream, EV(2, pl.n_epochs - 2), 0);
    }
    XG_CUDA(cudaGetLastError());
    if (staged_out) {
        const auto t_tail = std::chrono::steady_clock::now();
        xg_coo_owner *o = new xg_coo_owner();
        memset(&o->m, 0, sizeof(o->m));
        o->ctx = ctx;
        auto give_up = [&](int code, const std::string &msg) {
            cudaStreamSynchronize(ctx->d2h_stream);
            cudaStreamSynchronize(ctx->stream);
            for (void *q : o->bufs) ctx->pinned_put(q);
            ctx->pinned_put(h_cur);
            delete o;
            return ctx->fail(code, msg);
        };
        int64_t cap = ctx->fx_nnz_hint > 0 ? ctx->fx_nnz_hint + ctx->fx_nnz_hint / 8 + 1024 : 0;
        cap = std::min<int64_t>(cap, pl.staging_cap + 1);
        // "narrow" results: column and count of an entry in one 32-bit word (16 bits each; counts of
        // 65535 and more go to a side list).  Halves the bytes of the result copy, which is what a
        // host shared by several GPUs runs out of first.
        bool narrow = ctx->narrow_rows && n_cols <= 65536;
        const int64_t OVER_CAP = 1 << 20;
        uint32_t *d_packed = nullptr;
        long long *d_over_idx = nullptr;
        int32_t *d_over_val = nullptr;
        unsigned int *d_over_n = nullptr;
        if (narrow) {
            d_packed = (uint32_t *)ctx->get("fx_packed", sizeof(uint32_t) * (size_t)(pl.staging_cap + 1));
            d_over_idx = (long long *)ctx->get("fx_over_idx", sizeof(long long) * (size_t)OVER_CAP);
            d_over_val = (int32_t *)ctx->get("fx_over_val", sizeof(int32_t) * (size_t)OVER_CAP);
            d_over_n = (unsigned int *)ctx->get("fx_over_n", 16);
            if (!d_packed || !d_over_idx || !d_over_val || !d_over_n) return give_up(XG_E_CUDA, ctx->err);
            cudaMemsetAsync(d_over_n, 0, 4, ctx->d2h_stream);
        }
        int32_t *h_col = nullptr, *h_val = nullptr;
        uint32_t *h_packed = nullptr;
        auto host_buffers = [&](int64_t n_cap) {
            for (void *q : o->bufs) ctx->pinned_put(q);
            o->bufs.clear();
            h_col = h_val = nullptr;
            h_packed = nullptr;
            if (narrow) {
                h_packed = (uint32_t *)ctx->pinned_get((size_t)n_cap * 4);
                if (h_packed) o->bufs.push_back(h_packed);
                return h_packed != nullptr;
            }
            h_col = (int32_t *)ctx->pinned_get((size_t)n_cap * 4);
            h_val = (int32_t *)ctx->pinned_get((size_t)n_cap * 4);
            if (h_col) o->bufs.push_back(h_col);
            if (h_val) o->bufs.push_back(h_val);
            return h_col && h_val;
        };
        auto queue_rows = [&](int64_t from, int64_t to) {       // staging entries [from, to) -> host
            const size_t n = (size_t)(to - from);
            if (narrow) {
                k_pack_rows<<<(unsigned)((n + 255) / 256), 256, 0, ctx->d2h_stream>>>(st_col, st_val, d_packed, from, to, d_over_idx,
                                                                                    d_over_val, d_over_n, (unsigned int)OVER_CAP);
                cudaMemcpyAsync(h_packed + from, d_packed + from, n * 4, cudaMemcpyDeviceToHost, ctx->d2h_stream);
            } else {
                cudaMemcpyAsync(h_col + from, st_col + from, n * 4, cudaMemcpyDeviceToHost, ctx->d2h_stream);
                cudaMemcpyAsync(h_val + from, st_val + from, n * 4, cudaMemcpyDeviceToHost, ctx->d2h_stream);
            }
        };
        if (cap > 0 && !host_buffers(cap)) return give_up(XG_E_NOMEM, "out of pinned host memory for the result");
        int64_t done = 0;            // entries already queued for the host
        bool fits = cap > 0;
        for (int32_t e = 0; e < pl.n_epochs && fits; e++) {
            cudaError_t ce = cudaEventSynchronize(EV(3, e));
            if (ce != cudaSuccess) return give_up(XG_E_CUDA, std::string("basefc: ") + cudaGetErrorString(ce));
            const int64_t cur = (int64_t)h_cur[e];
            if (cur > cap) {
                fits = false;        // more rows than the last call: finish with one copy at the end
                break;
            }
            if (cur > done) {
                queue_rows(done, cur);
                done = cur;
            }
        }
        cudaError_t ce = cudaStreamSynchronize(ctx->stream);      // every epoch has finished
        if (ce != cudaSuccess) return give_up(XG_E_CUDA, std::string("basefc: ") + cudaGetErrorString(ce));
        unsigned long long nnz_u = 0;
        XG_CUDA(cudaMemcpy(&nnz_u, cursor, 8, cudaMemcpyDeviceToHost));
        const int64_t nnz = (int64_t)nnz_u;
        if (!fits || nnz > cap) {
            cudaStreamSynchronize(ctx->d2h_stream);
            cap = nnz + nnz / 8 + 1024;
            if (!host_buffers(cap)) return give_up(XG_E_NOMEM, "out of pinned host memory for the result");
            if (narrow) cudaMemsetAsync(d_over_n, 0, 4, ctx->d2h_stream);
            done = 0;
        }
        if (nnz > done) queue_rows(done, nnz);
        unsigned int n_over = 0;
        if (narrow) {
            cudaMemcpyAsync(&n_over, d_over_n, 4, cudaMemcpyDeviceToHost, ctx->d2h_stream);
            ce = cudaStreamSynchronize(ctx->d2h_stream);
            if (ce != cudaSuccess) return give_up(XG_E_CUDA, std::string("result D2H: ") + cudaGetErrorString(ce));
            if ((int64_t)n_over > OVER_CAP) {       // too many large counts for the side list: plain 32-bit columns
                narrow = false;
                if (!host_buffers(cap)) return give_up(XG_E_NOMEM, "out of pinned host memory for the result");
                if (nnz > 0) queue_rows(0, nnz);
                n_over = 0;
            }
        }
        long long *h_over_idx = nullptr;
        int32_t *h_over_val = nullptr;
        if (narrow) {
            h_over_idx = (long long *)ctx->pinned_get(((size_t)n_over + 1) * 8);
            h_over_val = (int32_t *)ctx->pinned_get(((size_t)n_over + 1) * 4);
            if (h_over_idx) o->bufs.push_back(h_over_idx);
            if (h_over_val) o->bufs.push_back(h_over_val);
            if (!h_over_idx || !h_over_val) return give_up(XG_E_NOMEM, "out of pinned host memory for the result");
            if (n_over) {
                cudaMemcpyAsync(h_over_idx, d_over_idx, (size_t)n_over * 8, cudaMemcpyDeviceToHost, ctx->d2h_stream);
                cudaMemcpyAsync(h_over_val, d_over_val, (size_t)n_over * 4, cudaMemcpyDeviceToHost, ctx->d2h_stream);
            }
        }
        int64_t *h_beg = (int64_t *)ctx->pinned_get((size_t)(n_rows + 1) * 8);
        int32_t *h_cnt = (int32_t *)ctx->pinned_get((size_t)(n_rows + 1) * 4);
        if (h_beg) o->bufs.push_back(h_beg);
        if (h_cnt) o->bufs.push_back(h_cnt);
        if (!h_beg || !h_cnt) return give_up(XG_E_NOMEM, "out of pinned host memory for the result");
        cudaMemcpyAsync(h_beg, seg_base, (size_t)n_rows * 8, cudaMemcpyDeviceToHost, ctx->d2h_stream);
        cudaMemcpyAsync(h_cnt, seg_nnz, (size_t)n_rows * 4, cudaMemcpyDeviceToHost, ctx->d2h_stream);
        cudaEventRecord(ctx->ev[3], ctx->stream);
        ce = cudaStreamSynchronize(ctx->d2h_stream);
        if (ce != cudaSuccess) return give_up(XG_E_CUDA, std::string("result D2H: ") + cudaGetErrorString(ce));
        cudaEventSynchronize(ctx->ev[3]);
        ctx->pinned_put(h_cur);
        ctx->fx_nnz_hint = nnz;
        ctx->timing[4] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_tail).count();
        o->m.nnz = nnz;
        o->m.n_rows = n_rows;
        o->m.n_cols = n_cols;
        o->m.col = h_col;
        o->m.val = h_val;
        o->m.colval16 = h_packed;
        o->m.n_over = narrow ? (int64_t)n_over : 0;
        o->m.over_idx = (const int64_t *)h_over_idx;
        o->m.over_val = h_over_val;
        o->m.row_beg = h_beg;
        o->m.row_cnt = h_cnt;
        *out = &o->m;
    } else if ((rc = xg_staging_to_coo(ctx, "fx", n_rows, n_cols, seg_base, seg_nnz, st_col, st_val, out, &launches))) {
        return rc;
    }

    float t_all = 0;
    const double t_d2h = ctx->timing[4];
    double t_cnt = 0;
    cudaEventElapsedTime(&t_all, ctx->ev[0], ctx->ev[3]);
    int n_cnt = 0;
    for (int32_t e = 0; e < pl.n_epochs; e++) {
        float t = 0;
        cudaEventElapsedTime(&t, EV(1, e), EV(2, e));     // per-launch duration on its own stream
        t_cnt += t;
        n_cnt++;
    }
    float t_span = 0;
    cudaEventElapsedTime(&t_span, ev_init, ctx->ev[3]);
    ctx->timing[0] = t_all;      // device time of the call: planning kernels, zero, count, finalize, gather
    ctx->timing[1] = t_cnt;      // sum over epochs of the counting kernel
    ctx->timing[2] = launches;
    ctx->timing[3] = t_span;     // epochs + scan + gather (no host planning in between)
    ctx->timing[4] = t_d2h;
    ctx->timing[5] = pl.n_epochs;
    ctx->timing[6] = (double)pl.pool_bytes;
    ctx->timing[7] = (double)pl.staging_cap;
    ctx->timing[8] = ms_index;
    ctx->timing[9] = ms_windows;
    ctx->timing[10] = ms_plan;
    ctx->timing[11] = ms_upload;
    ctx->timing[12] = ms_since(t_call);
    ctx->timing[13] = (double)h2d_bytes;
    return XG_OK;
}

extern "C" int xg_basefc(xg_ctx *ctx, const xg_dreads *rd, const xg_features *feats,
                         const xg_barcodes *cells, const xg_params *par, xg_coo **out) {
    return basefc_run(ctx, rd, nullptr, feats, cells, par, out);
}

extern "C" int xg_basefc_host(xg_ctx *ctx, const xg_reads *h, const xg_features *feats,
                              const xg_barcodes *cells, const xg_params *par, xg_coo **out) {
    if (!ctx || !ctx->stream) return ctx ? ctx->fail(XG_E_CUDA, "context has no device") : XG_E_ARG;
    if (!h) return ctx->fail(XG_E_ARG, "xg_basefc_host: null argument");
    XG_CUDA(cudaSetDevice(ctx->device));
    cudaPointerAttributes at;
    if (h->n_reads > 0 && (cudaPointerGetAttributes(&at, h->pos_end) != cudaSuccess || at.type != cudaMemoryTypeHost)) {
        cudaGetLastError();
        return ctx->fail(XG_E_ARG, "xg_basefc_host: record arrays must be pinned host memory");
    }
    xg_dreads *d = new xg_dreads();
    d->n_reads = h->n_reads;
    d->n_cigar = h->n_cigar;
    d->n_runs = h->n_runs;
    d->n_tiles = h->n_tiles;
    d->max_aln_len = h->max_aln_len;
    d->max_span = h->max_span;
    d->h_runs.assign(h->runs, h->runs + h->n_runs);
    d->h_tiles.assign(h->tiles, h->tiles + h->n_tiles);
    d->pooled = true;
    const size_t n = (size_t)h->n_reads;
    d->pos_end = (int2 *)ctx->dev_get(n * 8 + 16);
    d->fmq = (uint32_t *)ctx->dev_get(n * 4 + 16);
    d->cig_off = (uint32_t *)ctx->dev_get((n + 1) * 4 + 16);
    d->keys = (ulonglong2 *)ctx->dev_get(n * 16 + 16);
    d->cigar = (uint32_t *)ctx->dev_get((size_t)h->n_cigar * 4 + 16);
    d->runs = (xg_run *)ctx->dev_get((size_t)h->n_runs * sizeof(xg_run) + 16);
    d->tiles = (xg_tile *)ctx->dev_get((size_t)h->n_tiles * sizeof(xg_tile) + 16);
    int rc = XG_OK;
    if (!d->pos_end || !d->fmq || !d->cig_off || !d->keys || !d->cigar || !d->runs || !d->tiles) {
        rc = ctx->fail(XG_E_CUDA, "out of device memory for the read batch");
    } else {
        cudaMemcpyAsync(d->runs, h->runs, (size_t)h->n_runs * sizeof(xg_run), cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(d->tiles, h->tiles, (size_t)h->n_tiles * sizeof(xg_tile), cudaMemcpyHostToDevice, ctx->stream);
        rc = xg_make_tile_pmax(ctx, d);
        if (!rc) rc = basefc_run(ctx, d, h, feats, cells, par, out);
    }
    cudaStreamSynchronize(ctx->stream);
    xg_dreads_free(ctx, d);
    return rc;
}

// gpu_decode.cu -- BAM -> read batch entirely on the device: BGZF inflate + record parse.
//
// Same result, array for array, as the host decoder (decode.cpp: xg_decode_bams followed by
// xg_upload_reads); replaces what the reference gets from pysam/htslib on its counting paths
// (pysam.AlignmentFile + fetch(), xcltk/rdr/fc/core.py:75,100; xcltk/baf/fc/core.py:60,99).
// The compressed file crosses PCIe once (~65 B/read instead of ~250 B/read inflated), every
// BGZF block is inflated by one thread, and the records are parsed where they land.
//
// Block-parallel parsing needs every BGZF block to start at a record boundary.  htslib
// writes BAM that way (bam_write1 flushes the block before a record that does not fit, and
// the header ends with a flush), so files from samtools / cellranger / STARsolo qualify; the
// walk kernel verifies it and any other layout -- or a key that must be interned on the host
// (query names, non-ACGTN barcodes, numeric tags) -- returns XG_E_UNSUPPORTED so that the
// caller uses xg_decode_bams + xg_upload_reads instead.  Both decoders are decoders: neither
// counts anything, and the counting kernels stay device-only.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "bamfile.hpp"
#include "common.cuh"
#include "inflate.cuh"
#include "owner.hpp"

namespace {


// ---- unaligned little-endian loads (buffers are padded by 8 bytes) --------------------------
__device__ __forceinline__ uint32_t ld32u(const uint8_t *p) {
    const uintptr_t a = (uintptr_t)p;
    const uint32_t *w = (const uint32_t *)(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3) * 8;
    const uint32_t lo = w[0];
    if (sh == 0) return lo;
    return __funnelshift_r(lo, w[1], sh);
}
__device__ __forceinline__ uint32_t ld16u(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

struct BlkInfo {
    uint32_t n_all;      // records in the block
    uint32_t n_kept;     // ... on a contig the caller asked for
    uint32_t cig;        // cigar words the kept records store
    uint32_t seq;        // sequence words
    unsigned long long first_key, last_key;   // (tid, pos) of the first / last record; unplaced sort last
    uint32_t status;     // 0 ok, 1 block does not hold whole records, 2 corrupt record, 3 unsorted
    uint32_t n_starts;   // run-start entries the block will emit (first kept record + tid changes)
};

// One group of S lanes per BGZF block (inflate.cuh); INFLATE_THREADS / S blocks per CTA.
constexpr int INFLATE_THREADS = 128;
template <int S>
__global__ void __launch_bounds__(INFLATE_THREADS) k_inflate(const uint8_t *comp, const BgzfBlockDev *blocks,
                                                             int32_t n_blocks, int *n_bad) {
    extern __shared__ __align__(16) unsigned char inflate_smem[];
    xg_inflate::GroupSmem *gs = reinterpret_cast<xg_inflate::GroupSmem *>(inflate_smem);
    const int grp = threadIdx.x / S;
    const int b = blockIdx.x * (INFLATE_THREADS / S) + grp;
    if (b >= n_blocks) return;
    const BgzfBlockDev bk = blocks[b];
    if (bk.isize == 0) return;
    const int n = xg_inflate::inflate_group<S>(gs[grp], comp + bk.coff, bk.clen, bk.uptr, bk.isize);
    if ((threadIdx.x & (S - 1)) == 0 && n != (int)bk.isize) atomicAdd(n_bad, 1);
}

template <int S>
void launch_inflate_s(cudaStream_t st, const uint8_t *comp, const BgzfBlockDev *blocks, int32_t n_blocks, int *n_bad) {
    constexpr int per_cta = INFLATE_THREADS / S;
    const size_t smem = per_cta * sizeof(xg_inflate::GroupSmem);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_inflate<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    k_inflate<S><<<(unsigned)((n_blocks + per_cta - 1) / per_cta), INFLATE_THREADS, smem, st>>>(comp, blocks, n_blocks,
                                                                                              n_bad);
}

inline void launch_inflate(cudaStream_t st, const uint8_t *comp, const BgzfBlockDev *blocks, int32_t n_blocks,
                           int *n_bad) {
    if (n_blocks <= 0) return;
    static const int S = [] {
        const char *e = getenv("XG_INFLATE_S");
        return e ? atoi(e) : 32;
    }();
    if (S == 16) launch_inflate_s<16>(st, comp, blocks, n_blocks, n_bad);     // experiments only: slower
    else launch_inflate_s<32>(st, comp, blocks, n_blocks, n_bad);
}

__device__ __forceinline__ unsigned long long sort_key(int32_t tid, int32_t pos) {
    if (tid < 0) return 0x7fffffff00000000ull;
    return ((unsigned long long)(uint32_t)tid << 32) | (uint32_t)pos;
}

struct RecGeom {
    int32_t pos, end, aln;
    uint32_t fmq, n_words, seq_words, n_cig, l_name, l_seq;
    bool ok;
};

// decode.cpp: rec_info()
__device__ __forceinline__ RecGeom rec_geom(const uint8_t *r, uint32_t bs, bool want_seq) {
    RecGeom o;
    o.pos = (int32_t)ld32u(r + 8);
    const uint32_t w12 = ld32u(r + 12), w16 = ld32u(r + 16);
    o.l_name = w12 & 0xff;
    const uint32_t mapq = (w12 >> 8) & 0xff;
    o.n_cig = w16 & 0xffff;
    const uint32_t flag = w16 >> 16;
    o.l_seq = ld32u(r + 20);
    const unsigned long long need = 36ull + o.l_name + 4ull * o.n_cig + ((unsigned long long)o.l_seq + 1) / 2 + o.l_seq;
    o.ok = need <= (unsigned long long)bs + 4ull && o.l_seq < 0x40000000u;
    if (!o.ok) return o;
    const uint8_t *cig = r + 36 + o.l_name;
    long long rlen = 0, aln = 0;
    uint32_t first_op = 15;
    for (uint32_t i = 0; i < o.n_cig; i++) {
        const uint32_t w = ld32u(cig + 4 * i), op = w & 15, l = w >> 4;
        if (i == 0) first_op = op;
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += l;
        if (op == 0 || op == 7 || op == 8) aln += l;
    }
    if (flag & 4) rlen = 0;
    if (rlen == 0) rlen = 1;
    o.end = (int32_t)(o.pos + rlen);
    o.aln = (int32_t)aln;
    const bool simple = o.n_cig == 1 && (first_op == 0 || first_op == 7 || first_op == 8) && aln == rlen && !(flag & 4);
    uint32_t ncw;
    if (simple) {
        ncw = 0;
        o.n_words = 0;
    } else if (o.n_cig == 0) {
        ncw = 1;
        o.n_words = 1;
    } else if (o.n_cig < 255) {
        ncw = o.n_cig;
        o.n_words = o.n_cig;
    } else {
        ncw = 255;
        o.n_words = o.n_cig + 1;
    }
    o.fmq = flag | (mapq << 16) | (ncw << 24);
    o.seq_words = want_seq ? (((o.l_seq + 1) / 2 + 3) / 4) : 0;
    return o;
}

// One thread per BGZF block: check that the block holds whole records in coordinate order and
// size what the kept ones will store.
__global__ void k_walk(const BgzfBlockDev *blocks, int32_t n_blocks, unsigned long long hdr_end,
                       const int32_t *tid_map, int32_t n_ref, int want_seq, BlkInfo *info, int *maxes) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const BgzfBlockDev bk = blocks[b];
    BlkInfo bi;
    bi.n_all = bi.n_kept = bi.cig = bi.seq = 0;
    bi.first_key = ~0ull;
    bi.last_key = 0;
    bi.status = 0;
    bi.n_starts = 0;
    if (bk.uoff + bk.isize <= hdr_end) {
        info[b] = bi;
        return;
    }
    uint32_t off = bk.uoff < hdr_end ? (uint32_t)(hdr_end - bk.uoff) : 0u;      // offsets within the block
    const uint32_t end = bk.isize;
    int32_t max_aln = 0, max_span = 0, prev_kept_tid = -2;
    while (off < end) {
        if (off + 4 > end) {
            bi.status = 1;
            break;
        }
        const uint8_t *r = bk.uptr + off;
        const uint32_t bs = ld32u(r);
        if (bs < 32) {
            bi.status = 2;
            break;
        }
        if ((unsigned long long)off + 4ull + bs > end) {
            bi.status = 1;
            break;
        }
        const int32_t tid = (int32_t)ld32u(r + 4), pos = (int32_t)ld32u(r + 8);
        if (tid >= n_ref) {
            bi.status = 2;
            break;
        }
        const unsigned long long key = sort_key(tid, pos);
        if (bi.n_all == 0) bi.first_key = key;
        else if (key < bi.last_key) bi.status = 3;
        bi.last_key = key;
        bi.n_all++;
        if (tid >= 0 && tid_map[tid] >= 0) {
            const RecGeom g = rec_geom(r, bs, want_seq != 0);
            if (!g.ok) {
                bi.status = 2;
                break;
            }
            bi.n_kept++;
            bi.cig += g.n_words;
            bi.seq += g.seq_words;
            max_aln = max(max_aln, g.aln);
            max_span = max(max_span, g.end - g.pos);
            if (tid != prev_kept_tid) bi.n_starts++;
            prev_kept_tid = tid;
        }
        off += 4u + bs;
    }
    info[b] = bi;
    if (max_aln > 0) atomicMax(&maxes[0], max_aln);
    if (max_span > 0) atomicMax(&maxes[1], max_span);
}

// keys.hpp: key_pack().  Returns false when the string needs the host's intern table.
__device__ __forceinline__ bool pack_key(const uint8_t *s, const uint8_t *end, bool nul_terminated, int n_fixed,
                                         unsigned long long *out) {
    unsigned long long k = 0;
    int bits = 0;
    for (int i = 0; nul_terminated ? (s + i < end) : (i < n_fixed); i++) {
        const uint32_t c = s[i];
        if (nul_terminated && c == 0) break;
        int code;
        switch (c) {
            case 'A': code = 1; break;
            case 'C': code = 2; break;
            case 'G': code = 3; break;
            case 'T': code = 4; break;
            case 'N': code = 5; break;
            case '-': code = 6; break;
            default: code = (c >= '0' && c <= '9') ? 7 : -1;
        }
        if (code < 0) return false;
        if (code < 7) {
            if (bits + 3 > 63) return false;
            k |= (unsigned long long)code << (63 - bits - 3);
            bits += 3;
        } else {
            if (bits + 7 > 63) return false;
            k |= (unsigned long long)((7u << 4) | (c - '0')) << (63 - bits - 7);
            bits += 7;
        }
    }
    *out = k;
    return true;
}

// decode.cpp: tag_key() for the value at `t` (type byte first)
__device__ __forceinline__ bool tag_key_dev(const uint8_t *t, const uint8_t *end, bool is_cell, unsigned long long *out) {
    const uint32_t typ = t[0];
    if (typ == 'Z' || typ == 'H') return pack_key(t + 1, end, true, 0, out);
    if (typ == 'A') return pack_key(t + 1, end, false, 1, out);
    if (is_cell) {
        *out = XG_KEY_NOMATCH;
        return true;
    }
    return false;    // numeric UMI: interned under a type-tagged spelling on the host
}

struct RunStart {
    long long rec;      // global record index
    int32_t tid;
    int32_t bam;
};

struct ExtractArgs {
    const BgzfBlockDev *blocks;
    const BlkInfo *info;
    const unsigned long long *rec_base;   // per block: global index of its first kept record
    const unsigned long long *cig_base;
    const unsigned long long *seq_base;
    const int32_t *tid_map;
    int32_t n_blocks, bam_idx;
    unsigned long long hdr_end;
    int want_seq, has_cell, has_umi;
    uint32_t cell_tag, umi_tag;           // two characters, little endian
    int2 *pos_end;
    uint32_t *fmq, *cig_off, *seq_off, *cigar, *seq;
    ulonglong2 *keys;
    RunStart *starts;
    int *n_starts;
    int *n_need_host;
};

__global__ void k_extract(const ExtractArgs a) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.n_blocks) return;
    const BlkInfo bi = a.info[b];
    if (bi.n_kept == 0) return;
    const BgzfBlockDev bk = a.blocks[b];
    uint32_t off = bk.uoff < a.hdr_end ? (uint32_t)(a.hdr_end - bk.uoff) : 0u;
    const uint32_t end = bk.isize;
    unsigned long long g = a.rec_base[b], co = a.cig_base[b], so = a.seq_base[b];
    int32_t prev_tid = -2;
    int need_host = 0;
    while (off < end) {
        const uint8_t *r = bk.uptr + off;
        const uint32_t bs = ld32u(r);
        const int32_t tid = (int32_t)ld32u(r + 4);
        off += 4u + bs;
        if (tid < 0 || a.tid_map[tid] < 0) continue;
        const RecGeom q = rec_geom(r, bs, a.want_seq != 0);
        a.pos_end[g] = make_int2(q.pos, q.end);
        a.fmq[g] = q.fmq;
        const uint8_t *cig = r + 36 + q.l_name;
        const uint32_t ncw = q.fmq >> 24;
        if (ncw == 0) {
            a.cig_off[g] = (uint32_t)co;
        } else if (q.n_cig == 0) {
            a.cig_off[g] = (uint32_t)co;
            a.cigar[co++] = 6;
        } else {
            if (ncw == 255) a.cigar[co++] = q.n_cig;
            a.cig_off[g] = (uint32_t)co;
            for (uint32_t i = 0; i < q.n_cig; i++) a.cigar[co++] = ld32u(cig + 4 * i);
        }
        const uint8_t *sq = cig + 4ull * q.n_cig;
        const uint32_t seq_bytes = (q.l_seq + 1) / 2;
        if (a.want_seq) {
            a.seq_off[g] = q.seq_words ? (uint32_t)so : 0xFFFFFFFFu;
            for (uint32_t w = 0; w < q.seq_words; w++) {
                uint32_t v = ld32u(sq + 4 * w);
                const uint32_t left = seq_bytes - 4 * w;
                if (left < 4) v &= (1u << (8 * left)) - 1u;
                a.seq[so++] = v;
            }
        }
        // aux fields: one pass finds both tags (first occurrence wins, as bam_aux_get)
        const uint8_t *p = sq + seq_bytes + q.l_seq, *rend = r + 4 + bs;
        const uint8_t *t_cell = nullptr, *t_umi = nullptr;
        int want = (a.has_cell ? 1 : 0) + (a.has_umi ? 1 : 0);
        while (want > 0 && p + 3 <= rend) {
            const uint32_t tag = ld16u(p), typ = p[2];
            if (a.has_cell && !t_cell && tag == a.cell_tag) {
                t_cell = p + 2;
                want--;
            }
            // host order: the cell tag is looked up first, then the UMI tag, each from the start;
            // identical tags resolve to the same field
            if (a.has_umi && !t_umi && tag == a.umi_tag) {
                t_umi = p + 2;
                want--;
            }
            const uint8_t *v = p + 3;
            unsigned long long sz;
            if (typ == 'A' || typ == 'c' || typ == 'C') sz = 1;
            else if (typ == 's' || typ == 'S') sz = 2;
            else if (typ == 'i' || typ == 'I' || typ == 'f') sz = 4;
            else if (typ == 'Z' || typ == 'H') {
                const uint8_t *z = v;
                while (z < rend && *z) z++;
                if (z >= rend) break;
                sz = (unsigned long long)(z - v) + 1;
            } else if (typ == 'B') {
                if (v + 5 > rend) break;
                const uint32_t st = v[0], cnt = ld32u(v + 1);
                const unsigned long long es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
                sz = 5 + es * cnt;
            } else break;
            p = v + sz;
        }
        unsigned long long ck = XG_KEY_NONE, uk = XG_KEY_NONE;
        if (t_cell && !tag_key_dev(t_cell, rend, true, &ck)) need_host++;
        if (a.has_umi) {
            if (t_umi && !tag_key_dev(t_umi, rend, false, &uk)) need_host++;
        } else {
            if (!pack_key(r + 36, rend, false, q.l_name ? (int)q.l_name - 1 : 0, &uk)) need_host++;
        }
        a.keys[g] = make_ulonglong2(ck, uk);
        if (tid != prev_tid) {
            const int s = atomicAdd(a.n_starts, 1);
            a.starts[s] = RunStart{(long long)g, tid, a.bam_idx};
            prev_tid = tid;
        }
        g++;
    }
    if (need_host) atomicAdd(a.n_need_host, need_host);
}

__global__ void k_tile_index2(const int2 *pos_end, xg_tile *tiles, int32_t n_tiles) {
    int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= n_tiles) return;
    int lane = threadIdx.x & 31;
    xg_tile tl = tiles[t];
    int32_t m = INT32_MIN;
    for (int k = lane; k < tl.n_rec; k += 32) m = max(m, pos_end[tl.rec_beg + k].y);
    for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
    if (lane == 0) {
        tiles[t].first_pos = pos_end[tl.rec_beg].x;
        tiles[t].max_end = m;
    }
}

struct DevBam {
    std::vector<uint8_t *> slabs;           // inflated bytes; a block lies within one slab
    BgzfBlockDev *blocks = nullptr;
    BlkInfo *info = nullptr;
    int32_t *tid_map = nullptr;
    unsigned long long *bases = nullptr;    // 3 x n_blocks
    int32_t n_blocks = 0, n_ref = 0;
    uint64_t hdr_end = 0, usize = 0;
    std::vector<BlkInfo> h_info;
    int64_t n_kept = 0, n_all = 0, cig = 0, seq = 0, n_starts = 0;
    int32_t max_aln = 0, max_span = 0;
    xg_ctx *ctx = nullptr;
    void release() {          // buffers go back to the context's device pool (reused by the next call)
        for (uint8_t *p : slabs) ctx->dev_put(p);
        slabs.clear();
        ctx->dev_put(blocks);
        ctx->dev_put(info);
        ctx->dev_put(tid_map);
        ctx->dev_put(bases);
        blocks = nullptr;
        info = nullptr;
        tid_map = nullptr;
        bases = nullptr;
    }
};

double now_ms();
bool g_lap_on = false;          // XG_DECODE_TIMING: host-side phase times on stderr
double g_lap_t = 0;
void lap(const char *what) {
    if (!g_lap_on) return;
    const double t = now_ms();
    fprintf(stderr, "[device decode] %-14s %8.2f ms\n", what, t - g_lap_t);
    g_lap_t = t;
}

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---- file -> HBM -> inflated, pipelined -------------------------------------------------------
// The compressed file is read straight into two pinned staging buffers (several pread threads
// per chunk) and copied to the device chunk by chunk on the copy stream; the file is never held
// in host memory.  The BGZF block index is built from each chunk while it is staged (the
// unscanned tail of a chunk -- a block cut by the chunk boundary, < 64 KiB + header -- is carried
// in front of the next one), and the blocks that are complete are inflated on the compute
// stream as soon as their chunk has landed: disk read, PCIe copy and inflate overlap.
// Inflated bytes go to slabs sized from the first chunk's compression ratio.
constexpr size_t STAGE_HEAD = 128u << 10;
size_t stage_bytes() {          // XG_STAGE_BYTES: tests use small chunks to exercise the carry
    const char *e = getenv("XG_STAGE_BYTES");
    const size_t v = e ? (size_t)atoll(e) : (64u << 20);
    return std::max<size_t>(v, STAGE_HEAD) & ~(size_t)4095;
}

bool pread_all(int fd, uint8_t *dst, size_t len, uint64_t off) {
    while (len) {
        ssize_t got = pread(fd, dst, len, (off_t)off);
        if (got <= 0) return false;
        dst += got;
        off += (uint64_t)got;
        len -= (size_t)got;
    }
    return true;
}

bool pread_parallel(int fd, uint8_t *dst, size_t len, uint64_t off, int n_threads) {
    if (len < (4u << 20) || n_threads <= 1) return pread_all(fd, dst, len, off);
    std::atomic<bool> ok(true);
    std::vector<std::thread> th;
    const size_t slice = ((len + n_threads - 1) / n_threads + 4095) & ~(size_t)4095;
    for (size_t b = 0; b < len; b += slice)
        th.emplace_back([=, &ok] {
            if (!pread_all(fd, dst + b, std::min(slice, len - b), off + b)) ok = false;
        });
    for (auto &t : th) t.join();
    return ok;
}

// comp: device buffer of csize + 16 bytes.  On success every block's inflate kernel has been
// queued on ctx->stream; db.{blocks, slabs, n_blocks, usize, hdr_end, n_ref} are set.
int stream_and_inflate(xg_ctx *ctx, const char *path, int fd, uint64_t csize, uint8_t *comp, size_t cap_blocks,
                       int *d_bad, DevBam &db, double *t_read) {
    const size_t STAGE_BYTES = stage_bytes();
    uint8_t *stage[2] = {(uint8_t *)ctx->pinned_get(STAGE_HEAD + STAGE_BYTES), (uint8_t *)ctx->pinned_get(STAGE_HEAD + STAGE_BYTES)};
    cudaEvent_t done[2] = {nullptr, nullptr};
    auto finish = [&](int code, const std::string &msg) {
        cudaStreamSynchronize(ctx->copy_stream);       // staging buffers may still be in flight
        for (int k = 0; k < 2; k++) {
            if (stage[k]) ctx->pinned_put(stage[k]);
            if (done[k]) cudaEventDestroy(done[k]);
        }
        return code ? ctx->fail(code, msg) : XG_OK;
    };
    if (!stage[0] || !stage[1]) return finish(XG_E_NOMEM, "out of pinned host memory for the staging buffers");
    cudaEventCreateWithFlags(&done[0], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&done[1], cudaEventDisableTiming);
    const int n_threads = (int)std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
    uint64_t next_off = 0, uoff = 0;
    const uint8_t *prev_data = nullptr;
    size_t prev_len = 0, n_blocks = 0;
    bool header_done = false;
    xg_dec::Header hdr;
    std::vector<xg_dec::BgzfBlock> hb;            // the blocks of the current chunk (host descriptors)
    std::vector<BgzfBlockDev> dv;
    uint8_t *slab = nullptr;
    size_t slab_cap = 0, slab_used = 0;
    if (csize == 0) return finish(XG_E_FORMAT, std::string("'") + path + "' is not BGZF (empty file)");
    uint64_t k = 0;
    for (uint64_t c0 = 0; c0 < csize; c0 += STAGE_BYTES, k++) {
        const int si = (int)(k & 1);
        const size_t len = (size_t)std::min<uint64_t>(STAGE_BYTES, csize - c0);
        if (k >= 2) cudaEventSynchronize(done[si]);
        uint8_t *data = stage[si] + STAGE_HEAD;
        const size_t carry = (size_t)(c0 - next_off);
        if (carry) memcpy(data - carry, prev_data + prev_len - carry, carry);
        const double t0 = now_ms();
        if (!pread_parallel(fd, data, len, c0, n_threads)) return finish(XG_E_IO, std::string("short read on '") + path + "'");
        *t_read += now_ms() - t0;
        cudaMemcpyAsync(comp + c0, data, len, cudaMemcpyHostToDevice, ctx->copy_stream);
        const uint64_t view_end = c0 + len;
        hb.clear();
        while (next_off < view_end) {
            const uint8_t *p = data - (c0 - next_off);      // next_off >= c0 - carry
            uint32_t total = 0, hl = 0;
            int rc = xg_dec::bgzf_block_header(p, view_end - next_off, &total, &hl);
            if (rc == 0 && next_off + total > view_end) rc = 1;
            if (rc == 1) {
                if (view_end == csize) return finish(XG_E_FORMAT, std::string("truncated BGZF block in '") + path + "'");
                break;
            }
            if (rc < 0) return finish(XG_E_FORMAT, std::string("'") + path + "' is not BGZF (bad block header)");
            xg_dec::BgzfBlock b;
            b.coff = next_off + hl;
            b.clen = total - hl - 8;
            memcpy(&b.isize, p + total - 4, 4);
            b.uoff = uoff;
            if (b.isize > 65536) return finish(XG_E_FORMAT, "BGZF block larger than 64 KiB");
            uoff += b.isize;
            hb.push_back(b);
            next_off += total;
        }
        if (view_end - next_off > STAGE_HEAD) return finish(XG_E_FORMAT, "BGZF block larger than 64 KiB");
        if (!header_done) {
            // the BAM header: inflate leading blocks on the host until it parses
            size_t nb = 1;
            while (true) {
                const size_t take = std::min(nb, hb.size());
                xg_dec::Bytes u;
                int rc = xg_dec::inflate_blocks(data, c0, hb.data(), take, u, 1);
                if (rc) return finish(rc, xg_host_last_error());
                rc = xg_dec::parse_header(u, hdr, path);
                if (rc == XG_OK) break;
                if (rc != XG_E_LIMIT) return finish(rc, xg_host_last_error());
                if (take >= hb.size())
                    return finish(view_end == csize ? XG_E_FORMAT : XG_E_UNSUPPORTED,
                                  std::string("BAM header of '") + path + "' does not end within the first staged chunk");
                nb *= 2;
            }
            header_done = true;
        }
        if (n_blocks + hb.size() > cap_blocks)
            return finish(XG_E_UNSUPPORTED, std::string("'") + path + "' has unusually small BGZF blocks");
        // place the chunk's blocks in the slabs, hand their descriptors to the device, inflate
        dv.resize(hb.size());
        for (size_t i = 0; i < hb.size(); i++) {
            if (!slab || slab_used + hb[i].isize + 16 > slab_cap) {
                // what is left of the file at the ratio seen so far (+10%), at least 64 MiB
                const double ratio = next_off ? (double)uoff / (double)next_off : 4.0;
                size_t want = (size_t)((double)(csize - std::min<uint64_t>(csize, hb[i].coff)) * ratio * 1.1) + (64u << 20);
                size_t free_b = 0, total_b = 0;
                cudaMemGetInfo(&free_b, &total_b);
                if (want + (2ull << 30) > free_b + ctx->dev_idle_bytes())
                    return finish(XG_E_UNSUPPORTED, "BAM too large to inflate on this device in one piece");
                const double ta = now_ms();
                slab = (uint8_t *)ctx->dev_get(want);
                ctx->timing[9] += now_ms() - ta;
                if (!slab) return finish(XG_E_CUDA, "out of device memory for the inflated BAM");
                db.slabs.push_back(slab);
                slab_cap = want;
                slab_used = 0;
            }
            dv[i].coff = hb[i].coff;
            dv[i].clen = hb[i].clen;
            dv[i].isize = hb[i].isize;
            dv[i].uoff = hb[i].uoff;
            dv[i].uptr = slab + slab_used;
            slab_used += hb[i].isize;
        }
        if (!dv.empty())
            cudaMemcpyAsync(db.blocks + n_blocks, dv.data(), dv.size() * sizeof(BgzfBlockDev), cudaMemcpyHostToDevice,
                            ctx->copy_stream);
        cudaEventRecord(done[si], ctx->copy_stream);
        cudaStreamWaitEvent(ctx->stream, done[si], 0);
        launch_inflate(ctx->stream, comp, db.blocks + n_blocks, (int32_t)dv.size(), d_bad);
        n_blocks += hb.size();
        prev_data = data;
        prev_len = len;
    }
    db.n_blocks = (int32_t)n_blocks;
    db.usize = uoff;
    db.hdr_end = hdr.end_off;
    db.n_ref = (int32_t)hdr.names.size();
    return finish(XG_OK, "");
}

// Stream one BAM's compressed bytes to the device, inflate them and walk the records.  On
// success db holds the inflated stream and the per-block sizing.
int inflate_and_walk(xg_ctx *ctx, const char *path, const int32_t *tid_map, int32_t tid_map_len, int want_seq,
                     int *d_counters, DevBam &db, double *t_read, double *t_h2d, double *t_walk) {
    const int fd = open(path, O_RDONLY);
    struct stat stt;
    if (fd < 0 || fstat(fd, &stt) != 0) {
        if (fd >= 0) close(fd);
        return ctx->fail(XG_E_IO, std::string("cannot open '") + path + "'");
    }
    const size_t csize = (size_t)stt.st_size;
    db.ctx = ctx;
    if (!ctx->copy_stream && cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
        close(fd);
        return ctx->fail(XG_E_CUDA, "cannot create the copy stream");
    }
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    if (csize + (2ull << 30) > free_b + ctx->dev_idle_bytes()) {
        close(fd);
        return ctx->fail(XG_E_UNSUPPORTED, "BAM too large to inflate on this device in one piece");
    }
    // htslib fills blocks to ~64 KiB of payload; a file whose blocks average under 1 KiB
    // compressed is not worth a device pass
    const size_t cap_blocks = csize / 1024 + 4096;
    double t_a0 = now_ms();
    uint8_t *comp = (uint8_t *)ctx->dev_get(csize + 16);
    db.blocks = (BgzfBlockDev *)ctx->dev_get((cap_blocks + 1) * sizeof(BgzfBlockDev));
    db.info = (BlkInfo *)ctx->dev_get((cap_blocks + 1) * sizeof(BlkInfo));
    db.bases = (unsigned long long *)ctx->dev_get((3 * cap_blocks + 1) * 8);
    ctx->timing[9] += now_ms() - t_a0;
    auto bail = [&](int code, const std::string &msg) {
        cudaStreamSynchronize(ctx->stream);
        ctx->dev_put(comp);
        db.release();
        return ctx->fail(code, msg);
    };
    if (!comp || !db.blocks || !db.info || !db.bases) {
        close(fd);
        return bail(XG_E_CUDA, "out of device memory for the compressed BAM");
    }
    cudaStream_t st = ctx->stream;
    cudaMemsetAsync(d_counters, 0, 8 * sizeof(int), st);
    cudaEventRecord(ctx->ev[6], st);
    cudaStreamWaitEvent(ctx->copy_stream, ctx->ev[6], 0);      // comp may be a recycled buffer still in use
    int rc = stream_and_inflate(ctx, path, fd, csize, comp, cap_blocks, d_counters + 0, db, t_read);
    close(fd);
    cudaEventRecord(ctx->ev[7], st);                            // all inflate kernels queued before this
    if (rc) {
        const std::string msg = ctx->err;
        return bail(rc, msg);
    }
    lap("stream");
    if (tid_map_len < db.n_ref) return bail(XG_E_ARG, "tid_map shorter than the BAM's contig list");
    const size_t nb = (size_t)db.n_blocks;
    db.tid_map = (int32_t *)ctx->dev_get(((size_t)db.n_ref + 1) * 4);
    if (!db.tid_map) return bail(XG_E_CUDA, "out of device memory");
    cudaMemcpyAsync(db.tid_map, tid_map, (size_t)db.n_ref * 4, cudaMemcpyHostToDevice, st);
    if (nb) {
        k_walk<<<(unsigned)((nb + 63) / 64), 64, 0, st>>>(db.blocks, db.n_blocks, db.hdr_end, db.tid_map, db.n_ref, want_seq,
                                                          db.info, d_counters + 2);
    }
    cudaEventRecord(ctx->ev[4], st);
    db.h_info.resize(nb);
    int h_cnt[8];
    cudaMemcpyAsync(h_cnt, d_counters, sizeof(h_cnt), cudaMemcpyDeviceToHost, st);
    if (nb) cudaMemcpyAsync(db.h_info.data(), db.info, nb * sizeof(BlkInfo), cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return bail(XG_E_CUDA, std::string("device inflate: ") + cudaGetErrorString(e));
    lap("inflate sync");
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]);
    *t_h2d += ms;                                               // read + copy + inflate, overlapped
    cudaEventElapsedTime(&ms, ctx->ev[7], ctx->ev[4]);
    *t_walk += ms;
    ctx->dev_put(comp);
    comp = nullptr;
    if (h_cnt[0]) return bail(XG_E_FORMAT, std::string("BGZF inflate failed (corrupt block) in '") + path + "'");
    // per-block checks + sizes
    std::vector<unsigned long long> bases(3 * nb);
    unsigned long long last_key = 0;
    for (size_t b = 0; b < nb; b++) {
        const BlkInfo &bi = db.h_info[b];
        if (bi.status == 2) return bail(XG_E_FORMAT, std::string("corrupt BAM record in '") + path + "'");
        if (bi.status == 3) return bail(XG_E_FORMAT, std::string("'") + path + "' is not coordinate sorted");
        if (bi.status == 1)
            return bail(XG_E_UNSUPPORTED, std::string("records of '") + path + "' cross BGZF block boundaries");
        if (bi.n_all) {
            if (bi.first_key < last_key) return bail(XG_E_FORMAT, std::string("'") + path + "' is not coordinate sorted");
            last_key = bi.last_key;
        }
        bases[b] = (unsigned long long)db.n_kept;
        bases[nb + b] = (unsigned long long)db.cig;
        bases[2 * nb + b] = (unsigned long long)db.seq;
        db.n_all += bi.n_all;
        db.n_kept += bi.n_kept;
        db.cig += bi.cig;
        db.seq += bi.seq;
        db.n_starts += bi.n_starts;
    }
    db.h_info.clear();
    db.h_info.shrink_to_fit();
    // stash the block-local bases; the caller adds the per-BAM offsets in the kernel arguments
    if (nb) {
        e = cudaMemcpy(db.bases, bases.data(), 3 * nb * 8, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return bail(XG_E_CUDA, std::string("device decode: ") + cudaGetErrorString(e));
    }
    db.max_aln = h_cnt[2];
    db.max_span = h_cnt[3];
    lap("block sizes");
    return XG_OK;
}

__global__ void k_add_base(unsigned long long *v, int64_t n, unsigned long long add) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] += add;
}

}  // namespace

extern "C" {

// Testing / validation entry: inflate a whole BGZF file on the device, return the bytes.
int xg_bgzf_inflate_device(xg_ctx *ctx, const char *path, uint8_t *out, int64_t cap, int64_t *n_out) {
    if (!ctx || !ctx->stream) return ctx ? ctx->fail(XG_E_CUDA, "context has no device") : XG_E_ARG;
    if (!path || !n_out) return ctx->fail(XG_E_ARG, "xg_bgzf_inflate_device: null argument");
    XG_CUDA(cudaSetDevice(ctx->device));
    xg_dec::Bytes f;
    std::vector<xg_dec::BgzfBlock> blocks;
    int rc = xg_dec::read_file(path, f);
    if (!rc) rc = xg_dec::scan_bgzf(f, blocks, path);
    if (rc) return ctx->fail(rc, xg_host_last_error());
    const uint64_t usize = blocks.empty() ? 0 : blocks.back().uoff + blocks.back().isize;
    *n_out = (int64_t)usize;
    if (!out || cap < (int64_t)usize) return XG_OK;      // size query
    uint8_t *comp = nullptr, *ubuf = nullptr;
    BgzfBlockDev *dblk = nullptr;
    XG_GET(cnt, int, "gd_counters", 8);
    if (cudaMalloc(&comp, f.size() + 16) != cudaSuccess || cudaMalloc(&ubuf, usize + 16) != cudaSuccess ||
        cudaMalloc(&dblk, (blocks.size() + 1) * sizeof(BgzfBlockDev)) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(comp);
        cudaFree(ubuf);
        cudaFree(dblk);
        return ctx->fail(XG_E_CUDA, "out of device memory");
    }
    cudaStream_t st = ctx->stream;
    cudaMemcpyAsync(comp, f.data(), f.size(), cudaMemcpyHostToDevice, st);
    std::vector<BgzfBlockDev> dv(blocks.size());
    for (size_t i = 0; i < blocks.size(); i++) {
        dv[i].coff = blocks[i].coff;
        dv[i].clen = blocks[i].clen;
        dv[i].isize = blocks[i].isize;
        dv[i].uoff = blocks[i].uoff;
        dv[i].uptr = ubuf + blocks[i].uoff;
    }
    cudaMemcpyAsync(dblk, dv.data(), dv.size() * sizeof(BgzfBlockDev), cudaMemcpyHostToDevice, st);
    cudaMemsetAsync(cnt, 0, 8 * sizeof(int), st);
    cudaEventRecord(ctx->ev[0], st);
    launch_inflate(st, comp, dblk, (int32_t)blocks.size(), cnt);
    cudaEventRecord(ctx->ev[1], st);
    int bad = 0;
    cudaMemcpyAsync(&bad, cnt, 4, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(out, ubuf, usize, cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    float ms = 0;
    if (e == cudaSuccess) cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->timing[0] = ms;
    cudaFree(comp);
    cudaFree(ubuf);
    cudaFree(dblk);
    if (e != cudaSuccess) return ctx->fail(XG_E_CUDA, std::string("device inflate: ") + cudaGetErrorString(e));
    if (bad) return ctx->fail(XG_E_FORMAT, "BGZF inflate failed (corrupt block)");
    return XG_OK;
}

int xg_decode_bams_device(xg_ctx *ctx, int32_t n_bams, const char *const *paths, const int32_t *const *tid_map,
                          const int32_t *tid_map_len, const char *cell_tag, const char *umi_tag, int32_t want_seq,
                          xg_dreads **out, int64_t *n_records_seen) {
    if (!ctx || !ctx->stream) return ctx ? ctx->fail(XG_E_CUDA, "context has no device") : XG_E_ARG;
    if (n_bams < 0 || !out) return ctx->fail(XG_E_ARG, "xg_decode_bams_device: bad argument");
    if (cell_tag && strlen(cell_tag) != 2) return ctx->fail(XG_E_ARG, "cell tag must have 2 characters");
    if (umi_tag && strlen(umi_tag) != 2) return ctx->fail(XG_E_ARG, "UMI tag must have 2 characters");
    XG_CUDA(cudaSetDevice(ctx->device));
    const double t_begin = now_ms();
    for (double &t : ctx->timing) t = 0;
    g_lap_on = getenv("XG_DECODE_TIMING") != nullptr;
    g_lap_t = t_begin;
    XG_GET(cnt, int, "gd_counters", 8);
    std::vector<DevBam> bams((size_t)n_bams);
    auto release_all = [&] {
        for (auto &b : bams)
            if (b.ctx) b.release();
    };
    double t_read = 0, t_h2d = 0, t_walk = 0;
    int64_t n_total = 0, n_seen = 0, cig_total = 0, seq_total = 0, n_starts = 0;
    int32_t max_aln = 0, max_span = 0;
    cudaEventRecord(ctx->ev[2], ctx->stream);
    for (int32_t b = 0; b < n_bams; b++) {
        int rc = inflate_and_walk(ctx, paths[b], tid_map[b], tid_map_len[b], want_seq, cnt, bams[b], &t_read, &t_h2d, &t_walk);
        if (rc) {
            release_all();
            return rc;
        }
        n_total += bams[b].n_kept;
        n_seen += bams[b].n_all;
        cig_total += bams[b].cig;
        seq_total += bams[b].seq;
        n_starts += bams[b].n_starts;
        max_aln = std::max(max_aln, bams[b].max_aln);
        max_span = std::max(max_span, bams[b].max_span);
    }
    lap("inflate+walk");
    if (cig_total >= (1LL << 32) || seq_total >= (1LL << 32)) {
        release_all();
        return ctx->fail(XG_E_LIMIT, "batch too large for 32-bit stream offsets; decode fewer reads per batch");
    }
    xg_dreads *d = new xg_dreads();
    d->pooled = true;
    d->n_reads = n_total;
    d->n_cigar = cig_total;
    d->n_seq_words = seq_total;
    d->max_aln_len = max_aln;
    d->max_span = max_span;
    auto fail_free = [&](int code, const std::string &msg) {
        release_all();
        xg_dreads_free(ctx, d);
        return ctx->fail(code, msg);
    };
    const size_t n = (size_t)n_total;
    d->pos_end = (int2 *)ctx->dev_get(n * 8 + 16);
    d->fmq = (uint32_t *)ctx->dev_get(n * 4 + 16);
    d->cig_off = (uint32_t *)ctx->dev_get((n + 1) * 4 + 16);
    d->keys = (ulonglong2 *)ctx->dev_get(n * 16 + 16);
    d->cigar = (uint32_t *)ctx->dev_get((size_t)cig_total * 4 + 16);
    if (want_seq) {
        d->seq_off = (uint32_t *)ctx->dev_get(n * 4 + 16);
        d->seq = (uint32_t *)ctx->dev_get((size_t)seq_total * 4 + 16);
    }
    RunStart *d_starts = nullptr;
    if (!d->pos_end || !d->fmq || !d->cig_off || !d->keys || !d->cigar || (want_seq && (!d->seq_off || !d->seq)) ||
        cudaMalloc(&d_starts, ((size_t)n_starts + 1) * sizeof(RunStart)) != cudaSuccess) {
        cudaGetLastError();
        return fail_free(XG_E_CUDA, "out of device memory for the read batch");
    }
    lap("alloc");
    cudaStream_t st = ctx->stream;
    cudaMemsetAsync(cnt, 0, 8 * sizeof(int), st);
    cudaEventRecord(ctx->ev[0], st);
    int64_t rec0 = 0, cig0 = 0, seq0 = 0;
    for (int32_t b = 0; b < n_bams; b++) {
        DevBam &db = bams[b];
        const int64_t nb = db.n_blocks;
        if (nb && db.n_kept) {
            if (rec0) k_add_base<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(db.bases, nb, (unsigned long long)rec0);
            if (cig0) k_add_base<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(db.bases + nb, nb, (unsigned long long)cig0);
            if (seq0) k_add_base<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(db.bases + 2 * nb, nb, (unsigned long long)seq0);
            ExtractArgs a;
            a.blocks = db.blocks;
            a.info = db.info;
            a.rec_base = db.bases;
            a.cig_base = db.bases + nb;
            a.seq_base = db.bases + 2 * nb;
            a.tid_map = db.tid_map;
            a.n_blocks = db.n_blocks;
            a.bam_idx = b;
            a.hdr_end = db.hdr_end;
            a.want_seq = want_seq;
            a.has_cell = cell_tag != nullptr;
            a.has_umi = umi_tag != nullptr;
            a.cell_tag = cell_tag ? ((uint32_t)(uint8_t)cell_tag[0] | ((uint32_t)(uint8_t)cell_tag[1] << 8)) : 0;
            a.umi_tag = umi_tag ? ((uint32_t)(uint8_t)umi_tag[0] | ((uint32_t)(uint8_t)umi_tag[1] << 8)) : 0;
            a.pos_end = d->pos_end;
            a.fmq = d->fmq;
            a.cig_off = d->cig_off;
            a.seq_off = d->seq_off;
            a.cigar = d->cigar;
            a.seq = d->seq;
            a.keys = d->keys;
            a.starts = d_starts;
            a.n_starts = cnt + 4;
            a.n_need_host = cnt + 5;
            k_extract<<<(unsigned)((nb + 63) / 64), 64, 0, st>>>(a);
        }
        rec0 += db.n_kept;
        cig0 += db.cig;
        seq0 += db.seq;
    }
    cudaEventRecord(ctx->ev[1], st);
    const uint32_t sentinel = (uint32_t)cig_total;
    cudaMemcpyAsync(d->cig_off + n, &sentinel, 4, cudaMemcpyHostToDevice, st);
    int h_cnt[8];
    cudaMemcpyAsync(h_cnt, cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    lap("extract");
    release_all();
    if (e != cudaSuccess) {
        cudaFree(d_starts);
        return fail_free(XG_E_CUDA, std::string("device decode: ") + cudaGetErrorString(e));
    }
    {
        float xms = 0;
        cudaEventElapsedTime(&xms, ctx->ev[0], ctx->ev[1]);
        ctx->timing[3] = xms;               // extract kernels
    }
    if (h_cnt[5]) {
        cudaFree(d_starts);
        return fail_free(XG_E_UNSUPPORTED, "cell / UMI keys need the host intern table (" + std::to_string(h_cnt[5]) +
                                               " values are not short ACGTN-/digit strings)");
    }
    std::vector<RunStart> starts((size_t)h_cnt[4]);
    if (!starts.empty()) cudaMemcpy(starts.data(), d_starts, starts.size() * sizeof(RunStart), cudaMemcpyDeviceToHost);
    cudaFree(d_starts);
    lap("release");
    std::sort(starts.begin(), starts.end(), [](const RunStart &x, const RunStart &y) { return x.rec < y.rec; });
    // runs: maximal stretches of kept records of one contig of one BAM (decode.cpp's run_tid)
    int64_t bam_end = 0;
    size_t si = 0;
    for (int32_t b = 0; b < n_bams; b++) {
        bam_end += bams[b].n_kept;
        int32_t run_tid = -2;
        for (; si < starts.size() && starts[si].rec < bam_end; si++) {
            if (starts[si].tid == run_tid) continue;
            run_tid = starts[si].tid;
            if (!d->h_runs.empty() && d->h_runs.back().rec_end < 0) d->h_runs.back().rec_end = starts[si].rec;
            xg_run r;
            r.bam_idx = b;
            r.gid = tid_map[b][run_tid];
            r.rec_beg = starts[si].rec;
            r.rec_end = -1;
            d->h_runs.push_back(r);
        }
        if (!d->h_runs.empty() && d->h_runs.back().rec_end < 0) d->h_runs.back().rec_end = bam_end;
    }
    for (size_t r = 0; r < d->h_runs.size(); r++)
        for (int64_t s = d->h_runs[r].rec_beg; s < d->h_runs[r].rec_end; s += XG_TILE) {
            xg_tile tl;
            tl.rec_beg = s;
            tl.n_rec = (int32_t)std::min<int64_t>(XG_TILE, d->h_runs[r].rec_end - s);
            tl.run = (int32_t)r;
            tl.first_pos = 0;
            tl.max_end = 0;
            d->h_tiles.push_back(tl);
        }
    d->n_runs = (int32_t)d->h_runs.size();
    d->n_tiles = (int32_t)d->h_tiles.size();
    d->runs = (xg_run *)ctx->dev_get((size_t)d->n_runs * sizeof(xg_run) + 16);
    d->tiles = (xg_tile *)ctx->dev_get((size_t)d->n_tiles * sizeof(xg_tile) + 16);
    if (!d->runs || !d->tiles) return fail_free(XG_E_CUDA, "out of device memory for the tile index");
    cudaMemcpyAsync(d->runs, d->h_runs.data(), (size_t)d->n_runs * sizeof(xg_run), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d->tiles, d->h_tiles.data(), (size_t)d->n_tiles * sizeof(xg_tile), cudaMemcpyHostToDevice, st);
    if (d->n_tiles) {
        k_tile_index2<<<(d->n_tiles + 7) / 8, 256, 0, st>>>(d->pos_end, d->tiles, d->n_tiles);
        cudaMemcpyAsync(d->h_tiles.data(), d->tiles, (size_t)d->n_tiles * sizeof(xg_tile), cudaMemcpyDeviceToHost, st);
    }
    cudaEventRecord(ctx->ev[3], st);
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail_free(XG_E_CUDA, std::string("device decode: ") + cudaGetErrorString(e));
    lap("runs+tiles");
    int rc = xg_make_tile_pmax(ctx, d);
    if (rc) {
        xg_dreads_free(ctx, d);
        return rc;
    }
    lap("pmax");
    d->bytes = (int64_t)(n * 32 + (size_t)cig_total * 4 + (size_t)seq_total * 4 + (want_seq ? n * 4 : 0));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]);
    ctx->timing[0] = ms;                    // device span incl. H2D of the compressed files
    ctx->timing[2] = t_walk;                // walk kernels
    ctx->timing[4] = t_h2d;                 // file read + H2D of the compressed bytes + inflate, pipelined
    ctx->timing[8] = t_read;                // file read + block scan on the host
    {
        const double t_f0 = now_ms();
        ctx->dev_trim(8ull << 30);          // keep small inputs' buffers for the next call, give the rest back
        ctx->timing[10] = now_ms() - t_f0;
    }
    ctx->timing[12] = now_ms() - t_begin;   // whole call
    if (n_records_seen) *n_records_seen = n_seen;
    *out = d;
    return XG_OK;
}

}  // extern "C"

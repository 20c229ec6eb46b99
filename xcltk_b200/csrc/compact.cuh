// compact.cuh -- dense per-row counters -> (row, col)-sorted sparse output.
//
// Replaces the emit loops `for i, smp in enumerate(conf.samples): if nu > 0: "%d\t%d\t%d"`
// (xcltk/rdr/fc/core.py:109-124, xcltk/baf/fc/core.py:84-113): columns ascending inside a
// row, rows ascending.  Two passes over the dense row: count, exclusive scan, ordered write.
#pragma once
#include <cstring>

#include "common.cuh"
#include "owner.hpp"

// V: functor  __device__ int operator()(int row, int col) const  -> value (emitted iff > 0)
template <class V>
__global__ void __launch_bounds__(256) k_row_nnz(V v, int32_t n_rows, int32_t n_cols, int32_t *row_nnz) {
    for (int row = blockIdx.x; row < n_rows; row += gridDim.x) {
        int cnt = 0;
        for (int c0 = 0; c0 < n_cols; c0 += blockDim.x) {
            int c = c0 + threadIdx.x;
            int ok = (c < n_cols) && (v(row, c) > 0);
            cnt += __syncthreads_count(ok);
        }
        if (threadIdx.x == 0) row_nnz[row] = cnt;
    }
}

// Exclusive scan of n int32 -> int64 (n+1 outputs); one CTA, carries across chunks.
static __global__ void __launch_bounds__(1024) k_exclusive_scan(const int32_t *in, int64_t *out, int32_t n) {
    __shared__ int64_t warp_sum[32], warp_excl[32];
    __shared__ int64_t carry_s, chunk_total;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < n; base += blockDim.x) {
        int i = base + threadIdx.x;
        int64_t x = (i < n) ? in[i] : 0, s = x;
        for (int d = 1; d < 32; d <<= 1) {
            int64_t y = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += y;
        }
        if (lane == 31) warp_sum[w] = s;
        __syncthreads();
        if (w == 0) {
            int64_t t = warp_sum[lane], ts = t;
            for (int d = 1; d < 32; d <<= 1) {
                int64_t y = __shfl_up_sync(0xffffffffu, ts, d);
                if (lane >= d) ts += y;
            }
            warp_excl[lane] = ts - t;
            if (lane == 31) chunk_total = ts;
        }
        __syncthreads();
        int64_t carry = carry_s;
        if (i < n) out[i] = carry + warp_excl[w] + (s - x);
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + chunk_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry_s;
}

template <class V>
__global__ void __launch_bounds__(256) k_row_write(V v, int32_t n_rows, int32_t n_cols,
                                                   const int64_t *row_ptr, int32_t *o_row,
                                                   int32_t *o_col, int32_t *o_val) {
    __shared__ int warp_cnt[8];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int row = blockIdx.x; row < n_rows; row += gridDim.x) {
        int64_t base = row_ptr[row];
        if (row_ptr[row + 1] == base) continue;
        for (int c0 = 0; c0 < n_cols; c0 += blockDim.x) {
            int c = c0 + threadIdx.x;
            int val = (c < n_cols) ? v(row, c) : 0;
            unsigned m = __ballot_sync(0xffffffffu, val > 0);
            if (lane == 0) warp_cnt[w] = __popc(m);
            __syncthreads();
            int before = 0, total = 0;
            for (int k = 0; k < 8; k++) {
                int x = warp_cnt[k];
                if (k < w) before += x;
                total += x;
            }
            if (val > 0) {
                int64_t o = base + before + __popc(m & ((1u << lane) - 1u));
                o_row[o] = row;
                o_col[o] = c;
                o_val[o] = val;
            }
            base += total;
            __syncthreads();
        }
    }
}

// Runs the three kernels on ctx->stream and copies the result into pinned host memory.
template <class V>
int xg_dense_to_coo(xg_ctx *ctx, V v, int32_t n_rows, int32_t n_cols, const char *tag,
                    xg_coo **out, int *n_launches) {
    std::string t(tag);
    XG_GET(row_nnz, int32_t, (t + "_row_nnz").c_str(), n_rows + 1);
    XG_GET(row_ptr, int64_t, (t + "_row_ptr").c_str(), n_rows + 1);
    int grid = n_rows < 1 ? 1 : (n_rows > 148 * 64 ? 148 * 64 : n_rows);
    k_row_nnz<V><<<grid, 256, 0, ctx->stream>>>(v, n_rows, n_cols, row_nnz);
    k_exclusive_scan<<<1, 1024, 0, ctx->stream>>>(row_nnz, row_ptr, n_rows);
    int64_t nnz = 0;
    XG_CUDA(cudaMemcpyAsync(&nnz, row_ptr + n_rows, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    XG_CUDA(cudaStreamSynchronize(ctx->stream));
    XG_GET(d_row, int32_t, (t + "_coo_row").c_str(), nnz + 1);
    XG_GET(d_col, int32_t, (t + "_coo_col").c_str(), nnz + 1);
    XG_GET(d_val, int32_t, (t + "_coo_val").c_str(), nnz + 1);
    k_row_write<V><<<grid, 256, 0, ctx->stream>>>(v, n_rows, n_cols, row_ptr, d_row, d_col, d_val);
    *n_launches += 3;
    XG_CUDA(cudaGetLastError());

    xg_coo_owner *o = new xg_coo_owner();
    memset(&o->m, 0, sizeof(o->m));
    void *h_row = nullptr, *h_col = nullptr, *h_val = nullptr, *h_ptr = nullptr;
    size_t nb = (size_t)(nnz > 0 ? nnz : 1) * 4;
    h_row = ctx->pinned_get(nb);
    h_col = ctx->pinned_get(nb);
    h_val = ctx->pinned_get(nb);
    h_ptr = ctx->pinned_get((size_t)(n_rows + 1) * 8);
    if (!h_row || !h_col || !h_val || !h_ptr) {
        for (void *p : {h_row, h_col, h_val, h_ptr})
            if (p) ctx->pinned_put(p);
        delete o;
        return ctx->fail(XG_E_NOMEM, "out of pinned host memory for the result");
    }
    o->bufs = {h_row, h_col, h_val, h_ptr};
    o->ctx = ctx;
    cudaEventRecord(ctx->ev[4], ctx->stream);
    if (nnz > 0) {
        cudaMemcpyAsync(h_row, d_row, (size_t)nnz * 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(h_col, d_col, (size_t)nnz * 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(h_val, d_val, (size_t)nnz * 4, cudaMemcpyDeviceToHost, ctx->stream);
    }
    cudaMemcpyAsync(h_ptr, row_ptr, (size_t)(n_rows + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream);
    cudaEventRecord(ctx->ev[5], ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        for (void *p : o->bufs) ctx->pinned_put(p);
        delete o;
        return ctx->fail(XG_E_CUDA, std::string("result D2H: ") + cudaGetErrorString(e));
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]);
    ctx->timing[4] += ms;
    o->m.nnz = nnz;
    o->m.n_rows = n_rows;
    o->m.n_cols = n_cols;
    o->m.row = (const int32_t *)h_row;
    o->m.col = (const int32_t *)h_col;
    o->m.val = (const int32_t *)h_val;
    o->m.row_ptr = (const int64_t *)h_ptr;
    *out = &o->m;
    return XG_OK;
}

// Rows were written to a staging area in completion order (seg_base / seg_nnz per row);
// copy them to their place in row order.
static __global__ void __launch_bounds__(256) k_gather_rows(int32_t n_rows, const int64_t *seg_base,
                                                            const int32_t *seg_nnz, const int64_t *row_ptr,
                                                            const int32_t *st_col, const int32_t *st_val,
                                                            int32_t *o_row, int32_t *o_col, int32_t *o_val) {
    for (int row = blockIdx.x; row < n_rows; row += gridDim.x) {
        const int n = seg_nnz[row];
        const int64_t src = seg_base[row], dst = row_ptr[row];
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            o_row[dst + k] = row;
            o_col[dst + k] = st_col[src + k];
            o_val[dst + k] = st_val[src + k];
        }
    }
}

// seg_nnz -> row_ptr (scan), staging -> (row, col)-sorted device COO (gather), -> pinned host.
// One host synchronisation for nnz.  `tag` names the scratch buffers.
static int xg_staging_to_coo(xg_ctx *ctx, const char *tag, int32_t n_rows, int32_t n_cols,
                             const int64_t *seg_base, const int32_t *seg_nnz, const int32_t *st_col,
                             const int32_t *st_val, xg_coo **out, int *n_launches) {
    std::string t(tag);
    XG_GET(row_ptr, int64_t, (t + "_row_ptr").c_str(), n_rows + 2);
    k_exclusive_scan<<<1, 1024, 0, ctx->stream>>>(seg_nnz, row_ptr, n_rows);
    int64_t nnz = 0;
    XG_CUDA(cudaMemcpyAsync(&nnz, row_ptr + n_rows, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    XG_CUDA(cudaStreamSynchronize(ctx->stream));
    XG_GET(d_row, int32_t, (t + "_coo_row").c_str(), nnz + 1);
    XG_GET(d_col, int32_t, (t + "_coo_col").c_str(), nnz + 1);
    XG_GET(d_val, int32_t, (t + "_coo_val").c_str(), nnz + 1);
    *n_launches += 1;
    if (nnz > 0) {
        k_gather_rows<<<n_rows < 148 * 32 ? n_rows : 148 * 32, 256, 0, ctx->stream>>>(
            n_rows, seg_base, seg_nnz, row_ptr, st_col, st_val, d_row, d_col, d_val);
        *n_launches += 1;
    }
    cudaEventRecord(ctx->ev[3], ctx->stream);
    XG_CUDA(cudaGetLastError());
    xg_coo_owner *o = new xg_coo_owner();
    memset(&o->m, 0, sizeof(o->m));
    void *hp[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t hs[4] = {(size_t)(nnz + 1) * 4, (size_t)(nnz + 1) * 4, (size_t)(nnz + 1) * 4, (size_t)(n_rows + 1) * 8};
    for (int k = 0; k < 4; k++)
        if (!(hp[k] = ctx->pinned_get(hs[k]))) {
            for (int q = 0; q < k; q++) ctx->pinned_put(hp[q]);
            delete o;
            return ctx->fail(XG_E_NOMEM, "out of pinned host memory for the result");
        }
    o->bufs = {hp[0], hp[1], hp[2], hp[3]};
    o->ctx = ctx;
    cudaEventRecord(ctx->ev[4], ctx->stream);
    if (nnz > 0) {
        cudaMemcpyAsync(hp[0], d_row, (size_t)nnz * 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(hp[1], d_col, (size_t)nnz * 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(hp[2], d_val, (size_t)nnz * 4, cudaMemcpyDeviceToHost, ctx->stream);
    }
    cudaMemcpyAsync(hp[3], row_ptr, (size_t)(n_rows + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream);
    cudaEventRecord(ctx->ev[5], ctx->stream);
    cudaError_t ce = cudaStreamSynchronize(ctx->stream);
    if (ce != cudaSuccess) {
        for (void *q : o->bufs) ctx->pinned_put(q);
        delete o;
        return ctx->fail(XG_E_CUDA, std::string("result D2H: ") + cudaGetErrorString(ce));
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]);
    ctx->timing[4] += ms;
    o->m.nnz = nnz;
    o->m.n_rows = n_rows;
    o->m.n_cols = n_cols;
    o->m.row = (const int32_t *)hp[0];
    o->m.col = (const int32_t *)hp[1];
    o->m.val = (const int32_t *)hp[2];
    o->m.row_ptr = (const int64_t *)hp[3];
    *out = &o->m;
    return XG_OK;
}

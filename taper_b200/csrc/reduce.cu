// Broadcast / reduction / layout kernels on [rows, cols] row-major matrices.
// Reference: add_broadcast (src/tensor.rs:636-704), sub_broadcast_rows (:707-770), sum (:890-1018),
// max/argmax (:1021-1088), transpose (:544-591), add_bias_4d (:1972-2031), transpose_4d (:2034-2076).
// All reductions are deterministic (fixed two-stage trees, no atomics).
#include "common.cuh"
#include <cfloat>

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- out[i,f] = a[i,f] + bias[f] (optional relu) ------------------------------------------------
__global__ void __launch_bounds__(kThreads)
add_broadcast_kernel(const float* __restrict__ a, const float* __restrict__ bias, float* __restrict__ out,
                     size_t total, int cols, int relu) {
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += stride) {
        float v = a[i] + __ldg(bias + (i % cols));
        out[i] = relu ? fmaxf(v, 0.0f) : v;
    }
}

__global__ void __launch_bounds__(kThreads)
add_broadcast_vec4_kernel(const float4* __restrict__ a, const float4* __restrict__ bias, float4* __restrict__ out,
                          size_t total4, int cols4, int relu) {
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total4; i += stride) {
        float4 v = __ldg(a + i);
        float4 b = __ldg(bias + (i % cols4));
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        out[i] = v;
    }
}

// ---- out[i,c] = a[i,c] - r[i] --------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
sub_rows_kernel(const float* __restrict__ a, const float* __restrict__ r, float* __restrict__ out, size_t total, int cols) {
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += stride)
        out[i] = a[i] - __ldg(r + (i / cols));
}

// ---- gin[i,c] (+)= g[i] | g[c] | g[0] ---------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
broadcast_bwd_kernel(const float* __restrict__ g, float* __restrict__ gin, size_t total, int cols, int mode, int accumulate) {
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += stride) {
        float v = mode == 0 ? __ldg(g + (i / cols)) : (mode == 1 ? __ldg(g + (i % cols)) : __ldg(g));
        gin[i] = accumulate ? gin[i] + v : v;
    }
}

// ---- column sums: stage 1 writes partial[split][cols]; stage 2 folds the splits --------------------
// Block = 32 columns x 8 row-lanes; rows of one split are strided over the 8 lanes (coalesced 128 B rows).
__global__ void __launch_bounds__(kThreads)
colsum_stage1(const float* __restrict__ g, float* __restrict__ partial, int rows, int cols, int rows_per_split) {
    __shared__ float sm[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + tx;
    const int r0 = blockIdx.y * rows_per_split;
    const int r1 = min(rows, r0 + rows_per_split);
    float acc = 0.0f;
    if (col < cols)
        for (int r = r0 + ty; r < r1; r += 8) acc += __ldg(g + (size_t)r * cols + col);
    sm[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && col < cols) {
        float s = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += sm[j][tx];
        partial[(size_t)blockIdx.y * cols + col] = s;
    }
}

// single-launch column sums: every (column block, row split) CTA parks its partial row, takes a ticket, and the last
// CTA of a column block folds the splits in order (deterministic) — no second kernel, no atomics on the data
__global__ void __launch_bounds__(kThreads)
colsum_fused_kernel(const float* __restrict__ g, float* __restrict__ partial, float* __restrict__ out, int* __restrict__ tickets,
                    int rows, int cols, int rows_per_split, float scale, int accumulate) {
    __shared__ float sm[8][33];
    __shared__ int s_last;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + tx;
    const int r0 = blockIdx.y * rows_per_split;
    const int r1 = min(rows, r0 + rows_per_split);
    float acc = 0.0f;
    if (col < cols)
        for (int r = r0 + ty; r < r1; r += 8) acc += __ldg(g + (size_t)r * cols + col);
    sm[ty][tx] = acc;
    __syncthreads();
    if (ty == 0) {
        float s = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += sm[j][tx];
        if (col < cols) partial[(size_t)blockIdx.y * cols + col] = s;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int prev = atomicAdd(tickets + blockIdx.x, 1);
        s_last = (prev == (int)gridDim.y - 1);
        if (s_last) tickets[blockIdx.x] = 0;          // ready for the next launch / graph replay
    }
    __syncthreads();
    if (s_last && ty == 0 && col < cols) {
        __threadfence();
        float s = 0.0f;
        for (int j0 = 0; j0 < (int)gridDim.y; j0 += 8) {       // split order, 8 loads in flight
            float q[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) q[u] = (j0 + u < (int)gridDim.y) ? __ldcg(partial + (size_t)(j0 + u) * cols + col) : 0.0f;
#pragma unroll
            for (int u = 0; u < 8; ++u) s += q[u];
        }
        s *= scale;
        out[col] = accumulate ? out[col] + s : s;
    }
}

__global__ void __launch_bounds__(kThreads)
fold_partials(const float* __restrict__ partial, float* __restrict__ out, int n, int splits, float scale, int accumulate) {
    int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    float s = 0.0f;
    for (int j0 = 0; j0 < splits; j0 += 8) {                   // split order, 8 loads in flight
        float q[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) q[u] = (j0 + u < splits) ? partial[(size_t)(j0 + u) * n + i] : 0.0f;
#pragma unroll
        for (int u = 0; u < 8; ++u) s += q[u];
    }
    s *= scale;
    out[i] = accumulate ? out[i] + s : s;
}

// ---- row sums: one warp per row -----------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
rowsum_kernel(const float* __restrict__ g, float* __restrict__ out, int rows, int cols, float scale, int accumulate) {
    const int warp = (blockIdx.x * kThreads + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * kThreads) >> 5;
    for (int r = warp; r < rows; r += nwarps) {
        float acc = 0.0f;
        for (int c = lane; c < cols; c += 32) acc += __ldg(g + (size_t)r * cols + c);
        acc = warp_sum(acc);
        if (lane == 0) {
            acc *= scale;
            out[r] = accumulate ? out[r] + acc : acc;
        }
    }
}

// ---- per-channel sums of g[n, c, hw]: block (c, split over n) ---------------------------------------------
__global__ void __launch_bounds__(kThreads)
bias_grad4d_stage1(const float* __restrict__ g, float* __restrict__ partial, int n, int c, int hw, int n_per_split) {
    __shared__ float sm[kThreads / 32];
    const int ch = blockIdx.x;
    const int n0 = blockIdx.y * n_per_split, n1 = min(n, n0 + n_per_split);
    float acc = 0.0f;
    if (hw >= kThreads) {
        for (int b = n0; b < n1; ++b) {
            const float* p = g + ((size_t)b * c + ch) * hw;
            for (int s = threadIdx.x; s < hw; s += kThreads) acc += __ldg(p + s);
        }
    } else {
        // small planes (7x7 after the last pool): flatten (image, pixel) so every thread has work
        const int span = (n1 - n0) * hw;
        for (int idx = threadIdx.x; idx < span; idx += kThreads) {
            const int b = idx / hw, sidx = idx - b * hw;
            acc += __ldg(g + ((size_t)(n0 + b) * c + ch) * hw + sidx);
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.0f;
#pragma unroll
        for (int j = 0; j < kThreads / 32; ++j) s += sm[j];
        partial[(size_t)blockIdx.y * c + ch] = s;
    }
}

// ---- y[n,c,hw] = x[n,c,hw] + bias[c] (optional relu) ----------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
add_bias4d_kernel(const float* __restrict__ x, const float* __restrict__ bias, float* __restrict__ y,
                  size_t total, int c, int hw, int relu) {
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += stride) {
        float v = x[i] + __ldg(bias + ((i / hw) % c));
        y[i] = relu ? fmaxf(v, 0.0f) : v;
    }
}

// ---- full sum: stage 1 per-block partials, stage 2 fold ----------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
sum_stage1(const float* __restrict__ x, float* __restrict__ partial, size_t n) {
    __shared__ float sm[kThreads / 32];
    float acc = 0.0f;
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) acc += __ldg(x + i);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.0f;
#pragma unroll
        for (int j = 0; j < kThreads / 32; ++j) s += sm[j];
        partial[blockIdx.x] = s;
    }
}

// ---- max over dim=1 (one warp per row), first max wins, NaN never selected ---------------------------------------
__global__ void __launch_bounds__(kThreads)
max_rows_kernel(const float* __restrict__ x, float* __restrict__ vals, float* __restrict__ idx, int rows, int cols) {
    const int warp = (blockIdx.x * kThreads + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * kThreads) >> 5;
    for (int r = warp; r < rows; r += nwarps) {
        float best = -INFINITY;
        int bi = INT_MAX;
        for (int c = lane; c < cols; c += 32) {
            float v = __ldg(x + (size_t)r * cols + c);
            if (v > best) { best = v; bi = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) {
            if (vals) vals[r] = best;
            if (idx) idx[r] = (bi == INT_MAX) ? 0.0f : (float)bi;
        }
    }
}

// ---- max over dim=0 (one thread per column) ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
max_cols_kernel(const float* __restrict__ x, float* __restrict__ vals, float* __restrict__ idx, int rows, int cols) {
    int c = blockIdx.x * kThreads + threadIdx.x;
    if (c >= cols) return;
    float best = -INFINITY;
    int bi = 0;
    for (int r = 0; r < rows; ++r) {
        float v = __ldg(x + (size_t)r * cols + c);
        if (v > best) { best = v; bi = r; }
    }
    if (vals) vals[c] = best;
    if (idx) idx[c] = (float)bi;
}

// ---- global max, last of equal maxima (Iterator::max_by) ---------------------------------------------------------------
__global__ void __launch_bounds__(1024)
max_all_kernel(const float* __restrict__ x, float* __restrict__ val, float* __restrict__ idx, size_t n) {
    __shared__ float sv[32];
    __shared__ long long si[32];
    float best = -INFINITY;
    long long bi = -1;
    for (size_t i = threadIdx.x; i < n; i += 1024) {
        float v = __ldg(x + i);
        if (v >= best) { best = v; bi = (long long)i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi > bi)) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int j = 1; j < 32; ++j)
            if (sv[j] > best || (sv[j] == best && si[j] > bi)) { best = sv[j]; bi = si[j]; }
        if (n == 0) { best = 0.0f; bi = 0; }
        val[0] = best;
        idx[0] = (float)(bi < 0 ? 0 : bi);
    }
}

// ---- y[j,i] (+)= x[i,j], 32x32 tiles through padded shared memory ----------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
transpose_kernel(const float* __restrict__ x, float* __restrict__ y, int rows, int cols, int accumulate) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        int r = r0 + ty + j, c = c0 + tx;
        if (r < rows && c < cols) tile[ty + j][tx] = __ldg(x + (size_t)r * cols + c);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        int orow = c0 + ty + j, ocol = r0 + tx;               // y is [cols, rows]
        if (orow < cols && ocol < rows) {
            size_t o = (size_t)orow * rows + ocol;
            float v = tile[tx][ty + j];
            y[o] = accumulate ? y[o] + v : v;
        }
    }
}

// ---- batched 2-D transpose: y[b, j, i] = x[b, i, j]  (NHWC<->NCHW with rows=H*W, cols=C or vice versa) ------------------------
__global__ void __launch_bounds__(kThreads)
transpose_batched_kernel(const float* __restrict__ x, float* __restrict__ y, int rows, int cols) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const size_t off = (size_t)blockIdx.z * rows * cols;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        int r = r0 + ty + j, c = c0 + tx;
        if (r < rows && c < cols) tile[ty + j][tx] = __ldg(x + off + (size_t)r * cols + c);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        int orow = c0 + ty + j, ocol = r0 + tx;
        if (orow < cols && ocol < rows) y[off + (size_t)orow * rows + ocol] = tile[tx][ty + j];
    }
}

int pick_splits(const tp_ctx* ctx, int blocks_x, int rows, int min_rows_per_split) {
    int target = (2 * ctx->sm_count + blocks_x - 1) / blocks_x;
    int max_splits = (rows + min_rows_per_split - 1) / min_rows_per_split;
    int s = target < max_splits ? target : max_splits;
    return s < 1 ? 1 : s;
}

}  // namespace

extern "C" {

int tp_add_broadcast_fwd(tp_ctx* ctx, const tp_buf* a, const tp_buf* bias, tp_buf* out, int rows, int cols, int relu) {
    TP_CHECK_ARG(ctx && rows >= 0 && cols > 0, "tp_add_broadcast_fwd: bad dims %d x %d", rows, cols);
    size_t total = (size_t)rows * cols;
    TP_NEED(a, total, "a"); TP_NEED(bias, cols, "bias"); TP_NEED(out, total, "out");
    if (!total) return TP_OK;
    bool vec = cols % 4 == 0 && !(((uintptr_t)a->ptr | (uintptr_t)bias->ptr | (uintptr_t)out->ptr) & 15);
    if (vec)
        add_broadcast_vec4_kernel<<<tp::grid_for(ctx, total / 4, kThreads), kThreads, 0, ctx->stream>>>(
            (const float4*)a->ptr, (const float4*)bias->ptr, (float4*)out->ptr, total / 4, cols / 4, relu);
    else
        add_broadcast_kernel<<<tp::grid_for(ctx, total, kThreads), kThreads, 0, ctx->stream>>>(a->ptr, bias->ptr, out->ptr, total, cols, relu);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

int tp_colsum(tp_ctx* ctx, const tp_buf* g, tp_buf* out, int rows, int cols, float scale, int accumulate) {
    TP_CHECK_ARG(ctx && rows >= 0 && cols > 0, "tp_colsum: bad dims %d x %d", rows, cols);
    TP_NEED(g, (size_t)rows * cols, "g"); TP_NEED(out, cols, "out");
    int bx = (cols + 31) / 32;
    int splits = pick_splits(ctx, bx, rows, 64);
    int rps = (rows + splits - 1) / splits;
    if (rps < 1) rps = 1;
    splits = rows > 0 ? (rows + rps - 1) / rps : 1;
    int rc = tp::ensure_scratch(ctx, (size_t)splits * cols * sizeof(float));
    if (rc) return rc;
    if (bx <= tp::kCounterGemm - tp::kCounterColsum) {
        colsum_fused_kernel<<<dim3(bx, splits), kThreads, 0, ctx->stream>>>(g->ptr, ctx->scratch, out->ptr,
                                                                           ctx->dev_counters + tp::kCounterColsum, rows, cols, rps,
                                                                           scale, accumulate);
        TP_LAUNCH_OK(ctx);
        return TP_OK;
    }
    colsum_stage1<<<dim3(bx, splits), kThreads, 0, ctx->stream>>>(g->ptr, ctx->scratch, rows, cols, rps);
    TP_LAUNCH_OK(ctx);
    fold_partials<<<(cols + kThreads - 1) / kThreads, kThreads, 0, ctx->stream>>>(ctx->scratch, out->ptr, cols, splits, scale, accumulate);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

int tp_rowsum(tp_ctx* ctx, const tp_buf* g, tp_buf* out, int rows, int cols, float scale, int accumulate) {
    TP_CHECK_ARG(ctx && rows >= 0 && cols > 0, "tp_rowsum: bad dims %d x %d", rows, cols);
    TP_NEED(g, (size_t)rows * cols, "g"); TP_NEED(out, rows, "out");
    if (!rows) return TP_OK;
    rowsum_kernel<<<tp::grid_for(ctx, (size_t)rows * 32, kThreads), kThreads, 0, ctx->stream>>>(g->ptr, out->ptr, rows, cols, scale, accumulate);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

int tp_sum_all(tp_ctx* ctx, const tp_buf* x, tp_buf* out, size_t n) {
    TP_CHECK_ARG(ctx, "tp_sum_all: NULL ctx");
    TP_NEED(x, n, "x"); TP_NEED(out, 1, "out");
    int blocks = tp::grid_for(ctx, n, kThreads * 4, 2);
    int rc = tp::ensure_scratch(ctx, (size_t)blocks * sizeof(float));
    if (rc) return rc;
    sum_stage1<<<blocks, kThreads, 0, ctx->stream>>>(x->ptr, ctx->scratch, n);
    TP_LAUNCH_OK(ctx);
    fold_partials<<<1, kThreads, 0, ctx->stream>>>(ctx->scratch, out->ptr, 1, blocks, 1.0f, 0);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

int tp_sub_broadcast_rows_fwd(tp_ctx* ctx, const tp_buf* a, const tp_buf* r, tp_buf* out, int rows, int cols) {
    TP_CHECK_ARG(ctx && rows >= 0 && cols > 0, "tp_sub_broadcast_rows_fwd: bad dims %d x %d", rows, cols);
    size_t total = (size_t)rows * cols;
    TP_NEED(a, total, "a"); TP_NEED(r, rows, "r"); TP_NEED(out, total, "out");
    if (!total) return TP_OK;
    sub_rows_kernel<<<tp::grid_for(ctx, total, kThreads), kThreads, 0, ctx->stream>>>(a->ptr, r->ptr, out->ptr, total, cols);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

int tp_broadcast_bwd(tp_ctx* ctx, const tp_buf* g, tp_buf* gin, int rows, int cols, int mode, int accumulate) {
    TP_CHECK_ARG(ctx && rows >= 0 && cols > 0 && mode >= 0 && mode <= 2, "tp_broadcast_bwd: bad arguments");
    size_t total = (size_t)rows * cols;
    TP_NEED(g, mode == 0 ? rows : (mode == 1 ? cols : 1), "g"); TP_NEED(gin, total, "gin");
    if (!total) return TP_OK;
    broadcast_bwd_kernel<<<tp::grid_for(ctx, total, kThreads), kThreads, 0, ctx->stream>>>(g->ptr, gin->ptr, total, cols, mode, accumulate);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

int tp_max_rows(tp_ctx* ctx, const tp_buf* x, tp_buf* vals, tp_buf* idx, int rows, int cols) {
    TP_CHECK_ARG(ctx && rows >= 0 && cols > 0, "tp_max_rows: bad dims %d x %d", rows, cols);
    TP_NEED(x, (size_t)rows * cols, "x");
    if (vals) TP_NEED(vals, rows, "vals");
    if (idx) TP_NEED(idx, rows, "idx");
    if (!rows) return TP_OK;
    max_rows_kernel<<<tp::grid_for(ctx, (size_t)rows * 32, kThreads), kThreads, 0, ctx->stream>>>(
        x->ptr, vals ? vals->ptr : nullptr, idx ? idx->ptr : nullptr, rows, cols);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

int tp_max_cols(tp_ctx* ctx, const tp_buf* x, tp_buf* vals, tp_buf* idx, int rows, int cols) {
    TP_CHECK_ARG(ctx && rows >= 0 && cols > 0, "tp_max_cols: bad dims %d x %d", rows, cols);
    TP_NEED(x, (size_t)rows * cols, "x");
    if (vals) TP_NEED(vals, cols, "vals");
    if (idx) TP_NEED(idx, cols, "idx");
    max_cols_kernel<<<(cols + kThreads - 1) / kThreads, kThreads, 0, ctx->stream>>>(
        x->ptr, vals ? vals->ptr : nullptr, idx ? idx->ptr : nullptr, rows, cols);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

int tp_max_all(tp_ctx* ctx, const tp_buf* x, tp_buf* val, tp_buf* idx, size_t n) {
    TP_CHECK_ARG(ctx, "tp_max_all: NULL ctx");
    TP_NEED(x, n, "x"); TP_NEED(val, 1, "val"); TP_NEED(idx, 1, "idx");
    max_all_kernel<<<1, 1024, 0, ctx->stream>>>(x->ptr, val->ptr, idx->ptr, n);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

int tp_transpose2d(tp_ctx* ctx, const tp_buf* x, tp_buf* y, int rows, int cols, int accumulate) {
    TP_CHECK_ARG(ctx && rows >= 0 && cols >= 0, "tp_transpose2d: bad dims %d x %d", rows, cols);
    size_t total = (size_t)rows * cols;
    TP_NEED(x, total, "x"); TP_NEED(y, total, "y");
    if (!total) return TP_OK;
    transpose_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32), kThreads, 0, ctx->stream>>>(x->ptr, y->ptr, rows, cols, accumulate);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

int tp_add_bias_4d(tp_ctx* ctx, const tp_buf* x, const tp_buf* bias, tp_buf* y, int n, int c, int hw, int relu) {
    TP_CHECK_ARG(ctx && n >= 0 && c > 0 && hw > 0, "tp_add_bias_4d: bad dims");
    size_t total = (size_t)n * c * hw;
    TP_NEED(x, total, "x"); TP_NEED(bias, c, "bias"); TP_NEED(y, total, "y");
    if (!total) return TP_OK;
    add_bias4d_kernel<<<tp::grid_for(ctx, total, kThreads), kThreads, 0, ctx->stream>>>(x->ptr, bias->ptr, y->ptr, total, c, hw, relu);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

int tp_bias_grad_4d(tp_ctx* ctx, const tp_buf* g, tp_buf* gb, int n, int c, int hw, int accumulate) {
    TP_CHECK_ARG(ctx && n >= 0 && c > 0 && hw > 0, "tp_bias_grad_4d: bad dims");
    TP_NEED(g, (size_t)n * c * hw, "g"); TP_NEED(gb, c, "gb");
    int splits = pick_splits(ctx, c, n, 4);
    int nps = (n + splits - 1) / splits;
    if (nps < 1) nps = 1;
    int rc = tp::ensure_scratch(ctx, (size_t)splits * c * sizeof(float));
    if (rc) return rc;
    bias_grad4d_stage1<<<dim3(c, splits), kThreads, 0, ctx->stream>>>(g->ptr, ctx->scratch, n, c, hw, nps);
    TP_LAUNCH_OK(ctx);
    fold_partials<<<(c + kThreads - 1) / kThreads, kThreads, 0, ctx->stream>>>(ctx->scratch, gb->ptr, c, splits, 1.0f, accumulate);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

int tp_nhwc_to_nchw(tp_ctx* ctx, const tp_buf* x, tp_buf* y, int n, int h, int w, int c) {
    TP_CHECK_ARG(ctx && n >= 0 && h > 0 && w > 0 && c > 0, "tp_nhwc_to_nchw: bad dims");
    size_t total = (size_t)n * h * w * c;
    TP_NEED(x, total, "x"); TP_NEED(y, total, "y");
    if (!total) return TP_OK;
    TP_CHECK_ARG(n <= 65535, "tp_nhwc_to_nchw: batch %d exceeds grid.z limit", n);
    int rows = h * w, cols = c;
    transpose_batched_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32, n), kThreads, 0, ctx->stream>>>(x->ptr, y->ptr, rows, cols);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

int tp_nchw_to_nhwc(tp_ctx* ctx, const tp_buf* x, tp_buf* y, int n, int c, int h, int w) {
    TP_CHECK_ARG(ctx && n >= 0 && h > 0 && w > 0 && c > 0, "tp_nchw_to_nhwc: bad dims");
    size_t total = (size_t)n * h * w * c;
    TP_NEED(x, total, "x"); TP_NEED(y, total, "y");
    if (!total) return TP_OK;
    TP_CHECK_ARG(n <= 65535, "tp_nchw_to_nhwc: batch %d exceeds grid.z limit", n);
    int rows = c, cols = h * w;
    transpose_batched_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32, n), kThreads, 0, ctx->stream>>>(x->ptr, y->ptr, rows, cols);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

}  // extern "C"

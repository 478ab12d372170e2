// max_pool2d / avg_pool2d forward and backward, NCHW fp32.
// Reference: src/tensor.rs:1391-1521 (max: strict '>' from -inf scanning kh then kw, absolute argmax index,
// backward zeroes each input plane then scatter-adds — overwrite semantics) and
// src/tensor.rs:1524-1660 (avg: divisor kh*kw including padding, backward accumulates).
// Backward kernels are gather-form: each input element sums the output windows that cover it, in the
// reference's ascending (oh, ow) order — deterministic, no atomics, no separate memset pass.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

struct PoolGeom {
    int n, c, h, w, kh, kw, sh, sw, ph, pw, ho, wo;
};

__global__ void __launch_bounds__(kThreads)
maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int* __restrict__ arg, PoolGeom g, size_t total) {
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += stride) {
        // total <= 2^31 (make_geom): 32-bit index arithmetic
        const unsigned int iu = (unsigned int)i;
        const unsigned int tq = iu / (unsigned int)g.wo;
        int ow = (int)(iu - tq * (unsigned int)g.wo);
        const unsigned int pl = tq / (unsigned int)g.ho;      // n*C + c
        int oh = (int)(tq - pl * (unsigned int)g.ho);
        size_t base = (size_t)pl * g.h * g.w;
        float best = -INFINITY;
        size_t bi = base;                              // "any valid default" (src/tensor.rs:1432)
        for (int kr = 0; kr < g.kh; ++kr) {
            int ih = oh * g.sh + kr - g.ph;
            if (ih < 0 || ih >= g.h) continue;
            for (int kc = 0; kc < g.kw; ++kc) {
                int iw = ow * g.sw + kc - g.pw;
                if (iw < 0 || iw >= g.w) continue;
                size_t idx = base + (size_t)ih * g.w + iw;
                float v = __ldg(x + idx);
                if (v > best) { best = v; bi = idx; }
            }
        }
        y[i] = best;
        arg[i] = (int)bi;
    }
}

__global__ void __launch_bounds__(kThreads)
maxpool_bwd_gather_kernel(const float* __restrict__ gout, const int* __restrict__ arg, float* __restrict__ gin,
                          PoolGeom g, size_t total) {
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += stride) {
        // total <= 2^31 (make_geom): 32-bit index arithmetic
        const unsigned int iu = (unsigned int)i;
        const unsigned int tq = iu / (unsigned int)g.w;
        int iw = (int)(iu - tq * (unsigned int)g.w);
        const unsigned int plane = tq / (unsigned int)g.h;
        int ih = (int)(tq - plane * (unsigned int)g.h);
        int oh_lo = (ih + g.ph - g.kh + 1 + g.sh - 1);
        oh_lo = oh_lo <= 0 ? 0 : oh_lo / g.sh;
        int oh_hi = min(g.ho - 1, (ih + g.ph) / g.sh);
        int ow_lo = (iw + g.pw - g.kw + 1 + g.sw - 1);
        ow_lo = ow_lo <= 0 ? 0 : ow_lo / g.sw;
        int ow_hi = min(g.wo - 1, (iw + g.pw) / g.sw);
        float acc = 0.0f;
        for (int oh = oh_lo; oh <= oh_hi; ++oh)
            for (int ow = ow_lo; ow <= ow_hi; ++ow) {
                size_t o = (plane * g.ho + oh) * g.wo + ow;
                if (__ldg(arg + o) == (int)i) acc += __ldg(gout + o);
            }
        gin[i] = acc;                                  // plane zeroed, then accumulated (A6): overwrite
    }
}

// degenerate geometry (a window can lie entirely in the padding): literal zero + scatter
__global__ void __launch_bounds__(kThreads)
maxpool_bwd_scatter_kernel(const float* __restrict__ gout, const int* __restrict__ arg, float* __restrict__ gin, size_t total_out) {
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total_out; i += stride)
        atomicAdd(gin + __ldg(arg + i), __ldg(gout + i));
}

__global__ void __launch_bounds__(kThreads)
avgpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, PoolGeom g, size_t total) {
    const size_t stride = (size_t)gridDim.x * kThreads;
    const float pool = (float)(g.kh * g.kw);
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += stride) {
        int ow = (int)(i % g.wo);
        size_t t = i / g.wo;
        int oh = (int)(t % g.ho);
        size_t base = (t / g.ho) * g.h * g.w;
        float s = 0.0f;
        for (int kr = 0; kr < g.kh; ++kr) {
            int ih = oh * g.sh + kr - g.ph;
            if (ih < 0 || ih >= g.h) continue;
            for (int kc = 0; kc < g.kw; ++kc) {
                int iw = ow * g.sw + kc - g.pw;
                if (iw < 0 || iw >= g.w) continue;
                s += __ldg(x + base + (size_t)ih * g.w + iw);
            }
        }
        y[i] = s / pool;
    }
}

// global-average fast path (kernel == whole plane): one warp per (n,c) plane, coalesced reads
__global__ void __launch_bounds__(kThreads)
avgpool_global_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int planes, int hw) {
    const int warp = (blockIdx.x * kThreads + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * kThreads) >> 5;
    for (int p = warp; p < planes; p += nwarps) {
        float s = 0.0f;
        for (int i = lane; i < hw; i += 32) s += __ldg(x + (size_t)p * hw + i);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) y[p] = s / (float)hw;
    }
}

__global__ void __launch_bounds__(kThreads)
avgpool_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gin, PoolGeom g, size_t total, int accumulate) {
    const size_t stride = (size_t)gridDim.x * kThreads;
    const float pool = (float)(g.kh * g.kw);
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += stride) {
        // total <= 2^31 (make_geom): 32-bit index arithmetic
        const unsigned int iu = (unsigned int)i;
        const unsigned int tq = iu / (unsigned int)g.w;
        int iw = (int)(iu - tq * (unsigned int)g.w);
        const unsigned int plane = tq / (unsigned int)g.h;
        int ih = (int)(tq - plane * (unsigned int)g.h);
        int oh_lo = (ih + g.ph - g.kh + 1 + g.sh - 1);
        oh_lo = oh_lo <= 0 ? 0 : oh_lo / g.sh;
        int oh_hi = min(g.ho - 1, (ih + g.ph) / g.sh);
        int ow_lo = (iw + g.pw - g.kw + 1 + g.sw - 1);
        ow_lo = ow_lo <= 0 ? 0 : ow_lo / g.sw;
        int ow_hi = min(g.wo - 1, (iw + g.pw) / g.sw);
        float acc = 0.0f;
        for (int oh = oh_lo; oh <= oh_hi; ++oh)
            for (int ow = ow_lo; ow <= ow_hi; ++ow)
                acc += __ldg(gout + ((size_t)plane * g.ho + oh) * g.wo + ow) / pool;     // g/pool_size per tap (src/tensor.rs:1628)
        gin[i] = accumulate ? gin[i] + acc : acc;
    }
}

int make_geom(const tp_pool_desc* d, PoolGeom* g, const char* fn) {
    TP_CHECK_ARG(d, "%s: NULL descriptor", fn);
    TP_CHECK_ARG(d->n >= 0 && d->c > 0 && d->h > 0 && d->w > 0 && d->kh > 0 && d->kw > 0 && d->stride_h > 0 &&
                     d->stride_w > 0 && d->pad_h >= 0 && d->pad_w >= 0,
                 "%s: invalid pooling descriptor", fn);
    TP_CHECK_ARG(d->h + 2 * d->pad_h >= d->kh && d->w + 2 * d->pad_w >= d->kw, "%s: window larger than padded input", fn);
    g->n = d->n; g->c = d->c; g->h = d->h; g->w = d->w; g->kh = d->kh; g->kw = d->kw;
    g->sh = d->stride_h; g->sw = d->stride_w; g->ph = d->pad_h; g->pw = d->pad_w;
    g->ho = (d->h + 2 * d->pad_h - d->kh) / d->stride_h + 1;
    g->wo = (d->w + 2 * d->pad_w - d->kw) / d->stride_w + 1;
    TP_CHECK_ARG((size_t)d->n * d->c * d->h * d->w <= 0x7fffffffULL, "%s: input too large for int32 argmax indices", fn);
    return TP_OK;
}

}  // namespace

extern "C" {

int tp_pool_out_dims(const tp_pool_desc* d, int* h_out, int* w_out) {
    PoolGeom g;
    int rc = make_geom(d, &g, "tp_pool_out_dims");
    if (rc) return rc;
    if (h_out) *h_out = g.ho;
    if (w_out) *w_out = g.wo;
    return TP_OK;
}

int tp_maxpool2d_fwd(tp_ctx* ctx, const tp_buf* x, tp_buf* y, tp_buf* argmax_i32, const tp_pool_desc* d) {
    TP_CHECK_ARG(ctx, "tp_maxpool2d_fwd: NULL ctx");
    PoolGeom g;
    int rc = make_geom(d, &g, "tp_maxpool2d_fwd");
    if (rc) return rc;
    size_t total = (size_t)g.n * g.c * g.ho * g.wo;
    TP_NEED(x, (size_t)g.n * g.c * g.h * g.w, "x"); TP_NEED(y, total, "y"); TP_NEED(argmax_i32, total, "argmax");
    if (!total) return TP_OK;
    maxpool_fwd_kernel<<<tp::grid_for(ctx, total, kThreads), kThreads, 0, ctx->stream>>>(x->ptr, y->ptr, (int*)argmax_i32->ptr, g, total);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

int tp_maxpool2d_bwd(tp_ctx* ctx, const tp_buf* gout, const tp_buf* argmax_i32, tp_buf* gin, const tp_pool_desc* d) {
    TP_CHECK_ARG(ctx, "tp_maxpool2d_bwd: NULL ctx");
    PoolGeom g;
    int rc = make_geom(d, &g, "tp_maxpool2d_bwd");
    if (rc) return rc;
    size_t total_out = (size_t)g.n * g.c * g.ho * g.wo, total_in = (size_t)g.n * g.c * g.h * g.w;
    TP_NEED(gout, total_out, "gout"); TP_NEED(argmax_i32, total_out, "argmax"); TP_NEED(gin, total_in, "gin");
    if (!total_in) return TP_OK;
    if (g.ph >= g.kh || g.pw >= g.kw) {
        TP_CUDA(cudaMemsetAsync(gin->ptr, 0, total_in * sizeof(float), ctx->stream));
        if (total_out) {
            maxpool_bwd_scatter_kernel<<<tp::grid_for(ctx, total_out, kThreads), kThreads, 0, ctx->stream>>>(
                gout->ptr, (const int*)argmax_i32->ptr, gin->ptr, total_out);
            TP_LAUNCH_OK(ctx);
        }
        return TP_OK;
    }
    maxpool_bwd_gather_kernel<<<tp::grid_for(ctx, total_in, kThreads), kThreads, 0, ctx->stream>>>(
        gout->ptr, (const int*)argmax_i32->ptr, gin->ptr, g, total_in);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

int tp_avgpool2d_fwd(tp_ctx* ctx, const tp_buf* x, tp_buf* y, const tp_pool_desc* d) {
    TP_CHECK_ARG(ctx, "tp_avgpool2d_fwd: NULL ctx");
    PoolGeom g;
    int rc = make_geom(d, &g, "tp_avgpool2d_fwd");
    if (rc) return rc;
    size_t total = (size_t)g.n * g.c * g.ho * g.wo;
    TP_NEED(x, (size_t)g.n * g.c * g.h * g.w, "x"); TP_NEED(y, total, "y");
    if (!total) return TP_OK;
    if (g.kh == g.h && g.kw == g.w && g.ph == 0 && g.pw == 0) {
        int planes = g.n * g.c;
        avgpool_global_fwd_kernel<<<tp::grid_for(ctx, (size_t)planes * 32, kThreads), kThreads, 0, ctx->stream>>>(x->ptr, y->ptr, planes, g.h * g.w);
    } else {
        avgpool_fwd_kernel<<<tp::grid_for(ctx, total, kThreads), kThreads, 0, ctx->stream>>>(x->ptr, y->ptr, g, total);
    }
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

int tp_avgpool2d_bwd(tp_ctx* ctx, const tp_buf* gout, tp_buf* gin, const tp_pool_desc* d, int accumulate) {
    TP_CHECK_ARG(ctx, "tp_avgpool2d_bwd: NULL ctx");
    PoolGeom g;
    int rc = make_geom(d, &g, "tp_avgpool2d_bwd");
    if (rc) return rc;
    size_t total_in = (size_t)g.n * g.c * g.h * g.w;
    TP_NEED(gout, (size_t)g.n * g.c * g.ho * g.wo, "gout"); TP_NEED(gin, total_in, "gin");
    if (!total_in) return TP_OK;
    avgpool_bwd_kernel<<<tp::grid_for(ctx, total_in, kThreads), kThreads, 0, ctx->stream>>>(gout->ptr, gin->ptr, g, total_in, accumulate);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

}  // extern "C"

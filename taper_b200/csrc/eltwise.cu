// Elementwise family: HBM-bound, 128-bit vectorised, grid sized in multiples of the SM count.
// Reference: `pub mod simd` (src/tensor.rs:14-234), operators and their backward closures
// (src/ops.rs:8-151, 312-496), exp/log (src/tensor.rs:1091-1169).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;   // 4 independent 128-bit loads in flight per thread per input

struct OpAdd { __device__ float operator()(float a, float b, float) const { return a + b; } };
struct OpSub { __device__ float operator()(float a, float b, float) const { return a - b; } };
struct OpMul { __device__ float operator()(float a, float b, float) const { return a * b; } };
struct OpDiv { __device__ float operator()(float a, float b, float) const { return a / b; } };
struct OpScale { float s; __device__ float operator()(float a, float, float) const { return s * a; } };
struct OpRelu { __device__ float operator()(float a, float, float) const { return fmaxf(a, 0.0f); } };
// relu bwd: a = x, b = gout
struct OpReluBwd { __device__ float operator()(float x, float g, float) const { return x > 0.0f ? g : 0.0f; } };
struct OpExp { __device__ float operator()(float a, float, float) const { return expf(a); } };
struct OpLog { __device__ float operator()(float a, float, float) const { return logf(a); } };
// div bwd wrt b: a = gout, b = a, c = b  ->  -gout*a/(b*b)
struct OpDivBwdB { __device__ float operator()(float g, float a, float b) const { return -(g * a / (b * b)); } };
struct OpFill { float v; __device__ float operator()(float, float, float) const { return v; } };
// ---- sigmoid / pow / mean / binary cross-entropy (XOR demo, src/main.rs) -------------------------------
// sigmoid, the reference's two-branch form (src/tensor.rs:601-608)
struct OpSigmoid {
    __device__ float operator()(float x, float, float) const {
        if (x > 0.0f) { const float e = expf(-x); return 1.0f / (1.0f + e); }
        const float e = expf(x);
        return e / (1.0f + e);
    }
};
struct OpSigmoidBwd { __device__ float operator()(float g, float s, float) const { return g * s * (1.0f - s); } };      // :627
struct OpPow { float e; __device__ float operator()(float x, float, float) const { return powf(x, e); } };               // :1177
struct OpPowBwd { float e; __device__ float operator()(float g, float x, float) const { return g * e * powf(x, e - 1.0f); } };   // :1199
struct OpDivBy { float d; __device__ float operator()(float a, float, float) const { return a / d; } };
// bce term -(y ln p + (1-y) ln(1-p)) with p clamped to [1e-7, 1-1e-7]  (src/loss.rs:7, 19-21)
struct OpBceTerm {
    __device__ float operator()(float p, float y, float) const {
        const float pi = fminf(fmaxf(p, 1e-7f), 1.0f - 1e-7f);
        return -(y * logf(pi) + (1.0f - y) * logf(1.0f - pi));
    }
};
// g * (-(y/p - (1-y)/(1-p))) / n  with the upstream scalar g read from the device  (src/loss.rs:52)
struct OpBceBwdP {
    const float* g; float n;
    __device__ float operator()(float p, float y, float) const {
        const float pi = fminf(fmaxf(p, 1e-7f), 1.0f - 1e-7f);
        return __ldg(g) * (-(y / pi - (1.0f - y) / (1.0f - pi))) / n;
    }
};
// g * (ln(1-p) - ln p) / n  (src/loss.rs:66)
struct OpBceBwdT {
    const float* g; float n;
    __device__ float operator()(float p, float, float) const {
        const float pi = fminf(fmaxf(p, 1e-7f), 1.0f - 1e-7f);
        return __ldg(g) * (logf(1.0f - pi) - logf(pi)) / n;
    }
};
// gin (+)= g[0] / n for every element  (mean backward, src/tensor.rs:786-795)
struct OpMeanBwd { const float* g; float n; __device__ float operator()(float, float, float) const { return __ldg(g) / n; } };

template <int NIN, class Op>
__global__ void __launch_bounds__(kThreads)
ew_vec4(Op op, const float4* __restrict__ a, const float4* __restrict__ b, const float4* __restrict__ c,
        float4* __restrict__ out, size_t n4, int accumulate) {
    const size_t stride = (size_t)gridDim.x * kThreads * kUnroll;
    for (size_t base = (size_t)blockIdx.x * kThreads * kUnroll + threadIdx.x; base < n4; base += stride) {
        float4 va[kUnroll], vb[kUnroll], vc[kUnroll], vo[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            size_t i = base + (size_t)u * kThreads;
            if (i < n4) {
                if (NIN >= 1) va[u] = __ldg(a + i);
                if (NIN >= 2) vb[u] = __ldg(b + i);
                if (NIN >= 3) vc[u] = __ldg(c + i);
                if (accumulate) vo[u] = out[i];
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            size_t i = base + (size_t)u * kThreads;
            if (i < n4) {
                float4 r;
                r.x = op(NIN >= 1 ? va[u].x : 0.f, NIN >= 2 ? vb[u].x : 0.f, NIN >= 3 ? vc[u].x : 0.f);
                r.y = op(NIN >= 1 ? va[u].y : 0.f, NIN >= 2 ? vb[u].y : 0.f, NIN >= 3 ? vc[u].y : 0.f);
                r.z = op(NIN >= 1 ? va[u].z : 0.f, NIN >= 2 ? vb[u].z : 0.f, NIN >= 3 ? vc[u].z : 0.f);
                r.w = op(NIN >= 1 ? va[u].w : 0.f, NIN >= 2 ? vb[u].w : 0.f, NIN >= 3 ? vc[u].w : 0.f);
                if (accumulate) { r.x += vo[u].x; r.y += vo[u].y; r.z += vo[u].z; r.w += vo[u].w; }
                out[i] = r;
            }
        }
    }
}

template <int NIN, class Op>
__global__ void __launch_bounds__(kThreads)
ew_scalar(Op op, const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
          float* __restrict__ out, size_t n, int accumulate) {
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
        float r = op(NIN >= 1 ? a[i] : 0.f, NIN >= 2 ? b[i] : 0.f, NIN >= 3 ? c[i] : 0.f);
        out[i] = accumulate ? out[i] + r : r;
    }
}

// p[i] *= s in place (AdamW decoupled decay of grad-less params, src/optim.rs:154-161)
__global__ void __launch_bounds__(kThreads) scale_inplace(float* p, float s, size_t n) {
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) p[i] *= s;
}

inline bool aligned16(const void* p) { return p == nullptr || (((uintptr_t)p) & 15) == 0; }

template <int NIN, class Op>
int launch_ew(tp_ctx* ctx, Op op, const float* a, const float* b, const float* c, float* out, size_t n, int accumulate) {
    if (n == 0) return TP_OK;
    cudaSetDevice(ctx->device);
    size_t n4 = n / 4;
    bool vec = n4 > 0 && aligned16(a) && aligned16(b) && aligned16(c) && aligned16(out);
    if (vec) {
        int grid = tp::grid_for(ctx, n4, kThreads * kUnroll);
        ew_vec4<NIN, Op><<<grid, kThreads, 0, ctx->stream>>>(op, (const float4*)a, (const float4*)b, (const float4*)c,
                                                           (float4*)out, n4, accumulate);
        TP_LAUNCH_OK(ctx);
        size_t done = n4 * 4;
        if (done < n) {
            ew_scalar<NIN, Op><<<1, kThreads, 0, ctx->stream>>>(op, a ? a + done : a, b ? b + done : b, c ? c + done : c,
                                                               out + done, n - done, accumulate);
            TP_LAUNCH_OK(ctx);
        }
    } else {
        int grid = tp::grid_for(ctx, n, kThreads);
        ew_scalar<NIN, Op><<<grid, kThreads, 0, ctx->stream>>>(op, a, b, c, out, n, accumulate);
        TP_LAUNCH_OK(ctx);
    }
    return TP_OK;
}

}  // namespace

extern "C" {

#define TP_BINARY(name, OP)                                                                          \
    int name(tp_ctx* ctx, const tp_buf* a, const tp_buf* b, tp_buf* out, size_t n) {                 \
        TP_CHECK_ARG(ctx, #name ": NULL ctx");                                                       \
        TP_NEED(a, n, "a"); TP_NEED(b, n, "b"); TP_NEED(out, n, "out");                              \
        return launch_ew<2>(ctx, OP(), a->ptr, b->ptr, nullptr, out->ptr, n, 0);                     \
    }
TP_BINARY(tp_add, OpAdd)
TP_BINARY(tp_sub, OpSub)
TP_BINARY(tp_mul, OpMul)
TP_BINARY(tp_div, OpDiv)

int tp_accumulate(tp_ctx* ctx, tp_buf* dst, const tp_buf* src, float scale, size_t n, int accumulate) {
    TP_CHECK_ARG(ctx, "tp_accumulate: NULL ctx");
    TP_NEED(dst, n, "dst"); TP_NEED(src, n, "src");
    return launch_ew<1>(ctx, OpScale{scale}, src->ptr, nullptr, nullptr, dst->ptr, n, accumulate);
}

int tp_mul_bwd(tp_ctx* ctx, const tp_buf* gout, const tp_buf* other, tp_buf* gdst, size_t n, int accumulate) {
    TP_CHECK_ARG(ctx, "tp_mul_bwd: NULL ctx");
    TP_NEED(gout, n, "gout"); TP_NEED(other, n, "other"); TP_NEED(gdst, n, "gdst");
    return launch_ew<2>(ctx, OpMul(), gout->ptr, other->ptr, nullptr, gdst->ptr, n, accumulate);
}

int tp_div_bwd_a(tp_ctx* ctx, const tp_buf* gout, const tp_buf* b, tp_buf* ga, size_t n, int accumulate) {
    TP_CHECK_ARG(ctx, "tp_div_bwd_a: NULL ctx");
    TP_NEED(gout, n, "gout"); TP_NEED(b, n, "b"); TP_NEED(ga, n, "ga");
    return launch_ew<2>(ctx, OpDiv(), gout->ptr, b->ptr, nullptr, ga->ptr, n, accumulate);
}

int tp_div_bwd_b(tp_ctx* ctx, const tp_buf* gout, const tp_buf* a, const tp_buf* b, tp_buf* gb, size_t n, int accumulate) {
    TP_CHECK_ARG(ctx, "tp_div_bwd_b: NULL ctx");
    TP_NEED(gout, n, "gout"); TP_NEED(a, n, "a"); TP_NEED(b, n, "b"); TP_NEED(gb, n, "gb");
    return launch_ew<3>(ctx, OpDivBwdB(), gout->ptr, a->ptr, b->ptr, gb->ptr, n, accumulate);
}

int tp_relu_fwd(tp_ctx* ctx, const tp_buf* x, tp_buf* y, size_t n) {
    TP_CHECK_ARG(ctx, "tp_relu_fwd: NULL ctx");
    TP_NEED(x, n, "x"); TP_NEED(y, n, "y");
    return launch_ew<1>(ctx, OpRelu(), x->ptr, nullptr, nullptr, y->ptr, n, 0);
}

int tp_relu_bwd(tp_ctx* ctx, const tp_buf* x, const tp_buf* gout, tp_buf* gin, size_t n, int accumulate) {
    TP_CHECK_ARG(ctx, "tp_relu_bwd: NULL ctx");
    TP_NEED(x, n, "x"); TP_NEED(gout, n, "gout"); TP_NEED(gin, n, "gin");
    return launch_ew<2>(ctx, OpReluBwd(), x->ptr, gout->ptr, nullptr, gin->ptr, n, accumulate);
}

int tp_exp_fwd(tp_ctx* ctx, const tp_buf* x, tp_buf* y, size_t n) {
    TP_CHECK_ARG(ctx, "tp_exp_fwd: NULL ctx");
    TP_NEED(x, n, "x"); TP_NEED(y, n, "y");
    return launch_ew<1>(ctx, OpExp(), x->ptr, nullptr, nullptr, y->ptr, n, 0);
}

int tp_exp_bwd(tp_ctx* ctx, const tp_buf* y, const tp_buf* gout, tp_buf* gin, size_t n, int accumulate) {
    TP_CHECK_ARG(ctx, "tp_exp_bwd: NULL ctx");
    TP_NEED(y, n, "y"); TP_NEED(gout, n, "gout"); TP_NEED(gin, n, "gin");
    return launch_ew<2>(ctx, OpMul(), gout->ptr, y->ptr, nullptr, gin->ptr, n, accumulate);
}

int tp_log_fwd(tp_ctx* ctx, const tp_buf* x, tp_buf* y, size_t n) {
    TP_CHECK_ARG(ctx, "tp_log_fwd: NULL ctx");
    TP_NEED(x, n, "x"); TP_NEED(y, n, "y");
    return launch_ew<1>(ctx, OpLog(), x->ptr, nullptr, nullptr, y->ptr, n, 0);
}

int tp_log_bwd(tp_ctx* ctx, const tp_buf* x, const tp_buf* gout, tp_buf* gin, size_t n, int accumulate) {
    TP_CHECK_ARG(ctx, "tp_log_bwd: NULL ctx");
    TP_NEED(x, n, "x"); TP_NEED(gout, n, "gout"); TP_NEED(gin, n, "gin");
    return launch_ew<2>(ctx, OpDiv(), gout->ptr, x->ptr, nullptr, gin->ptr, n, accumulate);
}

int tp_sigmoid_fwd(tp_ctx* ctx, const tp_buf* x, tp_buf* y, size_t n) {
    TP_CHECK_ARG(ctx, "tp_sigmoid_fwd: NULL ctx");
    TP_NEED(x, n, "x"); TP_NEED(y, n, "y");
    return launch_ew<1>(ctx, OpSigmoid(), x->ptr, nullptr, nullptr, y->ptr, n, 0);
}

int tp_sigmoid_bwd(tp_ctx* ctx, const tp_buf* y, const tp_buf* gout, tp_buf* gin, size_t n, int accumulate) {
    TP_CHECK_ARG(ctx, "tp_sigmoid_bwd: NULL ctx");
    TP_NEED(y, n, "y"); TP_NEED(gout, n, "gout"); TP_NEED(gin, n, "gin");
    return launch_ew<2>(ctx, OpSigmoidBwd(), gout->ptr, y->ptr, nullptr, gin->ptr, n, accumulate);
}

int tp_pow_fwd(tp_ctx* ctx, const tp_buf* x, tp_buf* y, float exponent, size_t n) {
    TP_CHECK_ARG(ctx, "tp_pow_fwd: NULL ctx");
    TP_NEED(x, n, "x"); TP_NEED(y, n, "y");
    return launch_ew<1>(ctx, OpPow{exponent}, x->ptr, nullptr, nullptr, y->ptr, n, 0);
}

int tp_pow_bwd(tp_ctx* ctx, const tp_buf* x, const tp_buf* gout, tp_buf* gin, float exponent, size_t n, int accumulate) {
    TP_CHECK_ARG(ctx, "tp_pow_bwd: NULL ctx");
    TP_NEED(x, n, "x"); TP_NEED(gout, n, "gout"); TP_NEED(gin, n, "gin");
    return launch_ew<2>(ctx, OpPowBwd{exponent}, gout->ptr, x->ptr, nullptr, gin->ptr, n, accumulate);
}

int tp_mean_fwd(tp_ctx* ctx, const tp_buf* x, tp_buf* out, size_t n) {
    TP_CHECK_ARG(ctx && n > 0, "tp_mean_fwd: empty input");
    int rc = tp_sum_all(ctx, x, out, n);                       // sum: f32, then / len  (src/tensor.rs:773-775)
    if (rc) return rc;
    return launch_ew<1>(ctx, OpDivBy{(float)n}, out->ptr, nullptr, nullptr, out->ptr, 1, 0);
}

int tp_mean_bwd(tp_ctx* ctx, const tp_buf* gout, tp_buf* gin, size_t n, int accumulate) {
    TP_CHECK_ARG(ctx, "tp_mean_bwd: NULL ctx");
    TP_NEED(gout, 1, "gout"); TP_NEED(gin, n, "gin");
    return launch_ew<0>(ctx, OpMeanBwd{gout->ptr, (float)n}, nullptr, nullptr, nullptr, gin->ptr, n, accumulate);
}

int tp_bce_fwd(tp_ctx* ctx, const tp_buf* pred, const tp_buf* target, tp_buf* loss, size_t n) {
    TP_CHECK_ARG(ctx && n > 0, "tp_bce_fwd: empty input");
    TP_NEED(pred, n, "pred"); TP_NEED(target, n, "target"); TP_NEED(loss, 1, "loss");
    tp_buf* terms = nullptr;
    int rc = tp_buf_alloc(ctx, n, &terms);
    if (rc) return rc;
    rc = launch_ew<2>(ctx, OpBceTerm(), pred->ptr, target->ptr, nullptr, terms->ptr, n, 0);
    if (!rc) rc = tp_mean_fwd(ctx, terms, loss, n);            // acc / len  (src/loss.rs:23)
    tp_buf_release(terms);
    return rc;
}

int tp_bce_bwd(tp_ctx* ctx, const tp_buf* pred, const tp_buf* target, const tp_buf* gloss, tp_buf* gpred, tp_buf* gtarget, size_t n,
               int acc_pred, int acc_target) {
    TP_CHECK_ARG(ctx, "tp_bce_bwd: NULL ctx");
    TP_NEED(pred, n, "pred"); TP_NEED(target, n, "target"); TP_NEED(gloss, 1, "gloss");
    if (gpred) {
        TP_NEED(gpred, n, "gpred");
        int rc = launch_ew<2>(ctx, OpBceBwdP{gloss->ptr, (float)n}, pred->ptr, target->ptr, nullptr, gpred->ptr, n, acc_pred);
        if (rc) return rc;
    }
    if (gtarget) {
        TP_NEED(gtarget, n, "gtarget");
        int rc = launch_ew<1>(ctx, OpBceBwdT{gloss->ptr, (float)n}, pred->ptr, nullptr, nullptr, gtarget->ptr, n, acc_target);
        if (rc) return rc;
    }
    return TP_OK;
}

int tp_scale(tp_ctx* ctx, tp_buf* p, float s, size_t n) {
    TP_CHECK_ARG(ctx, "tp_scale: NULL ctx");
    TP_NEED(p, n, "p");
    if (n == 0) return TP_OK;
    scale_inplace<<<tp::grid_for(ctx, n, kThreads), kThreads, 0, ctx->stream>>>(p->ptr, s, n);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

int tp_buf_fill(tp_ctx* ctx, tp_buf* dst, float value, size_t n) {
    TP_CHECK_ARG(ctx, "tp_buf_fill: NULL ctx");
    TP_NEED(dst, n, "dst");
    if (value == 0.0f) {
        TP_CUDA(cudaMemsetAsync(dst->ptr, 0, n * sizeof(float), ctx->stream));
        return TP_OK;
    }
    return launch_ew<0>(ctx, OpFill{value}, nullptr, nullptr, nullptr, dst->ptr, n, 0);
}

}  // extern "C"

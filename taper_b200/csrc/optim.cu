// Optimizer steps: HBM-bound streaming kernels (SGD 12 B/param, Adam/AdamW 28 B/param), 128-bit vectorised.
// Reference: SGD::step (src/optim.rs:21-33), Adam::step (:83-113), AdamW::step (:148-168).
// The operation order of every formula follows the reference expression-for-expression and the
// kernels are compiled with -fmad=false so no multiply-add is contracted differently from the CPU code.
#include "common.cuh"
#include <cmath>
#include <cuda_bf16.h>

namespace {

constexpr int kThreads = 256;

struct AdamArgs {
    float step_size, beta1, beta2, eps, weight_decay, grad_scale, decay_factor;
    int decoupled;   // AdamW: p *= decay_factor first, wd = 0 inside
};

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, const AdamArgs& a) {
    if (a.decoupled) p *= a.decay_factor;                       // src/optim.rs:157-159
    if (a.grad_scale != 1.0f) g *= a.grad_scale;                // data-parallel mean of the summed gradient
    float gg = g + a.weight_decay * p;                          // src/optim.rs:98
    m = a.beta1 * m + (1.0f - a.beta1) * gg;                    // :101
    v = a.beta2 * v + (1.0f - a.beta2) * gg * gg;               // :104
    p -= a.step_size * m / (sqrtf(v) + a.eps);                  // :107
}

__global__ void __launch_bounds__(kThreads)
adam_vec4_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
                 size_t n4, AdamArgs a) {
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride) {
        float4 pp = p[i], gg = __ldg(g + i), mm = m[i], vv = v[i];
        adam_elem(pp.x, gg.x, mm.x, vv.x, a);
        adam_elem(pp.y, gg.y, mm.y, vv.y, a);
        adam_elem(pp.z, gg.z, mm.z, vv.z, a);
        adam_elem(pp.w, gg.w, mm.w, vv.w, a);
        p[i] = pp; m[i] = mm; v[i] = vv;
    }
}

__global__ void __launch_bounds__(kThreads)
adam_scalar_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                   size_t n, AdamArgs a) {
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
        float pp = p[i], mm = m[i], vv = v[i];
        adam_elem(pp, g[i], mm, vv, a);
        p[i] = pp; m[i] = mm; v[i] = vv;
    }
}

__global__ void __launch_bounds__(kThreads)
sgd_kernel(float* __restrict__ p, const float* __restrict__ g, size_t n, float lr, float grad_scale) {
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
        float gg = g[i];
        if (grad_scale != 1.0f) gg *= grad_scale;
        p[i] -= lr * gg;                                        // src/optim.rs:29
    }
}

// SGD with the learning rate read from device memory: a captured (CUDA-graph) step follows SGD::set_lr without re-capture
__global__ void __launch_bounds__(kThreads)
sgd_dev_kernel(float* __restrict__ p, const float* __restrict__ g, size_t n, const float* __restrict__ lr1, float grad_scale) {
    const float lr = *lr1;
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
        float gg = g[i];
        if (grad_scale != 1.0f) gg *= grad_scale;
        p[i] -= lr * gg;                                        // src/optim.rs:29
    }
}

// ---- device-resident optimizer state: lets a captured (CUDA-graph) step advance t and follow set_lr
// without the host patching kernel arguments.  hyper = {t (int bits), lr, b1, b2, eps, wd, step_size, decay}
enum { H_T = 0, H_LR, H_B1, H_B2, H_EPS, H_WD, H_SS, H_DECAY, H_COUNT };

__device__ __forceinline__ float powi_dev(float a, int b) {     // f32::powi (compiler-rt __powisf2)
    float r = 1.0f;
    unsigned int e = (unsigned int)b;
    while (true) {
        if (e & 1u) r *= a;
        e >>= 1;
        if (e == 0) break;
        a *= a;
    }
    return r;
}

__global__ void adam_advance_kernel(float* __restrict__ h) {
    int t = __float_as_int(h[H_T]) + 1;                        // self.t += 1  (src/optim.rs:86)
    h[H_T] = __int_as_float(t);
    float bc1 = 1.0f - powi_dev(h[H_B1], t);                   // :88
    float bc2 = 1.0f - powi_dev(h[H_B2], t);                   // :89
    h[H_SS] = h[H_LR] * (sqrtf(bc2) / bc1);                    // :90
    h[H_DECAY] = 1.0f - h[H_LR] * h[H_WD];                     // AdamW  :157
}

__global__ void set_scalar_kernel(float* __restrict__ dst, float v) { *dst = v; }

__global__ void __launch_bounds__(kThreads)
adam_dev_vec4_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
                     size_t n4, const float* __restrict__ h, float grad_scale, int decoupled) {
    AdamArgs a{h[H_SS], h[H_B1], h[H_B2], h[H_EPS], decoupled ? 0.0f : h[H_WD], grad_scale, h[H_DECAY],
               (decoupled && h[H_WD] > 0.0f) ? 1 : 0};
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride) {
        float4 pp = p[i], gg = __ldg(g + i), mm = m[i], vv = v[i];
        adam_elem(pp.x, gg.x, mm.x, vv.x, a);
        adam_elem(pp.y, gg.y, mm.y, vv.y, a);
        adam_elem(pp.z, gg.z, mm.z, vv.z, a);
        adam_elem(pp.w, gg.w, mm.w, vv.w, a);
        p[i] = pp; m[i] = mm; v[i] = vv;
    }
}

__global__ void __launch_bounds__(kThreads)
adam_dev_scalar_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                       size_t n, const float* __restrict__ h, float grad_scale, int decoupled) {
    AdamArgs a{h[H_SS], h[H_B1], h[H_B2], h[H_EPS], decoupled ? 0.0f : h[H_WD], grad_scale, h[H_DECAY],
               (decoupled && h[H_WD] > 0.0f) ? 1 : 0};
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
        float pp = p[i], mm = m[i], vv = v[i];
        adam_elem(pp, g[i], mm, vv, a);
        p[i] = pp; m[i] = mm; v[i] = vv;
    }
}

// p *= decay (hyper[H_DECAY]) — AdamW's decoupled decay of parameters that have no gradient this step
__global__ void __launch_bounds__(kThreads)
decay_dev_kernel(float* __restrict__ p, size_t n, const float* __restrict__ h) {
    if (!(h[H_WD] > 0.0f)) return;
    const float d = h[H_DECAY];
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) p[i] *= d;
}

// Multi-tensor form for arenas where some parameters have no gradient this step (the reference's conv weights):
// one launch over up to kMaxSeg slices of the flat arena.  mode 1: Adam / AdamW step; mode 2: AdamW's decoupled decay only
// (parameters without a gradient still decay, src/optim.rs:154-161).  blockIdx.y = slice.
constexpr int kMaxSeg = 32;
struct SegTable {
    long long off[kMaxSeg];          // element offset of the slice in the arenas (multiple of 4)
    int n4[kMaxSeg];                 // float4 count (slices are padded to 4 with zeros)
    int mode[kMaxSeg];
};

__global__ void __launch_bounds__(kThreads)
adam_dev_segments_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                         const __grid_constant__ SegTable t, const float* __restrict__ h, float grad_scale, int decoupled) {
    const int sgm = blockIdx.y;
    const int mode = t.mode[sgm];
    const size_t n4 = (size_t)t.n4[sgm];
    float4* p4 = reinterpret_cast<float4*>(p + t.off[sgm]);
    const size_t stride = (size_t)gridDim.x * kThreads;
    if (mode == 2) {
        if (!(h[H_WD] > 0.0f)) return;
        const float d = h[H_DECAY];
        for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride) {
            float4 pp = p4[i];
            pp.x *= d; pp.y *= d; pp.z *= d; pp.w *= d;
            p4[i] = pp;
        }
        return;
    }
    AdamArgs a{h[H_SS], h[H_B1], h[H_B2], h[H_EPS], decoupled ? 0.0f : h[H_WD], grad_scale, h[H_DECAY],
               (decoupled && h[H_WD] > 0.0f) ? 1 : 0};
    const float4* g4 = reinterpret_cast<const float4*>(g + t.off[sgm]);
    float4* m4 = reinterpret_cast<float4*>(m + t.off[sgm]);
    float4* v4 = reinterpret_cast<float4*>(v + t.off[sgm]);
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride) {
        float4 pp = p4[i], gg = __ldg(g4 + i), mm = m4[i], vv = v4[i];
        adam_elem(pp.x, gg.x, mm.x, vv.x, a);
        adam_elem(pp.y, gg.y, mm.y, vv.y, a);
        adam_elem(pp.z, gg.z, mm.z, vv.z, a);
        adam_elem(pp.w, gg.w, mm.w, vv.w, a);
        p4[i] = pp; m4[i] = mm; v4[i] = vv;
    }
}

int launch_adam(tp_ctx* ctx, tp_buf* p, const tp_buf* g, tp_buf* m, tp_buf* v, size_t n, const AdamArgs& a, const char* fn) {
    TP_CHECK_ARG(ctx, "%s: NULL ctx", fn);
    TP_NEED(p, n, "p"); TP_NEED(g, n, "g"); TP_NEED(m, n, "m"); TP_NEED(v, n, "v");
    if (!n) return TP_OK;
    bool vec = !(((uintptr_t)p->ptr | (uintptr_t)g->ptr | (uintptr_t)m->ptr | (uintptr_t)v->ptr) & 15);
    size_t n4 = vec ? n / 4 : 0;
    if (n4) {
        adam_vec4_kernel<<<tp::grid_for(ctx, n4, kThreads), kThreads, 0, ctx->stream>>>(
            (float4*)p->ptr, (const float4*)g->ptr, (float4*)m->ptr, (float4*)v->ptr, n4, a);
        TP_LAUNCH_OK(ctx);
    }
    size_t done = n4 * 4;
    if (done < n) {
        adam_scalar_kernel<<<tp::grid_for(ctx, n - done, kThreads), kThreads, 0, ctx->stream>>>(
            p->ptr + done, g->ptr + done, m->ptr + done, v->ptr + done, n - done, a);
        TP_LAUNCH_OK(ctx);
    }
    return TP_OK;
}


// ---- optimizer step that also refreshes the bf16 hi/lo planes of the parameters (operands of the bf16x3 GEMMs, gemm_bx3.cu):
// the same arithmetic as adam_dev_vec4_kernel / sgd_kernel, plus 4 bytes per parameter of stores.
__device__ __forceinline__ void split_store4_opt(uint16_t* hi, uint16_t* lo, const float (&v)[4]) {
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        h[j] = __float2bfloat16_rn(v[j]);
        l[j] = __float2bfloat16_rn(v[j] - __bfloat162float(h[j]));
    }
    uint2 ph, pl;
    ph.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
    ph.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
    pl.x = (uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16);
    pl.y = (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16);
    *(uint2*)hi = ph;
    *(uint2*)lo = pl;
}

__global__ void __launch_bounds__(kThreads)
adam_dev_split_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
                      size_t n4, const float* __restrict__ h, float grad_scale, int decoupled, uint16_t* __restrict__ hi,
                      uint16_t* __restrict__ lo, unsigned long long* stamp) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (stamp && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        *stamp = gt;
    }
    AdamArgs a{h[H_SS], h[H_B1], h[H_B2], h[H_EPS], decoupled ? 0.0f : h[H_WD], grad_scale, h[H_DECAY],
               (decoupled && h[H_WD] > 0.0f) ? 1 : 0};
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride) {
        float4 pp = p[i], gg = __ldg(g + i), mm = m[i], vv = v[i];
        adam_elem(pp.x, gg.x, mm.x, vv.x, a);
        adam_elem(pp.y, gg.y, mm.y, vv.y, a);
        adam_elem(pp.z, gg.z, mm.z, vv.z, a);
        adam_elem(pp.w, gg.w, mm.w, vv.w, a);
        p[i] = pp; m[i] = mm; v[i] = vv;
        const float o[4] = {pp.x, pp.y, pp.z, pp.w};
        split_store4_opt(hi + 4 * i, lo + 4 * i, o);
    }
}

__global__ void __launch_bounds__(kThreads)
sgd_split_kernel(float4* __restrict__ p, const float4* __restrict__ g, size_t n4, float lr, float grad_scale,
                 uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, unsigned long long* stamp) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (stamp && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        *stamp = gt;
    }
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride) {
        float4 pp = p[i], gg = __ldg(g + i);
        if (grad_scale != 1.0f) { gg.x *= grad_scale; gg.y *= grad_scale; gg.z *= grad_scale; gg.w *= grad_scale; }
        pp.x -= lr * gg.x; pp.y -= lr * gg.y; pp.z -= lr * gg.z; pp.w -= lr * gg.w;      // src/optim.rs:29
        p[i] = pp;
        const float o[4] = {pp.x, pp.y, pp.z, pp.w};
        split_store4_opt(hi + 4 * i, lo + 4 * i, o);
    }
}

}  // namespace

namespace tp {

// optimizer kind: 0 SGD, 1 Adam, 2 AdamW.  n is a multiple of 4 (arena length); all pointers 16-byte aligned.
int optimizer_step_split(tp_ctx* ctx, int kind, float* p, const float* g, float* m, float* v, const float* hyper, float sgd_lr,
                         float grad_scale, size_t n, uint16_t* hi, uint16_t* lo, bool pdl, unsigned long long* stamp) {
    if (!n) return TP_OK;
    const size_t n4 = n / 4;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid_for(ctx, n4, kThreads));
    cfg.blockDim = dim3(kThreads);
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    if (kind == 0) {
        TP_CUDA(cudaLaunchKernelEx(&cfg, sgd_split_kernel, (float4*)p, (const float4*)g, n4, sgd_lr, grad_scale, hi, lo, stamp));
    } else {
        TP_CUDA(cudaLaunchKernelEx(&cfg, adam_dev_split_kernel, (float4*)p, (const float4*)g, (float4*)m, (float4*)v, n4, hyper,
                                   grad_scale, kind == 2 ? 1 : 0, hi, lo, stamp));
    }
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

}  // namespace tp

namespace {

// f32::powi == compiler-rt __powisf2: square-and-multiply in f32
float powi_f32(float a, int b) {
    const bool recip = b < 0;
    float r = 1.0f;
    unsigned int e = recip ? (unsigned int)(-(long long)b) : (unsigned int)b;
    while (true) {
        if (e & 1u) r *= a;
        e >>= 1;
        if (e == 0) break;
        a *= a;
    }
    return recip ? 1.0f / r : r;
}

}  // namespace

extern "C" {

float tp_adam_step_size(float lr, float beta1, float beta2, int t) {
    volatile float bc1 = 1.0f - powi_f32(beta1, t);             // src/optim.rs:88
    volatile float bc2 = 1.0f - powi_f32(beta2, t);             // :89
    volatile float ratio = sqrtf(bc2) / bc1;
    return lr * ratio;                                          // :90
}

int tp_sgd_step(tp_ctx* ctx, tp_buf* p, const tp_buf* g, float lr, float grad_scale, size_t n) {
    TP_CHECK_ARG(ctx, "tp_sgd_step: NULL ctx");
    TP_NEED(p, n, "p"); TP_NEED(g, n, "g");
    if (!n) return TP_OK;
    sgd_kernel<<<tp::grid_for(ctx, n, kThreads), kThreads, 0, ctx->stream>>>(p->ptr, g->ptr, n, lr, grad_scale);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

int tp_adam_step(tp_ctx* ctx, tp_buf* p, const tp_buf* g, tp_buf* m, tp_buf* v, float step_size, float beta1, float beta2,
                 float eps, float weight_decay, float grad_scale, size_t n) {
    AdamArgs a{step_size, beta1, beta2, eps, weight_decay, grad_scale, 1.0f, 0};
    return launch_adam(ctx, p, g, m, v, n, a, "tp_adam_step");
}

int tp_adam_hyper_init(tp_ctx* ctx, tp_buf* hyper, float lr, float beta1, float beta2, float eps, float weight_decay) {
    TP_CHECK_ARG(ctx, "tp_adam_hyper_init: NULL ctx");
    TP_NEED(hyper, H_COUNT, "hyper");
    float h[H_COUNT] = {0.0f, lr, beta1, beta2, eps, weight_decay, 0.0f, 1.0f};
    return tp_buf_upload(ctx, hyper, h, H_COUNT);
}

int tp_sgd_step_dev(tp_ctx* ctx, tp_buf* p, const tp_buf* g, const tp_buf* lr1, float grad_scale, size_t n) {
    TP_CHECK_ARG(ctx, "tp_sgd_step_dev: NULL ctx");
    TP_NEED(p, n, "p"); TP_NEED(g, n, "g"); TP_NEED(lr1, 1, "lr");
    if (!n) return TP_OK;
    sgd_dev_kernel<<<tp::grid_for(ctx, n, kThreads), kThreads, 0, ctx->stream>>>(p->ptr, g->ptr, n, lr1->ptr, grad_scale);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

int tp_buf_set_scalar(tp_ctx* ctx, tp_buf* buf, size_t index, float value) {
    TP_CHECK_ARG(ctx, "tp_buf_set_scalar: NULL ctx");
    TP_NEED(buf, index + 1, "buf");
    set_scalar_kernel<<<1, 1, 0, ctx->stream>>>(buf->ptr + index, value);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

int tp_adam_hyper_set_lr(tp_ctx* ctx, tp_buf* hyper, float lr) {
    TP_CHECK_ARG(ctx, "tp_adam_hyper_set_lr: NULL ctx");
    TP_NEED(hyper, H_COUNT, "hyper");
    set_scalar_kernel<<<1, 1, 0, ctx->stream>>>(hyper->ptr + H_LR, lr);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

int tp_adam_advance(tp_ctx* ctx, tp_buf* hyper) {
    TP_CHECK_ARG(ctx, "tp_adam_advance: NULL ctx");
    TP_NEED(hyper, H_COUNT, "hyper");
    adam_advance_kernel<<<1, 1, 0, ctx->stream>>>(hyper->ptr);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

int tp_adam_step_dev(tp_ctx* ctx, tp_buf* p, const tp_buf* g, tp_buf* m, tp_buf* v, const tp_buf* hyper, float grad_scale,
                     int decoupled, size_t n) {
    TP_CHECK_ARG(ctx, "tp_adam_step_dev: NULL ctx");
    TP_NEED(p, n, "p"); TP_NEED(g, n, "g"); TP_NEED(m, n, "m"); TP_NEED(v, n, "v"); TP_NEED(hyper, H_COUNT, "hyper");
    if (!n) return TP_OK;
    bool vec = !(((uintptr_t)p->ptr | (uintptr_t)g->ptr | (uintptr_t)m->ptr | (uintptr_t)v->ptr) & 15);
    size_t n4 = vec ? n / 4 : 0;
    if (n4) {
        adam_dev_vec4_kernel<<<tp::grid_for(ctx, n4, kThreads), kThreads, 0, ctx->stream>>>(
            (float4*)p->ptr, (const float4*)g->ptr, (float4*)m->ptr, (float4*)v->ptr, n4, hyper->ptr, grad_scale, decoupled);
        TP_LAUNCH_OK(ctx);
    }
    size_t done = n4 * 4;
    if (done < n) {
        adam_dev_scalar_kernel<<<tp::grid_for(ctx, n - done, kThreads), kThreads, 0, ctx->stream>>>(
            p->ptr + done, g->ptr + done, m->ptr + done, v->ptr + done, n - done, hyper->ptr, grad_scale, decoupled);
        TP_LAUNCH_OK(ctx);
    }
    return TP_OK;
}

int tp_adam_step_segments(tp_ctx* ctx, tp_buf* p, const tp_buf* g, tp_buf* m, tp_buf* v, const tp_buf* hyper, float grad_scale,
                          int decoupled, const int64_t* offsets, const int64_t* lengths, const int* modes, int n_segments) {
    TP_CHECK_ARG(ctx && offsets && lengths && modes && n_segments >= 0, "tp_adam_step_segments: bad arguments");
    TP_NEED(hyper, H_COUNT, "hyper");
    TP_CHECK_ARG(p && g && m && v, "tp_adam_step_segments: NULL arena");
    TP_CHECK_ARG(!(((uintptr_t)p->ptr | (uintptr_t)g->ptr | (uintptr_t)m->ptr | (uintptr_t)v->ptr) & 15), "tp_adam_step_segments: arenas must be 16-byte aligned");
    // `s` is the scan position: skipped slices (mode 0 / empty) do not count against a launch's table, so the next launch
    // must resume where this one stopped, not kMaxSeg further (a slice would otherwise be stepped twice)
    for (int s = 0; s < n_segments;) {
        SegTable t{};
        int cnt = 0, max_n4 = 0;
        for (; s < n_segments && cnt < kMaxSeg; ++s) {
            if (modes[s] == 0 || lengths[s] <= 0) continue;
            const int64_t n_pad = (lengths[s] + 3) & ~(int64_t)3;
            TP_CHECK_ARG(offsets[s] >= 0 && offsets[s] % 4 == 0 && (size_t)(offsets[s] + n_pad) <= p->n && (size_t)(offsets[s] + n_pad) <= g->n &&
                             (size_t)(offsets[s] + n_pad) <= m->n && (size_t)(offsets[s] + n_pad) <= v->n,
                         "tp_adam_step_segments: slice %d outside the arenas or not 16-byte aligned", s);
            t.off[cnt] = offsets[s];
            t.n4[cnt] = (int)(n_pad / 4);
            t.mode[cnt] = modes[s];
            if (t.n4[cnt] > max_n4) max_n4 = t.n4[cnt];
            ++cnt;
        }
        if (!cnt) continue;
        dim3 grid(tp::grid_for(ctx, max_n4, kThreads, 2), cnt);
        adam_dev_segments_kernel<<<grid, kThreads, 0, ctx->stream>>>(p->ptr, g->ptr, m->ptr, v->ptr, t, hyper->ptr, grad_scale, decoupled);
        TP_LAUNCH_OK(ctx);
    }
    return TP_OK;
}

int tp_decay_dev(tp_ctx* ctx, tp_buf* p, const tp_buf* hyper, size_t n) {
    TP_CHECK_ARG(ctx, "tp_decay_dev: NULL ctx");
    TP_NEED(p, n, "p"); TP_NEED(hyper, H_COUNT, "hyper");
    if (!n) return TP_OK;
    decay_dev_kernel<<<tp::grid_for(ctx, n, kThreads), kThreads, 0, ctx->stream>>>(p->ptr, n, hyper->ptr);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

int tp_adamw_step(tp_ctx* ctx, tp_buf* p, const tp_buf* g, tp_buf* m, tp_buf* v, float step_size, float beta1, float beta2,
                  float eps, float decay_factor, float grad_scale, size_t n) {
    AdamArgs a{step_size, beta1, beta2, eps, 0.0f, grad_scale, decay_factor, 1};
    return launch_adam(ctx, p, g, m, v, n, a, "tp_adamw_step");
}

}  // extern "C"

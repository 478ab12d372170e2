// Softmax / cross-entropy / accuracy.  The reference composes log_softmax from six tensor ops
// (max, sub_broadcast_rows, exp, sum, log, sub_broadcast_rows — src/loss.rs:101-126) and gathers the NLL on
// the host (src/loss.rs:152-165); here it is one row-parallel pass (one warp per row) plus a
// deterministic two-stage mean.  The backward is the reference's direct formula (src/loss.rs:174-191).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarpsPerBlock = kThreads / 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// `t as usize` (src/loss.rs:160): saturating float->unsigned cast, NaN -> 0
__device__ __forceinline__ unsigned int class_of(float t) {
    if (!(t > 0.0f)) return 0u;
    if (t >= 4294967040.0f) return 0xffffffffu;
    return (unsigned int)t;
}

// mode 0: log_softmax, 1: softmax.  If targets != NULL also emits per-block NLL partial sums.
__global__ void __launch_bounds__(kThreads)
softmax_rows_kernel(const float* __restrict__ x, const float* __restrict__ targets, float* __restrict__ out,
                    float* __restrict__ partial, int* __restrict__ err, int rows, int cols, int mode) {
    __shared__ float sm[kWarpsPerBlock];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int row = blockIdx.x * kWarpsPerBlock + wid;
    float nll = 0.0f;
    if (row < rows) {
        const float* xr = x + (size_t)row * cols;
        float m = -INFINITY;
        for (int c = lane; c < cols; c += 32) {           // row max, strict '>' from -inf (src/tensor.rs:1062)
            float v = __ldg(xr + c);
            if (v > m) m = v;
        }
        m = warp_max(m);
        float s = 0.0f;
        for (int c = lane; c < cols; c += 32) s += expf(__ldg(xr + c) - m);
        s = warp_sum(s);
        float* o = out + (size_t)row * cols;
        if (mode == 0) {
            float ls = logf(s);
            for (int c = lane; c < cols; c += 32) o[c] = (__ldg(xr + c) - m) - ls;
            if (targets && lane == 0) {
                unsigned int cls = class_of(__ldg(targets + row));
                if (cls >= (unsigned int)cols) { atomicExch(err, 1); cls = cols - 1; }
                nll = -((__ldg(xr + cls) - m) - ls);
            }
        } else {
            for (int c = lane; c < cols; c += 32) o[c] = expf(__ldg(xr + c) - m) / s;
        }
    }
    if (partial) {
        if (lane == 0) sm[wid] = nll;
        __syncthreads();
        if (threadIdx.x == 0) {
            float acc = 0.0f;
#pragma unroll
            for (int j = 0; j < kWarpsPerBlock; ++j) acc += sm[j];   // rows ascending inside the block
            partial[blockIdx.x] = acc;
        }
    }
}

// Fused training-step head: log_softmax + NLL mean (src/loss.rs:152-165) + accuracy count (src/loss.rs:271-290) in ONE
// launch.  One warp per row; per-block partial sums are parked, the last block to finish folds them in block order
// (deterministic) and writes loss[0] = sum / rows and correct[0].
__global__ void __launch_bounds__(kThreads)
xent_acc_fused_kernel(const float* __restrict__ x, const float* __restrict__ targets, float* __restrict__ logp,
                      float* __restrict__ partial, float* __restrict__ loss, float* __restrict__ correct, int* __restrict__ err,
                      int* __restrict__ ticket, int rows, int cols) {
    __shared__ float sm_nll[kWarpsPerBlock], sm_hit[kWarpsPerBlock];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int row = blockIdx.x * kWarpsPerBlock + wid;
    float nll = 0.0f, hit = 0.0f;
    if (row < rows) {
        const float* xr = x + (size_t)row * cols;
        float m = -INFINITY;
        int bi = INT_MAX;
        for (int c = lane; c < cols; c += 32) {           // row max + first-index argmax, strict '>' (src/tensor.rs:1062)
            float v = __ldg(xr + c);
            if (v > m) { m = v; bi = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, m, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > m || (ov == m && oi < bi)) { m = ov; bi = oi; }
        }
        float s = 0.0f;
        for (int c = lane; c < cols; c += 32) s += expf(__ldg(xr + c) - m);
        s = warp_sum(s);
        const float ls = logf(s);
        float* o = logp + (size_t)row * cols;
        for (int c = lane; c < cols; c += 32) o[c] = (__ldg(xr + c) - m) - ls;
        if (lane == 0) {
            const float t = __ldg(targets + row);
            unsigned int cls = class_of(t);
            if (cls >= (unsigned int)cols) { atomicExch(err, 1); cls = cols - 1; }
            nll = -((__ldg(xr + cls) - m) - ls);
            const float idx = (bi == INT_MAX) ? 0.0f : (float)bi;
            if (fabsf(idx - t) < 1e-6f) hit = 1.0f;                              // src/loss.rs:284
        }
    }
    if (lane == 0) { sm_nll[wid] = nll; sm_hit[wid] = hit; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.0f, h = 0.0f;
#pragma unroll
        for (int j = 0; j < kWarpsPerBlock; ++j) { a += sm_nll[j]; h += sm_hit[j]; }      // rows ascending inside the block
        partial[2 * blockIdx.x] = a;
        partial[2 * blockIdx.x + 1] = h;
        __threadfence();
        int prev = atomicAdd(ticket, 1);
        s_last = (prev == (int)gridDim.x - 1);
        if (s_last) *ticket = 0;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        float a = 0.0f, h = 0.0f;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += kThreads) { a += __ldcg(partial + 2 * i); h += __ldcg(partial + 2 * i + 1); }
        a = warp_sum(a); h = warp_sum(h);
        if (lane == 0) { sm_nll[wid] = a; sm_hit[wid] = h; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float sa = 0.0f, sh = 0.0f;
#pragma unroll
            for (int j = 0; j < kWarpsPerBlock; ++j) { sa += sm_nll[j]; sh += sm_hit[j]; }
            loss[0] = sa / (float)rows;                                          // acc / b as f32 (src/loss.rs:164)
            if (correct) correct[0] = sh;
        }
    }
}

// out[0] = (sum_j partial[j]) / divisor, j ascending (single thread block, fixed tree => deterministic)
__global__ void __launch_bounds__(kThreads)
fold_scalar_kernel(const float* __restrict__ partial, float* __restrict__ out, int n, float divisor) {
    __shared__ float sm[kWarpsPerBlock];
    float acc = 0.0f;
    for (int i = threadIdx.x; i < n; i += kThreads) acc += partial[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.0f;
#pragma unroll
        for (int j = 0; j < kWarpsPerBlock; ++j) s += sm[j];
        out[0] = s / divisor;          // acc / b as f32 (src/loss.rs:164)
    }
}

__global__ void __launch_bounds__(kThreads)
xent_bwd_kernel(const float* __restrict__ logp, const float* __restrict__ targets, const float* __restrict__ gloss,
                float* __restrict__ glogits, size_t total, int cols, float inv_b, int accumulate) {
    const float scale = __ldg(gloss) * inv_b;        // g[0] / B  (src/loss.rs:186)
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += stride) {
        size_t r = i / cols;
        unsigned int c = (unsigned int)(i - r * cols);
        float p = expf(__ldg(logp + i));
        if (class_of(__ldg(targets + r)) == c) p -= 1.0f;
        p *= scale;
        glogits[i] = accumulate ? glogits[i] + p : p;
    }
}

__global__ void __launch_bounds__(kThreads)
accuracy_kernel(const float* __restrict__ pred, const float* __restrict__ targets, float* __restrict__ partial,
                int rows, int cols) {
    __shared__ float sm[kWarpsPerBlock];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int row = blockIdx.x * kWarpsPerBlock + wid;
    float hit = 0.0f;
    if (row < rows) {
        const float* xr = pred + (size_t)row * cols;
        float best = -INFINITY;
        int bi = INT_MAX;
        for (int c = lane; c < cols; c += 32) {
            float v = __ldg(xr + c);
            if (v > best) { best = v; bi = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        float idx = (bi == INT_MAX) ? 0.0f : (float)bi;
        if (lane == 0 && fabsf(idx - __ldg(targets + row)) < 1e-6f) hit = 1.0f;    // src/loss.rs:284
    }
    if (lane == 0) sm[wid] = hit;
    __syncthreads();
    if (threadIdx.x == 0) {
        float acc = 0.0f;
#pragma unroll
        for (int j = 0; j < kWarpsPerBlock; ++j) acc += sm[j];
        partial[blockIdx.x] = acc;
    }
}

int softmax_common(tp_ctx* ctx, const tp_buf* x, const tp_buf* targets, tp_buf* out, tp_buf* loss, int rows, int cols, int mode,
                   const char* fn) {
    TP_CHECK_ARG(ctx && rows >= 0 && cols > 0, "%s: bad dims %d x %d", fn, rows, cols);
    size_t total = (size_t)rows * cols;
    TP_NEED(x, total, "x"); TP_NEED(out, total, "out");
    if (targets) { TP_NEED(targets, rows, "targets"); TP_NEED(loss, 1, "loss"); }
    int blocks = (rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (blocks < 1) blocks = 1;
    float* partial = nullptr;
    if (targets) {
        int rc = tp::ensure_scratch(ctx, (size_t)blocks * sizeof(float));
        if (rc) return rc;
        partial = ctx->scratch;
    }
    softmax_rows_kernel<<<blocks, kThreads, 0, ctx->stream>>>(x->ptr, targets ? targets->ptr : nullptr, out->ptr, partial,
                                                             ctx->dev_error, rows, cols, mode);
    TP_LAUNCH_OK(ctx);
    if (targets) {
        fold_scalar_kernel<<<1, kThreads, 0, ctx->stream>>>(partial, loss->ptr, blocks, rows ? (float)rows : 1.0f);
        TP_LAUNCH_OK(ctx);
    }
    return TP_OK;
}

}  // namespace

extern "C" {

int tp_log_softmax_fwd(tp_ctx* ctx, const tp_buf* x, tp_buf* logp, int rows, int cols) {
    return softmax_common(ctx, x, nullptr, logp, nullptr, rows, cols, 0, "tp_log_softmax_fwd");
}

int tp_softmax_fwd(tp_ctx* ctx, const tp_buf* x, tp_buf* p, int rows, int cols) {
    return softmax_common(ctx, x, nullptr, p, nullptr, rows, cols, 1, "tp_softmax_fwd");
}

int tp_softmax_xent_fwd(tp_ctx* ctx, const tp_buf* logits, const tp_buf* targets, tp_buf* logp, tp_buf* loss, int rows, int cols) {
    TP_CHECK_ARG(targets && loss, "tp_softmax_xent_fwd: targets and loss are required");
    return softmax_common(ctx, logits, targets, logp, loss, rows, cols, 0, "tp_softmax_xent_fwd");
}

int tp_softmax_xent_acc_fwd(tp_ctx* ctx, const tp_buf* logits, const tp_buf* targets, tp_buf* logp, tp_buf* loss, tp_buf* correct,
                            int rows, int cols) {
    TP_CHECK_ARG(ctx && rows > 0 && cols > 0, "tp_softmax_xent_acc_fwd: bad dims %d x %d", rows, cols);
    size_t total = (size_t)rows * cols;
    TP_NEED(logits, total, "logits"); TP_NEED(targets, rows, "targets"); TP_NEED(logp, total, "logp"); TP_NEED(loss, 1, "loss");
    if (correct) TP_NEED(correct, 1, "correct");
    int blocks = (rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
    int rc = tp::ensure_scratch(ctx, (size_t)blocks * 2 * sizeof(float));
    if (rc) return rc;
    xent_acc_fused_kernel<<<blocks, kThreads, 0, ctx->stream>>>(logits->ptr, targets->ptr, logp->ptr, ctx->scratch, loss->ptr,
                                                               correct ? correct->ptr : nullptr, ctx->dev_error,
                                                               ctx->dev_counters + tp::kCounterXent, rows, cols);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

int tp_softmax_xent_bwd(tp_ctx* ctx, const tp_buf* logp, const tp_buf* targets, const tp_buf* gloss, tp_buf* glogits,
                        int rows, int cols, int accumulate) {
    TP_CHECK_ARG(ctx && rows >= 0 && cols > 0, "tp_softmax_xent_bwd: bad dims %d x %d", rows, cols);
    size_t total = (size_t)rows * cols;
    TP_NEED(logp, total, "logp"); TP_NEED(targets, rows, "targets"); TP_NEED(gloss, 1, "gloss"); TP_NEED(glogits, total, "glogits");
    if (!total) return TP_OK;
    xent_bwd_kernel<<<tp::grid_for(ctx, total, kThreads), kThreads, 0, ctx->stream>>>(
        logp->ptr, targets->ptr, gloss->ptr, glogits->ptr, total, cols, 1.0f / (float)rows, accumulate);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

int tp_accuracy_count(tp_ctx* ctx, const tp_buf* pred, const tp_buf* targets, tp_buf* correct, int rows, int cols) {
    TP_CHECK_ARG(ctx && rows >= 0 && cols > 0, "tp_accuracy_count: bad dims %d x %d", rows, cols);
    TP_NEED(pred, (size_t)rows * cols, "pred"); TP_NEED(targets, rows, "targets"); TP_NEED(correct, 1, "correct");
    int blocks = (rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (blocks < 1) blocks = 1;
    int rc = tp::ensure_scratch(ctx, (size_t)blocks * sizeof(float));
    if (rc) return rc;
    accuracy_kernel<<<blocks, kThreads, 0, ctx->stream>>>(pred->ptr, targets->ptr, ctx->scratch, rows, cols);
    TP_LAUNCH_OK(ctx);
    fold_scalar_kernel<<<1, kThreads, 0, ctx->stream>>>(ctx->scratch, correct->ptr, blocks, 1.0f);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

}  // extern "C"

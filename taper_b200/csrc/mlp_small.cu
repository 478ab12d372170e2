// A chain of small Linear(+ReLU) layers (every width <= 128: the classifier head behind the CNN's global average pool,
// examples/train_mnist_cnn.rs:80-100: 128-128-64-10) as TWO launches — forward and backward — instead of one tensor-core
// launch per matrix product plus a kernel per elementwise step.  At these sizes (0.05 GFLOP per pass at batch 1024) a
// tcgen05 launch is nothing but its own latency chain (TMA -> mbarrier -> MMA -> commit -> TMEM load, 7-11 us each, 15 launches
// per step); exact fp32 FFMA on the CUDA cores finishes the whole chain in a few microseconds with every weight in shared memory.
//
// Replaces, per layer, Linear::forward (src/nn.rs:54-60: transpose src/tensor.rs:544-591, matmul src/ops.rs:200-228,
// add_broadcast src/tensor.rs:636-704), ReLU (src/ops.rs:312-374) and their backward closures (matmul backward
// src/ops.rs:254-291, bias column sums src/tensor.rs:680-691, ReLU mask src/ops.rs:358-370).
//
// A CTA owns kRows consecutive batch rows.  Forward: all weights transposed into shared memory ([in][out + 1]: the inner
// loop reads consecutive outputs conflict-free), activations of the CTA's rows staged per layer, every layer's output written
// to global (the backward needs them).  Backward: per layer dW / db partials of the CTA's rows and the gradient of the
// layer's input (masked by the previous ReLU); a fold kernel sums the per-CTA partials in CTA order (deterministic).
// Products are accumulated in ascending k with fmaf, bias added last — the order of the reference's matmul + add_broadcast.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kRows = 32;            // batch rows per CTA
constexpr int kMaxL = 4;
constexpr int kMaxW = 128;

struct MlpArgs {
    int L, batch;
    int dims[kMaxL + 1];
    int relu[kMaxL];
    const float* W[kMaxL];           // [out, in] row-major (src/nn.rs:44)
    const float* b[kMaxL];           // may be NULL
    const float* x;                  // [batch, dims[0]]
    float* act[kMaxL];               // act[l]: output of layer l, [batch, dims[l + 1]]
};

__device__ __forceinline__ int pad4(int n) { return (n + 3) & ~3; }
__device__ __forceinline__ int wpitch(int out) { return pad4(out) + 4; }       // transposed-weight row pitch: 16-byte aligned, not a multiple of 32 banks

// Register tiles of 4 rows x 4 outputs (16 independent FMA chains per thread: the loops are latency-bound otherwise — one
// accumulator per thread measured 122 us for this kernel at batch 1024, against ~6 us of issue time).
__global__ void __launch_bounds__(kThreads)
mlp_small_fwd_kernel(const __grid_constant__ MlpArgs a) {
    extern __shared__ __align__(16) float sm[];
    // layout: transposed weights of every layer ([in][wpitch(out)], columns >= out zero), then two activation buffers [kRows][kMaxW]
    float* wt[kMaxL];
    float* p = sm;
    for (int l = 0; l < a.L; ++l) {
        wt[l] = p;
        p += a.dims[l] * wpitch(a.dims[l + 1]);
    }
    float* cur = p;
    float* nxt = p + kRows * kMaxW;
    const int r0 = blockIdx.x * kRows;
    const int rows = min(kRows, a.batch - r0);
    for (int l = 0; l < a.L; ++l) {
        const int in = a.dims[l], out = a.dims[l + 1], wp = wpitch(out);
        for (int i = threadIdx.x; i < in * wp; i += kThreads) wt[l][i] = 0.0f;
    }
    for (int i = threadIdx.x; i < 2 * kRows * kMaxW; i += kThreads) cur[i] = 0.0f;
    __syncthreads();
    for (int l = 0; l < a.L; ++l) {
        const int in = a.dims[l], out = a.dims[l + 1], wp = wpitch(out);
        // eight independent loads in flight per thread (one load per iteration leaves the kernel waiting on L2 latency)
#pragma unroll 8
        for (int i = threadIdx.x; i < in * out; i += kThreads) {
            const int o = i / in, k = i - o * in;
            wt[l][k * wp + o] = __ldg(a.W[l] + i);
        }
    }
    {
        const int in = a.dims[0];
#pragma unroll 8
        for (int i = threadIdx.x; i < rows * in; i += kThreads) {
            const int r = i / in, k = i - r * in;
            cur[r * kMaxW + k] = __ldg(a.x + (size_t)r0 * in + i);
        }
    }
    __syncthreads();
    for (int l = 0; l < a.L; ++l) {
        const int in = a.dims[l], out = a.dims[l + 1], wp = wpitch(out);
        const int og_n = pad4(out) / 4;
        for (int t = threadIdx.x; t < (kRows / 4) * og_n; t += kThreads) {
            const int rg = t / og_n, og = t - rg * og_n;
            const float* xr = cur + rg * 4 * kMaxW;
            const float* w = wt[l] + og * 4;
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
            for (int k = 0; k < in; ++k) {
                const float4 wv = *(const float4*)(w + k * wp);
                const float x0 = xr[k], x1 = xr[kMaxW + k], x2 = xr[2 * kMaxW + k], x3 = xr[3 * kMaxW + k];
                acc[0][0] = fmaf(x0, wv.x, acc[0][0]); acc[0][1] = fmaf(x0, wv.y, acc[0][1]); acc[0][2] = fmaf(x0, wv.z, acc[0][2]); acc[0][3] = fmaf(x0, wv.w, acc[0][3]);
                acc[1][0] = fmaf(x1, wv.x, acc[1][0]); acc[1][1] = fmaf(x1, wv.y, acc[1][1]); acc[1][2] = fmaf(x1, wv.z, acc[1][2]); acc[1][3] = fmaf(x1, wv.w, acc[1][3]);
                acc[2][0] = fmaf(x2, wv.x, acc[2][0]); acc[2][1] = fmaf(x2, wv.y, acc[2][1]); acc[2][2] = fmaf(x2, wv.z, acc[2][2]); acc[2][3] = fmaf(x2, wv.w, acc[2][3]);
                acc[3][0] = fmaf(x3, wv.x, acc[3][0]); acc[3][1] = fmaf(x3, wv.y, acc[3][1]); acc[3][2] = fmaf(x3, wv.z, acc[3][2]); acc[3][3] = fmaf(x3, wv.w, acc[3][3]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = rg * 4 + i;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int o = og * 4 + j;
                    if (o < out) {
                        float v = acc[i][j];
                        if (a.b[l]) v += __ldg(a.b[l] + o);
                        if (a.relu[l]) v = fmaxf(v, 0.0f);
                        nxt[r * kMaxW + o] = v;
                        if (r < rows) a.act[l][(size_t)(r0 + r) * out + o] = v;
                    }
                }
            }
        }
        __syncthreads();
        float* t2 = cur; cur = nxt; nxt = t2;
    }
}

struct MlpBwdArgs {
    int L, batch;
    int dims[kMaxL + 1];
    int relu[kMaxL];
    const float* W[kMaxL];
    const float* x;
    const float* act[kMaxL];
    const float* gout;               // [batch, dims[L]] gradient of the chain's output
    float* dx;                       // [batch, dims[0]] or NULL
    int acc_dx;
    float* partial;                  // [ctas][n_params]: per layer dW [out, in] then db [out]
    long long p_off[kMaxL];          // offset of layer l's dW inside a CTA's partial block (db follows at + out * in)
    long long n_params;
};

__global__ void __launch_bounds__(kThreads)
mlp_small_bwd_kernel(const __grid_constant__ MlpBwdArgs a) {
    extern __shared__ __align__(16) float sm[];
    // gz [kRows][kMaxW] (gradient of the current layer's pre-activation), gin [kRows][kMaxW], xin [kRows][kMaxW] (rows >= the CTA's
    // row count and columns beyond the widths stay zero), W [out][kMaxW]
    float* gz = sm;
    float* gin = gz + kRows * kMaxW;
    float* xin = gin + kRows * kMaxW;
    float* w = xin + kRows * kMaxW;
    const int r0 = blockIdx.x * kRows;
    const int rows = min(kRows, a.batch - r0);
    float* part = a.partial + (size_t)blockIdx.x * a.n_params;
    for (int i = threadIdx.x; i < 3 * kRows * kMaxW; i += kThreads) sm[i] = 0.0f;
    __syncthreads();
    {
        const int out = a.dims[a.L];
#pragma unroll 4
        for (int i = threadIdx.x; i < rows * out; i += kThreads) {
            const int r = i / out, o = i - r * out;
            float g = __ldg(a.gout + (size_t)r0 * out + i);
            if (a.relu[a.L - 1]) g = __ldg(a.act[a.L - 1] + (size_t)r0 * out + i) > 0.0f ? g : 0.0f;
            gz[r * kMaxW + o] = g;
        }
    }
    for (int l = a.L - 1; l >= 0; --l) {
        const int in = a.dims[l], out = a.dims[l + 1];
        const float* xin_g = l == 0 ? a.x : a.act[l - 1];
        for (int i = threadIdx.x; i < kRows * kMaxW; i += kThreads) xin[i] = 0.0f;
        __syncthreads();
#pragma unroll 8
        for (int i = threadIdx.x; i < rows * in; i += kThreads) {
            const int r = i / in, k = i - r * in;
            xin[r * kMaxW + k] = __ldg(xin_g + (size_t)r0 * in + i);
        }
#pragma unroll 8
        for (int i = threadIdx.x; i < out * kMaxW; i += kThreads) {
            const int o = i / kMaxW, k = i - o * kMaxW;
            w[i] = k < in ? __ldg(a.W[l] + (size_t)o * in + k) : 0.0f;
        }
        __syncthreads();
        // dW[o][k] partial = sum_r gz[r][o] * xin[r][k]   (src/ops.rs:280-291): 4 x 4 register tiles over (o, k)
        {
            const int kg_n = pad4(in) / 4, og_n = pad4(out) / 4;
            for (int t = threadIdx.x; t < og_n * kg_n; t += kThreads) {
                const int og = t / kg_n, kg = t - og * kg_n;
                float acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
                for (int r = 0; r < kRows; ++r) {
                    const float4 g4 = *(const float4*)(gz + r * kMaxW + og * 4);
                    const float4 x4 = *(const float4*)(xin + r * kMaxW + kg * 4);
                    acc[0][0] = fmaf(g4.x, x4.x, acc[0][0]); acc[0][1] = fmaf(g4.x, x4.y, acc[0][1]); acc[0][2] = fmaf(g4.x, x4.z, acc[0][2]); acc[0][3] = fmaf(g4.x, x4.w, acc[0][3]);
                    acc[1][0] = fmaf(g4.y, x4.x, acc[1][0]); acc[1][1] = fmaf(g4.y, x4.y, acc[1][1]); acc[1][2] = fmaf(g4.y, x4.z, acc[1][2]); acc[1][3] = fmaf(g4.y, x4.w, acc[1][3]);
                    acc[2][0] = fmaf(g4.z, x4.x, acc[2][0]); acc[2][1] = fmaf(g4.z, x4.y, acc[2][1]); acc[2][2] = fmaf(g4.z, x4.z, acc[2][2]); acc[2][3] = fmaf(g4.z, x4.w, acc[2][3]);
                    acc[3][0] = fmaf(g4.w, x4.x, acc[3][0]); acc[3][1] = fmaf(g4.w, x4.y, acc[3][1]); acc[3][2] = fmaf(g4.w, x4.z, acc[3][2]); acc[3][3] = fmaf(g4.w, x4.w, acc[3][3]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int o = og * 4 + i, k = kg * 4 + j;
                        if (o < out && k < in) part[a.p_off[l] + (long long)o * in + k] = acc[i][j];
                    }
            }
        }
        // db[o] = sum_r gz[r][o]  (src/tensor.rs:680-691)
        for (int o = threadIdx.x; o < out; o += kThreads) {
            float acc = 0.0f;
            for (int r = 0; r < kRows; ++r) acc += gz[r * kMaxW + o];
            part[a.p_off[l] + (long long)out * in + o] = acc;
        }
        // gradient of the layer's input: gin[r][k] = sum_o gz[r][o] * W[o][k]  (src/ops.rs:254-265), then the previous ReLU's mask
        if (l > 0 || a.dx) {
            const int kg_n = pad4(in) / 4;
            for (int t = threadIdx.x; t < (kRows / 4) * kg_n; t += kThreads) {
                const int rg = t / kg_n, kg = t - rg * kg_n;
                float acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
                const float* g = gz + rg * 4 * kMaxW;
                for (int o = 0; o < out; ++o) {
                    const float4 w4 = *(const float4*)(w + o * kMaxW + kg * 4);
                    const float g0 = g[o], g1 = g[kMaxW + o], g2 = g[2 * kMaxW + o], g3 = g[3 * kMaxW + o];
                    acc[0][0] = fmaf(g0, w4.x, acc[0][0]); acc[0][1] = fmaf(g0, w4.y, acc[0][1]); acc[0][2] = fmaf(g0, w4.z, acc[0][2]); acc[0][3] = fmaf(g0, w4.w, acc[0][3]);
                    acc[1][0] = fmaf(g1, w4.x, acc[1][0]); acc[1][1] = fmaf(g1, w4.y, acc[1][1]); acc[1][2] = fmaf(g1, w4.z, acc[1][2]); acc[1][3] = fmaf(g1, w4.w, acc[1][3]);
                    acc[2][0] = fmaf(g2, w4.x, acc[2][0]); acc[2][1] = fmaf(g2, w4.y, acc[2][1]); acc[2][2] = fmaf(g2, w4.z, acc[2][2]); acc[2][3] = fmaf(g2, w4.w, acc[2][3]);
                    acc[3][0] = fmaf(g3, w4.x, acc[3][0]); acc[3][1] = fmaf(g3, w4.y, acc[3][1]); acc[3][2] = fmaf(g3, w4.z, acc[3][2]); acc[3][3] = fmaf(g3, w4.w, acc[3][3]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int r = rg * 4 + i, k = kg * 4 + j;
                        if (k < in) {
                            float v = acc[i][j];
                            if (l > 0 && a.relu[l - 1]) v = xin[r * kMaxW + k] > 0.0f ? v : 0.0f;
                            gin[r * kMaxW + k] = v;
                            if (l == 0 && r < rows) {
                                float* d = a.dx + (size_t)(r0 + r) * in + k;
                                *d = a.acc_dx ? *d + v : v;
                            }
                        }
                    }
            }
        }
        __syncthreads();
        float* t2 = gz; gz = gin; gin = t2;
        // the next layer's gz columns beyond its width must be zero (they feed 4-wide tiles): the buffer that becomes gin is
        // rewritten per element below, the one that became gz was written for k < in only over zeros of the same or a wider layer
        if (l > 0) {
            const int nin = a.dims[l];                           // width of the new gz
            for (int i = threadIdx.x; i < kRows * kMaxW; i += kThreads) {
                const int k = i & (kMaxW - 1);
                if (k >= nin) gz[i] = 0.0f;
            }
            __syncthreads();
        }
    }
}

// dst[i] (+)= sum over ctas of partial[cta][off + i], ctas in order; one launch for all tensors
struct FoldArgs {
    const float* partial;
    long long n_params;
    int ctas, count;
    float* dst[2 * kMaxL];
    long long off[2 * kMaxL], len[2 * kMaxL];
    int acc[2 * kMaxL];
};
__global__ void __launch_bounds__(kThreads)
mlp_small_fold_kernel(const __grid_constant__ FoldArgs f) {
    const int t = blockIdx.y;
    if (t >= f.count || !f.dst[t]) return;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < f.len[t]; i += (long long)gridDim.x * kThreads) {
        const float* src = f.partial + f.off[t] + i;
        float s = 0.0f;
        int c = 0;
        for (; c + 8 <= f.ctas; c += 8) {                     // eight loads in flight; the additions stay in CTA order
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldg(src + (size_t)(c + u) * f.n_params);
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
        for (; c < f.ctas; ++c) s += __ldg(src + (size_t)c * f.n_params);
        f.dst[t][i] = f.acc[t] ? f.dst[t][i] + s : s;
    }
}

bool shapes_ok(int L, const int* dims, int batch, size_t* fwd_smem, size_t* bwd_smem) {
    if (L < 1 || L > kMaxL || batch < 1) return false;
    size_t wsum = 0, wmax = 0;
    for (int l = 0; l <= L; ++l)
        if (dims[l] < 1 || dims[l] > kMaxW) return false;
    for (int l = 0; l < L; ++l) {
        wsum += (size_t)dims[l] * (((dims[l + 1] + 3) & ~3) + 4);
        const size_t w = (size_t)dims[l + 1] * kMaxW;
        if (w > wmax) wmax = w;
    }
    *fwd_smem = (wsum + 2 * kRows * kMaxW) * sizeof(float);
    *bwd_smem = (3 * kRows * kMaxW + wmax) * sizeof(float);
    return *fwd_smem <= 200 * 1024 && *bwd_smem <= 200 * 1024;
}

struct TmpBuf {
    tp_buf* b = nullptr;
    ~TmpBuf() { if (b) tp_buf_release(b); }
};

}  // namespace

extern "C" {

int tp_mlp_small_supported(int n_layers, const int* dims, int batch) {
    size_t a, b;
    return dims && shapes_ok(n_layers, dims, batch, &a, &b) ? 1 : 0;
}

int tp_mlp_small_fwd(tp_ctx* ctx, const tp_buf* x, int n_layers, const int* dims, const tp_buf* const* weights,
                     const tp_buf* const* biases, const int* relu, tp_buf* const* acts, int batch) {
    TP_CHECK_ARG(ctx && x && dims && weights && biases && relu && acts, "tp_mlp_small_fwd: NULL argument");
    size_t fs, bs;
    TP_CHECK_ARG(shapes_ok(n_layers, dims, batch, &fs, &bs), "tp_mlp_small_fwd: 1..%d layers of width <= %d", kMaxL, kMaxW);
    TP_NEED(x, (size_t)batch * dims[0], "x");
    MlpArgs a{};
    a.L = n_layers; a.batch = batch;
    for (int l = 0; l <= n_layers; ++l) a.dims[l] = dims[l];
    for (int l = 0; l < n_layers; ++l) {
        TP_NEED(weights[l], (size_t)dims[l] * dims[l + 1], "weight");
        if (biases[l]) TP_NEED(biases[l], (size_t)dims[l + 1], "bias");
        TP_NEED(acts[l], (size_t)batch * dims[l + 1], "activation");
        a.W[l] = weights[l]->ptr;
        a.b[l] = biases[l] ? biases[l]->ptr : nullptr;
        a.relu[l] = relu[l] ? 1 : 0;
        a.act[l] = acts[l]->ptr;
    }
    a.x = x->ptr;
    cudaSetDevice(ctx->device);
    static int attr[16] = {};
    const int dev = ctx->device < 16 ? ctx->device : 15;
    if (ctx->device >= 16 || attr[dev] < (int)fs) {
        TP_CUDA(cudaFuncSetAttribute(mlp_small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fs));
        attr[dev] = (int)fs;
    }
    mlp_small_fwd_kernel<<<(batch + kRows - 1) / kRows, kThreads, fs, ctx->stream>>>(a);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

/* gout: gradient of the last layer's output (after its ReLU if relu[L-1]); dw[l] / db[l] / dx may be NULL; acc_*: 0 store, 1 add */
int tp_mlp_small_bwd(tp_ctx* ctx, const tp_buf* x, int n_layers, const int* dims, const tp_buf* const* weights, const int* relu,
                     const tp_buf* const* acts, const tp_buf* gout, tp_buf* dx, tp_buf* const* dw, tp_buf* const* db, int acc_dx,
                     const int* acc_dw, const int* acc_db, int batch) {
    TP_CHECK_ARG(ctx && x && dims && weights && relu && acts && gout && dw && db && acc_dw && acc_db, "tp_mlp_small_bwd: NULL argument");
    size_t fs, bs;
    TP_CHECK_ARG(shapes_ok(n_layers, dims, batch, &fs, &bs), "tp_mlp_small_bwd: 1..%d layers of width <= %d", kMaxL, kMaxW);
    TP_NEED(x, (size_t)batch * dims[0], "x"); TP_NEED(gout, (size_t)batch * dims[n_layers], "gout");
    if (dx) TP_NEED(dx, (size_t)batch * dims[0], "dx");
    MlpBwdArgs a{};
    a.L = n_layers; a.batch = batch;
    for (int l = 0; l <= n_layers; ++l) a.dims[l] = dims[l];
    long long np = 0;
    for (int l = 0; l < n_layers; ++l) {
        TP_NEED(weights[l], (size_t)dims[l] * dims[l + 1], "weight");
        TP_NEED(acts[l], (size_t)batch * dims[l + 1], "activation");
        if (dw[l]) TP_NEED(dw[l], (size_t)dims[l] * dims[l + 1], "dw");
        if (db[l]) TP_NEED(db[l], (size_t)dims[l + 1], "db");
        a.W[l] = weights[l]->ptr;
        a.act[l] = acts[l]->ptr;
        a.relu[l] = relu[l] ? 1 : 0;
        a.p_off[l] = np;
        np += (long long)dims[l] * dims[l + 1] + dims[l + 1];
    }
    a.n_params = np;
    a.x = x->ptr;
    a.gout = gout->ptr;
    a.dx = dx ? dx->ptr : nullptr;
    a.acc_dx = acc_dx;
    const int ctas = (batch + kRows - 1) / kRows;
    TmpBuf part;
    int rc = tp_buf_alloc(ctx, (size_t)ctas * np, &part.b);
    if (rc) return rc;
    a.partial = part.b->ptr;
    cudaSetDevice(ctx->device);
    static int attr[16] = {};
    const int dev = ctx->device < 16 ? ctx->device : 15;
    if (ctx->device >= 16 || attr[dev] < (int)bs) {
        TP_CUDA(cudaFuncSetAttribute(mlp_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bs));
        attr[dev] = (int)bs;
    }
    mlp_small_bwd_kernel<<<ctas, kThreads, bs, ctx->stream>>>(a);
    TP_LAUNCH_OK(ctx);
    FoldArgs f{};
    f.partial = part.b->ptr;
    f.n_params = np;
    f.ctas = ctas;
    f.count = 2 * n_layers;
    for (int l = 0; l < n_layers; ++l) {
        f.dst[2 * l] = dw[l] ? dw[l]->ptr : nullptr;
        f.off[2 * l] = a.p_off[l];
        f.len[2 * l] = (long long)dims[l] * dims[l + 1];
        f.acc[2 * l] = acc_dw[l];
        f.dst[2 * l + 1] = db[l] ? db[l]->ptr : nullptr;
        f.off[2 * l + 1] = a.p_off[l] + (long long)dims[l] * dims[l + 1];
        f.len[2 * l + 1] = dims[l + 1];
        f.acc[2 * l + 1] = acc_db[l];
    }
    mlp_small_fold_kernel<<<dim3(64, 2 * n_layers), kThreads, 0, ctx->stream>>>(f);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

}  // extern "C"

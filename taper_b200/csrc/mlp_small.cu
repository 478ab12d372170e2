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
//
// A CTA owns kRows = 16 consecutive batch rows (64 CTAs at batch 1024).  Every layer's weights are brought into shared memory
// once per CTA with 16-byte cp.async copies in their own [out][in] layout (row pitch in + 4 floats: a quarter-warp's 128-bit
// reads of eight consecutive rows fall into eight different bank groups), all in flight at once — the first version staged
// them through registers with a transpose and spent most of its 30 us waiting on those loads with 8 warps per SM.
// Forward: register tiles of 2 rows x 4 outputs walk k four at a time (128-bit shared loads of the activations, broadcast
// within a warp, and of four weight rows); every layer's output is written to global (the backward needs it).  Backward: per
// layer the dW / db partials of the CTA's rows (4 x 4 tiles over (o, k), contraction over the 16 rows) and the gradient of
// the layer's input (2 rows x 4 inputs, masked by the previous ReLU); a fold kernel sums the per-CTA partials in CTA order
// (deterministic).  Products are accumulated in ascending k with fmaf, bias added last — the order of the reference's
// matmul + add_broadcast.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kRows = 16;            // batch rows per CTA
constexpr int kMaxL = 4;
constexpr int kMaxW = 128;
constexpr int kXP = kMaxW + 4;       // row pitch of the staged activations / gradients (floats)

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
__device__ __forceinline__ int wpitch(int in) { return pad4(in) + 4; }         // weight row pitch in shared memory (floats)
__device__ __forceinline__ int wrows(int out) { return pad4(out); }            // rows held per layer (rows >= out are zero)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// weights of layer (in, out) at `w` -> dst[pad4(out)][wpitch(in)]; pad columns and pad rows zero
__device__ __forceinline__ void stage_weights(float* dst, const float* __restrict__ w, int in, int out) {
    const int wp = wpitch(in), in4 = pad4(in);
    if ((in & 3) == 0 && (((uintptr_t)w) & 15) == 0) {
        const int vec = in >> 2;
        for (int i = threadIdx.x; i < out * vec; i += kThreads) {
            const int o = i / vec, v = i - o * vec;
            cp_async16(dst + o * wp + 4 * v, w + (size_t)o * in + 4 * v);
        }
    } else {
#pragma unroll 4
        for (int i = threadIdx.x; i < out * in4; i += kThreads) {
            const int o = i / in4, k = i - o * in4;
            dst[o * wp + k] = k < in ? __ldg(w + (size_t)o * in + k) : 0.0f;
        }
    }
    for (int i = threadIdx.x; i < (wrows(out) - out) * in4; i += kThreads) {
        const int o = out + i / in4, k = i % in4;
        dst[o * wp + k] = 0.0f;
    }
}

// `rows` rows of a [batch, width] matrix -> dst[kRows][kXP]; pad columns (up to pad4(width)) and rows >= rows zero
__device__ __forceinline__ void stage_rows(float* dst, const float* __restrict__ src, int rows, int width) {
    const int w4 = pad4(width);
    if ((width & 3) == 0 && (((uintptr_t)src) & 15) == 0) {
        const int vec = width >> 2;
        for (int i = threadIdx.x; i < kRows * vec; i += kThreads) {
            const int r = i / vec, v = i - r * vec;
            if (r < rows) cp_async16(dst + r * kXP + 4 * v, src + (size_t)r * width + 4 * v);
            else *(float4*)(dst + r * kXP + 4 * v) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {
        for (int i = threadIdx.x; i < kRows * w4; i += kThreads) {
            const int r = i / w4, k = i - r * w4;
            dst[r * kXP + k] = (r < rows && k < width) ? __ldg(src + (size_t)r * width + k) : 0.0f;
        }
    }
}

#define TP_FMA4(acc, xv, wv)                 \
    do {                                     \
        acc = fmaf((xv).x, (wv).x, acc);     \
        acc = fmaf((xv).y, (wv).y, acc);     \
        acc = fmaf((xv).z, (wv).z, acc);     \
        acc = fmaf((xv).w, (wv).w, acc);     \
    } while (0)

__global__ void __launch_bounds__(kThreads)
mlp_small_fwd_kernel(const __grid_constant__ MlpArgs a) {
    extern __shared__ __align__(16) float sm[];
    // layout: the weights of every layer ([pad4(out)][wpitch(in)]), then two activation buffers [kRows][kXP]
    float* ws[kMaxL];
    float* p = sm;
    for (int l = 0; l < a.L; ++l) {
        ws[l] = p;
        p += wrows(a.dims[l + 1]) * wpitch(a.dims[l]);
    }
    float* cur = p;
    float* nxt = p + kRows * kXP;
    const int r0 = blockIdx.x * kRows;
    const int rows = min(kRows, a.batch - r0);
    for (int l = 0; l < a.L; ++l) stage_weights(ws[l], a.W[l], a.dims[l], a.dims[l + 1]);
    stage_rows(cur, a.x + (size_t)r0 * a.dims[0], rows, a.dims[0]);
    cp_async_wait_all();
    __syncthreads();
    for (int l = 0; l < a.L; ++l) {
        const int in4 = pad4(a.dims[l]), out = a.dims[l + 1], wp = wpitch(a.dims[l]);
        const int og_n = pad4(out) >> 2;                                         // a thread's four outputs: og + og_n * j
        const float* bias = a.b[l];
        for (int t = threadIdx.x; t < (kRows / 2) * og_n; t += kThreads) {
            const int rg = t / og_n, og = t - rg * og_n;
            const float* x0 = cur + (2 * rg) * kXP;
            const float* x1 = x0 + kXP;
            const float* w0 = ws[l] + (og) * wp;                                 // rows >= out exist (zero) up to pad4(out)
            const float* w1 = ws[l] + (og + og_n) * wp;
            const float* w2 = ws[l] + (og + 2 * og_n) * wp;
            const float* w3 = ws[l] + (og + 3 * og_n) * wp;
            float acc[2][4], bj[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)                               // requested before the k loop: the round trip hides behind it
                bj[j] = (bias && og + og_n * j < out) ? __ldg(bias + og + og_n * j) : 0.0f;
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
#pragma unroll 2
            for (int k = 0; k < in4; k += 4) {
                const float4 a0 = *(const float4*)(x0 + k), a1 = *(const float4*)(x1 + k);
                const float4 v0 = *(const float4*)(w0 + k), v1 = *(const float4*)(w1 + k);
                const float4 v2 = *(const float4*)(w2 + k), v3 = *(const float4*)(w3 + k);
                TP_FMA4(acc[0][0], a0, v0); TP_FMA4(acc[0][1], a0, v1); TP_FMA4(acc[0][2], a0, v2); TP_FMA4(acc[0][3], a0, v3);
                TP_FMA4(acc[1][0], a1, v0); TP_FMA4(acc[1][1], a1, v1); TP_FMA4(acc[1][2], a1, v2); TP_FMA4(acc[1][3], a1, v3);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int o = og + og_n * j;
                if (o >= out) continue;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int r = 2 * rg + i;
                    float v = acc[i][j];
                    if (bias) v += bj[j];
                    if (a.relu[l]) v = fmaxf(v, 0.0f);
                    nxt[r * kXP + o] = v;
                    if (r < rows) a.act[l][(size_t)(r0 + r) * out + o] = v;
                }
            }
        }
        // pad columns of the next layer's input must be zero
        for (int i = threadIdx.x; i < kRows * (pad4(out) - out); i += kThreads) {
            const int r = i / (pad4(out) - out), k = out + i % (pad4(out) - out);
            nxt[r * kXP + k] = 0.0f;
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
    // the weights of every layer ([pad4(out)][wpitch(in)]), then gz (gradient of the current layer's pre-activation), gin, xin:
    // [kRows][kXP] each; rows >= the CTA's row count and pad columns are zero
    float* ws[kMaxL];
    float* p = sm;
    for (int l = 0; l < a.L; ++l) {
        ws[l] = p;
        p += wrows(a.dims[l + 1]) * wpitch(a.dims[l]);
    }
    float* gz = p;
    float* gin = gz + kRows * kXP;
    float* xin = gin + kRows * kXP;
    const int r0 = blockIdx.x * kRows;
    const int rows = min(kRows, a.batch - r0);
    float* part = a.partial + (size_t)blockIdx.x * a.n_params;
    for (int l = 0; l < a.L; ++l) stage_weights(ws[l], a.W[l], a.dims[l], a.dims[l + 1]);
    {
        const int out = a.dims[a.L], o4 = pad4(out);
        const bool relu = a.relu[a.L - 1] != 0;
        for (int i = threadIdx.x; i < kRows * o4; i += kThreads) {
            const int r = i / o4, o = i - r * o4;
            float g = 0.0f;
            if (r < rows && o < out) {
                g = __ldg(a.gout + (size_t)(r0 + r) * out + o);
                if (relu) g = __ldg(a.act[a.L - 1] + (size_t)(r0 + r) * out + o) > 0.0f ? g : 0.0f;
            }
            gz[r * kXP + o] = g;
        }
    }
    for (int l = a.L - 1; l >= 0; --l) {
        const int in = a.dims[l], out = a.dims[l + 1], in4 = pad4(in), out4 = pad4(out), wp = wpitch(in);
        const float* xin_g = l == 0 ? a.x : a.act[l - 1];
        stage_rows(xin, xin_g + (size_t)r0 * in, rows, in);
        cp_async_wait_all();
        __syncthreads();
        // dW[o][k] partial = sum_r gz[r][o] * xin[r][k]   (src/ops.rs:280-291): 4 x 4 register tiles over (o, k)
        {
            const int kg_n = in4 >> 2, og_n = out4 >> 2;
            float* dst = part + a.p_off[l];
            for (int t = threadIdx.x; t < og_n * kg_n; t += kThreads) {
                const int og = t / kg_n, kg = t - og * kg_n;
                float acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
#pragma unroll 4
                for (int r = 0; r < kRows; ++r) {
                    const float4 g4 = *(const float4*)(gz + r * kXP + og * 4);
                    const float4 x4 = *(const float4*)(xin + r * kXP + kg * 4);
                    acc[0][0] = fmaf(g4.x, x4.x, acc[0][0]); acc[0][1] = fmaf(g4.x, x4.y, acc[0][1]); acc[0][2] = fmaf(g4.x, x4.z, acc[0][2]); acc[0][3] = fmaf(g4.x, x4.w, acc[0][3]);
                    acc[1][0] = fmaf(g4.y, x4.x, acc[1][0]); acc[1][1] = fmaf(g4.y, x4.y, acc[1][1]); acc[1][2] = fmaf(g4.y, x4.z, acc[1][2]); acc[1][3] = fmaf(g4.y, x4.w, acc[1][3]);
                    acc[2][0] = fmaf(g4.z, x4.x, acc[2][0]); acc[2][1] = fmaf(g4.z, x4.y, acc[2][1]); acc[2][2] = fmaf(g4.z, x4.z, acc[2][2]); acc[2][3] = fmaf(g4.z, x4.w, acc[2][3]);
                    acc[3][0] = fmaf(g4.w, x4.x, acc[3][0]); acc[3][1] = fmaf(g4.w, x4.y, acc[3][1]); acc[3][2] = fmaf(g4.w, x4.z, acc[3][2]); acc[3][3] = fmaf(g4.w, x4.w, acc[3][3]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int o = og * 4 + i;
                    if (o >= out) continue;
                    if ((in & 3) == 0) {
                        *(float4*)(dst + (long long)o * in + kg * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (kg * 4 + j < in) dst[(long long)o * in + kg * 4 + j] = acc[i][j];
                    }
                }
            }
        }
        // db[o] = sum_r gz[r][o]  (src/tensor.rs:680-691)
        for (int o = threadIdx.x; o < out; o += kThreads) {
            float acc = 0.0f;
#pragma unroll
            for (int r = 0; r < kRows; ++r) acc += gz[r * kXP + o];
            part[a.p_off[l] + (long long)out * in + o] = acc;
        }
        // gradient of the layer's input: gin[r][k] = sum_o gz[r][o] * W[o][k]  (src/ops.rs:254-265), then the previous ReLU's
        // mask: 2 rows x 4 inputs per thread, o four at a time (rows >= out of the staged weights are zero)
        if (l > 0 || a.dx) {
            const int kg_n = in4 >> 2;
            const bool mask = l > 0 && a.relu[l - 1] != 0;
            for (int t = threadIdx.x; t < (kRows / 2) * kg_n; t += kThreads) {
                const int rg = t / kg_n, kg = t - rg * kg_n;
                const float* g0 = gz + (2 * rg) * kXP;
                const float* g1 = g0 + kXP;
                const float* w = ws[l] + kg * 4;
                float acc[2][4];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
#pragma unroll 2
                for (int o = 0; o < out4; o += 4) {
                    const float4 ga = *(const float4*)(g0 + o), gb = *(const float4*)(g1 + o);
                    const float4 wa = *(const float4*)(w + (o) * wp), wb = *(const float4*)(w + (o + 1) * wp);
                    const float4 wc = *(const float4*)(w + (o + 2) * wp), wd = *(const float4*)(w + (o + 3) * wp);
                    acc[0][0] = fmaf(ga.x, wa.x, acc[0][0]); acc[0][1] = fmaf(ga.x, wa.y, acc[0][1]); acc[0][2] = fmaf(ga.x, wa.z, acc[0][2]); acc[0][3] = fmaf(ga.x, wa.w, acc[0][3]);
                    acc[1][0] = fmaf(gb.x, wa.x, acc[1][0]); acc[1][1] = fmaf(gb.x, wa.y, acc[1][1]); acc[1][2] = fmaf(gb.x, wa.z, acc[1][2]); acc[1][3] = fmaf(gb.x, wa.w, acc[1][3]);
                    acc[0][0] = fmaf(ga.y, wb.x, acc[0][0]); acc[0][1] = fmaf(ga.y, wb.y, acc[0][1]); acc[0][2] = fmaf(ga.y, wb.z, acc[0][2]); acc[0][3] = fmaf(ga.y, wb.w, acc[0][3]);
                    acc[1][0] = fmaf(gb.y, wb.x, acc[1][0]); acc[1][1] = fmaf(gb.y, wb.y, acc[1][1]); acc[1][2] = fmaf(gb.y, wb.z, acc[1][2]); acc[1][3] = fmaf(gb.y, wb.w, acc[1][3]);
                    acc[0][0] = fmaf(ga.z, wc.x, acc[0][0]); acc[0][1] = fmaf(ga.z, wc.y, acc[0][1]); acc[0][2] = fmaf(ga.z, wc.z, acc[0][2]); acc[0][3] = fmaf(ga.z, wc.w, acc[0][3]);
                    acc[1][0] = fmaf(gb.z, wc.x, acc[1][0]); acc[1][1] = fmaf(gb.z, wc.y, acc[1][1]); acc[1][2] = fmaf(gb.z, wc.z, acc[1][2]); acc[1][3] = fmaf(gb.z, wc.w, acc[1][3]);
                    acc[0][0] = fmaf(ga.w, wd.x, acc[0][0]); acc[0][1] = fmaf(ga.w, wd.y, acc[0][1]); acc[0][2] = fmaf(ga.w, wd.z, acc[0][2]); acc[0][3] = fmaf(ga.w, wd.w, acc[0][3]);
                    acc[1][0] = fmaf(gb.w, wd.x, acc[1][0]); acc[1][1] = fmaf(gb.w, wd.y, acc[1][1]); acc[1][2] = fmaf(gb.w, wd.z, acc[1][2]); acc[1][3] = fmaf(gb.w, wd.w, acc[1][3]);
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int r = 2 * rg + i;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int k = kg * 4 + j;
                        float v = acc[i][j];
                        if (mask) v = xin[r * kXP + k] > 0.0f ? v : 0.0f;
                        if (k >= in) v = 0.0f;
                        gin[r * kXP + k] = v;
                        if (l == 0 && r < rows && k < in) {
                            float* d = a.dx + (size_t)(r0 + r) * in + k;
                            *d = a.acc_dx ? *d + v : v;
                        }
                    }
                }
            }
        }
        __syncthreads();
        float* t2 = gz; gz = gin; gin = t2;
    }
}

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
    size_t wsum = 0;
    for (int l = 0; l <= L; ++l)
        if (dims[l] < 1 || dims[l] > kMaxW) return false;
    for (int l = 0; l < L; ++l) wsum += (size_t)((dims[l + 1] + 3) & ~3) * (((dims[l] + 3) & ~3) + 4);
    *fwd_smem = (wsum + 2 * kRows * kXP) * sizeof(float);
    *bwd_smem = (wsum + 3 * kRows * kXP) * sizeof(float);
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
        a.p_off[l] = np;                          // a multiple of 4: the kernel stores dW partial rows as float4 when in % 4 == 0
        np += (long long)dims[l] * dims[l + 1] + dims[l + 1];
        np = (np + 3) & ~3LL;
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

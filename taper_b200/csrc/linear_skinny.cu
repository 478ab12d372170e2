// Linear layers with a skinny output (out_features <= 16: the 10-logit classifier head of every MNIST model).
// A 128-wide tensor-core tile is pointless there (N = 10), and the three backward products are ~1 MFLOP each, so
// launch count and latency are what matter: forward is one launch, the whole backward (dX = dY*W, dW = dY^T*X,
// db = colsum(dY), optional ReLU mask on dY) is ONE launch with two block roles, deterministic (fixed fold order).
// Exact fp32 FMA arithmetic on the CUDA cores.  Reference: Linear::forward (src/nn.rs:54-60) and the backward closures
// of matmul / transpose / add_broadcast (src/ops.rs:238-294, src/tensor.rs:575-586, 680-691).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxOut = 16;
constexpr int kColsPerBlock = 64, kRowGroups = kThreads / kColsPerBlock;      // dW role of the backward kernel

// ---- forward: one warp per row.  W (out x in, <= 96 KB) is staged once per block in shared memory with 128-bit loads that
// are all in flight together, then every warp streams its row of X (float4) against it; scalar path for in_features % 4 != 0
// or unaligned operands --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
skinny_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                  float* __restrict__ y, int batch, int in_f, int out_f, int relu, int vec) {
    extern __shared__ __align__(16) float sw[];                // [out_f][in_f]
    const int nw = out_f * in_f;
    if (vec) {
        const float4* w4 = reinterpret_cast<const float4*>(w);
        float4* s4 = reinterpret_cast<float4*>(sw);
        for (int i0 = threadIdx.x; i0 < nw / 4; i0 += kThreads * 8) {
            float4 q[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u * kThreads;
                q[u] = i < nw / 4 ? __ldg(w4 + i) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u * kThreads;
                if (i < nw / 4) s4[i] = q[u];
            }
        }
    } else {
        for (int i = threadIdx.x; i < nw; i += kThreads) sw[i] = __ldg(w + i);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * kThreads + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * kThreads) >> 5;
    for (int r = warp; r < batch; r += nwarps) {
        const float* xr = x + (size_t)r * in_f;
        float acc[kMaxOut];
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o) acc[o] = 0.0f;
        if (vec) {
            const int in4 = in_f >> 2;
#pragma unroll 4
            for (int c = lane; c < in4; c += 32) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(xr) + c);
#pragma unroll
                for (int o = 0; o < kMaxOut; ++o) {
                    if (o < out_f) {
                        const float4 wv = *(reinterpret_cast<const float4*>(sw + (size_t)o * in_f) + c);
                        acc[o] = fmaf(a.x, wv.x, acc[o]);
                        acc[o] = fmaf(a.y, wv.y, acc[o]);
                        acc[o] = fmaf(a.z, wv.z, acc[o]);
                        acc[o] = fmaf(a.w, wv.w, acc[o]);
                    }
                }
            }
        } else {
            for (int k = lane; k < in_f; k += 32) {
                const float xv = __ldg(xr + k);
#pragma unroll
                for (int o = 0; o < kMaxOut; ++o)
                    if (o < out_f) acc[o] = fmaf(xv, sw[o * in_f + k], acc[o]);
            }
        }
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o) {
            if (o < out_f) {
                float v = acc[o];
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
                acc[o] = v;
            }
        }
        if (lane < out_f) {
            float v = 0.0f;
#pragma unroll
            for (int o = 0; o < kMaxOut; ++o) if (o == lane) v = acc[o];
            if (bias) v += __ldg(bias + lane);
            if (relu) v = fmaxf(v, 0.0f);
            y[(size_t)r * out_f + lane] = v;
        }
    }
}

// ---- backward: blocks [0, dx_blocks) compute dX rows, the rest compute dW/db partials over row splits ------------
struct BwdArgs {
    const float* x;       // [B, in]
    const float* w;       // [out, in]
    const float* dy;      // [B, out]
    const float* mask_y;  // optional [B, out]: dy := dy * [y > 0]
    float* dx;            // optional [B, in]
    float* dw;            // optional [out, in]
    float* db;            // optional [out]
    float* partial;       // [col_blocks][row_splits][out + 1... ] see below
    int* tickets;         // one per column block
    int batch, in_f, out_f;
    int dx_blocks, col_blocks, row_splits, rows_per_split;
    int acc_dx, acc_dw, acc_db;
};

__device__ __forceinline__ float masked(const BwdArgs& a, int r, int o) {
    float g = __ldg(a.dy + (size_t)r * a.out_f + o);
    if (a.mask_y && !(__ldg(a.mask_y + (size_t)r * a.out_f + o) > 0.0f)) g = 0.0f;
    return g;
}

__global__ void __launch_bounds__(kThreads)
skinny_bwd_kernel(BwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    __shared__ int s_last;
    const int tid = threadIdx.x;
    if ((int)blockIdx.x < a.dx_blocks) {
        // dX[r, i] (+)= sum_o dY[r, o] * W[o, i]      (N,N; src/ops.rs:254-265 composed with transpose bwd)
        float* sw = smem;                                      // [out][in]
        if (a.in_f % 4 == 0 && !((uintptr_t)a.w & 15)) {       // 128-bit staging loads, eight in flight per thread
            const int n4 = a.out_f * a.in_f / 4;
            const float4* w4 = reinterpret_cast<const float4*>(a.w);
            float4* s4 = reinterpret_cast<float4*>(sw);
            for (int i0 = tid; i0 < n4; i0 += kThreads * 8) {
                float4 q[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) q[u] = (i0 + u * kThreads < n4) ? __ldg(w4 + i0 + u * kThreads) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
                for (int u = 0; u < 8; ++u) if (i0 + u * kThreads < n4) s4[i0 + u * kThreads] = q[u];
            }
        } else {
            for (int i = tid; i < a.out_f * a.in_f; i += kThreads) sw[i] = __ldg(a.w + i);
        }
        __syncthreads();
        const int lane = tid & 31;
        const int warp = (blockIdx.x * kThreads + tid) >> 5;
        const int nwarps = (a.dx_blocks * kThreads) >> 5;
        for (int r = warp; r < a.batch; r += nwarps) {
            float g[kMaxOut];
#pragma unroll
            for (int o = 0; o < kMaxOut; ++o) g[o] = o < a.out_f ? masked(a, r, o) : 0.0f;
            float* dxr = a.dx + (size_t)r * a.in_f;
            for (int i = lane; i < a.in_f; i += 32) {
                float v = 0.0f;
#pragma unroll
                for (int o = 0; o < kMaxOut; ++o)
                    if (o < a.out_f) v = fmaf(g[o], sw[o * a.in_f + i], v);
                dxr[i] = a.acc_dx ? dxr[i] + v : v;
            }
        }
        return;
    }
    // dW[o, i] (+)= sum_r dY[r, o] * X[r, i] ;  db[o] (+)= sum_r dY[r, o]      (T,N; src/ops.rs:280-291, src/tensor.rs:680-691)
    // A block owns kColsPerBlock (64) input columns of one row split; its 256 threads are 64 columns x 4 row groups, each
    // thread walking its rows 8 at a time (8 independent loads in flight), the row groups folded through shared memory.
    const int b = blockIdx.x - a.dx_blocks;
    const int cb = b % a.col_blocks, rs = b / a.col_blocks;
    const int cg = tid & (kColsPerBlock - 1), rg = tid / kColsPerBlock;
    const int i = cb * kColsPerBlock + cg;                     // input column owned by this thread
    const int r0 = rs * a.rows_per_split, r1 = min(a.batch, r0 + a.rows_per_split);
    float* sdy = smem;                                         // [rows_per_split][out]
    float* red = smem + (size_t)a.rows_per_split * a.out_f;    // [kRowGroups][kMaxOut][kColsPerBlock]
    for (int e = tid; e < (r1 - r0) * a.out_f; e += kThreads) sdy[e] = masked(a, r0 + e / a.out_f, e % a.out_f);
    __syncthreads();
    float acc[kMaxOut];
#pragma unroll
    for (int o = 0; o < kMaxOut; ++o) acc[o] = 0.0f;
    if (i < a.in_f && a.dw) {
        for (int rb = r0 + rg; rb < r1; rb += kRowGroups * 8) {
            float xv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = rb + kRowGroups * u;
                xv[u] = r < r1 ? __ldg(a.x + (size_t)r * a.in_f + i) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = rb + kRowGroups * u;
                if (r < r1) {
                    const float* g = sdy + (r - r0) * a.out_f;
#pragma unroll
                    for (int o = 0; o < kMaxOut; ++o)
                        if (o < a.out_f) acc[o] = fmaf(g[o], xv[u], acc[o]);
                }
            }
        }
    }
#pragma unroll
    for (int o = 0; o < kMaxOut; ++o) red[(rg * kMaxOut + o) * kColsPerBlock + cg] = acc[o];
    __syncthreads();
    // park the partials: layout [cb][rs][out][kColsPerBlock] (+ a [out] tail per (cb, rs) for db, written by column block 0)
    const size_t per = (size_t)a.out_f * kColsPerBlock + kMaxOut;
    float* mine = a.partial + ((size_t)cb * a.row_splits + rs) * per;
    if (rg == 0) {
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o) {
            if (o < a.out_f) {
                float v = 0.0f;
#pragma unroll
                for (int q = 0; q < kRowGroups; ++q) v += red[(q * kMaxOut + o) * kColsPerBlock + cg];      // row-group order
                mine[o * kColsPerBlock + cg] = v;
            }
        }
    }
    if (cb == 0 && a.db && tid < a.out_f) {
        float s = 0.0f;
        for (int r = r0; r < r1; ++r) s += sdy[(r - r0) * a.out_f + tid];      // rows ascending
        mine[(size_t)a.out_f * kColsPerBlock + tid] = s;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        int prev = atomicAdd(a.tickets + cb, 1);
        s_last = (prev == a.row_splits - 1);
        if (s_last) a.tickets[cb] = 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const float* base = a.partial + (size_t)cb * a.row_splits * per;
    if (rg == 0 && i < a.in_f && a.dw) {
        // split order; the loads of four splits for every output row are in flight before the first add
        float s[kMaxOut];
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o) s[o] = 0.0f;
        for (int z0 = 0; z0 < a.row_splits; z0 += 4) {
            float q[4][kMaxOut];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int o = 0; o < kMaxOut; ++o)
                    q[u][o] = (o < a.out_f && z0 + u < a.row_splits) ? __ldcg(base + (size_t)(z0 + u) * per + o * kColsPerBlock + cg) : 0.0f;
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int o = 0; o < kMaxOut; ++o) s[o] += q[u][o];
        }
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o) {
            if (o < a.out_f) {
                float* d = a.dw + (size_t)o * a.in_f + i;
                *d = a.acc_dw ? *d + s[o] : s[o];
            }
        }
    }
    if (cb == 0 && a.db && tid < a.out_f) {
        float s = 0.0f;
        for (int z0 = 0; z0 < a.row_splits; z0 += 8) {
            float q[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                q[u] = (z0 + u < a.row_splits) ? __ldcg(base + (size_t)(z0 + u) * per + (size_t)a.out_f * kColsPerBlock + tid) : 0.0f;
#pragma unroll
            for (int u = 0; u < 8; ++u) s += q[u];
        }
        a.db[tid] = a.acc_db ? a.db[tid] + s : s;
    }
}

}  // namespace

namespace tp {

bool linear_skinny_ok(int batch, int in_f, int out_f) {
    return out_f <= kMaxOut && (size_t)out_f * in_f * sizeof(float) <= 96 * 1024 && batch > 0;
}

int linear_skinny_fwd(tp_ctx* ctx, const float* x, const float* w, const float* b, float* y, int batch, int in_f, int out_f,
                      int relu) {
    cudaSetDevice(ctx->device);
    if (!ctx->attr_skinny) {
        TP_CUDA(cudaFuncSetAttribute(skinny_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        TP_CUDA(cudaFuncSetAttribute(skinny_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        ctx->attr_skinny = true;
    }
    int blocks = (batch * 32 + kThreads - 1) / kThreads;
    if (blocks > ctx->sm_count) blocks = ctx->sm_count;       // every block stages W once: no more blocks than SMs
    const int vec = (in_f % 4 == 0) && !(((uintptr_t)x | (uintptr_t)w) & 15);
    const size_t smem = (size_t)out_f * in_f * sizeof(float);
    skinny_fwd_kernel<<<blocks, kThreads, smem, ctx->stream>>>(x, w, b, y, batch, in_f, out_f, relu, vec);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

int linear_skinny_bwd(tp_ctx* ctx, const float* x, const float* w, const float* dy, const float* mask_y, float* dx, float* dw,
                      float* db, int batch, int in_f, int out_f, int acc_dx, int acc_dw, int acc_db) {
    cudaSetDevice(ctx->device);
    if (!ctx->attr_skinny) {
        TP_CUDA(cudaFuncSetAttribute(skinny_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        TP_CUDA(cudaFuncSetAttribute(skinny_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        ctx->attr_skinny = true;
    }
    BwdArgs a;
    a.x = x; a.w = w; a.dy = dy; a.mask_y = mask_y; a.dx = dx; a.dw = dw; a.db = db;
    a.batch = batch; a.in_f = in_f; a.out_f = out_f;
    a.acc_dx = acc_dx; a.acc_dw = acc_dw; a.acc_db = acc_db;
    a.dx_blocks = 0;
    if (dx) {
        a.dx_blocks = (batch * 32 + kThreads - 1) / kThreads;
        if (a.dx_blocks > ctx->sm_count) a.dx_blocks = ctx->sm_count;
    }
    a.col_blocks = 0; a.row_splits = 0; a.rows_per_split = 0;
    a.partial = nullptr; a.tickets = nullptr;
    if (dw || db) {
        a.col_blocks = (in_f + kColsPerBlock - 1) / kColsPerBlock;
        // few, long row splits: the last block of a column block folds all of them, so the fold (splits x out loads per
        // thread) must stay short; 8 splits x col_blocks CTAs already cover the dW work in a few microseconds
        int splits = (ctx->sm_count + a.col_blocks - 1) / a.col_blocks;
        if (splits > 8) splits = 8;
        if (splits < 1) splits = 1;
        int rps = (batch + splits - 1) / splits;
        if (rps < 16) rps = 16;
        if (rps > 1024) rps = 1024;                            // dY chunk of a split is staged in shared memory
        a.rows_per_split = rps;
        a.row_splits = (batch + rps - 1) / rps;
        if (a.col_blocks > kCounterGemm - kCounterColsum) {
            set_error("linear_skinny_bwd: in_features %d too large", in_f);
            return TP_ERR_UNSUPPORTED;
        }
        size_t per = (size_t)out_f * kColsPerBlock + kMaxOut;
        int rc = ensure_scratch(ctx, (size_t)a.col_blocks * a.row_splits * per * sizeof(float));
        if (rc) return rc;
        a.partial = ctx->scratch;
        a.tickets = ctx->dev_counters + kCounterColsum;        // same stream, never concurrent with tp_colsum
    }
    int blocks = a.dx_blocks + a.col_blocks * a.row_splits;
    if (blocks == 0) return TP_OK;
    size_t smem_dx = dx ? (size_t)out_f * in_f * sizeof(float) : 0;
    size_t smem_dw = ((size_t)a.rows_per_split * out_f + (size_t)kRowGroups * kMaxOut * kColsPerBlock) * sizeof(float);
    size_t smem = smem_dx > smem_dw ? smem_dx : smem_dw;
    skinny_bwd_kernel<<<blocks, kThreads, smem, ctx->stream>>>(a);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

}  // namespace tp

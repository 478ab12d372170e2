// Exact-fp32 GEMM on the CUDA cores (FFMA), all four transpose combinations, split-K with a
// deterministic fold.  Serves (a) the shapes where a 128-wide tcgen05 tile is pointless
// (N = 10 logits, K = 10 / K = 9 contractions), (b) gemm_mode 0 (bit-faithful fp32 products), and
// (c) operands whose leading dimension is not 16-byte aligned and so cannot be described to TMA.
// Replaces matrixmultiply::sgemm / cblas_sgemm behind sgemm_rowmajor (src/gemm.rs:8-49, 72-119).
#include "common.cuh"

namespace {

struct EpiArgs {
    const float* bias;
    const float* relu_mask;
    int relu;
};

// Packed fp32 FMA (Blackwell FFMA2, PTX fma.rn.f32x2): two IEEE fused multiply-adds per issue slot, bit-identical to fmaf.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void fma2(f32x2& d, f32x2 a, f32x2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }

__device__ __forceinline__ float apply_epilogue(float v, const EpiArgs& ep, size_t idx, int col) {
    if (ep.bias) v += __ldg(ep.bias + col);
    if (ep.relu) v = fmaxf(v, 0.0f);
    if (ep.relu_mask) v = __ldg(ep.relu_mask + idx) > 0.0f ? v : 0.0f;
    return v;
}

// C tile BMxBN per CTA, BK-deep smem stages, each thread owns a TMxTN micro-tile.
// A(i,kk) = A[i*ars + kk*acs], B(kk,j) = B[kk*brs + j*bcs].
template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_simt_kernel(int m, int n, int k, float alpha, const float* __restrict__ A, long ars, long acs,
                 const float* __restrict__ B, long brs, long bcs, float beta, float* __restrict__ C,
                 EpiArgs ep, float* __restrict__ partial, int k_per_split, int* __restrict__ tickets) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int PAD = 4;
    __shared__ __align__(16) float As[BK][BM + PAD];
    __shared__ __align__(16) float Bs[BK][BN + PAD];

    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * k_per_split;
    const int kend = min(k, kbeg + k_per_split);

    constexpr bool kPacked = (TN % 2 == 0);                    // accumulate pairs along n with FFMA2
    constexpr int TN2 = kPacked ? TN / 2 : 1;
    float acc[TM][TN];
    f32x2 acc2[TM][TN2];
#pragma unroll
    for (int a = 0; a < TM; ++a) {
#pragma unroll
        for (int b = 0; b < TN; ++b) acc[a][b] = 0.0f;
#pragma unroll
        for (int b = 0; b < TN2; ++b) acc2[a][b] = 0ull;
    }

    const bool a_k_contig = (acs == 1);
    const bool b_k_contig = (brs == 1);

    for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
        for (int e = tid; e < BM * BK; e += NT) {
            int ii, kk;
            if (a_k_contig) { kk = e % BK; ii = e / BK; } else { ii = e % BM; kk = e / BM; }
            int gi = i0 + ii, gk = k0 + kk;
            As[kk][ii] = (gi < m && gk < kend) ? __ldg(A + gi * ars + gk * acs) : 0.0f;
        }
#pragma unroll
        for (int e = tid; e < BN * BK; e += NT) {
            int jj, kk;
            if (b_k_contig) { kk = e % BK; jj = e / BK; } else { jj = e % BN; kk = e / BN; }
            int gj = j0 + jj, gk = k0 + kk;
            Bs[kk][jj] = (gj < n && gk < kend) ? __ldg(B + gk * brs + gj * bcs) : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float av[TM], bv[TN];
#pragma unroll
            for (int a = 0; a < TM; ++a) av[a] = As[kk][ty * TM + a];
#pragma unroll
            for (int b = 0; b < TN; ++b) bv[b] = Bs[kk][tx * TN + b];
            if constexpr (kPacked) {
                // pairs along n through FFMA2: half the issue slots of the scalar loop, same bits
                f32x2 bp[TN2];
#pragma unroll
                for (int b = 0; b < TN2; ++b) bp[b] = pack2(bv[2 * b], bv[2 * b + 1]);
#pragma unroll
                for (int a = 0; a < TM; ++a) {
                    const f32x2 aa = pack2(av[a], av[a]);
#pragma unroll
                    for (int b = 0; b < TN2; ++b) fma2(acc2[a][b], aa, bp[b]);
                }
            } else {
#pragma unroll
                for (int a = 0; a < TM; ++a)
#pragma unroll
                    for (int b = 0; b < TN; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
            }
        }
        __syncthreads();
    }

    if constexpr (kPacked) {
#pragma unroll
        for (int a = 0; a < TM; ++a)
#pragma unroll
            for (int b = 0; b < TN2; ++b) unpack2(acc2[a][b], acc[a][2 * b], acc[a][2 * b + 1]);
    }
#pragma unroll
    for (int a = 0; a < TM; ++a) {
        int gi = i0 + ty * TM + a;
        if (gi >= m) continue;
#pragma unroll
        for (int b = 0; b < TN; ++b) {
            int gj = j0 + tx * TN + b;
            if (gj >= n) continue;
            size_t idx = (size_t)gi * n + gj;
            if (partial) {
                partial[(size_t)blockIdx.z * m * n + idx] = acc[a][b];
            } else {
                float v = alpha * acc[a][b];
                if (beta != 0.0f) v += beta * C[idx];
                C[idx] = apply_epilogue(v, ep, idx, gj);
            }
        }
    }
    if (partial && tickets) {
        // split-K in one launch: the last CTA of this output tile folds the partial tiles in split order (deterministic)
        __shared__ int s_last;
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            int* tk = tickets + blockIdx.y * gridDim.x + blockIdx.x;
            int prev = atomicAdd(tk, 1);
            s_last = (prev == (int)gridDim.z - 1);
            if (s_last) *tk = 0;
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            const size_t mn = (size_t)m * n;
#pragma unroll
            for (int a = 0; a < TM; ++a) {
                int gi = i0 + ty * TM + a;
                if (gi >= m) continue;
#pragma unroll
                for (int b = 0; b < TN; ++b) {
                    int gj = j0 + tx * TN + b;
                    if (gj >= n) continue;
                    size_t idx = (size_t)gi * n + gj;
                    float sum = 0.0f;                  // split order; 8 loads in flight per output instead of one round trip per split
                    for (int z0 = 0; z0 < (int)gridDim.z; z0 += 8) {
                        float q[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) q[u] = (z0 + u < (int)gridDim.z) ? __ldcg(partial + (size_t)(z0 + u) * mn + idx) : 0.0f;
#pragma unroll
                        for (int u = 0; u < 8; ++u) sum += q[u];
                    }
                    float v = alpha * sum;
                    if (beta != 0.0f) v += beta * C[idx];
                    C[idx] = apply_epilogue(v, ep, idx, gj);
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256)
splitk_fold_kernel(const float* __restrict__ partial, float* __restrict__ C, size_t mn, int n, int splits,
                   float alpha, float beta, EpiArgs ep) {
    const size_t stride = (size_t)gridDim.x * 256;
    for (size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x; idx < mn; idx += stride) {
        float s = 0.0f;
        for (int z0 = 0; z0 < splits; z0 += 8) {
            float q[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) q[u] = (z0 + u < splits) ? partial[(size_t)(z0 + u) * mn + idx] : 0.0f;
#pragma unroll
            for (int u = 0; u < 8; ++u) s += q[u];
        }
        float v = alpha * s;
        if (beta != 0.0f) v += beta * C[idx];
        C[idx] = apply_epilogue(v, ep, idx, (int)(idx % n));
    }
}

template <int BM, int BN, int BK, int TM, int TN>
int launch(tp_ctx* ctx, int m, int n, int k, float alpha, const float* A, long ars, long acs, const float* B, long brs,
           long bcs, float beta, float* C, const EpiArgs& ep) {
    constexpr int NT = (BM / TM) * (BN / TN);
    int gx = (n + BN - 1) / BN, gy = (m + BM - 1) / BM;
    long tiles = (long)gx * gy;
    int splits = 1;
    if (tiles < ctx->sm_count && k >= 64) {
        splits = (int)((2L * ctx->sm_count + tiles - 1) / tiles);
        int max_splits = k / 32;                         // at least two BK steps per CTA
        if (splits > max_splits) splits = max_splits;
        if (splits > 64) splits = 64;
        if (splits < 1) splits = 1;
    }
    int kps = ((k + splits - 1) / splits + BK - 1) / BK * BK;
    if (kps < BK) kps = BK;
    splits = (k + kps - 1) / kps;
    if (splits < 1) splits = 1;
    float* partial = nullptr;
    int* tickets = nullptr;
    if (splits > 1) {
        int rc = tp::ensure_scratch(ctx, (size_t)splits * m * n * sizeof(float));
        if (rc) return rc;
        partial = ctx->scratch;
        if (tiles <= tp::kNumCounters - tp::kCounterGemm) tickets = ctx->dev_counters + tp::kCounterGemm;
    }
    gemm_simt_kernel<BM, BN, BK, TM, TN><<<dim3(gx, gy, splits), NT, 0, ctx->stream>>>(
        m, n, k, alpha, A, ars, acs, B, brs, bcs, beta, C, ep, partial, kps, tickets);
    TP_LAUNCH_OK(ctx);
    if (splits > 1 && !tickets) {
        size_t mn = (size_t)m * n;
        splitk_fold_kernel<<<tp::grid_for(ctx, mn, 256), 256, 0, ctx->stream>>>(partial, C, mn, n, splits, alpha, beta, ep);
        TP_LAUNCH_OK(ctx);
    }
    return TP_OK;
}

}  // namespace

namespace tp {

int gemm_simt(tp_ctx* ctx, int ta, int tb, int m, int n, int k, float alpha, const float* a, const float* b,
              float beta, float* c, const Epilogue& ep) {
    if (m == 0 || n == 0) return TP_OK;
    cudaSetDevice(ctx->device);
    // op(A) is m x k: N -> lda = k (row stride k, col stride 1); T -> stored k x m (row stride 1, col stride m)
    long ars = ta ? 1 : k, acs = ta ? m : 1;
    // op(B) is k x n: N -> ldb = n; T -> stored n x k
    long brs = tb ? 1 : n, bcs = tb ? k : 1;
    EpiArgs e{ep.bias, ep.relu_mask, ep.relu};
    long tiles64 = (long)((m + 63) / 64) * ((n + 63) / 64);
    if (n <= 16)
        return launch<128, 16, 16, 8, 1>(ctx, m, n, k, alpha, a, ars, acs, b, brs, bcs, beta, c, e);
    if (m <= 16)
        return launch<16, 128, 16, 1, 8>(ctx, m, n, k, alpha, a, ars, acs, b, brs, bcs, beta, c, e);
    if (tiles64 >= 2L * ctx->sm_count)
        return launch<128, 128, 8, 8, 8>(ctx, m, n, k, alpha, a, ars, acs, b, brs, bcs, beta, c, e);
    return launch<64, 64, 16, 4, 4>(ctx, m, n, k, alpha, a, ars, acs, b, brs, bcs, beta, c, e);
}

}  // namespace tp

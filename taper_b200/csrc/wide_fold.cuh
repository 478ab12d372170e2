// Bias-gradient fold + step bookkeeping of the wide step plan (step_wide.cu), shared with the bf16x3 GEMM kernel
// (gemm_bx3.cu), whose grouped weight-gradient launch runs it on a few extra CTAs instead of paying a kernel slot for it:
//   * sums the per-CTA partial column sums (bias gradients, src/tensor.rs:680-691) in a fixed order into the gradient arena,
//   * publishes {loss, #correct} (src/loss.rs:157-164, 271-290) to the device result and the host's pinned slot,
//   * advances Adam's t / step size (src/optim.rs:86-90, 157) and the dataset cursor (src/data/mnist.rs:369-384).
#pragma once
#include <cstdint>
#include "../../include/taper_b200.h"

namespace tpfold {

constexpr int kMaxFold = 2 * TP_STEP_MAX_LAYERS + 4;
enum { H_T = 0, H_LR, H_B1, H_B2, H_EPS, H_WD, H_SS, H_DECAY, H_COUNT };       // same layout as optim.cu

struct FoldEntry {
    float* dst;                      // n outputs
    const float* src;                // partial j of output i at src[j * stride + i]
    int n, parts;
    long long stride;
    int first_block;                 // logical blocks [first_block, first_block + ceil(n / 32)) serve this entry
};
// per plan (device memory, written once)
struct FoldTable {
    FoldEntry e[kMaxFold];
    int n_entries, n_blocks;
    const float* lh_part;            // [lh_parts][2]  {sum of NLL, hits} per head CTA
    int lh_parts;
    int B;
    float* result;                   // device {loss, correct}
    float* hyper;                    // Adam state or NULL (SGD)
    const int* err;                  // sticky device error word of the context
};
// per step (kernel argument)
struct FoldStep {
    const FoldTable* table;          // NULL: nothing to fold
    float* result_host;              // optional mapped pinned {loss, correct, seq, err}
    unsigned int result_seq;
    int* cursor;                     // dataset cursor or NULL
    int cursor_delta, cursor_mod;
    unsigned long long* stamp;
};

__device__ __forceinline__ float powi_dev(float a, int b) {     // f32::powi, as optim.cu
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

// Worker `worker` of `n_workers` CTAs of kT threads.  A logical block folds 32 outputs (one per lane); warp w sums partials
// w, w + W, w + 2W, ... with eight loads in flight and the W warp sums are added in warp order: a fixed association order,
// so the result does not depend on timing.  `red` is kT floats of shared memory.
template <int kT>
__device__ __forceinline__ void fold_worker(const FoldStep& st, int worker, int n_workers, float* red) {
    constexpr int W = kT / 32;
    const FoldTable& tb = *st.table;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int b = worker; b < tb.n_blocks; b += n_workers) {
        int ei = 0;
#pragma unroll 1
        for (int i = 1; i < tb.n_entries; ++i)
            if (b >= tb.e[i].first_block) ei = i;
        const FoldEntry e = tb.e[ei];
        const int i = (b - e.first_block) * 32 + lane;
        float s = 0.0f;
        if (i < e.n) {
            for (int j0 = wid; j0 < e.parts; j0 += W * 8) {
                float x[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int j = j0 + u * W;
                    x[u] = j < e.parts ? __ldcg(e.src + (size_t)j * e.stride + i) : 0.0f;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) s += x[u];
            }
        }
        red[wid * 32 + lane] = s;
        __syncthreads();
        if (wid == 0 && i < e.n) {
            float t = red[lane];
#pragma unroll
            for (int w = 1; w < W; ++w) t += red[w * 32 + lane];
            e.dst[i] = t;
        }
        __syncthreads();
    }
    if (worker == 0 && wid == 0) {
        // loss / accuracy: lane l sums head CTAs l, l + 32, ... (ascending), the 32 lane sums are added in lane order
        float n = 0.0f, h = 0.0f;
        for (int j = lane; j < tb.lh_parts; j += 32) { n += __ldcg(tb.lh_part + 2 * j); h += __ldcg(tb.lh_part + 2 * j + 1); }
        float nt = 0.0f, ht = 0.0f;
#pragma unroll
        for (int l = 0; l < 32; ++l) { nt += __shfl_sync(0xffffffffu, n, l); ht += __shfl_sync(0xffffffffu, h, l); }
        if (lane == 0) {
            const float loss = nt / (float)tb.B;                                 // src/loss.rs:164
            tb.result[0] = loss;
            tb.result[1] = ht;
            if (st.result_host) {
                st.result_host[0] = loss;
                st.result_host[1] = ht;
                st.result_host[3] = __int_as_float(__ldcg(tb.err));              // 1: a label outside [0, classes) (the reference panics)
                if (st.result_seq) {
                    __threadfence_system();
                    ((volatile unsigned int*)st.result_host)[2] = st.result_seq;
                }
            }
            if (tb.hyper) {                                                      // Adam::step prologue (src/optim.rs:86-90, 157)
                float* hy = tb.hyper;
                const int tt = __float_as_int(hy[H_T]) + 1;
                hy[H_T] = __int_as_float(tt);
                const float bc1 = 1.0f - powi_dev(hy[H_B1], tt);
                const float bc2 = 1.0f - powi_dev(hy[H_B2], tt);
                hy[H_SS] = hy[H_LR] * (sqrtf(bc2) / bc1);
                hy[H_DECAY] = 1.0f - hy[H_LR] * hy[H_WD];
            }
            if (st.cursor) *st.cursor = (int)(((long long)*st.cursor + st.cursor_delta) % st.cursor_mod);
        }
    }
}

}  // namespace tpfold

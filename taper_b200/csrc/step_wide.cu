// Wide device tape: the whole training step of a WIDE MLP (chain of Linear(+ReLU) ending in a classifier of <= 16 classes)
// as a fixed plan of tensor-core kernels chained with programmatic dependent launch, behind the same tp_step_* ABI as the
// persistent-kernel device tape (tape_step.cu), which keeps the small models (< 1.5 GFLOP per step).
//
// Replaces the loop body of Trainer::train_epoch (src/train.rs:106-138) for models like BASELINE configs[3]
// (784-1024-1024-10, batch 1024): Linear::forward (src/nn.rs:54-60: transpose, matmul src/ops.rs:200-228, add_broadcast
// src/tensor.rs:636-704), ReLU (src/ops.rs:312-374), cross_entropy_loss + accuracy (src/loss.rs:136-195, 271-290), the backward
// closures (dY.W, dY^T.X src/ops.rs:254-291, bias column sums src/tensor.rs:680-691, ReLU mask src/ops.rs:358-370) and the
// optimizer step (src/optim.rs:21-33, 83-113, 148-168).
//
// Plan for L layers (kernels per step = 2L + 4):
//   input   gather the batch rows out of the resident dataset (or take the host-fed batch; fp32 or u8 pixels / 255,
//           src/data/mnist.rs:225, 276-309) and write them as bf16 hi/lo planes (the GEMM operand format, gemm_bx3.cu)
//   fwd l   act_l = relu(in_l . W_l^T + b_l)        bf16x3 tcgen05 GEMM; the epilogue also writes act_l's hi/lo planes
//   head    logits = act . W_last^T + b (exact fp32), log-softmax, NLL, first-max accuracy, dlogits = (softmax - onehot) / B,
//           dZ = (dlogits . W_last) * [act > 0] and dlogits as hi/lo planes, per-CTA partials of db_last and colsum(dZ)
//   dW_last = dlogits^T . act                          (T,N, a 16-row A operand: the classifier's weight gradient)
//   bwd l   dZ_{l-1} = (dZ_l . W_l) * [act_{l-1} > 0]  (N,N; ReLU mask, hi/lo planes and column-sum partials in the epilogue)
//           dW_l = dZ_l^T . in_l                       (T,N; straight into the gradient arena)
//   fold    sums the bias-gradient partials in a fixed order into the gradient arena, publishes {loss, correct},
//           advances Adam's t / step size and the dataset cursor
//   [allreduce of the gradient arena when data-parallel]
//   opt     SGD / Adam / AdamW over the flat arena; also rewrites the parameters' hi/lo planes for the next step
// No kernel of the plan reads an fp32 operand through the tensor pipe and nothing is transposed or copied in memory.
#include "common.cuh"
#include "tc_ptx.cuh"
#include "wide_fold.cuh"

#include <cuda_bf16.h>
#include <cstdlib>
#include <vector>

namespace {

using namespace tcptx;
using namespace tpfold;

constexpr int kThreads = 256;
constexpr int kMaxOut = 16;
constexpr int kHeadRowsMax = 16;     // rows of the batch per head CTA and pass (staged in shared memory)

__device__ __forceinline__ unsigned int class_of(float t) {     // `t as usize` (src/loss.rs:160)
    if (!(t > 0.0f)) return 0u;
    if (t >= 4294967040.0f) return 0xffffffffu;
    return (unsigned int)t;
}

__device__ __forceinline__ void split2(float v, uint16_t& hi, uint16_t& lo) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi = __bfloat16_as_ushort(h);
    lo = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(h)));
}

__device__ __forceinline__ void split_store4(uint16_t* hi, uint16_t* lo, float a, float b, float c, float d) {
    uint16_t h[4], l[4];
    split2(a, h[0], l[0]); split2(b, h[1], l[1]); split2(c, h[2], l[2]); split2(d, h[3], l[3]);
    *(uint2*)hi = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
    *(uint2*)lo = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
}

// profiling: the latest time any CTA of the grid passes this point (the slot is zeroed by the host before the run)
__device__ __forceinline__ void stamp_max(unsigned long long* stamp) {
    if (stamp && threadIdx.x == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        atomicMax(stamp, gt);
    }
}

__device__ __forceinline__ void stamp_now(unsigned long long* stamp) {
    if (stamp && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        *stamp = gt;
    }
}

// ---- input: batch rows -> bf16 hi/lo planes (+ labels) --------------------------------------------------------------
struct InputArgs {
    const void* x;                   // [rows or n_perm, cols] fp32 or u8
    const float* labels;
    const int* perm;                 // NULL: x / labels are the batch itself
    const int* cursor;
    int cursor_value;                // >= 0: host mirror of *cursor
    int n_perm, rows, cols, is_u8;
    uint16_t* hi;
    uint16_t* lo;
    float* y;                        // [rows]
    unsigned long long* stamp;
};

__global__ void __launch_bounds__(kThreads)
wide_input_kernel(InputArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    stamp_now(a.stamp);
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * kThreads + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * kThreads) >> 5;
    const int start = a.perm ? (a.cursor_value >= 0 ? a.cursor_value : __ldg(a.cursor)) : 0;
    for (int r = warp; r < a.rows; r += nwarps) {
        const int src = a.perm ? __ldg(a.perm + (start + r) % a.n_perm) : r;
        uint16_t* hi = a.hi + (size_t)r * a.cols;
        uint16_t* lo = a.lo + (size_t)r * a.cols;
        if (a.is_u8) {
            // MNIST pixels as stored on disk: f32 = u8 / 255 (src/data/mnist.rs:225); 8 pixels per lane and load, four loads
            // in flight per lane (a 784-pixel row is one round trip to HBM, not four)
            const uint2* s = (const uint2*)((const unsigned char*)a.x + (size_t)src * a.cols);
            const int n8 = a.cols / 8;
            for (int c0 = 0; c0 < n8; c0 += 128) {
                uint2 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int c = c0 + u * 32 + lane;
                    v[u] = c < n8 ? __ldg(s + c) : make_uint2(0u, 0u);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int c = c0 + u * 32 + lane;
                    if (c >= n8) continue;
                    const uint32_t w[2] = {v[u].x, v[u].y};
#pragma unroll
                    for (int hlf = 0; hlf < 2; ++hlf) {
                        const float f0 = (float)(w[hlf] & 0xffu) / 255.0f, f1 = (float)((w[hlf] >> 8) & 0xffu) / 255.0f;
                        const float f2 = (float)((w[hlf] >> 16) & 0xffu) / 255.0f, f3 = (float)(w[hlf] >> 24) / 255.0f;
                        split_store4(hi + 8 * c + 4 * hlf, lo + 8 * c + 4 * hlf, f0, f1, f2, f3);
                    }
                }
            }
        } else {
            const float4* s = (const float4*)((const float*)a.x + (size_t)src * a.cols);
            const int n4 = a.cols / 4;
            for (int c0 = 0; c0 < n4; c0 += 256) {
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int c = c0 + u * 32 + lane;
                    v[u] = c < n4 ? __ldg(s + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int c = c0 + u * 32 + lane;
                    if (c < n4) split_store4(hi + 4 * c, lo + 4 * c, v[u].x, v[u].y, v[u].z, v[u].w);
                }
            }
        }
        if (lane == 0) a.y[r] = __ldg(a.labels + src);
    }
}

// ---- head -----------------------------------------------------------------------------------------------------------------
struct HeadArgs {
    const float* act;                // [B, K] fp32 input of the classifier
    const float* w;                  // [C, K]
    const float* bias;               // [C] or NULL
    const float* y;                  // [B] labels
    int B, K, C, relu_mask;          // relu_mask: dZ *= [act > 0]
    int rows_per_cta;
    float inv_b;
    uint16_t* dz_hi;                 // [B, K] planes of dZ
    uint16_t* dz_lo;
    uint16_t* dl_hi;                 // [B, 16] planes of dlogits (zero padded): A operand of the dW_last GEMM
    uint16_t* dl_lo;
    float* db_part;                  // [grid][kMaxOut]
    float* cs_part;                  // [grid][K]   column sums of dZ
    float* lh_part;                  // [grid][2]   {sum of NLL, hits}
    int* err;
    unsigned long long* stamp;
};

// One CTA owns rows_per_cta rows of the batch.  Phase A (a warp per row): logits, log-softmax, loss / accuracy, dlogits.
// Phase B (a thread per 4 columns): dZ rows and their column sums, with W_last's columns in registers.  Everything a row
// needs is staged once in shared memory.  CT = compile-time class count (weights of classes >= C are zero padding).
template <int CT>
__global__ void __launch_bounds__(kThreads)
wide_head_kernel(HeadArgs a) {
    extern __shared__ __align__(16) float sm[];
    float* w_s = sm;                                  // [CT][K]
    float* a_s = w_s + CT * a.K;                      // [kHeadRowsMax][K]
    float* dl_s = a_s + kHeadRowsMax * a.K;           // [kHeadRowsMax][kMaxOut]
    float* red_s = dl_s + kHeadRowsMax * kMaxOut;     // [8][2]
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int K4 = a.K / 4;
    pdl_launch_dependents();
    // the classifier's parameters were written by the previous step's optimizer kernel, which completed before the kernel
    // ahead of this one could start: safe to stage before the dependency wait
    for (int i0 = t; i0 < CT * K4; i0 += kThreads * 5) {                   // five loads in flight per thread
        float4 v[5];
#pragma unroll
        for (int u = 0; u < 5; ++u) {
            const int i = i0 + u * kThreads;
            v[u] = (i < CT * K4 && i < a.C * K4) ? __ldg((const float4*)a.w + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 5; ++u)
            if (i0 + u * kThreads < CT * K4) ((float4*)w_s)[i0 + u * kThreads] = v[u];
    }
    float bias_r[CT];
#pragma unroll
    for (int c = 0; c < CT; ++c) bias_r[c] = (a.bias && c < a.C) ? __ldg(a.bias + c) : 0.0f;
    pdl_wait();
    stamp_now(a.stamp);
    const int r0 = blockIdx.x * a.rows_per_cta;
    const int r1 = min(a.B, r0 + a.rows_per_cta);
    float4 wreg[CT];
    float4 csacc = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool colthread = t < K4;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < CT; ++c) wreg[c] = colthread ? ((const float4*)(w_s + c * a.K))[t] : make_float4(0.f, 0.f, 0.f, 0.f);
    float nll_sum = 0.0f, hit_sum = 0.0f;             // lane 0 of each warp
    float dbacc = 0.0f;                               // thread c < C: db_last[c]
    for (int rb = r0; rb < r1; rb += kHeadRowsMax) {
        const int nr = min(kHeadRowsMax, r1 - rb);
        // ---- phase A ----
        for (int rr = wid; rr < nr; rr += kThreads / 32) {
            const int row = rb + rr;
            const float4* xr = (const float4*)(a.act + (size_t)row * a.K);
            const float tgt = __ldg(a.y + row);
            float logit[CT];
#pragma unroll
            for (int c = 0; c < CT; ++c) logit[c] = 0.0f;
            float4 xv[8];                                                      // K <= 1024: the whole row in one round trip
#pragma unroll
            for (int u = 0; u < 8; ++u) xv[u] = u * 32 + lane < K4 ? __ldg(xr + u * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = u * 32 + lane;
                if (j < K4) {
                    const float4 x = xv[u];
                    ((float4*)(a_s + rr * a.K))[j] = x;
#pragma unroll
                    for (int c = 0; c < CT; ++c) {
                        const float4 w = ((const float4*)(w_s + c * a.K))[j];
                        logit[c] = fmaf(x.x, w.x, logit[c]); logit[c] = fmaf(x.y, w.y, logit[c]);
                        logit[c] = fmaf(x.z, w.z, logit[c]); logit[c] = fmaf(x.w, w.w, logit[c]);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < CT; ++c) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) logit[c] += __shfl_xor_sync(0xffffffffu, logit[c], o);
                logit[c] += bias_r[c];
            }
            // every lane holds all logits: log-softmax (src/loss.rs:101-126), NLL (:158-164), first-max accuracy (:271-290)
            float m = -INFINITY;
            int bi = 0;
#pragma unroll
            for (int c = 0; c < CT; ++c)
                if (c < a.C && logit[c] > m) { m = logit[c]; bi = c; }
            float s = 0.0f;
#pragma unroll
            for (int c = 0; c < CT; ++c)
                if (c < a.C) s += expf(logit[c] - m);
            const float ls = logf(s);
            unsigned int cls = class_of(tgt);
            if (cls >= (unsigned int)a.C) { if (lane == 0) atomicExch(a.err, 1); cls = a.C - 1; }
            if (lane == 0) {
                float lp_t = 0.0f;
#pragma unroll
                for (int c = 0; c < CT; ++c)
                    if (c == (int)cls) lp_t = (logit[c] - m) - ls;
                nll_sum += -lp_t;
                if (fabsf((float)bi - tgt) < 1e-6f) hit_sum += 1.0f;             // src/loss.rs:284
            }
            // dlogits = (exp(logp) - onehot) * (1 / B)  (src/loss.rs:174-191, upstream gradient 1)
            if (lane < kMaxOut) {
                float d = 0.0f;
#pragma unroll
                for (int c = 0; c < CT; ++c)
                    if (c == lane && c < a.C) d = (expf((logit[c] - m) - ls) - (c == (int)cls ? 1.0f : 0.0f)) * a.inv_b;
                dl_s[rr * kMaxOut + lane] = d;
                uint16_t h, l;
                split2(d, h, l);
                a.dl_hi[(size_t)row * kMaxOut + lane] = h;
                a.dl_lo[(size_t)row * kMaxOut + lane] = l;
            }
        }
        __syncthreads();
        // ---- phase B ----
        if (colthread) {
            for (int rr = 0; rr < nr; ++rr) {
                const float4 x = ((const float4*)(a_s + rr * a.K))[t];
                const float4 d0 = ((const float4*)(dl_s + rr * kMaxOut))[0], d1 = ((const float4*)(dl_s + rr * kMaxOut))[1];
                const float4 d2 = ((const float4*)(dl_s + rr * kMaxOut))[2], d3 = ((const float4*)(dl_s + rr * kMaxOut))[3];
                const float dv[16] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w, d2.x, d2.y, d2.z, d2.w, d3.x, d3.y, d3.z, d3.w};
                float4 dz = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int c = 0; c < CT; ++c) {
                    dz.x = fmaf(dv[c], wreg[c].x, dz.x); dz.y = fmaf(dv[c], wreg[c].y, dz.y);
                    dz.z = fmaf(dv[c], wreg[c].z, dz.z); dz.w = fmaf(dv[c], wreg[c].w, dz.w);
                }
                if (a.relu_mask) {                                             // src/ops.rs:358-370
                    dz.x = x.x > 0.0f ? dz.x : 0.0f; dz.y = x.y > 0.0f ? dz.y : 0.0f;
                    dz.z = x.z > 0.0f ? dz.z : 0.0f; dz.w = x.w > 0.0f ? dz.w : 0.0f;
                }
                csacc.x += dz.x; csacc.y += dz.y; csacc.z += dz.z; csacc.w += dz.w;
                const size_t g = (size_t)(rb + rr) * a.K + 4 * t;
                split_store4(a.dz_hi + g, a.dz_lo + g, dz.x, dz.y, dz.z, dz.w);
            }
        }
        if (t < a.C)
            for (int rr = 0; rr < nr; ++rr) dbacc += dl_s[rr * kMaxOut + t];
        __syncthreads();
    }
    // ---- this CTA's partials ----
    if (colthread) ((float4*)(a.cs_part + (size_t)blockIdx.x * a.K))[t] = csacc;
    if (t < kMaxOut) a.db_part[blockIdx.x * kMaxOut + t] = t < a.C ? dbacc : 0.0f;
    if (lane == 0) { red_s[2 * wid] = nll_sum; red_s[2 * wid + 1] = hit_sum; }
    __syncthreads();
    if (t == 0) {
        float n = 0.0f, h = 0.0f;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) { n += red_s[2 * w]; h += red_s[2 * w + 1]; }
        a.lh_part[2 * blockIdx.x] = n;
        a.lh_part[2 * blockIdx.x + 1] = h;
    }
}

// ---- fold: partials -> gradient arena, results, optimizer counters (wide_fold.cuh) ------------------------------------
// Runs on a few extra CTAs of the grouped weight-gradient launch (gemm_bx3.cu); this stand-alone kernel serves plans whose
// weight-gradient GEMMs cannot take it along.
__global__ void __launch_bounds__(kThreads)
wide_fold_kernel(const FoldStep st) {
    __shared__ float red[kThreads];
    pdl_launch_dependents();
    pdl_wait();
    stamp_now(st.stamp);
    fold_worker<kThreads>(st, (int)blockIdx.x, (int)gridDim.x, red);
}

// ---- data-parallel: two-phase gradient exchange over NVLink peer memory + optimizer, one kernel -------------------------
// Every rank owns a window (tp_xchg_*): flags + slot rows, mapped into every peer.  The arena is cut into `world` slices and
// every slice into gridDim.x chunks; CTA c serves chunk c of every slice and only ever waits for CTA c of the other ranks
// (no grid-wide wait, no co-residency assumption):
//   A  push my gradient chunk of slice j into rank j's window (row = my rank), then raise flag1[me][c] there   (reduce-scatter)
//   B  wait for flag1[q][c] of every q, sum the world's contributions to MY slice in rank order, push the reduced chunk
//      into every rank's result row, raise flag2[me][c] there                                                  (all-gather)
//   C  wait for flag2[q][c] of every q, apply SGD / Adam / AdamW to chunk c of every slice and rewrite the bf16 planes
// Per rank 2 * (world-1)/world * arena bytes cross NVLink (one-shot pushing of whole arenas would be (world-1) * arena);
// each slice is summed by exactly one rank in rank order, so replicas stay bit-identical.  Double-buffered by step parity.
constexpr int kMaxWorld = 8;
constexpr long long kPeerSpinLimit = 40000000000LL;            // ~20 s: a peer that never arrives must not hang this GPU for ever
struct XchgArgs {
    float* P; const float* G; float* M; float* V; const float* hyper;
    uint16_t* hi; uint16_t* lo;
    long long arena4, slice4, chunk4;    // float4 counts: arena, per slice, per (slice, chunk)
    long long row;                       // floats between the rows of a slot array
    int opt_kind;
    float sgd_lr, grad_scale;
    int world, rank, items;
    unsigned int seq;
    const unsigned int* my_flags;        // [world][items]: flag1 at [q][c], flag2 at [q][gridDim.x + c]
    float* my_slots;                     // this parity: [world][row]; row q = {recv from q: slice4*4 floats, result slice q: slice4*4}
    unsigned int* peer_flags[kMaxWorld];
    float* peer_slots[kMaxWorld];
    int* err;
    unsigned long long* stamp;
    unsigned long long* stamp_x;     // profiling: 16 more slots for the kernel's own phases (CTA 0's stamps, latest-CTA stamps)
    // float4 ranges of the arena the weight-gradient GEMMs' epilogues have already pushed (Bx3Push): phase A skips them
    int n_pushed;
    long long pushed_lo[4], pushed_hi[4];
};

struct AdamArgsW {
    float step_size, beta1, beta2, eps, weight_decay, grad_scale, decay_factor;
    int decoupled;
};
// identical to optim.cu's adam_elem (expression order follows src/optim.rs:93-110; -fmad=false)
__device__ __forceinline__ void adam_elem_w(float& p, float g, float& m, float& v, const AdamArgsW& a) {
    if (a.decoupled) p *= a.decay_factor;
    if (a.grad_scale != 1.0f) g *= a.grad_scale;
    float gg = g + a.weight_decay * p;
    m = a.beta1 * m + (1.0f - a.beta1) * gg;
    v = a.beta2 * v + (1.0f - a.beta2) * gg * gg;
    p -= a.step_size * m / (sqrtf(v) + a.eps);
}

__device__ __forceinline__ bool wait_flags(const XchgArgs& a, int flag_index) {
    const int t = threadIdx.x;
    if (t < a.world && t != a.rank) {
        const unsigned int* f = a.my_flags + (size_t)t * a.items + flag_index;
        unsigned int cur;
        const long long t0 = clock64();
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(cur) : "l"(f) : "memory");
            if ((int)(cur - a.seq) < 0 && clock64() - t0 > kPeerSpinLimit) { atomicExch(a.err, 3); break; }
        } while ((int)(cur - a.seq) < 0);
    }
    __syncthreads();
    return __ldcg(a.err) < 2;
}

__device__ __forceinline__ void raise_flags(const XchgArgs& a, int flag_index) {
    __syncthreads();                     // the CTA's data stores are ordered before the release below
    const int t = threadIdx.x;
    if (t < a.world && t != a.rank)
        asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(a.peer_flags[t] + (size_t)a.rank * a.items + flag_index), "r"(a.seq) : "memory");
}

__global__ void __launch_bounds__(kThreads)
wide_xchg_opt_kernel(const __grid_constant__ XchgArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    stamp_now(a.stamp);
    const int t = threadIdx.x, c = blockIdx.x, C_ = gridDim.x;
    const long long c0 = (long long)c * a.chunk4;
    const float4* G4 = reinterpret_cast<const float4*>(a.G);
    auto chunk_len = [&](int q) {        // float4 of chunk c that exist in slice q
        const long long s_len = min(a.slice4, a.arena4 - (long long)q * a.slice4);
        long long n = min(a.chunk4, s_len - c0);
        return n > 0 ? n : 0;
    };
    // ---- A: my contributions to the other ranks' slices ----
    // One flat index space over (slice j, vector i of chunk c): a thread keeps four independent load -> remote store pairs in
    // flight whatever the world size (one loop per peer was a chain of world - 1 dependent L2 round trips per CTA: 5.4 us at 8 GPUs)
    const int ch = (int)a.chunk4;
    const int flat = a.world * ch;
    __shared__ int s_len[kMaxWorld];                               // vectors of chunk c that exist in slice q
    if (t < kMaxWorld) s_len[t] = t < a.world ? (int)chunk_len(t) : 0;
    __syncthreads();
    {
        bool all_pushed = a.n_pushed > 0;                          // every vector of this CTA's chunks already pushed by the GEMMs?
        for (int j = 0; j < a.world && all_pushed; ++j) {
            if (j == a.rank) continue;
            const long long n = chunk_len(j), e0 = (long long)j * a.slice4 + c0;
            bool covered = n == 0;
            for (int r = 0; r < a.n_pushed; ++r) covered = covered || (e0 >= a.pushed_lo[r] && e0 + n <= a.pushed_hi[r]);
            all_pushed = covered;
        }
        if (!all_pushed) {
            const int flat_a = flat - ch;                          // the peers' slices only
            for (int x0 = t; x0 < flat_a; x0 += kThreads * 4) {
                float4 v[4];
                float4* dst[4];
                bool live[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int x = x0 + u * kThreads;
                    const int jp = x / ch, i = x - jp * ch;
                    const int j = min(jp + (jp >= a.rank ? 1 : 0), a.world - 1);
                    live[u] = x < flat_a && i < s_len[j];
                    const long long e = (long long)j * a.slice4 + c0 + i;         // arena position (float4)
                    for (int r = 0; r < a.n_pushed; ++r)
                        if (e >= a.pushed_lo[r] && e < a.pushed_hi[r]) live[u] = false;
                    if (live[u]) {
                        v[u] = __ldcg(G4 + e);
                        dst[u] = reinterpret_cast<float4*>(a.peer_slots[j] + (long long)a.rank * a.row) + c0 + i;
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) if (live[u]) *dst[u] = v[u];
            }
        }
    }
    stamp_now(a.stamp_x ? a.stamp_x + 1 : nullptr);          // CTA 0's phase stamps (profiling only): pushed | flag raised | peers seen | ...
    raise_flags(a, c);
    stamp_now(a.stamp_x ? a.stamp_x + 2 : nullptr);
    // ---- B: reduce my slice in rank order, hand the result to everybody ----
    if (!wait_flags(a, c)) return;       // a peer never arrived: its slot is stale, apply nothing
    stamp_now(a.stamp_x ? a.stamp_x + 3 : nullptr);
    stamp_max(a.stamp_x ? a.stamp_x + 7 : nullptr);
    {
        const long long n = chunk_len(a.rank);
        const long long res_off = a.slice4;                    // result area of a row starts after its recv area
        for (long long i0 = t; i0 < n; i0 += kThreads * 2) {                 // two output vectors x world loads in flight per thread
            float4 q[2][kMaxWorld];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const long long i = i0 + u * kThreads;
#pragma unroll
                for (int r = 0; r < kMaxWorld; ++r)
                    if (r < a.world && i < n)
                        q[u][r] = (r == a.rank) ? __ldcg(G4 + (long long)a.rank * a.slice4 + c0 + i)
                                                : __ldcg(reinterpret_cast<const float4*>(a.my_slots + (long long)r * a.row) + c0 + i);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const long long i = i0 + u * kThreads;
                if (i >= n) continue;
                float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int r = 0; r < kMaxWorld; ++r)                                // rank order on every rank: one rank sums a slice
                    if (r < a.world) { sum.x += q[u][r].x; sum.y += q[u][r].y; sum.z += q[u][r].z; sum.w += q[u][r].w; }
#pragma unroll
                for (int r = 0; r < kMaxWorld; ++r)
                    if (r < a.world) {
                        float* base = (r == a.rank) ? a.my_slots : a.peer_slots[r];
                        reinterpret_cast<float4*>(base + (long long)a.rank * a.row)[res_off + c0 + i] = sum;
                    }
            }
        }
    }
    stamp_now(a.stamp_x ? a.stamp_x + 4 : nullptr);
    raise_flags(a, C_ + c);
    // ---- C: optimizer on chunk c of every slice ----
    if (!wait_flags(a, C_ + c)) return;
    stamp_now(a.stamp_x ? a.stamp_x + 5 : nullptr);
    stamp_max(a.stamp_x ? a.stamp_x + 8 : nullptr);
    AdamArgsW aa{};
    if (a.opt_kind != 0) {
        const float* h = a.hyper;
        const int decoupled = a.opt_kind == 2;
        aa.step_size = h[H_SS]; aa.beta1 = h[H_B1]; aa.beta2 = h[H_B2]; aa.eps = h[H_EPS];
        aa.weight_decay = decoupled ? 0.0f : h[H_WD];
        aa.grad_scale = a.grad_scale;
        aa.decay_factor = h[H_DECAY];
        aa.decoupled = (decoupled && h[H_WD] > 0.0f) ? 1 : 0;
    }
    float4* P4 = reinterpret_cast<float4*>(a.P);
    float4* M4 = reinterpret_cast<float4*>(a.M);
    float4* V4 = reinterpret_cast<float4*>(a.V);
    // flat index space over (slice q, vector i of chunk c) again: four (gradient, p, m, v) quadruples in flight per thread
    for (int x0 = t; x0 < flat; x0 += kThreads * 4) {
        float4 gg[4], pp[4], mm[4], vv[4];
        long long ee[4];
        bool live[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int x = x0 + u * kThreads;
            const int q = x / ch, i = x - q * ch;
            live[u] = x < flat && i < s_len[q];
            ee[u] = (long long)q * a.slice4 + c0 + i;
            if (live[u]) {
                gg[u] = __ldcg(reinterpret_cast<const float4*>(a.my_slots + (long long)q * a.row) + a.slice4 + c0 + i);
                pp[u] = P4[ee[u]];
                if (a.opt_kind != 0) { mm[u] = M4[ee[u]]; vv[u] = V4[ee[u]]; }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (!live[u]) continue;
            if (a.opt_kind == 0) {
                if (a.grad_scale != 1.0f) { gg[u].x *= a.grad_scale; gg[u].y *= a.grad_scale; gg[u].z *= a.grad_scale; gg[u].w *= a.grad_scale; }
                pp[u].x -= a.sgd_lr * gg[u].x; pp[u].y -= a.sgd_lr * gg[u].y;                                      // src/optim.rs:29
                pp[u].z -= a.sgd_lr * gg[u].z; pp[u].w -= a.sgd_lr * gg[u].w;
            } else {
                adam_elem_w(pp[u].x, gg[u].x, mm[u].x, vv[u].x, aa);
                adam_elem_w(pp[u].y, gg[u].y, mm[u].y, vv[u].y, aa);
                adam_elem_w(pp[u].z, gg[u].z, mm[u].z, vv[u].z, aa);
                adam_elem_w(pp[u].w, gg[u].w, mm[u].w, vv[u].w, aa);
                M4[ee[u]] = mm[u]; V4[ee[u]] = vv[u];
            }
            P4[ee[u]] = pp[u];
            split_store4(a.hi + 4 * ee[u], a.lo + 4 * ee[u], pp[u].x, pp[u].y, pp[u].z, pp[u].w);
        }
    }
    stamp_now(a.stamp_x ? a.stamp_x + 6 : nullptr);
    stamp_max(a.stamp_x ? a.stamp_x + 9 : nullptr);
}

template <typename Kern, typename Arg>
int launch_pdl(tp_ctx* ctx, Kern kern, dim3 grid, size_t smem, const Arg& arg, bool pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    TP_CUDA(cudaLaunchKernelEx(&cfg, kern, arg));
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

}  // namespace

namespace tp {

// what the plan needs to know about a connected tp_xchg window (filled by tape_step.cu, where the type lives)
struct WideXchg {
    int rank = 0, world = 1, items = 0;
    size_t arena_len = 0, row = 0, slots_off = 0;
    unsigned char* window = nullptr;
    unsigned char* peers[8] = {};
    unsigned int seq = 0;            // flag value of this run
};

struct WidePlan {
    tp_ctx* ctx = nullptr;
    tp_step_desc d{};
    float *P = nullptr, *G = nullptr, *M = nullptr, *V = nullptr, *hyper = nullptr, *result = nullptr;
    unsigned char* block = nullptr;  // one device allocation
    size_t bytes = 0;
    long long arena_plane = 0;
    uint16_t* w_split = nullptr;     // planes of the whole parameter arena
    uint16_t* x_split = nullptr;     // planes of the input batch
    float* y = nullptr;
    std::vector<float*> act;         // fp32 activations of the hidden layers
    std::vector<uint16_t*> act_split, dz_split;
    std::vector<float*> cs_part;     // column-sum partials of dz_split[l] (bias gradient of hidden layer l)
    std::vector<int> cs_parts;
    uint16_t* dl_split = nullptr;    // planes of dlogits [B, 16]
    float *db_part = nullptr, *lh_part = nullptr;
    Bx3Launch dw_last{};             // dW_last = dlogits^T . act   (T,N) on the tensor cores
    int head_grid = 0, head_rows = 0;
    size_t head_smem = 0;
    std::vector<Bx3Launch> fwd, dx, dw;   // dx[l] produces dz[l-1] (unused for l = 0)
    FoldTable fold{};                // host copy; the device copy lives in the plan's block
    FoldTable* fold_dev = nullptr;
    bool fold_in_group = true;       // the grouped dW launch runs the fold on extra CTAs
    bool push_in_gemm = false;       // data parallel: the dW epilogues push other ranks' slices into their exchange windows
                                     // (TAPER_WIDE_PUSH_IN_GEMM=1; measured neutral at 2 GPUs: the NVLink stores then lengthen the GEMM)
    bool pdl = true;
    bool weights_fresh = false;
    unsigned long long* stamps = nullptr;    // [2][16] %globaltimer per kernel of the last two steps (tp_step_set_profile)
    bool profile = false;
    unsigned int runs = 0;
    int kernels_per_step = 0;        // launches of the last run (the fold rides on the dW launch when it can)
};

namespace {
struct Carver {
    unsigned char* base = nullptr;
    size_t off = 0;
    template <typename T> T* take(size_t count) {
        off = (off + 1023) & ~(size_t)1023;
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

double wide_flops(const tp_step_desc* d) {
    double f = 0.0;
    for (int l = 0; l < d->n_layers; ++l) f += 2.0 * d->batch * (double)d->dims[l] * d->dims[l + 1];
    return 3.0 * f;
}
}  // namespace

bool wide_supported(const tp_step_desc* d, const char** why) {
    auto no = [&](const char* w) { if (why) *why = w; return false; };
    if (!d) return no("NULL descriptor");
    const int L = d->n_layers;
    if (L < 2 || L > TP_STEP_MAX_LAYERS) return no("layer count (the wide plan needs at least one hidden layer)");
    if (d->batch < 1 || d->batch > 65536) return no("batch size");
    if (d->optimizer < 0 || d->optimizer > 2) return no("optimizer kind");
    if (d->dims[L] < 1 || d->dims[L] > kMaxOut) return no("classifier wider than 16");
    if (d->dims[L - 1] > 1024) return no("classifier input wider than 1024");
    for (int l = 0; l < L; ++l) {
        if (d->dims[l] < 8 || d->dims[l] % 8) return no("feature width not a multiple of 8");
        if (d->w_off[l] % 4 || (d->b_off[l] >= 0 && d->b_off[l] % 4)) return no("parameter slice not 16-byte aligned");
        if (l + 1 < L && d->w_off[l] % 8) return no("hidden weight slice not 16-byte aligned as bf16");
    }
    if (d->arena_len % 4) return no("arena length not a multiple of 4");
    return true;
}

// which implementation tp_step_create picks: the plan for wide models, the persistent kernel for small ones
bool wide_preferred(const tp_step_desc* d) {
    static const int force = [] { const char* v = getenv("TAPER_STEP_WIDE"); return v && *v ? atoi(v) : -1; }();
    if (!wide_supported(d, nullptr)) return false;
    if (force == 0) return false;
    if (force == 1) return true;
    return wide_flops(d) > 1.5e9;
}

static int wide_build(WidePlan* w, Carver& c) {
    const tp_step_desc& d = w->d;
    const int L = d.n_layers, B = d.batch;
    const int C = d.dims[L], KH = d.dims[L - 1];
    w->arena_plane = (d.arena_len + 7) & ~(long long)7;
    w->w_split = c.take<uint16_t>(2 * (size_t)w->arena_plane);
    w->x_split = c.take<uint16_t>(2 * (size_t)B * d.dims[0]);
    w->y = c.take<float>(B);
    w->act.assign(L - 1, nullptr); w->act_split.assign(L - 1, nullptr); w->dz_split.assign(L - 1, nullptr);
    w->cs_part.assign(L - 1, nullptr); w->cs_parts.assign(L - 1, 0);
    for (int l = 0; l + 1 < L; ++l) {
        const size_t n = (size_t)B * d.dims[l + 1];
        w->act[l] = c.take<float>(n);
        w->act_split[l] = c.take<uint16_t>(2 * n);
        w->dz_split[l] = c.take<uint16_t>(2 * n);
    }
    // head: rows per CTA so that the grid covers the SMs once (at most kHeadRowsMax rows are staged per pass)
    int rows = (B + w->ctx->sm_count - 1) / w->ctx->sm_count;
    if (rows < 8) rows = B >= 8 * 32 ? 8 : (rows < 1 ? 1 : rows);
    w->head_rows = rows;
    w->head_grid = (B + rows - 1) / rows;
    w->head_smem = ((size_t)(C == 10 ? 10 : kMaxOut) * KH + (size_t)kHeadRowsMax * KH + kHeadRowsMax * kMaxOut + 16) * sizeof(float);
    w->dl_split = c.take<uint16_t>(2 * (size_t)B * kMaxOut);
    w->db_part = c.take<float>((size_t)w->head_grid * kMaxOut);
    w->cs_part[L - 2] = c.take<float>((size_t)w->head_grid * KH);
    w->cs_parts[L - 2] = w->head_grid;
    w->lh_part = c.take<float>((size_t)w->head_grid * 2);
    // column-sum partials of the dX GEMMs (sized for the worst tiling: 128-row tiles x 8 K-splits)
    for (int l = 0; l + 2 < L; ++l) w->cs_part[l] = c.take<float>((size_t)((B + 127) / 128) * 8 * d.dims[l + 1]);
    w->fold_dev = c.take<FoldTable>(1);
    if (!c.base) return TP_OK;

    tp_ctx* ctx = w->ctx;
    w->fwd.assign(L - 1, Bx3Launch{}); w->dx.assign(L - 1, Bx3Launch{}); w->dw.assign(L - 1, Bx3Launch{});
    // the weight-gradient GEMMs (every layer's dW and the classifier's) are independent and share launches: one tile
    // width and one K-split for all of them, chosen for the total tile count
    const int dw_bn = 128;
    long dw_tiles = (KH + dw_bn - 1) / dw_bn;          // dW_last: one row of tiles
    for (int l = 0; l + 1 < L; ++l) dw_tiles += (long)((d.dims[l + 1] + 127) / 128) * ((d.dims[l] + dw_bn - 1) / dw_bn);
    const int dw_splits = bx3_best_splits(ctx, dw_bn, dw_tiles, B);
    for (int l = 0; l + 1 < L; ++l) {
        const int in = d.dims[l], out = d.dims[l + 1];
        const uint16_t* in_split = l == 0 ? w->x_split : w->act_split[l - 1];
        const long long in_plane = (long long)B * in;
        // act_l = relu(in . W_l^T + b_l)   (N,T)
        Bx3Epilogue ef;
        ef.bias = d.b_off[l] >= 0 ? w->P + d.b_off[l] : nullptr;
        ef.relu = d.relu[l];
        ef.c_split = w->act_split[l];
        int rc = bx3_prepare(ctx, 0, 1, B, out, in, 1.0f, in_split, in_plane, w->w_split + d.w_off[l], w->arena_plane, 0.0f, w->act[l], ef,
                             &w->fwd[l]);
        if (rc) return rc;
        w->fwd[l].b_early = true;                      // W_l's planes: written by the previous step's optimizer kernel
        // dW_l = dZ_l^T . in   (T,N) straight into the gradient arena
        Bx3Epilogue ew;
        rc = bx3_prepare(ctx, 1, 0, out, in, B, 1.0f, w->dz_split[l], (long long)B * out, in_split, in_plane, 0.0f, w->G + d.w_off[l], ew,
                         &w->dw[l], dw_bn, dw_splits);
        if (rc) return rc;
        // the weight-gradient GEMMs run after the whole dX chain (one grouped launch): the forward activations are old by
        // then, and so is every dZ except the one the last dX kernel has just produced (dZ_0)
        w->dw[l].b_early = true;
        w->dw[l].a_early = l > 0;
        if (l > 0) {
            // dZ_{l-1} = (dZ_l . W_l) * [act_{l-1} > 0]   (N,N); planes + column sums (db_{l-1}) in the epilogue
            Bx3Epilogue ex;
            ex.relu_mask = d.relu[l - 1] ? w->act[l - 1] : nullptr;
            ex.c_split = w->dz_split[l - 1];
            ex.colsum_part = d.b_off[l - 1] >= 0 ? w->cs_part[l - 1] : nullptr;
            rc = bx3_prepare(ctx, 0, 0, B, in, out, 1.0f, w->dz_split[l], (long long)B * out, w->w_split + d.w_off[l], w->arena_plane, 0.0f,
                             nullptr, ex, &w->dx[l]);
            if (rc) return rc;
            w->dx[l].b_early = true;
            w->cs_parts[l - 1] = w->dx[l].tiles_m * w->dx[l].splits;
        }
    }
    {
        // dW_last[C, KH] = dlogits^T . act  (T,N): the A operand is declared 16 wide (zero padded planes); rows >= C are
        // computed and not stored
        Bx3Epilogue el;
        int rc = bx3_prepare(ctx, 1, 0, kMaxOut, KH, B, 1.0f, w->dl_split, (long long)B * kMaxOut, w->act_split[L - 2], (long long)B * KH, 0.0f,
                             w->G + d.w_off[L - 1], el, &w->dw_last, dw_bn, dw_splits);
        if (rc) return rc;
        w->dw_last.m = C;
        w->dw_last.b_early = true;
        w->dw_last.a_early = L > 2;                    // dlogits come from the head; a dX kernel sits in between unless L == 2
    }
    // fold table (device copy in the plan's block)
    FoldTable& f = w->fold;
    f = FoldTable{};
    int nb = 0;
    auto add = [&](float* dst, const float* src, int n, int parts, long long stride) {
        FoldEntry& e = f.e[f.n_entries++];
        e.dst = dst; e.src = src; e.n = n; e.parts = parts; e.stride = stride; e.first_block = nb;
        nb += (n + 31) / 32;
    };
    if (d.b_off[L - 1] >= 0) add(w->G + d.b_off[L - 1], w->db_part, C, w->head_grid, kMaxOut);
    for (int l = 0; l + 1 < L; ++l)
        if (d.b_off[l] >= 0) add(w->G + d.b_off[l], w->cs_part[l], d.dims[l + 1], w->cs_parts[l], d.dims[l + 1]);
    f.n_blocks = nb;
    f.lh_part = w->lh_part; f.lh_parts = w->head_grid; f.B = B;
    f.result = w->result;
    f.err = ctx->dev_error;
    f.hyper = d.optimizer != 0 ? w->hyper : nullptr;
    if (cudaMemcpyAsync(w->fold_dev, &f, sizeof f, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        cudaGetLastError();
        set_error("tp_step_create: uploading the fold table failed");
        return TP_ERR_CUDA;
    }
    { const char* v = getenv("TAPER_WIDE_FOLD_IN_GROUP"); w->fold_in_group = !(v && v[0] == '0'); }
    { const char* v = getenv("TAPER_WIDE_PUSH_IN_GEMM"); w->push_in_gemm = v && v[0] == '1'; }
    return TP_OK;
}

int wide_create(tp_ctx* ctx, const tp_step_desc* desc, float* P, float* G, float* M, float* V, float* hyper, float* result,
                WidePlan** out) {
    const char* why = nullptr;
    if (!wide_supported(desc, &why)) {
        set_error("tp_step_create: unsupported wide step (%s)", why ? why : "?");
        return TP_ERR_UNSUPPORTED;
    }
    cudaSetDevice(ctx->device);
    WidePlan* w = new WidePlan();
    w->ctx = ctx; w->d = *desc;
    w->P = P; w->G = G; w->M = M; w->V = V; w->hyper = hyper; w->result = result;
    { const char* v = getenv("TAPER_WIDE_PDL"); w->pdl = !(v && v[0] == '0'); }
    Carver sizing;
    wide_build(w, sizing);
    w->bytes = sizing.off + 1024;
    if (cudaMalloc(&w->block, w->bytes) != cudaSuccess) {
        cudaGetLastError();
        set_error("tp_step_create: cudaMalloc(%zu) failed", w->bytes);
        delete w;
        return TP_ERR_OOM;
    }
    cudaMemsetAsync(w->block, 0, w->bytes, ctx->stream);
    Carver place;
    place.base = w->block;
    int rc = wide_build(w, place);
    if (!rc) {
        int optin = 0;
        if (cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device) != cudaSuccess || (size_t)optin < w->head_smem ||
            cudaFuncSetAttribute(wide_head_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin) != cudaSuccess ||
            cudaFuncSetAttribute(wide_head_kernel<kMaxOut>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin) != cudaSuccess) {
            cudaGetLastError();
            set_error("tp_step_create: %zu bytes of shared memory not available for the head kernel", w->head_smem);
            rc = TP_ERR_CUDA;
        }
    }
    if (rc) {
        cudaFree(w->block);
        delete w;
        return rc;
    }
    *out = w;
    return TP_OK;
}

int wide_set_profile(WidePlan* w, int on) {
    cudaSetDevice(w->ctx->device);
    if (on && !w->stamps) {
        TP_CUDA(cudaMalloc(&w->stamps, 64 * sizeof(unsigned long long)));
        TP_CUDA(cudaMemsetAsync(w->stamps, 0, 64 * sizeof(unsigned long long), w->ctx->stream));
    }
    w->profile = on != 0;
    return TP_OK;
}

// out[0..15] = stamps of the step before last, out[16..31] = of the last one (slots: input, fwd..., head, bwd..., fold, optimizer)
int wide_read_profile(WidePlan* w, long long* out, size_t cap, int* slots) {
    if (!w->stamps || cap < 32) { set_error("tp_step_read_profile: profiling is off or the buffer is too small"); return TP_ERR_INVALID; }
    cudaSetDevice(w->ctx->device);
    TP_CUDA(cudaStreamSynchronize(w->ctx->stream));
    unsigned long long h[64];                          // per parity: 16 kernel slots, then 16 slots of the exchange kernel's phases
    TP_CUDA(cudaMemcpy(h, w->stamps, sizeof h, cudaMemcpyDeviceToHost));
    const int last = (w->runs + 1) & 1;                // parity of the last run
    for (int i = 0; i < 16; ++i) { out[i] = (long long)h[(last ^ 1) * 32 + i]; out[16 + i] = (long long)h[last * 32 + i]; }
    if (cap >= 64)                                     // out[32..47] / out[48..63]: the exchange kernel's phase stamps of the same two runs
        for (int i = 0; i < 16; ++i) { out[32 + i] = (long long)h[(last ^ 1) * 32 + 16 + i]; out[48 + i] = (long long)h[last * 32 + 16 + i]; }
    if (slots) *slots = 16;
    return TP_OK;
}

void wide_destroy(WidePlan* w) {
    if (!w) return;
    cudaSetDevice(w->ctx->device);
    cudaStreamSynchronize(w->ctx->stream);
    if (w->stamps) cudaFree(w->stamps);
    if (w->block) cudaFree(w->block);
    delete w;
}

// the parameters were changed by somebody else (host upload, broadcast, another step): re-split them
int wide_refresh(WidePlan* w) {
    w->weights_fresh = false;
    return TP_OK;
}

void wide_info(const WidePlan* w, int* n_phases, int* n_jobs, int* grid) {
    const int L = w->d.n_layers;
    if (n_phases) *n_phases = w->kernels_per_step ? w->kernels_per_step : 2 * L + 1;
    if (n_jobs) *n_jobs = 3 * (L - 1) + 5;
    if (grid) *grid = w->ctx->sm_count;
}

int wide_run(WidePlan* w, const void* x, int x_is_u8, const float* labels, const int* perm, int* cursor, int n_perm, int cursor_value,
             float sgd_lr, float grad_scale, float* result_host, unsigned int result_seq, const WideXchg* xc) {
    tp_ctx* ctx = w->ctx;
    const tp_step_desc& d = w->d;
    const int L = d.n_layers, B = d.batch;
    cudaSetDevice(ctx->device);
    const bool pdl = w->pdl;
    int rc = TP_OK;
    if (!w->weights_fresh) {
        rc = split_bf16(ctx, w->P, w->w_split, (size_t)d.arena_len, w->arena_plane, false);
        if (rc) return rc;
        w->weights_fresh = true;
    }
    const uint64_t launches0 = ctx->launches;
    unsigned long long* st = w->profile ? w->stamps + (w->runs & 1) * 32 : nullptr;
    int slot = 0;
    auto next_stamp = [&]() { return st ? st + (slot < 15 ? slot++ : 15) : nullptr; };
    InputArgs ia{};
    ia.stamp = next_stamp();
    ia.x = x; ia.labels = labels; ia.perm = perm; ia.cursor = cursor; ia.cursor_value = perm ? cursor_value : -1;
    ia.n_perm = n_perm; ia.rows = B; ia.cols = d.dims[0]; ia.is_u8 = x_is_u8;
    ia.hi = w->x_split; ia.lo = w->x_split + (size_t)B * d.dims[0]; ia.y = w->y;
    rc = launch_pdl(ctx, wide_input_kernel, dim3(grid_for(ctx, (size_t)B * 32, kThreads, 2)), 0, ia, pdl);
    if (rc) return rc;
    for (int l = 0; l + 1 < L; ++l) {
        w->fwd[l].stamp = next_stamp();
        rc = bx3_launch(ctx, w->fwd[l], pdl);
        if (rc) return rc;
    }
    HeadArgs ha{};
    ha.stamp = next_stamp();
    ha.act = w->act[L - 2]; ha.w = w->P + d.w_off[L - 1]; ha.bias = d.b_off[L - 1] >= 0 ? w->P + d.b_off[L - 1] : nullptr;
    ha.y = w->y; ha.B = B; ha.K = d.dims[L - 1]; ha.C = d.dims[L]; ha.relu_mask = d.relu[L - 2];
    ha.rows_per_cta = w->head_rows; ha.inv_b = 1.0f / (float)B;
    ha.dz_hi = w->dz_split[L - 2]; ha.dz_lo = w->dz_split[L - 2] + (size_t)B * d.dims[L - 1];
    ha.dl_hi = w->dl_split; ha.dl_lo = w->dl_split + (size_t)B * kMaxOut;
    ha.db_part = w->db_part; ha.cs_part = w->cs_part[L - 2]; ha.lh_part = w->lh_part;
    ha.err = ctx->dev_error;
    rc = d.dims[L] == 10 ? launch_pdl(ctx, wide_head_kernel<10>, dim3(w->head_grid), w->head_smem, ha, pdl)
                         : launch_pdl(ctx, wide_head_kernel<kMaxOut>, dim3(w->head_grid), w->head_smem, ha, pdl);
    if (rc) return rc;
    FoldStep fs{};
    fs.table = w->fold_dev;
    fs.result_host = result_host;
    fs.result_seq = result_host ? result_seq : 0u;
    fs.cursor = perm ? cursor : nullptr;
    fs.cursor_delta = B; fs.cursor_mod = n_perm > 0 ? n_perm : 1;
    Bx3Push push{};
    if (xc && w->push_in_gemm) {
        // data parallel: the weight-gradient epilogues store every vector that belongs to another rank's slice straight into that
        // rank's exchange window — the reduce-scatter's NVLink traffic rides on the GEMMs' stores
        const size_t par = xc->seq & 1u;
        const long long slice4 = (d.arena_len / 4 + xc->world - 1) / xc->world;
        push.base = w->G;
        push.slice = (unsigned int)(slice4 * 4);
        push.world = xc->world; push.rank = xc->rank;
        for (int r = 0; r < xc->world; ++r)
            push.peer[r] = r == xc->rank ? nullptr
                                         : reinterpret_cast<float*>(xc->peers[r] + xc->slots_off) + par * xc->world * xc->row + (size_t)xc->rank * xc->row;
    }
    const Bx3Push* pp = push.world > 1 ? &push : nullptr;
    bool folded = false;
    if (pp && L >= 3) {
        // ... and the weight gradients are INTERLEAVED with the dX chain instead of following it: dW_l only needs dZ_l, so it is
        // launched as soon as dZ_l exists, and its pushes cross NVLink while the next dX / dW kernels compute (at 8 GPUs the
        // pushes of all gradients at the end of the backward pass cost 25-35 us of exposed NVLink time).  One more launch than the
        // grouped form.  A dW kernel now directly follows the producer of its dZ operand: no early request of that operand.
        Bx3Launch first_last = w->dw_last, first_dw = w->dw[L - 2];
        first_last.a_early = false; first_dw.a_early = false;
        first_last.stamp = next_stamp(); first_dw.stamp = nullptr;
        const Bx3Launch* g1[2] = {&first_last, &first_dw};
        rc = bx3_launch_group(ctx, g1, 2, pdl, nullptr, nullptr, pp);
        if (rc) return rc;
        for (int l = L - 2; l > 0; --l) {
            w->dx[l].stamp = next_stamp();
            rc = bx3_launch(ctx, w->dx[l], pdl);                   // dZ_{l-1}
            if (rc) return rc;
            Bx3Launch g = w->dw[l - 1];
            g.a_early = false;
            g.stamp = next_stamp();
            const Bx3Launch* one[1] = {&g};
            const bool last = l == 1;
            rc = bx3_launch_group(ctx, one, 1, pdl, last && w->fold_in_group ? &fs : nullptr, last ? &folded : nullptr, pp);
            if (rc) return rc;
        }
    } else {
        for (int l = L - 2; l > 0; --l) {                  // the dX chain: dZ_{L-2} -> ... -> dZ_0
            w->dx[l].stamp = next_stamp();
            rc = bx3_launch(ctx, w->dx[l], pdl);
            if (rc) return rc;
        }
        // every weight gradient: independent of each other, so compatible ones share a launch — and a few extra CTAs of that
        // launch fold the bias-gradient partials and publish the step's results (everything they read is complete by then)
        const Bx3Launch* all[TP_STEP_MAX_LAYERS + 1];
        int n = 0;
        w->dw_last.stamp = next_stamp();
        all[n++] = &w->dw_last;
        for (int l = L - 2; l >= 0; --l) { w->dw[l].stamp = nullptr; all[n++] = &w->dw[l]; }
        rc = bx3_launch_group(ctx, all, n, pdl, w->fold_in_group ? &fs : nullptr, &folded, pp);
        if (rc) return rc;
    }
    if (!folded) {
        fs.stamp = next_stamp();
        rc = launch_pdl(ctx, wide_fold_kernel, dim3(w->fold.n_blocks < 2 * ctx->sm_count ? w->fold.n_blocks : 2 * ctx->sm_count), 0, fs, pdl);
        if (rc) return rc;
    }
    if (xc) {
        // allreduce + optimizer as one kernel over NVLink peer memory (reduce-scatter, all-gather, update)
        const size_t par = xc->seq & 1u;
        XchgArgs xa{};
        xa.P = w->P; xa.G = w->G; xa.M = w->M; xa.V = w->V; xa.hyper = w->hyper;
        xa.hi = w->w_split; xa.lo = w->w_split + w->arena_plane;
        xa.arena4 = d.arena_len / 4;
        xa.slice4 = (xa.arena4 + xc->world - 1) / xc->world;
        // one wave: every CTA spins on its peers' flags, so all of them must be resident at once (two per SM by registers);
        // a second wave would only start after the first has been through both exchange phases
        static const int per_sm = [] { const char* v = getenv("TAPER_XCHG_CTAS_PER_SM"); return v && *v ? atoi(v) : 2; }();
        int grid = per_sm * ctx->sm_count < xc->items / 2 ? per_sm * ctx->sm_count : xc->items / 2;
        if ((long long)grid > xa.slice4) grid = (int)xa.slice4;
        xa.chunk4 = (xa.slice4 + grid - 1) / grid;
        xa.row = (long long)xc->row;
        xa.opt_kind = d.optimizer; xa.sgd_lr = sgd_lr; xa.grad_scale = grad_scale;
        xa.world = xc->world; xa.rank = xc->rank; xa.items = xc->items; xa.seq = xc->seq;
        xa.my_flags = reinterpret_cast<const unsigned int*>(xc->window);
        xa.my_slots = reinterpret_cast<float*>(xc->window + xc->slots_off) + par * xc->world * xc->row;
        for (int r = 0; r < xc->world; ++r) {
            xa.peer_flags[r] = reinterpret_cast<unsigned int*>(xc->peers[r]);
            xa.peer_slots[r] = reinterpret_cast<float*>(xc->peers[r] + xc->slots_off) + par * xc->world * xc->row;
        }
        xa.err = ctx->dev_error;
        xa.stamp = next_stamp();
        xa.stamp_x = st ? st + 16 : nullptr;
        xa.n_pushed = 0;
        if (w->push_in_gemm) {
            auto pushed = [&](const Bx3Launch& g) {
                if (!g.c || xa.n_pushed >= 4) return;
                const long long lo = (g.c - w->G) / 4;
                xa.pushed_lo[xa.n_pushed] = lo;
                xa.pushed_hi[xa.n_pushed] = lo + (long long)g.m * g.n / 4;
                xa.n_pushed++;
            };
            pushed(w->dw_last);
            for (int l = L - 2; l >= 0; --l) pushed(w->dw[l]);
        }
        w->runs++;
        rc = launch_pdl(ctx, wide_xchg_opt_kernel, dim3(grid), 0, xa, pdl);
        w->kernels_per_step = (int)(ctx->launches - launches0);
        return rc;
    }
    if (d.data_parallel && ctx->world > 1 && ctx->nccl_comm) {
        // sum of the per-rank mean gradients; the optimizer folds 1 / world (grad_scale)
        tp_buf view;
        view.ctx = ctx; view.ptr = w->G; view.n = (size_t)d.arena_len; view.external = true;
        rc = tp_allreduce_sum(ctx, &view, (size_t)d.arena_len);
        if (rc) return rc;
    }
    unsigned long long* opt_stamp = next_stamp();
    w->runs++;
    rc = optimizer_step_split(ctx, d.optimizer, w->P, w->G, w->M, w->V, w->hyper, sgd_lr, grad_scale, (size_t)d.arena_len, w->w_split,
                              w->w_split + w->arena_plane, pdl, opt_stamp);
    w->kernels_per_step = (int)(ctx->launches - launches0);
    return rc;
}

}  // namespace tp

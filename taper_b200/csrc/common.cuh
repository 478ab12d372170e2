// Internal definitions shared by every translation unit of libtaper_b200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include "../../include/taper_b200.h"

namespace tp {

constexpr int kNumCounters = 1024;
constexpr int kCounterXent = 0;      // fused softmax-xent fold
constexpr int kCounterColsum = 16;   // [16, 528): one per 32-column block of tp_colsum
constexpr int kCounterGemm = 528;    // [528, 1024): one per output tile of the split-K CUDA-core GEMM

void set_error(const char* fmt, ...);

struct Allocator {
    // Size-bucketed caching allocator.  Single stream per context, so a block freed by the host
    // and re-used by a later launch is ordered after every earlier use (no events needed).
    // A context has one main pool plus, while a CUDA graph is being captured, a private pool that
    // the finished graph then owns: memory a captured step uses as scratch is never handed to
    // anyone else while the graph can still be replayed.
    std::unordered_map<size_t, std::vector<void*>> free_lists;
    std::unordered_set<void*> all_blocks;
    size_t in_use = 0, reserved = 0;
    size_t live_bufs = 0;            // tp_bufs currently allocated from this pool
    bool orphaned = false;           // owner (graph) is gone; delete when live_bufs reaches 0
    static size_t bucket(size_t bytes);
    void* alloc(size_t bytes, size_t* cap, Allocator* steal_from = nullptr, bool relaxed_capture = false);
    void free(void* p, size_t cap);
    void release_all();
};

}  // namespace tp

struct tp_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;      // input prefetch: H2D copies that overlap the compute stream (created on first use)
    tp::Allocator alloc;
    tp::Allocator* capture_pool = nullptr;   // non-null between tp_graph_begin and tp_graph_end
    uint64_t launches = 0;
    int gemm_mode = 1;               // 0 exact fp32 SIMT, 1 3xTF32 tcgen05, 2 1xTF32 tcgen05, 3 bf16x3 tcgen05
    int* dev_error = nullptr;        // sticky device-side error flag
    int* dev_counters = nullptr;     // kNumCounters zero-initialised "last block done" tickets (each user resets its own)
    void* pinned = nullptr;          // staging ring for pageable uploads
    size_t pinned_bytes = 0;
    cudaEvent_t pinned_ev = nullptr;
    bool capturing = false;
    // scratch for split-K / two-stage reductions (grown on demand, never during capture)
    float* scratch = nullptr;
    size_t scratch_bytes = 0;
    std::vector<float*> retired_scratch;
    // communication
    void* nccl_comm = nullptr;
    int rank = 0, world = 1;
    void* tc_state = nullptr;        // tensor-map cache of the tcgen05 GEMM path
    void* bx3_state = nullptr;       // cluster residency table of the bf16x3 GEMM path
    // per-context "function attribute already set" flags (cudaFuncSetAttribute is per device, contexts may sit on different ones)
    bool attr_skinny = false;
    bool attr_conv[3] = {false, false, false};
};

struct tp_buf {
    tp_ctx* ctx = nullptr;
    float* ptr = nullptr;
    size_t n = 0;
    size_t cap = 0;                  // allocator bucket bytes (0 for views / external)
    tp::Allocator* pool = nullptr;   // pool the block came from
    std::atomic<int> rc{1};
    tp_buf* parent = nullptr;        // for slices
    bool external = false;
};

#define TP_CHECK_ARG(cond, ...)                         \
    do {                                                \
        if (!(cond)) {                                  \
            tp::set_error(__VA_ARGS__);                 \
            return TP_ERR_INVALID;                      \
        }                                               \
    } while (0)

#define TP_CUDA(expr)                                                                   \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            tp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),       \
                          __FILE__, __LINE__);                                          \
            return TP_ERR_CUDA;                                                         \
        }                                                                               \
    } while (0)

// call after every kernel launch
#define TP_LAUNCH_OK(ctx)                                                               \
    do {                                                                                \
        (ctx)->launches++;                                                              \
        cudaError_t _e = cudaPeekAtLastError();                                         \
        if (_e != cudaSuccess) {                                                        \
            cudaGetLastError();                                                         \
            tp::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),   \
                          __FILE__, __LINE__);                                          \
            return TP_ERR_CUDA;                                                         \
        }                                                                               \
    } while (0)

#define TP_NEED(buf, count, name)                                                        \
    TP_CHECK_ARG((buf) != nullptr && (buf)->n >= (size_t)(count),                        \
                 "%s: buffer '%s' is NULL or shorter than %zu elements", __func__, name, (size_t)(count))

namespace tp {

// Grid sizing for HBM-bound grid-stride kernels: a multiple of the SM count, capped by the work.
inline int grid_for(const tp_ctx* ctx, size_t work_items, int threads, int ctas_per_sm = 8) {
    size_t need = (work_items + threads - 1) / threads;
    size_t cap = (size_t)ctx->sm_count * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

int ensure_scratch(tp_ctx* ctx, size_t bytes);

// internal GEMM entry points (device pointers), see gemm.cu
struct Epilogue {
    const float* bias = nullptr;     // per output column n
    int relu = 0;                    // max(x, 0) after bias
    const float* relu_mask = nullptr;// multiply by [mask[m,n] > 0] (same layout as C)
    float* colsum = nullptr;         // unused by the SIMT path
};
int gemm_rowmajor(tp_ctx* ctx, int ta, int tb, int m, int n, int k, float alpha, const float* a,
                  const float* b, float beta, float* c, const Epilogue& ep);
int gemm_simt(tp_ctx* ctx, int ta, int tb, int m, int n, int k, float alpha, const float* a,
              const float* b, float beta, float* c, const Epilogue& ep);
// returns TP_ERR_UNSUPPORTED when the shape/alignment cannot go through TMA + tcgen05
int gemm_tc(tp_ctx* ctx, int ta, int tb, int m, int n, int k, float alpha, const float* a,
            const float* b, float beta, float* c, const Epilogue& ep, int mode);
void gemm_tc_destroy(tp_ctx* ctx);
// Implicit-GEMM convolution forward on the tcgen05 path: y[n, cout, ho, wo] = (im2col(x)[M, K] * w2[K, cout]) + bias (+ ReLU)
// without ever materialising the im2col matrix (the A tiles are gathered from the NCHW input straight into swizzled shared
// memory) or the NHWC product (the epilogue writes NCHW).
// Returns TP_ERR_UNSUPPORTED when the shape cannot go this way (caller falls back to im2col + GEMM).
struct ConvShape {
    int n, c, h, w, cout, kh, kw, sh, sw, ph, pw, dh, dw, ho, wo, K;
};
int gemm_tc_conv_fwd(tp_ctx* ctx, const float* x, const float* w2, const float* bias, int relu, float* y, const ConvShape& g);


// ---- stack of 3x3 / s1 / p1 convolutions (+bias, optional ReLU, optional 2x2 max-pool each) on the tcgen05 bf16x3 path over NHWC
// bf16 hi/lo planes (conv_bx3.cu); NCHW fp32 in and out.  TP_ERR_UNSUPPORTED (nothing launched) when a shape does not fit.
int conv_stack_fwd(tp_ctx* ctx, const float* x, int N, int C0, int H, int W, int n_layers, const float* const* w2,
                   const float* const* bias, const int* cout, const int* pool, const int* relu, float* y,
                   float* gap_mean = nullptr, float* gap_cnt = nullptr);       // gap_mean: global average pool on top (y unused)

// input gradient of a 3x3 / s1 / p1 convolution on the same kernel (dZ = gy * [relu_mask_y > 0] re-laid as planes, weights transposed
// and mirrored); NCHW fp32 in and out.  TP_ERR_UNSUPPORTED when the shape does not fit (nothing launched).
int conv_bx3_dx(tp_ctx* ctx, const float* gy, const float* relu_mask_y, const float* w2, float* dx, int N, int Cin, int H, int W,
                int Cout, int accumulate);

// weight gradient of a 3x3 / s1 / p1 convolution as an implicit GEMM over pixels (both operands MN-major planes, two taps per MMA)
int conv_bx3_dw(tp_ctx* ctx, const float* x, const float* gy, const float* relu_mask_y, float* dw, int N, int Cin, int H, int W, int Cout,
                int accumulate);

// ---- bf16x3 tensor-core GEMM on pre-split operands (gemm_bx3.cu) -----------------------------------------------------
// A "split" tensor holds an fp32 tensor of n elements as two bf16 planes: hi = rn_bf16(x) at [0, n) and
// lo = rn_bf16(x - hi) `plane` elements further (plane >= n, multiple of 8).
struct Bx3Epilogue {
    const float* bias = nullptr;     // per output column
    int relu = 0;
    const float* relu_mask = nullptr;// fp32 [m,n]: out = mask > 0 ? out : 0
    float* colsum_part = nullptr;    // [tiles_m * splits][n] partial column sums of the stored values
    uint16_t* c_split = nullptr;     // bf16 planes of the output (operand of the next GEMM)
    long long c_plane = 0;           // 0: m * n
};
struct Bx3Launch {                   // a prepared launch: tensor maps encoded once, replayed every step
    CUtensorMap ma, mb;
    int m, n, k, bn, splits, tiles_m, tiles_n;
    bool a_mn, b_mn, drain;
    float alpha, beta;
    float* c;
    uint16_t* c_split;
    long long c_plane;
    const float* bias;
    const float* relu_mask;
    int relu;
    float* colsum_part;
    unsigned long long* stamp;       // profiling: where the kernel stores %globaltimer after its dependency wait (NULL: off)
    bool a_early, b_early;           // under PDL: the operand was complete before the PREVIOUS kernel started (weights, older
                                     // activations), so its first pipeline stages may be requested before the dependency wait
};
int bx3_prepare(tp_ctx* ctx, int ta, int tb, int m, int n, int k, float alpha, const uint16_t* a_split, long long a_plane,
                const uint16_t* b_split, long long b_plane, float beta, float* c, const Bx3Epilogue& ep, Bx3Launch* out,
                int want_bn = 0, int want_splits = 0);                 // 0: tile width / K-split from the cost model
int bx3_best_splits(tp_ctx* ctx, int bn, long tiles, int k);           // K-split for tiles that share one launch
int bx3_launch(tp_ctx* ctx, const Bx3Launch& L, bool pdl);
// compatible problems share a launch; `tail` (a tpfold::FoldStep, wide_fold.cuh) rides along on extra CTAs of the LAST launch
// when that launch leaves SMs free (*tail_done says whether it did)
// data parallel: the epilogue also stores every output vector that lies in another rank's slice of the gradient arena into
// that rank's exchange window over NVLink (the reduce-scatter push of step_wide.cu's exchange rides on the GEMM's stores)
struct Bx3Push {
    const float* base = nullptr;     // the local gradient arena; outputs outside [base, base + world * slice) are not pushed
    float* peer[8] = {};             // peer[q]: where rank q receives MY contribution to its slice (NULL for q == rank)
    unsigned int slice = 0;          // floats per slice (a multiple of 4)
    int world = 0, rank = 0;         // world <= 1: off
};
int bx3_launch_group(tp_ctx* ctx, const Bx3Launch* const* Ls, int count, bool pdl, const void* tail = nullptr, bool* tail_done = nullptr,
                     const Bx3Push* push = nullptr);
int split_bf16(tp_ctx* ctx, const float* src, uint16_t* dst, size_t n, long long plane, bool pdl);
// fp32 operands: splits both into temporaries first.  TP_ERR_UNSUPPORTED when the shape cannot go through TMA.
int gemm_bx3(tp_ctx* ctx, int ta, int tb, int m, int n, int k, float alpha, const float* a, const float* b, float beta,
             float* c, const Epilogue& ep);
void gemm_bx3_destroy(tp_ctx* ctx);
// optimizer step over a flat arena that also rewrites the parameters' bf16 hi/lo planes (optim.cu); kind 0 SGD, 1 Adam, 2 AdamW
int optimizer_step_split(tp_ctx* ctx, int kind, float* p, const float* g, float* m, float* v, const float* hyper, float sgd_lr,
                         float grad_scale, size_t n, uint16_t* hi, uint16_t* lo, bool pdl, unsigned long long* stamp = nullptr);

// Linear layers with out_features <= 16 (classifier heads), see linear_skinny.cu
bool linear_skinny_ok(int batch, int in_f, int out_f);
int linear_skinny_fwd(tp_ctx* ctx, const float* x, const float* w, const float* b, float* y, int batch, int in_f, int out_f, int relu);
int linear_skinny_bwd(tp_ctx* ctx, const float* x, const float* w, const float* dy, const float* mask_y, float* dx, float* dw,
                      float* db, int batch, int in_f, int out_f, int acc_dx, int acc_dw, int acc_db);

}  // namespace tp

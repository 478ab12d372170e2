// Device-resident dataset and batch gather.  Reference: MNISTDataset::get_batch (src/data/mnist.rs:276-309,
// a rayon row gather of 784-float images) driven by DataLoader's shuffled index list (:326-385).
// HBM-bound: each gathered row is read once and written once (8 B per element + 8 B per label).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

// one warp per row; 128-bit loads when the row is 16-byte aligned (cols % 4 == 0)
__global__ void __launch_bounds__(kThreads)
gather_batch_kernel(const float* __restrict__ images, const float* __restrict__ labels, const int* __restrict__ perm,
                    const int* __restrict__ cursor, float* __restrict__ dst_x, float* __restrict__ dst_y,
                    int rows, int cols, int n_perm, int vec) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * kThreads + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * kThreads) >> 5;
    const int start = __ldg(cursor);
    for (int r = warp; r < rows; r += nwarps) {
        const int src = __ldg(perm + (start + r) % n_perm);
        const float* s = images + (size_t)src * cols;
        float* d = dst_x + (size_t)r * cols;
        if (vec) {
            const float4* s4 = (const float4*)s;
            float4* d4 = (float4*)d;
            for (int c = lane; c < cols / 4; c += 32) d4[c] = __ldg(s4 + c);
        } else {
            for (int c = lane; c < cols; c += 32) d[c] = __ldg(s + c);
        }
        if (lane == 0 && labels) dst_y[r] = __ldg(labels + src);
    }
}

__global__ void cursor_advance_kernel(int* cursor, int delta, int modulo) {
    *cursor = (int)(((long long)*cursor + delta) % modulo);
}

}  // namespace

extern "C" {

int tp_gather_batch(tp_ctx* ctx, const tp_buf* images, const tp_buf* labels, const tp_buf* perm_i32, const tp_buf* cursor_i32,
                    tp_buf* dst_x, tp_buf* dst_y, int rows, int cols, int n_perm) {
    TP_CHECK_ARG(ctx && rows >= 0 && cols > 0 && n_perm > 0, "tp_gather_batch: bad dims");
    TP_NEED(images, (size_t)cols, "images"); TP_NEED(perm_i32, n_perm, "perm"); TP_NEED(cursor_i32, 1, "cursor");
    TP_NEED(dst_x, (size_t)rows * cols, "dst_x");
    if (labels) TP_NEED(dst_y, rows, "dst_y");
    if (!rows) return TP_OK;
    int vec = cols % 4 == 0 && !(((uintptr_t)images->ptr | (uintptr_t)dst_x->ptr) & 15);
    gather_batch_kernel<<<tp::grid_for(ctx, (size_t)rows * 32, kThreads), kThreads, 0, ctx->stream>>>(
        images->ptr, labels ? labels->ptr : nullptr, (const int*)perm_i32->ptr, (const int*)cursor_i32->ptr, dst_x->ptr,
        labels ? dst_y->ptr : nullptr, rows, cols, n_perm, vec);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

int tp_cursor_advance(tp_ctx* ctx, tp_buf* cursor_i32, int delta, int modulo) {
    TP_CHECK_ARG(ctx && modulo > 0, "tp_cursor_advance: bad arguments");
    TP_NEED(cursor_i32, 1, "cursor");
    cursor_advance_kernel<<<1, 1, 0, ctx->stream>>>((int*)cursor_i32->ptr, delta, modulo);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

}  // extern "C"

/* taper_b200.h — C ABI of the B200-native backend for taper's tape-evaluation hot path.
 *
 * This is the drop-in boundary: what a Rust `extern "C"` block in taper would bind in place of its
 * CPU kernels (see INTEGRATION.md for the shim).  Every entry point cites the reference
 * interface it replaces as `path:line` relative to the reference root (vaibhawvipul/taper @ aea74b46).
 *
 * Conventions
 *  - Every call returns 0 (TP_OK) on success; otherwise an error code, with a thread-local
 *    message available from tp_last_error().  The reference panics on the same conditions
 *    (assert!/unwrap, e.g. src/ops.rs:11-15, 201-208); the host shim turns non-zero into a panic.
 *  - No C++ exceptions cross this boundary; no torch / CUDA types appear in signatures.
 *  - tp_ctx  : 1 host thread : 1 device : 1 CUDA stream (: 1 NCCL rank).  Thread-affine — mirrors the
 *              reference's `thread_local!` tape (src/tape.rs:6-9).  Different contexts are independent.
 *  - tp_buf  : intrusively refcounted device buffer of 4-byte elements (fp32 unless stated) ==
 *              the reference's `Arc<RwLock<Vec<f32>>>` (src/tensor.rs:236-244).  The callee never
 *              frees caller buffers.  A NULL tp_buf* for an optional argument means "absent".
 *  - All ops are asynchronous with respect to the host and ordered on the context's stream, except
 *    tp_buf_download / tp_sync / tp_ctx_destroy which synchronise.
 *  - `accumulate` arguments: 0 = first touch (dst  = value, the reference's lazily zero-allocated
 *    grad followed by `+=`, e.g. src/ops.rs:126-128), 1 = dst += value.
 *  - Matrices are dense row-major fp32, exactly as in the reference.
 */
#ifndef TAPER_B200_H
#define TAPER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TP_ABI_VERSION 1

enum {
    TP_OK = 0,
    TP_ERR_INVALID = 1,     /* shape / argument check failed (reference: assert!/panic!) */
    TP_ERR_CUDA = 2,        /* CUDA runtime / driver error                               */
    TP_ERR_OOM = 3,
    TP_ERR_COMM = 4,        /* NCCL / peer-memory error                                  */
    TP_ERR_UNSUPPORTED = 5
};

typedef struct tp_ctx tp_ctx;
typedef struct tp_buf tp_buf;
typedef struct tp_graph tp_graph;
typedef struct tp_event tp_event;

/* ---------------------------------------------------------------------------------------------
 * Runtime: context, stream, buffers  (replaces Vec<f32> allocation + mimalloc, src/main.rs:7-10)
 * ------------------------------------------------------------------------------------------- */
int         tp_abi_version(void);
const char* tp_last_error(void);
int  tp_device_count(int* count);
int  tp_ctx_create(int device, tp_ctx** out);
int  tp_ctx_destroy(tp_ctx* ctx);
int  tp_sync(tp_ctx* ctx);
void* tp_ctx_stream(tp_ctx* ctx);                  /* the cudaStream_t, for interop              */
int  tp_ctx_device(tp_ctx* ctx);
int  tp_ctx_sm_count(tp_ctx* ctx);
/* bytes currently handed out / bytes held by the caching allocator */
int  tp_ctx_mem_stats(tp_ctx* ctx, size_t* in_use, size_t* reserved);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
int  tp_ctx_launch_count(tp_ctx* ctx, uint64_t* count);

int  tp_buf_alloc(tp_ctx* ctx, size_t n, tp_buf** out);             /* n 4-byte elements       */
int  tp_buf_wrap(tp_ctx* ctx, void* device_ptr, size_t n, tp_buf** out); /* external, not owned */
int  tp_buf_slice(tp_buf* parent, size_t offset, size_t n, tp_buf** out); /* view; retains parent */
int  tp_buf_retain(tp_buf* buf);
int  tp_buf_release(tp_buf* buf);
void*  tp_buf_ptr(const tp_buf* buf);
size_t tp_buf_len(const tp_buf* buf);
int  tp_buf_upload(tp_ctx* ctx, tp_buf* dst, const void* host, size_t n);    /* Tensor::new     src/tensor.rs:470-478 */
int  tp_buf_upload_pinned(tp_ctx* ctx, tp_buf* dst, const void* pinned_host, size_t n); /* no staging copy */
int  tp_buf_download(tp_ctx* ctx, const tp_buf* src, void* host, size_t n);  /* Tensor::data()  src/tensor.rs:493-496 */
int  tp_buf_copy(tp_ctx* ctx, tp_buf* dst, const tp_buf* src, size_t n);     /* Vec::clone, e.g. reshape src/tensor.rs:814 */
int  tp_buf_fill(tp_ctx* ctx, tp_buf* dst, float value, size_t n);           /* vec![v; n], e.g. backward seed src/tensor.rs:521 */
/* asynchronous device->pinned-host copy on the context's stream; pair with tp_event_record / tp_event_sync */
int  tp_buf_download_async(tp_ctx* ctx, const tp_buf* src, void* pinned_host, size_t n);
int  tp_host_alloc_pinned(size_t bytes, void** out);
int  tp_host_free_pinned(void* p);

/* Events on the context's stream (timing with CUDA events; host waits for one step's result only). */
int  tp_event_create(tp_ctx* ctx, tp_event** out);
int  tp_event_record(tp_ctx* ctx, tp_event* ev);
int  tp_event_sync(tp_event* ev);
int  tp_event_elapsed_ms(tp_event* start, tp_event* stop, float* ms);
int  tp_event_destroy(tp_event* ev);

/* Input prefetch: a second (copy) stream per context so the host->device copy of step i+1's batch overlaps
 * step i's kernels.  tp_copy_* enqueue on the copy stream; events order the two streams:
 *   copy stream : tp_copy_wait_event(done[i-N]) ; tp_copy_upload_pinned(x[i%N]) ; tp_copy_event_record(ready[i])
 *   main stream : tp_stream_wait_event(ready[i]) ; <step i> ; tp_event_record(done[i])                         */
int  tp_copy_upload_pinned(tp_ctx* ctx, tp_buf* dst, const void* pinned_host, size_t n);
int  tp_copy_event_record(tp_ctx* ctx, tp_event* ev);
int  tp_copy_wait_event(tp_ctx* ctx, tp_event* ev);
int  tp_stream_wait_event(tp_ctx* ctx, tp_event* ev);

/* CUDA-graph capture of everything enqueued on the context's stream between begin and end. */
int  tp_graph_begin(tp_ctx* ctx);
int  tp_graph_end(tp_ctx* ctx, tp_graph** out);
int  tp_graph_launch(tp_ctx* ctx, tp_graph* g);
int  tp_graph_destroy(tp_graph* g);

/* ---------------------------------------------------------------------------------------------
 * The reference's existing operator boundary  (src/gemm.rs:8-19 cblas / :72-83 matrixmultiply;
 * re-exported at src/lib.rs:13).  C[m,n] = alpha*op(A)*op(B) + beta*C, row-major,
 * lda = k (N) | m (T), ldb = n (N) | k (T), ldc = n  (src/gemm.rs:21-29).  trans: 0 = N, 1 = T.
 * ------------------------------------------------------------------------------------------- */
int tp_sgemm_rowmajor(tp_ctx* ctx, int trans_a, int trans_b, int m, int n, int k, float alpha,
                      const tp_buf* a, const tp_buf* b, float beta, tp_buf* c);
/* GEMM math mode for the tcgen05 path: 0 = exact fp32 (CUDA-core FFMA), 1 = 3xTF32 split
 * (fp32-accurate on the tensor cores, default), 2 = 1xTF32 (throughput mode), 3 = bf16x3 (operands split into
 * bf16 hi + lo, three kind::f16 MMAs per product: ~1e-5 of |C|inf, half the tensor-pipe cost of 3xTF32; shapes TMA
 * cannot describe run as mode 1). */
int tp_set_gemm_mode(tp_ctx* ctx, int mode);
int tp_get_gemm_mode(tp_ctx* ctx, int* mode);
/* fp32 -> the pre-split operand format of mode 3: dst holds bf16 hi = rn(x) at [0, n) and lo = rn(x - hi) at
 * [plane, plane + n) with plane = n rounded up to 8 (dst is a tp_buf of >= plane floats = 2 * plane bf16). */
int tp_split_bf16(tp_ctx* ctx, const tp_buf* src, tp_buf* dst, size_t n);

/* ---------------------------------------------------------------------------------------------
 * Elementwise family  (src/tensor.rs:14-234 `pub mod simd`; src/ops.rs:8-151, 312-496)
 * ------------------------------------------------------------------------------------------- */
int tp_add(tp_ctx*, const tp_buf* a, const tp_buf* b, tp_buf* out, size_t n);   /* src/ops.rs:8-30   */
int tp_sub(tp_ctx*, const tp_buf* a, const tp_buf* b, tp_buf* out, size_t n);   /* src/ops.rs:377-396 */
int tp_mul(tp_ctx*, const tp_buf* a, const tp_buf* b, tp_buf* out, size_t n);   /* src/ops.rs:54-71  */
int tp_div(tp_ctx*, const tp_buf* a, const tp_buf* b, tp_buf* out, size_t n);   /* src/ops.rs:440-456 */
/* dst (+)= scale*src : accumulate_grad (scale=1) / accumulate_grad_scaled  src/ops.rs:124-151 */
int tp_accumulate(tp_ctx*, tp_buf* dst, const tp_buf* src, float scale, size_t n, int accumulate);
/* gdst (+)= gout*other           Mul backward  src/ops.rs:79-115 */
int tp_mul_bwd(tp_ctx*, const tp_buf* gout, const tp_buf* other, tp_buf* gdst, size_t n, int accumulate);
/* ga (+)= gout/b ; gb (+)= -gout*a/(b*b)     Div backward  src/ops.rs:466-491 */
int tp_div_bwd_a(tp_ctx*, const tp_buf* gout, const tp_buf* b, tp_buf* ga, size_t n, int accumulate);
int tp_div_bwd_b(tp_ctx*, const tp_buf* gout, const tp_buf* a, const tp_buf* b, tp_buf* gb, size_t n, int accumulate);
int tp_relu_fwd(tp_ctx*, const tp_buf* x, tp_buf* y, size_t n);                 /* src/ops.rs:312-349 */
/* gin (+)= x>0 ? gout : 0        src/ops.rs:358-370 */
int tp_relu_bwd(tp_ctx*, const tp_buf* x, const tp_buf* gout, tp_buf* gin, size_t n, int accumulate);
int tp_exp_fwd(tp_ctx*, const tp_buf* x, tp_buf* y, size_t n);                  /* src/tensor.rs:1091-1101 */
int tp_exp_bwd(tp_ctx*, const tp_buf* y, const tp_buf* gout, tp_buf* gin, size_t n, int accumulate); /* :1109-1127 */
int tp_log_fwd(tp_ctx*, const tp_buf* x, tp_buf* y, size_t n);                  /* src/tensor.rs:1136-1143 */
int tp_log_bwd(tp_ctx*, const tp_buf* x, const tp_buf* gout, tp_buf* gin, size_t n, int accumulate); /* :1150-1163 */

/* ---------------------------------------------------------------------------------------------
 * Broadcast / reduction / layout ops on [rows, cols] matrices
 * ------------------------------------------------------------------------------------------- */
/* out[i,f] = a[i,f] + bias[f] (optionally ReLU)      add_broadcast  src/tensor.rs:636-663 */
int tp_add_broadcast_fwd(tp_ctx*, const tp_buf* a, const tp_buf* bias, tp_buf* out, int rows, int cols, int relu);
/* out[f] (+)= scale * sum_i g[i,f]                   add_broadcast bwd  src/tensor.rs:680-691;  sum(dim=0) :890-940 */
int tp_colsum(tp_ctx*, const tp_buf* g, tp_buf* out, int rows, int cols, float scale, int accumulate);
/* out[i] (+)= scale * sum_c g[i,c]                   sub_broadcast_rows bwd (scale=-1)  src/tensor.rs:748-761; sum(dim=1) */
int tp_rowsum(tp_ctx*, const tp_buf* g, tp_buf* out, int rows, int cols, float scale, int accumulate);
/* out[0] = sum x                                     sum(None)  src/tensor.rs:994-996 */
int tp_sum_all(tp_ctx*, const tp_buf* x, tp_buf* out, size_t n);
/* out[i,c] = a[i,c] - r[i]                           sub_broadcast_rows  src/tensor.rs:707-737 */
int tp_sub_broadcast_rows_fwd(tp_ctx*, const tp_buf* a, const tp_buf* r, tp_buf* out, int rows, int cols);
/* gin[i,c] (+)= g[i] (mode 0) | g[c] (mode 1) | g[0] (mode 2)     sum backward  src/tensor.rs:942-990, 1003-1010 */
int tp_broadcast_bwd(tp_ctx*, const tp_buf* g, tp_buf* gin, int rows, int cols, int mode, int accumulate);
/* row-wise (dim=1) / column-wise (dim=0) max with first-max-wins indices stored as f32  src/tensor.rs:1021-1069 */
int tp_max_rows(tp_ctx*, const tp_buf* x, tp_buf* vals, tp_buf* idx, int rows, int cols);
int tp_max_cols(tp_ctx*, const tp_buf* x, tp_buf* vals, tp_buf* idx, int rows, int cols);
/* global max, LAST of equal maxima (Iterator::max_by)  src/tensor.rs:1072-1080 */
int tp_max_all(tp_ctx*, const tp_buf* x, tp_buf* val, tp_buf* idx, size_t n);
/* y[j,i] (+)= x[i,j]                                 transpose fwd/bwd  src/tensor.rs:544-591 */
int tp_transpose2d(tp_ctx*, const tp_buf* x, tp_buf* y, int rows, int cols, int accumulate);

/* ---------------------------------------------------------------------------------------------
 * Fused Linear  (src/nn.rs:54-60 = transpose src/tensor.rs:544 + matmul src/ops.rs:200 +
 * add_broadcast src/tensor.rs:636; backward = the three closures of those ops)
 *   fwd:  Y[B,out] = X[B,in] * W[out,in]^T + b[out]   (optionally ReLU; b may be NULL)
 *   bwd:  dX[B,in] (+)= dY * W ;  dW[out,in] (+)= dY^T * X ;  db[out] (+)= colsum(dY)
 *         any of dX/dW/db may be NULL (operand does not require grad).
 *         relu_mask_y != NULL applies dY := dY * [Y>0] first (ReLU backward src/ops.rs:358-370 fused).
 * ------------------------------------------------------------------------------------------- */
int tp_linear_fwd(tp_ctx*, const tp_buf* x, const tp_buf* w, const tp_buf* b, tp_buf* y,
                  int batch, int in_features, int out_features, int relu);
int tp_linear_bwd(tp_ctx*, const tp_buf* x, const tp_buf* w, const tp_buf* dy, const tp_buf* relu_mask_y,
                  tp_buf* dx, tp_buf* dw, tp_buf* db, int batch, int in_features, int out_features,
                  int acc_dx, int acc_dw, int acc_db);

/* ---------------------------------------------------------------------------------------------
 * Softmax / cross-entropy / accuracy  (src/loss.rs:82-195, 271-290)
 * ------------------------------------------------------------------------------------------- */
int tp_log_softmax_fwd(tp_ctx*, const tp_buf* x, tp_buf* logp, int rows, int cols);   /* src/loss.rs:101-126 */
int tp_softmax_fwd(tp_ctx*, const tp_buf* x, tp_buf* p, int rows, int cols);          /* src/loss.rs:82-98 (intent, A13) */
/* fused: logp = log_softmax(logits); loss[0] = -(1/B) sum_i logp[i, (int)t_i]   src/loss.rs:152-165.
 * A target outside [0, cols) sets the context's sticky device error flag (reference asserts, :161). */
int tp_softmax_xent_fwd(tp_ctx*, const tp_buf* logits, const tp_buf* targets, tp_buf* logp, tp_buf* loss,
                        int rows, int cols);
/* the same plus accuracy's correct count (src/loss.rs:271-290; correct may be NULL) in ONE launch: the head of a training
 * step (src/train.rs:112-115).  Block partials are folded in block order by the last block to finish (deterministic). */
int tp_softmax_xent_acc_fwd(tp_ctx*, const tp_buf* logits, const tp_buf* targets, tp_buf* logp, tp_buf* loss, tp_buf* correct,
                            int rows, int cols);
/* glogits (+)= (exp(logp) - onehot(t)) * gloss[0]/B                               src/loss.rs:174-191 */
int tp_softmax_xent_bwd(tp_ctx*, const tp_buf* logp, const tp_buf* targets, const tp_buf* gloss,
                        tp_buf* glogits, int rows, int cols, int accumulate);
/* correct[0] = #{ i : |argmax_c pred[i,c] - t_i| < 1e-6 } as f32                    src/loss.rs:271-290 */
int tp_accuracy_count(tp_ctx*, const tp_buf* pred, const tp_buf* targets, tp_buf* correct, int rows, int cols);
/* reads and clears the sticky device error flag (0 = none, 1 = target class out of bounds) */
int tp_ctx_device_error(tp_ctx* ctx, int* flag);
/* the same as an asynchronous copy on the context's stream into (pinned) host memory: ordered after the kernels enqueued so far */
int tp_ctx_device_error_async(tp_ctx* ctx, int* host_flag);

/* ---------------------------------------------------------------------------------------------
 * Convolution and pooling, NCHW fp32  (src/tensor.rs:1221-1285, 1391-1660, 1663-1780, 1972-2076)
 * Weight buffer is the reference's [C_out,C_in,kh,kw] allocation REINTERPRETED as [K=C_in*kh*kw, C_out]
 * (src/tensor.rs:1262): element (k, co) at flat k*C_out + co.
 * ------------------------------------------------------------------------------------------- */
typedef struct tp_conv_desc {
    int n, c_in, h, w;          /* input  [N, C_in, H, W]          */
    int c_out, kh, kw;          /* weight [C_out, C_in, kh, kw]    */
    int stride_h, stride_w, pad_h, pad_w, dil_h, dil_w;
} tp_conv_desc;
int tp_conv2d_out_dims(const tp_conv_desc* d, int* h_out, int* w_out);     /* src/tensor.rs:1254-1255 */
/* col[(n,oh,ow), ci*kh*kw + kr*kw + kc]                im2col  src/tensor.rs:1663-1780 */
int tp_im2col(tp_ctx*, const tp_buf* x, tp_buf* col, const tp_conv_desc* d);
/* gx (+)= col2im(gcol)  — adjoint of im2col (the link the reference drops at src/tensor.rs:1725) */
int tp_col2im(tp_ctx*, const tp_buf* gcol, tp_buf* gx, const tp_conv_desc* d, int accumulate);
/* y[N,C_out,Ho,Wo] = conv(x, w) + b, optional ReLU     conv2d / conv2d_relu  src/tensor.rs:1221-1285, 1379-1389 */
int tp_conv2d_fwd(tp_ctx*, const tp_buf* x, const tp_buf* w, const tp_buf* b, tp_buf* y,
                  const tp_conv_desc* d, int relu);
/* gy is the gradient w.r.t. the conv output (after undoing ReLU if fused: pass relu_mask_y).
 *   db[co]   (+)= sum_{n,oh,ow} gy                       add_bias_4d backward  src/tensor.rs:2003-2027
 *   dw[K,Co] (+)= col^T * gy_nhwc ;  dx (+)= col2im(gy_nhwc * w^T)   (full adjoint; NULL to skip —
 *   the reference computes neither, Appendix A1).  For 3x3 / s1 / p1 layers with channel counts that are multiples of 32 both
 *   run as implicit GEMMs on the tensor cores (dx: the forward kernel on the masked gradient with mirrored, transposed
 *   weights; dw: a contraction over pixels on MN-major bf16 hi/lo planes) — neither the im2col matrix nor its gradient exists */
int tp_conv2d_bwd(tp_ctx*, const tp_buf* x, const tp_buf* w, const tp_buf* gy, const tp_buf* relu_mask_y,
                  tp_buf* dx, tp_buf* dw, tp_buf* db, const tp_conv_desc* d,
                  int acc_dx, int acc_dw, int acc_db);
/* A chain of n_layers (<= 4) small Linear(+ReLU) layers — every width dims[l] <= 128, dims has n_layers + 1 entries — as one
 * forward and one backward launch (+ a deterministic fold of per-CTA partials): exact fp32 FFMA with all weights in shared
 * memory.  Replaces, per layer, Linear::forward (src/nn.rs:54-60: transpose src/tensor.rs:544-591, matmul src/ops.rs:200-228,
 * add_broadcast src/tensor.rs:636-704), ReLU (src/ops.rs:312-374) and their backward closures (src/ops.rs:254-291, 358-370,
 * src/tensor.rs:680-691) for classifier heads like the example CNN's 128-128-64-10 (examples/train_mnist_cnn.rs:80-100).
 *   fwd: acts[l] [batch, dims[l+1]] = relu?(in_l . W_l^T + b_l), in_0 = x, in_l = acts[l-1]; weights[l] is [dims[l+1], dims[l]].
 *   bwd: gout = gradient of acts[n_layers-1]; dw[l] / db[l] / dx (each may be NULL) are stored (acc 0) or added to (acc 1). */
int tp_mlp_small_supported(int n_layers, const int* dims, int batch);
int tp_mlp_small_fwd(tp_ctx*, const tp_buf* x, int n_layers, const int* dims, const tp_buf* const* weights,
                     const tp_buf* const* biases, const int* relu, tp_buf* const* acts, int batch);
int tp_mlp_small_bwd(tp_ctx*, const tp_buf* x, int n_layers, const int* dims, const tp_buf* const* weights, const int* relu,
                     const tp_buf* const* acts, const tp_buf* gout, tp_buf* dx, tp_buf* const* dw, tp_buf* const* db,
                     int acc_dx, const int* acc_dw, const int* acc_db, int batch);
/* A stack of n_layers (<= 8) 3x3 / stride 1 / pad 1 convolutions, each + bias (+ ReLU when relu[l]) and, when pool[l],
 * followed by a 2x2 / stride 2 max-pool, evaluated back to back:  x [N, C_in, H, W] fp32 NCHW -> y NCHW fp32 (the last
 * layer's output, pooled if pool[last]).  Replaces the chain conv2d_relu -> max_pool2d -> conv2d_relu ... that
 * Sequential::forward runs over examples/train_mnist_cnn.rs:35-100 (Tensor::conv2d src/tensor.rs:1221-1285, conv2d_relu
 * :1379-1389, max_pool2d :1391-1464) when nothing but the last output is observed — the strict-reference tape, where
 * im2col drops the autograd link, so no intermediate activation or pooling index is ever read again.
 * weights[l]: the reference's [C_out, C_in, 3, 3] buffer read as [K = ci*9 + kr*3 + kc, C_out] (Appendix A2); biases[l]
 * may be NULL.  Intermediate activations stay in the tensor cores' operand format (NHWC bf16 hi/lo pairs); products are
 * bf16x3 (three bf16 MMAs per fp32 product, ~1e-5 of |y|inf), accumulation fp32.
 * Shapes: C_in % 32 == 0 or C_in * 9 <= 36 (direct fp32 first layer, needs n_layers >= 2); every C_out in {32, 64, 128};
 * W < 32.  Anything else: TP_ERR_UNSUPPORTED and nothing has been launched (callers run the layers one by one). */
int tp_conv_stack_fwd(tp_ctx*, const tp_buf* x, int n, int c_in, int h, int w, int n_layers,
                      const tp_buf* const* weights, const tp_buf* const* biases, const int* c_out,
                      const int* pool, const int* relu, tp_buf* y);
/* The same stack with AdaptiveAvgPool2d::global on top (src/nn.rs:670-686): mean[n,c] over the last layer's plane and, when
 * cnt != NULL, cnt[n,c] = #{units of the plane > 0}; the last layer's activation itself is not written when its tiles hold whole
 * images (7x7 and smaller: the pooling happens in the conv epilogue), otherwise it goes through a scratch buffer. */
int tp_conv_stack_gap_fwd(tp_ctx*, const tp_buf* x, int n, int c_in, int h, int w, int n_layers,
                          const tp_buf* const* weights, const tp_buf* const* biases, const int* c_out,
                          const int* pool, const int* relu, tp_buf* mean, tp_buf* cnt);
/* Global average pool fused with the per-plane count of positive units:  mean[n,c] = sum_p y[n,c,p] / hw
 * (AdaptiveAvgPool2d::global -> avg_pool2d, src/nn.rs:670-686, src/tensor.rs:1524-1590), cnt[n,c] = #{p : y[n,c,p] > 0}
 * (NULL to skip).  cnt is all that the backward of [conv bias -> ReLU -> global average pool] needs: */
int tp_gap_count_fwd(tp_ctx*, const tp_buf* y, tp_buf* mean, tp_buf* cnt, int n, int c, int hw);
/* gb[c] (+)= sum_n (g[n,c] / hw) * cnt[n,c]  — avg_pool2d backward (g / hw to every unit, src/tensor.rs:1600-1655), ReLU
 * backward (src/ops.rs:358-370) and add_bias_4d backward (sum over n,h,w, src/tensor.rs:2003-2027) as one reduction over
 * [N, C]; cnt == NULL: no ReLU between the bias and the pool.  Deterministic (fixed summation order). */
int tp_gap_relu_bias_grad(tp_ctx*, const tp_buf* g, const tp_buf* cnt, tp_buf* gb, int n, int c, int hw, int accumulate);
/* y = x + bias[c] per channel; gb[c] (+)= sum_{n,hw} g     add_bias_4d  src/tensor.rs:1972-2031 */
int tp_add_bias_4d(tp_ctx*, const tp_buf* x, const tp_buf* bias, tp_buf* y, int n, int c, int hw, int relu);
int tp_bias_grad_4d(tp_ctx*, const tp_buf* g, tp_buf* gb, int n, int c, int hw, int accumulate);
/* y[N,C,H,W] = x[N,H,W,C]                               transpose_4d([0,3,1,2])  src/tensor.rs:2034-2076 */
int tp_nhwc_to_nchw(tp_ctx*, const tp_buf* x, tp_buf* y, int n, int h, int w, int c);
int tp_nchw_to_nhwc(tp_ctx*, const tp_buf* x, tp_buf* y, int n, int c, int h, int w);

typedef struct tp_pool_desc {
    int n, c, h, w, kh, kw, stride_h, stride_w, pad_h, pad_w;
} tp_pool_desc;
int tp_pool_out_dims(const tp_pool_desc* d, int* h_out, int* w_out);       /* src/tensor.rs:1406-1407 */
/* max_pool2d: strict '>' scan kh then kw from -inf, first max wins; argmax = absolute flat input
 * index as int32 (reference: usize)                          src/tensor.rs:1391-1464 */
int tp_maxpool2d_fwd(tp_ctx*, const tp_buf* x, tp_buf* y, tp_buf* argmax_i32, const tp_pool_desc* d);
/* gin plane := 0 then gin[argmax] += gout   (overwrites, Appendix A6)   src/tensor.rs:1476-1516 */
int tp_maxpool2d_bwd(tp_ctx*, const tp_buf* gout, const tp_buf* argmax_i32, tp_buf* gin, const tp_pool_desc* d);
/* avg_pool2d: window sum over in-bounds taps / (kh*kw)       src/tensor.rs:1524-1590; bwd accumulates :1600-1655 */
int tp_avgpool2d_fwd(tp_ctx*, const tp_buf* x, tp_buf* y, const tp_pool_desc* d);
int tp_avgpool2d_bwd(tp_ctx*, const tp_buf* gout, tp_buf* gin, const tp_pool_desc* d, int accumulate);

/* ---------------------------------------------------------------------------------------------
 * the remaining elementwise / loss ops of the reference's XOR demo (src/main.rs) and MSE path
 * ------------------------------------------------------------------------------------------- */
int tp_sigmoid_fwd(tp_ctx*, const tp_buf* x, tp_buf* y, size_t n);                            /* src/tensor.rs:594-610 */
int tp_sigmoid_bwd(tp_ctx*, const tp_buf* y, const tp_buf* gout, tp_buf* gin, size_t n, int accumulate);   /* g*s*(1-s)  :617-629 */
int tp_pow_fwd(tp_ctx*, const tp_buf* x, tp_buf* y, float exponent, size_t n);                /* powf; sqrt = pow(0.5)  :1172-1211 */
int tp_pow_bwd(tp_ctx*, const tp_buf* x, const tp_buf* gout, tp_buf* gin, float exponent, size_t n, int accumulate);
int tp_mean_fwd(tp_ctx*, const tp_buf* x, tp_buf* out1, size_t n);                            /* sum / len  :772-776 */
int tp_mean_bwd(tp_ctx*, const tp_buf* gout1, tp_buf* gin, size_t n, int accumulate);         /* gin += g/len  :784-796 */
/* bce_loss: -mean(y ln p + (1-y) ln(1-p)), p clamped to [1e-7, 1-1e-7]; gradients w.r.t. predictions and (optionally) targets
 * scaled by the upstream scalar gloss[0]                                                     src/loss.rs:6-72 */
int tp_bce_fwd(tp_ctx*, const tp_buf* pred, const tp_buf* target, tp_buf* loss1, size_t n);
int tp_bce_bwd(tp_ctx*, const tp_buf* pred, const tp_buf* target, const tp_buf* gloss1, tp_buf* gpred, tp_buf* gtarget, size_t n,
               int acc_pred, int acc_target);

/* ---------------------------------------------------------------------------------------------
 * Optimizer steps  (src/optim.rs:21-33, 83-113, 148-168).  grad_scale multiplies g first
 * (1/world for data-parallel averaging; 1 otherwise).
 * ------------------------------------------------------------------------------------------- */
int tp_sgd_step(tp_ctx*, tp_buf* p, const tp_buf* g, float lr, float grad_scale, size_t n);
/* the same with lr read from lr1[0] on the device, so that a captured (CUDA-graph) step follows SGD::set_lr; tp_buf_set_scalar
 * stores one float on the context's stream */
int tp_sgd_step_dev(tp_ctx*, tp_buf* p, const tp_buf* g, const tp_buf* lr1, float grad_scale, size_t n);
int tp_buf_set_scalar(tp_ctx*, tp_buf* buf, size_t index, float value);
/* g' = g*grad_scale + wd*p; m = b1*m+(1-b1)*g'; v = b2*v+(1-b2)*g'*g'; p -= step_size*m/(sqrt(v)+eps)
 * step_size = lr*sqrt(1-b2^t)/(1-b1^t) is computed by the caller (tp_adam_step_size). */
int tp_adam_step(tp_ctx*, tp_buf* p, const tp_buf* g, tp_buf* m, tp_buf* v, float step_size,
                 float beta1, float beta2, float eps, float weight_decay, float grad_scale, size_t n);
/* fused AdamW: p *= (1 - lr*wd) first (src/optim.rs:154-161), then Adam with wd = 0 */
int tp_adamw_step(tp_ctx*, tp_buf* p, const tp_buf* g, tp_buf* m, tp_buf* v, float step_size,
                  float beta1, float beta2, float eps, float decay_factor, float grad_scale, size_t n);
int tp_scale(tp_ctx*, tp_buf* p, float s, size_t n);            /* AdamW decay of grad-less params */
/* Device-resident Adam state so that a captured (CUDA-graph) step replays without the host patching
 * arguments: hyper[8] = { t (int32 bits), lr, beta1, beta2, eps, weight_decay, step_size, decay }.
 *   tp_adam_advance : t += 1; step_size = lr*sqrt(1-b2^t)/(1-b1^t); decay = 1 - lr*wd   (src/optim.rs:86-90, 157)
 *   tp_adam_step_dev: Adam::step body (decoupled=0, src/optim.rs:93-110) or AdamW (decoupled=1, :154-164)
 *   tp_decay_dev    : p *= decay for parameters without a gradient under AdamW (:154-161)            */
int tp_adam_hyper_init(tp_ctx*, tp_buf* hyper, float lr, float beta1, float beta2, float eps, float weight_decay);
int tp_adam_hyper_set_lr(tp_ctx*, tp_buf* hyper, float lr);     /* Adam::set_lr  src/optim.rs:125-127 */
int tp_adam_advance(tp_ctx*, tp_buf* hyper);
int tp_adam_step_dev(tp_ctx*, tp_buf* p, const tp_buf* g, tp_buf* m, tp_buf* v, const tp_buf* hyper,
                     float grad_scale, int decoupled, size_t n);
int tp_decay_dev(tp_ctx*, tp_buf* p, const tp_buf* hyper, size_t n);
/* the same over up to any number of slices of the flat arenas in one launch per 32 slices: modes[i] = 1 Adam / AdamW step,
 * 2 AdamW decay only (a parameter whose gradient is None still decays, src/optim.rs:154-161), 0 skip (Adam skips
 * gradient-less parameters, :93).  Slices start at multiples of 4 elements and are padded to 4 with zeros. */
int tp_adam_step_segments(tp_ctx*, tp_buf* p, const tp_buf* g, tp_buf* m, tp_buf* v, const tp_buf* hyper, float grad_scale,
                          int decoupled, const int64_t* offsets, const int64_t* lengths, const int* modes, int n_segments);
/* lr * sqrt(1 - b2^t) / (1 - b1^t) with f32::powi semantics   src/optim.rs:88-90 */
float tp_adam_step_size(float lr, float beta1, float beta2, int t);

/* ---------------------------------------------------------------------------------------------
 * Device-resident dataset: MNISTDataset::get_batch + DataLoader (src/data/mnist.rs:276-309, 326-385).
 *   dst_x[r, :] = images[perm[(cursor + r) % n_perm], :],  dst_y[r] = labels[perm[...]]   for r < rows
 *   perm is an int32 index buffer (the shuffled order); cursor is a 1-element int32 device counter that
 *   tp_cursor_advance moves by `delta` modulo `modulo`, so a captured step walks the dataset by itself.
 * ------------------------------------------------------------------------------------------- */
int tp_gather_batch(tp_ctx*, const tp_buf* images, const tp_buf* labels, const tp_buf* perm_i32, const tp_buf* cursor_i32,
                    tp_buf* dst_x, tp_buf* dst_y, int rows, int cols, int n_perm);
int tp_cursor_advance(tp_ctx*, tp_buf* cursor_i32, int delta, int modulo);

/* ---------------------------------------------------------------------------------------------
 * Device tape: one whole training step of a small MLP as ONE persistent kernel (one CTA per SM, grid barrier between phases).
 * Replaces the loop body of Trainer::train_epoch (src/train.rs:106-138: Tape::reset, model.forward
 * [Linear src/nn.rs:54-60, ReLU src/ops.rs:312-374], cross_entropy_loss src/loss.rs:136-195, accuracy
 * src/loss.rs:271-290, loss.backward src/tape.rs:106-127, optimizer.step src/optim.rs:21-33 / 83-113 /
 * 148-168, zero_grad) for models that are a chain of Linear(+ReLU) layers ending in a classifier of at
 * most 16 classes.  The job list is compiled once per (model, batch size); the kernel walks it phase by
 * phase with a grid barrier between dependent phases.  Exact fp32 (FFMA), deterministic.
 *   tp_step_supported : 1 if the fused step can run this description (feature widths multiples of 4,
 *                       classes <= 16, batch <= 4096, < 1.5 GFLOP per step: larger models are faster
 *                       on the tcgen05 GEMM path)
 *   tp_step_run       : perm_i32 == NULL: x [batch,in] / labels [batch] are the batch (host-fed);
 *                       otherwise x / labels are the resident dataset [n_perm,...] and rows
 *                       perm[(cursor + r) % n_perm] are gathered in-kernel; cursor advances by batch.
 *                       cursor_value >= 0 is the caller's mirror of *cursor (saves the kernel a dependent
 *                       load), -1 reads the device word.  result_host (optional) is a pinned, device-
 *                       mapped {loss, correct, seq} slot (3 words) the kernel also writes, so no separate D2H copy
 *                       is needed; with result_seq != 0 the kernel stores it to word 2 after the other two
 *                       (system-scope fence in between), so the host can poll the slot instead of a CUDA event.
 *                       Writes result = {loss, #correct}; Adam state (hyper, see tp_adam_hyper_init)
 *                       advances on the device exactly as tp_adam_advance + tp_adam_step_dev would.
 * ------------------------------------------------------------------------------------------- */
#define TP_STEP_MAX_LAYERS 8
typedef struct tp_step tp_step;
typedef struct tp_step_desc {
    int n_layers;                            /* Linear layers; the last one is the classifier head */
    int dims[TP_STEP_MAX_LAYERS + 1];        /* in, hidden..., classes                              */
    int relu[TP_STEP_MAX_LAYERS];            /* ReLU after layer l (ignored for the last)           */
    int batch;
    int optimizer;                           /* 0 SGD, 1 Adam, 2 AdamW                              */
    int64_t w_off[TP_STEP_MAX_LAYERS];       /* offset of W_l [out,in] in the flat arenas           */
    int64_t b_off[TP_STEP_MAX_LAYERS];       /* offset of b_l [out], or -1 if the layer has no bias */
    int64_t arena_len;                       /* elements in params / grads / m / v                  */
    int materialize_grads;                   /* 1: leave the folded gradients in the grads arena (for inspection); 0: the
                                                optimizer phase sums the split-K partials itself                       */
    int data_parallel;                       /* wide plan: 1 = sum the gradient arena over the context's NCCL communicator
                                                (tp_comm_init) between the fold and the optimizer kernels             */
} tp_step_desc;
/* Data-parallel gradient exchange INSIDE the step kernel, over NVLink peer memory (one process per GPU on one node;
 * no counterpart in the reference, which is single-process).  Every rank owns a window {per-slice flags, gradient slots
 * [2 parities][world source ranks][arena_len]} allocated with cudaMalloc and exported as a 64-byte cudaIpcMemHandle; the
 * launcher distributes the handles (torch.distributed / MPI / files) and tp_xchg_connect maps the peers' windows.
 * In a step created with an exchange, the optimizer phase of the kernel becomes allreduce + optimizer in one: the CTA that
 * owns a 256-float4 slice of the arena folds its local gradient slice, stores it into its slot of every peer's window,
 * raises that slice's flag there (st.release.sys), waits for the peers' flags of the same slice (ld.acquire.sys) and sums the
 * world's slices in rank order (replicas stay bit-identical) before SGD / Adam.  No grid-wide or host-side wait is involved.
 * Every rank must run the same sequence of steps. */
typedef struct tp_xchg tp_xchg;
int tp_xchg_create(tp_ctx* ctx, size_t arena_len, int rank, int world, tp_xchg** out);
int tp_xchg_handle(tp_xchg* x, void* out64);
int tp_xchg_connect(tp_xchg* x, const void* handles_world_x_64, int world);
int tp_xchg_destroy(tp_xchg* x);

int tp_step_supported(const tp_step_desc* desc);
int tp_step_create(tp_ctx* ctx, const tp_step_desc* desc, tp_buf* params, tp_buf* grads, tp_buf* m, tp_buf* v,
                   tp_buf* hyper, tp_buf* result, tp_xchg* xchg, tp_step** out);
int tp_step_run(tp_ctx* ctx, tp_step* step, const tp_buf* x, const tp_buf* labels, const tp_buf* perm_i32,
                tp_buf* cursor_i32, int n_perm, int cursor_value, float sgd_lr, float grad_scale, float* result_host,
                uint32_t result_seq);
/* Wide models (at least one hidden layer, feature widths multiples of 8, >= 1.5 GFLOP per step; TAPER_STEP_WIDE=0/1 forces the
 * choice) run as a fixed plan of tcgen05 kernels chained with programmatic dependent launch instead of the persistent kernel:
 * bf16x3 GEMMs on pre-split operands (tp_set_gemm_mode 3: ~1e-5 of |C|inf) with bias / ReLU / ReLU-mask / bias-gradient /
 * operand-split epilogues, a fused classifier head and a fused optimizer step; data-parallel runs sum the gradient arena with
 * the context's NCCL communicator (tp_comm_init).  Same entry points, same results contract.
 *   tp_step_kind    : 0 unsupported, 1 persistent kernel, 2 wide plan: what tp_step_create would build
 *   tp_step_run_u8  : as tp_step_run with x holding u8 pixels (MNIST's on-disk format; the plan divides by 255 on the
 *                     device, src/data/mnist.rs:225): 4x fewer bytes over PCIe / HBM.  Wide plan only.
 *   tp_step_refresh : tell the step that the parameter arena was written by somebody else (upload, broadcast, another step):
 *                     the plan re-derives its bf16 operand planes on the next run */
int tp_step_kind(const tp_step_desc* desc);
int tp_step_run_u8(tp_ctx* ctx, tp_step* step, const tp_buf* x_u8, const tp_buf* labels, const tp_buf* perm_i32,
                   tp_buf* cursor_i32, int n_perm, int cursor_value, float sgd_lr, float grad_scale, float* result_host,
                   uint32_t result_seq);
int tp_step_refresh(tp_step* step);
int tp_step_is_wide(const tp_step* step);                            /* 1: this step is the wide plan */
int tp_step_info(const tp_step* step, int* n_phases, int* n_jobs, int* grid);
/* per-CTA SM-clock stamps of the last run: [grid][slots] = entry, setup done, then {work done, barrier passed}
 * for each phase (evidence for profiles/: where the step's time goes) */
int tp_step_set_profile(tp_step* step, int on);
int tp_step_read_profile(tp_step* step, int64_t* out, size_t cap, int* slots);
int tp_step_destroy(tp_step* step);

/* ---------------------------------------------------------------------------------------------
 * Data-parallel gradient exchange (no counterpart in the reference: it is single-process).
 * One rank per context.  `unique_id` is the 128-byte ncclUniqueId produced by tp_comm_unique_id on
 * rank 0 and distributed by the launcher (torch.distributed / MPI / files).
 * ------------------------------------------------------------------------------------------- */
int tp_comm_unique_id(void* out128);
int tp_comm_init(tp_ctx* ctx, int rank, int world, const void* unique_id128);
int tp_comm_destroy(tp_ctx* ctx);
int tp_allreduce_sum(tp_ctx* ctx, tp_buf* buf, size_t n);       /* in place, fp32, on the ctx stream */
int tp_broadcast(tp_ctx* ctx, tp_buf* buf, size_t n, int root);

#ifdef __cplusplus
}
#endif
#endif /* TAPER_B200_H */

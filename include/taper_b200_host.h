/* taper_b200_host.h — flat C entry points of the C++ host layer (taper::nn / optim / train in
 * taper_b200/csrc/host/taper.hpp), for drivers that cannot include C++: the Python test-suite, bench.py
 * and, later, a Rust shim that wants whole-module granularity instead of op granularity.
 *
 * The kernel-level boundary is include/taper_b200.h; everything here is composed from it.  Same
 * conventions: every call returns 0 on success, tp_last_error() holds the message otherwise (the
 * reference panics); one host thread : one context : one stream.  Citations are path:line in
 * vaibhawvipul/taper @ aea74b46.
 */
#ifndef TAPER_B200_HOST_H
#define TAPER_B200_HOST_H

#include "taper_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tp_model tp_model;       /* nn::Sequential            src/nn.rs:130-162 */
typedef struct tp_trainer tp_trainer;   /* train::Trainer            src/train.rs:73-293 */
typedef struct tp_dataset tp_dataset;   /* data::MNISTDataset        src/data/mnist.rs:21-25 */
typedef struct tp_loader tp_loader;     /* data::DataLoader          src/data/mnist.rs:326-385 */
typedef struct tp_scheduler tp_scheduler;   /* optim::LRScheduler    src/optim.rs:184-352 */
typedef struct tp_tensor tp_tensor;     /* Tensor (a shared handle: clone = Tensor::clone)   src/tensor.rs:236-244 */
typedef struct tp_optimizer tp_optimizer;   /* optim::{SGD, Adam, AdamW}                     src/optim.rs:8-181 */

/* device of this thread's context (call before anything else on the thread); the context itself */
int tp_host_set_device(int device);
int tp_host_ctx(tp_ctx** out);
/* switches (pass -1 to leave one unchanged):
 *   conv_full_adjoint      0 = reproduce the reference's cut conv autograd chain (default), 1 = full adjoint
 *   fuse_linear_relu       1 = Sequential runs Linear+ReLU as one fused launch (default)
 *   reference_op_sequence  1 = Linear records transpose -> matmul -> add_broadcast exactly as src/nn.rs:54-60
 *   gemm_mode              0 = fp32 FFMA, 1 = 3xTF32 tcgen05 (default), 2 = 1xTF32 tcgen05, 3 = bf16x3 tcgen05 */
int tp_host_config(int conv_full_adjoint, int fuse_linear_relu, int reference_op_sequence, int gemm_mode);
/* 1 (default) = Sequential runs a chain of >= 2 [Conv2d(ReLU) 3x3/s1/p1 (+ MaxPool2d 2x2)] layers as one tp_conv_stack_fwd
 * (strict-reference conv autograd only: nothing but the chain's last output is read again); 0 = layer by layer,
 * as Sequential::forward does in the reference (src/nn.rs:149-151) */
int tp_host_config_conv_stack(int fuse_conv_stack);
/* 1 (default) = Sequential runs a chain of >= 2 Linear(+ReLU) layers of width <= 128 as one forward / one backward launch
 * (tp_mlp_small_*); 0 = one fused Linear launch per layer */
int tp_host_config_small_mlp(int fuse_small_mlp);

/* Sequential from a comma-separated layer list (the constructors of src/nn.rs, src/activation.rs):
 *   linear:IN:OUT[:nobias] | relu | sigmoid | conv:CIN:COUT:K:STRIDE:PAD | conv_relu:CIN:COUT:K:STRIDE:PAD |
 *   maxpool:K:STRIDE | avgpool:K:STRIDE | gap (AdaptiveAvgPool2d::global) | flatten | dropout:PERCENT
 * Weights are drawn from the reference's init distributions with a seeded generator. */
int tp_model_create(const char* spec, uint64_t seed, tp_model** out);
int tp_model_destroy(tp_model* m);
int tp_model_num_params(tp_model* m, int* count);                    /* Module::parameters  src/nn.rs:17 */
int tp_model_param_info(tp_model* m, int index, size_t* numel, int* ndim, size_t* dims4);
int tp_model_set_param(tp_model* m, int index, const float* host, size_t n);
int tp_model_get_param(tp_model* m, int index, float* host, size_t n);
int tp_model_get_grad(tp_model* m, int index, float* host, size_t n, int* has_grad);   /* has_grad = 0 <=> None */
int tp_model_zero_grad(tp_model* m);
/* Module::forward on a host batch; out receives the logits (out_cap floats available) */
int tp_model_forward(tp_model* m, const float* x, const size_t* shape, int ndim, float* out, size_t out_cap, size_t* out_n);
/* Tape::reset; forward; cross_entropy_loss; accuracy; loss.backward()  — no optimizer (src/train.rs:108-121) */
int tp_model_loss_backward(tp_model* m, const float* x, const size_t* shape, int ndim, const float* labels,
                           float* loss, float* correct, size_t* tape_len);

/* Tape::reset; forward; loss(out, targets); loss.backward() with loss_kind "bce" (src/loss.rs:6-72), "mse" (:75-80) or
 * "ce_onehot" (:202-245); targets have the model output's shape.  Gradients are read with tp_model_get_grad. */
int tp_model_regression_backward(tp_model* m, const float* x, const size_t* shape, int ndim, const float* targets,
                                 size_t n_targets, const char* loss_kind, float* loss);

/* Trainer over a model; optimizer is "sgd" | "adam" | "adamw" (src/optim.rs).  The trainer shares the model's
 * parameters (they are re-homed into one flat arena). */
int tp_trainer_create(tp_model* m, const char* optimizer, float lr, float beta1, float beta2, float eps,
                      float weight_decay, tp_trainer** out);
int tp_trainer_destroy(tp_trainer* t);
int tp_trainer_set_lr(tp_trainer* t, float lr);                      /* Adam::set_lr  src/optim.rs:125-127 */
int tp_trainer_set_use_graph(tp_trainer* t, int on);
/* fused device step (tp_step_* in taper_b200.h: the whole loop body as one persistent kernel); on by default, used
 * whenever model, optimizer and batch qualify, otherwise the step runs through the tape + CUDA-graph path */
int tp_trainer_set_use_fused(tp_trainer* t, int on);
int tp_trainer_fused_steps(tp_trainer* t, uint64_t* count);
/* one train_epoch iteration (src/train.rs:106-138), synchronous, host inputs */
int tp_trainer_step(tp_trainer* t, const float* images, const float* labels, size_t batch, const size_t* sample_shape,
                    int ndim, float* loss, float* correct);
/* the same, asynchronous: inputs are copied from (pinned) host memory on the stream; results are read with
 * tp_trainer_fetch in FIFO order (at most 8 outstanding) */
int tp_trainer_step_async(tp_trainer* t, const float* images, const float* labels, size_t batch,
                          const size_t* sample_shape, int ndim, int pinned);
/* device-resident dataset + on-device batch gather (MNISTDataset::get_batch, src/data/mnist.rs:276-309); perm: the sample
 * order, EXACTLY n entries (DataLoader's shuffled indices, src/data/mnist.rs:326-358), or NULL for 0..n-1 */
int tp_trainer_load_dataset(tp_trainer* t, const float* images, const float* labels, size_t n,
                            const size_t* sample_shape, int ndim, const uint32_t* perm);
int tp_trainer_step_resident(tp_trainer* t, size_t batch);
/* the same two entry points with u8 pixels (MNIST's on-disk format; the reference divides by 255 at load time,
 * src/data/mnist.rs:225): the wide step plan (tp_step_run_u8) divides on the device, so a quarter of the bytes cross PCIe
 * and sit in HBM; models without a wide plan widen a host-fed u8 batch on the host and refuse a u8 resident dataset */
int tp_trainer_step_async_u8(tp_trainer* t, const void* images_u8, const float* labels, size_t batch,
                             const size_t* sample_shape, int ndim, int pinned);
int tp_trainer_load_dataset_u8(tp_trainer* t, const void* images_u8, const float* labels, size_t n,
                               const size_t* sample_shape, int ndim, const uint32_t* perm);
/* 2 if the last fused step ran the wide tcgen05 plan, 1 the persistent kernel, 0 none yet */
int tp_trainer_fused_kind(tp_trainer* t, int* kind);
int tp_trainer_fetch(tp_trainer* t, float* loss, float* correct);
int tp_trainer_pending(tp_trainer* t, size_t* count);
/* Trainer::evaluate body (src/train.rs:156-166) on one host batch */
int tp_trainer_eval(tp_trainer* t, const float* images, const float* labels, size_t batch, const size_t* sample_shape,
                    int ndim, float* loss, float* correct);
int tp_trainer_save_checkpoint(tp_trainer* t, const char* path);     /* src/train.rs:264-292 */
int tp_trainer_load_checkpoint(tp_trainer* t, const char* path);
/* data-parallel replica: NCCL rank binding, parameter broadcast; gradients are allreduced inside every step */
int tp_trainer_comm_init(tp_trainer* t, int rank, int world, const void* unique_id128);
int tp_trainer_broadcast_params(tp_trainer* t, int root);
/* NVLink peer-memory gradient exchange for the fused device step (tp_xchg_* in taper_b200.h), after tp_trainer_comm_init:
 * export this rank's 64-byte window handle, all-gather the handles in rank order with the launcher's transport, connect. */
int tp_trainer_peer_handle(tp_trainer* t, void* out64);
int tp_trainer_peer_connect(tp_trainer* t, const void* handles_world_x_64);
int tp_trainer_graph_replays(tp_trainer* t, uint64_t* count);

/* ---- the data path and the epoch-level trainer calls (src/data/mnist.rs, src/train.rs:98-261) ------------------------------
 * tp_dataset_from_arrays : samples [n, cols] as f32 in [0,1] or (is_u8) as the raw u8 pixels, labels f32 [n]
 * tp_dataset_load_mnist  : the IDX files `dir`/{train,test}_{images,labels} (src/data/mnist.rs:184-273; no download)
 * tp_loader_create       : DataLoader::new(dataset, batch_size, shuffle) (:335-348); the loader shares the dataset.  Batches are
 *                          gathered by worker threads into pinned host buffers ahead of the trainer (get_batch, :276-309).
 * tp_loader_set_sample_shape : shape of one sample as the model wants it (default [cols]; a CNN takes [1,28,28]) */
int tp_dataset_from_arrays(const void* images, int is_u8, const float* labels, size_t n, size_t cols, tp_dataset** out);
int tp_dataset_load_mnist(const char* dir, int train, tp_dataset** out);
int tp_dataset_len(tp_dataset* d, size_t* n);
int tp_dataset_destroy(tp_dataset* d);
int tp_loader_create(tp_dataset* d, size_t batch_size, int shuffle, uint64_t seed, tp_loader** out);
int tp_loader_set_sample_shape(tp_loader* l, const size_t* sample_shape, int ndim);
int tp_loader_num_batches(tp_loader* l, size_t* count);                                  /* :360-362 */
int tp_loader_destroy(tp_loader* l);
/* Trainer::train_epoch (src/train.rs:98-144): loader.reset(), then one step per batch; returns (sum of batch losses /
 * num_batches, correct / samples).  max_batches > 0 stops the epoch early (benchmarks).  Raw u8 pixels cross PCIe when the
 * dataset has them and the model's step is the wide plan; otherwise f32, as the reference's loader produces. */
int tp_trainer_train_epoch(tp_trainer* t, tp_loader* l, size_t max_batches, float* loss, float* acc);
int tp_trainer_evaluate(tp_trainer* t, tp_loader* l, float* loss, float* acc);          /* src/train.rs:147-172 */
/* LR schedulers (src/optim.rs:190-352): kind "step" (p1 = gamma, n = step_size) | "exponential" (p1 = gamma) |
 * "cosine" (p1 = min_lr, n = t_max) | "plateau" (p1 = factor, p2 = min_lr, n = patience, mode "min" | "max") */
int tp_scheduler_create(const char* kind, float base_lr, float p1, float p2, size_t n, const char* mode, tp_scheduler** out);
int tp_scheduler_step(tp_scheduler* s, int has_metric, float metric);
int tp_scheduler_get_lr(tp_scheduler* s, float* lr);
int tp_scheduler_destroy(tp_scheduler* s);
int tp_trainer_set_scheduler(tp_trainer* t, tp_scheduler* s);                            /* Trainer::new(.., scheduler) :83-95 */
/* Trainer::fit (src/train.rs:175-261): per epoch train_epoch, evaluate, scheduler.step(val_loss) + optimizer.set_lr, metrics;
 * stops early above 99 % validation accuracy (:246-249) */
int tp_trainer_fit(tp_trainer* t, tp_loader* train_loader, tp_loader* val_loader, size_t epochs, int verbose);
/* Metrics (src/train.rs:10-71): which = 0 train_loss, 1 train_acc, 2 val_loss, 3 val_acc, 4 epoch_times */
int tp_trainer_metrics(tp_trainer* t, int which, float* out, size_t cap, size_t* count);
int tp_trainer_get_lr(tp_trainer* t, float* lr);
/* sticky device-side error word of the trainer's context (0 = none): 1 label outside [0, classes), 2 grid barrier timeout,
 * 3 peer exchange timeout.  fetch / train_epoch raise on a non-zero word; this reads it without raising. */
int tp_trainer_device_error(tp_trainer* t, int* code);

/* ---- Tensor / Tape / Module / loss / optimizer at the granularity of the reference's own API ----------------------------
 * What a Rust (or any FFI) shim binds so that examples/train_mnist.rs-style code runs unmodified: every call records the
 * same tape node the reference records (src/tape.rs:51-101); tensors live on the device, tp_tensor_data reads them back.
 *   tp_tensor_new        Tensor::new(data, shape) [.requires_grad()]                 src/tensor.rs:470-487
 *   tp_tensor_unary      op = relu | exp | log | sigmoid | mean | transpose | pow | sqrt   (arg = exponent for pow)
 *   tp_tensor_binary     op = add | sub | mul | div | matmul | add_broadcast | sub_broadcast_rows
 *   tp_tensor_sum/argmax dim < 0: over all elements                                  src/tensor.rs:890-1088
 *   tp_module_forward    Module::forward on a device tensor (src/nn.rs:10-12); tp_model_parameter = parameters()[i]
 *   tp_loss              kind = cross_entropy | cross_entropy_onehot | bce | mse     src/loss.rs:6-245
 *   tp_optimizer_*       SGD / Adam / AdamW over tensor handles                      src/optim.rs:8-181 */
int tp_tensor_new(const float* data, const size_t* shape, int ndim, int requires_grad, tp_tensor** out);
int tp_tensor_clone(tp_tensor* t, tp_tensor** out);
int tp_tensor_free(tp_tensor* t);
int tp_tensor_ndim(tp_tensor* t, int* ndim);
int tp_tensor_shape(tp_tensor* t, size_t* dims, int cap);
int tp_tensor_numel(tp_tensor* t, size_t* n);
int tp_tensor_data(tp_tensor* t, float* out, size_t n);                                  /* Tensor::data      :493-496 */
int tp_tensor_set_data(tp_tensor* t, const float* data, size_t n);                       /* data_mut          :499-501 */
int tp_tensor_grad(tp_tensor* t, float* out, size_t n, int* has_grad);                   /* Tensor::grad      :512-518 */
int tp_tensor_requires_grad(tp_tensor* t, int* flag);
int tp_tensor_zero_grad(tp_tensor* t);                                                   /* :531-533 */
int tp_tensor_backward(tp_tensor* t);                                                    /* :520-529 */
int tp_tensor_unary(const char* op, tp_tensor* x, float arg, tp_tensor** out);
int tp_tensor_binary(const char* op, tp_tensor* a, tp_tensor* b, tp_tensor** out);
int tp_tensor_reshape(tp_tensor* x, const size_t* shape, int ndim, tp_tensor** out);     /* :803-840 */
int tp_tensor_flatten(tp_tensor* x, size_t start_dim, tp_tensor** out);                  /* :842-858 */
int tp_tensor_sum(tp_tensor* x, int dim, int keepdim, tp_tensor** out);
int tp_tensor_argmax(tp_tensor* x, int dim, tp_tensor** out);
int tp_tape_reset(void);                                                                 /* Tape::reset       src/tape.rs:43-49 */
int tp_tape_len(size_t* nodes);
int tp_module_forward(tp_model* m, tp_tensor* x, tp_tensor** out);
int tp_model_parameter(tp_model* m, int index, tp_tensor** out);
int tp_loss(const char* kind, tp_tensor* predictions, tp_tensor* targets, tp_tensor** out);
int tp_accuracy(tp_tensor* predictions, tp_tensor* targets, float* acc);                 /* src/loss.rs:271-290 */
int tp_optimizer_create(const char* kind, tp_tensor* const* params, int n_params, float lr, float beta1, float beta2, float eps,
                        float weight_decay, tp_optimizer** out);
int tp_optimizer_step(tp_optimizer* o);
int tp_optimizer_zero_grad(tp_optimizer* o);
int tp_optimizer_set_lr(tp_optimizer* o, float lr);
int tp_optimizer_get_lr(tp_optimizer* o, float* lr);
int tp_optimizer_destroy(tp_optimizer* o);

#ifdef __cplusplus
}
#endif
#endif /* TAPER_B200_HOST_H */

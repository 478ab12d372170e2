// taper.hpp — C++ host layer mirroring taper's Rust API one-to-one (Tensor / Tape / nn::Module /
// loss / optim / train), sitting above the C ABI in include/taper_b200.h exactly as a Rust shim
// would.  The reference's toolchain (cargo/rustc) is not available in the build image, so the host
// side is C++; names, argument meaning and error behaviour (panic -> std::runtime_error) follow the
// reference.  Citations are `path:line` in vaibhawvipul/taper @ aea74b46.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <memory>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "taper_b200.h"

namespace taper {

using Shape = std::vector<size_t>;
using Pair = std::pair<size_t, size_t>;

// ---- device context (one per host thread, like the reference's thread-local tape) -------------
tp_ctx* ctx();
void set_device(int device);            // must be called before the first tensor is created on this thread
void synchronize();
void check(int rc);                     // non-zero status -> throw (the reference panics)

// Switches documented in DESIGN.md
struct Config {
    // the reference drops the tape links inside conv2d, so conv weights and
    // conv inputs never receive gradients.  false (default) reproduces the reference bit-for-bit in
    // structure; true computes the full adjoint (dW = col^T dY, dX = col2im(dY W^T)).
    static bool& conv_full_adjoint();
    // Sequential peephole: Linear followed by ReLU runs as one fused launch (bias+ReLU in the GEMM
    // epilogue, ReLU mask folded into the backward).  Numerically identical to the unfused ops.
    static bool& fuse_linear_relu();
    // true: nn::Linear records the reference's literal op sequence transpose -> matmul -> add_broadcast
    // (src/nn.rs:54-60, three tape nodes) instead of the single fused node.  Same numbers, more launches.
    static bool& reference_op_sequence();
    // Sequential peephole: a run of >= 2 [Conv2d / Conv2dReLU (3x3, stride 1, pad 1)] (+ MaxPool2d 2x2) layers whose input
    // carries no gradient runs as one stack on the tensor cores (tp_conv_stack_fwd: NHWC bf16 hi/lo activations between the
    // layers, pooling in the conv epilogue).  Only under the strict-reference conv autograd (A1), where nothing but the
    // stack's last output is ever read again; off under reference_op_sequence.  Same numbers within the bf16x3 bound.
    static bool& fuse_conv_stack();
    // Sequential peephole: a run of >= 2 Linear(+ReLU) layers whose widths are all <= 128 (a classifier head) runs as one
    // forward and one backward launch of exact fp32 FFMA kernels (tp_mlp_small_*) instead of one tensor-core launch per product.
    static bool& fuse_small_mlp();
};

struct TensorImpl;

struct ConvStackLayer;
struct MlpLayer;

// ---- Tensor  (src/tensor.rs:236-244, 469-541) -----------------------------------------------------
class Tensor {
public:
    Tensor() = default;
    static Tensor create(const std::vector<float>& data, const Shape& shape);   // Tensor::new  :470-478
    static Tensor from_host(const float* data, const Shape& shape);
    static Tensor scalar(float v);                                              // :480-482
    static Tensor zeros(const Shape& shape);
    static Tensor empty(const Shape& shape);                                    // uninitialised device buffer
    static Tensor randn(const Shape& shape, uint64_t seed);                     // src/ops.rs:301-309 (seeded here)
    static Tensor adopt(tp_buf* buf, const Shape& shape);                       // takes ownership of buf

    Tensor requires_grad() const;                 // `fn requires_grad(mut self) -> Self`  :484-487
    bool needs_grad() const { return requires_grad_; }
    void set_requires_grad(bool v) { requires_grad_ = v; }

    const Shape& shape() const;                   // :489-491
    size_t numel() const;
    bool defined() const { return (bool)impl_; }
    const std::vector<float>& data() const;       // :493-496 — lazily synchronised host mirror (syncs the stream)
    float item() const;                           // data()[0]
    void set_data(const std::vector<float>& v) const;   // data_mut() :499-501 — writes through to the device
    std::optional<Tensor> grad() const;           // :512-518 (clone of the gradient, or None)
    void set_grad(const std::vector<float>& g) const;   // `grad` is a pub field, src/tensor.rs:241
    void zero_grad() const;                       // :531-533
    void backward() const;                        // :520-529

    tp_buf* buf() const;                          // device buffer (borrowed)
    tp_buf* grad_buf() const;                     // device gradient buffer or NULL
    std::shared_ptr<TensorImpl> impl() const { return impl_; }

    // ops — each records a tape node exactly when the reference does
    Tensor matmul(const Tensor& other) const;                     // src/ops.rs:200-298
    Tensor relu() const;                                          // src/ops.rs:312-374
    Tensor transpose() const;                                     // src/tensor.rs:544-591
    Tensor add_broadcast(const Tensor& other) const;              // src/tensor.rs:636-704
    Tensor sub_broadcast_rows(const Tensor& other) const;         // src/tensor.rs:707-770
    Tensor reshape(const Shape& shape) const;                     // src/tensor.rs:803-840 (copies, A11)
    Tensor flatten(size_t start_dim) const;                       // src/tensor.rs:842-858
    Tensor view(const Shape& shape) const { return reshape(shape); }   // src/tensor.rs:1214
    Tensor sum(std::optional<size_t> dim, bool keepdim) const;    // src/tensor.rs:890-1018
    std::pair<Tensor, Tensor> max(std::optional<size_t> dim) const;   // src/tensor.rs:1021-1083
    Tensor argmax(std::optional<size_t> dim) const;               // src/tensor.rs:1086-1088
    Tensor sigmoid() const;                                       // src/tensor.rs:594-634
    Tensor mean() const;                                          // src/tensor.rs:772-800
    Tensor pow(float exponent) const;                             // src/tensor.rs:1172-1206
    Tensor sqrt() const { return pow(0.5f); }                     // src/tensor.rs:1209-1211
    Tensor exp() const;                                           // src/tensor.rs:1091-1133
    Tensor log() const;                                           // src/tensor.rs:1136-1169
    Tensor conv2d(const Tensor& weight, const Tensor* bias, Pair stride, Pair padding, Pair dilation) const;       // :1221-1285
    Tensor conv2d_relu(const Tensor& weight, const Tensor* bias, Pair stride, Pair padding, Pair dilation) const;  // :1379-1389
    Tensor max_pool2d(Pair kernel, std::optional<Pair> stride, Pair padding) const;   // :1391-1521
    Tensor avg_pool2d(Pair kernel, std::optional<Pair> stride, Pair padding) const;   // :1524-1660

    // fused Linear (+ReLU): one node standing for transpose+matmul+add_broadcast(+relu) of src/nn.rs:54-60
    Tensor linear(const Tensor& weight, const Tensor* bias, bool relu) const;
    // a stack of 3x3 / s1 / p1 convolutions (+bias, +ReLU, + 2x2 max-pool) as one fused forward (Config::fuse_conv_stack);
    // records the one node the strict-reference tape can ever deliver through: the last layer's bias gradient.
    // Returns an undefined Tensor when the shapes are outside the fused kernels (the caller runs the layers one by one).
    // gap: 0 = the stack's NCHW output; 1 = followed by a global average pool -> [N, C, 1, 1]; 2 = and Flatten(1) -> [N, C]
    Tensor conv_stack(const std::vector<ConvStackLayer>& layers, int gap = 0) const;
    // a chain of small Linear(+ReLU) layers as one node (Config::fuse_small_mlp); undefined Tensor when the widths do not qualify
    Tensor mlp_chain(const std::vector<MlpLayer>& layers) const;

private:
    Tensor conv2d_impl(const Tensor& weight, const Tensor* bias, Pair stride, Pair padding, Pair dilation, bool relu) const;
    std::shared_ptr<TensorImpl> impl_;
    bool requires_grad_ = false;
    friend struct TensorImpl;
};

struct ConvStackLayer {           // one layer of Tensor::conv_stack: conv 3x3 / s1 / p1 + bias (+ ReLU) (+ 2x2 / s2 max-pool)
    Tensor weight;
    std::optional<Tensor> bias;
    bool relu = true, pool = false;
};

struct MlpLayer {                 // one layer of Tensor::mlp_chain: x . W^T + b (+ ReLU)
    Tensor weight;
    std::optional<Tensor> bias;
    bool relu = false;
};

Tensor operator+(const Tensor& a, const Tensor& b);     // src/ops.rs:8-52
Tensor operator-(const Tensor& a, const Tensor& b);     // src/ops.rs:377-420
Tensor operator*(const Tensor& a, const Tensor& b);     // src/ops.rs:54-120
Tensor operator/(const Tensor& a, const Tensor& b);     // src/ops.rs:440-496

// ---- Tape  (src/tape.rs) ------------------------------------------------------------------------------
class Tape {
public:
    static void ensure_active();                                                            // :34-40
    static void reset();                                                                    // :43-49
    static void push_binary_op(const Tensor& a, const Tensor& b, const Tensor& out, std::function<void()> fn);   // :51-76
    static void push_unary_op(const Tensor& input, const Tensor& out, std::function<void()> fn);                 // :78-101
    static size_t len();
    // moves the recorded closures out (a captured training step keeps them, and the tensors they
    // own, alive for as long as its CUDA graph can be replayed)
    static std::vector<std::function<void()>> take();
};
void backward(size_t final_node_id);                                                        // src/tape.rs:106-127

// ---- nn  (src/nn.rs, src/activation.rs) -------------------------------------------------------------------
namespace nn {

struct Module {
    virtual ~Module() = default;
    virtual Tensor forward(const Tensor& input) const = 0;          // src/nn.rs:10-12
    virtual std::vector<Tensor> parameters() const { return {}; }
    virtual const char* kind() const { return "module"; }
    // false: forward() draws fresh host-side state every call (Dropout's mask), so a step containing it must not be
    // captured into a CUDA graph and replayed
    virtual bool capturable() const { return true; }
};

struct Linear : Module {                                             // src/nn.rs:28-78
    Tensor weight;                   // [out, in]
    std::optional<Tensor> bias;      // [out]
    Linear(size_t in_features, size_t out_features, bool with_bias, uint64_t seed = 0);
    Tensor forward(const Tensor& input) const override;
    Tensor forward_fused_relu(const Tensor& input) const;
    std::vector<Tensor> parameters() const override;
    const char* kind() const override { return "linear"; }
};

struct ReLU : Module {                                               // src/activation.rs:7-21
    Tensor forward(const Tensor& input) const override { return input.relu(); }
    const char* kind() const override { return "relu"; }
};

struct Sigmoid : Module {                                            // src/activation.rs:37-51
    Tensor forward(const Tensor& input) const override { return input.sigmoid(); }
    const char* kind() const override { return "sigmoid"; }
};

struct Dropout : Module {                                            // src/nn.rs:775-827 (mask from a seeded generator here)
    float p;
    bool training = true;
    mutable uint64_t rng_state;
    explicit Dropout(float p_, uint64_t seed = 0);
    void eval() { training = false; }
    void train() { training = true; }
    Tensor forward(const Tensor& input) const override;
    const char* kind() const override { return "dropout"; }
    bool capturable() const override { return !training || p == 0.0f || p == 1.0f; }
};

struct Sequential : Module {                                         // src/nn.rs:130-162
    std::vector<std::shared_ptr<Module>> layers;
    Sequential() = default;
    explicit Sequential(std::vector<std::shared_ptr<Module>> l) : layers(std::move(l)) {}
    Tensor forward(const Tensor& input) const override;
    std::vector<Tensor> parameters() const override;
    const char* kind() const override { return "sequential"; }
    bool capturable() const override {
        for (auto& l : layers) if (!l->capturable()) return false;
        return true;
    }
};

struct Conv2d : Module {                                             // src/nn.rs:180-354 (groups == 1)
    Tensor weight;                   // [C_out, C_in, kh, kw]
    std::optional<Tensor> bias;
    Pair stride, padding, dilation;
    size_t groups;
    Conv2d(size_t in_channels, size_t out_channels, Pair kernel_size, std::optional<Pair> stride,
           std::optional<Pair> padding, std::optional<Pair> dilation, std::optional<size_t> groups, bool bias,
           uint64_t seed = 0);
    Tensor forward(const Tensor& input) const override;
    std::vector<Tensor> parameters() const override;
    const char* kind() const override { return "conv2d"; }
};

struct Conv2dReLU : Conv2d {                                         // src/nn.rs:433-490
    using Conv2d::Conv2d;
    Tensor forward(const Tensor& input) const override;
    const char* kind() const override { return "conv2d_relu"; }
};

struct MaxPool2d : Module {                                          // src/nn.rs:508-549
    Pair kernel_size;
    std::optional<Pair> stride;
    Pair padding;
    MaxPool2d(Pair k, std::optional<Pair> s, std::optional<Pair> p) : kernel_size(k), stride(s), padding(p.value_or(Pair{0, 0})) {}
    Tensor forward(const Tensor& input) const override { return input.max_pool2d(kernel_size, stride, padding); }
    const char* kind() const override { return "maxpool2d"; }
};

struct AvgPool2d : Module {                                          // src/nn.rs:570-653
    Pair kernel_size;
    std::optional<Pair> stride;
    Pair padding;
    AvgPool2d(Pair k, std::optional<Pair> s, std::optional<Pair> p) : kernel_size(k), stride(s), padding(p.value_or(Pair{0, 0})) {}
    Tensor forward(const Tensor& input) const override { return input.avg_pool2d(kernel_size, stride, padding); }
    const char* kind() const override { return "avgpool2d"; }
};

struct AdaptiveAvgPool2d : Module {                                  // src/nn.rs:655-697
    Pair output_size;
    explicit AdaptiveAvgPool2d(Pair out) : output_size(out) {}
    static AdaptiveAvgPool2d global() { return AdaptiveAvgPool2d({1, 1}); }
    Tensor forward(const Tensor& input) const override;
    const char* kind() const override { return "adaptive_avgpool2d"; }
};

struct Flatten : Module {                                            // src/nn.rs:730-756
    size_t start_dim;
    explicit Flatten(std::optional<size_t> sd = std::nullopt) : start_dim(sd.value_or(1)) {}
    Tensor forward(const Tensor& input) const override { return input.flatten(start_dim); }
    const char* kind() const override { return "flatten"; }
};

}  // namespace nn

// ---- loss  (src/loss.rs) -----------------------------------------------------------------------------------
namespace loss {
Tensor softmax(const Tensor& x, int dim);                            // :82-98 (row-wise intent, A13)
Tensor log_softmax(const Tensor& x, int dim);                        // :101-126 — composed op-for-op
Tensor cross_entropy_loss(const Tensor& logits, const Tensor& targets);   // :136-195 — fused fwd, direct bwd
float accuracy(const Tensor& predictions, const Tensor& targets);    // :271-290
Tensor bce_loss(const Tensor& predictions, const Tensor& targets);   // :6-72
Tensor mse_loss(const Tensor& predictions, const Tensor& targets);   // :75-80  ((p - t) * (p - t)).mean(), composed as in the reference
Tensor cross_entropy_loss_onehot(const Tensor& logits, const Tensor& targets);   // :202-245
Tensor one_hot(const Tensor& indices, size_t num_classes);           // :248-268
// device-side variant: correct count as a [1] tensor, no host sync (used by the captured train step)
Tensor accuracy_count(const Tensor& predictions, const Tensor& targets);
// variants writing into a caller-provided [1] tensor (the trainer's result slot)
Tensor cross_entropy_loss_into(const Tensor& logits, const Tensor& targets, const Tensor& out);
void accuracy_count_into(const Tensor& predictions, const Tensor& targets, const Tensor& out);
// cross_entropy_loss + accuracy's correct count in one launch (the head of a training step, src/train.rs:112-115)
Tensor cross_entropy_with_accuracy_into(const Tensor& logits, const Tensor& targets, const Tensor& out, const Tensor* correct_out);
}  // namespace loss

// ---- optim  (src/optim.rs) ----------------------------------------------------------------------------------
namespace optim {

// Flat parameter / gradient / moment arenas shared by the optimizers: params are re-homed into one
// contiguous buffer so the step is one launch and the data-parallel exchange is one allreduce.
struct Arena;

struct Optimizer {                                                   // :3-6
    virtual ~Optimizer() = default;
    virtual void step() = 0;
    virtual void zero_grad() = 0;
    // data-parallel hooks (no counterpart in the reference): the flat gradient arena to allreduce and
    // the 1/world factor the step kernel folds in
    virtual std::shared_ptr<Arena> arena() const = 0;
    virtual void set_grad_scale(float s) = 0;
    virtual void set_lr(float) {}
    // a replayed CUDA graph updated the parameters without the host-side step() running: invalidate host mirrors
    void mark_parameters_updated() const;
    // hooks of the fused device step (tp_step_*, the whole train_epoch loop body as one persistent kernel)
    virtual int kind() const = 0;                 // tp_step_desc.optimizer: 0 SGD, 1 Adam, 2 AdamW
    virtual float lr() const = 0;
    virtual float grad_scale() const = 0;
    virtual void note_device_step() = 0;          // the device ran one step(): host-side counters and mirrors follow
};
// Describe `model` + `opt` as a tp_step_desc if the model is a chain Linear[,ReLU]...Linear whose parameters are exactly
// the optimizer's arena; false otherwise.  bufs receives {params, grads, m, v, hyper} (m, v, hyper NULL for SGD).
bool describe_fused_step(const nn::Module& model, const Optimizer& opt, size_t batch, tp_step_desc* desc, tp_buf* bufs[5]);
tp_buf* arena_grad_buf(const std::shared_ptr<Arena>& a);
tp_buf* arena_param_buf(const std::shared_ptr<Arena>& a);
size_t arena_total(const std::shared_ptr<Arena>& a);

class SGD : public Optimizer {                                       // :8-40 (momentum ignored, :14-17)
public:
    SGD(std::vector<Tensor> params, float lr, std::optional<float> momentum = std::nullopt);
    ~SGD() override;
    void step() override;
    void zero_grad() override;
    void set_grad_scale(float s) override { grad_scale_ = s; }
    std::shared_ptr<Arena> arena() const override { return arena_; }
    void set_lr(float lr) override;               // also updates the device-side copy a captured step reads
    int kind() const override { return 0; }
    float lr() const override { return lr_; }
    float grad_scale() const override { return grad_scale_; }
    void note_device_step() override { mark_parameters_updated(); }
private:
    std::vector<Tensor> params_;
    float lr_;
    float grad_scale_ = 1.0f;
    std::shared_ptr<Arena> arena_;
    tp_buf* lr_dev_ = nullptr;
};

class Adam : public Optimizer {                                      // :43-128
public:
    Adam(std::vector<Tensor> params, float lr, std::optional<std::pair<float, float>> betas = std::nullopt,
         std::optional<float> eps = std::nullopt, std::optional<float> weight_decay = std::nullopt);
    ~Adam() override;
    void step() override;                                            // :83-113
    void zero_grad() override;                                       // :115-119
    float get_lr() const { return lr_; }                             // :121-123
    void set_lr(float lr) override;                                  // :125-127 (also updates the device-side state)
    void set_grad_scale(float s) override { grad_scale_ = s; }       // 1/world for data-parallel averaging
    std::shared_ptr<Arena> arena() const override { return arena_; }
    size_t t() const { return t_; }
    int kind() const override { return 1; }
    float lr() const override { return lr_; }
    float grad_scale() const override { return grad_scale_; }
    void note_device_step() override { t_ += 1; mark_parameters_updated(); }
protected:
    void step_impl(bool decoupled);
    std::vector<Tensor> params_;
    float lr_, beta1_, beta2_, eps_, weight_decay_;
    float grad_scale_ = 1.0f;
    size_t t_ = 0;
    std::shared_ptr<Arena> arena_;
    friend class AdamW;
};

class AdamW : public Optimizer {                                     // :131-181
public:
    AdamW(std::vector<Tensor> params, float lr, std::optional<std::pair<float, float>> betas = std::nullopt,
          std::optional<float> eps = std::nullopt, std::optional<float> weight_decay = std::nullopt);
    void step() override;                                            // :148-168
    void zero_grad() override { adam_.zero_grad(); }
    float get_lr() const { return adam_.get_lr(); }
    void set_lr(float lr) override { adam_.set_lr(lr); }
    void set_grad_scale(float s) override { adam_.set_grad_scale(s); }
    std::shared_ptr<Arena> arena() const override { return adam_.arena(); }
    int kind() const override { return 2; }
    float lr() const override { return adam_.lr(); }
    float grad_scale() const override { return adam_.grad_scale(); }
    void note_device_step() override { adam_.note_device_step(); }
private:
    Adam adam_;
};

struct LRScheduler {                                                 // :184-187
    virtual ~LRScheduler() = default;
    virtual void step(std::optional<float> metrics) = 0;
    virtual float get_lr() const = 0;
};
struct StepLR : LRScheduler {                                        // :190-219
    float current_lr, gamma; size_t step_size, current_epoch = 0;
    StepLR(float base_lr, size_t step_size_, float gamma_) : current_lr(base_lr), gamma(gamma_), step_size(step_size_) {}
    void step(std::optional<float>) override;
    float get_lr() const override { return current_lr; }
};
struct ExponentialLR : LRScheduler {                                 // :221-244
    float current_lr, gamma;
    ExponentialLR(float base_lr, float gamma_) : current_lr(base_lr), gamma(gamma_) {}
    void step(std::optional<float>) override { current_lr *= gamma; }
    float get_lr() const override { return current_lr; }
};
struct CosineAnnealingLR : LRScheduler {                             // :246-285
    float base_lr, min_lr, current_lr; size_t t_max, current_epoch = 0;
    CosineAnnealingLR(float base, size_t tmax, std::optional<float> minlr)
        : base_lr(base), min_lr(minlr.value_or(0.0f)), current_lr(base), t_max(tmax) {}
    void step(std::optional<float>) override;
    float get_lr() const override { return current_lr; }
};
struct ReduceLROnPlateau : LRScheduler {                             // :287-352
    float current_lr, factor, min_lr, best_metric; size_t patience, patience_counter = 0; std::string mode;
    ReduceLROnPlateau(float initial_lr, float factor_, size_t patience_, std::optional<float> min_lr_, std::optional<std::string> mode_);
    void step(std::optional<float> metrics) override;
    float get_lr() const override { return current_lr; }
};

}  // namespace optim

// ---- data-parallel group (no counterpart in the reference) -----------------------------------------------------
namespace dist {
void init(int rank, int world, const void* nccl_unique_id128);       // binds NCCL to this thread's context
int rank();
int world();
void allreduce_sum(tp_buf* buf, size_t n);
void broadcast(tp_buf* buf, size_t n, int root);
}  // namespace dist

// ---- data  (src/data/mnist.rs) ------------------------------------------------------------------------------------
namespace data {

struct MNISTDataset {                                                // :21-25
    std::vector<float> images;       // [N, cols] in [0, 1]  (src/data/mnist.rs:225: u8 / 255); may be empty when only the raw
                                     // pixels are kept (from_arrays with u8 input)
    std::vector<uint8_t> images_u8;  // [N, cols] the pixels as stored on disk; dropped by normalize()
    std::vector<float> labels;       // [N]
    size_t cols = 784;
    bool train = true;
    MNISTDataset() = default;
    MNISTDataset(bool train, const std::string& data_dir = "./data/mnist");     // :29-58 (IDX files must exist; no download)
    static MNISTDataset synthetic(size_t n, uint64_t seed);          // MNIST-shaped U[0,1) images, uniform labels
    // caller-provided samples: f32 [n, cols] or u8 [n, cols] (is_u8), labels f32 [n]
    static MNISTDataset from_arrays(const void* images, bool is_u8, const float* labels, size_t n, size_t cols);
    size_t len() const { return labels.size(); }                     // :312-314
    void normalize(float mean, float std);                           // :317-322
    bool has_u8() const { return !images_u8.empty(); }
    void ensure_f32();                                               // materialise `images` from the raw pixels if needed
};

class DataLoader {                                                   // :326-385
public:
    DataLoader(MNISTDataset dataset, size_t batch_size, bool shuffle, uint64_t seed = 0);
    DataLoader(std::shared_ptr<MNISTDataset> dataset, size_t batch_size, bool shuffle, uint64_t seed = 0);
    ~DataLoader();
    DataLoader(const DataLoader&) = delete;
    DataLoader& operator=(const DataLoader&) = delete;
    void reset();                                                    // :350-358 (reshuffles)
    size_t num_batches() const { return (dataset_->len() + batch_size_ - 1) / batch_size_; }   // :360-362
    // Iterator::next (:369-384): false at the end of the epoch.  The batch is gathered into host vectors
    // (last batch partial, :377).
    bool next(std::vector<float>& images, std::vector<float>& labels, size_t& batch);
    const MNISTDataset& dataset() const { return *dataset_; }
    const std::vector<uint32_t>& indices() const { return indices_; }
    size_t batch_size() const { return batch_size_; }
    // shape of one sample as the model wants it (default {cols}; a CNN takes {1, 28, 28}, examples/train_mnist_cnn.rs:162)
    Shape sample_shape;

    // The same iteration as a pipeline: worker threads gather the batches of the epoch (get_batch, :276-309, is a rayon
    // row gather in the reference) into a ring of PINNED host buffers, in order, ahead of the consumer, so the trainer's
    // H2D copies are asynchronous.  u8: deliver the raw pixels (the device divides by 255) instead of f32.
    struct Batch { const void* images; const float* labels; size_t batch; int slot; bool u8; };
    void start_prefetch(bool u8, size_t max_batches = 0);            // after reset(); max_batches 0 = the whole epoch
    bool next_pinned(Batch& b);                                      // false at the end of the epoch
    void release(int slot);                                          // the consumer is done with the slot's memory
    void stop_prefetch();
private:
    struct Pipe;
    std::shared_ptr<MNISTDataset> dataset_;
    size_t batch_size_;
    bool shuffle_;
    std::vector<uint32_t> indices_;
    size_t current_ = 0;
    uint64_t rng_state_;
    std::unique_ptr<Pipe> pipe_;
};

}  // namespace data

// ---- train  (src/train.rs) -----------------------------------------------------------------------------------------
namespace train {

struct Metrics {                                                     // :10-71
    std::vector<float> train_loss, train_acc, val_loss, val_acc, epoch_times;
    void print_last() const;
    void plot_summary() const;
};

struct StepResult { float loss; float correct; };

class Trainer {                                                      // :73-293
public:
    std::shared_ptr<nn::Module> model;
    std::shared_ptr<optim::Optimizer> optimizer;                     // the reference hard-wires Adam (:76)
    std::shared_ptr<optim::LRScheduler> scheduler;
    Metrics metrics;
    Trainer(std::shared_ptr<nn::Module> m, std::shared_ptr<optim::Optimizer> o, std::shared_ptr<optim::LRScheduler> s = nullptr);
    ~Trainer();

    // One iteration of the train_epoch loop body (:106-138): Tape::reset, forward, cross-entropy, accuracy,
    // backward, [gradient allreduce], optimizer.step, zero_grad.  After one eager iteration per batch
    // shape the whole body is captured into a CUDA graph and replayed.
    //   train_batch            host inputs (pageable), synchronous: returns (loss, #correct)
    //   train_batch_async      host inputs (pinned => true async H2D), result via fetch() in FIFO order
    //   load_dataset + train_batch_resident   inputs gathered on the device from a resident dataset
    StepResult train_batch(const float* images, const float* labels, size_t batch, const Shape& sample_shape);
    void train_batch_async(const float* images, const float* labels, size_t batch, const Shape& sample_shape, bool pinned);
    void load_dataset(const float* images, const float* labels, size_t n, const Shape& sample_shape, const uint32_t* perm);
    // the same with u8 pixels, MNIST's on-disk format (src/data/mnist.rs:225 divides by 255 at load time; here the wide step
    // plan does it on the device, so a quarter of the bytes cross PCIe / sit in HBM)
    void train_batch_async_u8(const uint8_t* images, const float* labels, size_t batch, const Shape& sample_shape, bool pinned);
    void load_dataset_u8(const uint8_t* images, const float* labels, size_t n, const Shape& sample_shape, const uint32_t* perm);
    void train_batch_resident(size_t batch);
    StepResult fetch();
    size_t pending() const;
    StepResult eval_batch(const float* images, const float* labels, size_t batch, const Shape& sample_shape);   // :156-166 body

    std::pair<float, float> train_epoch(data::DataLoader& loader, size_t max_batches = 0);  // :98-144 (max_batches 0 = all)
    std::pair<float, float> evaluate(data::DataLoader& loader);     // :147-172
    void fit(data::DataLoader& train_loader, data::DataLoader& val_loader, size_t epochs, bool verbose);   // :175-261

    void save_checkpoint(const std::string& path) const;             // :264-292 text format
    void load_checkpoint(const std::string& path) const;             // loader for the same format (the reference has none)

    // data-parallel: bind this thread's context to a NCCL rank; gradients are summed across ranks after
    // backward and the optimizer folds 1/world.  broadcast_parameters makes every replica start equal.
    void init_data_parallel(int rank, int world, const void* nccl_unique_id128);
    void broadcast_parameters(int root);
    // NVLink peer-memory gradient exchange for the fused device step (tp_xchg_*): each rank exports the 64-byte IPC
    // handle of its window, the launcher all-gathers them (rank order), connect maps the peers.  Without it a
    // data-parallel trainer uses the tape + CUDA-graph path with the NCCL allreduce.
    void peer_exchange_handle(void* out64);
    void peer_exchange_connect(const void* handles_world_x_64);

    void set_use_graph(bool v) { use_graph_ = v; }
    uint64_t graph_replays() const { return graph_replays_; }
    // fused device step (one persistent kernel per training step, tp_step_*): used whenever the model / optimizer /
    // batch qualify (describe_fused_step + tp_step_supported); everything else takes the tape + CUDA-graph path
    void set_use_fused(bool v) { use_fused_ = v; }
    uint64_t fused_steps() const { return fused_steps_; }
    int fused_kind() const;                                          // 2 wide tcgen05 plan, 1 persistent kernel, 0 none run yet
private:
    struct Impl;
    void train_batch_async_impl(const void* images, const float* labels, size_t batch, const Shape& sample_shape, bool pinned, bool u8);
    void load_dataset_impl(const void* images, const float* labels, size_t n, const Shape& sample_shape, const uint32_t* perm, bool u8);
    std::unique_ptr<Impl> p_;
    bool use_graph_ = true;
    uint64_t graph_replays_ = 0;
    bool use_fused_ = true;
    uint64_t fused_steps_ = 0;
};

}  // namespace train

}  // namespace taper

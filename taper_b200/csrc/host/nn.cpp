// Host layer, part 2: nn modules, losses, optimizers, LR schedulers, the data-parallel group.
// Mirrors src/nn.rs, src/activation.rs, src/loss.rs and src/optim.rs of the reference.
#include <cstring>
#include "taper_internal.hpp"

#include <cmath>
#include <random>

namespace taper {

// =====================================================================================================
// nn  (src/nn.rs)
// =====================================================================================================
namespace nn {

static std::vector<float> uniform_init(size_t n, float bound, uint64_t seed) {
    // Uniform::new_inclusive(-bound, bound) drawn from thread_rng in the reference (src/nn.rs:36-41);
    // seeded here so runs are reproducible.  Parity tests inject weights explicitly.
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<float> d(-bound, bound);
    std::vector<float> v(n);
    for (auto& x : v) x = d(rng);
    return v;
}

Linear::Linear(size_t in_features, size_t out_features, bool with_bias, uint64_t seed) {
    float scale = std::sqrt(2.0f / (float)in_features);                          // src/nn.rs:36
    weight = Tensor::create(uniform_init(in_features * out_features, scale, seed), {out_features, in_features}).requires_grad();
    if (with_bias) bias = Tensor::zeros({out_features}).requires_grad();         // src/nn.rs:46-47
}

Tensor Linear::forward(const Tensor& input) const {
    if (Config::reference_op_sequence()) {
        Tensor out = input.matmul(weight.transpose());                           // src/nn.rs:55
        if (bias) out = out.add_broadcast(*bias);                                // src/nn.rs:56-58
        return out;
    }
    return input.linear(weight, bias ? &*bias : nullptr, false);
}

Tensor Linear::forward_fused_relu(const Tensor& input) const {
    return input.linear(weight, bias ? &*bias : nullptr, true);
}

std::vector<Tensor> Linear::parameters() const {
    std::vector<Tensor> p{weight};
    if (bias) p.push_back(*bias);
    return p;
}

Dropout::Dropout(float p_, uint64_t seed) : p(p_), rng_state(seed ? seed : 0x9E3779B97F4A7C15ull) {
    if (!(p >= 0.0f && p <= 1.0f)) panic("Dropout probability must be between 0 and 1");
}

Tensor Dropout::forward(const Tensor& input) const {                             // src/nn.rs:799-822
    if (!training || p == 0.0f) return input;
    if (p == 1.0f) return Tensor::zeros(input.shape());
    // the reference draws the mask from thread_rng on the host (:808-816); a seeded xorshift keeps runs reproducible
    std::vector<float> mask(input.numel());
    const float scale = 1.0f / (1.0f - p);
    for (auto& m : mask) {
        rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
        float u = (float)(rng_state >> 40) * (1.0f / 16777216.0f);
        m = u > p ? scale : 0.0f;
    }
    return input * Tensor::create(mask, input.shape());
}

Tensor Sequential::forward(const Tensor& input) const {                         // src/nn.rs:149-151
    Tensor x = input;
    const bool fuse = Config::fuse_linear_relu() && !Config::reference_op_sequence();
    const bool fuse_conv = Config::fuse_conv_stack() && !Config::reference_op_sequence() && !Config::conv_full_adjoint();
    for (size_t i = 0; i < layers.size(); ++i) {
        if (fuse_conv && x.shape().size() == 4 && !x.needs_grad()) {
            // peephole: [Conv2d(ReLU) 3x3 s1 p1 (+ MaxPool2d 2x2 s2)]+  ->  one stack on the tensor cores
            std::vector<ConvStackLayer> run;
            size_t j = i;
            while (j < layers.size() && run.size() < 8) {
                auto* cv = dynamic_cast<const Conv2d*>(layers[j].get());
                if (!cv || cv->weight.shape()[2] != 3 || cv->weight.shape()[3] != 3 || cv->stride != Pair{1, 1} ||
                    cv->padding != Pair{1, 1} || cv->dilation != Pair{1, 1} || cv->groups != 1)
                    break;
                ConvStackLayer ly;
                ly.weight = cv->weight;
                ly.bias = cv->bias;
                ly.relu = dynamic_cast<const Conv2dReLU*>(cv) != nullptr;
                ++j;
                if (j < layers.size()) {
                    auto* mp = dynamic_cast<const MaxPool2d*>(layers[j].get());
                    if (mp && mp->kernel_size == Pair{2, 2} && mp->stride.value_or(mp->kernel_size) == Pair{2, 2} && mp->padding == Pair{0, 0}) {
                        ly.pool = true;
                        ++j;
                    }
                }
                run.push_back(ly);
            }
            if (run.size() >= 2) {
                // ... + AdaptiveAvgPool2d::global (+ Flatten(1)) ride along: pooled features and the count of positive units
                // per plane are all that forward and backward need of the last activation
                int gap = 0;
                if (j < layers.size()) {
                    auto* ap = dynamic_cast<const AdaptiveAvgPool2d*>(layers[j].get());
                    if (ap && ap->output_size == Pair{1, 1}) {
                        gap = 1;
                        ++j;
                        if (j < layers.size()) {
                            auto* fl = dynamic_cast<const Flatten*>(layers[j].get());
                            if (fl && fl->start_dim == 1) { gap = 2; ++j; }
                        }
                    }
                }
                Tensor y = x.conv_stack(run, gap);
                if (y.defined()) {
                    x = y;
                    i = j - 1;
                    continue;
                }
            }
        }
        if (fuse && Config::fuse_small_mlp() && x.shape().size() == 2) {
            // peephole: [Linear (+ ReLU)]{2,4} with every width <= 128 -> one forward / one backward launch (tp_mlp_small_*)
            std::vector<MlpLayer> run;
            size_t j = i;
            size_t width = x.shape()[1];
            while (j < layers.size() && run.size() < 4) {
                auto* lin = dynamic_cast<const Linear*>(layers[j].get());
                if (!lin || width > 128 || lin->weight.shape()[0] > 128 || lin->weight.shape()[1] != width) break;
                MlpLayer ly;
                ly.weight = lin->weight;
                ly.bias = lin->bias;
                ++j;
                if (j < layers.size() && dynamic_cast<const ReLU*>(layers[j].get())) {
                    ly.relu = true;
                    ++j;
                }
                width = lin->weight.shape()[0];
                run.push_back(ly);
            }
            if (run.size() >= 2) {
                Tensor y = x.mlp_chain(run);
                if (y.defined()) {
                    x = y;
                    i = j - 1;
                    continue;
                }
            }
        }
        if (fuse && i + 1 < layers.size()) {
            // peephole: Linear followed by ReLU -> bias + ReLU in the GEMM epilogue, mask folded into backward
            auto* lin = dynamic_cast<const Linear*>(layers[i].get());
            auto* act = dynamic_cast<const ReLU*>(layers[i + 1].get());
            if (lin && act) {
                x = lin->forward_fused_relu(x);
                ++i;
                continue;
            }
        }
        x = layers[i]->forward(x);
    }
    return x;
}

std::vector<Tensor> Sequential::parameters() const {                             // src/nn.rs:153-155
    std::vector<Tensor> out;
    for (auto& l : layers)
        for (auto& p : l->parameters()) out.push_back(p);
    return out;
}

Conv2d::Conv2d(size_t in_channels, size_t out_channels, Pair kernel_size, std::optional<Pair> stride_,
               std::optional<Pair> padding_, std::optional<Pair> dilation_, std::optional<size_t> groups_, bool with_bias,
               uint64_t seed)
    : stride(stride_.value_or(Pair{1, 1})), padding(padding_.value_or(Pair{0, 0})),
      dilation(dilation_.value_or(Pair{1, 1})), groups(groups_.value_or(1)) {
    if (groups != 1) panic("Conv2d: groups > 1 is outside the hot path (the reference's grouped path records no gradients)");
    size_t fan_in = in_channels * kernel_size.first * kernel_size.second / groups;
    float bound = std::sqrt(2.0f / (float)fan_in) * std::sqrt(3.0f);             // src/nn.rs:219-221
    weight = Tensor::create(uniform_init(out_channels * in_channels * kernel_size.first * kernel_size.second, bound, seed),
                            {out_channels, in_channels, kernel_size.first, kernel_size.second}).requires_grad();
    if (with_bias) bias = Tensor::zeros({out_channels}).requires_grad();
}

Tensor Conv2d::forward(const Tensor& input) const {                              // src/nn.rs:279-289
    return input.conv2d(weight, bias ? &*bias : nullptr, stride, padding, dilation);
}

std::vector<Tensor> Conv2d::parameters() const {
    std::vector<Tensor> p{weight};
    if (bias) p.push_back(*bias);
    return p;
}

Tensor Conv2dReLU::forward(const Tensor& input) const {                          // src/nn.rs:470-479
    return input.conv2d_relu(weight, bias ? &*bias : nullptr, stride, padding, dilation);
}

Tensor AdaptiveAvgPool2d::forward(const Tensor& input) const {                   // src/nn.rs:670-686
    if (input.shape().size() != 4) panic("AdaptiveAvgPool2d expects [N,C,H,W]");
    size_t kh = input.shape()[2] / output_size.first, kw = input.shape()[3] / output_size.second;
    return input.avg_pool2d({kh, kw}, Pair{kh, kw}, {0, 0});
}

}  // namespace nn

// =====================================================================================================
// loss  (src/loss.rs)
// =====================================================================================================
namespace loss {

Tensor log_softmax(const Tensor& x, int dim) {                                   // src/loss.rs:101-126, op for op
    int nd = (int)x.shape().size();
    int d = dim < 0 ? nd + dim : dim;
    if (d != nd - 1 || nd != 2) panic("Only last-dim log_softmax on [B,C] is supported");
    Tensor max_vals = x.max((size_t)d).first;                                    // no tape node
    Tensor shifted = x.sub_broadcast_rows(max_vals);
    Tensor sum_exp = shifted.exp().sum((size_t)d, true);
    Tensor log_sum = sum_exp.log();
    return shifted.sub_broadcast_rows(log_sum);
}

Tensor softmax(const Tensor& x, int dim) {
    // The reference body uses non-broadcasting `-` and `/` and panics for C > 1; this is the
    // evident intent: a row-wise stable softmax (single fused kernel, no tape node).
    int nd = (int)x.shape().size();
    int d = dim < 0 ? nd + dim : dim;
    if (d != nd - 1 || nd != 2) panic("Only last-dim softmax on [B,C] is supported");
    Tensor out = Tensor::empty(x.shape());
    check(tp_softmax_fwd(ctx(), x.buf(), out.buf(), (int)x.shape()[0], (int)x.shape()[1]));
    return out;
}

// ---- BCE / MSE / one-hot cross-entropy ---------------------------------------------------------------------
Tensor bce_loss(const Tensor& predictions, const Tensor& targets) {             // src/loss.rs:6-72
    if (predictions.numel() != targets.numel()) panic("bce_loss: predictions and targets must match in length");
    size_t n = predictions.numel();
    Tensor out = Tensor::empty({1});
    check(tp_bce_fwd(ctx(), predictions.buf(), targets.buf(), out.buf(), n));
    if (predictions.needs_grad() || targets.needs_grad()) {
        out.set_requires_grad(true);
        Tensor p = predictions, t = targets;
        Tape::push_binary_op(p, t, out, [p, t, out, n]() {
            tp_buf* g = out.grad_buf();
            if (!g) return;
            int ap = 0, at = 0;
            tp_buf* gp = p.needs_grad() ? p.impl()->grad_for_write(&ap) : nullptr;
            tp_buf* gt = t.needs_grad() ? t.impl()->grad_for_write(&at) : nullptr;
            check(tp_bce_bwd(ctx(), p.buf(), t.buf(), g, gp, gt, n, ap, at));
        });
    }
    return out;
}

Tensor mse_loss(const Tensor& predictions, const Tensor& targets) {             // src/loss.rs:75-80
    Tensor diff = predictions - targets;
    Tensor squared = diff * diff;
    return squared.mean();
}

Tensor one_hot(const Tensor& indices, size_t num_classes) {                      // src/loss.rs:248-268 (host loop, as the reference)
    if (indices.shape().size() != 1) panic("Indices must be 1D");
    size_t b = indices.shape()[0];
    const std::vector<float>& idx = indices.data();
    std::vector<float> oh(b * num_classes, 0.0f);
    for (size_t i = 0; i < b; ++i) {
        size_t c = idx[i] > 0.0f ? (size_t)idx[i] : 0;
        if (c >= num_classes) panic("Index %zu out of bounds for %zu classes", c, num_classes);
        oh[i * num_classes + c] = 1.0f;
    }
    return Tensor::create(oh, {b, num_classes});
}

Tensor cross_entropy_loss_onehot(const Tensor& logits, const Tensor& targets) { // src/loss.rs:202-245
    if (logits.shape() != targets.shape()) panic("Logits and targets shapes must match");
    if (logits.shape().size() != 2) panic("Must be 2D tensors");
    size_t batch = logits.shape()[0], total = logits.numel();
    // forward: -sum(targets * log_softmax(logits)) / batch; the reference reads the sum back and builds a fresh scalar (:219),
    // so the composed ops below it carry no gradient into the loss: the backward is the direct closure (:227-240)
    Tensor logits_ng = logits;
    logits_ng.set_requires_grad(false);
    Tensor tg = targets;
    tg.set_requires_grad(false);
    Tensor logp = log_softmax(logits_ng, -1);
    Tensor sum = (tg * logp).sum(std::nullopt, false);
    Tensor out = Tensor::empty({1});
    check(tp_accumulate(ctx(), out.buf(), sum.buf(), -1.0f / (float)batch, 1, 0));
    if (logits.needs_grad()) {
        out.set_requires_grad(true);
        Tensor lg = logits;
        Tape::push_unary_op(lg, out, [lg, tg, logp, out, batch, total]() {
            tp_buf* g = out.grad_buf();
            if (!g) return;
            // (softmax - targets) * g / batch: softmax = exp(logp); two scaled accumulations with the device scalar g
            Tensor probs = Tensor::empty(lg.shape());
            check(tp_exp_fwd(ctx(), logp.buf(), probs.buf(), total));
            Tensor diff = Tensor::empty(lg.shape());
            check(tp_sub(ctx(), probs.buf(), tg.buf(), diff.buf(), total));
            Tensor gb = Tensor::empty({1});                  // g / batch
            check(tp_accumulate(ctx(), gb.buf(), g, 1.0f / (float)batch, 1, 0));
            Tensor scaled = Tensor::empty(lg.shape());
            check(tp_broadcast_bwd(ctx(), gb.buf(), scaled.buf(), 1, (int)total, 2, 0));      // every element = g / batch
            int acc;
            tp_buf* gin = lg.impl()->grad_for_write(&acc);
            check(tp_mul_bwd(ctx(), diff.buf(), scaled.buf(), gin, total, acc));
        });
    }
    return out;
}

Tensor cross_entropy_loss_into(const Tensor& logits, const Tensor& targets, const Tensor& out) {
    return cross_entropy_with_accuracy_into(logits, targets, out, nullptr);
}

Tensor cross_entropy_with_accuracy_into(const Tensor& logits, const Tensor& targets, const Tensor& out, const Tensor* correct_out) {
    const Shape& ts = targets.shape();
    if (!(ts.size() == 1 || (ts.size() == 2 && ts[1] == 1))) panic("targets must be [B] or [B,1]");     // src/loss.rs:137-141
    if (logits.shape().size() != 2 || logits.shape()[0] != ts[0]) panic("logits must be [B,C] with the targets' batch size");
    int rows = (int)logits.shape()[0], cols = (int)logits.shape()[1];
    Tensor logp = Tensor::empty(logits.shape());
    // fused log_softmax + NLL mean (src/loss.rs:152-165); the six log_softmax nodes the reference records are
    // dead in backward, so a single node carrying the direct gradient is recorded instead.
    if (rows > 0)
        check(tp_softmax_xent_acc_fwd(ctx(), logits.buf(), targets.buf(), logp.buf(), out.buf(), correct_out ? correct_out->buf() : nullptr,
                                      rows, cols));
    else
        check(tp_softmax_xent_fwd(ctx(), logits.buf(), targets.buf(), logp.buf(), out.buf(), rows, cols));
    Tensor o = out;
    if (logits.needs_grad()) {
        o.set_requires_grad(true);
        Tensor lg = logits, t = targets;
        Tape::push_unary_op(lg, o, [lg, t, logp, o, rows, cols]() {             // src/loss.rs:174-191
            tp_buf* g = o.grad_buf();
            if (!g) return;
            int acc;
            tp_buf* gl = lg.impl()->grad_for_write(&acc);
            check(tp_softmax_xent_bwd(ctx(), logp.buf(), t.buf(), g, gl, rows, cols, acc));
        });
    }
    return o;
}

Tensor cross_entropy_loss(const Tensor& logits, const Tensor& targets) {         // src/loss.rs:136-195
    return cross_entropy_loss_into(logits, targets, Tensor::empty({1}));
}

void accuracy_count_into(const Tensor& predictions, const Tensor& targets, const Tensor& out) {
    if (predictions.shape().size() != 2 || predictions.shape()[0] != targets.shape()[0]) panic("accuracy: batch size mismatch");
    check(tp_accuracy_count(ctx(), predictions.buf(), targets.buf(), out.buf(), (int)predictions.shape()[0], (int)predictions.shape()[1]));
}

Tensor accuracy_count(const Tensor& predictions, const Tensor& targets) {
    Tensor out = Tensor::empty({1});
    accuracy_count_into(predictions, targets, out);
    return out;
}

float accuracy(const Tensor& predictions, const Tensor& targets) {              // src/loss.rs:271-290
    float correct = accuracy_count(predictions, targets).item();
    return correct / (float)targets.shape()[0];
}

}  // namespace loss

// =====================================================================================================
// optim  (src/optim.rs)
// =====================================================================================================
namespace optim {

// Flat arenas: parameters, gradients and Adam moments each live in ONE contiguous device buffer (every
// parameter starts on a 16-byte boundary).  The optimizer step is then one fused launch and the
// data-parallel exchange is one allreduce over the gradient arena.  Tensor handles keep working because
// their buffers are re-pointed at slices of the arena.
struct Arena {
    std::vector<Tensor> params;
    std::vector<size_t> off;
    size_t total = 0;
    tp_buf *p = nullptr, *g = nullptr, *m = nullptr, *v = nullptr, *hyper = nullptr;
    std::vector<tp_buf*> ps, gs, ms, vs;          // per-parameter slices

    Arena(std::vector<Tensor> prm, bool moments) : params(std::move(prm)) {
        for (auto& t : params) {
            off.push_back(total);
            total += (t.numel() + 3) & ~(size_t)3;
        }
        tp_ctx* c = ctx();
        size_t cap = total ? total : 4;
        check(tp_buf_alloc(c, cap, &p));
        check(tp_buf_alloc(c, cap, &g));
        check(tp_buf_fill(c, p, 0.0f, cap));
        check(tp_buf_fill(c, g, 0.0f, cap));
        if (moments) {
            check(tp_buf_alloc(c, cap, &m));
            check(tp_buf_alloc(c, cap, &v));
            check(tp_buf_fill(c, m, 0.0f, cap));                                 // src/optim.rs:62-71
            check(tp_buf_fill(c, v, 0.0f, cap));
            check(tp_buf_alloc(c, 8, &hyper));
        }
        for (size_t i = 0; i < params.size(); ++i) {
            TensorImpl& im = *params[i].impl();
            size_t n = im.n;
            tp_buf *sp, *sg, *sm = nullptr, *sv = nullptr;
            check(tp_buf_slice(p, off[i], n, &sp));
            check(tp_buf_slice(g, off[i], n, &sg));
            check(tp_buf_copy(c, sp, im.buf, n));
            if (im.has_grad && im.grad) check(tp_buf_copy(c, sg, im.grad, n));
            tp_buf_release(im.buf);
            if (im.grad) tp_buf_release(im.grad);
            im.buf = sp;  tp_buf_retain(sp);
            im.grad = sg; tp_buf_retain(sg);
            if (moments) {
                check(tp_buf_slice(m, off[i], n, &sm));
                check(tp_buf_slice(v, off[i], n, &sv));
            }
            ps.push_back(sp); gs.push_back(sg); ms.push_back(sm); vs.push_back(sv);
        }
    }
    ~Arena() {
        for (auto* b : ps) tp_buf_release(b);
        for (auto* b : gs) tp_buf_release(b);
        for (auto* b : ms) tp_buf_release(b);
        for (auto* b : vs) tp_buf_release(b);
        tp_buf_release(p); tp_buf_release(g); tp_buf_release(m); tp_buf_release(v); tp_buf_release(hyper);
    }
    bool all_have_grad() const {
        for (auto& t : params) if (!t.impl()->has_grad) return false;
        return true;
    }
    void bump_versions() { for (auto& t : params) t.impl()->version++; }
};

void Optimizer::mark_parameters_updated() const { arena()->bump_versions(); }

tp_buf* arena_grad_buf(const std::shared_ptr<Arena>& a) { return a->g; }
size_t arena_total(const std::shared_ptr<Arena>& a) { return a->total; }
tp_buf* arena_param_buf(const std::shared_ptr<Arena>& a) { return a->p; }

bool describe_fused_step(const nn::Module& model, const Optimizer& opt, size_t batch, tp_step_desc* d, tp_buf* bufs[5]) {
    auto* seq = dynamic_cast<const nn::Sequential*>(&model);
    if (!seq || Config::reference_op_sequence()) return false;
    std::shared_ptr<Arena> a = opt.arena();
    if (!a) return false;
    std::memset(d, 0, sizeof(*d));
    auto offset_of = [&](const Tensor& t, int64_t* off) {
        for (size_t i = 0; i < a->params.size(); ++i)
            if (a->params[i].impl() == t.impl()) { *off = (int64_t)a->off[i]; return t.needs_grad(); }
        return false;
    };
    size_t matched = 0;
    int L = 0;
    const auto& layers = seq->layers;
    for (size_t i = 0; i < layers.size(); ++i) {
        auto* lin = dynamic_cast<const nn::Linear*>(layers[i].get());
        if (!lin || L >= TP_STEP_MAX_LAYERS) return false;
        const Shape& ws = lin->weight.shape();
        if (ws.size() != 2) return false;
        if (L == 0) d->dims[0] = (int)ws[1];
        else if ((size_t)d->dims[L] != ws[1]) return false;
        d->dims[L + 1] = (int)ws[0];
        if (!offset_of(lin->weight, &d->w_off[L])) return false;
        matched++;
        d->b_off[L] = -1;
        if (lin->bias) {
            if (!offset_of(*lin->bias, &d->b_off[L])) return false;
            matched++;
        }
        bool relu = i + 1 < layers.size() && dynamic_cast<const nn::ReLU*>(layers[i + 1].get()) != nullptr;
        d->relu[L] = relu ? 1 : 0;
        if (relu) ++i;
        ++L;
    }
    if (L == 0 || d->relu[L - 1] || matched != a->params.size()) return false;
    d->n_layers = L;
    d->batch = (int)batch;
    d->optimizer = opt.kind();
    d->arena_len = (int64_t)a->total;
    bufs[0] = a->p; bufs[1] = a->g; bufs[2] = a->m; bufs[3] = a->v; bufs[4] = a->hyper;
    return tp_step_supported(d) != 0;
}

// ---- SGD  (src/optim.rs:8-40) ----------------------------------------------------------------------
SGD::SGD(std::vector<Tensor> params, float lr, std::optional<float> /*momentum: ignored, :14-17*/)
    : params_(params), lr_(lr), arena_(std::make_shared<Arena>(std::move(params), false)) {
    // the learning rate lives on the device (like Adam's hyper buffer): a captured step follows set_lr
    check(tp_buf_alloc(ctx(), 4, &lr_dev_));
    check(tp_buf_set_scalar(ctx(), lr_dev_, 0, lr_));
}
SGD::~SGD() { tp_buf_release(lr_dev_); }

void SGD::set_lr(float lr) {
    lr_ = lr;
    check(tp_buf_set_scalar(ctx(), lr_dev_, 0, lr));
}

void SGD::step() {                                                               // :21-33
    Arena& a = *arena_;
    if (a.all_have_grad()) {
        check(tp_sgd_step_dev(ctx(), a.p, a.g, lr_dev_, grad_scale_, a.total));
    } else {
        for (size_t i = 0; i < a.params.size(); ++i)
            if (a.params[i].impl()->has_grad) check(tp_sgd_step_dev(ctx(), a.ps[i], a.gs[i], lr_dev_, grad_scale_, a.params[i].numel()));
    }
    a.bump_versions();
}

void SGD::zero_grad() { for (auto& p : params_) p.zero_grad(); }                 // :35-39

// ---- Adam  (src/optim.rs:43-128) -----------------------------------------------------------------------
Adam::Adam(std::vector<Tensor> params, float lr, std::optional<std::pair<float, float>> betas, std::optional<float> eps,
           std::optional<float> weight_decay)
    : params_(params), lr_(lr), beta1_(betas ? betas->first : 0.9f), beta2_(betas ? betas->second : 0.999f),
      eps_(eps.value_or(1e-8f)), weight_decay_(weight_decay.value_or(0.0f)),
      arena_(std::make_shared<Arena>(std::move(params), true)) {
    check(tp_adam_hyper_init(ctx(), arena_->hyper, lr_, beta1_, beta2_, eps_, weight_decay_));
}
Adam::~Adam() = default;

void Adam::set_lr(float lr) {                                                    // :125-127
    lr_ = lr;
    check(tp_adam_hyper_set_lr(ctx(), arena_->hyper, lr));
}

void Adam::step_impl(bool decoupled) {
    Arena& a = *arena_;
    tp_ctx* c = ctx();
    t_ += 1;                                                                     // :86 (even if every grad is None)
    check(tp_adam_advance(c, a.hyper));                                          // same increment, on the device
    if (a.all_have_grad()) {
        check(tp_adam_step_dev(c, a.p, a.g, a.m, a.v, a.hyper, grad_scale_, decoupled ? 1 : 0, a.total));
    } else {
        // some parameters have no gradient (None): Adam skips them, AdamW still decays them — one multi-tensor launch
        std::vector<int64_t> offs, lens;
        std::vector<int> modes;
        for (size_t i = 0; i < a.params.size(); ++i) {
            offs.push_back((int64_t)a.off[i]);
            lens.push_back((int64_t)a.params[i].numel());
            modes.push_back(a.params[i].impl()->has_grad ? 1 : (decoupled ? 2 : 0));
        }
        check(tp_adam_step_segments(c, a.p, a.g, a.m, a.v, a.hyper, grad_scale_, decoupled ? 1 : 0, offs.data(), lens.data(), modes.data(),
                                    (int)offs.size()));
    }
    a.bump_versions();
}

void Adam::step() { step_impl(false); }                                          // :83-113
void Adam::zero_grad() { for (auto& p : params_) p.zero_grad(); }                // :115-119

// ---- AdamW  (src/optim.rs:131-181) ---------------------------------------------------------------------
AdamW::AdamW(std::vector<Tensor> params, float lr, std::optional<std::pair<float, float>> betas, std::optional<float> eps,
             std::optional<float> weight_decay)
    : adam_(std::move(params), lr, betas, eps, weight_decay) {}

void AdamW::step() { adam_.step_impl(true); }                                    // :148-168: p *= 1 - lr*wd, then Adam with wd = 0

// ---- LR schedulers  (src/optim.rs:184-352), host scalars -------------------------------------------------
void StepLR::step(std::optional<float>) {
    current_epoch += 1;
    if (current_epoch % step_size == 0) current_lr *= gamma;
}

void CosineAnnealingLR::step(std::optional<float>) {
    current_epoch += 1;
    float progress = (float)current_epoch / (float)t_max;
    float cos_val = (1.0f + std::cos(progress * 3.14159265358979323846f)) / 2.0f;
    current_lr = min_lr + (base_lr - min_lr) * cos_val;
}

ReduceLROnPlateau::ReduceLROnPlateau(float initial_lr, float factor_, size_t patience_, std::optional<float> min_lr_,
                                     std::optional<std::string> mode_)
    : current_lr(initial_lr), factor(factor_), min_lr(min_lr_.value_or(1e-6f)), patience(patience_), mode(mode_.value_or("min")) {
    best_metric = mode == "min" ? INFINITY : -INFINITY;
}

void ReduceLROnPlateau::step(std::optional<float> metrics) {
    if (!metrics) return;
    bool improved = mode == "min" ? *metrics < best_metric : *metrics > best_metric;
    if (improved) {
        best_metric = *metrics;
        patience_counter = 0;
    } else {
        patience_counter += 1;
        if (patience_counter >= patience) {
            current_lr = std::max(current_lr * factor, min_lr);
            patience_counter = 0;
        }
    }
}

}  // namespace optim

// =====================================================================================================
// dist: data-parallel group (no counterpart in the reference)
// =====================================================================================================
namespace dist {
namespace {
thread_local int g_rank = 0, g_world = 1;
}
void init(int rank, int world, const void* nccl_unique_id128) {
    check(tp_comm_init(ctx(), rank, world, nccl_unique_id128));
    g_rank = rank;
    g_world = world;
}
int rank() { return g_rank; }
int world() { return g_world; }
void allreduce_sum(tp_buf* buf, size_t n) { check(tp_allreduce_sum(ctx(), buf, n)); }
void broadcast(tp_buf* buf, size_t n, int root) { check(tp_broadcast(ctx(), buf, n, root)); }
}  // namespace dist

}  // namespace taper

// Flat C entry points of the host layer (include/taper_b200_host.h).  No exception crosses the boundary.
#include "taper_internal.hpp"
#include "../../../include/taper_b200_host.h"
#include "../common.cuh"

#include <cstring>
#include <sstream>

using namespace taper;

struct tp_model {
    std::shared_ptr<nn::Sequential> seq;
    std::vector<Tensor> params;
};

struct tp_trainer {
    std::shared_ptr<train::Trainer> tr;
};

struct tp_dataset {
    std::shared_ptr<data::MNISTDataset> ds;
};

struct tp_loader {
    std::unique_ptr<data::DataLoader> ld;
};

struct tp_scheduler {
    std::shared_ptr<optim::LRScheduler> sc;
};

struct tp_tensor {
    Tensor t;
};

struct tp_optimizer {
    std::shared_ptr<optim::Optimizer> opt;
};

namespace {

template <class F>
int guarded(F&& f) {
    try {
        f();
        return TP_OK;
    } catch (const std::exception& e) {
        std::string msg = e.what();          // copy first: set_error overwrites the thread-local a nested what() may alias
        tp::set_error("%s", msg.c_str());
        return TP_ERR_INVALID;
    }
}

std::vector<std::string> split(const std::string& s, char sep) {
    std::vector<std::string> out;
    std::stringstream ss(s);
    std::string item;
    while (std::getline(ss, item, sep)) out.push_back(item);
    return out;
}

size_t num(const std::vector<std::string>& f, size_t i, const std::string& layer) {
    if (i >= f.size()) panic("layer '%s': missing field %zu", layer.c_str(), i);
    return (size_t)std::stoul(f[i]);
}

Shape to_shape(const size_t* dims, int ndim) { return Shape(dims, dims + ndim); }

}  // namespace

extern "C" {

int tp_host_set_device(int device) { return guarded([&] { set_device(device); }); }

int tp_host_ctx(tp_ctx** out) {
    return guarded([&] {
        if (!out) panic("tp_host_ctx: NULL out pointer");
        *out = ctx();
    });
}

int tp_host_config(int conv_full_adjoint, int fuse_linear_relu, int reference_op_sequence, int gemm_mode) {
    return guarded([&] {
        if (conv_full_adjoint >= 0) Config::conv_full_adjoint() = conv_full_adjoint != 0;
        if (fuse_linear_relu >= 0) Config::fuse_linear_relu() = fuse_linear_relu != 0;
        if (reference_op_sequence >= 0) Config::reference_op_sequence() = reference_op_sequence != 0;
        if (gemm_mode >= 0) check(tp_set_gemm_mode(ctx(), gemm_mode));
    });
}

int tp_host_config_conv_stack(int fuse_conv_stack) {
    return guarded([&] { Config::fuse_conv_stack() = fuse_conv_stack != 0; });
}

int tp_host_config_small_mlp(int fuse_small_mlp) {
    return guarded([&] { Config::fuse_small_mlp() = fuse_small_mlp != 0; });
}

int tp_model_create(const char* spec, uint64_t seed, tp_model** out) {
    return guarded([&] {
        if (!spec || !out) panic("tp_model_create: NULL argument");
        std::vector<std::shared_ptr<nn::Module>> layers;
        uint64_t s = seed;
        for (auto& item : split(spec, ',')) {
            auto f = split(item, ':');
            if (f.empty() || f[0].empty()) continue;
            const std::string& k = f[0];
            if (k == "linear") {
                bool bias = !(f.size() > 3 && f[3] == "nobias");
                layers.push_back(std::make_shared<nn::Linear>(num(f, 1, item), num(f, 2, item), bias, s++));
            } else if (k == "relu") {
                layers.push_back(std::make_shared<nn::ReLU>());
            } else if (k == "sigmoid") {
                layers.push_back(std::make_shared<nn::Sigmoid>());
            } else if (k == "conv" || k == "conv_relu") {
                size_t cin = num(f, 1, item), cout = num(f, 2, item), ks = num(f, 3, item), st = num(f, 4, item), pd = num(f, 5, item);
                if (k == "conv")
                    layers.push_back(std::make_shared<nn::Conv2d>(cin, cout, Pair{ks, ks}, Pair{st, st}, Pair{pd, pd}, std::nullopt, std::nullopt, true, s++));
                else
                    layers.push_back(std::make_shared<nn::Conv2dReLU>(cin, cout, Pair{ks, ks}, Pair{st, st}, Pair{pd, pd}, std::nullopt, std::nullopt, true, s++));
            } else if (k == "maxpool") {
                size_t ks = num(f, 1, item), st = num(f, 2, item);
                layers.push_back(std::make_shared<nn::MaxPool2d>(Pair{ks, ks}, Pair{st, st}, std::nullopt));
            } else if (k == "avgpool") {
                size_t ks = num(f, 1, item), st = num(f, 2, item);
                layers.push_back(std::make_shared<nn::AvgPool2d>(Pair{ks, ks}, Pair{st, st}, std::nullopt));
            } else if (k == "gap") {
                layers.push_back(std::make_shared<nn::AdaptiveAvgPool2d>(nn::AdaptiveAvgPool2d::global()));
            } else if (k == "flatten") {
                layers.push_back(std::make_shared<nn::Flatten>(1));
            } else if (k == "dropout") {                      // dropout:PERCENT (src/nn.rs:775-827)
                layers.push_back(std::make_shared<nn::Dropout>((float)num(f, 1, item) / 100.0f, s++));
            } else {
                panic("tp_model_create: unknown layer '%s'", item.c_str());
            }
        }
        auto* m = new tp_model();
        m->seq = std::make_shared<nn::Sequential>(std::move(layers));
        m->params = m->seq->parameters();
        *out = m;
    });
}

int tp_model_destroy(tp_model* m) {
    return guarded([&] { delete m; });
}

int tp_model_num_params(tp_model* m, int* count) {
    return guarded([&] {
        if (!m || !count) panic("tp_model_num_params: NULL argument");
        *count = (int)m->params.size();
    });
}

static Tensor& param_at(tp_model* m, int index) {
    if (!m || index < 0 || (size_t)index >= m->params.size()) panic("parameter index %d out of range", index);
    return m->params[index];
}

int tp_model_param_info(tp_model* m, int index, size_t* numel, int* ndim, size_t* dims4) {
    return guarded([&] {
        Tensor& p = param_at(m, index);
        if (numel) *numel = p.numel();
        if (ndim) *ndim = (int)p.shape().size();
        if (dims4) for (size_t i = 0; i < p.shape().size() && i < 4; ++i) dims4[i] = p.shape()[i];
    });
}

int tp_model_set_param(tp_model* m, int index, const float* host, size_t n) {
    return guarded([&] {
        Tensor& p = param_at(m, index);
        if (!host || n != p.numel()) panic("tp_model_set_param: expected %zu floats", p.numel());
        p.set_data(std::vector<float>(host, host + n));
    });
}

int tp_model_get_param(tp_model* m, int index, float* host, size_t n) {
    return guarded([&] {
        Tensor& p = param_at(m, index);
        if (!host || n != p.numel()) panic("tp_model_get_param: expected %zu floats", p.numel());
        std::memcpy(host, p.data().data(), n * sizeof(float));
    });
}

int tp_model_get_grad(tp_model* m, int index, float* host, size_t n, int* has_grad) {
    return guarded([&] {
        Tensor& p = param_at(m, index);
        tp_buf* g = p.grad_buf();
        if (has_grad) *has_grad = g ? 1 : 0;
        if (g && host) {
            if (n != p.numel()) panic("tp_model_get_grad: expected %zu floats", p.numel());
            check(tp_buf_download(ctx(), g, host, n));
        }
    });
}

int tp_model_zero_grad(tp_model* m) {
    return guarded([&] {
        if (!m) panic("tp_model_zero_grad: NULL model");
        for (auto& p : m->params) p.zero_grad();
    });
}

int tp_model_forward(tp_model* m, const float* x, const size_t* shape, int ndim, float* out, size_t out_cap, size_t* out_n) {
    return guarded([&] {
        if (!m || !x || !shape || !out) panic("tp_model_forward: NULL argument");
        Tensor in = Tensor::from_host(x, to_shape(shape, ndim));
        Tensor y = m->seq->forward(in);
        if (out_n) *out_n = y.numel();
        if (y.numel() > out_cap) panic("tp_model_forward: output needs %zu floats, %zu available", y.numel(), out_cap);
        std::memcpy(out, y.data().data(), y.numel() * sizeof(float));
        Tape::reset();
    });
}

int tp_model_loss_backward(tp_model* m, const float* x, const size_t* shape, int ndim, const float* labels, float* loss,
                           float* correct, size_t* tape_len) {
    return guarded([&] {
        if (!m || !x || !shape || !labels) panic("tp_model_loss_backward: NULL argument");
        Tape::reset();
        Shape s = to_shape(shape, ndim);
        Tensor in = Tensor::from_host(x, s);
        Tensor t = Tensor::from_host(labels, {s[0]});
        Tensor logits = m->seq->forward(in);
        Tensor l = loss::cross_entropy_loss(logits, t);
        Tensor c = loss::accuracy_count(logits, t);
        if (tape_len) *tape_len = Tape::len();
        l.backward();
        if (loss) *loss = l.item();
        if (correct) *correct = c.item();
        Tape::reset();
    });
}

int tp_model_regression_backward(tp_model* m, const float* x, const size_t* shape, int ndim, const float* targets, size_t n_targets,
                                 const char* loss_kind, float* loss) {
    return guarded([&] {
        if (!m || !x || !shape || !targets || !loss_kind) panic("tp_model_regression_backward: NULL argument");
        Tape::reset();
        Shape s = to_shape(shape, ndim);
        Tensor in = Tensor::from_host(x, s);
        Tensor out = m->seq->forward(in);
        if (out.numel() != n_targets) panic("tp_model_regression_backward: %zu outputs but %zu targets", out.numel(), n_targets);
        Tensor t = Tensor::from_host(targets, out.shape());
        std::string k = loss_kind;
        Tensor l;
        if (k == "bce") l = loss::bce_loss(out, t);
        else if (k == "mse") l = loss::mse_loss(out, t);
        else if (k == "ce_onehot") l = loss::cross_entropy_loss_onehot(out, t);
        else panic("tp_model_regression_backward: unknown loss '%s'", loss_kind);
        l.backward();
        if (loss) *loss = l.item();
        Tape::reset();
    });
}

int tp_trainer_create(tp_model* m, const char* optimizer, float lr, float beta1, float beta2, float eps, float weight_decay,
                      tp_trainer** out) {
    return guarded([&] {
        if (!m || !optimizer || !out) panic("tp_trainer_create: NULL argument");
        std::shared_ptr<optim::Optimizer> opt;
        std::string k = optimizer;
        if (k == "sgd") opt = std::make_shared<optim::SGD>(m->params, lr);
        else if (k == "adam") opt = std::make_shared<optim::Adam>(m->params, lr, std::make_pair(beta1, beta2), eps, weight_decay);
        else if (k == "adamw") opt = std::make_shared<optim::AdamW>(m->params, lr, std::make_pair(beta1, beta2), eps, weight_decay);
        else panic("tp_trainer_create: unknown optimizer '%s'", optimizer);
        auto* t = new tp_trainer();
        t->tr = std::make_shared<train::Trainer>(m->seq, opt, nullptr);
        *out = t;
    });
}

int tp_trainer_destroy(tp_trainer* t) {
    return guarded([&] { delete t; });
}

#define TRAINER(t) do { if (!(t) || !(t)->tr) panic("%s: NULL trainer", __func__); } while (0)

int tp_trainer_set_lr(tp_trainer* t, float lr) {
    return guarded([&] { TRAINER(t); t->tr->optimizer->set_lr(lr); });
}

int tp_trainer_set_use_graph(tp_trainer* t, int on) {
    return guarded([&] { TRAINER(t); t->tr->set_use_graph(on != 0); });
}

int tp_trainer_step(tp_trainer* t, const float* images, const float* labels, size_t batch, const size_t* sample_shape, int ndim,
                    float* loss, float* correct) {
    return guarded([&] {
        TRAINER(t);
        train::StepResult r = t->tr->train_batch(images, labels, batch, to_shape(sample_shape, ndim));
        if (loss) *loss = r.loss;
        if (correct) *correct = r.correct;
    });
}

int tp_trainer_step_async(tp_trainer* t, const float* images, const float* labels, size_t batch, const size_t* sample_shape,
                          int ndim, int pinned) {
    return guarded([&] { TRAINER(t); t->tr->train_batch_async(images, labels, batch, to_shape(sample_shape, ndim), pinned != 0); });
}

int tp_trainer_load_dataset(tp_trainer* t, const float* images, const float* labels, size_t n, const size_t* sample_shape,
                            int ndim, const uint32_t* perm) {
    return guarded([&] { TRAINER(t); t->tr->load_dataset(images, labels, n, to_shape(sample_shape, ndim), perm); });
}

int tp_trainer_step_async_u8(tp_trainer* t, const void* images_u8, const float* labels, size_t batch, const size_t* sample_shape,
                             int ndim, int pinned) {
    return guarded([&] {
        TRAINER(t);
        t->tr->train_batch_async_u8(static_cast<const uint8_t*>(images_u8), labels, batch, to_shape(sample_shape, ndim), pinned != 0);
    });
}

int tp_trainer_load_dataset_u8(tp_trainer* t, const void* images_u8, const float* labels, size_t n, const size_t* sample_shape,
                               int ndim, const uint32_t* perm) {
    return guarded([&] {
        TRAINER(t);
        t->tr->load_dataset_u8(static_cast<const uint8_t*>(images_u8), labels, n, to_shape(sample_shape, ndim), perm);
    });
}

int tp_trainer_fused_kind(tp_trainer* t, int* kind) {
    return guarded([&] { TRAINER(t); if (kind) *kind = t->tr->fused_kind(); });
}

int tp_trainer_step_resident(tp_trainer* t, size_t batch) {
    return guarded([&] { TRAINER(t); t->tr->train_batch_resident(batch); });
}

int tp_trainer_fetch(tp_trainer* t, float* loss, float* correct) {
    return guarded([&] {
        TRAINER(t);
        train::StepResult r = t->tr->fetch();
        if (loss) *loss = r.loss;
        if (correct) *correct = r.correct;
    });
}

int tp_trainer_pending(tp_trainer* t, size_t* count) {
    return guarded([&] { TRAINER(t); if (count) *count = t->tr->pending(); });
}

int tp_trainer_eval(tp_trainer* t, const float* images, const float* labels, size_t batch, const size_t* sample_shape, int ndim,
                    float* loss, float* correct) {
    return guarded([&] {
        TRAINER(t);
        train::StepResult r = t->tr->eval_batch(images, labels, batch, to_shape(sample_shape, ndim));
        if (loss) *loss = r.loss;
        if (correct) *correct = r.correct;
    });
}

int tp_trainer_save_checkpoint(tp_trainer* t, const char* path) {
    return guarded([&] { TRAINER(t); t->tr->save_checkpoint(path); });
}

int tp_trainer_load_checkpoint(tp_trainer* t, const char* path) {
    return guarded([&] { TRAINER(t); t->tr->load_checkpoint(path); });
}

int tp_trainer_comm_init(tp_trainer* t, int rank, int world, const void* unique_id128) {
    return guarded([&] { TRAINER(t); t->tr->init_data_parallel(rank, world, unique_id128); });
}

int tp_trainer_broadcast_params(tp_trainer* t, int root) {
    return guarded([&] { TRAINER(t); t->tr->broadcast_parameters(root); });
}

int tp_trainer_peer_handle(tp_trainer* t, void* out64) {
    return guarded([&] { TRAINER(t); t->tr->peer_exchange_handle(out64); });
}

int tp_trainer_peer_connect(tp_trainer* t, const void* handles_world_x_64) {
    return guarded([&] { TRAINER(t); t->tr->peer_exchange_connect(handles_world_x_64); });
}

int tp_trainer_set_use_fused(tp_trainer* t, int on) {
    return guarded([&] { TRAINER(t); t->tr->set_use_fused(on != 0); });
}

int tp_trainer_fused_steps(tp_trainer* t, uint64_t* count) {
    return guarded([&] { TRAINER(t); if (count) *count = t->tr->fused_steps(); });
}

int tp_trainer_graph_replays(tp_trainer* t, uint64_t* count) {
    return guarded([&] { TRAINER(t); if (count) *count = t->tr->graph_replays(); });
}

// ---- data path, epoch-level trainer calls, schedulers ----------------------------------------------------------------------
int tp_dataset_from_arrays(const void* images, int is_u8, const float* labels, size_t n, size_t cols, tp_dataset** out) {
    return guarded([&] {
        if (!out) panic("tp_dataset_from_arrays: NULL out pointer");
        auto* d = new tp_dataset();
        d->ds = std::make_shared<data::MNISTDataset>(data::MNISTDataset::from_arrays(images, is_u8 != 0, labels, n, cols));
        *out = d;
    });
}

int tp_dataset_load_mnist(const char* dir, int train, tp_dataset** out) {
    return guarded([&] {
        if (!dir || !out) panic("tp_dataset_load_mnist: NULL argument");
        auto* d = new tp_dataset();
        d->ds = std::make_shared<data::MNISTDataset>(train != 0, std::string(dir));
        *out = d;
    });
}

int tp_dataset_len(tp_dataset* d, size_t* n) {
    return guarded([&] { if (!d || !n) panic("tp_dataset_len: NULL argument"); *n = d->ds->len(); });
}

int tp_dataset_destroy(tp_dataset* d) {
    return guarded([&] { delete d; });
}

int tp_loader_create(tp_dataset* d, size_t batch_size, int shuffle, uint64_t seed, tp_loader** out) {
    return guarded([&] {
        if (!d || !out) panic("tp_loader_create: NULL argument");
        auto* l = new tp_loader();
        l->ld.reset(new data::DataLoader(d->ds, batch_size, shuffle != 0, seed));
        *out = l;
    });
}

int tp_loader_set_sample_shape(tp_loader* l, const size_t* sample_shape, int ndim) {
    return guarded([&] {
        if (!l || !sample_shape || ndim < 1) panic("tp_loader_set_sample_shape: bad arguments");
        Shape s = to_shape(sample_shape, ndim);
        if (shape_numel(s) != l->ld->dataset().cols) panic("tp_loader_set_sample_shape: %zu elements per sample, the dataset has %zu", shape_numel(s), l->ld->dataset().cols);
        l->ld->sample_shape = s;
    });
}

int tp_loader_num_batches(tp_loader* l, size_t* count) {
    return guarded([&] { if (!l || !count) panic("tp_loader_num_batches: NULL argument"); *count = l->ld->num_batches(); });
}

int tp_loader_destroy(tp_loader* l) {
    return guarded([&] { delete l; });
}

int tp_trainer_train_epoch(tp_trainer* t, tp_loader* l, size_t max_batches, float* loss, float* acc) {
    return guarded([&] {
        TRAINER(t);
        if (!l) panic("tp_trainer_train_epoch: NULL loader");
        auto r = t->tr->train_epoch(*l->ld, max_batches);
        if (loss) *loss = r.first;
        if (acc) *acc = r.second;
    });
}

int tp_trainer_evaluate(tp_trainer* t, tp_loader* l, float* loss, float* acc) {
    return guarded([&] {
        TRAINER(t);
        if (!l) panic("tp_trainer_evaluate: NULL loader");
        auto r = t->tr->evaluate(*l->ld);
        if (loss) *loss = r.first;
        if (acc) *acc = r.second;
    });
}

int tp_scheduler_create(const char* kind, float base_lr, float p1, float p2, size_t n, const char* mode, tp_scheduler** out) {
    return guarded([&] {
        if (!kind || !out) panic("tp_scheduler_create: NULL argument");
        std::string k = kind;
        auto* s = new tp_scheduler();
        if (k == "step") s->sc = std::make_shared<optim::StepLR>(base_lr, n, p1);
        else if (k == "exponential") s->sc = std::make_shared<optim::ExponentialLR>(base_lr, p1);
        else if (k == "cosine") s->sc = std::make_shared<optim::CosineAnnealingLR>(base_lr, n, p1);
        else if (k == "plateau")
            s->sc = std::make_shared<optim::ReduceLROnPlateau>(base_lr, p1, n, p2, mode ? std::optional<std::string>(mode) : std::nullopt);
        else { delete s; panic("tp_scheduler_create: unknown scheduler '%s'", kind); }
        *out = s;
    });
}

int tp_scheduler_step(tp_scheduler* s, int has_metric, float metric) {
    return guarded([&] {
        if (!s) panic("tp_scheduler_step: NULL scheduler");
        s->sc->step(has_metric ? std::optional<float>(metric) : std::nullopt);
    });
}

int tp_scheduler_get_lr(tp_scheduler* s, float* lr) {
    return guarded([&] { if (!s || !lr) panic("tp_scheduler_get_lr: NULL argument"); *lr = s->sc->get_lr(); });
}

int tp_scheduler_destroy(tp_scheduler* s) {
    return guarded([&] { delete s; });
}

int tp_trainer_set_scheduler(tp_trainer* t, tp_scheduler* s) {
    return guarded([&] { TRAINER(t); t->tr->scheduler = s ? s->sc : nullptr; });
}

int tp_trainer_fit(tp_trainer* t, tp_loader* train_loader, tp_loader* val_loader, size_t epochs, int verbose) {
    return guarded([&] {
        TRAINER(t);
        if (!train_loader || !val_loader) panic("tp_trainer_fit: NULL loader");
        t->tr->fit(*train_loader->ld, *val_loader->ld, epochs, verbose != 0);
    });
}

int tp_trainer_metrics(tp_trainer* t, int which, float* out, size_t cap, size_t* count) {
    return guarded([&] {
        TRAINER(t);
        const train::Metrics& m = t->tr->metrics;
        const std::vector<float>* v = which == 0 ? &m.train_loss : which == 1 ? &m.train_acc : which == 2 ? &m.val_loss
                                    : which == 3 ? &m.val_acc : which == 4 ? &m.epoch_times : nullptr;
        if (!v) panic("tp_trainer_metrics: which must be 0..4");
        if (count) *count = v->size();
        if (out) for (size_t i = 0; i < v->size() && i < cap; ++i) out[i] = (*v)[i];
    });
}

int tp_trainer_get_lr(tp_trainer* t, float* lr) {
    return guarded([&] { TRAINER(t); if (lr) *lr = t->tr->optimizer->lr(); });
}

// ---- Tensor / Tape / Module / loss / optimizer handles -----------------------------------------------------------------------
#define TENSOR(x) do { if (!(x) || !(x)->t.defined()) panic("%s: NULL tensor", __func__); } while (0)

static tp_tensor* wrap(const Tensor& t) {
    auto* h = new tp_tensor();
    h->t = t;
    return h;
}

int tp_tensor_new(const float* data, const size_t* shape, int ndim, int requires_grad, tp_tensor** out) {
    return guarded([&] {
        if (!data || !shape || ndim < 1 || !out) panic("tp_tensor_new: bad arguments");
        Tensor t = Tensor::from_host(data, to_shape(shape, ndim));
        *out = wrap(requires_grad ? t.requires_grad() : t);
    });
}

int tp_tensor_clone(tp_tensor* t, tp_tensor** out) {
    return guarded([&] { TENSOR(t); if (!out) panic("tp_tensor_clone: NULL out pointer"); *out = wrap(t->t); });
}

int tp_tensor_free(tp_tensor* t) {
    return guarded([&] { delete t; });
}

int tp_tensor_ndim(tp_tensor* t, int* ndim) {
    return guarded([&] { TENSOR(t); if (ndim) *ndim = (int)t->t.shape().size(); });
}

int tp_tensor_shape(tp_tensor* t, size_t* dims, int cap) {
    return guarded([&] {
        TENSOR(t);
        const Shape& s = t->t.shape();
        if (!dims || cap < (int)s.size()) panic("tp_tensor_shape: buffer for %zu dimensions needed", s.size());
        for (size_t i = 0; i < s.size(); ++i) dims[i] = s[i];
    });
}

int tp_tensor_numel(tp_tensor* t, size_t* n) {
    return guarded([&] { TENSOR(t); if (n) *n = t->t.numel(); });
}

int tp_tensor_data(tp_tensor* t, float* out, size_t n) {
    return guarded([&] {
        TENSOR(t);
        if (!out || n != t->t.numel()) panic("tp_tensor_data: expected a buffer of %zu floats", t->t.numel());
        const std::vector<float>& d = t->t.data();
        std::memcpy(out, d.data(), n * sizeof(float));
    });
}

int tp_tensor_set_data(tp_tensor* t, const float* data, size_t n) {
    return guarded([&] {
        TENSOR(t);
        if (!data || n != t->t.numel()) panic("tp_tensor_set_data: expected %zu floats", t->t.numel());
        t->t.set_data(std::vector<float>(data, data + n));
    });
}

int tp_tensor_grad(tp_tensor* t, float* out, size_t n, int* has_grad) {
    return guarded([&] {
        TENSOR(t);
        auto g = t->t.grad();
        if (has_grad) *has_grad = g ? 1 : 0;
        if (g && out) {
            if (n != t->t.numel()) panic("tp_tensor_grad: expected a buffer of %zu floats", t->t.numel());
            std::memcpy(out, g->data().data(), n * sizeof(float));
        }
    });
}

int tp_tensor_requires_grad(tp_tensor* t, int* flag) {
    return guarded([&] { TENSOR(t); if (flag) *flag = t->t.needs_grad() ? 1 : 0; });
}

int tp_tensor_zero_grad(tp_tensor* t) {
    return guarded([&] { TENSOR(t); t->t.zero_grad(); });
}

int tp_tensor_backward(tp_tensor* t) {
    return guarded([&] { TENSOR(t); t->t.backward(); });
}

int tp_tensor_unary(const char* op, tp_tensor* x, float arg, tp_tensor** out) {
    return guarded([&] {
        TENSOR(x);
        if (!op || !out) panic("tp_tensor_unary: NULL argument");
        std::string k = op;
        const Tensor& t = x->t;
        if (k == "relu") *out = wrap(t.relu());
        else if (k == "exp") *out = wrap(t.exp());
        else if (k == "log") *out = wrap(t.log());
        else if (k == "sigmoid") *out = wrap(t.sigmoid());
        else if (k == "mean") *out = wrap(t.mean());
        else if (k == "transpose") *out = wrap(t.transpose());
        else if (k == "pow") *out = wrap(t.pow(arg));
        else if (k == "sqrt") *out = wrap(t.sqrt());
        else panic("tp_tensor_unary: unknown op '%s'", op);
    });
}

int tp_tensor_binary(const char* op, tp_tensor* a, tp_tensor* b, tp_tensor** out) {
    return guarded([&] {
        TENSOR(a); TENSOR(b);
        if (!op || !out) panic("tp_tensor_binary: NULL argument");
        std::string k = op;
        if (k == "add") *out = wrap(a->t + b->t);
        else if (k == "sub") *out = wrap(a->t - b->t);
        else if (k == "mul") *out = wrap(a->t * b->t);
        else if (k == "div") *out = wrap(a->t / b->t);
        else if (k == "matmul") *out = wrap(a->t.matmul(b->t));
        else if (k == "add_broadcast") *out = wrap(a->t.add_broadcast(b->t));
        else if (k == "sub_broadcast_rows") *out = wrap(a->t.sub_broadcast_rows(b->t));
        else panic("tp_tensor_binary: unknown op '%s'", op);
    });
}

int tp_tensor_reshape(tp_tensor* x, const size_t* shape, int ndim, tp_tensor** out) {
    return guarded([&] {
        TENSOR(x);
        if (!shape || ndim < 1 || !out) panic("tp_tensor_reshape: bad arguments");
        *out = wrap(x->t.reshape(to_shape(shape, ndim)));
    });
}

int tp_tensor_flatten(tp_tensor* x, size_t start_dim, tp_tensor** out) {
    return guarded([&] { TENSOR(x); if (!out) panic("tp_tensor_flatten: NULL out pointer"); *out = wrap(x->t.flatten(start_dim)); });
}

int tp_tensor_sum(tp_tensor* x, int dim, int keepdim, tp_tensor** out) {
    return guarded([&] {
        TENSOR(x);
        if (!out) panic("tp_tensor_sum: NULL out pointer");
        *out = wrap(x->t.sum(dim < 0 ? std::nullopt : std::optional<size_t>((size_t)dim), keepdim != 0));
    });
}

int tp_tensor_argmax(tp_tensor* x, int dim, tp_tensor** out) {
    return guarded([&] {
        TENSOR(x);
        if (!out) panic("tp_tensor_argmax: NULL out pointer");
        *out = wrap(x->t.argmax(dim < 0 ? std::nullopt : std::optional<size_t>((size_t)dim)));
    });
}

int tp_tape_reset(void) {
    return guarded([&] { Tape::reset(); });
}

int tp_tape_len(size_t* nodes) {
    return guarded([&] { if (nodes) *nodes = Tape::len(); });
}

int tp_module_forward(tp_model* m, tp_tensor* x, tp_tensor** out) {
    return guarded([&] {
        TENSOR(x);
        if (!m || !out) panic("tp_module_forward: NULL argument");
        *out = wrap(m->seq->forward(x->t));
    });
}

int tp_model_parameter(tp_model* m, int index, tp_tensor** out) {
    return guarded([&] {
        if (!m || !out) panic("tp_model_parameter: NULL argument");
        *out = wrap(param_at(m, index));
    });
}

int tp_loss(const char* kind, tp_tensor* predictions, tp_tensor* targets, tp_tensor** out) {
    return guarded([&] {
        TENSOR(predictions); TENSOR(targets);
        if (!kind || !out) panic("tp_loss: NULL argument");
        std::string k = kind;
        if (k == "cross_entropy") *out = wrap(loss::cross_entropy_loss(predictions->t, targets->t));
        else if (k == "cross_entropy_onehot") *out = wrap(loss::cross_entropy_loss_onehot(predictions->t, targets->t));
        else if (k == "bce") *out = wrap(loss::bce_loss(predictions->t, targets->t));
        else if (k == "mse") *out = wrap(loss::mse_loss(predictions->t, targets->t));
        else panic("tp_loss: unknown loss '%s'", kind);
    });
}

int tp_accuracy(tp_tensor* predictions, tp_tensor* targets, float* acc) {
    return guarded([&] { TENSOR(predictions); TENSOR(targets); if (acc) *acc = loss::accuracy(predictions->t, targets->t); });
}

int tp_optimizer_create(const char* kind, tp_tensor* const* params, int n_params, float lr, float beta1, float beta2, float eps,
                        float weight_decay, tp_optimizer** out) {
    return guarded([&] {
        if (!kind || !params || n_params < 1 || !out) panic("tp_optimizer_create: bad arguments");
        std::vector<Tensor> ps;
        for (int i = 0; i < n_params; ++i) { TENSOR(params[i]); ps.push_back(params[i]->t); }
        std::string k = kind;
        auto* o = new tp_optimizer();
        if (k == "sgd") o->opt = std::make_shared<optim::SGD>(ps, lr);
        else if (k == "adam") o->opt = std::make_shared<optim::Adam>(ps, lr, std::make_pair(beta1, beta2), eps, weight_decay);
        else if (k == "adamw") o->opt = std::make_shared<optim::AdamW>(ps, lr, std::make_pair(beta1, beta2), eps, weight_decay);
        else { delete o; panic("tp_optimizer_create: unknown optimizer '%s'", kind); }
        *out = o;
    });
}

#define OPTIMIZER(o) do { if (!(o) || !(o)->opt) panic("%s: NULL optimizer", __func__); } while (0)

int tp_optimizer_step(tp_optimizer* o) {
    return guarded([&] { OPTIMIZER(o); o->opt->step(); });
}

int tp_optimizer_zero_grad(tp_optimizer* o) {
    return guarded([&] { OPTIMIZER(o); o->opt->zero_grad(); });
}

int tp_optimizer_set_lr(tp_optimizer* o, float lr) {
    return guarded([&] { OPTIMIZER(o); o->opt->set_lr(lr); });
}

int tp_optimizer_get_lr(tp_optimizer* o, float* lr) {
    return guarded([&] { OPTIMIZER(o); if (lr) *lr = o->opt->lr(); });
}

int tp_optimizer_destroy(tp_optimizer* o) {
    return guarded([&] { delete o; });
}

int tp_trainer_device_error(tp_trainer* t, int* code) {
    return guarded([&] {
        TRAINER(t);
        int c = 0;
        check(tp_ctx_device_error(ctx(), &c));
        if (code) *code = c;
    });
}

}  // extern "C"

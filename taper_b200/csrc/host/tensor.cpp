// Host layer, part 1: device context, Tensor, Tape and the tape-recorded ops.
// Mirrors src/tensor.rs, src/ops.rs and src/tape.rs of the reference one-to-one; every op enqueues
// kernels on this thread's stream through the C ABI (include/taper_b200.h) and records the same
// backward closure the reference records.  No arithmetic happens on the host.
#include "taper_internal.hpp"

#include <cstdarg>
#include <cstdio>
#include <random>

namespace taper {

// ---- thread-local state -------------------------------------------------------------------------
namespace {
struct ThreadState {
    tp_ctx* ctx = nullptr;
    int device = 0;
    std::vector<std::function<void()>> tape;      // src/tape.rs:16-23
};
thread_local ThreadState g_ts;
}  // namespace

void check(int rc) {
    if (rc != TP_OK) throw std::runtime_error(std::string("taper_b200: ") + tp_last_error());
}

[[noreturn]] void panic(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw std::runtime_error(buf);
}

void set_device(int device) {
    if (g_ts.ctx) {
        if (g_ts.device == device) return;
        panic("set_device(%d): this thread already owns a context on device %d", device, g_ts.device);
    }
    g_ts.device = device;
}

tp_ctx* ctx() {
    if (!g_ts.ctx) check(tp_ctx_create(g_ts.device, &g_ts.ctx));
    return g_ts.ctx;
}

void synchronize() { check(tp_sync(ctx())); }

bool& Config::conv_full_adjoint() { static thread_local bool v = false; return v; }
bool& Config::fuse_linear_relu() { static thread_local bool v = true; return v; }
bool& Config::fuse_conv_stack() { static thread_local bool v = true; return v; }
bool& Config::fuse_small_mlp() { static thread_local bool v = true; return v; }
bool& Config::reference_op_sequence() { static thread_local bool v = false; return v; }

// ---- TensorImpl -----------------------------------------------------------------------------------
TensorImpl::~TensorImpl() {
    if (buf) tp_buf_release(buf);
    if (grad) tp_buf_release(grad);
}

tp_buf* TensorImpl::grad_for_write(int* accumulate) {
    // `if g.is_none() { *g = Some(vec![0.0; n]) }` then `+=`  (e.g. src/ops.rs:126-128, 250-253):
    // the first write after zero_grad stores (accumulate = 0), later writes add (accumulate = 1).
    if (!grad) check(tp_buf_alloc(ctx(), n ? n : 1, &grad));
    *accumulate = has_grad ? 1 : 0;
    has_grad = true;
    grad_version++;
    return grad;
}

size_t shape_numel(const Shape& s) {
    size_t n = 1;
    for (size_t d : s) n *= d;
    return n;
}

static std::shared_ptr<TensorImpl> make_impl(const Shape& shape) {
    auto im = std::make_shared<TensorImpl>();
    im->shape = shape;
    im->n = shape_numel(shape);
    check(tp_buf_alloc(ctx(), im->n ? im->n : 1, &im->buf));
    return im;
}

Tensor Tensor::empty(const Shape& shape) {
    Tensor t;
    t.impl_ = make_impl(shape);
    return t;
}

Tensor Tensor::from_host(const float* data, const Shape& shape) {
    Tensor t = empty(shape);
    if (t.impl_->n) check(tp_buf_upload(ctx(), t.impl_->buf, data, t.impl_->n));
    return t;
}

Tensor Tensor::create(const std::vector<float>& data, const Shape& shape) {
    if (data.size() != shape_numel(shape)) panic("Tensor::new: data length %zu does not match shape", data.size());
    return from_host(data.data(), shape);
}

Tensor Tensor::scalar(float v) { return create({v}, {1}); }

Tensor Tensor::zeros(const Shape& shape) {
    Tensor t = empty(shape);
    check(tp_buf_fill(ctx(), t.impl_->buf, 0.0f, t.impl_->n));
    return t;
}

Tensor Tensor::randn(const Shape& shape, uint64_t seed) {
    std::mt19937_64 rng(seed);
    std::normal_distribution<float> d(0.0f, 1.0f);
    std::vector<float> v(shape_numel(shape));
    for (auto& x : v) x = d(rng);
    return create(v, shape);
}

Tensor Tensor::adopt(tp_buf* buf, const Shape& shape) {
    Tensor t;
    t.impl_ = std::make_shared<TensorImpl>();
    t.impl_->shape = shape;
    t.impl_->n = shape_numel(shape);
    t.impl_->buf = buf;
    return t;
}

Tensor Tensor::requires_grad() const {
    Tensor t = *this;
    t.requires_grad_ = true;
    return t;
}

const Shape& Tensor::shape() const { return impl_->shape; }
size_t Tensor::numel() const { return impl_->n; }
tp_buf* Tensor::buf() const { return impl_->buf; }
tp_buf* Tensor::grad_buf() const { return impl_->has_grad ? impl_->grad : nullptr; }

const std::vector<float>& Tensor::data() const {
    TensorImpl& im = *impl_;
    if (im.host_version != im.version) {
        im.host.resize(im.n);
        if (im.n) check(tp_buf_download(ctx(), im.buf, im.host.data(), im.n));
        im.host_version = im.version;
    }
    return im.host;
}

float Tensor::item() const { return data().at(0); }

void Tensor::set_data(const std::vector<float>& v) const {
    if (v.size() != impl_->n) panic("set_data: length mismatch");
    if (impl_->n) check(tp_buf_upload(ctx(), impl_->buf, v.data(), impl_->n));
    impl_->version++;
}

std::optional<Tensor> Tensor::grad() const {
    if (!impl_->has_grad) return std::nullopt;
    Tensor g = empty(impl_->shape);
    check(tp_buf_copy(ctx(), g.impl_->buf, impl_->grad, impl_->n));
    return g;
}

void Tensor::set_grad(const std::vector<float>& g) const {
    if (g.size() != impl_->n) panic("set_grad: length mismatch");
    int acc;
    tp_buf* gb = impl_->grad_for_write(&acc);
    if (impl_->n) check(tp_buf_upload(ctx(), gb, g.data(), impl_->n));
}

void Tensor::zero_grad() const { impl_->has_grad = false; }      // grad := None  (src/tensor.rs:531-533)

void Tensor::backward() const {
    // grad := ones(len); node 0 is "no node" in the reference (src/tensor.rs:520-529).  Nodes are
    // stamped id+1 here, so the first recorded op can be a root too (deliberate fix).
    int acc;
    tp_buf* g = impl_->grad_for_write(&acc);
    check(tp_buf_fill(ctx(), g, 1.0f, impl_->n));
    size_t node = impl_->tape_node;
    if (node != 0) taper::backward(node - 1);
}

// ---- Tape  (src/tape.rs) ---------------------------------------------------------------------------
void Tape::ensure_active() {}
void Tape::reset() { g_ts.tape.clear(); }
size_t Tape::len() { return g_ts.tape.size(); }

static void tape_push(const Tensor& out, std::function<void()> fn) {
    g_ts.tape.push_back(std::move(fn));
    out.impl()->tape_node = g_ts.tape.size();                  // index + 1
}

void Tape::push_binary_op(const Tensor& a, const Tensor& b, const Tensor& out, std::function<void()> fn) {
    if (!(a.needs_grad() || b.needs_grad())) return;           // src/tape.rs:55-57
    tape_push(out, std::move(fn));
}

void Tape::push_unary_op(const Tensor& input, const Tensor& out, std::function<void()> fn) {
    if (!input.needs_grad()) return;                           // src/tape.rs:82-84
    tape_push(out, std::move(fn));
}

std::vector<std::function<void()>> Tape::take() {
    std::vector<std::function<void()>> t;
    t.swap(g_ts.tape);
    return t;
}

void backward(size_t final_node_id) {
    // Clone the closure list first: a closure may record new nodes while it runs (src/tape.rs:108-126)
    auto& tape = g_ts.tape;
    if (tape.empty()) return;
    size_t end = std::min(final_node_id, tape.size() - 1);
    std::vector<std::function<void()>> fns(tape.begin(), tape.begin() + end + 1);
    for (size_t i = fns.size(); i-- > 0;) fns[i]();
}

// ---- elementwise operators  (src/ops.rs:8-120, 377-496) --------------------------------------------
namespace {

enum class Bin { Add, Sub, Mul, Div };

Tensor binary(const Tensor& a, const Tensor& b, Bin op) {
    if (a.numel() != b.numel()) panic("Tensor dimensions must match for elementwise op (%zu vs %zu)", a.numel(), b.numel());
    size_t n = a.numel();
    Tensor out = Tensor::empty(a.shape());
    tp_ctx* c = ctx();
    switch (op) {
        case Bin::Add: check(tp_add(c, a.buf(), b.buf(), out.buf(), n)); break;
        case Bin::Sub: check(tp_sub(c, a.buf(), b.buf(), out.buf(), n)); break;
        case Bin::Mul: check(tp_mul(c, a.buf(), b.buf(), out.buf(), n)); break;
        case Bin::Div: check(tp_div(c, a.buf(), b.buf(), out.buf(), n)); break;
    }
    if (a.needs_grad() || b.needs_grad()) {
        out.set_requires_grad(true);
        Tape::push_binary_op(a, b, out, [a, b, out, op, n]() {
            tp_buf* g = out.grad_buf();
            if (!g) return;                                    // `if let Some(gout)`
            tp_ctx* c = ctx();
            int acc;
            if (a.needs_grad()) {
                tp_buf* ga = a.impl()->grad_for_write(&acc);
                switch (op) {
                    case Bin::Add: case Bin::Sub: check(tp_accumulate(c, ga, g, 1.0f, n, acc)); break;
                    case Bin::Mul: check(tp_mul_bwd(c, g, b.buf(), ga, n, acc)); break;
                    case Bin::Div: check(tp_div_bwd_a(c, g, b.buf(), ga, n, acc)); break;
                }
            }
            if (b.needs_grad()) {
                tp_buf* gb = b.impl()->grad_for_write(&acc);
                switch (op) {
                    case Bin::Add: check(tp_accumulate(c, gb, g, 1.0f, n, acc)); break;
                    case Bin::Sub: check(tp_accumulate(c, gb, g, -1.0f, n, acc)); break;
                    case Bin::Mul: check(tp_mul_bwd(c, g, a.buf(), gb, n, acc)); break;
                    case Bin::Div: check(tp_div_bwd_b(c, g, a.buf(), b.buf(), gb, n, acc)); break;
                }
            }
        });
    }
    return out;
}

}  // namespace

Tensor operator+(const Tensor& a, const Tensor& b) { return binary(a, b, Bin::Add); }
Tensor operator-(const Tensor& a, const Tensor& b) { return binary(a, b, Bin::Sub); }
Tensor operator*(const Tensor& a, const Tensor& b) { return binary(a, b, Bin::Mul); }
Tensor operator/(const Tensor& a, const Tensor& b) { return binary(a, b, Bin::Div); }

// ---- matmul  (src/ops.rs:200-298) ---------------------------------------------------------------------
Tensor Tensor::matmul(const Tensor& other) const {
    if (shape().size() != 2 || other.shape().size() != 2) panic("matmul: both operands must be 2-D");
    int m = (int)shape()[0], k = (int)shape()[1], k2 = (int)other.shape()[0], n = (int)other.shape()[1];
    if (k != k2) panic("matmul: inner dimensions must match (%d vs %d)", k, k2);
    Tensor out = Tensor::empty({(size_t)m, (size_t)n});
    check(tp_sgemm_rowmajor(ctx(), 0, 0, m, n, k, 1.0f, buf(), other.buf(), 0.0f, out.buf()));     // :215-226
    if (needs_grad() || other.needs_grad()) {
        out.set_requires_grad(true);
        Tensor a = *this, b = other;
        Tape::push_binary_op(a, b, out, [a, b, out, m, n, k]() {
            tp_buf* g = out.grad_buf();
            if (!g) return;
            int acc;
            if (a.needs_grad()) {                              // dA += dC * B^T  (N,T)  :254-265
                tp_buf* ga = a.impl()->grad_for_write(&acc);
                check(tp_sgemm_rowmajor(ctx(), 0, 1, m, k, n, 1.0f, g, b.buf(), acc ? 1.0f : 0.0f, ga));
            }
            if (b.needs_grad()) {                              // dB += A^T * dC  (T,N)  :280-291
                tp_buf* gb = b.impl()->grad_for_write(&acc);
                check(tp_sgemm_rowmajor(ctx(), 1, 0, k, n, m, 1.0f, a.buf(), g, acc ? 1.0f : 0.0f, gb));
            }
        });
    }
    return out;
}

// ---- relu  (src/ops.rs:312-374) --------------------------------------------------------------------------
Tensor Tensor::relu() const {
    Tensor out = Tensor::empty(shape());
    check(tp_relu_fwd(ctx(), buf(), out.buf(), numel()));
    if (needs_grad()) {
        out.set_requires_grad(true);
        Tensor x = *this;
        Tape::push_unary_op(x, out, [x, out]() {
            tp_buf* g = out.grad_buf();
            if (!g) return;
            int acc;
            tp_buf* gin = x.impl()->grad_for_write(&acc);
            check(tp_relu_bwd(ctx(), x.buf(), g, gin, x.numel(), acc));
        });
    }
    return out;
}

// ---- transpose  (src/tensor.rs:544-591) -----------------------------------------------------------------
Tensor Tensor::transpose() const {
    if (shape().size() != 2) panic("Can only transpose 2D tensors");
    int r = (int)shape()[0], c = (int)shape()[1];
    Tensor out = Tensor::empty({(size_t)c, (size_t)r});
    check(tp_transpose2d(ctx(), buf(), out.buf(), r, c, 0));
    if (needs_grad()) {
        out.set_requires_grad(true);
        Tensor x = *this;
        Tape::push_unary_op(x, out, [x, out, r, c]() {
            tp_buf* g = out.grad_buf();
            if (!g) return;
            int acc;
            tp_buf* gin = x.impl()->grad_for_write(&acc);
            check(tp_transpose2d(ctx(), g, gin, c, r, acc));   // gin[i,j] += g[j,i]  :575-586
        });
    }
    return out;
}

// ---- add_broadcast  (src/tensor.rs:636-704) ---------------------------------------------------------------
Tensor Tensor::add_broadcast(const Tensor& other) const {
    if (shape() == other.shape()) return *this + other;
    if (shape().size() != 2 || other.shape().size() != 1) panic("Unsupported broadcasting shapes");
    int rows = (int)shape()[0], cols = (int)shape()[1];
    if ((size_t)cols != other.shape()[0]) panic("Last dimension must match for broadcasting");
    Tensor out = Tensor::empty(shape());
    check(tp_add_broadcast_fwd(ctx(), buf(), other.buf(), out.buf(), rows, cols, 0));
    if (needs_grad() || other.needs_grad()) {
        out.set_requires_grad(true);
        Tensor a = *this, b = other;
        Tape::push_binary_op(a, b, out, [a, b, out, rows, cols]() {
            tp_buf* g = out.grad_buf();
            if (!g) return;
            int acc;
            if (a.needs_grad()) {
                tp_buf* ga = a.impl()->grad_for_write(&acc);
                check(tp_accumulate(ctx(), ga, g, 1.0f, a.numel(), acc));
            }
            if (b.needs_grad()) {                              // db[f] = sum_i g[i,f]  :680-691
                tp_buf* gb = b.impl()->grad_for_write(&acc);
                check(tp_colsum(ctx(), g, gb, rows, cols, 1.0f, acc));
            }
        });
    }
    return out;
}

// ---- sub_broadcast_rows  (src/tensor.rs:707-770) ------------------------------------------------------------
Tensor Tensor::sub_broadcast_rows(const Tensor& other) const {
    if (shape() == other.shape()) return *this - other;
    if (shape().size() != 2 || other.shape().size() != 2 || other.shape()[0] != shape()[0] || other.shape()[1] != 1)
        panic("sub_broadcast_rows expects [B,C] - [B,1]");
    int rows = (int)shape()[0], cols = (int)shape()[1];
    Tensor out = Tensor::empty(shape());
    check(tp_sub_broadcast_rows_fwd(ctx(), buf(), other.buf(), out.buf(), rows, cols));
    if (needs_grad() || other.needs_grad()) {
        out.set_requires_grad(true);
        Tensor a = *this, r = other;
        Tape::push_binary_op(a, r, out, [a, r, out, rows, cols]() {
            tp_buf* g = out.grad_buf();
            if (!g) return;
            int acc;
            if (a.needs_grad()) {
                tp_buf* ga = a.impl()->grad_for_write(&acc);
                check(tp_accumulate(ctx(), ga, g, 1.0f, a.numel(), acc));
            }
            if (r.needs_grad()) {                              // gr[i] += -sum_c g[i,c]  :748-761
                tp_buf* gr = r.impl()->grad_for_write(&acc);
                check(tp_rowsum(ctx(), g, gr, rows, cols, -1.0f, acc));
            }
        });
    }
    return out;
}

// ---- reshape / flatten  (src/tensor.rs:803-858) --------------------------------------------------------------
Tensor Tensor::reshape(const Shape& new_shape) const {
    if (shape_numel(new_shape) != numel()) panic("Cannot reshape tensor of %zu elements", numel());
    // The reference clones the data (A11); a view of the same refcounted buffer is observationally
    // identical because op outputs are never mutated in place.  The gradient still flows through a node.
    check(tp_buf_retain(buf()));
    Tensor out = Tensor::adopt(buf(), new_shape);
    if (needs_grad()) {
        out.set_requires_grad(true);
        Tensor x = *this;
        Tape::push_unary_op(x, out, [x, out]() {
            tp_buf* g = out.grad_buf();
            if (!g) return;
            int acc;
            tp_buf* gin = x.impl()->grad_for_write(&acc);
            check(tp_accumulate(ctx(), gin, g, 1.0f, x.numel(), acc));      // :822-832
        });
    }
    return out;
}

Tensor Tensor::flatten(size_t start_dim) const {
    if (start_dim >= shape().size()) panic("start_dim out of bounds");
    Shape s(shape().begin(), shape().begin() + start_dim);
    size_t rest = 1;
    for (size_t i = start_dim; i < shape().size(); ++i) rest *= shape()[i];
    s.push_back(rest);
    return reshape(s);
}

// ---- sum  (src/tensor.rs:890-1018) ------------------------------------------------------------------------------
Tensor Tensor::sum(std::optional<size_t> dim, bool keepdim) const {
    Tensor x = *this;
    if (!dim.has_value() || (shape().size() == 1 && *dim == 0)) {
        Tensor out = Tensor::empty({1});
        check(tp_sum_all(ctx(), buf(), out.buf(), numel()));
        if (needs_grad()) {
            out.set_requires_grad(true);
            Tape::push_unary_op(x, out, [x, out]() {
                tp_buf* g = out.grad_buf();
                if (!g) return;
                int acc;
                tp_buf* gin = x.impl()->grad_for_write(&acc);
                check(tp_broadcast_bwd(ctx(), g, gin, 1, (int)x.numel(), 2, acc));     // :1003-1010
            });
        }
        return out;
    }
    size_t d = *dim;
    if (d >= shape().size()) panic("sum: dim out of range");
    // view as [outer, shape[d], inner]; the kernels cover the first and the last dimension
    size_t outer = 1, inner = 1;
    for (size_t i = 0; i < d; ++i) outer *= shape()[i];
    for (size_t i = d + 1; i < shape().size(); ++i) inner *= shape()[i];
    Shape os;
    for (size_t i = 0; i < shape().size(); ++i) {
        if (i == d) { if (keepdim) os.push_back(1); }
        else os.push_back(shape()[i]);
    }
    if (os.empty()) os.push_back(1);
    Tensor out = Tensor::empty(os);
    int rows, cols, mode;
    if (inner == 1) { rows = (int)outer; cols = (int)shape()[d]; mode = 0; check(tp_rowsum(ctx(), buf(), out.buf(), rows, cols, 1.0f, 0)); }
    else if (outer == 1) { rows = (int)shape()[d]; cols = (int)inner; mode = 1; check(tp_colsum(ctx(), buf(), out.buf(), rows, cols, 1.0f, 0)); }
    else panic("sum over a middle dimension is not on the hot path (TP_ERR_UNSUPPORTED)");
    if (needs_grad()) {
        out.set_requires_grad(true);
        Tape::push_unary_op(x, out, [x, out, rows, cols, mode]() {
            tp_buf* g = out.grad_buf();
            if (!g) return;
            int acc;
            tp_buf* gin = x.impl()->grad_for_write(&acc);
            check(tp_broadcast_bwd(ctx(), g, gin, rows, cols, mode, acc));            // :942-990
        });
    }
    return out;
}

// ---- max / argmax  (src/tensor.rs:1021-1088); no tape node ---------------------------------------------------------
std::pair<Tensor, Tensor> Tensor::max(std::optional<size_t> dim) const {
    if (!dim.has_value()) {
        Tensor v = Tensor::empty({1}), i = Tensor::empty({1});
        check(tp_max_all(ctx(), buf(), v.buf(), i.buf(), numel()));
        return {v, i};
    }
    if (shape().size() == 1 && *dim == 0) {
        Tensor v = Tensor::empty({1}), i = Tensor::empty({1});
        check(tp_max_rows(ctx(), buf(), v.buf(), i.buf(), 1, (int)numel()));
        return {v, i};
    }
    if (shape().size() != 2 || *dim > 1) panic("max(dim) is defined for 1-D/2-D tensors only (SURVEY A7)");
    int rows = (int)shape()[0], cols = (int)shape()[1];
    if (*dim == 1) {
        Tensor v = Tensor::empty({(size_t)rows, 1}), i = Tensor::empty({(size_t)rows, 1});
        check(tp_max_rows(ctx(), buf(), v.buf(), i.buf(), rows, cols));
        return {v, i};
    }
    Tensor v = Tensor::empty({1, (size_t)cols}), i = Tensor::empty({1, (size_t)cols});
    check(tp_max_cols(ctx(), buf(), v.buf(), i.buf(), rows, cols));
    return {v, i};
}

Tensor Tensor::argmax(std::optional<size_t> dim) const { return max(dim).second; }

// ---- exp / log  (src/tensor.rs:1091-1169) ----------------------------------------------------------------------------
Tensor Tensor::exp() const {
    Tensor out = Tensor::empty(shape());
    check(tp_exp_fwd(ctx(), buf(), out.buf(), numel()));
    if (needs_grad()) {
        out.set_requires_grad(true);
        Tensor x = *this;
        Tape::push_unary_op(x, out, [x, out]() {
            tp_buf* g = out.grad_buf();
            if (!g) return;
            int acc;
            tp_buf* gin = x.impl()->grad_for_write(&acc);
            check(tp_exp_bwd(ctx(), out.buf(), g, gin, x.numel(), acc));
        });
    }
    return out;
}

// ---- sigmoid / mean / pow  (src/tensor.rs:594-634, 772-800, 1172-1211) ------------------------------------------
Tensor Tensor::sigmoid() const {
    Tensor out = Tensor::empty(shape());
    check(tp_sigmoid_fwd(ctx(), buf(), out.buf(), numel()));
    if (needs_grad()) {
        out.set_requires_grad(true);
        Tensor x = *this;
        Tape::push_unary_op(x, out, [x, out]() {
            tp_buf* g = out.grad_buf();
            if (!g) return;
            int acc;
            tp_buf* gin = x.impl()->grad_for_write(&acc);
            check(tp_sigmoid_bwd(ctx(), out.buf(), g, gin, x.numel(), acc));     // gin += g * s * (1 - s)
        });
    }
    return out;
}

Tensor Tensor::mean() const {
    if (numel() == 0) panic("mean of an empty tensor");
    Tensor out = Tensor::empty({1});
    check(tp_mean_fwd(ctx(), buf(), out.buf(), numel()));
    if (needs_grad()) {
        out.set_requires_grad(true);
        Tensor x = *this;
        Tape::push_unary_op(x, out, [x, out]() {
            tp_buf* g = out.grad_buf();
            if (!g) return;
            int acc;
            tp_buf* gin = x.impl()->grad_for_write(&acc);
            check(tp_mean_bwd(ctx(), g, gin, x.numel(), acc));                   // gin += g[0] / n
        });
    }
    return out;
}

Tensor Tensor::pow(float exponent) const {
    Tensor out = Tensor::empty(shape());
    check(tp_pow_fwd(ctx(), buf(), out.buf(), exponent, numel()));
    if (needs_grad()) {
        out.set_requires_grad(true);
        Tensor x = *this;
        Tape::push_unary_op(x, out, [x, out, exponent]() {
            tp_buf* g = out.grad_buf();
            if (!g) return;
            int acc;
            tp_buf* gin = x.impl()->grad_for_write(&acc);
            check(tp_pow_bwd(ctx(), x.buf(), g, gin, exponent, x.numel(), acc)); // gin += g * n * x^(n-1)
        });
    }
    return out;
}

Tensor Tensor::log() const {
    Tensor out = Tensor::empty(shape());
    check(tp_log_fwd(ctx(), buf(), out.buf(), numel()));
    if (needs_grad()) {
        out.set_requires_grad(true);
        Tensor x = *this;
        Tape::push_unary_op(x, out, [x, out]() {
            tp_buf* g = out.grad_buf();
            if (!g) return;
            int acc;
            tp_buf* gin = x.impl()->grad_for_write(&acc);
            check(tp_log_bwd(ctx(), x.buf(), g, gin, x.numel(), acc));
        });
    }
    return out;
}

// ---- fused Linear: one node for transpose + matmul + add_broadcast (+ relu) of src/nn.rs:54-60 ---------------------------
Tensor Tensor::linear(const Tensor& weight, const Tensor* bias, bool relu) const {
    if (shape().size() != 2 || weight.shape().size() != 2) panic("linear: input [B,in] and weight [out,in] must be 2-D");
    int batch = (int)shape()[0], fin = (int)shape()[1], fout = (int)weight.shape()[0];
    if ((size_t)fin != weight.shape()[1]) panic("linear: inner dimensions must match (%d vs %zu)", fin, weight.shape()[1]);
    if (bias && (bias->shape().size() != 1 || bias->shape()[0] != (size_t)fout)) panic("Last dimension must match for broadcasting");
    Tensor out = Tensor::empty({(size_t)batch, (size_t)fout});
    check(tp_linear_fwd(ctx(), buf(), weight.buf(), bias ? bias->buf() : nullptr, out.buf(), batch, fin, fout, relu ? 1 : 0));
    bool any = needs_grad() || weight.needs_grad() || (bias && bias->needs_grad());
    if (any) {
        out.set_requires_grad(true);
        Tensor x = *this, w = weight;
        std::optional<Tensor> b;
        if (bias) b = *bias;
        tape_push(out, [x, w, b, out, batch, fin, fout, relu]() {
            tp_buf* g = out.grad_buf();
            if (!g) return;
            int ax = 0, aw = 0, ab = 0;
            tp_buf* gx = x.needs_grad() ? x.impl()->grad_for_write(&ax) : nullptr;
            tp_buf* gw = w.needs_grad() ? w.impl()->grad_for_write(&aw) : nullptr;
            tp_buf* gb = (b && b->needs_grad()) ? b->impl()->grad_for_write(&ab) : nullptr;
            check(tp_linear_bwd(ctx(), x.buf(), w.buf(), g, relu ? out.buf() : nullptr, gx, gw, gb, batch, fin, fout, ax, aw, ab));
        });
    }
    return out;
}

// ---- a chain of small Linear(+ReLU) layers as one node (tp_mlp_small_fwd / bwd) ---------------------------------------------
Tensor Tensor::mlp_chain(const std::vector<MlpLayer>& layers) const {
    const int L = (int)layers.size();
    if (shape().size() != 2 || L < 1 || L > 4) return Tensor();
    const int batch = (int)shape()[0];
    int dims[5];
    dims[0] = (int)shape()[1];
    for (int l = 0; l < L; ++l) {
        const Tensor& w = layers[l].weight;
        if (w.shape().size() != 2 || (int)w.shape()[1] != dims[l]) return Tensor();
        if (layers[l].bias && (layers[l].bias->shape().size() != 1 || layers[l].bias->shape()[0] != w.shape()[0])) return Tensor();
        dims[l + 1] = (int)w.shape()[0];
    }
    if (!tp_mlp_small_supported(L, dims, batch)) return Tensor();
    const tp_buf* wb[4];
    const tp_buf* bb[4];
    tp_buf* ab[4];
    int re[4];
    std::vector<Tensor> acts;
    bool any = needs_grad();
    for (int l = 0; l < L; ++l) {
        acts.push_back(Tensor::empty({(size_t)batch, (size_t)dims[l + 1]}));
        wb[l] = layers[l].weight.buf();
        bb[l] = layers[l].bias ? layers[l].bias->buf() : nullptr;
        ab[l] = acts[l].buf();
        re[l] = layers[l].relu ? 1 : 0;
        any = any || layers[l].weight.needs_grad() || (layers[l].bias && layers[l].bias->needs_grad());
    }
    check(tp_mlp_small_fwd(ctx(), buf(), L, dims, wb, bb, re, ab, batch));
    Tensor out = acts[L - 1];
    if (any) {
        out.set_requires_grad(true);
        Tensor x = *this;
        std::vector<MlpLayer> ls = layers;
        std::vector<int> dv(dims, dims + L + 1);
        tape_push(out, [x, ls, acts, out, dv, batch]() {
            tp_buf* g = out.grad_buf();
            if (!g) return;
            const int L = (int)ls.size();
            const tp_buf* wb[4];
            const tp_buf* ab[4];
            tp_buf* dw[4];
            tp_buf* db[4];
            int re[4], aw[4] = {0, 0, 0, 0}, abi[4] = {0, 0, 0, 0}, ax = 0;
            for (int l = 0; l < L; ++l) {
                wb[l] = ls[l].weight.buf();
                ab[l] = acts[l].buf();
                re[l] = ls[l].relu ? 1 : 0;
                dw[l] = ls[l].weight.needs_grad() ? ls[l].weight.impl()->grad_for_write(&aw[l]) : nullptr;
                db[l] = (ls[l].bias && ls[l].bias->needs_grad()) ? ls[l].bias->impl()->grad_for_write(&abi[l]) : nullptr;
            }
            tp_buf* gx = x.needs_grad() ? x.impl()->grad_for_write(&ax) : nullptr;
            check(tp_mlp_small_bwd(ctx(), x.buf(), L, dv.data(), wb, re, ab, g, gx, dw, db, ax, aw, abi, batch));
        });
    }
    return out;
}

// ---- conv2d  (src/tensor.rs:1221-1285, 1379-1389) ------------------------------------------------------------------------------
Tensor Tensor::conv2d_impl(const Tensor& weight, const Tensor* bias, Pair stride, Pair padding, Pair dilation, bool relu) const {
    if (shape().size() != 4 || weight.shape().size() != 4) panic("conv2d: input [N,C,H,W] and weight [Co,Ci,kh,kw] must be 4-D");
    if (shape()[1] != weight.shape()[1]) panic("Input and weight channel dimensions must match");
    tp_conv_desc d;
    d.n = (int)shape()[0]; d.c_in = (int)shape()[1]; d.h = (int)shape()[2]; d.w = (int)shape()[3];
    d.c_out = (int)weight.shape()[0]; d.kh = (int)weight.shape()[2]; d.kw = (int)weight.shape()[3];
    d.stride_h = (int)stride.first; d.stride_w = (int)stride.second;
    d.pad_h = (int)padding.first; d.pad_w = (int)padding.second;
    d.dil_h = (int)dilation.first; d.dil_w = (int)dilation.second;
    if (bias && (bias->shape().size() != 1 || bias->shape()[0] != (size_t)d.c_out)) panic("Bias must be 1D with C_out elements");
    int ho, wo;
    check(tp_conv2d_out_dims(&d, &ho, &wo));
    Tensor out = Tensor::empty({(size_t)d.n, (size_t)d.c_out, (size_t)ho, (size_t)wo});
    check(tp_conv2d_fwd(ctx(), buf(), weight.buf(), bias ? bias->buf() : nullptr, out.buf(), &d, relu ? 1 : 0));
    const bool full = Config::conv_full_adjoint();
    // strict_reference (A1): im2col and transpose_4d drop the tape links, so only the bias node
    // (add_bias_4d, src/tensor.rs:2003-2027) delivers a gradient and nothing crosses the layer.
    bool bias_rg = bias && bias->needs_grad();
    bool any = full ? (needs_grad() || weight.needs_grad() || bias_rg) : bias_rg;
    if (any) {
        out.set_requires_grad(true);
        Tensor x = *this, w = weight;
        std::optional<Tensor> b;
        if (bias) b = *bias;
        tape_push(out, [x, w, b, out, d, relu, full]() {
            tp_buf* g = out.grad_buf();
            if (!g) return;
            int ax = 0, aw = 0, ab = 0;
            tp_buf* gx = (full && x.needs_grad()) ? x.impl()->grad_for_write(&ax) : nullptr;
            tp_buf* gw = (full && w.needs_grad()) ? w.impl()->grad_for_write(&aw) : nullptr;
            tp_buf* gb = (b && b->needs_grad()) ? b->impl()->grad_for_write(&ab) : nullptr;
            check(tp_conv2d_bwd(ctx(), x.buf(), w.buf(), g, relu ? out.buf() : nullptr, gx, gw, gb, &d, ax, aw, ab));
        });
    }
    return out;
}

Tensor Tensor::conv2d(const Tensor& weight, const Tensor* bias, Pair stride, Pair padding, Pair dilation) const {
    return conv2d_impl(weight, bias, stride, padding, dilation, false);
}

Tensor Tensor::conv2d_relu(const Tensor& weight, const Tensor* bias, Pair stride, Pair padding, Pair dilation) const {
    // The reference runs conv2d then relu (src/tensor.rs:1387-1388); here ReLU is the conv epilogue and its
    // mask is folded into the backward.  With no bias in strict mode the reference's output does not
    // require grad, which conv2d_impl reproduces.
    return conv2d_impl(weight, bias, stride, padding, dilation, true);
}

// ---- a stack of conv2d_relu (+ max_pool2d) layers as one fused forward (tp_conv_stack_fwd) ------------------------------------
// Strict-reference autograd only: im2col / transpose_4d drop the tape links, so of everything these layers record
// only the LAST layer's add_bias_4d node (src/tensor.rs:2003-2027) ever sees a gradient — through that layer's ReLU
// (src/ops.rs:358-370) and, when it is pooled, through max_pool2d's scatter (src/tensor.rs:1476-1516), which hands every
// pooled gradient to exactly one pre-pool unit whose ReLU gate equals [pooled output > 0].  Hence
//     bias.grad[c] (+)= sum_{n, p} g[n, c, p] * [out[n, c, p] > 0]          (out = the stack's output, pooled or not)
// and no intermediate activation or pooling index is needed.
Tensor Tensor::conv_stack(const std::vector<ConvStackLayer>& layers, int gap) const {
    if (shape().size() != 4 || layers.empty() || layers.size() > 8 || needs_grad() || Config::conv_full_adjoint()) return Tensor();
    const int L = (int)layers.size();
    const tp_buf* wb[8];
    const tp_buf* bb[8];
    int co[8], po[8], re[8];
    size_t c = shape()[1], h = shape()[2], w = shape()[3];
    for (int l = 0; l < L; ++l) {
        const Tensor& wt = layers[l].weight;
        if (wt.shape().size() != 4 || wt.shape()[1] != c || wt.shape()[2] != 3 || wt.shape()[3] != 3) return Tensor();
        if (layers[l].bias && (layers[l].bias->shape().size() != 1 || layers[l].bias->shape()[0] != wt.shape()[0])) return Tensor();
        wb[l] = wt.buf();
        bb[l] = layers[l].bias ? layers[l].bias->buf() : nullptr;
        co[l] = (int)wt.shape()[0];
        po[l] = layers[l].pool ? 1 : 0;
        re[l] = layers[l].relu ? 1 : 0;
        c = wt.shape()[0];
        if (layers[l].pool) { h /= 2; w /= 2; }
    }
    if (!h || !w) return Tensor();
    const size_t n = shape()[0];
    const ConvStackLayer& last = layers[L - 1];
    const bool bias_rg = last.bias && last.bias->needs_grad();
    const bool relu = last.relu;
    const int nn = (int)n, cc = (int)c, hw = (int)(h * w);
    // the closures keep every weight alive like the reference's per-layer closures do
    std::vector<Tensor> keep;
    for (auto& ly : layers) keep.push_back(ly.weight);
    if (gap) {
        // AdaptiveAvgPool2d::global (+ Flatten) on top (src/nn.rs:670-686, 743-745): avg_pool2d backward hands g / hw to every unit
        // of a plane (src/tensor.rs:1600-1655), so with cnt[n, c] = #{units > 0} of the plane
        //     bias.grad[c] (+)= sum_n (g[n, c] / hw) * cnt[n, c]
        // and the [N, C, H, W] gradient, its ReLU mask pass and the conv output itself are never needed again.
        Tensor feat = Tensor::empty(gap == 2 ? Shape{n, c} : Shape{n, c, 1, 1});
        Tensor cnt;
        if (bias_rg && relu) cnt = Tensor::empty({n, c});
        int rc = tp_conv_stack_gap_fwd(ctx(), buf(), nn, (int)shape()[1], (int)shape()[2], (int)shape()[3], L, wb, bb, co, po, re, feat.buf(),
                                       cnt.defined() ? cnt.buf() : nullptr);
        if (rc == TP_ERR_UNSUPPORTED) return Tensor();
        check(rc);
        if (bias_rg) {
            feat.set_requires_grad(true);
            Tensor b = *last.bias;
            tape_push(feat, [b, feat, cnt, nn, cc, hw, keep]() {
                tp_buf* g = feat.grad_buf();
                if (!g) return;
                int ab = 0;
                tp_buf* gb = b.impl()->grad_for_write(&ab);
                check(tp_gap_relu_bias_grad(ctx(), g, cnt.defined() ? cnt.buf() : nullptr, gb, nn, cc, hw, ab));
            });
        }
        return feat;
    }
    Tensor out = Tensor::empty({n, c, h, w});
    int rc = tp_conv_stack_fwd(ctx(), buf(), nn, (int)shape()[1], (int)shape()[2], (int)shape()[3], L, wb, bb, co, po, re, out.buf());
    if (rc == TP_ERR_UNSUPPORTED) return Tensor();
    check(rc);
    if (bias_rg) {
        out.set_requires_grad(true);
        Tensor b = *last.bias;
        tape_push(out, [b, out, relu, nn, cc, hw, keep]() {
            tp_buf* g = out.grad_buf();
            if (!g) return;
            int ab = 0;
            tp_buf* gb = b.impl()->grad_for_write(&ab);
            if (relu) {
                tp_buf* masked = nullptr;
                check(tp_buf_alloc(ctx(), (size_t)nn * cc * hw, &masked));
                int rc2 = tp_relu_bwd(ctx(), out.buf(), g, masked, (size_t)nn * cc * hw, 0);
                if (!rc2) rc2 = tp_bias_grad_4d(ctx(), masked, gb, nn, cc, hw, ab);
                tp_buf_release(masked);
                check(rc2);
            } else {
                check(tp_bias_grad_4d(ctx(), g, gb, nn, cc, hw, ab));
            }
        });
    }
    return out;
}

// ---- pooling  (src/tensor.rs:1391-1660) ---------------------------------------------------------------------------------------------
static tp_pool_desc pool_desc(const Tensor& x, Pair k, std::optional<Pair> s, Pair p) {
    if (x.shape().size() != 4) panic("pooling expects a 4-D [N,C,H,W] tensor");
    Pair st = s.value_or(k);
    tp_pool_desc d;
    d.n = (int)x.shape()[0]; d.c = (int)x.shape()[1]; d.h = (int)x.shape()[2]; d.w = (int)x.shape()[3];
    d.kh = (int)k.first; d.kw = (int)k.second; d.stride_h = (int)st.first; d.stride_w = (int)st.second;
    d.pad_h = (int)p.first; d.pad_w = (int)p.second;
    return d;
}

Tensor Tensor::max_pool2d(Pair kernel, std::optional<Pair> stride, Pair padding) const {
    tp_pool_desc d = pool_desc(*this, kernel, stride, padding);
    int ho, wo;
    check(tp_pool_out_dims(&d, &ho, &wo));
    Tensor out = Tensor::empty({(size_t)d.n, (size_t)d.c, (size_t)ho, (size_t)wo});
    Tensor arg = Tensor::empty(out.shape());                   // int32 argmax indices (reference: usize, :1416)
    check(tp_maxpool2d_fwd(ctx(), buf(), out.buf(), arg.buf(), &d));
    if (needs_grad()) {
        out.set_requires_grad(true);
        Tensor x = *this;
        Tape::push_unary_op(x, out, [x, out, arg, d]() {
            tp_buf* g = out.grad_buf();
            if (!g) return;
            int acc;
            tp_buf* gin = x.impl()->grad_for_write(&acc);      // zeroes the plane, then scatters: overwrite (A6)
            check(tp_maxpool2d_bwd(ctx(), g, arg.buf(), gin, &d));
        });
    }
    return out;
}

Tensor Tensor::avg_pool2d(Pair kernel, std::optional<Pair> stride, Pair padding) const {
    tp_pool_desc d = pool_desc(*this, kernel, stride, padding);
    int ho, wo;
    check(tp_pool_out_dims(&d, &ho, &wo));
    Tensor out = Tensor::empty({(size_t)d.n, (size_t)d.c, (size_t)ho, (size_t)wo});
    check(tp_avgpool2d_fwd(ctx(), buf(), out.buf(), &d));
    if (needs_grad()) {
        out.set_requires_grad(true);
        Tensor x = *this;
        Tape::push_unary_op(x, out, [x, out, d]() {
            tp_buf* g = out.grad_buf();
            if (!g) return;
            int acc;
            tp_buf* gin = x.impl()->grad_for_write(&acc);
            check(tp_avgpool2d_bwd(ctx(), g, gin, &d, acc));
        });
    }
    return out;
}

}  // namespace taper

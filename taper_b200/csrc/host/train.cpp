// Host layer, part 3: the training loop (src/train.rs), the MNIST data path (src/data/mnist.rs) and
// CUDA-graph capture of one whole training step.
#include "taper_internal.hpp"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <thread>
#include <cstring>
#include <deque>
#include <fstream>
#include <map>
#include <numeric>
#include <sstream>

namespace taper {

// =====================================================================================================
// data  (src/data/mnist.rs)
// =====================================================================================================
namespace data {

namespace {
uint64_t splitmix64(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

std::vector<unsigned char> read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) panic("Failed to open %s", path.c_str());
    return std::vector<unsigned char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

uint32_t be32(const unsigned char* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }
}  // namespace

MNISTDataset::MNISTDataset(bool train_, const std::string& dir) : train(train_) {
    // IDX parser (src/data/mnist.rs:184-273).  The reference downloads missing files (:60-181); there is
    // no network here, so missing files are an error.
    auto img = read_file(dir + (train ? "/train_images" : "/test_images"));
    auto lab = read_file(dir + (train ? "/train_labels" : "/test_labels"));
    if (img.size() < 16 || be32(img.data()) != 0x00000803) panic("Invalid magic number for images");
    size_t n = be32(img.data() + 4), rows = be32(img.data() + 8), cols_ = be32(img.data() + 12);
    if (rows != 28 || cols_ != 28) panic("Unexpected image size: %zux%zu", rows, cols_);
    if (img.size() != 16 + n * 784) panic("Image file size mismatch");
    if (lab.size() < 8 || be32(lab.data()) != 0x00000801) panic("Invalid magic number for labels");
    if (be32(lab.data() + 4) != n || lab.size() != 8 + n) panic("Label file size mismatch");
    cols = 784;
    images.resize(n * 784);
    for (size_t i = 0; i < n * 784; ++i) images[i] = (float)img[16 + i] / 255.0f;          // :225
    images_u8.assign(img.begin() + 16, img.end());                                          // the same pixels, undivided
    labels.resize(n);
    for (size_t i = 0; i < n; ++i) labels[i] = (float)lab[8 + i];                           // :268
}

MNISTDataset MNISTDataset::synthetic(size_t n, uint64_t seed) {
    MNISTDataset d;
    d.images.resize(n * 784);
    d.labels.resize(n);
    uint64_t s = seed;
    for (auto& v : d.images) v = (float)(splitmix64(s) >> 40) * (1.0f / 16777216.0f);
    for (auto& v : d.labels) v = (float)(splitmix64(s) % 10);
    return d;
}

MNISTDataset MNISTDataset::from_arrays(const void* images_, bool is_u8, const float* labels_, size_t n, size_t cols_) {
    if (!images_ || !labels_ || !cols_) panic("MNISTDataset::from_arrays: NULL or empty input");
    MNISTDataset d;
    d.cols = cols_;
    if (is_u8) {
        const uint8_t* p = static_cast<const uint8_t*>(images_);
        d.images_u8.assign(p, p + n * cols_);
    } else {
        const float* p = static_cast<const float*>(images_);
        d.images.assign(p, p + n * cols_);
    }
    d.labels.assign(labels_, labels_ + n);
    return d;
}

void MNISTDataset::ensure_f32() {
    if (!images.empty() || images_u8.empty()) return;
    images.resize(images_u8.size());
    for (size_t i = 0; i < images_u8.size(); ++i) images[i] = (float)images_u8[i] / 255.0f;      // :225
}

void MNISTDataset::normalize(float mean, float std) {
    ensure_f32();
    for (auto& p : images) p = (p - mean) / std;
    images_u8.clear();
    images_u8.shrink_to_fit();
}

// ---- pinned prefetch pipeline ---------------------------------------------------------------------------------------------------
// Worker threads live as long as the loader (parked between epochs): starting an epoch costs a notify, not a thread spawn.
struct DataLoader::Pipe {
    static constexpr int kSlots = 12;                // > the trainer's result ring (8) + its staging depth
    struct Slot {
        void* img = nullptr;
        float* lab = nullptr;
        size_t batch = 0;
        long turn = 0;                               // batch index allowed to fill this slot next
        long ready = -1;                             // batch index the slot holds
    };
    Slot slots[kSlots];
    size_t img_bytes = 0;
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv;
    bool quit = false;
    uint64_t gen = 0;                                // epoch generation the workers should be running
    uint64_t aborted = 0;                            // generations <= this one are cancelled
    int finished = 0;                                // workers done with generation `gen`
    // epoch parameters (written under mu before gen is bumped)
    bool u8 = false;
    size_t n_batches = 0, next_consume = 0, bs = 0, cols = 0, n = 0;
    const uint32_t* idx = nullptr;
    const MNISTDataset* ds = nullptr;

    void worker(int w, int T) {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return quit || gen != seen; });
                if (quit) return;
                seen = gen;
            }
            for (size_t b = (size_t)w; b < n_batches; b += (size_t)T) {
                Slot& s = slots[b % kSlots];
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return quit || aborted >= seen || s.turn == (long)b; });
                    if (quit || aborted >= seen) break;
                }
                const size_t first = b * bs, count = std::min(bs, n - first);
                if (u8) {
                    uint8_t* dst = static_cast<uint8_t*>(s.img);
                    for (size_t i = 0; i < count; ++i) std::memcpy(dst + i * cols, &ds->images_u8[(size_t)idx[first + i] * cols], cols);
                } else {
                    float* dst = static_cast<float*>(s.img);
                    for (size_t i = 0; i < count; ++i)
                        std::memcpy(dst + i * cols, &ds->images[(size_t)idx[first + i] * cols], cols * sizeof(float));
                }
                for (size_t i = 0; i < count; ++i) s.lab[i] = ds->labels[idx[first + i]];
                {
                    std::lock_guard<std::mutex> g(mu);
                    s.batch = count;
                    s.ready = (long)b;
                }
                cv.notify_all();
            }
            {
                std::lock_guard<std::mutex> g(mu);
                finished++;
            }
            cv.notify_all();
        }
    }
    ~Pipe() {
        {
            std::lock_guard<std::mutex> g(mu);
            quit = true;
        }
        cv.notify_all();
        for (auto& t : workers) t.join();
        for (auto& s : slots) {
            if (s.img) tp_host_free_pinned(s.img);
            if (s.lab) tp_host_free_pinned(s.lab);
        }
    }
};

DataLoader::DataLoader(MNISTDataset dataset, size_t batch_size, bool shuffle, uint64_t seed)
    : DataLoader(std::make_shared<MNISTDataset>(std::move(dataset)), batch_size, shuffle, seed) {}

DataLoader::DataLoader(std::shared_ptr<MNISTDataset> dataset, size_t batch_size, bool shuffle, uint64_t seed)
    : dataset_(std::move(dataset)), batch_size_(batch_size), shuffle_(shuffle), rng_state_(seed) {
    if (!dataset_ || !batch_size_) panic("DataLoader: NULL dataset or zero batch size");
    sample_shape = {dataset_->cols};
    indices_.resize(dataset_->len());
    std::iota(indices_.begin(), indices_.end(), 0u);
    if (shuffle_) reset();
}

DataLoader::~DataLoader() { stop_prefetch(); }

void DataLoader::reset() {
    stop_prefetch();
    current_ = 0;
    if (shuffle_)                                                                           // Fisher-Yates (:353-357)
        for (size_t i = indices_.size(); i > 1; --i) std::swap(indices_[i - 1], indices_[splitmix64(rng_state_) % i]);
}

bool DataLoader::next(std::vector<float>& images, std::vector<float>& labels, size_t& batch) {
    if (current_ >= dataset_->len()) return false;
    dataset_->ensure_f32();
    const size_t cols = dataset_->cols;
    size_t end = std::min(current_ + batch_size_, dataset_->len());
    batch = end - current_;
    images.resize(batch * cols);
    labels.resize(batch);
    for (size_t i = 0; i < batch; ++i) {                                                    // get_batch (:276-309)
        size_t idx = indices_[current_ + i];
        std::memcpy(&images[i * cols], &dataset_->images[idx * cols], cols * sizeof(float));
        labels[i] = dataset_->labels[idx];
    }
    current_ = end;
    return true;
}

void DataLoader::stop_prefetch() {
    if (!pipe_ || pipe_->gen == 0) return;
    Pipe& p = *pipe_;
    std::unique_lock<std::mutex> lk(p.mu);
    p.aborted = p.gen;                               // cancel whatever is left of the current epoch
    p.cv.notify_all();
    p.cv.wait(lk, [&] { return p.finished == (int)p.workers.size(); });
}

void DataLoader::start_prefetch(bool u8, size_t max_batches) {
    stop_prefetch();
    if (u8 && !dataset_->has_u8()) panic("DataLoader: the dataset holds no u8 pixels");
    if (!u8) dataset_->ensure_f32();
    if (!pipe_) pipe_.reset(new Pipe());
    Pipe& p = *pipe_;
    const size_t cols = dataset_->cols;
    const size_t need = batch_size_ * cols * (u8 ? 1 : sizeof(float));
    if (p.img_bytes < need) {
        for (auto& s : p.slots) {
            if (s.img) tp_host_free_pinned(s.img);
            if (s.lab) tp_host_free_pinned(s.lab);
            s.img = nullptr; s.lab = nullptr;
            void* q = nullptr;
            check(tp_host_alloc_pinned((need + 63) & ~(size_t)63, &q));
            s.img = q;
            check(tp_host_alloc_pinned(((batch_size_ * sizeof(float)) + 63) & ~(size_t)63, &q));
            s.lab = static_cast<float*>(q);
        }
        p.img_bytes = need;
    }
    if (p.workers.empty()) {
        unsigned hw = std::thread::hardware_concurrency();
        const int T = hw >= 16 ? 4 : hw >= 8 ? 2 : 1;
        for (int w = 0; w < T; ++w) p.workers.emplace_back([&p, w, T]() { p.worker(w, T); });
    }
    size_t nb = num_batches();
    if (max_batches && max_batches < nb) nb = max_batches;
    {
        std::lock_guard<std::mutex> g(p.mu);
        p.u8 = u8;
        p.n_batches = nb;
        p.next_consume = 0;
        p.bs = batch_size_; p.cols = cols; p.n = dataset_->len();
        p.idx = indices_.data();
        p.ds = dataset_.get();
        for (int i = 0; i < Pipe::kSlots; ++i) { p.slots[i].turn = i; p.slots[i].ready = -1; }
        p.finished = 0;
        p.gen++;
    }
    p.cv.notify_all();
}

bool DataLoader::next_pinned(Batch& out) {
    if (!pipe_) panic("DataLoader::next_pinned: start_prefetch first");
    Pipe& p = *pipe_;
    const size_t b = p.next_consume;
    if (b >= p.n_batches) return false;
    Pipe::Slot& s = p.slots[b % Pipe::kSlots];
    {
        std::unique_lock<std::mutex> lk(p.mu);
        p.cv.wait(lk, [&] { return s.ready == (long)b; });
    }
    out.images = s.img; out.labels = s.lab; out.batch = s.batch; out.slot = (int)(b % Pipe::kSlots); out.u8 = p.u8;
    p.next_consume = b + 1;
    current_ = std::min(dataset_->len(), (b + 1) * batch_size_);
    return true;
}

void DataLoader::release(int slot) {
    if (!pipe_ || slot < 0 || slot >= Pipe::kSlots) return;
    Pipe& p = *pipe_;
    {
        std::lock_guard<std::mutex> g(p.mu);
        p.slots[slot].turn = p.slots[slot].ready + Pipe::kSlots;
    }
    p.cv.notify_all();
}

}  // namespace data

// =====================================================================================================
// train  (src/train.rs)
// =====================================================================================================
namespace train {

void Metrics::print_last() const {
    if (train_loss.empty() || val_loss.empty()) return;
    std::printf("Train Loss: %.4f | Train Acc: %.2f%% | Val Loss: %.4f | Val Acc: %.2f%%\n", train_loss.back(),
                train_acc.back() * 100.0f, val_loss.back(), val_acc.back() * 100.0f);
}

void Metrics::plot_summary() const {
    std::printf("\nTraining Summary:\n==================================================\n");
    if (!train_acc.empty()) {
        std::printf("Best Train Accuracy: %.2f%%\n", *std::max_element(train_acc.begin(), train_acc.end()) * 100.0f);
        if (!val_acc.empty()) std::printf("Best Val Accuracy: %.2f%%\n", *std::max_element(val_acc.begin(), val_acc.end()) * 100.0f);
        std::printf("Final Train Accuracy: %.2f%%\n", train_acc.back() * 100.0f);
        if (!val_acc.empty()) std::printf("Final Val Accuracy: %.2f%%\n", val_acc.back() * 100.0f);
        if (!epoch_times.empty()) {
            float total = std::accumulate(epoch_times.begin(), epoch_times.end(), 0.0f);
            std::printf("Total Training Time: %.2fs\nAverage Epoch Time: %.2fs\n", total, total / (float)epoch_times.size());
        }
    }
    std::printf("==================================================\n");
}

namespace {
constexpr size_t kRing = 8;          // result slots in flight (host may run this many steps ahead)
constexpr size_t kStage = 3;         // input staging buffers per batch shape (H2D of step i+1 overlaps step i)

struct Slot {                        // per batch-shape persistent state
    Tensor x, y;                     // device inputs the step reads
    // pinned-host inputs: copied on the context's copy stream into a rotating staging buffer
    Tensor sx[kStage], sy[kStage];
    tp_event* ready[kStage] = {};    // copy stream: staging buffer k holds the batch
    tp_event* done[kStage] = {};     // main stream: the reader of staging buffer k has been enqueued and finished
    bool done_valid[kStage] = {};
    uint64_t staged = 0;
    int eager_runs = 0;
    tp_graph* graph = nullptr;       // captured step (host-fed variant)
    tp_graph* graph_resident = nullptr;      // captured gather + step (device-resident dataset)
    std::vector<std::function<void()>> keep, keep_resident;    // tape closures (own the step's tensors)
};
}  // namespace

struct Trainer::Impl {
    std::map<std::vector<size_t>, Slot> slots;               // key: full input shape [B, sample...]
    Tensor result;                                           // device {loss, correct}
    Tensor loss_view, correct_view;
    float* ring = nullptr;                                   // pinned kRing x 4 floats: {loss, correct, seq bits, pad}
    tp_event* events[kRing] = {};
    uint32_t slot_seq[kRing] = {};                           // != 0: the step kernel publishes this value in the slot (host polls)
    uint32_t next_seq = 1;
    size_t head = 0, tail = 0;                               // FIFO of outstanding results
    // resident dataset
    tp_buf *ds_images = nullptr, *ds_labels = nullptr, *ds_perm = nullptr, *ds_cursor = nullptr;
    size_t ds_n = 0;
    Shape ds_sample;
    bool ds_u8 = false;                                      // ds_images holds u8 pixels (MNIST on-disk format)
    int world = 1;
    // fused device step per batch size (NULL once a size is known not to qualify)
    std::map<std::pair<size_t, size_t>, tp_step*> fused;
    tp_xchg* xchg = nullptr;                                 // NVLink peer-memory gradient exchange (world > 1)
    bool xchg_connected = false;

    // fused device step per (batch, sample width): the persistent-kernel tape for small MLPs, the tcgen05 kernel plan for wide
    // ones (tp_step_kind).  Data-parallel: the persistent kernel needs the connected peer window, the plan uses NCCL.
    tp_step* fused_step(Trainer& tr, size_t batch, const Shape& sample_shape) {
        if (sample_shape.size() != 1) return nullptr;
        const std::pair<size_t, size_t> key{batch, sample_shape[0]};
        auto it = fused.find(key);
        if (it != fused.end()) return it->second;
        tp_step* st = nullptr;
        tp_step_desc d;
        tp_buf* b[5] = {};
        if (optim::describe_fused_step(*tr.model, *tr.optimizer, batch, &d, b) && (size_t)d.dims[0] == sample_shape[0]) {
            const int kind = tp_step_kind(&d);
            if (kind == 1 && world != 1 && !xchg_connected) return nullptr;      // not cached: the window may still be connected
            d.materialize_grads = 0;
            d.data_parallel = world > 1 ? 1 : 0;
            // a step that qualifies on paper but cannot be built on this device (no cooperative launch, shared memory) simply
            // stays on the tape + graph path; with an exchange every rank must agree, so there a failure is an error
            // data-parallel: the persistent kernel needs the connected window; the plan uses it when there is one (two-phase
            // peer-memory exchange fused with the optimizer) and the NCCL communicator otherwise
            int rc = tp_step_create(ctx(), &d, b[0], b[1], b[2], b[3], b[4], result.buf(), (world > 1 && xchg_connected) ? xchg : nullptr, &st);
            if (rc != TP_OK) {
                if (world > 1) check(rc);
                std::fprintf(stderr, "taper_b200: fused step unavailable (%s); using the tape + CUDA-graph path\n", tp_last_error());
                st = nullptr;
            }
        }
        fused[key] = st;
        return st;
    }

    // The wide plan keeps bf16 operand planes of the parameters, refreshed by its own optimizer kernel.  Anything else that
    // writes the parameters (set_data, checkpoint load, broadcast, a step on another path) bumps their versions: the plan is
    // told to re-derive the planes whenever the versions are not the ones it left behind.
    std::vector<float> u8_scratch;
    bool fused_is_wide(tp_step* st) const { return st && tp_step_is_wide(st) == 1; }
    tp_step* last_fused = nullptr;
    uint64_t fused_stamp = 0;
    uint64_t param_stamp(Trainer& tr) const {
        uint64_t s = 0;
        for (auto& t : tr.model->parameters()) s += t.impl()->version;
        return s;
    }
    void before_fused(Trainer& tr, tp_step* st) {
        if (st != last_fused || param_stamp(tr) != fused_stamp) check(tp_step_refresh(st));
    }
    void after_fused(Trainer& tr, tp_step* st) {
        last_fused = st;
        fused_stamp = param_stamp(tr);
    }

    size_t ds_cursor_host = 0;                               // host mirror of the device cursor

    float* next_result_slot() const { return ring + 4 * (head % kRing); }
    uint32_t next_result_seq() {                             // never 0
        if (next_seq == 0) next_seq = 1;
        return next_seq;
    }
    // in_kernel: the step kernel writes {loss, correct} and then the sequence word into next_result_slot() (mapped pinned
    // memory): no D2H copy and no CUDA event on the stream — fetch() polls the slot
    void enqueue_result(bool in_kernel = false) {
        size_t i = head % kRing;
        if (in_kernel) {
            slot_seq[i] = next_seq++;
        } else {
            slot_seq[i] = 0;
            check(tp_buf_download_async(ctx(), result.buf(), ring + 4 * i, 2));
            check(tp_ctx_device_error_async(ctx(), reinterpret_cast<int*>(ring + 4 * i + 3)));
            check(tp_event_record(ctx(), events[i]));
        }
        head++;
    }

    ~Impl() {
        tp_sync(ctx());
        for (auto& kv : fused) tp_step_destroy(kv.second);
        tp_xchg_destroy(xchg);
        for (auto& kv : slots) {
            tp_graph_destroy(kv.second.graph);
            tp_graph_destroy(kv.second.graph_resident);
            for (size_t k = 0; k < kStage; ++k) {
                tp_event_destroy(kv.second.ready[k]);
                tp_event_destroy(kv.second.done[k]);
            }
        }
        slots.clear();
        for (auto* e : events) tp_event_destroy(e);
        if (ring) tp_host_free_pinned(ring);
        tp_buf_release(ds_images); tp_buf_release(ds_labels); tp_buf_release(ds_perm); tp_buf_release(ds_cursor);
    }
};

Trainer::Trainer(std::shared_ptr<nn::Module> m, std::shared_ptr<optim::Optimizer> o, std::shared_ptr<optim::LRScheduler> s)
    : model(std::move(m)), optimizer(std::move(o)), scheduler(std::move(s)), p_(new Impl()) {
    p_->result = Tensor::zeros({2});
    tp_buf *l, *c;
    check(tp_buf_slice(p_->result.buf(), 0, 1, &l));
    check(tp_buf_slice(p_->result.buf(), 1, 1, &c));
    p_->loss_view = Tensor::adopt(l, {1});
    p_->correct_view = Tensor::adopt(c, {1});
    void* ring = nullptr;
    check(tp_host_alloc_pinned(kRing * 4 * sizeof(float), &ring));
    std::memset(ring, 0, kRing * 4 * sizeof(float));
    p_->ring = (float*)ring;
    for (auto& e : p_->events) check(tp_event_create(ctx(), &e));
}

Trainer::~Trainer() = default;

void Trainer::init_data_parallel(int rank, int world, const void* uid) {
    dist::init(rank, world, uid);
    Impl& p = *p_;
    // steps captured / compiled for the old world carry its gradient scale (a by-value kernel argument) and no exchange
    check(tp_sync(ctx()));
    for (auto& kv : p.slots) {
        tp_graph_destroy(kv.second.graph); kv.second.graph = nullptr; kv.second.keep.clear();
        tp_graph_destroy(kv.second.graph_resident); kv.second.graph_resident = nullptr; kv.second.keep_resident.clear();
        kv.second.eager_runs = 0;
    }
    for (auto& kv : p.fused) tp_step_destroy(kv.second);
    p.fused.clear();
    p.last_fused = nullptr;
    p.world = world;
    optimizer->set_grad_scale(1.0f / (float)world);
}

void Trainer::peer_exchange_handle(void* out64) {
    Impl& p = *p_;
    if (p.world < 2) panic("Trainer::peer_exchange_handle: call init_data_parallel first");
    if (!p.xchg) check(tp_xchg_create(ctx(), optim::arena_total(optimizer->arena()), dist::rank(), p.world, &p.xchg));
    check(tp_xchg_handle(p.xchg, out64));
}

void Trainer::peer_exchange_connect(const void* handles) {
    Impl& p = *p_;
    if (!p.xchg) panic("Trainer::peer_exchange_connect: no exchange window (peer_exchange_handle first)");
    check(tp_xchg_connect(p.xchg, handles, p.world));
    p.xchg_connected = true;
    for (auto& kv : p.fused) tp_step_destroy(kv.second);    // steps compiled before the exchange existed
    p.fused.clear();
    p.last_fused = nullptr;
}

void Trainer::broadcast_parameters(int root) {
    auto a = optimizer->arena();
    dist::broadcast(optim::arena_param_buf(a), optim::arena_total(a), root);
    for (auto& t : model->parameters()) t.impl()->version++;
}

namespace {

// The loop body of Trainer::train_epoch (src/train.rs:106-138) on device-resident inputs.
void step_body(Trainer& tr, const Tensor& x, const Tensor& y, const Tensor& loss_out, const Tensor& correct_out, int world) {
    Tape::reset();                                                               // :108
    Tensor logits = tr.model->forward(x);                                        // :111
    // the loss tensor is a fresh handle on the result slot each step: its grad state must start as None
    loss_out.impl()->has_grad = false;
    loss_out.impl()->tape_node = 0;
    Tensor loss = loss::cross_entropy_with_accuracy_into(logits, y, loss_out, &correct_out);      // :112-115, one launch
    // loss.backward() (:121) = grad := ones, then the tape walk.  The result slot's gradient buffer holds 1.0 from its
    // first allocation on and nothing ever writes it again, so the per-step fill is skipped.
    TensorImpl& li = *loss.impl();
    if (!li.grad) {
        int acc;
        check(tp_buf_fill(ctx(), li.grad_for_write(&acc), 1.0f, 1));
    }
    li.has_grad = true;
    if (li.tape_node != 0) taper::backward(li.tape_node - 1);
    if (world > 1) {                                                             // sum of per-rank mean gradients
        auto a = tr.optimizer->arena();
        dist::allreduce_sum(optim::arena_grad_buf(a), optim::arena_total(a));
    }
    tr.optimizer->step();                                                        // :124
    tr.optimizer->zero_grad();                                                   // :125
}

Shape full_shape(size_t batch, const Shape& sample) {
    Shape s{batch};
    s.insert(s.end(), sample.begin(), sample.end());
    return s;
}

}  // namespace

size_t Trainer::pending() const { return p_->head - p_->tail; }

int Trainer::fused_kind() const { return !p_->last_fused ? 0 : (p_->fused_is_wide(p_->last_fused) ? 2 : 1); }

StepResult Trainer::fetch() {
    if (p_->head == p_->tail) panic("Trainer::fetch: no step outstanding");
    size_t i = p_->tail % kRing;
    if (p_->slot_seq[i]) {
        // the step kernel publishes the slot itself: spin on the sequence word (bounded: a lost kernel must not hang the host)
        volatile uint32_t* w = reinterpret_cast<volatile uint32_t*>(p_->ring + 4 * i + 2);
        auto t0 = std::chrono::steady_clock::now();
        uint64_t spins = 0;
        while (*w != p_->slot_seq[i]) {
            if ((++spins & 0xFFFF) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(30)) {
                check(tp_sync(ctx()));
                if (*w != p_->slot_seq[i]) panic("Trainer::fetch: the step kernel did not publish its result");
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
    } else {
        check(tp_event_sync(p_->events[i]));
    }
    StepResult r{p_->ring[4 * i], p_->ring[4 * i + 1]};
    int err;
    std::memcpy(&err, p_->ring + 4 * i + 3, sizeof err);
    p_->tail++;
    if (err != 0) {
        // the reference panics on a label outside the class range (`logp[i * c + t]`, src/loss.rs:160-162); barrier / peer
        // timeouts mean the step did not complete.  The flag is sticky: the context does not train on.
        const char* why = err == 1 ? "a label is outside [0, classes)" : err == 2 ? "a grid barrier of the fused step timed out"
                        : err == 3 ? "a peer never delivered its gradient slice (exchange timeout)" : "unknown device error";
        panic("Trainer::fetch: device error %d: %s", err, why);
    }
    return r;
}

// enqueue: [optional gather] + step (eager for the first iteration of a shape, then captured, then replayed),
// followed by the asynchronous read-back of {loss, correct} into the next ring slot.
void Trainer::train_batch_async(const float* images, const float* labels, size_t batch, const Shape& sample_shape, bool pinned) {
    train_batch_async_impl(images, labels, batch, sample_shape, pinned, false);
}

void Trainer::train_batch_async_u8(const uint8_t* images, const float* labels, size_t batch, const Shape& sample_shape, bool pinned) {
    train_batch_async_impl(images, labels, batch, sample_shape, pinned, true);
}

void Trainer::train_batch_async_impl(const void* images_any, const float* labels, size_t batch, const Shape& sample_shape, bool pinned,
                                     bool u8) {
    Impl& p = *p_;
    if (pending() >= kRing) panic("Trainer: %zu steps outstanding; call fetch()", kRing);
    Shape fs_shape = full_shape(batch, sample_shape);
    tp_ctx* c = ctx();
    tp_step* fs = use_fused_ ? p.fused_step(*this, batch, sample_shape) : nullptr;
    if (u8) {
        // MNIST's on-disk pixel format (src/data/mnist.rs:225 converts to f32 at load time): the bytes cross PCIe as they are
        // and the step divides by 255 on the device
        if (!fs || !p.fused_is_wide(fs)) {
            // no wide plan for this model: widen on the host side of the boundary into the f32 path
            const uint8_t* src = static_cast<const uint8_t*>(images_any);
            p.u8_scratch.resize(batch * shape_numel(sample_shape));
            for (size_t i = 0; i < p.u8_scratch.size(); ++i) p.u8_scratch[i] = (float)src[i] / 255.0f;
            train_batch_async_impl(p.u8_scratch.data(), labels, batch, sample_shape, false, false);
            return;
        }
    }
    const float* images = static_cast<const float*>(images_any);
    Slot& s = p.slots[fs_shape];
    if (!s.x.defined()) {
        s.x = Tensor::empty(fs_shape);
        s.y = Tensor::empty({batch});
    }
    tp_buf *xin = s.x.buf(), *yin = s.y.buf();
    size_t k = 0;
    if (pinned) {
        // H2D on the copy stream into staging buffer k while earlier steps still run on the main stream
        k = s.staged % kStage;
        if (!s.sx[k].defined()) {
            s.sx[k] = Tensor::empty(fs_shape);
            s.sy[k] = Tensor::empty({batch});
            check(tp_event_create(c, &s.ready[k]));
            check(tp_event_create(c, &s.done[k]));
        }
        if (s.done_valid[k]) check(tp_copy_wait_event(c, s.done[k]));
        check(tp_copy_upload_pinned(c, s.sx[k].buf(), images, u8 ? (s.x.numel() + 3) / 4 : s.x.numel()));
        check(tp_copy_upload_pinned(c, s.sy[k].buf(), labels, batch));
        check(tp_copy_event_record(c, s.ready[k]));
        check(tp_stream_wait_event(c, s.ready[k]));
        if (fs) {
            xin = s.sx[k].buf();
            yin = s.sy[k].buf();
        } else {                                           // a captured graph reads fixed buffers: device-to-device hop
            check(tp_buf_copy(c, s.x.buf(), s.sx[k].buf(), s.x.numel()));
            check(tp_buf_copy(c, s.y.buf(), s.sy[k].buf(), batch));
            check(tp_event_record(c, s.done[k]));
            s.done_valid[k] = true;
        }
        s.staged++;
    } else {
        check(tp_buf_upload(c, s.x.buf(), images, u8 ? (s.x.numel() + 3) / 4 : s.x.numel()));
        check(tp_buf_upload(c, s.y.buf(), labels, batch));
    }
    if (fs) {
        // the whole loop body (src/train.rs:106-138) as one persistent kernel walking the compiled tape (or, wide models, as
        // a plan of tcgen05 kernels)
        p.before_fused(*this, fs);
        if (u8)
            check(tp_step_run_u8(c, fs, xin, yin, nullptr, nullptr, 0, -1, optimizer->lr(), optimizer->grad_scale(), p.next_result_slot(),
                                 p.next_result_seq()));
        else
            check(tp_step_run(c, fs, xin, yin, nullptr, nullptr, 0, -1, optimizer->lr(), optimizer->grad_scale(), p.next_result_slot(),
                              p.next_result_seq()));
        if (pinned) {
            check(tp_event_record(c, s.done[k]));
            s.done_valid[k] = true;
        }
        optimizer->note_device_step();
        p.after_fused(*this, fs);
        fused_steps_++;
        p.enqueue_result(true);
        return;
    }
    if (s.graph) {
        check(tp_graph_launch(c, s.graph));
        optimizer->note_device_step();                     // the replay advanced t / the parameters on the device
        graph_replays_++;
    } else if (use_graph_ && s.eager_runs >= 1 && model->capturable()) {
        // second iteration of this shape: record it.  Every buffer the step allocates comes from a pool the
        // graph owns, and the closures (which own the activations) are kept alive with the graph.
        check(tp_graph_begin(c));
        try {
            step_body(*this, s.x, s.y, p.loss_view, p.correct_view, p.world);
        } catch (...) {
            tp_graph* g = nullptr;
            tp_graph_end(c, &g);
            tp_graph_destroy(g);
            throw;
        }
        s.keep = Tape::take();
        check(tp_graph_end(c, &s.graph));
        check(tp_graph_launch(c, s.graph));
        graph_replays_++;
    } else {
        step_body(*this, s.x, s.y, p.loss_view, p.correct_view, p.world);
        s.eager_runs++;
    }
    p.enqueue_result();
}

StepResult Trainer::train_batch(const float* images, const float* labels, size_t batch, const Shape& sample_shape) {
    train_batch_async(images, labels, batch, sample_shape, false);
    return fetch();
}

void Trainer::load_dataset(const float* images, const float* labels, size_t n, const Shape& sample_shape, const uint32_t* perm) {
    load_dataset_impl(images, labels, n, sample_shape, perm, false);
}

void Trainer::load_dataset_u8(const uint8_t* images, const float* labels, size_t n, const Shape& sample_shape, const uint32_t* perm) {
    load_dataset_impl(images, labels, n, sample_shape, perm, true);
}

void Trainer::load_dataset_impl(const void* images, const float* labels, size_t n, const Shape& sample_shape, const uint32_t* perm, bool u8) {
    Impl& p = *p_;
    tp_ctx* c = ctx();
    check(tp_sync(c));
    for (auto& kv : p.slots) {                                  // graphs that reference the old dataset
        tp_graph_destroy(kv.second.graph_resident);
        kv.second.graph_resident = nullptr;
        kv.second.keep_resident.clear();
    }
    tp_buf_release(p.ds_images); tp_buf_release(p.ds_labels); tp_buf_release(p.ds_perm); tp_buf_release(p.ds_cursor);
    size_t cols = shape_numel(sample_shape);
    const size_t img_words = u8 ? (n * cols + 3) / 4 : n * cols;     // u8 pixels are stored as they are (4 per 32-bit word)
    check(tp_buf_alloc(c, img_words, &p.ds_images));
    check(tp_buf_alloc(c, n, &p.ds_labels));
    check(tp_buf_alloc(c, n, &p.ds_perm));
    check(tp_buf_alloc(c, 1, &p.ds_cursor));
    check(tp_buf_upload(c, p.ds_images, images, img_words));
    p.ds_u8 = u8;
    check(tp_buf_upload(c, p.ds_labels, labels, n));
    std::vector<uint32_t> ident;
    if (!perm) {
        ident.resize(n);
        std::iota(ident.begin(), ident.end(), 0u);
        perm = ident.data();
    }
    check(tp_buf_upload(c, p.ds_perm, perm, n));
    check(tp_buf_fill(c, p.ds_cursor, 0.0f, 1));
    p.ds_n = n;
    p.ds_cursor_host = 0;
    p.ds_sample = sample_shape;
}

void Trainer::train_batch_resident(size_t batch) {
    Impl& p = *p_;
    if (!p.ds_images) panic("Trainer::train_batch_resident: load_dataset has not been called");
    if (pending() >= kRing) panic("Trainer: %zu steps outstanding; call fetch()", kRing);
    tp_step* fst = use_fused_ ? p.fused_step(*this, batch, p.ds_sample) : nullptr;
    if (p.ds_u8 && !(fst && p.fused_is_wide(fst))) panic("Trainer: a u8 resident dataset is read by the wide step plan only (load f32 pixels for this model)");
    if (fst) {
        // batch rows are gathered out of the resident dataset inside the step kernel; the cursor advances there too
        p.before_fused(*this, fst);
        if (p.ds_u8)
            check(tp_step_run_u8(ctx(), fst, p.ds_images, p.ds_labels, p.ds_perm, p.ds_cursor, (int)p.ds_n, (int)p.ds_cursor_host,
                                 optimizer->lr(), optimizer->grad_scale(), p.next_result_slot(), p.next_result_seq()));
        else
            check(tp_step_run(ctx(), fst, p.ds_images, p.ds_labels, p.ds_perm, p.ds_cursor, (int)p.ds_n, (int)p.ds_cursor_host,
                              optimizer->lr(), optimizer->grad_scale(), p.next_result_slot(), p.next_result_seq()));
        p.ds_cursor_host = (p.ds_cursor_host + batch) % p.ds_n;
        optimizer->note_device_step();
        p.after_fused(*this, fst);
        fused_steps_++;
        p.enqueue_result(true);
        return;
    }
    p.ds_cursor_host = (p.ds_cursor_host + batch) % p.ds_n;
    Shape fs = full_shape(batch, p.ds_sample);
    Slot& s = p.slots[fs];
    if (!s.x.defined()) {
        s.x = Tensor::empty(fs);
        s.y = Tensor::empty({batch});
    }
    tp_ctx* c = ctx();
    int cols = (int)shape_numel(p.ds_sample);
    auto body = [&]() {
        // MNISTDataset::get_batch on the device (src/data/mnist.rs:276-309), then the cursor moves on
        check(tp_gather_batch(c, p.ds_images, p.ds_labels, p.ds_perm, p.ds_cursor, s.x.buf(), s.y.buf(), (int)batch, cols, (int)p.ds_n));
        check(tp_cursor_advance(c, p.ds_cursor, (int)batch, (int)p.ds_n));
        step_body(*this, s.x, s.y, p.loss_view, p.correct_view, p.world);
    };
    if (s.graph_resident) {
        check(tp_graph_launch(c, s.graph_resident));
        optimizer->note_device_step();
        graph_replays_++;
    } else if (use_graph_ && s.eager_runs >= 1 && model->capturable()) {
        check(tp_graph_begin(c));
        try {
            body();
        } catch (...) {
            tp_graph* g = nullptr;
            tp_graph_end(c, &g);
            tp_graph_destroy(g);
            throw;
        }
        s.keep_resident = Tape::take();
        check(tp_graph_end(c, &s.graph_resident));
        check(tp_graph_launch(c, s.graph_resident));
        graph_replays_++;
    } else {
        body();
        s.eager_runs++;
    }
    p.enqueue_result();
}

StepResult Trainer::eval_batch(const float* images, const float* labels, size_t batch, const Shape& sample_shape) {
    // body of Trainer::evaluate (src/train.rs:156-166): forward + loss + accuracy.  The reference records tape
    // nodes here too (no no-grad mode, A10); they are dropped right away.
    Shape fs = full_shape(batch, sample_shape);
    Tensor x = Tensor::from_host(images, fs), y = Tensor::from_host(labels, {batch});
    Tensor logits = model->forward(x);
    Tensor l = loss::cross_entropy_loss(logits, y);
    Tensor cnt = loss::accuracy_count(logits, y);
    StepResult r{l.item(), cnt.item()};
    Tape::reset();
    return r;
}

std::pair<float, float> Trainer::train_epoch(data::DataLoader& loader, size_t max_batches) {        // src/train.rs:98-144
    float total_loss = 0.0f;
    size_t total_correct = 0, total_samples = 0;
    loader.reset();
    size_t num_batches = loader.num_batches();
    if (max_batches && max_batches < num_batches) num_batches = max_batches;
    const Shape sample = loader.sample_shape;
    // raw u8 pixels cross PCIe when the dataset still has them and this model's step is the wide plan (which divides by 255
    // on the device); everything else is fed f32, as the reference's loader produces (src/data/mnist.rs:225)
    bool u8 = false;
    if (loader.dataset().has_u8() && use_fused_ && sample.size() == 1) {
        tp_step* st = p_->fused_step(*this, std::min(loader.batch_size(), loader.dataset().len()), sample);
        u8 = p_->fused_is_wide(st);
    }
    loader.start_prefetch(u8, max_batches);
    std::deque<std::pair<size_t, int>> inflight;                                // {batch size, pinned slot}
    auto drain_one = [&]() {
        StepResult r = fetch();
        // `(acc * batch_size as f32) as usize` with acc = correct / total in f32 (src/loss.rs:289, src/train.rs:117): the
        // reference's round trip through the ratio truncates, so a count may come back one short — reproduced
        const float bsz = (float)inflight.front().first;
        total_correct += (size_t)((r.correct / bsz) * bsz);
        total_samples += inflight.front().first;
        loader.release(inflight.front().second);                                 // its H2D copy finished before its step ran
        inflight.pop_front();
        total_loss += r.loss;                                                    // :127
    };
    data::DataLoader::Batch b;
    while (loader.next_pinned(b)) {
        if (pending() >= kRing - 1) drain_one();
        if (b.u8) train_batch_async_u8(static_cast<const uint8_t*>(b.images), b.labels, b.batch, sample, true);
        else train_batch_async(static_cast<const float*>(b.images), b.labels, b.batch, sample, true);
        inflight.emplace_back(b.batch, b.slot);
    }
    while (pending()) drain_one();
    loader.stop_prefetch();
    if (!total_samples) return {0.0f, 0.0f};
    return {total_loss / (float)num_batches, (float)total_correct / (float)total_samples};
}

std::pair<float, float> Trainer::evaluate(data::DataLoader& loader) {           // src/train.rs:147-172
    float total_loss = 0.0f;
    size_t total_correct = 0, total_samples = 0;
    loader.reset();
    size_t num_batches = loader.num_batches();
    std::vector<float> images, labels;
    size_t batch = 0;
    while (loader.next(images, labels, batch)) {
        StepResult r = eval_batch(images.data(), labels.data(), batch, loader.sample_shape);
        total_correct += (size_t)((r.correct / (float)batch) * (float)batch);    // as in train_epoch (src/train.rs:160)
        total_samples += batch;
        total_loss += r.loss;
    }
    return {total_loss / (float)num_batches, (float)total_correct / (float)total_samples};
}

void Trainer::fit(data::DataLoader& train_loader, data::DataLoader& val_loader, size_t epochs, bool verbose) {   // :175-261
    std::printf("Starting training for %zu epochs\n============================================================\n", epochs);
    for (size_t epoch = 0; epoch < epochs; ++epoch) {
        auto t0 = std::chrono::steady_clock::now();
        if (verbose) std::printf("\nEpoch %zu/%zu\n", epoch + 1, epochs);
        auto [train_loss, train_acc] = train_epoch(train_loader);
        auto [val_loss, val_acc] = evaluate(val_loader);
        if (scheduler) {                                                         // :212-216
            scheduler->step(val_loss);
            optimizer->set_lr(scheduler->get_lr());
        }
        metrics.train_loss.push_back(train_loss);
        metrics.train_acc.push_back(train_acc);
        metrics.val_loss.push_back(val_loss);
        metrics.val_acc.push_back(val_acc);
        metrics.epoch_times.push_back(std::chrono::duration<float>(std::chrono::steady_clock::now() - t0).count());
        if (verbose) {
            std::printf("\nEpoch %zu - Train Loss: %.4f | Train Acc: %.2f%% | Val Loss: %.4f | Val Acc: %.2f%% | Time: %.2fs\n",
                        epoch + 1, train_loss, train_acc * 100.0f, val_loss, val_acc * 100.0f, metrics.epoch_times.back());
            if (scheduler) std::printf("   Learning Rate: %.6f\n", scheduler->get_lr());
        }
        if (val_acc > 0.99f) {                                                   // :246-249
            std::printf("\nReached 99%% validation accuracy! Stopping early.\n");
            break;
        }
    }
    metrics.plot_summary();
}

void Trainer::save_checkpoint(const std::string& path) const {                  // src/train.rs:264-292
    std::ofstream f(path);
    if (!f) panic("save_checkpoint: cannot create %s", path.c_str());
    auto params = model->parameters();
    f << params.size() << "\n";
    char buf[64];
    for (auto& p : params) {
        f << p.shape().size();
        for (size_t d : p.shape()) f << " " << d;
        f << "\n";
        for (float v : p.data()) {
            std::snprintf(buf, sizeof buf, "%.9g", v);                           // round-trips an f32 exactly
            f << buf << "\n";
        }
    }
}

void Trainer::load_checkpoint(const std::string& path) const {
    std::ifstream f(path);
    if (!f) panic("load_checkpoint: cannot open %s", path.c_str());
    auto params = model->parameters();
    size_t count = 0;
    f >> count;
    if (count != params.size()) panic("load_checkpoint: file has %zu parameters, model has %zu", count, params.size());
    for (auto& p : params) {
        size_t nd = 0;
        f >> nd;
        Shape s(nd);
        for (auto& d : s) f >> d;
        if (s != p.shape()) panic("load_checkpoint: parameter shape mismatch");
        std::vector<float> v(p.numel());
        for (auto& x : v) f >> x;
        if (!f) panic("load_checkpoint: truncated file");
        p.set_data(v);
    }
}

}  // namespace train

}  // namespace taper

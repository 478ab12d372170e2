// Runtime: context, stream, caching device allocator, refcounted buffers, CUDA-graph capture.
// Replaces the reference's Vec<f32> allocation (src/tensor.rs:470-478) and mimalloc global
// allocator (src/main.rs:7-10) on the device side.
#include "common.cuh"
#include <cstdarg>
#include <mutex>

namespace tp {

static thread_local std::string g_err;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
}

size_t Allocator::bucket(size_t bytes) {
    if (bytes == 0) bytes = 1;
    if (bytes <= (1u << 20)) return (bytes + 511) & ~(size_t)511;            // 512 B granules
    if (bytes <= (64u << 20)) return (bytes + (1u << 20) - 1) & ~(size_t)((1u << 20) - 1);  // 1 MiB
    return (bytes + (16u << 20) - 1) & ~(size_t)((16u << 20) - 1);           // 16 MiB
}

void* Allocator::alloc(size_t bytes, size_t* cap, Allocator* steal_from, bool relaxed_capture) {
    size_t b = bucket(bytes);
    *cap = b;
    auto it = free_lists.find(b);
    if (it != free_lists.end() && !it->second.empty()) {
        void* p = it->second.back();
        it->second.pop_back();
        in_use += b;
        return p;
    }
    if (steal_from) {                 // graph-private pool: adopt a cached block of the main pool
        auto st = steal_from->free_lists.find(b);
        if (st != steal_from->free_lists.end() && !st->second.empty()) {
            void* p = st->second.back();
            st->second.pop_back();
            steal_from->all_blocks.erase(p);
            steal_from->reserved -= b;
            all_blocks.insert(p);
            reserved += b;
            in_use += b;
            return p;
        }
    }
    // cudaMalloc is not allowed on a thread that is capturing unless the capture mode is relaxed
    cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
    if (relaxed_capture) cudaThreadExchangeStreamCaptureMode(&mode);
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, b);
    if (e != cudaSuccess && !relaxed_capture) {
        cudaGetLastError();
        // give cached blocks back to the driver and retry once
        for (auto& kv : free_lists) {
            for (void* q : kv.second) {
                cudaFree(q);
                reserved -= kv.first;
                all_blocks.erase(q);
            }
            kv.second.clear();
        }
        e = cudaMalloc(&p, b);
    }
    if (relaxed_capture) cudaThreadExchangeStreamCaptureMode(&mode);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    all_blocks.insert(p);
    reserved += b;
    in_use += b;
    return p;
}

void Allocator::free(void* p, size_t cap) {
    free_lists[cap].push_back(p);
    in_use -= cap;
}

void Allocator::release_all() {
    for (void* p : all_blocks) cudaFree(p);
    all_blocks.clear();
    free_lists.clear();
    in_use = reserved = 0;
}

int ensure_scratch(tp_ctx* ctx, size_t bytes) {
    if (ctx->scratch_bytes >= bytes) return TP_OK;
    if (ctx->capturing) {
        set_error("scratch growth (%zu bytes) requested during graph capture; run one eager step first", bytes);
        return TP_ERR_INVALID;
    }
    size_t want = bytes < (8u << 20) ? (8u << 20) : bytes;
    float* p = nullptr;
    // The old block is kept until the context dies: a captured CUDA graph may still point at it.
    if (ctx->scratch) ctx->retired_scratch.push_back(ctx->scratch);
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
    TP_CUDA(cudaMalloc(&p, want));
    ctx->scratch = p;
    ctx->scratch_bytes = want;
    return TP_OK;
}

}  // namespace tp

struct tp_event {
    cudaEvent_t ev = nullptr;
};

struct tp_graph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    uint64_t launches = 0;           // kernels recorded inside the graph
    tp::Allocator* pool = nullptr;   // memory the captured work allocated; owned by the graph
    int device = 0;
};

static void drop_pool(tp::Allocator* pool) {
    if (!pool) return;
    if (pool->live_bufs == 0) {
        pool->release_all();
        delete pool;
    } else {
        pool->orphaned = true;       // a tp_buf allocated during capture outlives the graph
    }
}

extern "C" {

int tp_abi_version(void) { return TP_ABI_VERSION; }
const char* tp_last_error(void) { return tp::g_err.c_str(); }

int tp_device_count(int* count) {
    TP_CHECK_ARG(count, "tp_device_count: NULL out pointer");
    TP_CUDA(cudaGetDeviceCount(count));
    return TP_OK;
}

int tp_ctx_create(int device, tp_ctx** out) {
    TP_CHECK_ARG(out, "tp_ctx_create: NULL out pointer");
    int count = 0;
    TP_CUDA(cudaGetDeviceCount(&count));
    TP_CHECK_ARG(device >= 0 && device < count, "tp_ctx_create: device %d out of range (%d visible)", device, count);
    TP_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    TP_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        tp::set_error("tp_ctx_create: device %d is sm_%d%d; this library contains sm_100a code only",
                      device, prop.major, prop.minor);
        return TP_ERR_UNSUPPORTED;
    }
    tp_ctx* ctx = new tp_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    TP_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    TP_CUDA(cudaMalloc(&ctx->dev_error, sizeof(int)));
    TP_CUDA(cudaMemsetAsync(ctx->dev_error, 0, sizeof(int), ctx->stream));
    TP_CUDA(cudaMalloc(&ctx->dev_counters, tp::kNumCounters * sizeof(int)));
    TP_CUDA(cudaMemsetAsync(ctx->dev_counters, 0, tp::kNumCounters * sizeof(int), ctx->stream));
    ctx->pinned_bytes = 8u << 20;
    TP_CUDA(cudaMallocHost(&ctx->pinned, ctx->pinned_bytes));
    TP_CUDA(cudaEventCreateWithFlags(&ctx->pinned_ev, cudaEventDisableTiming));
    *out = ctx;
    return TP_OK;
}

int tp_ctx_destroy(tp_ctx* ctx) {
    if (!ctx) return TP_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamDestroy(ctx->copy_stream);
    }
    tp_comm_destroy(ctx);
    tp::gemm_tc_destroy(ctx);
    tp::gemm_bx3_destroy(ctx);
    ctx->alloc.release_all();
    if (ctx->scratch) cudaFree(ctx->scratch);
    for (float* q : ctx->retired_scratch) cudaFree(q);
    if (ctx->dev_error) cudaFree(ctx->dev_error);
    if (ctx->dev_counters) cudaFree(ctx->dev_counters);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->pinned_ev) cudaEventDestroy(ctx->pinned_ev);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return TP_OK;
}

int tp_sync(tp_ctx* ctx) {
    TP_CHECK_ARG(ctx, "tp_sync: NULL ctx");
    if (ctx->copy_stream) TP_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    TP_CUDA(cudaStreamSynchronize(ctx->stream));
    return TP_OK;
}

void* tp_ctx_stream(tp_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int tp_ctx_device(tp_ctx* ctx) { return ctx ? ctx->device : -1; }
int tp_ctx_sm_count(tp_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

int tp_ctx_mem_stats(tp_ctx* ctx, size_t* in_use, size_t* reserved) {
    TP_CHECK_ARG(ctx, "tp_ctx_mem_stats: NULL ctx");
    if (in_use) *in_use = ctx->alloc.in_use;
    if (reserved) *reserved = ctx->alloc.reserved;
    return TP_OK;
}

int tp_ctx_launch_count(tp_ctx* ctx, uint64_t* count) {
    TP_CHECK_ARG(ctx && count, "tp_ctx_launch_count: NULL argument");
    *count = ctx->launches;
    return TP_OK;
}

int tp_ctx_device_error_async(tp_ctx* ctx, int* host_flag) {
    TP_CHECK_ARG(ctx && host_flag, "tp_ctx_device_error_async: NULL argument");
    TP_CUDA(cudaMemcpyAsync(host_flag, ctx->dev_error, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    return TP_OK;
}

int tp_ctx_device_error(tp_ctx* ctx, int* flag) {
    TP_CHECK_ARG(ctx && flag, "tp_ctx_device_error: NULL argument");
    TP_CUDA(cudaMemcpyAsync(flag, ctx->dev_error, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    TP_CUDA(cudaStreamSynchronize(ctx->stream));
    if (*flag) TP_CUDA(cudaMemsetAsync(ctx->dev_error, 0, sizeof(int), ctx->stream));
    return TP_OK;
}

int tp_buf_alloc(tp_ctx* ctx, size_t n, tp_buf** out) {
    TP_CHECK_ARG(ctx && out, "tp_buf_alloc: NULL argument");
    cudaSetDevice(ctx->device);
    size_t cap = 0;
    tp::Allocator* pool = ctx->capture_pool ? ctx->capture_pool : &ctx->alloc;
    void* p = ctx->capture_pool ? pool->alloc(n * sizeof(float), &cap, &ctx->alloc, true)
                                : pool->alloc(n * sizeof(float), &cap);
    if (!p) {
        tp::set_error("tp_buf_alloc: out of device memory allocating %zu bytes (reserved %zu)",
                      n * sizeof(float), ctx->alloc.reserved);
        return TP_ERR_OOM;
    }
    tp_buf* b = new tp_buf();
    b->ctx = ctx;
    b->ptr = (float*)p;
    b->n = n;
    b->cap = cap;
    b->pool = pool;
    pool->live_bufs++;
    *out = b;
    return TP_OK;
}

int tp_buf_wrap(tp_ctx* ctx, void* device_ptr, size_t n, tp_buf** out) {
    TP_CHECK_ARG(ctx && out && device_ptr, "tp_buf_wrap: NULL argument");
    tp_buf* b = new tp_buf();
    b->ctx = ctx;
    b->ptr = (float*)device_ptr;
    b->n = n;
    b->external = true;
    *out = b;
    return TP_OK;
}

int tp_buf_slice(tp_buf* parent, size_t offset, size_t n, tp_buf** out) {
    TP_CHECK_ARG(parent && out, "tp_buf_slice: NULL argument");
    TP_CHECK_ARG(offset + n <= parent->n, "tp_buf_slice: [%zu, %zu) exceeds parent length %zu", offset, offset + n, parent->n);
    tp_buf* b = new tp_buf();
    b->ctx = parent->ctx;
    b->ptr = parent->ptr + offset;
    b->n = n;
    b->parent = parent;
    parent->rc.fetch_add(1);
    *out = b;
    return TP_OK;
}

int tp_buf_retain(tp_buf* buf) {
    TP_CHECK_ARG(buf, "tp_buf_retain: NULL buffer");
    buf->rc.fetch_add(1);
    return TP_OK;
}

int tp_buf_release(tp_buf* buf) {
    if (!buf) return TP_OK;
    if (buf->rc.fetch_sub(1) == 1) {
        if (buf->parent) tp_buf_release(buf->parent);
        else if (!buf->external) {
            tp::Allocator* pool = buf->pool;
            pool->free(buf->ptr, buf->cap);
            pool->live_bufs--;
            if (pool->orphaned && pool->live_bufs == 0) {
                pool->release_all();
                delete pool;
            }
        }
        delete buf;
    }
    return TP_OK;
}

void* tp_buf_ptr(const tp_buf* buf) { return buf ? buf->ptr : nullptr; }
size_t tp_buf_len(const tp_buf* buf) { return buf ? buf->n : 0; }

int tp_buf_upload(tp_ctx* ctx, tp_buf* dst, const void* host, size_t n) {
    TP_CHECK_ARG(ctx && host, "tp_buf_upload: NULL argument");
    TP_NEED(dst, n, "dst");
    // a captured copy out of the shared staging ring would replay whatever the ring holds at replay time
    TP_CHECK_ARG(!ctx->capturing, "tp_buf_upload: host data cannot be uploaded while a CUDA graph is being captured");
    const char* src = (const char*)host;
    char* d = (char*)dst->ptr;
    size_t bytes = n * sizeof(float);
    // stage pageable memory through the pinned ring so the copy is a true async DMA
    while (bytes) {
        size_t chunk = bytes < ctx->pinned_bytes ? bytes : ctx->pinned_bytes;
        TP_CUDA(cudaEventSynchronize(ctx->pinned_ev));
        memcpy(ctx->pinned, src, chunk);
        TP_CUDA(cudaMemcpyAsync(d, ctx->pinned, chunk, cudaMemcpyHostToDevice, ctx->stream));
        TP_CUDA(cudaEventRecord(ctx->pinned_ev, ctx->stream));
        src += chunk;
        d += chunk;
        bytes -= chunk;
    }
    return TP_OK;
}

int tp_buf_upload_pinned(tp_ctx* ctx, tp_buf* dst, const void* pinned_host, size_t n) {
    TP_CHECK_ARG(ctx && pinned_host, "tp_buf_upload_pinned: NULL argument");
    TP_NEED(dst, n, "dst");
    TP_CUDA(cudaMemcpyAsync(dst->ptr, pinned_host, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    return TP_OK;
}

// ---- copy stream: host->device input prefetch that overlaps the compute stream --------------------------------------
static int ensure_copy_stream(tp_ctx* ctx) {
    if (ctx->copy_stream) return TP_OK;
    cudaSetDevice(ctx->device);
    TP_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    return TP_OK;
}

int tp_copy_upload_pinned(tp_ctx* ctx, tp_buf* dst, const void* pinned_host, size_t n) {
    TP_CHECK_ARG(ctx && pinned_host, "tp_copy_upload_pinned: NULL argument");
    TP_NEED(dst, n, "dst");
    int rc = ensure_copy_stream(ctx);
    if (rc) return rc;
    TP_CUDA(cudaMemcpyAsync(dst->ptr, pinned_host, n * sizeof(float), cudaMemcpyHostToDevice, ctx->copy_stream));
    return TP_OK;
}

int tp_copy_event_record(tp_ctx* ctx, tp_event* ev) {
    TP_CHECK_ARG(ctx && ev, "tp_copy_event_record: NULL argument");
    int rc = ensure_copy_stream(ctx);
    if (rc) return rc;
    TP_CUDA(cudaEventRecord(ev->ev, ctx->copy_stream));
    return TP_OK;
}

int tp_copy_wait_event(tp_ctx* ctx, tp_event* ev) {
    TP_CHECK_ARG(ctx && ev, "tp_copy_wait_event: NULL argument");
    int rc = ensure_copy_stream(ctx);
    if (rc) return rc;
    TP_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ev->ev, 0));
    return TP_OK;
}

int tp_stream_wait_event(tp_ctx* ctx, tp_event* ev) {
    TP_CHECK_ARG(ctx && ev, "tp_stream_wait_event: NULL argument");
    TP_CHECK_ARG(!ctx->capturing, "tp_stream_wait_event: not inside a graph capture");
    TP_CUDA(cudaStreamWaitEvent(ctx->stream, ev->ev, 0));
    return TP_OK;
}

int tp_buf_download(tp_ctx* ctx, const tp_buf* src, void* host, size_t n) {
    TP_CHECK_ARG(ctx && host, "tp_buf_download: NULL argument");
    TP_NEED(src, n, "src");
    TP_CUDA(cudaMemcpyAsync(host, src->ptr, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    TP_CUDA(cudaStreamSynchronize(ctx->stream));
    return TP_OK;
}

int tp_buf_copy(tp_ctx* ctx, tp_buf* dst, const tp_buf* src, size_t n) {
    TP_CHECK_ARG(ctx, "tp_buf_copy: NULL ctx");
    TP_NEED(dst, n, "dst");
    TP_NEED(src, n, "src");
    TP_CUDA(cudaMemcpyAsync(dst->ptr, src->ptr, n * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    return TP_OK;
}

int tp_buf_download_async(tp_ctx* ctx, const tp_buf* src, void* pinned_host, size_t n) {
    TP_CHECK_ARG(ctx && pinned_host, "tp_buf_download_async: NULL argument");
    TP_NEED(src, n, "src");
    TP_CUDA(cudaMemcpyAsync(pinned_host, src->ptr, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    return TP_OK;
}

int tp_event_create(tp_ctx* ctx, tp_event** out) {
    TP_CHECK_ARG(ctx && out, "tp_event_create: NULL argument");
    cudaSetDevice(ctx->device);
    tp_event* e = new tp_event();
    cudaError_t rc = cudaEventCreate(&e->ev);
    if (rc != cudaSuccess) {
        delete e;
        tp::set_error("cudaEventCreate failed: %s", cudaGetErrorString(rc));
        return TP_ERR_CUDA;
    }
    *out = e;
    return TP_OK;
}

int tp_event_record(tp_ctx* ctx, tp_event* ev) {
    TP_CHECK_ARG(ctx && ev, "tp_event_record: NULL argument");
    TP_CUDA(cudaEventRecord(ev->ev, ctx->stream));
    return TP_OK;
}

int tp_event_sync(tp_event* ev) {
    TP_CHECK_ARG(ev, "tp_event_sync: NULL event");
    TP_CUDA(cudaEventSynchronize(ev->ev));
    return TP_OK;
}

int tp_event_elapsed_ms(tp_event* start, tp_event* stop, float* ms) {
    TP_CHECK_ARG(start && stop && ms, "tp_event_elapsed_ms: NULL argument");
    TP_CUDA(cudaEventElapsedTime(ms, start->ev, stop->ev));
    return TP_OK;
}

int tp_event_destroy(tp_event* ev) {
    if (ev) {
        cudaEventDestroy(ev->ev);
        delete ev;
    }
    return TP_OK;
}

int tp_host_alloc_pinned(size_t bytes, void** out) {
    TP_CHECK_ARG(out, "tp_host_alloc_pinned: NULL out pointer");
    TP_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
    return TP_OK;
}

int tp_host_free_pinned(void* p) {
    if (p) TP_CUDA(cudaFreeHost(p));
    return TP_OK;
}

int tp_graph_begin(tp_ctx* ctx) {
    TP_CHECK_ARG(ctx && !ctx->capturing, "tp_graph_begin: NULL ctx or capture already active");
    cudaSetDevice(ctx->device);
    TP_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    ctx->capturing = true;
    ctx->capture_pool = new tp::Allocator();
    return TP_OK;
}

int tp_graph_end(tp_ctx* ctx, tp_graph** out) {
    TP_CHECK_ARG(ctx && out && ctx->capturing, "tp_graph_end: no capture active");
    ctx->capturing = false;
    tp_graph* g = new tp_graph();
    g->pool = ctx->capture_pool;
    g->device = ctx->device;
    ctx->capture_pool = nullptr;
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &g->graph);
    if (e != cudaSuccess || !g->graph) {
        cudaGetLastError();
        drop_pool(g->pool);
        delete g;
        tp::set_error("cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
        return TP_ERR_CUDA;
    }
    e = cudaGraphInstantiate(&g->exec, g->graph, 0);
    if (e != cudaSuccess) {
        cudaGetLastError();
        cudaGraphDestroy(g->graph);
        drop_pool(g->pool);
        delete g;
        tp::set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
        return TP_ERR_CUDA;
    }
    size_t nodes = 0;
    cudaGraphGetNodes(g->graph, nullptr, &nodes);
    std::vector<cudaGraphNode_t> ns(nodes);
    if (nodes) cudaGraphGetNodes(g->graph, ns.data(), &nodes);
    for (auto nd : ns) {
        cudaGraphNodeType t;
        if (cudaGraphNodeGetType(nd, &t) == cudaSuccess && t == cudaGraphNodeTypeKernel) g->launches++;
    }
    *out = g;
    return TP_OK;
}

int tp_graph_launch(tp_ctx* ctx, tp_graph* g) {
    TP_CHECK_ARG(ctx && g && g->exec, "tp_graph_launch: NULL argument");
    TP_CUDA(cudaGraphLaunch(g->exec, ctx->stream));
    ctx->launches += g->launches;
    return TP_OK;
}

int tp_graph_destroy(tp_graph* g) {
    if (!g) return TP_OK;
    cudaSetDevice(g->device);
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    drop_pool(g->pool);               // the caller synchronises before destroying a graph in flight
    delete g;
    return TP_OK;
}

}  // extern "C"

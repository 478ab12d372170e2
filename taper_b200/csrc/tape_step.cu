// tape_step.cu — a whole MLP training step as ONE persistent kernel: the "device tape".
//
// The reference's step (src/train.rs:106-138: Tape::reset, forward, cross_entropy_loss, accuracy, backward,
// optimizer.step, zero_grad) on the small MNIST MLPs is ~0.2 GFLOP: on a B200 every kernel of the eager / CUDA-graph
// path is launch- and latency-bound (5-12 us each, 12 launches).  Here the host compiles the recorded tape of such a
// model into a static job list and ONE persistent kernel (one CTA per SM) walks it phase by phase with a grid barrier
// between dependent phases (consecutive steps are chained by programmatic dependent launch):
//     fwd GEMMs (gathering the batch rows straight out of the resident dataset)  ->  head (fold of the fwd split-K partials,
//     logits, log-softmax, NLL, accuracy, dlogits, dX of the head, ReLU mask)  ->  backward GEMMs (dW with the bias
//     column-sum riding along, dX with the ReLU mask in the epilogue) + loss fold  ->  SGD / Adam / AdamW over the flat
//     arena (fold of the dW split-K partials; with data parallelism also the gradient exchange over NVLink peer memory).
// Arithmetic is exact fp32 FMA on the CUDA cores: at these sizes (512x128x784) the tensor-core path is bound by its own
// TMA/TMEM/commit latency chain, not by math (measured: 12 us for the tcgen05 kernel alone).  Larger models keep the
// tcgen05 GEMM path (gemm_tc.cu); tp_step_supported() draws the line.
// Deterministic: split-K partials are summed in split order by their consumer (head / optimizer job, or the last CTA of a
// tile where the result must be materialised), peer gradient slices in rank order; no float atomics.
// The step's {loss, #correct} are also published into a mapped pinned host slot followed by a sequence word, so the host
// needs neither a D2H copy nor a CUDA event per step.
#include "common.cuh"
#include <cmath>
#include <cstdlib>
#include <type_traits>

namespace {

constexpr int kThreads = 256;
constexpr int BM = 64, BN = 64, BK = 16, LDS = BM + 4;        // smem tile row pitch (floats)
constexpr int SUB = 4, SK = SUB * BK;                          // a pipeline stage = 4 sub-tiles = 64 k: one memory round trip
constexpr int kMaxOut = 16;
constexpr int kMaxJobs = 48, kMaxPhases = 20;
constexpr int kMaxBatch = 4096;
constexpr int kHeadRows = 8;                                   // one warp per row
constexpr int kMaxWorld = 8;
constexpr long long kPeerSpinLimit = 40000000000LL;            // ~20 s: a peer that never arrives must not hang this GPU for ever
constexpr int kProfSlots = 2 + 2 * kMaxPhases;               // entry, setup done, then {work done, barrier passed} per phase
constexpr long long kSpinLimit = 400000000LL;                  // ~0.2 s: a barrier that never completes must not hang the GPU

enum { JOB_GEMM = 0, JOB_HEAD = 1, JOB_LOSS = 2, JOB_OPT = 3 };
enum { H_T = 0, H_LR, H_B1, H_B2, H_EPS, H_WD, H_SS, H_DECAY, H_COUNT };       // same layout as optim.cu

struct Job {
    int kind, items;
    // GEMM  C[M,N] = A[M,K] * B[K,N]   (A element (m,k): a_kc ? A[row(m)*lda + k] : A[row(k)*lda + m]; same for B with n)
    const float* A; const float* B; float* C;
    int lda, ldb, ldc;
    int a_kc, b_kc;
    int a_input, b_input;            // operand is the step's input batch X (rows go through the gather index)
    int M, N, K, m_store;
    int tiles_m, tiles_n, splits, kchunk;
    const float* bias; int relu;
    const float* mask; int ldmask;   // C *= [mask > 0]
    float* partial;                  // [splits][tiles_m*BM][N]
    float* colsum;                   // optional [M]: sum_k A(m,k)  (bias gradient rides on the dW GEMM)
    float* cs_partial;               // [splits][tiles_m*BM]
    int* tickets;                    // [tiles_m*tiles_n]
    int defer_fold;                  // leave the split-K partials unfolded: the consumer (head / optimizer job) sums them
    // HEAD  (A = activations [M=batch, K=in], B = W [N=out, in], bias)
    const float* a_part;             // activations still as split-K partials [a_splits][a_stride] (+ a_bias, a_relu); the head
    int a_splits, a_relu;            //   folds them itself and materialises the rows into a_out
    long long a_stride;
    const float* a_bias;
    float* a_out;
    float* dlog;                     // [batch, 16]  zero padded
    float* dz;                       // optional [batch, in]
    int dz_mask;                     // dz *= [A > 0]
    float* nll; float* hit;          // [batch]
    float inv_b;
    // LOSS: nll, hit, M -> result {loss, correct}
    float* result;
    // OPT (one job per parameter tensor): gradient = g, or the sum over g_splits partials at stride g_stride
    float* p; float* g; float* m; float* v; int n4;
    const float* g_part; int g_splits; long long g_stride;
    long long g_off;                 // offset of this tensor in the gradient arena (peer windows share the layout)
};

struct StepParams {
    const Job* jobs;
    int n_jobs, n_phases;
    int phase_first[kMaxPhases + 1];
    const float* x;                  // [batch, in] or the resident dataset [n_perm, in]
    const float* labels;             // [batch] or [n_perm]
    const int* perm;                 // NULL: x / labels are the batch itself
    int* cursor;
    int n_perm, batch;
    int cursor_value;                // host mirror of *cursor (>= 0), or -1: read the device word
    float* result_host;              // optional mapped pinned {loss, correct, seq} slot of this step (no separate D2H copy)
    unsigned int result_seq;         // written (as bits) to result_host[2] after loss and correct: the host polls it
    unsigned int* bar;               // arrival counter of the grid barrier
    unsigned int bar_base;           // its value when this launch starts (tracked by the host)
    int opt_kind;                    // 0 SGD, 1 Adam, 2 AdamW
    float sgd_lr, grad_scale;
    float* hyper;
    int* err;
    long long* prof;                 // optional [grid][kProfSlots] SM-clock stamps (tp_step_set_profile)
    // data-parallel gradient exchange over NVLink peer memory (world > 1), fused into the optimizer phase: the CTA that owns
    // a slice of the arena pushes its local gradient slice into every peer's window, raises a per-slice flag there, waits for
    // the peers' flags on the same slice and sums the world's slices in rank order.  No grid-wide wait is involved.
    int world, rank;
    unsigned int xseq;               // this step's flag value (monotonic, identical on every rank)
    int x_items;                     // flags per source rank
    unsigned int* my_flags;          // [world][x_items] in the local window
    const float* my_slots;           // [world][arena_len] in the local window: this step's gradient buffers, one per source rank
    long long x_arena;               // arena_len
    unsigned int* peer_flags[kMaxWorld];     // row `rank` of every peer's flag array
    float* peer_slot[kMaxWorld];             // slot `rank` of every peer's window (this step's parity)
};

struct AdamArgs {
    float step_size, beta1, beta2, eps, weight_decay, grad_scale, decay_factor;
    int decoupled;
};

// identical to optim.cu's adam_elem (expression order follows src/optim.rs:93-110; -fmad=false)
__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, const AdamArgs& a) {
    if (a.decoupled) p *= a.decay_factor;
    if (a.grad_scale != 1.0f) g *= a.grad_scale;
    float gg = g + a.weight_decay * p;
    m = a.beta1 * m + (1.0f - a.beta1) * gg;
    v = a.beta2 * v + (1.0f - a.beta2) * gg * gg;
    p -= a.step_size * m / (sqrtf(v) + a.eps);
}

__device__ __forceinline__ float powi_dev(float a, int b) {     // f32::powi, as optim.cu
    float r = 1.0f;
    unsigned int e = (unsigned int)b;
    while (true) {
        if (e & 1u) r *= a;
        e >>= 1;
        if (e == 0) break;
        a *= a;
    }
    return r;
}

__device__ __forceinline__ unsigned int class_of(float t) {     // `t as usize` (src/loss.rs:160)
    if (!(t > 0.0f)) return 0u;
    if (t >= 4294967040.0f) return 0xffffffffu;
    return (unsigned int)t;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// Packed fp32 FMA (Blackwell FFMA2, PTX fma.rn.f32x2): two IEEE fused multiply-adds per issue slot, bit-identical to two
// fmaf calls.  The multiply loops of this kernel are issue-bound, so this halves their instruction count.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void fma2(f32x2& d, f32x2 a, f32x2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }

// ---- grid barrier: one monotonically increasing arrival counter.  The host passes the counter value at launch
// (bar_base); barrier k of this launch completes when the counter reaches bar_base + (k + 1) * gridDim.x (wrap-safe
// compare).  Arrival is a fire-and-forget release reduction, so a CTA pays one poll round trip, not an atomic's too.
__device__ __forceinline__ bool grid_sync(unsigned int* bar, unsigned int target, int* err, int* s_ok) {
    __syncthreads();
    if (threadIdx.x == 0) {
        *s_ok = 1;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(bar) : "memory");
        unsigned int cur;
        const long long t0 = clock64();
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cur) : "l"(bar) : "memory");
            if ((int)(cur - target) < 0 && clock64() - t0 > kSpinLimit) { atomicExch(err, 2); *s_ok = 0; break; }
        } while ((int)(cur - target) < 0);
    }
    __syncthreads();
    return *s_ok != 0;
}

// ---- GEMM tile loader: one float4 per thread per operand per BK step ------------------------------------------------
// KC (operand stored K-contiguous or not) is a run-time, CTA-uniform flag: one copy of the multiply / epilogue / fold code
// serves all three GEMM flavours (the kernel is instruction-fetch sensitive).
__device__ __forceinline__ float4 fetch(const bool KC, const float* __restrict__ P, int ld, const int* __restrict__ ridx, int mn0,
                                        int mn_lim, int k0, int kend, int tid) {
    float4 z = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (KC) {
        const int mn = mn0 + (tid >> 2), gk = k0 + ((tid & 3) << 2);
        if (mn < mn_lim && gk < kend) {
            const int row = ridx ? ridx[mn] : mn;
            z = ldcg4(P + (size_t)row * ld + gk);
        }
    } else {
        const int gk = k0 + (tid >> 4), mn = mn0 + ((tid & 15) << 2);
        if (gk < kend && mn < mn_lim) {
            const int row = ridx ? ridx[gk] : gk;
            z = ldcg4(P + (size_t)row * ld + mn);
        }
    }
    return z;
}

__device__ __forceinline__ void stash(const bool KC, float* __restrict__ S, float4 v, int tid) {          // S: [BK][LDS]
    if (KC) {
        const int mn = tid >> 2, kq = (tid & 3) << 2;
        S[(kq + 0) * LDS + mn] = v.x;
        S[(kq + 1) * LDS + mn] = v.y;
        S[(kq + 2) * LDS + mn] = v.z;
        S[(kq + 3) * LDS + mn] = v.w;
    } else {
        const int k = tid >> 4, mq = (tid & 15) << 2;
        *reinterpret_cast<float4*>(S + k * LDS + mq) = v;
    }
}

__device__ __forceinline__ void epilogue_store(const Job& j, int gm, int gn, float4 v) {
    if (j.bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(j.bias + gn));
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    if (j.relu) { v.x = fmaxf(v.x, 0.0f); v.y = fmaxf(v.y, 0.0f); v.z = fmaxf(v.z, 0.0f); v.w = fmaxf(v.w, 0.0f); }
    if (j.mask) {
        const float4 k = ldcg4(j.mask + (size_t)gm * j.ldmask + gn);
        v.x = k.x > 0.0f ? v.x : 0.0f; v.y = k.y > 0.0f ? v.y : 0.0f;
        v.z = k.z > 0.0f ? v.z : 0.0f; v.w = k.w > 0.0f ? v.w : 0.0f;
    }
    *reinterpret_cast<float4*>(j.C + (size_t)gm * j.ldc + gn) = v;
}

__device__ void gemm_item(const bool AKC, const bool BKC, const Job& j, int item, const StepParams& P, const int* __restrict__ ridx, float* __restrict__ As,
                          float* __restrict__ Bs, int* s_flag) {
    const int tid = threadIdx.x;
    const int tiles = j.tiles_m * j.tiles_n;
    const int split = item / tiles, tile = item - split * tiles;
    const int tm = tile / j.tiles_n, tn = tile - tm * j.tiles_n;
    const int m0 = tm * BM, n0 = tn * BN;
    const int kbeg = split * j.kchunk, kend = min(j.K, kbeg + j.kchunk);
    const float* A = j.a_input ? P.x : j.A;
    const float* B = j.b_input ? P.x : j.B;
    const int* ra = j.a_input ? ridx : nullptr;
    const int* rb = j.b_input ? ridx : nullptr;
    const int tx = tid & 15, ty = tid >> 4;
    const bool do_cs = (!AKC) && j.colsum && tn == 0 && tid < BM;
#define TP_GSTAMP(i) do { if (P.prof && tid == 0) P.prof[(size_t)blockIdx.x * kProfSlots + 24 + (AKC ? 0 : 8) + (i)] = clock64(); } while (0)
    TP_GSTAMP(0);

    f32x2 acc2[4][2];                                          // 4 x 4 accumulators as packed pairs along n
#pragma unroll
    for (int a = 0; a < 4; ++a) { acc2[a][0] = 0ull; acc2[a][1] = 0ull; }
    float cs = 0.0f;

    // Stage = SK (64) k-values per operand held in registers while the previous stage is multiplied out of shared
    // memory: one global-memory round trip per 64 k instead of per 16 (the loop is latency-, not bandwidth-bound).
    // The first TWO stages are requested before anything waits (the gathered rows come from HBM / cold TLB entries, ~2 us):
    // later stages are requested right after the previous one has been written to shared memory.
    const int ns = (kend - kbeg + SK - 1) / SK;
    float4 ra4[SUB], rb4[SUB];                                 // the stage that will be written to shared memory next
    {
        float4 r0a[SUB], r0b[SUB];
#pragma unroll
        for (int u = 0; u < SUB; ++u) {
            r0a[u] = fetch(AKC, A, j.lda, ra, m0, j.M, kbeg + u * BK, kend, tid);
            r0b[u] = fetch(BKC, B, j.ldb, rb, n0, j.N, kbeg + u * BK, kend, tid);
        }
        if (ns > 1) {
#pragma unroll
            for (int u = 0; u < SUB; ++u) {
                ra4[u] = fetch(AKC, A, j.lda, ra, m0, j.M, kbeg + SK + u * BK, kend, tid);
                rb4[u] = fetch(BKC, B, j.ldb, rb, n0, j.N, kbeg + SK + u * BK, kend, tid);
            }
        }
        TP_GSTAMP(1);
#pragma unroll
        for (int u = 0; u < SUB; ++u) {
            stash(AKC, As + u * (BK * LDS), r0a[u], tid);
            stash(BKC, Bs + u * (BK * LDS), r0b[u], tid);
        }
    }
    __syncthreads();
    TP_GSTAMP(2);
    for (int st = 0; st < ns; ++st) {
        const int cur = st & 1;
        const int k0 = kbeg + st * SK;
        const float* as = As + cur * (SK * LDS);
        const float* bs = Bs + cur * (SK * LDS);
        const int klen = min(SK, kend - k0);
#pragma unroll 8
        for (int kk = 0; kk < klen; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(as + kk * LDS + ty * 4);
            const float4 b4 = *reinterpret_cast<const float4*>(bs + kk * LDS + tx * 4);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
            const f32x2 b01 = pack2(b4.x, b4.y), b23 = pack2(b4.z, b4.w);
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const f32x2 aa = pack2(av[a], av[a]);
                fma2(acc2[a][0], aa, b01);
                fma2(acc2[a][1], aa, b23);
            }
        }
        if (do_cs) {
            for (int kk = 0; kk < klen; ++kk) cs += as[kk * LDS + tid];         // k ascending
        }
        if (st + 1 < ns) {
#pragma unroll
            for (int u = 0; u < SUB; ++u) {
                stash(AKC, As + (cur ^ 1) * (SK * LDS) + u * (BK * LDS), ra4[u], tid);
                stash(BKC, Bs + (cur ^ 1) * (SK * LDS) + u * (BK * LDS), rb4[u], tid);
            }
            if (st + 2 < ns) {
#pragma unroll
                for (int u = 0; u < SUB; ++u) {
                    ra4[u] = fetch(AKC, A, j.lda, ra, m0, j.M, k0 + 2 * SK + u * BK, kend, tid);
                    rb4[u] = fetch(BKC, B, j.ldb, rb, n0, j.N, k0 + 2 * SK + u * BK, kend, tid);
                }
            }
        }
        __syncthreads();
    }

    TP_GSTAMP(3);
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        unpack2(acc2[a][0], acc[a][0], acc[a][1]);
        unpack2(acc2[a][1], acc[a][2], acc[a][3]);
    }
    const int gn = n0 + tx * 4;
    if (j.splits == 1) {
        if (gn < j.N) {
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int gm = m0 + ty * 4 + a;
                if (gm < j.m_store) epilogue_store(j, gm, gn, make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]));
            }
        }
        if (do_cs && m0 + tid < j.m_store) j.colsum[m0 + tid] = cs;
        return;
    }
    // split-K: park the partial tile; the last CTA to arrive folds all splits in split order
    const int mpad = j.tiles_m * BM;
    if (gn < j.N) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int gm = m0 + ty * 4 + a;
            *reinterpret_cast<float4*>(j.partial + ((size_t)split * mpad + gm) * j.N + gn) =
                make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
        }
    }
    if (do_cs) j.cs_partial[(size_t)split * mpad + m0 + tid] = cs;
    TP_GSTAMP(4);
    if (j.defer_fold) return;                              // folded by the consumer after the grid barrier
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int prev = atomicAdd(j.tickets + tile, 1);
        const int last = (prev == j.splits - 1);
        if (last) j.tickets[tile] = 0;
        *s_flag = last;
    }
    __syncthreads();
    const int last = *s_flag;
    __syncthreads();                                       // s_flag may be rewritten by the next item
    if (!last) return;
    __threadfence();
    if (gn < j.N) {
        // all loads of a batch are issued before the first add: the fold costs ceil(splits/4) round trips, not 4*splits
        float4 s4[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) s4[a] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        const float* pbase = j.partial + (size_t)(m0 + ty * 4) * j.N + gn;
        const size_t zstride = (size_t)mpad * j.N;
#pragma unroll 1
        for (int z0 = 0; z0 < j.splits; z0 += 4) {
            float4 q[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int a = 0; a < 4; ++a)
                    q[u][a] = (z0 + u < j.splits) ? ldcg4(pbase + (size_t)(z0 + u) * zstride + (size_t)a * j.N)
                                                  : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
            for (int u = 0; u < 4; ++u)                        // split order
#pragma unroll
                for (int a = 0; a < 4; ++a) { s4[a].x += q[u][a].x; s4[a].y += q[u][a].y; s4[a].z += q[u][a].z; s4[a].w += q[u][a].w; }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int gm = m0 + ty * 4 + a;
            if (gm < j.m_store) epilogue_store(j, gm, gn, s4[a]);
        }
    }
    if (do_cs && m0 + tid < j.m_store) {
        float s = 0.0f;
        for (int z0 = 0; z0 < j.splits; z0 += 8) {
            float q[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) q[u] = (z0 + u < j.splits) ? __ldcg(j.cs_partial + (size_t)(z0 + u) * mpad + m0 + tid) : 0.0f;
#pragma unroll
            for (int u = 0; u < 8; ++u) s += q[u];
        }
        j.colsum[m0 + tid] = s;
    }
}

// ---- classifier head: one warp per batch row ----------------------------------------------------------------------
// logits = a.W^T + b (src/nn.rs:54-60), log_softmax + NLL (src/loss.rs:101-126, 152-165), accuracy hit (:271-290),
// dlogits = (exp(logp) - onehot) * (1/B) (src/loss.rs:174-191 with g(loss) = 1), dA = dlogits.W (src/ops.rs:254-265),
// optionally masked by the producing ReLU (src/ops.rs:358-370).
__device__ void head_item(const Job& j, int item, const StepParams& P, const int* __restrict__ ridx) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int r = item * kHeadRows + wid;
    if (r >= j.M) return;
    const int in_f = j.K, out_f = j.N, in4 = in_f >> 2;
    const float* arow = (j.a_input ? P.x + (size_t)(ridx ? ridx[r] : r) * j.lda : j.A + (size_t)r * j.lda);
    const float* W = j.B;
    const float* afin = j.a_part ? j.a_out + (size_t)r * in_f : arow;      // the materialised activation row
    // issue everything that does not depend on the activations first: label, head bias (their latency hides behind the fold)
    const float t = __ldg(P.labels + (P.perm ? ridx[r] : r));
    float hb[kMaxOut];
#pragma unroll
    for (int o = 0; o < kMaxOut; ++o) hb[o] = (j.bias && o < out_f) ? __ldg(j.bias + o) : 0.0f;
    float4 a_first = make_float4(0.0f, 0.0f, 0.0f, 0.0f);    // chunk `lane` of the row (the only one when in <= 128)
    float acc[kMaxOut];
#pragma unroll
    for (int o = 0; o < kMaxOut; ++o) acc[o] = 0.0f;
    // rolled on purpose: the kernel is instruction-fetch sensitive (every CTA runs each path once, cold)
#pragma unroll 1
    for (int c = lane; c < in4; c += 32) {
        {
            float4 wv[kMaxOut];                                // W column chunk: independent of the activations, issued first
#pragma unroll
            for (int o = 0; o < kMaxOut; ++o)
                if (o < out_f) wv[o] = __ldg(reinterpret_cast<const float4*>(W + (size_t)o * in_f) + c);
            float4 a;
            if (j.a_part) {
                // the producing GEMM left its split-K partials unfolded: sum them here (split order), add bias, ReLU
                a = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                const float* base = j.a_part + (size_t)r * in_f + 4 * c;
#pragma unroll 1
                for (int z0 = 0; z0 < j.a_splits; z0 += 8) {
                    float4 q[8];
#pragma unroll
                    for (int t = 0; t < 8; ++t)
                        q[t] = (z0 + t < j.a_splits) ? ldcg4(base + (size_t)(z0 + t) * j.a_stride) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
                    for (int t = 0; t < 8; ++t) { a.x += q[t].x; a.y += q[t].y; a.z += q[t].z; a.w += q[t].w; }
                }
                if (j.a_bias) {
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(j.a_bias) + c);
                    a.x += bb.x; a.y += bb.y; a.z += bb.z; a.w += bb.w;
                }
                if (j.a_relu) { a.x = fmaxf(a.x, 0.0f); a.y = fmaxf(a.y, 0.0f); a.z = fmaxf(a.z, 0.0f); a.w = fmaxf(a.w, 0.0f); }
                *reinterpret_cast<float4*>(j.a_out + (size_t)r * in_f + 4 * c) = a;
            } else {
                a = ldcg4(arow + 4 * c);
            }
            if (c == lane) a_first = a;
#pragma unroll
            for (int o = 0; o < kMaxOut; ++o) {
                if (o < out_f) {
                    const float4 w = wv[o];
                    acc[o] = fmaf(a.x, w.x, acc[o]);
                    acc[o] = fmaf(a.y, w.y, acc[o]);
                    acc[o] = fmaf(a.z, w.z, acc[o]);
                    acc[o] = fmaf(a.w, w.w, acc[o]);
                }
            }
        }
    }
    float mx = -INFINITY;
    int bi = 0;
#pragma unroll
    for (int o = 0; o < kMaxOut; ++o) {
        if (o < out_f) {
            float v = warp_sum(acc[o]);
            if (j.bias) v += hb[o];
            acc[o] = v;
            if (v > mx) { mx = v; bi = o; }                   // strict '>' from -inf, first max wins (src/tensor.rs:1062)
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int o = 0; o < kMaxOut; ++o)
        if (o < out_f) s += expf(acc[o] - mx);                // classes ascending (src/tensor.rs:890-1018 sum over dim 1)
    const float ls = logf(s);
    unsigned int cls = class_of(t);
    if (cls >= (unsigned int)out_f) { if (lane == 0) atomicExch(P.err, 1); cls = out_f - 1; }
    float dl[kMaxOut];
    float xc = 0.0f;
    const float scale = 1.0f * j.inv_b;                       // g(loss)[0] / B with the seeded g = 1 (src/loss.rs:186)
#pragma unroll
    for (int o = 0; o < kMaxOut; ++o) {
        dl[o] = 0.0f;
        if (o < out_f) {
            const float lp = (acc[o] - mx) - ls;
            float p = expf(lp);
            if ((unsigned int)o == cls) { p -= 1.0f; xc = acc[o]; }
            dl[o] = p * scale;
        }
    }
    if (lane == 0) {
        j.nll[r] = -((xc - mx) - ls);
        j.hit[r] = (fabsf((float)bi - t) < 1e-6f) ? 1.0f : 0.0f;       // src/loss.rs:284
    }
    if (lane < kMaxOut) {
        float mine = 0.0f;
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o) if (o == lane) mine = dl[o];
        j.dlog[(size_t)r * kMaxOut + lane] = mine;
    }
    if (j.dz) {
#pragma unroll 1
        for (int c = lane; c < in4; c += 32) {
            {
                float4 d = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
                for (int o = 0; o < kMaxOut; ++o) {
                    if (o < out_f) {
                        const float4 w = __ldg(reinterpret_cast<const float4*>(W + (size_t)o * in_f) + c);
                        d.x = fmaf(dl[o], w.x, d.x); d.y = fmaf(dl[o], w.y, d.y);
                        d.z = fmaf(dl[o], w.z, d.z); d.w = fmaf(dl[o], w.w, d.w);
                    }
                }
                if (j.dz_mask) {
                    const float4 a = (c == lane) ? a_first : ldcg4(afin + 4 * c);      // this lane wrote / read it above
                    d.x = a.x > 0.0f ? d.x : 0.0f; d.y = a.y > 0.0f ? d.y : 0.0f;
                    d.z = a.z > 0.0f ? d.z : 0.0f; d.w = a.w > 0.0f ? d.w : 0.0f;
                }
                *reinterpret_cast<float4*>(j.dz + (size_t)r * in_f + 4 * c) = d;
            }
        }
    }
}

// ---- loss = sum(nll) / B, correct = sum(hit)  (fixed tree: deterministic) ----------------------------------------
__device__ void loss_item(const Job& j, const StepParams& P, float* sm /* >= 16 floats */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float a = 0.0f, h = 0.0f;
    for (int r = threadIdx.x; r < j.M; r += kThreads) { a += __ldcg(j.nll + r); h += __ldcg(j.hit + r); }
    a = warp_sum(a); h = warp_sum(h);
    if (lane == 0) { sm[wid] = a; sm[8 + wid] = h; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float sa = 0.0f, sh = 0.0f;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) { sa += sm[w]; sh += sm[8 + w]; }
        const float loss = sa / (float)j.M;                   // acc / b as f32 (src/loss.rs:164)
        j.result[0] = loss;
        j.result[1] = sh;
        if (P.result_host) {
            P.result_host[0] = loss; P.result_host[1] = sh;
            P.result_host[3] = __int_as_float(__ldcg(P.err)); // 1: a label outside [0, classes) in this batch (the reference panics)
            if (P.result_seq) {                               // publish: the host spins on this word instead of a CUDA event
                __threadfence_system();
                *reinterpret_cast<volatile unsigned int*>(P.result_host + 2) = P.result_seq;
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void opt_item(const Job& j, int item, int phase_item, const StepParams& P, const AdamArgs& aa) {
    const int i = item * kThreads + threadIdx.x;
    const bool live = i < j.n4;
    float4 gg = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (live) {
        if (j.g_part) {                                       // fold the producing GEMM's split-K partials (split order)
            const float* base = j.g_part + 4 * (size_t)i;
#pragma unroll 1
            for (int z0 = 0; z0 < j.g_splits; z0 += 8) {
                float4 q[8];
#pragma unroll
                for (int t = 0; t < 8; ++t)
                    q[t] = (z0 + t < j.g_splits) ? ldcg4(base + (size_t)(z0 + t) * j.g_stride) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
                for (int t = 0; t < 8; ++t) { gg.x += q[t].x; gg.y += q[t].y; gg.z += q[t].z; gg.w += q[t].w; }
            }
        } else {
            gg = __ldcg(reinterpret_cast<const float4*>(j.g) + i);
        }
    }
    if (P.world > 1) {
        // all-reduce fused into the optimizer: push my slice to every peer, flag it, wait for theirs, sum in rank order
        const size_t e = (size_t)j.g_off + 4 * (size_t)i;
        if (live) {
#pragma unroll
            for (int r = 0; r < kMaxWorld; ++r)
                if (r < P.world && r != P.rank) *reinterpret_cast<float4*>(P.peer_slot[r] + e) = gg;
        }
        __syncthreads();
        const int tid = threadIdx.x;
        if (tid < P.world && tid != P.rank) {
            // release at system scope: cumulative over the whole CTA's slot stores (ordered before it by the barrier above)
            asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(P.peer_flags[tid] + phase_item), "r"(P.xseq) : "memory");
            const unsigned int* f = P.my_flags + (size_t)tid * P.x_items + phase_item;
            unsigned int cur;
            const long long t0 = clock64();
            do {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(cur) : "l"(f) : "memory");
                if ((int)(cur - P.xseq) < 0 && clock64() - t0 > kPeerSpinLimit) { atomicExch(P.err, 3); break; }
            } while ((int)(cur - P.xseq) < 0);
        }
        __syncthreads();
        if (__ldcg(P.err) >= 2) return;                      // a peer never arrived: its slot is stale, apply nothing
        if (live) {
            float4 q[kMaxWorld];
#pragma unroll
            for (int r = 0; r < kMaxWorld; ++r)
                q[r] = (r < P.world && r != P.rank) ? ldcg4(P.my_slots + (size_t)r * P.x_arena + e) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            float4 sum = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
            for (int r = 0; r < kMaxWorld; ++r) {             // rank order on every rank: replicas stay bit-identical
                if (r < P.world) {
                    const float4 v = (r == P.rank) ? gg : q[r];
                    sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
                }
            }
            gg = sum;
        }
    }
    if (!live) return;
    float4* p4 = reinterpret_cast<float4*>(j.p) + i;
    float4 pp = *p4;
    if (P.opt_kind == 0) {                                    // SGD: p -= lr * g (src/optim.rs:29)
        if (P.grad_scale != 1.0f) { gg.x *= P.grad_scale; gg.y *= P.grad_scale; gg.z *= P.grad_scale; gg.w *= P.grad_scale; }
        pp.x -= P.sgd_lr * gg.x; pp.y -= P.sgd_lr * gg.y; pp.z -= P.sgd_lr * gg.z; pp.w -= P.sgd_lr * gg.w;
        *p4 = pp;
        return;
    }
    float4* m4 = reinterpret_cast<float4*>(j.m) + i;
    float4* v4 = reinterpret_cast<float4*>(j.v) + i;
    float4 mm = *m4, vv = *v4;
    adam_elem(pp.x, gg.x, mm.x, vv.x, aa);
    adam_elem(pp.y, gg.y, mm.y, vv.y, aa);
    adam_elem(pp.z, gg.z, mm.z, vv.z, aa);
    adam_elem(pp.w, gg.w, mm.w, vv.w, aa);
    *p4 = pp; *m4 = mm; *v4 = vv;
}

__global__ void __launch_bounds__(kThreads, 1)
tape_step_kernel(const StepParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* As = reinterpret_cast<float*>(smem_raw);            // [2][SK][LDS]
    float* Bs = As + 2 * SK * LDS;                             // [2][SK][LDS]
    float* red = Bs + 2 * SK * LDS;                            // [16] reduction scratch + [16] optimizer state
    int* s_flag = reinterpret_cast<int*>(red + 32);            // [4]
    Job* sjobs = reinterpret_cast<Job*>(s_flag + 4);           // [n_jobs]
    int* ridx = reinterpret_cast<int*>(sjobs + P.n_jobs);      // [batch] when gathering
    const int tid = threadIdx.x;

#define TP_PROF(slot) do { if (P.prof && tid == 0) P.prof[(size_t)blockIdx.x * kProfSlots + (slot)] = clock64(); } while (0)
    // Programmatic dependent launch: the next step's CTAs may be scheduled as soon as SMs free up (they stage their job list
    // and gather index, which nothing in this step writes, and then wait for this grid to complete) — hides the ~3.5 us
    // between two back-to-back launches.  No-ops when the kernel was launched without the PDL attribute.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    TP_PROF(0);
    // setup: the three independent fetches (job list, gather index, optimizer state) are issued by different warps /
    // consumed late so that their latencies overlap
    const int* rix = P.perm ? ridx : nullptr;
    if (P.perm && P.cursor_value < 0) asm volatile("griddepcontrol.wait;" ::: "memory");      // the device cursor is written by the previous step
    if (tid < kThreads / 2 || !P.perm) {                       // stage the job list
        const int nthr = P.perm ? kThreads / 2 : kThreads;
        const int words = P.n_jobs * (int)(sizeof(Job) / 4);
        const int* src = reinterpret_cast<const int*>(P.jobs);
        int* dst = reinterpret_cast<int*>(sjobs);
#pragma unroll 4
        for (int i = tid; i < words; i += nthr) dst[i] = __ldg(src + i);
    } else {                                                   // MNISTDataset::get_batch as an index (src/data/mnist.rs:276-309)
        const int start = P.cursor_value >= 0 ? P.cursor_value : __ldcg(P.cursor);
#pragma unroll 4
        for (int r = tid - kThreads / 2; r < P.batch; r += kThreads / 2) {
            int idx = start + r;
            if (idx >= P.n_perm) idx %= P.n_perm;
            ridx[r] = __ldg(P.perm + idx);
        }
    }
    // everything below reads state the previous step (still draining under PDL) writes: parameters, moments, Adam state
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // optimizer state: fetched now (before CTA 0 advances it at the end), parked in shared memory, used in the last phase
    if (P.opt_kind != 0 && tid >= kThreads - H_COUNT) {
        const int q = tid - (kThreads - H_COUNT);
        red[16 + q] = __ldcg(P.hyper + q);
    }
    if (tid == 0) s_flag[3] = __ldcg(P.err);                  // sticky device error of an earlier step (same round trip)
    __syncthreads();
    TP_PROF(1);
    // A barrier or peer-exchange timeout (codes 2, 3) leaves gradients incomplete: nothing may be applied from them, in this
    // step or any later one.  The host sees the code in the result slot (fetch raises) or through tp_ctx_device_error.
    if (s_flag[3] >= 2) {
        if (blockIdx.x == 0 && tid == 0 && P.result_host) {
            P.result_host[0] = 0.0f; P.result_host[1] = 0.0f;
            P.result_host[3] = __int_as_float(s_flag[3]);
            if (P.result_seq) {
                __threadfence_system();
                *reinterpret_cast<volatile unsigned int*>(P.result_host + 2) = P.result_seq;
            }
        }
        return;
    }

    AdamArgs aa{};
    int t_new = 0;
    for (int ph = 0; ph < P.n_phases; ++ph) {
        if (ph == P.n_phases - 1) {
            if (P.opt_kind != 0) {                             // Adam::step prologue (src/optim.rs:86-90), as adam_advance_kernel
                const float h_t = red[16 + H_T], h_lr = red[16 + H_LR], h_b1 = red[16 + H_B1], h_b2 = red[16 + H_B2];
                const float h_eps = red[16 + H_EPS], h_wd = red[16 + H_WD];
                t_new = __float_as_int(h_t) + 1;
                const float bc1 = 1.0f - powi_dev(h_b1, t_new);
                const float bc2 = 1.0f - powi_dev(h_b2, t_new);
                aa.step_size = h_lr * (sqrtf(bc2) / bc1);
                aa.beta1 = h_b1; aa.beta2 = h_b2; aa.eps = h_eps;
                aa.decay_factor = 1.0f - h_lr * h_wd;
                const int decoupled = P.opt_kind == 2;
                aa.weight_decay = decoupled ? 0.0f : h_wd;
                aa.decoupled = (decoupled && h_wd > 0.0f) ? 1 : 0;
                aa.grad_scale = P.grad_scale;
            }
        }
        const int j0 = P.phase_first[ph], j1 = P.phase_first[ph + 1];
        int total = 0;
        for (int q = j0; q < j1; ++q) total += sjobs[q].items;
        for (int item = blockIdx.x; item < total; item += gridDim.x) {
            int q = j0, local = item;
            while (local >= sjobs[q].items) { local -= sjobs[q].items; ++q; }
            const Job& j = sjobs[q];
            switch (j.kind) {
                case JOB_GEMM:
                    gemm_item(j.a_kc != 0, j.b_kc != 0, j, local, P, rix, As, Bs, s_flag);
                    break;
                case JOB_HEAD: head_item(j, local, P, rix); break;
                case JOB_LOSS: loss_item(j, P, red); break;
                default: opt_item(j, local, item, P, aa); break;
            }
        }
        TP_PROF(2 + 2 * ph);
        if (ph + 1 < P.n_phases && !grid_sync(P.bar, P.bar_base + (unsigned int)(ph + 1) * gridDim.x, P.err, s_flag + 2))
            return;                                          // incomplete barrier: this CTA applies nothing from partial data
        TP_PROF(3 + 2 * ph);
    }
#undef TP_PROF
    if (blockIdx.x == 0 && tid == 0) {
        // every CTA read the cursor and the optimizer state before the first barrier
        if (P.perm) {
            const int start = P.cursor_value >= 0 ? P.cursor_value : *P.cursor;
            *P.cursor = (int)(((long long)start + P.batch) % P.n_perm);
        }
        if (P.opt_kind != 0) {
            P.hyper[H_T] = __int_as_float(t_new);
            P.hyper[H_SS] = aa.step_size;
            P.hyper[H_DECAY] = aa.decay_factor;
        }
    }
}

size_t smem_bytes(int n_jobs, int batch) {
    return (size_t)(4 * SK * LDS + 32 + 4) * sizeof(float) + (size_t)n_jobs * sizeof(Job) + (size_t)batch * sizeof(int) + 16;
}

}  // namespace

namespace tp {
struct WidePlan;
bool wide_supported(const tp_step_desc* d, const char** why);
bool wide_preferred(const tp_step_desc* d);
int wide_create(tp_ctx* ctx, const tp_step_desc* desc, float* P, float* G, float* M, float* V, float* hyper, float* result, WidePlan** out);
void wide_destroy(WidePlan* w);
int wide_refresh(WidePlan* w);
void wide_info(const WidePlan* w, int* n_phases, int* n_jobs, int* grid);
int wide_set_profile(WidePlan* w, int on);
int wide_read_profile(WidePlan* w, long long* out, size_t cap, int* slots);
struct WideXchg {
    int rank = 0, world = 1, items = 0;
    size_t arena_len = 0, row = 0, slots_off = 0;
    unsigned char* window = nullptr;
    unsigned char* peers[8] = {};
    unsigned int seq = 0;
};
int wide_run(WidePlan* w, const void* x, int x_is_u8, const float* labels, const int* perm, int* cursor, int n_perm, int cursor_value,
             float sgd_lr, float grad_scale, float* result_host, unsigned int result_seq, const WideXchg* xc);
}  // namespace tp

struct tp_xchg {
    tp_ctx* ctx = nullptr;
    int rank = 0, world = 1;
    size_t arena_len = 0;
    size_t row = 0;                      // floats per slot row: arena_len + 16 (the wide plan packs {recv, result} = 2 slices per row)
    int items = 0;                       // flags per source rank (>= optimizer-phase items of any step using this window)
    // window: [flags: world x items u32][slots: 2 parities x world source ranks x arena_len f32]
    unsigned char* window = nullptr;
    size_t bytes = 0, slots_off = 0;
    unsigned char* peers[kMaxWorld] = {};    // mapped base of every rank's window (own one included)
    bool connected = false;
    unsigned int seq = 0;                // steps run through this window (flag value of the next step is seq + 1)
};

struct tp_step {
    tp_ctx* ctx = nullptr;
    tp_xchg* xchg = nullptr;
    const Job* jobs_dev = nullptr;
    unsigned int bar_count = 0;                      // host mirror of the barrier's arrival counter
    tp_step_desc desc{};
    StepParams params{};
    std::vector<Job> jobs;
    void* dev_block = nullptr;       // one allocation: job list, barrier, tickets, scratch
    size_t dev_bytes = 0;
    int grid = 0;
    size_t smem = 0;
    tp_buf *p = nullptr, *g = nullptr, *m = nullptr, *v = nullptr, *hyper = nullptr, *result = nullptr;
    long long* prof = nullptr;
    tp::WidePlan* wide = nullptr;    // non-NULL: this step is the multi-kernel tcgen05 plan (step_wide.cu), not the persistent kernel
};

namespace {
// the window as the wide plan sees it; seq is the flag value of the run about to be launched
bool wide_xchg_of(const tp_step* s, tp::WideXchg* xc) {
    const tp_xchg* x = s->xchg;
    if (!x) return false;
    xc->rank = x->rank; xc->world = x->world; xc->items = x->items;
    xc->arena_len = x->arena_len; xc->row = x->row; xc->slots_off = x->slots_off;
    xc->window = x->window;
    for (int r = 0; r < x->world && r < 8; ++r) xc->peers[r] = x->peers[r];
    xc->seq = x->seq + 1;
    return true;
}
}  // namespace

namespace {

struct Carver {                      // bump allocator over one device block (two passes: size, then assign)
    unsigned char* base = nullptr;
    size_t off = 0;
    template <typename T> T* take(size_t count) {
        off = (off + 255) & ~(size_t)255;
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

double step_flops(const tp_step_desc* d) {
    double f = 0.0;
    for (int l = 0; l < d->n_layers; ++l) f += 2.0 * d->batch * (double)d->dims[l] * d->dims[l + 1];
    return 3.0 * f;
}

bool desc_ok(const tp_step_desc* d, const char** why) {
    auto no = [&](const char* w) { if (why) *why = w; return false; };
    if (!d) return no("NULL descriptor");
    if (d->n_layers < 1 || d->n_layers > TP_STEP_MAX_LAYERS) return no("layer count");
    if (d->batch < 1 || d->batch > kMaxBatch) return no("batch size");
    if (d->optimizer < 0 || d->optimizer > 2) return no("optimizer kind");
    const int L = d->n_layers;
    if (d->dims[L] < 1 || d->dims[L] > kMaxOut) return no("classifier wider than 16");
    if (d->dims[L - 1] > 1024) return no("classifier input wider than 1024");
    for (int l = 0; l < L; ++l) {
        if (d->dims[l] < 4 || d->dims[l] % 4) return no("feature width not a multiple of 4");
        if (d->dims[l] > 4096) return no("feature width > 4096");
        if (d->w_off[l] % 4 || (d->b_off[l] >= 0 && d->b_off[l] % 4)) return no("parameter slice not 16-byte aligned");
    }
    if (d->arena_len % 4) return no("arena length not a multiple of 4");
    if (step_flops(d) > 1.5e9) return no("step too large: the tcgen05 GEMM path is faster");
    return true;
}

// Build the job list.  Pass 1 (c.base == NULL) only sizes the device block.
void build(tp_step* s, Carver& c, int sms, float* G) {
    const tp_step_desc& d = s->desc;
    const int L = d.n_layers, B = d.batch;
    float* P = s->p->ptr;
    s->jobs.clear();
    StepParams& sp = s->params;
    int ph = 0;
    auto begin_phase = [&]() { sp.phase_first[ph++] = (int)s->jobs.size(); };
    auto pick_splits = [&](Job& j, int share) {
        const int tiles = j.tiles_m * j.tiles_n;
        int splits = share / tiles;
        if (splits < 1) splits = 1;
        int maxs = j.K / 32;
        if (maxs < 1) maxs = 1;
        if (splits > maxs) splits = maxs;
        int kchunk = ((j.K + splits - 1) / splits + BK - 1) / BK * BK;
        j.kchunk = kchunk;
        j.splits = (j.K + kchunk - 1) / kchunk;
        j.items = tiles * j.splits;
        j.tickets = c.take<int>(tiles);
        if (j.splits > 1) {
            j.partial = c.take<float>((size_t)j.splits * j.tiles_m * BM * j.N);
            if (j.colsum) j.cs_partial = c.take<float>((size_t)j.splits * j.tiles_m * BM);
        }
    };
    auto gemm = [&](int M, int N, int K) {
        Job j{};
        j.kind = JOB_GEMM;
        j.M = M; j.N = N; j.K = K; j.m_store = M;
        j.tiles_m = (M + BM - 1) / BM; j.tiles_n = (N + BN - 1) / BN;
        return j;
    };
    // activations act[l] = output of hidden layer l  [B, dims[l+1]];  dz[l] = gradient w.r.t. its pre-activation
    std::vector<float*> act(L, nullptr), dz(L, nullptr);
    for (int l = 0; l + 1 < L; ++l) {
        act[l] = c.take<float>((size_t)B * d.dims[l + 1]);
        dz[l] = c.take<float>((size_t)B * d.dims[l + 1]);
    }
    float* dlog = c.take<float>((size_t)B * kMaxOut);
    float* nll = c.take<float>(B);
    float* hit = c.take<float>(B);

    // ---- forward: hidden layers (Linear::forward src/nn.rs:54-60 [+ ReLU src/ops.rs:312-349]) -----------------------
    for (int l = 0; l + 1 < L; ++l) {
        begin_phase();
        Job j = gemm(B, d.dims[l + 1], d.dims[l]);
        j.a_kc = 1; j.b_kc = 1;
        if (l == 0) j.a_input = 1; else j.A = act[l - 1];
        j.lda = d.dims[l];
        j.B = P + d.w_off[l]; j.ldb = d.dims[l];
        j.C = act[l]; j.ldc = d.dims[l + 1];
        j.bias = d.b_off[l] >= 0 ? P + d.b_off[l] : nullptr;
        j.relu = d.relu[l];
        pick_splits(j, sms);
        if (l == L - 2 && j.splits > 1) j.defer_fold = 1;      // the head job folds the partials while it reads its rows
        s->jobs.push_back(j);
    }
    // ---- head ----------------------------------------------------------------------------------------------------------
    {
        begin_phase();
        Job j{};
        j.kind = JOB_HEAD;
        j.M = B; j.K = d.dims[L - 1]; j.N = d.dims[L];
        if (L == 1) j.a_input = 1; else j.A = act[L - 2];
        j.lda = d.dims[L - 1];
        if (L >= 2 && s->jobs.back().defer_fold) {
            const Job& f = s->jobs.back();
            j.a_part = f.partial; j.a_splits = f.splits; j.a_stride = (long long)f.tiles_m * BM * f.N;
            j.a_bias = f.bias; j.a_relu = f.relu; j.a_out = act[L - 2];
        }
        j.B = P + d.w_off[L - 1];
        j.bias = d.b_off[L - 1] >= 0 ? P + d.b_off[L - 1] : nullptr;
        j.dlog = dlog; j.nll = nll; j.hit = hit;
        j.inv_b = 1.0f / (float)B;
        if (L >= 2) { j.dz = dz[L - 2]; j.dz_mask = d.relu[L - 2]; }
        j.items = (B + kHeadRows - 1) / kHeadRows;
        s->jobs.push_back(j);
    }
    // ---- backward phases -------------------------------------------------------------------------------------------------
    // layer l (from the head down): dW_l = dZ_l^T . A_{l-1} with db_l = colsum(dZ_l) riding along (src/ops.rs:280-291,
    // src/tensor.rs:680-691) and, for hidden layers l > 0, dZ_{l-1} = (dZ_l . W_l) * [A_{l-1} > 0] (src/ops.rs:254-265,
    // 358-370).  The head's dX comes out of the head job, so the head's dW shares a phase with layer L-2's jobs.
    std::vector<Job> pending;
    auto flush_phase = [&]() {
        if (pending.empty()) return;
        begin_phase();
        int tiles_total = 0;
        for (auto& j : pending) if (j.kind == JOB_GEMM) tiles_total += j.tiles_m * j.tiles_n;
        for (auto& j : pending) {
            if (j.kind == JOB_GEMM) {
                int share = (int)((long long)sms * (j.tiles_m * j.tiles_n) / tiles_total);
                pick_splits(j, share);
                // parameter gradients: the optimizer job folds the partials itself unless something else (a gradient
                // exchange) needs them materialised in the arena
                if (!j.a_kc && j.splits > 1 && !d.materialize_grads) j.defer_fold = 1;
            }
            s->jobs.push_back(j);
        }
        pending.clear();
    };
    for (int l = L - 1; l >= 0; --l) {
        const bool head = (l == L - 1);
        const int out = d.dims[l + 1], in = d.dims[l];
        {
            Job j = gemm(head ? kMaxOut : out, in, B);
            j.m_store = out;
            j.a_kc = 0; j.b_kc = 0;
            j.A = head ? dlog : dz[l]; j.lda = head ? kMaxOut : out;
            if (l == 0) j.b_input = 1; else j.B = act[l - 1];
            j.ldb = in;
            j.C = G + d.w_off[l]; j.ldc = in;
            j.colsum = d.b_off[l] >= 0 ? G + d.b_off[l] : nullptr;
            pending.push_back(j);
        }
        if (l > 0 && !head) {
            Job j = gemm(B, in, out);
            j.a_kc = 1; j.b_kc = 0;
            j.A = dz[l]; j.lda = out;
            j.B = P + d.w_off[l]; j.ldb = in;
            j.C = dz[l - 1]; j.ldc = in;
            if (d.relu[l - 1]) { j.mask = act[l - 1]; j.ldmask = in; }
            pending.push_back(j);
        }
        if (head) {
            Job j{};
            j.kind = JOB_LOSS;
            j.items = 1;
            j.M = B; j.nll = nll; j.hit = hit; j.result = s->result->ptr;
            pending.push_back(j);
            if (L >= 2) continue;                              // layer L-2's jobs only need the head job's outputs too
        }
        flush_phase();
    }
    flush_phase();
    // ---- optimizer (src/optim.rs:21-33, 83-113, 148-168): one job per parameter tensor of the flat arena -----------------
    {
        std::vector<Job> dws;                                  // the dW jobs, to find each tensor's gradient source
        for (auto& q : s->jobs) if (q.kind == JOB_GEMM && !q.a_kc) dws.push_back(q);
        begin_phase();
        auto opt_job = [&](int64_t off, size_t n, const float* part, int splits, long long stride) {
            Job j{};
            j.kind = JOB_OPT;
            j.p = P + off; j.g = G + off;
            j.m = s->m ? s->m->ptr + off : nullptr;
            j.v = s->v ? s->v->ptr + off : nullptr;
            j.n4 = (int)((n + 3) / 4);
            j.items = (j.n4 + kThreads - 1) / kThreads;
            j.g_part = part; j.g_splits = splits; j.g_stride = stride;
            j.g_off = off;
            s->jobs.push_back(j);
        };
        for (int l = 0; l < L; ++l) {
            const Job* src = nullptr;
            for (auto& q : dws) if (q.C == G + d.w_off[l]) src = &q;
            const bool deferred = src && src->defer_fold;
            const long long mpad = src ? (long long)src->tiles_m * BM : 0;
            opt_job(d.w_off[l], (size_t)d.dims[l] * d.dims[l + 1], deferred ? src->partial : nullptr, deferred ? src->splits : 0,
                    deferred ? mpad * src->N : 0);
            if (d.b_off[l] >= 0)
                opt_job(d.b_off[l], (size_t)d.dims[l + 1], deferred ? src->cs_partial : nullptr, deferred ? src->splits : 0, mpad);
        }
    }
    sp.phase_first[ph] = (int)s->jobs.size();
    sp.n_phases = ph;
    sp.n_jobs = (int)s->jobs.size();
    sp.bar = c.take<unsigned int>(4);
    s->jobs_dev = c.take<Job>(s->jobs.size());
    sp.jobs = s->jobs_dev;
}

}  // namespace

extern "C" {

int tp_step_supported(const tp_step_desc* desc) {
    const char* why = nullptr;
    return (desc_ok(desc, &why) || tp::wide_supported(desc, &why)) ? 1 : 0;
}

int tp_step_kind(const tp_step_desc* desc) {
    const char* why = nullptr;
    if (tp::wide_preferred(desc)) return 2;
    if (desc_ok(desc, &why)) return 1;
    return tp::wide_supported(desc, &why) ? 2 : 0;
}

int tp_step_create(tp_ctx* ctx, const tp_step_desc* desc, tp_buf* params, tp_buf* grads, tp_buf* m, tp_buf* v, tp_buf* hyper,
                   tp_buf* result, tp_xchg* xchg, tp_step** out) {
    TP_CHECK_ARG(ctx && out, "tp_step_create: NULL argument");
    if (xchg) {
        TP_CHECK_ARG(xchg->ctx == ctx && xchg->connected, "tp_step_create: the exchange window is not connected");
        TP_CHECK_ARG(desc && (size_t)desc->arena_len == xchg->arena_len, "tp_step_create: the exchange window was sized for another arena");
    }
    const char* why = nullptr;
    const int kind = tp_step_kind(desc);
    if (kind == 0) desc_ok(desc, &why);
    TP_CHECK_ARG(kind != 0, "tp_step_create: unsupported step (%s)", why ? why : "?");
    TP_NEED(params, desc->arena_len, "params"); TP_NEED(grads, desc->arena_len, "grads"); TP_NEED(result, 2, "result");
    if (desc->optimizer != 0) {
        TP_NEED(m, desc->arena_len, "m"); TP_NEED(v, desc->arena_len, "v"); TP_NEED(hyper, H_COUNT, "hyper");
    }
    TP_CHECK_ARG(!(((uintptr_t)params->ptr | (uintptr_t)grads->ptr | (uintptr_t)(m ? m->ptr : nullptr) | (uintptr_t)(v ? v->ptr : nullptr)) & 15),
                 "tp_step_create: arenas must be 16-byte aligned");
    TP_CHECK_ARG(!ctx->capturing, "tp_step_create: not inside a graph capture");
    cudaSetDevice(ctx->device);
    const int L = desc->n_layers;
    for (int l = 0; l < L; ++l) {
        size_t wn = (size_t)desc->dims[l] * desc->dims[l + 1];
        TP_CHECK_ARG((int64_t)(desc->w_off[l] + wn) <= desc->arena_len, "tp_step_create: weight %d outside the arena", l);
        TP_CHECK_ARG(desc->b_off[l] < 0 || desc->b_off[l] + desc->dims[l + 1] <= desc->arena_len,
                     "tp_step_create: bias %d outside the arena", l);
    }
    tp_step* s = new tp_step();
    s->ctx = ctx;
    s->desc = *desc;
    s->p = params; s->g = grads; s->m = m; s->v = v; s->hyper = hyper; s->result = result;
    for (tp_buf* b : {params, grads, m, v, hyper, result}) if (b) tp_buf_retain(b);
    if (kind == 2) {
        // wide model: a plan of tcgen05 kernels (step_wide.cu); a data-parallel run exchanges gradients through the peer-memory
        // window if one is given (two-phase, fused with the optimizer), else sums the arena with the context's NCCL communicator
        int rc = tp::wide_create(ctx, desc, params->ptr, grads->ptr, m ? m->ptr : nullptr, v ? v->ptr : nullptr, hyper ? hyper->ptr : nullptr,
                                 result->ptr, &s->wide);
        if (rc != TP_OK) { tp_step_destroy(s); return rc; }
        s->xchg = xchg;                  // connected window: exchange + optimizer as one peer-memory kernel; NULL: NCCL
        *out = s;
        return TP_OK;
    }
    s->xchg = xchg;
    float* G0 = grads->ptr;
    auto fail = [&](int rc) { tp_step_destroy(s); return rc; };
    Carver sizing;
    build(s, sizing, ctx->sm_count, G0);
    if ((int)s->jobs.size() > kMaxJobs || s->params.n_phases > kMaxPhases) {
        tp::set_error("tp_step_create: model too deep for the fused step (%zu jobs)", s->jobs.size());
        return fail(TP_ERR_UNSUPPORTED);
    }
    s->dev_bytes = sizing.off + 256;
    if (cudaMalloc(&s->dev_block, s->dev_bytes) != cudaSuccess) {
        cudaGetLastError();
        tp::set_error("tp_step_create: cudaMalloc(%zu) failed", s->dev_bytes);
        return fail(TP_ERR_OOM);
    }
    if (cudaMemsetAsync(s->dev_block, 0, s->dev_bytes, ctx->stream) != cudaSuccess) return fail(TP_ERR_CUDA);
    Carver place;
    place.base = (unsigned char*)s->dev_block;
    build(s, place, ctx->sm_count, G0);
    if (cudaMemcpyAsync((void*)s->jobs_dev, s->jobs.data(), s->jobs.size() * sizeof(Job), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
        return fail(TP_ERR_CUDA);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return fail(TP_ERR_CUDA);      // s->jobs is pageable host memory
    s->params.batch = desc->batch;
    s->params.opt_kind = desc->optimizer;
    s->params.hyper = hyper ? hyper->ptr : nullptr;
    s->params.err = ctx->dev_error;
    s->smem = smem_bytes((int)s->jobs.size(), desc->batch);
    // The limit is a property of the kernel, not of this step: steps of several batch sizes coexist (a ragged last batch
    // compiles its own step), so it is raised to the device's opt-in maximum rather than set to this step's need — a smaller
    // later step must not lower it under an earlier one.
    int optin = 0;
    if (cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device) != cudaSuccess || (size_t)optin < s->smem ||
        cudaFuncSetAttribute(tape_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin) != cudaSuccess) {
        cudaGetLastError();
        tp::set_error("tp_step_create: %zu bytes of shared memory not available", s->smem);
        return fail(TP_ERR_CUDA);
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tape_step_kernel, kThreads, s->smem) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        tp::set_error("tp_step_create: the step kernel does not fit on an SM");
        return fail(TP_ERR_CUDA);
    }
    int max_items = 1;
    for (int ph = 0; ph < s->params.n_phases; ++ph) {
        int t = 0;
        for (int q = s->params.phase_first[ph]; q < s->params.phase_first[ph + 1]; ++q) t += s->jobs[q].items;
        if (t > max_items) max_items = t;
    }
    s->grid = ctx->sm_count < max_items ? ctx->sm_count : max_items;
    if (xchg) {
        int opt_items = 0;
        for (int q = s->params.phase_first[s->params.n_phases - 1]; q < s->params.phase_first[s->params.n_phases]; ++q) opt_items += s->jobs[q].items;
        if (opt_items > xchg->items) {
            tp::set_error("tp_step_create: %d optimizer slices but the exchange window has %d flags per rank", opt_items, xchg->items);
            return fail(TP_ERR_INVALID);
        }
    }
    *out = s;
    return TP_OK;
}

int tp_step_run(tp_ctx* ctx, tp_step* s, const tp_buf* x, const tp_buf* labels, const tp_buf* perm_i32, tp_buf* cursor_i32,
                int n_perm, int cursor_value, float sgd_lr, float grad_scale, float* result_host, unsigned int result_seq) {
    TP_CHECK_ARG(ctx && s && s->ctx == ctx, "tp_step_run: NULL or foreign step");
    TP_CHECK_ARG(!ctx->capturing, "tp_step_run: a cooperative launch cannot be captured into a CUDA graph");
    const int in = s->desc.dims[0];
    if (s->wide) {
        if (perm_i32) {
            TP_CHECK_ARG(n_perm > 0 && cursor_value < n_perm, "tp_step_run: empty dataset or cursor outside it");
            TP_NEED(x, (size_t)n_perm * in, "images"); TP_NEED(labels, n_perm, "labels"); TP_NEED(perm_i32, n_perm, "perm");
            TP_NEED(cursor_i32, 1, "cursor");
        } else {
            TP_NEED(x, (size_t)s->desc.batch * in, "x"); TP_NEED(labels, s->desc.batch, "labels");
        }
        TP_CHECK_ARG(!((uintptr_t)x->ptr & 15), "tp_step_run: input rows must be 16-byte aligned");
        tp::WideXchg xc;
        const bool have_x = wide_xchg_of(s, &xc);
        int rc = tp::wide_run(s->wide, x->ptr, 0, labels->ptr, perm_i32 ? (const int*)perm_i32->ptr : nullptr,
                              perm_i32 ? (int*)cursor_i32->ptr : nullptr, perm_i32 ? n_perm : 0, cursor_value, sgd_lr, grad_scale, result_host,
                              result_seq, have_x ? &xc : nullptr);
        if (rc == TP_OK && have_x) s->xchg->seq += 1;
        return rc;
    }
    StepParams p = s->params;
    if (perm_i32) {
        TP_CHECK_ARG(n_perm > 0, "tp_step_run: empty dataset");
        TP_NEED(x, (size_t)n_perm * in, "images"); TP_NEED(labels, n_perm, "labels"); TP_NEED(perm_i32, n_perm, "perm");
        TP_NEED(cursor_i32, 1, "cursor");
        p.perm = (const int*)perm_i32->ptr; p.cursor = (int*)cursor_i32->ptr; p.n_perm = n_perm;
    } else {
        TP_NEED(x, (size_t)s->desc.batch * in, "x"); TP_NEED(labels, s->desc.batch, "labels");
        p.perm = nullptr; p.cursor = nullptr; p.n_perm = 0;
    }
    TP_CHECK_ARG(!((uintptr_t)x->ptr & 15), "tp_step_run: input rows must be 16-byte aligned");
    p.x = x->ptr; p.labels = labels->ptr;
    p.sgd_lr = sgd_lr; p.grad_scale = grad_scale;
    p.cursor_value = perm_i32 ? cursor_value : -1;
    TP_CHECK_ARG(cursor_value < n_perm || !perm_i32, "tp_step_run: cursor_value %d outside the dataset", cursor_value);
    p.result_host = result_host;
    p.result_seq = result_host ? result_seq : 0u;
    p.world = 1; p.rank = 0;
    // host mirrors (barrier counter, exchange sequence) only advance once the launch has been accepted
    p.bar_base = s->bar_count;
    if (tp_xchg* x = s->xchg) {
        const unsigned int seq = x->seq + 1;
        const size_t par = seq & 1u;
        p.world = x->world; p.rank = x->rank; p.xseq = seq;
        p.x_items = x->items; p.x_arena = (long long)x->row;
        p.my_flags = reinterpret_cast<unsigned int*>(x->window);
        p.my_slots = reinterpret_cast<const float*>(x->window + x->slots_off) + par * x->world * x->row;
        for (int r = 0; r < x->world; ++r) {
            p.peer_flags[r] = reinterpret_cast<unsigned int*>(x->peers[r]) + (size_t)x->rank * x->items;
            p.peer_slot[r] = reinterpret_cast<float*>(x->peers[r] + x->slots_off) + (par * x->world + x->rank) * x->row;
        }
    }
    cudaSetDevice(ctx->device);
    void* args[] = {(void*)&p};
    // Default: an ordinary launch with the programmatic-stream-serialization attribute (PDL): the grid is at most one CTA per SM,
    // so once the previous step has drained every CTA is resident and the grid barrier is safe on a GPU this process has to
    // itself (a spin limit turns a violation into a sticky device error, never a hang).  TAPER_STEP_COOP=1 asks the driver
    // for a cooperative launch instead (co-residency guaranteed or the launch fails; no PDL overlap).
    static const bool coop = [] { const char* v = getenv("TAPER_STEP_COOP"); return v && v[0] == '1'; }();
    cudaError_t e;
    if (coop) {
        e = cudaLaunchCooperativeKernel((const void*)tape_step_kernel, dim3(s->grid), dim3(kThreads), args, s->smem, ctx->stream);
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(s->grid);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = s->smem;
        cfg.stream = ctx->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, tape_step_kernel, p);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        tp::set_error("tp_step_run: cooperative launch failed: %s", cudaGetErrorString(e));
        return TP_ERR_CUDA;
    }
    s->bar_count += (unsigned int)(s->params.n_phases - 1) * (unsigned int)s->grid;
    if (s->xchg) s->xchg->seq += 1;
    ctx->launches++;
    return TP_OK;
}

int tp_step_run_u8(tp_ctx* ctx, tp_step* s, const tp_buf* x_u8, const tp_buf* labels, const tp_buf* perm_i32, tp_buf* cursor_i32,
                   int n_perm, int cursor_value, float sgd_lr, float grad_scale, float* result_host, unsigned int result_seq) {
    TP_CHECK_ARG(ctx && s && s->ctx == ctx, "tp_step_run_u8: NULL or foreign step");
    if (!s->wide) {
        tp::set_error("tp_step_run_u8: u8 pixels are read by the wide step plan only");
        return TP_ERR_UNSUPPORTED;
    }
    const int in = s->desc.dims[0];
    const size_t rows = perm_i32 ? (size_t)n_perm : (size_t)s->desc.batch;
    TP_CHECK_ARG(!perm_i32 || (n_perm > 0 && cursor_value < n_perm), "tp_step_run_u8: empty dataset or cursor outside it");
    TP_NEED(x_u8, (rows * in + 3) / 4, "x_u8"); TP_NEED(labels, rows, "labels");
    if (perm_i32) { TP_NEED(perm_i32, n_perm, "perm"); TP_NEED(cursor_i32, 1, "cursor"); }
    TP_CHECK_ARG(!((uintptr_t)x_u8->ptr & 15), "tp_step_run_u8: input rows must be 16-byte aligned");
    tp::WideXchg xc;
    const bool have_x = wide_xchg_of(s, &xc);
    int rc = tp::wide_run(s->wide, x_u8->ptr, 1, labels->ptr, perm_i32 ? (const int*)perm_i32->ptr : nullptr,
                          perm_i32 ? (int*)cursor_i32->ptr : nullptr, perm_i32 ? n_perm : 0, cursor_value, sgd_lr, grad_scale, result_host,
                          result_seq, have_x ? &xc : nullptr);
    if (rc == TP_OK && have_x) s->xchg->seq += 1;
    return rc;
}

int tp_step_is_wide(const tp_step* s) { return s && s->wide ? 1 : 0; }

int tp_step_refresh(tp_step* s) {
    TP_CHECK_ARG(s, "tp_step_refresh: NULL step");
    return s->wide ? tp::wide_refresh(s->wide) : TP_OK;
}

int tp_xchg_create(tp_ctx* ctx, size_t arena_len, int rank, int world, tp_xchg** out) {
    TP_CHECK_ARG(ctx && out && arena_len > 0 && arena_len % 4 == 0, "tp_xchg_create: bad arguments");
    TP_CHECK_ARG(world >= 2 && world <= kMaxWorld && rank >= 0 && rank < world, "tp_xchg_create: world %d / rank %d (2..%d ranks)", world, rank, kMaxWorld);
    cudaSetDevice(ctx->device);
    tp_xchg* x = new tp_xchg();
    x->ctx = ctx; x->rank = rank; x->world = world; x->arena_len = arena_len;
    x->row = arena_len + 16;
    x->items = (int)(arena_len / 4 / kThreads) + 4 * TP_STEP_MAX_LAYERS + 8;      // one flag per 256-float4 optimizer slice
    const size_t flags = ((size_t)world * x->items * sizeof(unsigned int) + 255) & ~(size_t)255;
    x->slots_off = flags;
    x->bytes = flags + 2 * (size_t)world * x->row * sizeof(float);
    if (cudaMalloc((void**)&x->window, x->bytes) != cudaSuccess) {       // plain cudaMalloc: cudaIpcGetMemHandle needs an allocation base
        cudaGetLastError();
        tp::set_error("tp_xchg_create: cudaMalloc(%zu) failed", x->bytes);
        delete x;
        return TP_ERR_OOM;
    }
    cudaMemset(x->window, 0, x->bytes);
    x->peers[rank] = x->window;
    *out = x;
    return TP_OK;
}

int tp_xchg_handle(tp_xchg* x, void* out64) {
    TP_CHECK_ARG(x && out64, "tp_xchg_handle: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaSetDevice(x->ctx->device);
    cudaIpcMemHandle_t h;
    TP_CUDA(cudaIpcGetMemHandle(&h, x->window));
    std::memcpy(out64, &h, 64);
    return TP_OK;
}

int tp_xchg_connect(tp_xchg* x, const void* handles, int world) {
    TP_CHECK_ARG(x && handles && world == x->world, "tp_xchg_connect: bad arguments");
    cudaSetDevice(x->ctx->device);
    for (int r = 0; r < world; ++r) {
        if (r == x->rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, (const unsigned char*)handles + 64 * (size_t)r, 64);
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            tp::set_error("tp_xchg_connect: cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
            return TP_ERR_COMM;
        }
        x->peers[r] = (unsigned char*)ptr;
    }
    x->connected = true;
    return TP_OK;
}

int tp_xchg_destroy(tp_xchg* x) {
    if (!x) return TP_OK;
    cudaSetDevice(x->ctx->device);
    cudaStreamSynchronize(x->ctx->stream);
    for (int r = 0; r < x->world; ++r)
        if (r != x->rank && x->peers[r]) cudaIpcCloseMemHandle(x->peers[r]);
    if (x->window) cudaFree(x->window);
    delete x;
    return TP_OK;
}

int tp_step_set_profile(tp_step* s, int on) {
    TP_CHECK_ARG(s, "tp_step_set_profile: NULL step");
    if (s->wide) return tp::wide_set_profile(s->wide, on);
    cudaSetDevice(s->ctx->device);
    if (on && !s->prof) {
        size_t bytes = (size_t)s->grid * kProfSlots * sizeof(long long);
        TP_CUDA(cudaMalloc(&s->prof, bytes));
        TP_CUDA(cudaMemsetAsync(s->prof, 0, bytes, s->ctx->stream));
    }
    s->params.prof = on ? s->prof : nullptr;
    return TP_OK;
}

int tp_step_read_profile(tp_step* s, int64_t* out, size_t cap, int* slots) {
    if (s && s->wide) return tp::wide_read_profile(s->wide, (long long*)out, cap, slots);
    TP_CHECK_ARG(s && s->prof && out && cap >= (size_t)s->grid * kProfSlots, "tp_step_read_profile: profiling is off or the buffer is too small");
    cudaSetDevice(s->ctx->device);
    TP_CUDA(cudaStreamSynchronize(s->ctx->stream));
    TP_CUDA(cudaMemcpy(out, s->prof, (size_t)s->grid * kProfSlots * sizeof(long long), cudaMemcpyDeviceToHost));
    if (slots) *slots = kProfSlots;
    return TP_OK;
}

int tp_step_info(const tp_step* s, int* n_phases, int* n_jobs, int* grid) {
    TP_CHECK_ARG(s, "tp_step_info: NULL step");
    if (s->wide) { tp::wide_info(s->wide, n_phases, n_jobs, grid); return TP_OK; }
    if (n_phases) *n_phases = s->params.n_phases;
    if (n_jobs) *n_jobs = s->params.n_jobs;
    if (grid) *grid = s->grid;
    return TP_OK;
}

int tp_step_destroy(tp_step* s) {
    if (!s) return TP_OK;
    if (s->ctx) {
        cudaSetDevice(s->ctx->device);
        cudaStreamSynchronize(s->ctx->stream);
    }
    if (s->wide) tp::wide_destroy(s->wide);
    if (s->dev_block) cudaFree(s->dev_block);
    if (s->prof) cudaFree(s->prof);
    for (tp_buf* b : {s->p, s->g, s->m, s->v, s->hyper, s->result}) if (b) tp_buf_release(b);
    delete s;
    return TP_OK;
}

}  // extern "C"

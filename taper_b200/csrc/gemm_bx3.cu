// bf16x3 tensor-core GEMM for sm_100a: fp32 operands are held as PRE-SPLIT bf16 pairs x = hi + lo
// (hi = rn_bf16(x), lo = rn_bf16(x - hi): 16 significant bits) and the product is formed as
//     D = A_lo*B_hi + A_hi*B_lo + A_hi*B_hi          (tcgen05.mma kind::f16, fp32 accumulation in TMEM)
// i.e. three bf16 MMAs per product: |error| <= ~3 * 2^-16 per product (typically ~1e-5 of |C|inf), inside the path's
// 1e-4 fp32 parity bar, at HALF the tensor-pipe cost of 3xTF32 (bf16 runs at twice the TF32 rate) and with no
// operand-splitting pass inside the kernel: the hi/lo planes are written by whoever produced the tensor (a GEMM
// epilogue, the optimizer step, the batch gather — or tp_split_bf16 for a foreign fp32 buffer), so the main loop is
// a plain TMA -> mbarrier -> tcgen05.mma pipeline.  Same fp32 HBM footprint (2 x 2 bytes per element).
//
// Replaces matrixmultiply::sgemm / cblas_sgemm behind sgemm_rowmajor (src/gemm.rs:8-49, 72-119) for the reference's
// call shapes (N,N) / (N,T) / (T,N) (src/ops.rs:215-226, 254-265, 280-291) and the fused Linear (src/nn.rs:54-60).
// A transposed operand is described to the tensor core as MN-major (no transpose in memory).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-5 = accumulator
// read-out (and, DRAIN: a round-to-nearest fp32 drain of the double-buffered accumulator every 256 elements of K,
// which bounds the tensor core's truncating accumulation).  K is split over a thread-block cluster (grid z) with a
// deterministic DSMEM fold; the epilogue fuses alpha/beta, bias, ReLU, a ReLU mask, the bf16 hi/lo planes of the
// OUTPUT (operand of the next GEMM) and partial column sums (bias gradients), all stored with 128-bit accesses.
#include "common.cuh"
#include "tc_ptx.cuh"
#include "wide_fold.cuh"

#include <cuda_bf16.h>
#include <cstdlib>
#include <type_traits>

namespace {

using namespace tcptx;

constexpr int BM = 128;              // UMMA M (cta_group::1)
constexpr int BK = 64;               // bf16 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;           // bf16: 32 bytes per instruction
// threads per CTA: the main loop needs 6 warps (TMA, MMA, 4 x accumulator read-out); the fold / epilogue / store phase is
// instruction-latency bound (~60 instructions per output vector), so it gets 16 warps — except in the DRAIN variants, whose
// read-out warps hold BN accumulators in registers
template <bool DRAIN> constexpr int threads_of() { return DRAIN ? 256 : 512; }

struct Bx3Params {
    int m, n, k;
    int splits;
    float alpha, beta;
    float* c;                        // fp32 output [m,n] (may be NULL when only the split planes are wanted)
    uint16_t* c_split;               // bf16 hi plane [m,n]; lo plane at c_split + c_plane (NULL: none)
    long long c_plane;
    const float* bias;
    const float* relu_mask;
    int relu;
    float* colsum_part;              // [tiles_m * splits][n] partial column sums of the stored values (NULL: none)
    unsigned long long* stamp;       // optional: CTA 0 stores %globaltimer here once its dependency wait is over
    int tiles_n;                     // CTAs of the grid's x extent beyond this problem's tiles leave at once
    int y0;                          // first grid row (blockIdx.y) of this problem in a grouped launch
    int a_early, b_early;            // operand was complete before the previous kernel started: its first stages are
                                     // requested BEFORE the programmatic dependency wait
};

// Up to three problems of the same tile shape / operand majors / K-split in ONE launch (grid rows are concatenated): the
// independent weight-gradient GEMMs of a backward pass share a launch instead of paying one kernel slot each.
constexpr int kMaxGroup = 3;
struct Bx3Group {
    Bx3Params p[kMaxGroup];
    int count;
    // grid rows >= tail_y0 are not GEMM tiles: their first CTA runs the plan's fold / bookkeeping (wide_fold.cuh) on SMs
    // the GEMM tiles leave free, the others leave at once
    int tail_y0, tail_workers;
    tpfold::FoldStep tail;
    tp::Bx3Push push;                // data parallel: outputs in another rank's slice of the arena also go to that rank's window
};

// K-major operand : rows of 128 B (64 bf16 of K), 8 rows = one 1024 B swizzle atom -> SBO 1024; LBO unused.
// MN-major operand: rows of 128 B (64 bf16 of M/N) indexed by k, 8 k-rows = one atom -> SBO 1024 (next 8 k);
//                   LBO = distance between 64-element M/N chunks = 16384 (each chunk is one TMA box of 2 planes x 8 KB).
// Both use layout type 2 (SWIZZLE_128B, 16-byte chunks XOR row % 8 = TMA SWIZZLE_128B).
template <bool MN_MAJOR>
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    constexpr uint32_t lbo_bytes = MN_MAJOR ? 16384 : 16, sbo_bytes = 1024;
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

template <int BN>
struct Smem {
    static constexpr int kABytes = 2 * BM * 128;      // hi + lo planes
    static constexpr int kBBytes = 2 * BN * 128;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = (200 * 1024 / kStageBytes) > 6 ? 6 : (200 * 1024 / kStageBytes);
    static constexpr int kTotal = kStages * kStageBytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void split_store4(uint16_t* hi, uint16_t* lo, const float (&v)[4]) {
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        h[j] = __float2bfloat16_rn(v[j]);
        l[j] = __float2bfloat16_rn(v[j] - __bfloat162float(h[j]));
    }
    uint2 ph, pl;
    ph.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
    ph.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
    pl.x = (uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16);
    pl.y = (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16);
    *(uint2*)hi = ph;
    *(uint2*)lo = pl;
}

// development aid: per-phase SM-clock stamps of CTA (0,0,0), read back with tpdbg_bx3_times()
__device__ long long g_bx3_t[16];
#define BX3_T(slot) do { if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_bx3_t[slot] = clock64(); } while (0)

template <int BN, bool A_MN, bool B_MN, bool DRAIN>
__global__ void __launch_bounds__(threads_of<DRAIN>(), 1)
gemm_bx3_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_b0,
                const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_b1,
                const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_b2,
                const __grid_constant__ Bx3Group grp) {
    static_assert(!DRAIN || BN <= 128, "the register drain holds BN accumulators per thread");
    constexpr int kThreads = threads_of<DRAIN>();
    if (grp.tail_workers > 0 && (int)blockIdx.y >= grp.tail_y0) {
        if (blockIdx.x != 0 || blockIdx.z != 0) return;
        extern __shared__ __align__(1024) uint8_t tail_smem[];
        pdl_launch_dependents();
        pdl_wait();
        tpfold::fold_worker<kThreads>(grp.tail, (int)blockIdx.y - grp.tail_y0, grp.tail_workers, reinterpret_cast<float*>(tail_smem));
        return;
    }
    int gi = 0;
    if (grp.count > 1 && (int)blockIdx.y >= grp.p[1].y0) gi = 1;
    if (grp.count > 2 && (int)blockIdx.y >= grp.p[2].y0) gi = 2;
    const Bx3Params& p = grp.p[gi];
    if ((int)blockIdx.x >= p.tiles_n) return;         // (every CTA of a K-split cluster shares x and y: they leave together)
    const CUtensorMap* pma = gi == 0 ? &map_a0 : gi == 1 ? &map_a1 : &map_a2;
    const CUtensorMap* pmb = gi == 0 ? &map_b0 : gi == 1 ? &map_b1 : &map_b2;
    const int by = (int)blockIdx.y - p.y0;
    using S = Smem<BN>;
    constexpr int kStages = S::kStages;
    constexpr int kChunk = 4;                         // k-blocks (256 elements of K) per tensor-core accumulation (DRAIN)
    constexpr uint32_t kTmemCols = DRAIN ? 2 * BN : BN;
    constexpr int kPitch = BN + 4;                    // floats per staged accumulator row
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full_bar = (uint64_t*)(smem + kStages * S::kStageBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* acc_full = empty_bar + kStages;         // [2]
    uint64_t* acc_empty = acc_full + 2;               // [2] (DRAIN)
    uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = by * BM, n0 = blockIdx.x * BN;
    const int total_kb = (p.k + BK - 1) / BK;
    const int kb0 = (int)(((long long)blockIdx.z * total_kb) / p.splits);
    const int kb1 = (int)(((long long)(blockIdx.z + 1) * total_kb) / p.splits);
    const int nkb = kb1 - kb0;                        // >= 1 because splits <= total_kb
    if (threadIdx.x == 0) BX3_T(0);

    if (threadIdx.x == 0) {
        tma_prefetch_desc(pma);
        tma_prefetch_desc(pmb);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar + s, 1);
            mbar_init(empty_bar + s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(acc_full + b, 1);
            mbar_init(acc_empty + b, 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) BX3_T(1);
    auto a_st = [&](int s) { return smem + s * S::kStageBytes; };
    auto b_st = [&](int s) { return smem + s * S::kStageBytes + S::kABytes; };
    auto load_a = [&](int s, int k0) {
        if (A_MN) {
#pragma unroll
            for (int g = 0; g < BM / 64; ++g) tma_load_3d(a_st(s) + g * 16384, pma, full_bar + s, m0 + 64 * g, k0, 0);
        } else {
            tma_load_3d(a_st(s), pma, full_bar + s, k0, m0, 0);
        }
    };
    auto load_b = [&](int s, int k0) {
        if (B_MN) {
#pragma unroll
            for (int g = 0; g < BN / 64; ++g) tma_load_3d(b_st(s) + g * 16384, pmb, full_bar + s, n0 + 64 * g, k0, 0);
        } else {
            tma_load_3d(b_st(s), pmb, full_bar + s, k0, n0, 0);
        }
    };
    // programmatic dependent launch: everything above overlapped the previous kernel's tail; nothing below may touch memory
    // the previous kernel writes before the wait.  An operand that was complete before the previous kernel even started
    // (weights: written by the last step's optimizer; the forward activations in a backward GEMM) is requested for the
    // first pipeline stages right away, so its tiles are in flight while the previous kernel drains.
    pdl_launch_dependents();
    const int n_early = nkb < kStages ? nkb : kStages;
    if (warp == 0 && lane == 0 && (p.a_early || p.b_early)) {
        for (int i = 0; i < n_early; ++i) {
            mbar_expect_tx(full_bar + i, S::kStageBytes);
            const int k0 = (kb0 + i) * BK;
            if (p.a_early) load_a(i, k0);
            if (p.b_early) load_b(i, k0);
        }
    }
    pdl_wait();
    if (threadIdx.x == 0) BX3_T(2);
    if (p.stamp && threadIdx.x == 0 && blockIdx.x == 0 && by == 0 && blockIdx.z == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        *p.stamp = gt;
    }

    float* stage = (float*)smem;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            const bool early = p.a_early || p.b_early;
            for (int i = 0; i < nkb; ++i) {
                const int s = i % kStages;
                const uint32_t ph = (i / kStages) & 1;
                const int k0 = (kb0 + i) * BK;
                if (early && i < n_early) {            // the stage's transaction count is armed; only the late operand is missing
                    if (!p.a_early) load_a(s, k0);
                    if (!p.b_early) load_b(s, k0);
                    continue;
                }
                mbar_wait(empty_bar + s, ph ^ 1);
                mbar_expect_tx(full_bar + s, S::kStageBytes);
                load_a(s, k0);
                load_b(s, k0);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one elected lane) =====
        if (lane == 0) {
            // instruction descriptor: c = F32 (1 << 4), a/b = BF16 (1 << 7, 1 << 10), majors (15, 16), N >> 3 (17), M >> 4 (24)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                                   ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            constexpr uint32_t kStepA = A_MN ? 2048 : UMMA_K * 2;      // bytes per UMMA_K step inside a stage
            constexpr uint32_t kStepB = B_MN ? 2048 : UMMA_K * 2;
            constexpr uint32_t kLoA = A_MN ? 8192 : BM * 128;          // offset of the lo plane
            constexpr uint32_t kLoB = B_MN ? 8192 : BN * 128;
            uint32_t tmem_d = tmem_base;
            for (int i = 0; i < nkb; ++i) {
                const int s = i % kStages;
                const uint32_t ph = (i / kStages) & 1;
                uint32_t fresh = (i == 0) ? 1u : 0u;
                if (DRAIN && i % kChunk == 0) {
                    const int chunk = i / kChunk, b = chunk & 1;
                    mbar_wait(acc_empty + b, ((chunk >> 1) & 1) ^ 1);
                    tc_fence_after();
                    tmem_d = tmem_base + b * BN;
                    fresh = 1u;
                }
                mbar_wait(full_bar + s, ph);
                tc_fence_after();
                if (i == 0) BX3_T(3);
                if (i == nkb - 1) BX3_T(4);
                const uint32_t ah = smem_u32(a_st(s)), bh = smem_u32(b_st(s));
                const uint32_t al = ah + kLoA, bl = bh + kLoB;
#pragma unroll
                for (int kk = 0; kk < BK / UMMA_K; ++kk)              // small terms first
                    tc_mma_bf16(tmem_d, make_desc<A_MN>(al + kk * kStepA), make_desc<B_MN>(bh + kk * kStepB), idesc,
                                (fresh && kk == 0) ? 0u : 1u);
#pragma unroll
                for (int kk = 0; kk < BK / UMMA_K; ++kk)
                    tc_mma_bf16(tmem_d, make_desc<A_MN>(ah + kk * kStepA), make_desc<B_MN>(bl + kk * kStepB), idesc, 1u);
#pragma unroll
                for (int kk = 0; kk < BK / UMMA_K; ++kk)
                    tc_mma_bf16(tmem_d, make_desc<A_MN>(ah + kk * kStepA), make_desc<B_MN>(bh + kk * kStepB), idesc, 1u);
                tc_commit(empty_bar + s);
                if (DRAIN ? (i % kChunk == kChunk - 1 || i == nkb - 1) : (i == nkb - 1))
                    tc_commit(acc_full + (DRAIN ? ((i / kChunk) & 1) : 0));
            }
        }
    } else if (warp < 6) {
        // ===== warps 2-5: accumulator -> this CTA's shared memory (row pitch BN + 4 floats) =====
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        if constexpr (DRAIN) {
            float acc[BN];
#pragma unroll
            for (int j = 0; j < BN; ++j) acc[j] = 0.0f;
            const int nchunks = (nkb + kChunk - 1) / kChunk;
            for (int c = 0; c < nchunks; ++c) {
                const int b = c & 1;
                mbar_wait(acc_full + b, (c >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int c0 = 0; c0 < BN; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(taddr + b * BN + c0, v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[c0 + j] += __uint_as_float(v[j]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty + b);
            }
#pragma unroll
            for (int c0 = 0; c0 < BN; c0 += 4)
                *(float4*)(stage + row * kPitch + c0) = make_float4(acc[c0], acc[c0 + 1], acc[c0 + 2], acc[c0 + 3]);
        } else {
            mbar_wait(acc_full, 0);
            tc_fence_after();
            if (threadIdx.x == 64) BX3_T(5);
#pragma unroll 4
            for (int c0 = 0; c0 < BN; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(taddr + c0, v);
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *(float4*)(stage + row * kPitch + c0 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                            __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            }
            tc_fence_before();
        }
    }
    __syncwarp();
    if (threadIdx.x == 64) BX3_T(6);
    if (p.splits > 1) cluster_sync_all(); else __syncthreads();
    if (threadIdx.x == 64) BX3_T(7);

    // CTA z of the cluster owns rows [z*BM/S, (z+1)*BM/S) of the tile: fold the S partial tiles in split order through
    // distributed shared memory, apply the epilogue, store coalesced.  A thread keeps one group of four columns (so its
    // bias vector is loaded once) and walks down the rows in batches whose loads — partial tiles, ReLU mask, old C — are
    // all issued before the first use.
    {
        const int t = threadIdx.x;
        const int S_ = p.splits;
        const int rows_per = BM / S_;
        const int r_begin = (int)blockIdx.z * rows_per;
        constexpr int kVecPerRow = BN / 4;
        constexpr int kRowStep = kThreads / kVecPerRow;                  // 12 / 6 / 3 rows between a thread's vectors
        static_assert(kThreads % kVecPerRow == 0, "a thread must stay on one column group");
        const uint32_t stage_addr = smem_u32(smem);
        const bool want_cs = p.colsum_part != nullptr;
        const int c4 = (t % kVecPerRow) * 4, col = n0 + c4;
        const bool col_ok = col < p.n;                                   // n % 4 == 0 (checked on the host): a live vector is whole
        const int r_first = r_begin + t / kVecPerRow;
        const int r_end = r_begin + rows_per;
        float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias && col_ok) bs = __ldg((const float4*)(p.bias + col));
        const bool has_beta = p.beta != 0.0f;
        auto fold_store = [&](auto ku_tag) {
            constexpr int KU = decltype(ku_tag)::value;
            for (int rb = r_first; rb < r_end; rb += KU * kRowStep) {
                float4 acc[KU], mk[KU], cold[KU];
                uint32_t off[KU];
                bool live[KU];
#pragma unroll
                for (int u = 0; u < KU; ++u) {
                    const int r = rb + u * kRowStep;
                    off[u] = (uint32_t)(r * kPitch + c4) * 4;
                    live[u] = r < r_end && (m0 + r) < p.m && col_ok;
                    const size_t g = (size_t)(m0 + r) * p.n + col;
                    mk[u] = (p.relu_mask && live[u]) ? __ldg((const float4*)(p.relu_mask + g)) : make_float4(1.f, 1.f, 1.f, 1.f);
                    cold[u] = (has_beta && live[u]) ? *(const float4*)(p.c + g) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (S_ == 1) {
#pragma unroll
                    for (int u = 0; u < KU; ++u)
                        acc[u] = rb + u * kRowStep < r_end ? *(const float4*)(smem + off[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
                } else {
                    // partial tiles in split order 0, 1, 2, ...: the result does not depend on timing; two peers in flight
                    for (int z0 = 0; z0 < S_; z0 += 2) {
                        float4 x[2][KU];
#pragma unroll
                        for (int dz = 0; dz < 2; ++dz)
#pragma unroll
                            for (int u = 0; u < KU; ++u)
                                if (z0 + dz < S_ && rb + u * kRowStep < r_end)
                                    x[dz][u] = (z0 + dz == (int)blockIdx.z) ? *(const float4*)(smem + off[u])
                                                                            : ld_cluster_f4(map_to_cta(stage_addr + off[u], (uint32_t)(z0 + dz)));
#pragma unroll
                        for (int dz = 0; dz < 2; ++dz)
#pragma unroll
                            for (int u = 0; u < KU; ++u)
                                if (z0 + dz < S_ && rb + u * kRowStep < r_end) {
                                    if (z0 + dz == 0) acc[u] = x[dz][u];
                                    else { acc[u].x += x[dz][u].x; acc[u].y += x[dz][u].y; acc[u].z += x[dz][u].z; acc[u].w += x[dz][u].w; }
                                }
                    }
                }
#pragma unroll
                for (int u = 0; u < KU; ++u) {
                    const int r = rb + u * kRowStep;
                    if (r >= r_end) continue;
                    float o[4] = {acc[u].x, acc[u].y, acc[u].z, acc[u].w};
                    if (live[u]) {
                        const size_t g = (size_t)(m0 + r) * p.n + col;
                        const float cv[4] = {cold[u].x, cold[u].y, cold[u].z, cold[u].w}, bv[4] = {bs.x, bs.y, bs.z, bs.w};
                        const float mv[4] = {mk[u].x, mk[u].y, mk[u].z, mk[u].w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float v = p.alpha * o[j];
                            if (has_beta) v += p.beta * cv[j];
                            v += bv[j];
                            if (p.relu) v = fmaxf(v, 0.0f);
                            if (p.relu_mask) v = mv[j] > 0.0f ? v : 0.0f;
                            o[j] = v;
                        }
                        if (p.c) *(float4*)(p.c + g) = make_float4(o[0], o[1], o[2], o[3]);
                        if (grp.push.world > 1 && p.c) {
                            // reduce-scatter push: the owner of this vector's slice sums the ranks' contributions (step_wide.cu)
                            const unsigned int e = (unsigned int)((p.c + g) - grp.push.base);
                            const unsigned int q = e / grp.push.slice;
                            if ((int)q != grp.push.rank && q < (unsigned int)grp.push.world)
                                *(float4*)(grp.push.peer[q] + (e - q * grp.push.slice)) = make_float4(o[0], o[1], o[2], o[3]);
                        }
                        if (p.c_split) split_store4(p.c_split + g, p.c_split + p.c_plane + g, o);
                    } else {
                        o[0] = o[1] = o[2] = o[3] = 0.0f;
                    }
                    if (want_cs) *(float4*)(smem + off[u]) = make_float4(o[0], o[1], o[2], o[3]);      // own rows only: no peer reads them
                }
            }
        };
        if (S_ <= 2) fold_store(std::integral_constant<int, 4>{});
        else fold_store(std::integral_constant<int, 2>{});
        if (want_cs) {
            __syncthreads();
            float* dst = p.colsum_part + ((size_t)by * S_ + blockIdx.z) * p.n;
            for (int c = t; c < BN; c += kThreads) {
                if (n0 + c >= p.n) continue;
                float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
                int r = 0;
                for (; r + 4 <= rows_per; r += 4) {
                    s0 += stage[(r_begin + r) * kPitch + c]; s1 += stage[(r_begin + r + 1) * kPitch + c];
                    s2 += stage[(r_begin + r + 2) * kPitch + c]; s3 += stage[(r_begin + r + 3) * kPitch + c];
                }
                for (; r < rows_per; ++r) s0 += stage[(r_begin + r) * kPitch + c];
                dst[n0 + c] = (s0 + s1) + (s2 + s3);
            }
        }
    }
    __syncwarp();
    if (threadIdx.x == 64) BX3_T(8);
    if (p.splits > 1) cluster_sync_all();
    __syncthreads();
    if (threadIdx.x == 0) BX3_T(9);
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

// ---- fp32 -> bf16 hi/lo planes ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
split_bf16_kernel(const float* __restrict__ src, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, size_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t n4 = n / 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 x = __ldg((const float4*)src + i);
        const float v[4] = {x.x, x.y, x.z, x.w};
        split_store4(hi + 4 * i, lo + 4 * i, v);
    }
    for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const __nv_bfloat16 h = __float2bfloat16_rn(src[i]);
        hi[i] = __bfloat16_as_ushort(h);
        lo[i] = __bfloat16_as_ushort(__float2bfloat16_rn(src[i] - __bfloat162float(h)));
    }
}

// ---- host side -------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else
            cudaGetLastError();
    }
    return fn;
}

// stored matrix [rows, cols] bf16, two planes (hi, lo) `plane` elements apart; box = {64 elements, box_rows, 2 planes}
bool make_map(EncodeTiledFn enc, CUtensorMap* map, const uint16_t* ptr, int rows, int cols, long long plane, int box_rows) {
    cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 2};
    cuuint64_t gstride[2] = {(cuuint64_t)cols * 2, (cuuint64_t)plane * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)ptr, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <int BN, bool A_MN, bool B_MN, bool DRAIN>
int launch_t(tp_ctx* ctx, const tp::Bx3Launch* const* Ls, int count, bool pdl, const tpfold::FoldStep* tail, int tail_workers,
             const tp::Bx3Push* push) {
    auto kern = gemm_bx3_kernel<BN, A_MN, B_MN, DRAIN>;
    constexpr int smem = Smem<BN>::kTotal;
    static bool attr_set[16] = {};                    // per device
    if (ctx->device < 16 && !attr_set[ctx->device]) {
        TP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set[ctx->device] = true;
    } else if (ctx->device >= 16) {
        TP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    Bx3Group g{};
    g.count = count;
    if (push) g.push = *push;
    int rows = 0, max_tn = 0;
    for (int i = 0; i < count; ++i) {
        const tp::Bx3Launch& L = *Ls[i];
        Bx3Params& p = g.p[i];
        p.m = L.m; p.n = L.n; p.k = L.k; p.splits = L.splits;
        p.alpha = L.alpha; p.beta = L.beta;
        p.c = L.c; p.c_split = L.c_split; p.c_plane = L.c_plane;
        p.bias = L.bias; p.relu_mask = L.relu_mask; p.relu = L.relu;
        p.colsum_part = L.colsum_part;
        p.stamp = L.stamp;
        p.tiles_n = L.tiles_n;
        p.y0 = rows;
        p.a_early = L.a_early ? 1 : 0; p.b_early = L.b_early ? 1 : 0;
        rows += L.tiles_m;
        if (L.tiles_n > max_tn) max_tn = L.tiles_n;
    }
    g.tail_y0 = rows;
    g.tail_workers = 0;
    if (tail && tail_workers > 0) {
        g.tail_workers = tail_workers;
        g.tail = *tail;
        rows += tail_workers;
    }
    if (rows > 65535) return TP_ERR_UNSUPPORTED;
    const tp::Bx3Launch& L0 = *Ls[0];
    const tp::Bx3Launch& L1 = *Ls[count > 1 ? 1 : 0];
    const tp::Bx3Launch& L2 = *Ls[count > 2 ? 2 : 0];
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(max_tn, rows, L0.splits);
    cfg.blockDim = dim3(threads_of<DRAIN>());
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (L0.splits > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 1;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = L0.splits;
        ++na;
    }
    if (pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    TP_CUDA(cudaLaunchKernelEx(&cfg, kern, L0.ma, L0.mb, L1.ma, L1.mb, L2.ma, L2.mb, g));
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

template <int BN, bool DRAIN>
int launch_major(tp_ctx* ctx, const tp::Bx3Launch* const* Ls, int count, bool pdl, const tpfold::FoldStep* tail, int tw, const tp::Bx3Push* push) {
    const tp::Bx3Launch& L = *Ls[0];
    if (!L.a_mn && !L.b_mn) return launch_t<BN, false, false, DRAIN>(ctx, Ls, count, pdl, tail, tw, push);
    if (!L.a_mn && L.b_mn) return launch_t<BN, false, true, DRAIN>(ctx, Ls, count, pdl, tail, tw, push);
    if (L.a_mn && !L.b_mn) return launch_t<BN, true, false, DRAIN>(ctx, Ls, count, pdl, tail, tw, push);
    return launch_t<BN, true, true, DRAIN>(ctx, Ls, count, pdl, tail, tw, push);
}

template <int BN>
int max_clusters_of(int cluster) {
    int n = -1;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = cluster;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.blockDim = dim3(threads_of<false>());
    cfg.gridDim = dim3(1, 1, cluster);
    auto k = gemm_bx3_kernel<BN, false, false, false>;
    cfg.dynamicSmemBytes = Smem<BN>::kTotal;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<BN>::kTotal);
    if (cudaOccupancyMaxActiveClusters(&n, k, &cfg) != cudaSuccess) { cudaGetLastError(); n = -1; }
    return n;
}

struct Bx3State {
    int max_clusters[3][4];          // [BN 64/128/256][cluster 1/2/4/8]
};

Bx3State* state_of(tp_ctx* ctx) {
    if (ctx->bx3_state) return (Bx3State*)ctx->bx3_state;
    Bx3State* st = new Bx3State();
    const int dflt[4] = {ctx->sm_count, ctx->sm_count / 2, ctx->sm_count / 4 - 4, ctx->sm_count / 8 - 3};
    for (int b = 0; b < 3; ++b) {
        st->max_clusters[b][0] = ctx->sm_count;
        for (int si = 1; si < 4; ++si) {
            int q = b == 0 ? max_clusters_of<64>(1 << si) : b == 1 ? max_clusters_of<128>(1 << si) : max_clusters_of<256>(1 << si);
            st->max_clusters[b][si] = q > 0 ? q : dflt[si];
        }
    }
    ctx->bx3_state = st;
    return st;
}

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

}  // namespace

namespace tp {

void gemm_bx3_destroy(tp_ctx* ctx) {
    if (ctx->bx3_state) delete (Bx3State*)ctx->bx3_state;
    ctx->bx3_state = nullptr;
}

size_t bx3_split_elems(size_t n) { return 2 * ((n + 7) & ~(size_t)7); }

int split_bf16(tp_ctx* ctx, const float* src, uint16_t* dst, size_t n, long long plane, bool pdl) {
    if (!n) return TP_OK;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid_for(ctx, (n + 3) / 4, 256));
    cfg.blockDim = dim3(256);
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    TP_CUDA(cudaLaunchKernelEx(&cfg, split_bf16_kernel, src, dst, dst + plane, n));
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

namespace {
// SM cycles of one CTA wave-by-wave: waves * (fixed + k-blocks per CTA * t_kb(BN) + fold(S, BN))
long bx3_cost(const Bx3State* st, int bi, int si, long tiles, int kblocks) {
    static const long tkb[3] = {1150, 1530, 3200};    // measured (B200, 1024^3): cycles per k-block incl. pipeline stalls
    static const int cands[3] = {64, 128, 256};
    const int sp = 1 << si;
    const long kbpc = (kblocks + sp - 1) / sp;
    const long waves = (tiles + st->max_clusters[bi][si] - 1) / st->max_clusters[bi][si];
    return waves * (5000 + kbpc * tkb[bi] + (sp > 1 ? 1000 + 4 * cands[bi] + 500 * sp : 0));
}
}  // namespace

// K-split for `tiles` output tiles of width bn that will share one launch (bx3_launch_group)
int bx3_best_splits(tp_ctx* ctx, int bn, long tiles, int k) {
    cudaSetDevice(ctx->device);
    Bx3State* st = state_of(ctx);
    const int bi = bn == 64 ? 0 : bn == 128 ? 1 : 2;
    const int kblocks = (k + BK - 1) / BK;
    int best_s = 1;
    long best = -1;
    for (int si = 3; si >= 0; --si) {
        const int sp = 1 << si;
        if (sp > 1 && kblocks < 2 * sp) continue;
        const long c = bx3_cost(st, bi, si, tiles, kblocks);
        if (best < 0 || c < best) { best = c; best_s = sp; }
    }
    return best_s;
}

int bx3_prepare(tp_ctx* ctx, int ta, int tb, int m, int n, int k, float alpha, const uint16_t* a_split, long long a_plane,
                const uint16_t* b_split, long long b_plane, float beta, float* c, const Bx3Epilogue& ep, Bx3Launch* L,
                int want_bn, int want_splits) {
    if (m <= 0 || n <= 0 || k <= 0) return TP_ERR_UNSUPPORTED;
    const int a_rows = ta ? k : m, a_cols = ta ? m : k;       // A stored [m,k] (N) or [k,m] (T)
    const int b_rows = tb ? n : k, b_cols = tb ? k : n;       // B stored [k,n] (N) or [n,k] (T)
    // TMA: 16-byte aligned bases, row pitches and plane strides; the epilogue stores whole float4 vectors
    if ((a_cols & 7) || (b_cols & 7) || (n & 3) || (a_plane & 7) || (b_plane & 7)) return TP_ERR_UNSUPPORTED;
    if (((uintptr_t)a_split | (uintptr_t)b_split | (uintptr_t)c | (uintptr_t)ep.c_split | (uintptr_t)ep.bias | (uintptr_t)ep.relu_mask) & 15)
        return TP_ERR_UNSUPPORTED;
    if (!c && !ep.c_split) return TP_ERR_UNSUPPORTED;
    if (beta != 0.0f && !c) return TP_ERR_UNSUPPORTED;
    EncodeTiledFn enc = get_encode();
    if (!enc) return TP_ERR_UNSUPPORTED;
    cudaSetDevice(ctx->device);
    Bx3State* st = state_of(ctx);
    const int kblocks = (k + BK - 1) / BK;
    const int tiles_m = (m + BM - 1) / BM;
    // Tile width BN and K-split S (cluster size) from a cost model in SM cycles per CTA:
    //   waves * (fixed + k-blocks per CTA * t_kb(BN) + fold(S, BN)),  waves = ceil(tiles / co-resident clusters of S)
    // t_kb: 12 bf16 MMAs of 128 x BN x 16 per k-block (tensor-bound for BN = 256, shared-memory-port-bound below).
    static const int env_bn = env_int("TAPER_BX3_BN", 0), env_s = env_int("TAPER_BX3_SPLITS", 0);
    const int force_bn = want_bn ? want_bn : env_bn, force_s = want_splits ? want_splits : env_s;
    const int cands[3] = {64, 128, 256};
    int bn = 128, splits = 1;
    long best = -1;
    for (int bi = 2; bi >= 0; --bi) {
        const int cand = cands[bi];
        if (force_bn ? cand != force_bn : (cand > 64 && n <= cand / 2)) continue;
        const long t = (long)tiles_m * ((n + cand - 1) / cand);
        for (int si = 3; si >= 0; --si) {
            const int sp = 1 << si;
            if (force_s ? sp != force_s : (sp > 1 && kblocks < 2 * sp)) continue;
            if (sp > kblocks) continue;
            const long kbpc = (kblocks + sp - 1) / sp;
            if (cand == 256 && kbpc * BK > 2048) continue;       // no register drain for 256-wide tiles: bound the accumulation depth
            const long cost = bx3_cost(st, bi, si, t, kblocks);
            if (best < 0 || cost < best) { best = cost; bn = cand; splits = sp; }
        }
    }
    if (best < 0) return TP_ERR_UNSUPPORTED;
    const int tiles_n = (n + bn - 1) / bn;
    if (tiles_m > 65535) return TP_ERR_UNSUPPORTED;
    L->a_mn = ta != 0;
    L->b_mn = tb == 0;
    if (!make_map(enc, &L->ma, a_split, a_rows, a_cols, a_plane, L->a_mn ? 64 : BM)) return TP_ERR_UNSUPPORTED;
    if (!make_map(enc, &L->mb, b_split, b_rows, b_cols, b_plane, L->b_mn ? 64 : bn)) return TP_ERR_UNSUPPORTED;
    L->m = m; L->n = n; L->k = k;
    L->bn = bn; L->splits = splits; L->tiles_m = tiles_m; L->tiles_n = tiles_n;
    const int kbpc = (kblocks + splits - 1) / splits;
    L->drain = bn <= 128 && kbpc * BK > 1024;                    // deep accumulations: round-to-nearest drain every 256 of K
    L->alpha = alpha; L->beta = beta;
    L->c = c;
    L->c_split = ep.c_split; L->c_plane = ep.c_plane ? ep.c_plane : (long long)m * n;
    L->bias = ep.bias; L->relu_mask = ep.relu_mask; L->relu = ep.relu;
    L->colsum_part = ep.colsum_part;
    L->stamp = nullptr;
    L->a_early = L->b_early = false;
    return TP_OK;
}

static int launch_any(tp_ctx* ctx, const Bx3Launch* const* Ls, int count, bool pdl, const tpfold::FoldStep* tail = nullptr, int tw = 0,
                      const Bx3Push* push = nullptr) {
    cudaSetDevice(ctx->device);
    const Bx3Launch& L = *Ls[0];
    if (L.bn == 256) return launch_major<256, false>(ctx, Ls, count, pdl, tail, tw, push);
    if (L.bn == 128) return L.drain ? launch_major<128, true>(ctx, Ls, count, pdl, tail, tw, push) : launch_major<128, false>(ctx, Ls, count, pdl, tail, tw, push);
    return L.drain ? launch_major<64, true>(ctx, Ls, count, pdl, tail, tw, push) : launch_major<64, false>(ctx, Ls, count, pdl, tail, tw, push);
}

int bx3_launch(tp_ctx* ctx, const Bx3Launch& L, bool pdl) {
    const Bx3Launch* one[1] = {&L};
    return launch_any(ctx, one, 1, pdl);
}

// problems that share tile width, operand majors, K-split and drain mode go out as ONE launch (up to three); the rest follow
// one by one.  The tail (fold + bookkeeping of the wide plan) joins the last launch when that launch's tiles leave at least
// four SMs free in its first wave and it is not a K-split (cluster) launch; otherwise the caller launches it by itself.
int bx3_launch_group(tp_ctx* ctx, const Bx3Launch* const* Ls, int count, bool pdl, const void* tail, bool* tail_done, const Bx3Push* push) {
    if (tail_done) *tail_done = false;
    int i = 0;
    while (i < count) {
        int j = i + 1;
        while (j < count && j - i < kMaxGroup && Ls[j]->bn == Ls[i]->bn && Ls[j]->a_mn == Ls[i]->a_mn && Ls[j]->b_mn == Ls[i]->b_mn &&
               Ls[j]->splits == Ls[i]->splits && Ls[j]->drain == Ls[i]->drain && Ls[j]->k == Ls[i]->k)
            ++j;
        int tw = 0;
        if (tail && j == count && Ls[i]->splits == 1 && !Ls[i]->drain) {
            long ctas = 0;
            for (int q = i; q < j; ++q) ctas += (long)Ls[q]->tiles_m * Ls[q]->tiles_n;
            const long free_sms = (long)ctx->sm_count - ctas % ctx->sm_count;
            if (ctas % ctx->sm_count != 0 && free_sms >= 4) tw = (int)(free_sms < 16 ? free_sms : 16);
        }
        int rc = launch_any(ctx, Ls + i, j - i, pdl, tw ? static_cast<const tpfold::FoldStep*>(tail) : nullptr, tw, push);
        if (rc) return rc;
        if (tw && tail_done) *tail_done = true;
        i = j;
    }
    return TP_OK;
}

// sgemm_rowmajor on fp32 operands in bf16x3 mode: split both operands into temporaries, then the pre-split kernel
int gemm_bx3(tp_ctx* ctx, int ta, int tb, int m, int n, int k, float alpha, const float* a, const float* b, float beta,
             float* c, const Epilogue& ep) {
    if (m <= 0 || n <= 0) return TP_OK;
    if (k < 16 || (long)m * n * k < (1L << 16)) return TP_ERR_UNSUPPORTED;
    const int a_cols = ta ? m : k, b_cols = tb ? k : n;
    if ((a_cols & 7) || (b_cols & 7) || (n & 3)) return TP_ERR_UNSUPPORTED;
    if (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) return TP_ERR_UNSUPPORTED;
    const size_t na = (size_t)m * k, nb = (size_t)k * n;
    const long long pa = (long long)((na + 7) & ~(size_t)7), pb = (long long)((nb + 7) & ~(size_t)7);
    tp_buf *sa = nullptr, *sb = nullptr;
    int rc = tp_buf_alloc(ctx, (size_t)pa, &sa);               // pa bf16 pairs = pa floats of storage
    if (!rc) rc = tp_buf_alloc(ctx, (size_t)pb, &sb);
    Bx3Launch L;
    Bx3Epilogue e;
    e.bias = ep.bias; e.relu = ep.relu; e.relu_mask = ep.relu_mask;
    if (!rc) rc = bx3_prepare(ctx, ta, tb, m, n, k, alpha, (const uint16_t*)sa->ptr, pa, (const uint16_t*)sb->ptr, pb, beta, c, e, &L);
    if (!rc) rc = split_bf16(ctx, a, (uint16_t*)sa->ptr, na, pa, false);
    if (!rc) rc = split_bf16(ctx, b, (uint16_t*)sb->ptr, nb, pb, false);
    if (!rc) rc = bx3_launch(ctx, L, false);
    if (sa) tp_buf_release(sa);
    if (sb) tp_buf_release(sb);
    return rc;
}

}  // namespace tp

extern "C" int tpdbg_bx3_times(long long* out16) {
    return cudaMemcpyFromSymbol(out16, g_bx3_t, sizeof(long long) * 16) == cudaSuccess ? 0 : 1;
}

extern "C" {

int tp_split_bf16(tp_ctx* ctx, const tp_buf* src, tp_buf* dst, size_t n) {
    TP_CHECK_ARG(ctx, "tp_split_bf16: NULL ctx");
    TP_NEED(src, n, "src");
    const size_t plane = (n + 7) & ~(size_t)7;
    TP_NEED(dst, plane, "dst");
    return tp::split_bf16(ctx, src->ptr, (uint16_t*)dst->ptr, n, (long long)plane, false);
}

}  // extern "C"

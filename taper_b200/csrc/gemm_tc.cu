// TF32 tensor-core GEMM for sm_100a: tcgen05.mma (kind::tf32) with the accumulator in TMEM, operands
// staged in 128B-swizzled shared memory by TMA, mbarrier producer/consumer pipeline, warp-specialised
// (warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-5 = operand splitter + epilogue).
//
// Replaces matrixmultiply::sgemm / cblas_sgemm behind sgemm_rowmajor (src/gemm.rs:8-49, 72-119) for the
// three call shapes of the reference: (N,N) forward, (N,T) dA, (T,N) dB (src/ops.rs:215-226, 254-265,
// 280-291) plus the fused Linear (N,T).  No operand is ever transposed in memory: a transposed operand is
// described to the tensor core as MN-major (instruction-descriptor bits 15/16), a plain one as K-major.
//
// Math modes
//   1xTF32 (mode 2): operands are read as fp32 and used at TF32 precision (10-bit mantissa).
//   3xTF32 (mode 1): x = hi + lo with hi = x truncated to TF32 (what the tensor core does to an fp32 operand on its
//                    own) and lo = x - hi; D = A_lo*B_hi + A_hi*B_lo + A_hi*B_hi with fp32 accumulation: fp32-accurate
//                    products (~2^-21 relative) at one third of the MMA rate.  The lo tiles are produced in shared
//                    memory by four dedicated warps, so HBM traffic is unchanged.
// Small problems are split along K over a thread-block cluster (2/4/8 CTAs along grid z): every CTA parks its
// partial accumulator tile in its own shared memory and, after a cluster barrier, each CTA folds one slice of the
// rows over all peers through distributed shared memory in a fixed order (deterministic, no global scratch, one
// launch).  The epilogue always goes TMEM -> shared memory -> coalesced 128-bit global stores.
#include "common.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>

extern "C" int tpdbg_max_clusters(int bn, int cluster);

namespace {

constexpr int BM = 128;              // UMMA M (cta_group::1)
constexpr int BK = 32;               // floats per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 8;            // tf32: 32 bytes per instruction
constexpr int kMaxConvK = 1536;      // implicit-GEMM conv: largest contraction (Cin*kh*kw) the per-k table behind the barriers holds

using namespace tcptx;

// Shared-memory matrix descriptor (descriptor version 1 for sm_100):
//   bits [0,14) start address >> 4, [16,30) leading byte offset >> 4, [32,46) stride byte offset >> 4,
//   [46,48) version = 1, [61,64) layout type.
// K-major operand : layout type 2 (SWIZZLE_128B, 16-byte chunks XOR row % 8; TMA SWIZZLE_128B).  Rows of 128 B
//                   (32 floats of K); 8 rows form a 1024 B atom -> SBO = 1024; LBO unused.
// MN-major operand: 32-bit operands only exist in layout type 1 (SWIZZLE_128B_BASE32B: 32-byte chunks XOR row % 4;
//                   TMA SWIZZLE_128B_ATOM_32B).  Rows of 128 B (32 floats of M/N) indexed by k; 4 k-rows form a 512 B
//                   atom -> SBO = 512 (next 4 k), LBO = BK*128 = 4096 (next 32 elements of M/N = the next TMA box).
template <bool MN_MAJOR>
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    constexpr uint32_t lbo_bytes = MN_MAJOR ? BK * 128 : 16, sbo_bytes = MN_MAJOR ? 512 : 1024;
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(MN_MAJOR ? 1 : 2) << 61;
    return d;
}

// gather target for padding taps / k >= K of the implicit-GEMM convolution: every load is unconditional (no data-dependent
// branch between the table lookup and the global load), padding simply reads this zero
__device__ float g_zero_pad[4] = {0.0f, 0.0f, 0.0f, 0.0f};

// development aid: per-phase SM-clock timestamps of CTA (0,0,0), read back with tpdbg_gemm_times()
__device__ long long g_dbg_t[16];
#define DBG_T(slot) do { if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_dbg_t[slot] = clock64(); } while (0)

struct EpiArgs {
    const float* bias;
    const float* relu_mask;
    int relu;
};

struct GemmParams {
    int m, n, k;
    int splits;
    float alpha, beta;
    float* c;
    EpiArgs ep;
    // implicit-GEMM convolution (IM2COL kernels): A(m, k) = x[n, ci, oh*sh + kr*dh - ph, ow*sw + kc*dw - pw] or 0
    const float* x;
    tp::ConvShape g;
};

template <int BN, bool SPLIT3, bool CONV = false, bool OCC2 = false>
struct Smem {
    static constexpr int kABytes = BM * BK * 4;
    static constexpr int kBBytes = BN * BK * 4;
    static constexpr int kStageBytes = (kABytes + kBBytes) * (SPLIT3 ? 2 : 1);
    // as deep as ~200 KB allows (the TMA -> split -> MMA -> free round trip is ~2.5k cycles), at most 8.
    // Implicit-GEMM convolution: three stages only — the gather warps, not the MMAs, set the pace, and the shared memory
    // not taken is L1: a tile's input footprint (25-50 KB) is re-read once per filter tap and should hit there, not in L2.
    // OCC2: two stages so that two CTAs fit on an SM (one's setup / epilogue overlaps the other's main loop).
    static constexpr int kStages = OCC2 ? 2 : CONV ? 3 : (200 * 1024 / kStageBytes) > 8 ? 8 : (200 * 1024 / kStageBytes);
    static constexpr int kTotal = kStages * kStageBytes + 1024 /*alignment slack*/ + 512 /*barriers*/;
};

__device__ __forceinline__ float epilogue_elem(float acc, const GemmParams& p, size_t idx, int col) {
    float v = p.alpha * acc;
    if (p.beta != 0.0f) v += p.beta * p.c[idx];
    if (p.ep.bias) v += __ldg(p.ep.bias + col);
    if (p.ep.relu) v = fmaxf(v, 0.0f);
    if (p.ep.relu_mask) v = __ldg(p.ep.relu_mask + idx) > 0.0f ? v : 0.0f;
    return v;
}

// A_MN / B_MN: operand is MN-major in memory (trans_a = 1 / trans_b = 0 of sgemm_rowmajor).
//
// Warp roles        1xTF32 (192 threads)                 3xTF32 (320 threads)
//   warp 0          TMA producer                         TMA producer
//   warp 1          MMA issuer, TMEM owner               MMA issuer, TMEM owner
//   warps 2-5       epilogue                             hi/lo operand splitter
//   warps 6-9       -                                    accumulator drain + epilogue
//
// 3xTF32 is the fp32-parity mode, so it also bounds the depth of the tensor core's own accumulation: the
// hardware accumulator truncates, which biases long sums of same-signed products (MNIST pixels, post-ReLU
// activations) by ~K/3 * 2^-24.  Every kChunk k-blocks the TMEM accumulator (double-buffered) is drained
// into fp32 registers with round-to-nearest adds while the next chunk is already being multiplied.
// GATHER2 (implicit-GEMM convolution, BN <= 64): a second group of four gather warps (10-13) takes the upper half of every
// k-block, doubling the global loads in flight — the kernel is bound by the gather, not by the tensor pipe.
// OCC2 (implicit-GEMM convolution): two-stage pipeline and a register cap so that two CTAs share an SM — every tile is short
// (9-18 k-blocks), so per-CTA setup and epilogue are a large share of its life and are hidden behind the other CTA's main loop.
template <int BN, bool A_MN, bool B_MN, bool SPLIT3, bool IM2COL = false, bool GATHER2 = false, bool OCC2 = false>
__global__ void __launch_bounds__(SPLIT3 ? (GATHER2 ? 448 : 320) : 192, OCC2 ? 2 : 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, GemmParams p) {
    static_assert(!IM2COL || (SPLIT3 && !A_MN), "the im2col gather is done by the 3xTF32 splitter warps into a K-major A tile");
    static_assert(!GATHER2 || IM2COL, "the second gather group only exists in the implicit-GEMM convolution");
    using S = Smem<BN, SPLIT3, IM2COL, OCC2>;
    constexpr int kStages = S::kStages;
    constexpr int kChunk = 4;                         // k-blocks (128 elements of K) per tensor-core accumulation
    constexpr uint32_t kTmemCols = SPLIT3 ? (2 * BN < 32 ? 32 : 2 * BN) : (BN < 32 ? 32 : BN);
    constexpr int kPitch = BN + 4;                    // floats per staged accumulator row
    constexpr int kEpiWarp0 = SPLIT3 ? 6 : 2;         // first of the four epilogue warps
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // swizzle atoms need 1024 B alignment; offsetting the array (instead of rounding a generic pointer) keeps every
    // access below in the shared address space (LDS/STS, not generic LD/ST)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full_bar = (uint64_t*)(smem + kStages * S::kStageBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* split_bar = empty_bar + kStages;        // 3xTF32: hi/lo tiles ready for the MMA warp
    uint64_t* acc_full = split_bar + kStages;         // [2] accumulator buffer complete (tcgen05.commit)
    uint64_t* acc_empty = acc_full + 2;               // [2] accumulator buffer drained (3xTF32)
    uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);
    int* ktab = (int*)(((uintptr_t)(tmem_slot + 4) + 15) & ~(uintptr_t)15);      // IM2COL: per contraction index k: (input offset relative to the row's base) << 5 | tap

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int total_kb = (p.k + BK - 1) / BK;
    const int kb0 = (int)(((long long)blockIdx.z * total_kb) / p.splits);            // balanced: sizes differ by at most one
    const int kb1 = (int)(((long long)(blockIdx.z + 1) * total_kb) / p.splits);
    const int nkb = kb1 - kb0;                        // >= 1 because splits <= total_kb
    if (threadIdx.x == 0) DBG_T(0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar + s, 1);
            mbar_init(empty_bar + s, 1);
            mbar_init(split_bar + s, GATHER2 ? 256 : 128);      // every thread of the splitter / gather warps arrives
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(acc_full + b, 1);
            mbar_init(acc_empty + b, 4);              // one arrival per drain warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (IM2COL) {
        const int khw = p.g.kh * p.g.kw;
        const int kpad = (p.k + BK - 1) / BK * BK;
        for (int k = threadIdx.x; k < kpad; k += blockDim.x) {
            const int ci = k / khw, r = k - ci * khw, kr = r / p.g.kw, kc = r - kr * p.g.kw;
            // tap 31 is never valid (kh*kw <= 31): marks k >= K
            ktab[k] = k < p.k ? ((((ci * p.g.h + kr * p.g.dh) * p.g.w + kc * p.g.dw) << 5) | r) : 31;
        }
    }
    if (warp == 1) {                                  // TMEM allocation is owned by the MMA warp
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) DBG_T(1);

    auto a_hi = [&](int s) { return smem + s * S::kStageBytes; };
    auto b_hi = [&](int s) { return smem + s * S::kStageBytes + S::kABytes; };
    auto a_lo = [&](int s) { return smem + s * S::kStageBytes + S::kABytes + S::kBBytes; };
    auto b_lo = [&](int s) { return smem + s * S::kStageBytes + 2 * S::kABytes + S::kBBytes; };
    float* stage = (float*)smem;                      // accumulator tile staging (the pipeline stages are free by then)

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % kStages;
                const uint32_t ph = (i / kStages) & 1;
                mbar_wait(empty_bar + s, ph ^ 1);
                mbar_expect_tx(full_bar + s, (IM2COL ? 0 : S::kABytes) + S::kBBytes);
                const int k0 = (kb0 + i) * BK;
                if (IM2COL) {
                    // A is gathered by warps 2-5
                } else if (A_MN) {
#pragma unroll
                    for (int g = 0; g < BM / 32; ++g) tma_load_2d(a_hi(s) + g * (BK * 128), &map_a, full_bar + s, m0 + 32 * g, k0);
                } else {
                    tma_load_2d(a_hi(s), &map_a, full_bar + s, k0, m0);
                }
                if (B_MN) {
#pragma unroll
                    for (int g = 0; g < BN / 32; ++g) tma_load_2d(b_hi(s) + g * (BK * 128), &map_b, full_bar + s, n0 + 32 * g, k0);
                } else {
                    tma_load_2d(b_hi(s), &map_b, full_bar + s, k0, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one elected lane) =====
        if (lane == 0) {
            // instruction descriptor: c = F32 (bit 4), a/b = TF32 (2 << 7, 2 << 10), majors (15, 16), N >> 3 (17), M >> 4 (24)
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                                   ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            constexpr uint32_t kStepA = A_MN ? 1024 : UMMA_K * 4;      // bytes per UMMA_K step inside a stage
            constexpr uint32_t kStepB = B_MN ? 1024 : UMMA_K * 4;
            uint32_t tmem_d = tmem_base;
            for (int i = 0; i < nkb; ++i) {
                const int s = i % kStages;
                const uint32_t ph = (i / kStages) & 1;
                uint32_t fresh = (i == 0) ? 1u : 0u;  // first MMA of an accumulation overwrites
                if (SPLIT3 && i % kChunk == 0) {
                    const int chunk = i / kChunk, b = chunk & 1;
                    mbar_wait(acc_empty + b, ((chunk >> 1) & 1) ^ 1);      // buffer drained by the epilogue warps
                    tc_fence_after();
                    tmem_d = tmem_base + b * BN;
                    fresh = 1u;
                }
                mbar_wait((SPLIT3 ? split_bar : full_bar) + s, ph);
                tc_fence_after();
                if (i == 0) DBG_T(2);
                if (i == nkb - 1) DBG_T(3);
                const uint32_t ah = smem_u32(a_hi(s)), bh = smem_u32(b_hi(s));
                if (SPLIT3) {
                    const uint32_t al = smem_u32(a_lo(s)), bl = smem_u32(b_lo(s));
#pragma unroll
                    for (int kk = 0; kk < BK / UMMA_K; ++kk)          // small terms first
                        tc_mma_tf32(tmem_d, make_desc<A_MN>(al + kk * kStepA), make_desc<B_MN>(bh + kk * kStepB), idesc,
                                    (fresh && kk == 0) ? 0u : 1u);
#pragma unroll
                    for (int kk = 0; kk < BK / UMMA_K; ++kk)
                        tc_mma_tf32(tmem_d, make_desc<A_MN>(ah + kk * kStepA), make_desc<B_MN>(bl + kk * kStepB), idesc, 1u);
#pragma unroll
                    for (int kk = 0; kk < BK / UMMA_K; ++kk)
                        tc_mma_tf32(tmem_d, make_desc<A_MN>(ah + kk * kStepA), make_desc<B_MN>(bh + kk * kStepB), idesc, 1u);
                } else {
#pragma unroll
                    for (int kk = 0; kk < BK / UMMA_K; ++kk)
                        tc_mma_tf32(tmem_d, make_desc<A_MN>(ah + kk * kStepA), make_desc<B_MN>(bh + kk * kStepB), idesc,
                                    (fresh && kk == 0) ? 0u : 1u);
                }
                tc_commit(empty_bar + s);             // frees the smem stage when these MMAs retire
                if (SPLIT3 ? (i % kChunk == kChunk - 1 || i == nkb - 1) : (i == nkb - 1))
                    tc_commit(acc_full + (SPLIT3 ? ((i / kChunk) & 1) : 0));      // accumulation complete
            }
        }
    } else if (SPLIT3 && (warp < 6 || warp >= 10)) {
        // ===== warps 2-5 (3xTF32): split every landed stage into hi/lo tiles; warps 10-13: second gather group (GATHER2) =====
        const int half = warp >= 10 ? 1 : 0;          // which half of every k-block this gather group owns (GATHER2)
        const int t = (threadIdx.x - 64) & 127;       // 0..127: row of the A tile
        // IM2COL: this thread owns row t of the A tile = output pixel m0 + t; its 32 k-values per k-block are gathered from
        // the NCHW input (lanes = consecutive ow: coalesced) and written as hi (the fp32 value) and lo = x - trunc_tf32(x)
        // straight into the K-major SWIZZLE_128B layout the MMA descriptor expects (16-byte chunk j of row r at j ^ (r % 8)).
        const float* xb = nullptr;
        uint32_t tapmask = 0;
        if (IM2COL) {
            const int gm = m0 + t;
            if (gm < p.m) {
                const int ow = gm % p.g.wo, tq = gm / p.g.wo, oh = tq % p.g.ho, nb = tq / p.g.ho;
                const int ih0 = oh * p.g.sh - p.g.ph, iw0 = ow * p.g.sw - p.g.pw;
                xb = p.x + ((ptrdiff_t)nb * p.g.c * p.g.h + ih0) * p.g.w + iw0;
                for (int kr = 0; kr < p.g.kh; ++kr)
                    for (int kc = 0; kc < p.g.kw; ++kc) {
                        const int ih = ih0 + kr * p.g.dh, iw = iw0 + kc * p.g.dw;
                        if (ih >= 0 && ih < p.g.h && iw >= 0 && iw < p.g.w) tapmask |= 1u << (kr * p.g.kw + kc);
                    }
            }
        }
        if constexpr (IM2COL) {
            // two k-blocks of gathers are in flight per thread: block i + 1 is issued before block i is written out
            constexpr int kC = GATHER2 ? BK / 8 : BK / 4;      // 16-byte chunks of a row per gather thread and k-block
            const int c0 = GATHER2 ? half * kC : 0;
            auto gather = [&](int i, float (&v)[BK]) {
                const int4* kt4 = reinterpret_cast<const int4*>(ktab + (kb0 + i) * BK) + c0;
#pragma unroll
                for (int c = 0; c < kC; ++c) {
                    const int4 kt = kt4[c];
                    const int e4[4] = {kt.x, kt.y, kt.z, kt.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float* src = ((tapmask >> (e4[e] & 31)) & 1u) ? xb + (e4[e] >> 5) : g_zero_pad;
                        v[4 * c + e] = __ldg(src);
                    }
                }
            };
            auto commit = [&](int i, const float (&v)[BK]) {
                const int s = i % kStages;
                const uint32_t ph = (i / kStages) & 1;
                mbar_wait(empty_bar + s, ph ^ 1);     // the MMAs that read this stage last time have retired
                uint8_t* rhi = a_hi(s) + t * 128;
                uint8_t* rlo = a_lo(s) + t * 128;
#pragma unroll
                for (int c = 0; c < kC; ++c) {
                    float4 l;
                    l.x = v[4 * c + 0] - __uint_as_float(__float_as_uint(v[4 * c + 0]) & 0xFFFFE000u);
                    l.y = v[4 * c + 1] - __uint_as_float(__float_as_uint(v[4 * c + 1]) & 0xFFFFE000u);
                    l.z = v[4 * c + 2] - __uint_as_float(__float_as_uint(v[4 * c + 2]) & 0xFFFFE000u);
                    l.w = v[4 * c + 3] - __uint_as_float(__float_as_uint(v[4 * c + 3]) & 0xFFFFE000u);
                    const int pc = ((c0 + c) ^ (t & 7)) * 16;
                    *(float4*)(rhi + pc) = make_float4(v[4 * c + 0], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                    *(float4*)(rlo + pc) = l;
                }
                mbar_wait(full_bar + s, ph);          // the weight tile has landed: split it (each group takes half of it)
                const float4* bh4 = (const float4*)b_hi(s);
                float4* bl4 = (float4*)b_lo(s);
                for (int q = t + (GATHER2 ? half * 128 : 0); q < S::kBBytes / 16; q += (GATHER2 ? 256 : 128)) {
                    const float4 x = bh4[q];
                    float4 l;
                    l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
                    l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
                    l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
                    l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
                    bl4[q] = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(split_bar + s);
            };
            float va[BK], vb[BK];
            gather(0, va);
            for (int i = 0; i < nkb; i += 2) {
                if (i + 1 < nkb) gather(i + 1, vb);
                commit(i, va);
                if (i + 1 < nkb) {
                    if (i + 2 < nkb) gather(i + 2, va);
                    commit(i + 1, vb);
                }
            }
        } else
        for (int i = 0; i < nkb; ++i) {
            const int s = i % kStages;
            const uint32_t ph = (i / kStages) & 1;
            mbar_wait(full_bar + s, ph);
            // The tensor core truncates its fp32 operands to TF32 by itself (measured: scripts/tf32_round_probe.py), so the
            // landed tile already serves as `hi`; only lo = x - trunc(x) is written, elementwise at identical offsets so the
            // swizzled layout is preserved.
            const float4* hi4 = (const float4*)a_hi(s);       // A and B tiles are contiguous: [A_hi | B_hi | A_lo | B_lo]
            float4* lo4 = (float4*)a_lo(s);
            constexpr int kVec = (S::kABytes + S::kBBytes) / 16;
#pragma unroll 8
            for (int v = t; v < kVec; v += 128) {
                const float4 x = hi4[v];
                float4 l;
                l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
                l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
                l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
                l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
                lo4[v] = l;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
            mbar_arrive(split_bar + s);
        }
    } else {
        // ===== epilogue warps: accumulator -> fp32 registers -> this CTA's shared memory =====
        // Row pitch BN + 4 floats keeps the per-row 128-bit stores of a quarter-warp on distinct banks.
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        if (SPLIT3) {
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
                    for (int j = 0; j < 16; ++j) acc[c0 + j] += __uint_as_float(v[j]);      // round-to-nearest fp32
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
#pragma unroll
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
    // make the tiles visible (cluster-wide when K is split)
    __syncwarp();
    if (threadIdx.x == kEpiWarp0 * 32) DBG_T(4);
    if (p.splits > 1) cluster_sync_all(); else __syncthreads();
    if (threadIdx.x == kEpiWarp0 * 32) DBG_T(5);
    if constexpr (IM2COL) {
        // convolution epilogue: the tile's rows are output pixels, its columns output channels; write y[n, co, oh, ow] (NCHW,
        // what transpose_4d + add_bias_4d (+ relu) produce, src/tensor.rs:1275-1281, 1387-1388) straight from the staged
        // accumulators: for a fixed channel consecutive rows are consecutive addresses, so the stores stay coalesced
        constexpr int kT = GATHER2 ? 448 : 320;
        const int hw = p.g.ho * p.g.wo;
        // one float4 of four channels per thread-iteration: consecutive lanes = consecutive rows, so the shared-memory reads
        // (row pitch BN + 4 floats) are conflict-free and each of the four channel stores is a coalesced run
        for (int idx = threadIdx.x; idx < BM * (BN / 4); idx += kT) {
            const int cq = idx / BM, r = idx - cq * BM;
            const int gm = m0 + r, gc = n0 + 4 * cq;
            if (gm < p.m && gc < p.n) {
                const int nb = gm / hw, pix = gm - nb * hw;
                const float4 a = *(const float4*)(stage + r * kPitch + 4 * cq);
                const float av[4] = {a.x, a.y, a.z, a.w};
                float* yo = p.c + ((size_t)nb * p.n + gc) * hw + pix;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (gc + e < p.n) {
                        float v = av[e];
                        if (p.ep.bias) v += __ldg(p.ep.bias + gc + e);
                        if (p.ep.relu) v = fmaxf(v, 0.0f);
                        yo[(size_t)e * hw] = v;
                    }
                }
            }
        }
    } else {
        // CTA z of the cluster owns rows [z*BM/S, (z+1)*BM/S) of the tile: fold the S partial tiles in split order
        // through distributed shared memory, apply the epilogue, store coalesced.  Every warp of the CTA takes part.
        constexpr int kT = SPLIT3 ? 320 : 192;
        const int t = threadIdx.x;
        const int S_ = p.splits;
        const int rows_per = BM / S_;
        const int r_begin = (int)blockIdx.z * rows_per;
        constexpr int kVecPerRow = BN / 4;
        const uint32_t stage_addr = smem_u32(smem);
        const bool vec_ok = (p.n % 4 == 0);
        const int total_vec = rows_per * kVecPerRow;
        constexpr int kU = 4;                         // independent output vectors in flight per thread
        for (int base = t; base < total_vec; base += kT * kU) {
            float4 acc[kU];
            uint32_t off[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int idx = base + u * kT;
                const int r = r_begin + idx / kVecPerRow, c4 = (idx % kVecPerRow) * 4;
                off[u] = (uint32_t)(r * kPitch + c4) * 4;
                acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (S_ == 1) {
#pragma unroll
                for (int u = 0; u < kU; ++u)
                    if (base + u * kT < total_vec) acc[u] = *(const float4*)(smem + off[u]);
            } else {
                // all loads of a group of two peers (x 4 vectors) are issued before the first add (DSMEM latency ~500 cycles);
                // the adds run in split order 0, 1, 2, ... so the result does not depend on timing
                for (int z0 = 0; z0 < S_; z0 += 2) {
                    float4 x[2][kU];
#pragma unroll
                    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
                        for (int u = 0; u < kU; ++u)
                            if (z0 + dz < S_ && base + u * kT < total_vec)
                                x[dz][u] = ld_cluster_f4(map_to_cta(stage_addr + off[u], (uint32_t)(z0 + dz)));
#pragma unroll
                    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
                        for (int u = 0; u < kU; ++u)
                            if (z0 + dz < S_ && base + u * kT < total_vec) {
                                if (z0 + dz == 0) acc[u] = x[dz][u];
                                else { acc[u].x += x[dz][u].x; acc[u].y += x[dz][u].y; acc[u].z += x[dz][u].z; acc[u].w += x[dz][u].w; }
                            }
                }
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int idx = base + u * kT;
                if (idx >= total_vec) continue;
                const int r = r_begin + idx / kVecPerRow, c4 = (idx % kVecPerRow) * 4;
                const int gm = m0 + r, col = n0 + c4;
                if (gm < p.m && col < p.n) {
                    const size_t rowoff = (size_t)gm * p.n;
                    if (vec_ok && col + 4 <= p.n) {
                        float4 o;
                        o.x = epilogue_elem(acc[u].x, p, rowoff + col + 0, col + 0);
                        o.y = epilogue_elem(acc[u].y, p, rowoff + col + 1, col + 1);
                        o.z = epilogue_elem(acc[u].z, p, rowoff + col + 2, col + 2);
                        o.w = epilogue_elem(acc[u].w, p, rowoff + col + 3, col + 3);
                        *(float4*)(p.c + rowoff + col) = o;
                    } else {
                        const float a4[4] = {acc[u].x, acc[u].y, acc[u].z, acc[u].w};
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (col + j < p.n) p.c[rowoff + col + j] = epilogue_elem(a4[j], p, rowoff + col + j, col + j);
                    }
                }
            }
        }
    }
    // nobody leaves while a peer may still read its tile
    __syncwarp();
    if (threadIdx.x == kEpiWarp0 * 32) DBG_T(6);
    if (p.splits > 1) cluster_sync_all();
    __syncthreads();
    if (threadIdx.x == 0) DBG_T(7);
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

// ---- host side -------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TcState {
    EncodeTiledFn encode = nullptr;
    bool attr_set[5][2][2][2] = {};
    int max_clusters[4] = {148, 74, 33, 15};      // co-resident clusters of 1/2/4/8 CTAs (queried at first use)
};

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

// stored matrix [rows, cols] row-major fp32; box = {32 floats, box_rows}; OOB elements read as zero
bool make_map(EncodeTiledFn enc, CUtensorMap* map, const float* ptr, int rows, int cols, int box_rows, bool mn_major) {
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <int BN, bool A_MN, bool B_MN, bool SPLIT3>
int launch(tp_ctx* ctx, TcState* st, const CUtensorMap& ma, const CUtensorMap& mb, GemmParams& p, dim3 grid) {
    auto kern = gemm_tf32_kernel<BN, A_MN, B_MN, SPLIT3>;
    constexpr int smem = Smem<BN, SPLIT3>::kTotal;
    constexpr int bi = BN == 16 ? 0 : BN == 32 ? 1 : BN == 64 ? 2 : BN == 128 ? 3 : 4;
    if (!st->attr_set[bi][A_MN][B_MN][SPLIT3]) {
        TP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        st->attr_set[bi][A_MN][B_MN][SPLIT3] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(SPLIT3 ? 320 : 192);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;      // the K-splits of one output tile form a cluster
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = grid.z;
    cfg.attrs = attr;
    cfg.numAttrs = grid.z > 1 ? 1 : 0;
    TP_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, p));
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

template <int BN, bool SPLIT3>
int dispatch_major(tp_ctx* ctx, TcState* st, int ta, int tb, const CUtensorMap& ma, const CUtensorMap& mb, GemmParams& p, dim3 grid) {
    const bool a_mn = ta != 0, b_mn = tb == 0;
    if (!a_mn && !b_mn) return launch<BN, false, false, SPLIT3>(ctx, st, ma, mb, p, grid);
    if (!a_mn && b_mn) {
        if constexpr (BN >= 32) return launch<BN, false, true, SPLIT3>(ctx, st, ma, mb, p, grid);
    }
    if (a_mn && !b_mn) return launch<BN, true, false, SPLIT3>(ctx, st, ma, mb, p, grid);
    if (a_mn && b_mn) {
        if constexpr (BN >= 32) return launch<BN, true, true, SPLIT3>(ctx, st, ma, mb, p, grid);
    }
    return TP_ERR_UNSUPPORTED;
}

}  // namespace

namespace tp {

int gemm_tc(tp_ctx* ctx, int ta, int tb, int m, int n, int k, float alpha, const float* a, const float* b, float beta,
            float* c, const Epilogue& ep, int mode) {
    if (m <= 0 || n <= 0) return TP_OK;
    if (k <= 0) return TP_ERR_UNSUPPORTED;
    // TMA needs 16-byte aligned bases and row pitches: the contiguous dimension of each stored operand % 4 == 0
    const int a_rows = ta ? k : m, a_cols = ta ? m : k;       // A stored [m,k] (N) or [k,m] (T)
    const int b_rows = tb ? n : k, b_cols = tb ? k : n;       // B stored [k,n] (N) or [n,k] (T)
    if ((a_cols & 3) || (b_cols & 3) || (((uintptr_t)a | (uintptr_t)b) & 15)) return TP_ERR_UNSUPPORTED;
    if (k < 16 || (long)m * n * k < (1L << 16)) return TP_ERR_UNSUPPORTED;     // not worth a tensor-core tile
    EncodeTiledFn enc = get_encode();
    if (!enc) return TP_ERR_UNSUPPORTED;
    cudaSetDevice(ctx->device);
    TcState* st = (TcState*)ctx->tc_state;
    if (!st) {
        st = new TcState();
        st->encode = enc;
        st->max_clusters[0] = ctx->sm_count;
        for (int si = 1; si < 4; ++si) {
            int q = tpdbg_max_clusters(128, 1 << si);
            if (q > 0) st->max_clusters[si] = q;
        }
        ctx->tc_state = st;
    }
    const bool b_mn = tb == 0;
    const int kblocks = (k + BK - 1) / BK;
    const int tiles_m = (m + BM - 1) / BM;
    // Tile width BN and K-split S (cluster size) from a small cost model, in SM cycles per CTA:
    //   waves * (fixed + k-blocks per CTA * t_kb(BN) + cluster fold), waves = ceil(tiles / co-resident clusters of S)
    // t_kb is shared-memory bound in both modes (operand reads of the MMAs + the lo-tile pass in 3xTF32).  A cluster that
    // does not fit in the first wave doubles the kernel, hence the measured residency limits (148 / 74 / 33 / 15).
    const int bn_min = b_mn ? 32 : 16;
    // 128x256 tiles (1xTF32 only: the 3xTF32 stages would not fit) halve the B re-reads per MMA: with 128x128 SS-mode tiles the
    // MMA operand fetch (128 B/clk) plus the TMA fill (128 B/clk) are twice the 128 B/clk shared-memory port; 128x256 needs 192.
    int bn_max = n <= 16 ? 16 : n <= 32 ? 32 : n <= 64 ? 64 : (n >= 256 && mode == 2) ? 256 : 128;
    if (bn_max < bn_min) bn_max = bn_min;
    int bn = bn_max, splits = 1;
    long best = -1;
    for (int cand = bn_max; cand >= bn_min; cand >>= 1) {
        const long t = (long)tiles_m * ((n + cand - 1) / cand);
        const long tkb = mode == 1 ? 640 + 5 * cand : 256 + 2 * cand;
        for (int si = 3; si >= 0; --si) {
            const int sp = 1 << si;
            if (sp > 1 && kblocks < 2 * sp) continue;               // at least two k-blocks per CTA
            const long waves = (t + st->max_clusters[si] - 1) / st->max_clusters[si];
            const long kbpc = (kblocks + sp - 1) / sp;
            const long cost = waves * (3000 + kbpc * tkb + (sp > 1 ? 2500 : 0));
            if (best < 0 || cost < best) { best = cost; bn = cand; splits = sp; }
        }
    }
    const int tiles_n = (n + bn - 1) / bn;
    if (tiles_m > 65535) return TP_ERR_UNSUPPORTED;

    CUtensorMap ma, mb;
    if (!make_map(enc, &ma, a, a_rows, a_cols, ta ? 32 : BM, ta != 0)) return TP_ERR_UNSUPPORTED;
    if (!make_map(enc, &mb, b, b_rows, b_cols, b_mn ? 32 : bn, b_mn)) return TP_ERR_UNSUPPORTED;

    GemmParams p;
    p.m = m; p.n = n; p.k = k;
    p.splits = splits;
    p.alpha = alpha; p.beta = beta;
    p.c = c;
    p.ep = EpiArgs{ep.bias, ep.relu_mask, ep.relu};
    dim3 grid(tiles_n, tiles_m, splits);
    const bool split3 = mode == 1;
#define TP_BN(BNV)                                                                                          \
    return split3 ? dispatch_major<BNV, true>(ctx, st, ta, tb, ma, mb, p, grid)                             \
                  : dispatch_major<BNV, false>(ctx, st, ta, tb, ma, mb, p, grid)
    switch (bn) {
        case 16: TP_BN(16);
        case 32: TP_BN(32);
        case 64: TP_BN(64);
        case 256: return dispatch_major<256, false>(ctx, st, ta, tb, ma, mb, p, grid);
        default: TP_BN(128);
    }
#undef TP_BN
}

}  // namespace tp

// development aid: how many clusters of `cluster` CTAs of the 3xTF32 K-major kernel can be resident at once
extern "C" int tpdbg_max_clusters(int bn, int cluster) {
    int n = -1;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = cluster;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.blockDim = dim3(320);
    cfg.gridDim = dim3(1, 1, cluster);
#define TP_Q(BNV)                                                                                           \
    {                                                                                                       \
        auto k = gemm_tf32_kernel<BNV, false, false, true>;                                                 \
        cfg.dynamicSmemBytes = Smem<BNV, true>::kTotal;                                                     \
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<BNV, true>::kTotal);      \
        if (cudaOccupancyMaxActiveClusters(&n, k, &cfg) != cudaSuccess) { cudaGetLastError(); n = -1; }     \
    }
    if (bn == 16) TP_Q(16) else if (bn == 32) TP_Q(32) else if (bn == 64) TP_Q(64) else TP_Q(128)
#undef TP_Q
    return n;
}

extern "C" int tpdbg_gemm_times(long long* out16) {
    return cudaMemcpyFromSymbol(out16, g_dbg_t, sizeof(long long) * 16) == cudaSuccess ? 0 : 1;
}

namespace tp {

namespace {
template <int BN>
int launch_conv(tp_ctx* ctx, const CUtensorMap& mb, GemmParams& p, dim3 grid) {
    constexpr bool kG2 = false;                       // second gather group (448 threads): superseded by two CTAs per SM below
    constexpr bool kOcc2 = BN <= 64;                  // two co-resident CTAs: 2 x (2 stages + tables) of shared memory, <= 102 registers
    auto kern = gemm_tf32_kernel<BN, false, true, true, true, kG2, kOcc2>;
    constexpr int smem = Smem<BN, true, true, kOcc2>::kTotal + kMaxConvK * 4 + 64;
    bool& attr_set = ctx->attr_conv[BN == 32 ? 0 : BN == 64 ? 1 : 2];
    if (!attr_set) {
        TP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    kern<<<grid, kG2 ? 448 : 320, smem, ctx->stream>>>(mb /*unused A map*/, mb, p);
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}
}  // namespace

int gemm_tc_conv_fwd(tp_ctx* ctx, const float* x, const float* w2, const float* bias, int relu, float* y, const ConvShape& g) {
    const long long M = (long long)g.n * g.ho * g.wo;
    if (M <= 0 || M > 0x7fffffffLL || g.K < 16 || g.K > kMaxConvK) return TP_ERR_UNSUPPORTED;
    if (g.kh * g.kw > 31 || g.cout < 32 || (g.cout & 3) || ((uintptr_t)w2 & 15)) return TP_ERR_UNSUPPORTED;
    if ((long long)g.c * g.h * g.w >= (1LL << 26)) return TP_ERR_UNSUPPORTED;     // per-image offset must fit 26 bits of the k table
    EncodeTiledFn enc = get_encode();
    if (!enc) return TP_ERR_UNSUPPORTED;
    cudaSetDevice(ctx->device);
    const int bn = g.cout <= 32 ? 32 : g.cout <= 64 ? 64 : 128;
    const int tiles_m = (int)((M + BM - 1) / BM), tiles_n = (g.cout + bn - 1) / bn;
    if (tiles_m > 65535) return TP_ERR_UNSUPPORTED;
    CUtensorMap mb;
    if (!make_map(enc, &mb, w2, g.K, g.cout, 32, true)) return TP_ERR_UNSUPPORTED;     // weights [K, Cout]: MN-major B operand
    GemmParams p{};
    p.m = (int)M; p.n = g.cout; p.k = g.K;
    p.splits = 1;
    p.alpha = 1.0f; p.beta = 0.0f;
    p.c = y;
    p.ep = EpiArgs{bias, nullptr, relu};
    p.x = x;
    p.g = g;
    dim3 grid(tiles_n, tiles_m, 1);
    switch (bn) {
        case 32: return launch_conv<32>(ctx, mb, p, grid);
        case 64: return launch_conv<64>(ctx, mb, p, grid);
        default: return launch_conv<128>(ctx, mb, p, grid);
    }
}

void gemm_tc_destroy(tp_ctx* ctx) {
    TcState* st = (TcState*)ctx->tc_state;
    if (!st) return;
    delete st;
    ctx->tc_state = nullptr;
}

}  // namespace tp

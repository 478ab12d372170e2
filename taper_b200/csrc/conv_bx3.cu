// 3x3 / stride 1 / pad 1 convolution as an implicit GEMM on tcgen05 (bf16x3, fp32 accumulation in TMEM) over
// NHWC activations held as bf16 hi/lo pairs — and the chain of such layers (+ 2x2 max-pool) that the reference's CNNs
// are made of (examples/train_mnist_cnn.rs:35-100; Tensor::conv2d / conv2d_relu src/tensor.rs:1221-1285, 1379-1389;
// max_pool2d :1391-1464), run back to back without ever leaving that layout.
//
// Activation format ("planes"): element (n, y, x, c) of an [N, H, W, C] tensor, C % 32 == 0, lives at bf16 index
//     (((n*H + y)*W + x) * C/32 + c/32) * 64 + hl*32 + c%32,        hl = 0: hi = rn_bf16(v), hl = 1: lo = rn_bf16(v - hi)
// i.e. every pixel's 32-channel block is one 128-byte row [hi x32 | lo x32]: exactly one row of a K-major SWIZZLE_128B
// UMMA operand, same 4 bytes per element as fp32.
//
// Implicit GEMM: M = output pixels, N = C_out, K = 9 taps x C_in.  The A operand is never gathered: ONE TMA box
// {64 bf16, 1 channel block, Wp columns, Hp rows, G images} brings a zero-padded input patch (halo rows via negative /
// out-of-range coordinates, which TMA zero-fills) into shared memory as Hp*Wp consecutive 128-byte rows, and the A tile
// of tap (kr, kc) is that SAME patch read from a row offset:  start = patch + (kr*Wp + kc - 1) * 128 bytes.
// Wp (8 / 16 / 32) is the patch's row pitch in pixels, > W, so column W of every patch row is a zero pixel and serves
// as the right halo of its row and the left halo of the next one.  Row offsets that are not multiples of 8 rows leave
// the 1024-byte swizzle atom alignment; measured on B200 (scripts/conv_stack_probe.py): the tensor core applies the 128-byte
// swizzle to ABSOLUTE shared-memory address bits, exactly as TMA wrote them, so an unaligned start address with the
// descriptor's base-offset field left 0 reads the shifted rows correctly (shift_mode 2, the default), whereas setting the base
// offset to (start >> 7) & 7 (shift_mode 1) double-counts the phase and returns garbage.
// (shift_mode 0, the conservative variant kept for cross-checking: three patches per channel block loaded at x = -1, 0, +1,
// so every tap offset kr*Wp is atom-aligned — 3x the L2 -> shared-memory traffic, which then bounds the 28x28 layers.)
// M tile = 128 consecutive patch rows = 128/Wp image rows (or, for images smaller than that, G whole images); rows that
// fall on pad columns / halo rows compute junk that the epilogue drops (12.5 % of the MMA rows at 28x28, 23 % at 14x14 / 7x7 —
// the kernel is shared-memory-port bound, not MMA bound).
// Products (bf16x3: A_hi*W_hi + A_hi*W_lo + A_lo*W_hi, fp32 accumulation): per tap, 32-channel block and K = 16 step TWO MMAs,
//   A_hi x [W_hi ; W_lo]  (N = 2*C_out: hi*hi lands in accumulator columns [0, C_out), hi*lo in [C_out, 2*C_out))
//   A_lo x  W_hi          (N = C_out, into columns [0, C_out))
// and the epilogue adds the two column halves: the A tile — which dominates the shared-memory port for C_out <= 64 — is read
// twice per product instead of three times, and a third fewer MMAs go through the issuing threads.
// Weights [K = ci*9 + tap, C_out] (the reference's reinterpretation of the [C_out, C_in, 3, 3] buffer) are
// re-laid once per step as [ci/32][tap pair][W_hi rows x C_out ; W_lo rows x C_out][tap 2p x32 | tap 2p+1 x32] (128-byte
// K-major rows again; a tap is the row's first or second half) and either stay resident in shared memory for the whole kernel
// (<= 96 KB) or stream through a second ring, one tap pair per stage.
// Persistent CTAs (one per SM), warp-specialised: warp 0 TMA producer, warps 1-2 MMA issuers (two for C_out <= 64, tiles handled
// in pairs; one for C_out = 128) with TWO TMEM accumulators each, warps 3-6 epilogue (the next tiles multiply while a tile is
// read out): bias + ReLU (+ 2x2 max-pool through a shared-memory staging tile) and either the planes of the next layer or NCHW
// fp32 for the rest of the tape.
#include "common.cuh"
#include "tc_ptx.cuh"

#include <cuda_bf16.h>
#include <cstdlib>

namespace {

using namespace tcptx;

constexpr int kMaxIssuers = 2;             // warps 1..2 can issue MMAs (C_out = 128 uses one: its accumulators fill TMEM two at a time)
constexpr int kEpiWarp0 = 1 + kMaxIssuers;  // first of the four epilogue warps
constexpr int kConvThreads = 32 * (1 + kMaxIssuers + 4);
// MMA issuers and accumulators per issuer for a tile width.  Measured (scripts/conv_timeline.py): with one accumulator per issuer
// an issuer idles while the epilogue drains its tile (pipe utilisation ~60 %); with two it starts the next tile at once.  An
// accumulator is 2 * C_out TMEM columns ([hi*hi + lo*hi | hi*lo]) and TMEM has 512.
template <int BN> struct ConvCfg {
    static constexpr int kIssuers = BN <= 64 ? 2 : 1;
    static constexpr int kAccPer = 2;
    static constexpr int kAccs = kIssuers * kAccPer;
    static constexpr uint32_t kTmemCols = kAccs * 2 * BN < 32 ? 32 : kAccs * 2 * BN;
};
constexpr int kMaxRing = 8;
constexpr int kMaxPrep = 8;

struct ConvP {
    int N, H, W, CB, Cout;           // input [N, H, W, 32*CB] planes
    int Wp, R, Hp, G;                // patch pitch, image rows per tile, patch rows per image, images per tile
    int row_blocks, tiles;
    int img_rows;                    // Hp * Wp: patch rows per image
    int patch_bytes;                 // G * Hp * Wp * 128: bytes one TMA box delivers
    int patch_alloc;                 // bytes reserved per patch (>= patch_bytes and >= what 128 rows + the tap shifts read)
    int shift_mode;                  // 0: three patches (x = -1, 0, +1), atom-aligned taps; 1 / 2: one patch, row-shifted taps (base offset set / 0)
    int dbg;                         // development: bit 0 hi*hi products only, bit 1 epilogue reads the accumulator but stores nothing
    int a_stage_bytes, nA;
    int w_resident, b_stage_bytes, nB;
    int out_mode;                    // 0: planes (next conv), 1: NCHW fp32, 2: global average pool -> mean[N, C] + cnt[N, C]
    int warp_img[4];                 // out_mode 2: the image (within the tile) whose rows epilogue warp q holds, -1: none
    float* out_mean;
    float* out_cnt;                  // may be NULL
    int pool, relu;
    int accumulate;                  // NCHW output: out += result (gradient accumulation into a live grad, src/ops.rs:250-253)
    int Ho, Wo;                      // output spatial size (after the pool)
    const float* bias;
    uint16_t* out_planes;
    float* out_nchw;
};

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

// K-major SWIZZLE_128B operand descriptor, split in two 32-bit words so that stepping through a tile is one 32-bit add:
//   lo = start address >> 4 (bits 0-13) | LBO = 16 B (bit 16, unused by swizzled K-major layouts)
//   hi = SBO = 1024 B (8 rows of 128 B per swizzle atom) | descriptor version 1 (bit 46) | base offset (bits 49-51) | SWIZZLE_128B (2 << 61)
constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return (saddr >> 4) | 0x10000u; }      // saddr < 256 KB
__device__ __forceinline__ void mma_bf16_w(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate)
        : "memory");
}

// two floats -> packed bf16x2 (one cvt.rn.bf16x2.f32), low half = a
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&v);
}
// 8 values -> 16 bytes of hi = rn_bf16(v) and 16 bytes of lo = rn_bf16(v - hi); bf16 -> fp32 is a 16-bit shift
__device__ __forceinline__ void split8(const float* v, uint4* hi, uint4* lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        h[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
        const float r0 = v[2 * j] - __uint_as_float(h[j] << 16), r1 = v[2 * j + 1] - __uint_as_float(h[j] & 0xffff0000u);
        l[j] = pack_bf16(r0, r1);
    }
    *hi = make_uint4(h[0], h[1], h[2], h[3]);
    *lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// Sum 16 per-lane values over the 32 lanes of a warp with 16 shuffles (recursive halving: after the exchange with lane ^ 16 a
// lane keeps 8 of the 16 sums, then 4, 2, 1; a last exchange with lane ^ 1 completes them).  Returns, in every lane, the total
// of value index warp_sum16_index(lane).  Fixed association order: deterministic.
__device__ __forceinline__ int warp_sum16_index(int lane) { return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1); }
__device__ __forceinline__ float warp_sum16(float (&v)[16], int lane) {
#pragma unroll
    for (int n = 8, off = 16; n >= 1; n >>= 1, off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int j = 0; j < n; ++j) {
            const float send = upper ? v[j] : v[j + n];
            const float keep = upper ? v[j + n] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// development aid: SM-clock stamps of CTA 0's first 64 tiles, read back with tpdbg_conv_times():
//   [0] issuer saw its accumulator free  [1] issuer committed the tile  [2] epilogue saw the accumulator full
//   [3] epilogue released the accumulator  [4] producer issued the tile's first patch load  [5] issuer saw the first patch
__device__ long long g_conv_t[6][64];
#define CONV_T(kind, it) do { if (blockIdx.x == 0 && (it) < 64) g_conv_t[kind][it] = clock64(); } while (0)

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int BN>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3x3_bx3_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const __grid_constant__ ConvP p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* a_ring = smem;
    uint8_t* b_ring = a_ring + p.nA * p.a_stage_bytes;
    constexpr int kPairBytes = 2 * BN * 128;                          // one weight stage: a pair of taps, rows [W_hi x BN ; W_lo x BN]
    const int b_total = p.w_resident ? 5 * p.CB * kPairBytes : p.nB * kPairBytes;
    constexpr int kPitch = BN + 4;                                    // floats per staged row (pool)
    float* stage = (float*)(b_ring + b_total);
    uint64_t* bars = (uint64_t*)((uint8_t*)stage + (p.pool ? 128 * kPitch * 4 : p.out_mode == 2 ? 8 * BN * 4 : 0));
    uint64_t* a_full = bars;
    uint64_t* a_empty = a_full + kMaxRing;
    uint64_t* b_full = a_empty + kMaxRing;
    uint64_t* b_empty = b_full + kMaxRing;
    uint64_t* acc_full = b_empty + kMaxRing;                          // [4]
    uint64_t* acc_empty = acc_full + 4;                               // [4]
    uint64_t* w_full = acc_empty + 4;
    uint32_t* tmem_slot = (uint32_t*)(w_full + 1);
    constexpr int kMmaWarps = ConvCfg<BN>::kIssuers, kAccPer = ConvCfg<BN>::kAccPer;
    constexpr uint32_t kTmemCols = ConvCfg<BN>::kTmemCols;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_w);
        for (int s = 0; s < kMaxRing; ++s) {
            mbar_init(a_full + s, 1);
            mbar_init(a_empty + s, 1);
            mbar_init(b_full + s, 1);
            mbar_init(b_empty + s, kMmaWarps);
        }
        for (int b = 0; b < 4; ++b) {
            mbar_init(acc_full + b, 1);
            mbar_init(acc_empty + b, 4);
        }
        mbar_init(w_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (p.shift_mode) {
        // the row in front of every patch is the left halo of the patch's first pixel: zero, never written by TMA
        for (int i = threadIdx.x; i < p.nA * 256; i += kConvThreads)
            *(uint32_t*)(a_ring + (i >> 8) * p.a_stage_bytes + (i & 255) * 4) = 0u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the tensor core's reads
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // everything above overlapped the previous kernel; its output (this layer's input planes) is complete after the wait.
    // The dependents are released only after our own wait, so "complete" is transitive along the chain of launches.
    pdl_wait();
    pdl_launch_dependents();

    const int patch_off = p.shift_mode ? 1024 : 0;

    if (warp == 0) {
        // ===== TMA producer =====
        // Tiles are handled in PAIRS (it = 2P, 2P + 1): issuer w owns tile 2P + w, its own half of the patch ring (stages
        // w*nAe .. w*nAe + nAe - 1, filled and drained strictly in that issuer's order) and accumulator w.  Streamed weights go
        // through ONE ring that both issuers walk in lockstep — every stage (3 taps of one channel block) is multiplied into both
        // tiles of the pair and released by both (b_empty counts 2) — so each ring has exactly one consumer sequence: a parity
        // wait can never be more than one phase away from the barrier's state, and the weights are read from L2 once per pair.
        if (lane == 0) {
            if (p.w_resident) {
                mbar_expect_tx(w_full, 5 * p.CB * kPairBytes);
                for (int g = 0; g < 5 * p.CB; ++g) tma_load_3d(b_ring + g * kPairBytes, &map_w, w_full, 0, 0, g);
            }
            const int nAe = p.nA / kMmaWarps;
            const int my_tiles = ((int)p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
            const int npairs = (my_tiles + kMmaWarps - 1) / kMmaWarps;
            for (int P = 0; P < npairs; ++P) {
                for (int cb = 0; cb < p.CB; ++cb) {
                    const int k = P * p.CB + cb;
                    for (int w = 0; w < kMmaWarps; ++w) {
                        const int tile = blockIdx.x + (P * kMmaWarps + w) * gridDim.x;
                        if (tile >= p.tiles) continue;
                        const int ng = tile / p.row_blocks, rb = tile - ng * p.row_blocks;
                        const int n0 = ng * p.G, y0 = rb * p.R;
                        const int s = w * nAe + k % nAe;
                        mbar_wait(a_empty + s, ((k / nAe) & 1) ^ 1);
                        uint8_t* dst = a_ring + s * p.a_stage_bytes + patch_off;
                        if (cb == 0) CONV_T(4, P * kMmaWarps + w);
                        if (p.shift_mode) {
                            mbar_expect_tx(a_full + s, p.patch_bytes);
                            tma_load_5d(dst, &map_x, a_full + s, 0, cb, 0, y0 - 1, n0);
                        } else {
                            mbar_expect_tx(a_full + s, 3 * p.patch_bytes);
                            for (int kc = 0; kc < 3; ++kc) tma_load_5d(dst + kc * p.patch_alloc, &map_x, a_full + s, 0, cb, kc - 1, y0 - 1, n0);
                        }
                    }
                    if (!p.w_resident) {
                        for (int pr = 0; pr < 5; ++pr) {
                            const int bi = k * 5 + pr;
                            const int sb = bi % p.nB;
                            mbar_wait(b_empty + sb, ((bi / p.nB) & 1) ^ 1);
                            mbar_expect_tx(b_full + sb, kPairBytes);
                            tma_load_3d(b_ring + sb * kPairBytes, &map_w, b_full + sb, 0, 0, cb * 5 + pr);
                        }
                    }
                }
            }
        }
    } else if (warp <= kMaxIssuers) {
        // ===== MMA issuers: warp 1 + w multiplies tile 2P + w of every pair into accumulator w.  One thread issuing all MMAs is
        // latency-bound on its own instruction stream (measured ~95 clocks per MMA against the ~48 the shared-memory port allows
        // for 128 x 32 x 16): two issuers interleave their streams in the tensor pipe. =====
        if (lane == 0 && warp <= kMmaWarps) {
            const int mw = warp - 1;
            // instruction descriptors: D = F32 (1 << 4), A / B = BF16 (1 << 7, 1 << 10), both K-major, N >> 3 at 17, M >> 4 at 24.
            // Per tap and K = 16 step two MMAs instead of three:  A_hi x [W_hi ; W_lo] (N = 2 BN: hi*hi into columns [0, BN), hi*lo
            // into [BN, 2 BN)) and A_lo x W_hi (N = BN, into [0, BN)); the epilogue adds the two column halves.  The A tile — the
            // operand that dominates the shared-memory port for C_out <= 64 — is read twice per product instead of three times.
            const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t idesc2 = idesc_base | ((uint32_t)((2 * BN) >> 3) << 17), idesc1 = idesc_base | ((uint32_t)(BN >> 3) << 17);
            // descriptor-unit (16 B) offsets of the nine taps inside a patch stage
            uint32_t tap_off[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int kr = t / 3, kc = t - 3 * kr;
                tap_off[t] = p.shift_mode ? (uint32_t)((kr * p.Wp + kc - 1) * 8) : (uint32_t)((kc * p.patch_alloc + kr * p.Wp * 128) >> 4);
            }
            const int nAe = p.nA / kMmaWarps;
            const uint32_t a_step = (uint32_t)p.a_stage_bytes >> 4;
            const uint32_t a_lo0 = desc_lo(smem_u32(a_ring + patch_off)) + (uint32_t)(mw * nAe) * a_step;
            constexpr uint32_t kPairUnits = (uint32_t)kPairBytes >> 4;
            const uint32_t b_lo0 = desc_lo(smem_u32(b_ring));
            if (p.w_resident) {
                mbar_wait(w_full, 0);
                tc_fence_after();
            }
            const int my_tiles = ((int)p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
            const int npairs = (my_tiles + kMmaWarps - 1) / kMmaWarps;
            for (int P = 0; P < npairs; ++P) {
                const bool has = (int)(blockIdx.x + (P * kMmaWarps + mw) * gridDim.x) < p.tiles;
                const int acc = mw * kAccPer + P % kAccPer;               // this issuer's P-th tile goes to its accumulator P % kAccPer
                const uint32_t tmem_d = tmem_base + acc * 2 * BN;
                if (has) {
                    mbar_wait(acc_empty + acc, ((P / kAccPer) & 1) ^ 1);
                    tc_fence_after();
                    CONV_T(0, P * kMmaWarps + mw);
                }
                for (int cb = 0; cb < p.CB; ++cb) {
                    const int k = P * p.CB + cb;
                    const int sl = k % nAe;                             // stage inside this issuer's half of the patch ring
                    if (has) {
                        mbar_wait(a_full + mw * nAe + sl, (k / nAe) & 1);
                        tc_fence_after();
                        if (cb == 0) CONV_T(5, P * kMmaWarps + mw);
                    }
                    const uint32_t a_lo = a_lo0 + sl * a_step;
#pragma unroll
                    for (int pr = 0; pr < 5; ++pr) {                    // weight stage = taps 2 pr, 2 pr + 1
                        uint32_t b_lo;
                        int sb = 0;
                        if (p.w_resident) {
                            b_lo = b_lo0 + (uint32_t)(cb * 5 + pr) * kPairUnits;
                        } else {
                            const int bi = k * 5 + pr;
                            sb = bi % p.nB;
                            mbar_wait(b_full + sb, (bi / p.nB) & 1);
                            tc_fence_after();
                            b_lo = b_lo0 + (uint32_t)sb * kPairUnits;
                        }
                        if (has) {
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const int tap = 2 * pr + h;
                                if (tap < 9) {
                                    // patch rows are [hi x32 | lo x32] (hi at +0 / +2 descriptor units for the two K = 16 steps, lo at
                                    // +4 / +6); weight rows are [tap 2 pr x32 | tap 2 pr + 1 x32]
                                    const uint32_t al = a_lo + tap_off[tap], bl = b_lo + 4 * h;
                                    const uint32_t fresh = (cb == 0 && tap == 0) ? 1u : 0u;
#pragma unroll
                                    for (int kk = 0; kk < 2; ++kk) {
                                        mma_bf16_w(tmem_d, al + 2 * kk, kDescHi, bl + 2 * kk, kDescHi, idesc2, (fresh && kk == 0) ? 0u : 1u);
                                        if (!(p.dbg & 1)) mma_bf16_w(tmem_d, al + 4 + 2 * kk, kDescHi, bl + 2 * kk, kDescHi, idesc1, 1u);
                                    }
                                }
                            }
                        }
                        if (!p.w_resident) {
                            // release this issuer's share of the weight stage (an issuer without a tile in the last pair just arrives)
                            if (has) tc_commit(b_empty + sb);
                            else mbar_arrive(b_empty + sb);
                        }
                    }
                    if (has) tc_commit(a_empty + mw * nAe + sl);
                }
                if (has) {
                    tc_commit(acc_full + acc);
                    CONV_T(1, P * kMmaWarps + mw);
                }
            }
        }
    } else {
        // ===== four epilogue warps: TMEM lane quarter q = warp & 3, thread = tile row =====
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int img = r / p.img_rows, rem = r - img * p.img_rows;
        const int ri = rem / p.Wp, rj = rem - ri * p.Wp;
        const int et = threadIdx.x - 32 * kEpiWarp0;                      // 0..127 in warp order
        const int CBo = BN / 32;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
            const int ew = it % kMmaWarps, ej = it / kMmaWarps;           // issuer and that issuer's tile count
            const int buf = ew * kAccPer + ej % kAccPer;
            const int ng = tile / p.row_blocks, rb = tile - ng * p.row_blocks;
            const int n0 = ng * p.G, y0 = rb * p.R;
            const int n = n0 + img, y = y0 + ri, x = rj;
            const bool valid = img < p.G && n < p.N && ri < p.R && y < p.H && x < p.W;
            mbar_wait(acc_full + buf, (ej / kAccPer) & 1);
            tc_fence_after();
            if (threadIdx.x == 32 * kEpiWarp0) CONV_T(2, it);
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 2 * BN;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 16) {
                uint32_t v[16], v2[16];
                tmem_ld16(taddr + c0, v);
                tmem_ld16(taddr + BN + c0, v2);
                float o[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float t = (__uint_as_float(v[j]) + __uint_as_float(v2[j])) + (p.bias ? __ldg(p.bias + c0 + j) : 0.0f);
                    o[j] = p.relu ? fmaxf(t, 0.0f) : t;
                }
                if (p.dbg & 2) {
                } else if (p.out_mode == 2) {
                    // global average pool in the epilogue: per-warp sums (and counts of positive units) of the 16 channels over the
                    // warp's 32 tile rows -> shared memory; folded per image below
                    float sv[16], cv[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        sv[j] = valid ? o[j] : 0.0f;
                        cv[j] = (valid && o[j] > 0.0f) ? 1.0f : 0.0f;
                    }
                    const float ts = warp_sum16(sv, lane), tc = warp_sum16(cv, lane);
                    if (!(lane & 1)) {
                        stage[(q * 2 + 0) * BN + c0 + warp_sum16_index(lane)] = ts;
                        stage[(q * 2 + 1) * BN + c0 + warp_sum16_index(lane)] = tc;
                    }
                } else if (p.pool) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) *(float4*)(stage + r * kPitch + c0 + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
                } else if (valid) {
                    if (p.out_mode == 0) {
                        uint16_t* dst = p.out_planes + ((((size_t)n * p.H + y) * p.W + x) * CBo + (c0 >> 5)) * 64 + (c0 & 31);
                        uint4 h0, l0, h1, l1;
                        split8(o, &h0, &l0);
                        split8(o + 8, &h1, &l1);
                        *(uint4*)dst = h0;
                        *(uint4*)(dst + 8) = h1;
                        *(uint4*)(dst + 32) = l0;
                        *(uint4*)(dst + 40) = l1;
                    } else {
                        float* dst = p.out_nchw + (((size_t)n * BN + c0) * p.H + y) * p.W + x;
                        const size_t cs = (size_t)p.H * p.W;
#pragma unroll
                        for (int j = 0; j < 16; ++j) dst[j * cs] = p.accumulate ? dst[j * cs] + o[j] : o[j];
                    }
                }
            }
            // the accumulator buffer is free as soon as it has been read
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + buf);
            if (threadIdx.x == 32 * kEpiWarp0) CONV_T(3, it);
            if (p.out_mode == 2 && !(p.dbg & 2)) {
                epi_bar_sync();
                // fold the warps of each image in warp order (fixed: deterministic), one channel per thread
                for (int c = et; c < BN; c += 128) {
                    for (int g = 0; g < p.G; ++g) {
                        float ssum = 0.0f, scnt = 0.0f;
#pragma unroll
                        for (int w = 0; w < 4; ++w)
                            if (p.warp_img[w] == g) { ssum += stage[(w * 2 + 0) * BN + c]; scnt += stage[(w * 2 + 1) * BN + c]; }
                        const int gn = n0 + g;
                        if (gn < p.N) {
                            p.out_mean[(size_t)gn * BN + c] = ssum / (float)(p.H * p.W);
                            if (p.out_cnt) p.out_cnt[(size_t)gn * BN + c] = scnt;
                        }
                    }
                }
                epi_bar_sync();
            }
            if (p.pool && !(p.dbg & 2)) {
                epi_bar_sync();
                // 32 pooled slots per tile (slot = lane), BN/4 channels per thread (channel quarter = warp).  The tile's rows
                // are Wp-wide image rows stacked over the G images (Hp rows each, Hp even whenever G > 1, and R even): stacked rows
                // (2k, 2k+1) are the two source rows of pooled row k.
                const int Wh = p.Wp >> 1;
                const int pr = lane / Wh, pc = lane - pr * Wh;
                const int trow = 2 * pr;
                const int timg = trow / p.Hp, tri = trow - timg * p.Hp;
                const int src = trow * p.Wp + 2 * pc;
                const int pn = n0 + timg, py = (y0 + tri) >> 1, px = pc;
                const bool pvalid = timg < p.G && pn < p.N && (y0 + tri + 1) < p.H && (2 * pc + 1) < p.W;
                constexpr int kCh = BN / 4;
                const int cq = (et >> 5) * kCh;
                if (pvalid) {
                    float m[kCh];
#pragma unroll
                    for (int j = 0; j < kCh; j += 4) {
                        const float4 a = *(const float4*)(stage + src * kPitch + cq + j);
                        const float4 b = *(const float4*)(stage + (src + 1) * kPitch + cq + j);
                        const float4 c = *(const float4*)(stage + (src + p.Wp) * kPitch + cq + j);
                        const float4 d = *(const float4*)(stage + (src + p.Wp + 1) * kPitch + cq + j);
                        m[j] = fmaxf(fmaxf(a.x, b.x), fmaxf(c.x, d.x));
                        m[j + 1] = fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y));
                        m[j + 2] = fmaxf(fmaxf(a.z, b.z), fmaxf(c.z, d.z));
                        m[j + 3] = fmaxf(fmaxf(a.w, b.w), fmaxf(c.w, d.w));
                    }
                    if (p.out_mode == 0) {
                        uint16_t* dst = p.out_planes + ((((size_t)pn * p.Ho + py) * p.Wo + px) * CBo + (cq >> 5)) * 64 + (cq & 31);
#pragma unroll
                        for (int j = 0; j < kCh; j += 8) {
                            uint4 h, l;
                            split8(m + j, &h, &l);
                            *(uint4*)(dst + j) = h;
                            *(uint4*)(dst + 32 + j) = l;
                        }
                    } else {
                        float* dst = p.out_nchw + (((size_t)pn * BN + cq) * p.Ho + py) * p.Wo + px;
                        const size_t cs = (size_t)p.Ho * p.Wo;
#pragma unroll
                        for (int j = 0; j < kCh; ++j) dst[j * cs] = m[j];
                    }
                }
                epi_bar_sync();
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

// ---- weight gradient of a 3x3 / s1 / p1 convolution as an implicit GEMM (full adjoint; north_star's X^T . dY on the conv path) ----
//   dW[ci*9 + tap, co] = sum over output pixels r of  X[r shifted by tap, ci] * dZ[r, co]        (src/ops.rs:280-291 on the im2col matrix)
// The contraction runs over PIXELS, and both operands already sit in shared memory pixel-major: a row of an activation patch is
// [hi x32 | lo x32] of a 32-channel block — as an MN-major UMMA operand that is 64 "M" values per K index, and the same for the
// gradient tile as the N operand.  One MMA of M = 128 (two taps side by side: the second 64-value chunk is another tap's
// patch, LBO = the distance between the two starts), N = 64, K = 16 pixels therefore produces, for two taps at once, all four
// hi / lo blocks of the bf16x3 product: [X_hi ; X_lo]^T . [dZ_hi | dZ_lo]; the epilogue adds hi*hi + hi*lo + lo*hi.
// Patches come in the three x-shifted copies (x = -1, 0, +1) so that every tap offset is a whole number of swizzle atoms
// (MN-major operands with atom-aligned starts are the layout gemm_bx3.cu's T,N GEMMs already use).  A CTA owns one (ci block,
// co block) pair and a slice of the tiles, accumulates all of them in TMEM (5 tap pairs x 64 columns) and writes one partial
// [9][32][32]; conv_dw_fold_kernel sums the partials in CTA order (deterministic) into dW.
struct DwP {
    int N, H, W, CBx, CBz;            // X: [N, H, W, 32*CBx] planes, dZ: [N, H, W, 32*CBz] planes
    int Wp, R, Hp, G, row_blocks, tiles;
    int patch_bytes, patch_alloc;     // one X patch: delivered / reserved bytes
    int z_bytes;                      // dZ tile box bytes
    int stage_bytes;                  // 3 X patches + 1 dZ tile
    int slices;                       // CTAs per (ci block, co block) pair
    float* partial;                   // [pair][slice][9][32][32]
};

__global__ void __launch_bounds__(192, 1)
conv_dw_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_z, const __grid_constant__ DwP p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int kStages = 2;
    uint64_t* full = (uint64_t*)(smem + kStages * p.stage_bytes);
    uint64_t* empty = full + kStages;
    uint64_t* done = empty + kStages;
    uint32_t* tmem_slot = (uint32_t*)(done + 1);
    constexpr uint32_t kTmemCols = 512;                       // 5 tap pairs x 64 columns
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = blockIdx.y, cbx = pair / p.CBz, cbz = pair - cbx * p.CBz;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_z);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // rows of a stage that no TMA box covers are still read by the 128-row MMAs (they meet zero gradient rows or belong to pad
    // positions): they must hold finite values
    for (int i = threadIdx.x; i < kStages * p.stage_bytes / 16; i += 192) ((uint4*)smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();
    pdl_launch_dependents();
    const int my_tiles = ((int)p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    if (warp == 0) {
        if (lane == 0) {
            int i = 0;
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++i) {
                const int ng = tile / p.row_blocks, rb = tile - ng * p.row_blocks;
                const int n0 = ng * p.G, y0 = rb * p.R;
                const int s = i % kStages;
                mbar_wait(empty + s, ((i / kStages) & 1) ^ 1);
                uint8_t* st = smem + s * p.stage_bytes;
                mbar_expect_tx(full + s, 3 * p.patch_bytes + p.z_bytes);
                for (int kc = 0; kc < 3; ++kc) tma_load_5d(st + kc * p.patch_alloc, &map_x, full + s, 0, cbx, kc - 1, y0 - 1, n0);
                tma_load_5d(st + 3 * p.patch_alloc, &map_z, full + s, 0, cbz, 0, y0, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D = F32, A / B = BF16, both MN-major (bits 15, 16), N = 64, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            // tap pairs (first chunk, second chunk) as (kc, kr): horizontal neighbours share a row offset, the x = +1 column pairs up
            // vertically; the ninth tap is paired with itself (its second half is ignored)
            const int pa_kc[5] = {0, 0, 0, 2, 2}, pa_kr[5] = {0, 1, 2, 0, 2};
            const int pb_kc[5] = {1, 1, 1, 2, 2}, pb_kr[5] = {0, 1, 2, 1, 2};
            uint32_t a_off[5], a_lbo[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const int oa = pa_kc[q] * p.patch_alloc + pa_kr[q] * p.Wp * 128, ob = pb_kc[q] * p.patch_alloc + pb_kr[q] * p.Wp * 128;
                a_off[q] = (uint32_t)oa >> 4;
                a_lbo[q] = (uint32_t)((ob > oa ? ob - oa : 1024) >> 4) << 16;
            }
            int i = 0;
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++i) {
                const int s = i % kStages;
                mbar_wait(full + s, (i / kStages) & 1);
                tc_fence_after();
                const uint32_t x_lo = smem_u32(smem + s * p.stage_bytes) >> 4;
                const uint32_t z_lo = ((smem_u32(smem + s * p.stage_bytes + 3 * p.patch_alloc)) >> 4) | 0x10000u;
#pragma unroll 1
                for (int ks = 0; ks < 8; ++ks) {               // 16 pixels (tile rows) per step: 2048 bytes further in both operands
#pragma unroll
                    for (int q = 0; q < 5; ++q)
                        mma_bf16_w(tmem_base + q * 64, (x_lo + a_off[q] + ks * 128) | a_lbo[q], kDescHi, z_lo + ks * 128, kDescHi, idesc,
                                   (i == 0 && ks == 0) ? 0u : 1u);
                }
                tc_commit(empty + s);
            }
            tc_commit(done);
        }
    }
    // ===== epilogue (warps 2-5, once): fold hi / lo blocks of the five accumulators into the partial [9][32][32] =====
    if (warp >= 2) {
        const int q4 = warp & 3, row = q4 * 32 + lane;        // TMEM lane = M index: [tap A: hi 0-31, lo 32-63 | tap B: hi 64-95, lo 96-127]
        float* stg = (float*)smem;                              // the pipeline stages are idle by now: [128 rows][33] floats
        if (my_tiles > 0) {
            mbar_wait(done, 0);
            tc_fence_after();
        }
        float* out = p.partial + ((size_t)pair * gridDim.x + blockIdx.x) * (9 * 32 * 32);
        const int tap_of[5][2] = {{0, 1}, {3, 4}, {6, 7}, {2, 5}, {8, -1}};    // tap = kr*3 + kc of (chunk A, chunk B) per pair
        for (int q = 0; q < 5; ++q) {
            float v[32];
            if (my_tiles > 0) {
                uint32_t u[16];
                const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + q * 64;
                float a[64];
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 16) {
                    tmem_ld16(taddr + c0, u);
#pragma unroll
                    for (int j = 0; j < 16; ++j) a[c0 + j] = __uint_as_float(u[j]);
                }
                // hi rows (q4 even) contribute hi*hi + hi*lo, lo rows (q4 odd) lo*hi
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = (q4 & 1) ? a[j] : a[j] + a[32 + j];
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0.0f;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) stg[row * 33 + j] = v[j];
            asm volatile("bar.sync 1, 128;" ::: "memory");
            // thread t: channel ci = t % 32 of chunk t / 64 ... two passes over (chunk, ci) x 32 co
            const int t = threadIdx.x - 64;
            for (int e = t; e < 2 * 32 * 32; e += 128) {
                const int chunk = e >> 10, ci = (e >> 5) & 31, co = e & 31;
                const int tap = tap_of[q][chunk];
                if (tap < 0) continue;
                out[(tap * 32 + ci) * 32 + co] = stg[(chunk * 64 + ci) * 33 + co] + stg[(chunk * 64 + 32 + ci) * 33 + co];
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

// dW[(cbx*32 + ci)*9 + tap][cbz*32 + co] (+)= sum over slices of partial[pair][slice][tap][ci][co], slices in order
__global__ void __launch_bounds__(256)
conv_dw_fold_kernel(const float* __restrict__ partial, float* __restrict__ dw, int CBx, int CBz, int slices, int Cout, int accumulate) {
    pdl_wait();
    pdl_launch_dependents();
    const int total = CBx * CBz * 9 * 32 * 32;
    for (int e = blockIdx.x * 256 + threadIdx.x; e < total; e += gridDim.x * 256) {
        const int co = e & 31, ci = (e >> 5) & 31;
        const int t = e >> 10, tap = t % 9, pair = t / 9;
        const int cbx = pair / CBz, cbz = pair - cbx * CBz;
        const float* src = partial + (size_t)pair * slices * 9216 + (tap * 32 + ci) * 32 + co;
        float s = 0.0f;
        for (int k = 0; k < slices; ++k) s += src[(size_t)k * 9216];
        float* d = dw + (size_t)((cbx * 32 + ci) * 9 + tap) * Cout + cbz * 32 + co;
        *d = accumulate ? *d + s : s;
    }
}

// ---- weights [K = ci*9 + tap, C_out] fp32 -> the kernel's B operand:
//   [ci/32][tap pair p = tap/2 (5 pairs, the tenth tap is zero)][row r < 2*C_out][64 bf16 = tap 2p: 32 channels | tap 2p+1: 32 channels]
//   rows r < C_out hold hi = rn_bf16(w) of output channel r, rows r >= C_out hold lo = rn_bf16(w - hi) of channel r - C_out
struct WPrep {
    const float* w2[kMaxPrep];
    uint16_t* dst[kMaxPrep];
    int cin[kMaxPrep], cout[kMaxPrep];
    int adjoint[kMaxPrep];           // 1: the weights of the input-gradient convolution, W'[co*9 + (8 - tap), ci] = W[ci*9 + tap, co]
    int count;
};
__global__ void __launch_bounds__(256)
conv_w_planes_kernel(const __grid_constant__ WPrep wp) {
    pdl_wait();
    pdl_launch_dependents();
    const int l = blockIdx.y;
    if (l >= wp.count) return;
    const int cin = wp.cin[l], cout = wp.cout[l];
    const int total = cin * 10 * cout;                          // tap 9 of every channel: the zero half of the fifth pair
    for (int e = blockIdx.x * 256 + threadIdx.x; e < total; e += gridDim.x * 256) {
        const int co = e % cout, k = e / cout;                  // k = ci*10 + tap: coalesced reads along co
        const int ci = k / 10, tap = k - ci * 10;
        // adjoint: this layer's (ci, co) are the forward layer's (co, ci) and the taps are mirrored (dX = conv(dY, flipped W^T))
        const float v = tap >= 9 ? 0.0f
                        : wp.adjoint[l] ? __ldg(wp.w2[l] + (size_t)(co * 9 + (8 - tap)) * cin + ci)
                                        : __ldg(wp.w2[l] + (size_t)(ci * 9 + tap) * cout + co);
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(h));
        uint16_t* d = wp.dst[l] + ((size_t)((ci >> 5) * 5 + (tap >> 1)) * 2 * cout + co) * 64 + (tap & 1) * 32 + (ci & 31);
        d[0] = __bfloat16_as_ushort(h);
        d[(size_t)cout * 64] = __bfloat16_as_ushort(lo);
    }
}

// ---- NCHW fp32 -> planes (first layer of a stack whose input already has >= 32 channels; the single-layer eager op) --------
__global__ void __launch_bounds__(256)
nchw_to_planes_kernel(const float* __restrict__ x, const float* __restrict__ mask, uint16_t* __restrict__ out, int N, int C, int H, int W) {
    pdl_wait();
    pdl_launch_dependents();
    const unsigned int CG = C / 8;
    const unsigned int hw = (unsigned int)(H * W), total = (unsigned int)N * hw * CG;      // < 2^31 (checked on the host)
    for (unsigned int e = blockIdx.x * 256 + threadIdx.x; e < total; e += gridDim.x * 256) {
        // pixel fastest: a warp reads 32 consecutive x of one channel (coalesced), 8 channels per thread; 32-bit index math
        const unsigned int t = e / hw, pix = e - t * hw;
        const unsigned int n = t / CG, cg = t - n * CG;
        const float* src = x + ((size_t)n * C + cg * 8) * hw + pix;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldg(src + j * hw);
        if (mask) {                                          // ReLU backward on the way in: g * [y > 0]  (src/ops.rs:358-370)
            const float* ms = mask + (src - x);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __ldg(ms + j * hw) > 0.0f ? v[j] : 0.0f;
        }
        uint4 h, l;
        split8(v, &h, &l);
        uint16_t* dst = out + (((size_t)n * hw + pix) * (C / 32) + (cg >> 2)) * 64 + (cg & 3) * 8;
        *(uint4*)dst = h;
        *(uint4*)(dst + 32) = l;
    }
}

// ---- first layer with a tiny contraction (C_in * 9 <= 36: the C_in = 1 image layer): direct fp32 on the CUDA cores,
// bias + ReLU (+ 2x2 max-pool) fused, output written straight as planes.  Exact fp32, k order (ci, kr, kc) ascending.
// Replaces im2col + sgemm + transpose_4d + add_bias_4d + relu (+ max_pool2d) of src/tensor.rs:1221-1285, 1379-1464.
template <bool POOL, bool C1>
__global__ void __launch_bounds__(256)
conv_first_planes_kernel(const float* __restrict__ x, const float* __restrict__ w2, const float* __restrict__ bias,
                         uint16_t* __restrict__ out, int N, int Cin, int H, int W, int Cout, int relu) {
    constexpr bool pool = POOL;
    extern __shared__ __align__(16) float sw[];                // [K][Cout] then [Cout]
    const int K = Cin * 9;
    float* sb = sw + K * Cout;
    pdl_wait();
    pdl_launch_dependents();
    for (int i = threadIdx.x; i < K * Cout; i += 256) sw[i] = __ldg(w2 + i);
    for (int i = threadIdx.x; i < Cout; i += 256) sb[i] = bias ? __ldg(bias + i) : 0.0f;
    __syncthreads();
    const int Ho = pool ? H / 2 : H, Wo = pool ? W / 2 : W;
    const unsigned int CG = Cout / 8;
    constexpr int nsub = POOL ? 4 : 1, win = POOL ? 4 : 3;
    const unsigned int total = (unsigned int)N * Ho * Wo * CG;                               // < 2^31 (checked on the host)
    // C1 (one input channel, the image layer): the grid stride is a multiple of the channel-group count, so a thread keeps its
    // channel group for all its pixels and holds its 9 x 8 weights in registers — the per-tap shared-memory loads (4 LSU
    // wavefronts per 128-bit load, 18 loads per pixel) were what bound this kernel
    float4 wreg[C1 ? 9 : 1][2];
    if (C1) {
        const unsigned int cg0 = (blockIdx.x * 256 + threadIdx.x) % CG;
#pragma unroll
        for (int tap = 0; tap < (C1 ? 9 : 1); ++tap) {
            wreg[tap][0] = *(const float4*)(sw + tap * Cout + cg0 * 8);
            wreg[tap][1] = *(const float4*)(sw + tap * Cout + cg0 * 8 + 4);
        }
    }
    for (unsigned int e = blockIdx.x * 256 + threadIdx.x; e < total; e += gridDim.x * 256) {
        // 32-bit index arithmetic (64-bit divisions cost more than the nine taps)
        const unsigned int pix = e / CG, cg = e - pix * CG;
        const unsigned int rowi = pix / (unsigned int)Wo;
        const int xo = (int)(pix - rowi * (unsigned int)Wo);
        const unsigned int n = rowi / (unsigned int)Ho;
        const int yo = (int)(rowi - n * (unsigned int)Ho);
        const int y0 = (pool ? 2 * yo : yo) - 1, x0 = (pool ? 2 * xo : xo) - 1;
        float acc[nsub][8];
#pragma unroll
        for (int s = 0; s < nsub; ++s)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[s][j] = 0.0f;
        for (int ci = 0; ci < Cin; ++ci) {
            // the whole input window first (3x3, or 4x4 shared by the four pooled positions): every load is in flight before the
            // first multiply, instead of one round trip per tap
            const float* xp = x + ((size_t)n * Cin + ci) * (size_t)H * W;
            float in[win * win];
            if (y0 >= 0 && x0 >= 0 && y0 + win <= H && x0 + win <= W) {            // interior window: no bounds checks
                const float* wp0 = xp + y0 * W + x0;
#pragma unroll
                for (int r = 0; r < win; ++r)
#pragma unroll
                    for (int c = 0; c < win; ++c) in[r * win + c] = __ldg(wp0 + r * W + c);
            } else {
#pragma unroll
                for (int r = 0; r < win; ++r)
#pragma unroll
                    for (int c = 0; c < win; ++c) {
                        const int iy = y0 + r, ix = x0 + c;
                        in[r * win + c] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(xp + iy * W + ix) : 0.0f;
                    }
            }
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const float4 wa = C1 ? wreg[C1 ? tap : 0][0] : *(const float4*)(sw + (ci * 9 + tap) * Cout + cg * 8);
                const float4 wb = C1 ? wreg[C1 ? tap : 0][1] : *(const float4*)(sw + (ci * 9 + tap) * Cout + cg * 8 + 4);
#pragma unroll
                for (int s = 0; s < nsub; ++s) {
                    {
                        const float xv = in[((s >> 1) + tap / 3) * win + (s & 1) + tap % 3];
                        acc[s][0] = fmaf(xv, wa.x, acc[s][0]); acc[s][1] = fmaf(xv, wa.y, acc[s][1]);
                        acc[s][2] = fmaf(xv, wa.z, acc[s][2]); acc[s][3] = fmaf(xv, wa.w, acc[s][3]);
                        acc[s][4] = fmaf(xv, wb.x, acc[s][4]); acc[s][5] = fmaf(xv, wb.y, acc[s][5]);
                        acc[s][6] = fmaf(xv, wb.z, acc[s][6]); acc[s][7] = fmaf(xv, wb.w, acc[s][7]);
                    }
                }
            }
        }
        float best[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float bj = sb[cg * 8 + j];
            float v = acc[0][j] + bj;
            if (relu) v = fmaxf(v, 0.0f);
            best[j] = v;
#pragma unroll
            for (int s = 1; s < nsub; ++s) {
                {
                    float u = acc[s][j] + bj;
                    if (relu) u = fmaxf(u, 0.0f);
                    best[j] = fmaxf(best[j], u);
                }
            }
        }
        uint4 h, l;
        split8(best, &h, &l);
        uint16_t* dst = out + ((size_t)pix * (Cout / 32) + (cg >> 2)) * 64 + (cg & 3) * 8;
        *(uint4*)dst = h;
        *(uint4*)(dst + 32) = l;
    }
}

// ---- the image layer (C_in = 1, no pool) again, one thread per (image row, 8-channel group) walking along x -----------------
// Same arithmetic as conv_first_planes_kernel (fp32, taps ascending, bias last, ReLU, bf16 hi/lo planes) — bit-identical output —
// but the per-output overhead is gone: the 9 x 8 weights of the thread's channel group sit in registers as packed pairs and are
// multiplied with fma.rn.f32x2 (Blackwell FFMA2: two IEEE FMAs per issue slot), the 3 x 3 input window slides (three loads per
// pixel, the next column requested one pixel ahead), and the index arithmetic happens once per row.  ~100 instructions per 8
// outputs instead of ~265: the kernel goes from issue-bound (54 us for the 102 MB of the first activation at batch 1024)
// towards the HBM write time.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2f(float lo, float hi) {
    f32x2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2f(f32x2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void fma2f(f32x2_t& d, f32x2_t a, f32x2_t b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }

__global__ void __launch_bounds__(256)
conv_first_rows_kernel(const float* __restrict__ x, const float* __restrict__ w2, const float* __restrict__ bias,
                       uint16_t* __restrict__ out, int N, int H, int W, int Cout, int relu) {
    pdl_wait();
    pdl_launch_dependents();
    const unsigned int CG = Cout / 8;                          // 256 % CG == 0 (host): a thread's channel group is fixed
    const unsigned int cg = threadIdx.x % CG;
    f32x2_t wp[9][4];
    float bj[8];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        const float4 wa = __ldg((const float4*)(w2 + tap * Cout + cg * 8)), wb = __ldg((const float4*)(w2 + tap * Cout + cg * 8 + 4));
        wp[tap][0] = pack2f(wa.x, wa.y); wp[tap][1] = pack2f(wa.z, wa.w);
        wp[tap][2] = pack2f(wb.x, wb.y); wp[tap][3] = pack2f(wb.z, wb.w);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) bj[j] = bias ? __ldg(bias + cg * 8 + j) : 0.0f;
    const unsigned int total = (unsigned int)N * H * CG;       // < 2^31 (checked on the host)
    const unsigned int CBo = Cout / 32;
    for (unsigned int e = blockIdx.x * 256 + threadIdx.x; e < total; e += gridDim.x * 256) {
        const unsigned int rowi = e / CG;                      // (n, y)
        const unsigned int n = rowi / (unsigned int)H;
        const int y = (int)(rowi - n * (unsigned int)H);
        const float* r1 = x + (size_t)rowi * W;
        const bool up = y > 0, dn = y + 1 < H;
        const float* r0 = r1 - W;
        const float* r2 = r1 + W;
        // window columns: c?0 = x - 1, c?1 = x, c?2 = x + 1; nx? = column x + 2 (in flight)
        float c00 = 0.f, c10 = 0.f, c20 = 0.f;
        float c01 = up ? __ldg(r0) : 0.f, c11 = __ldg(r1), c21 = dn ? __ldg(r2) : 0.f;
        float c02 = 0.f, c12 = 0.f, c22 = 0.f;
        if (W > 1) { c02 = up ? __ldg(r0 + 1) : 0.f; c12 = __ldg(r1 + 1); c22 = dn ? __ldg(r2 + 1) : 0.f; }
        uint16_t* dst = out + ((size_t)rowi * W * CBo + (cg >> 2)) * 64 + (cg & 3) * 8;
        for (int xo = 0; xo < W; ++xo) {
            float n0 = 0.f, n1 = 0.f, n2 = 0.f;
            if (xo + 2 < W) { n0 = up ? __ldg(r0 + xo + 2) : 0.f; n1 = __ldg(r1 + xo + 2); n2 = dn ? __ldg(r2 + xo + 2) : 0.f; }
            f32x2_t acc[4] = {0ull, 0ull, 0ull, 0ull};
            const float win[9] = {c00, c01, c02, c10, c11, c12, c20, c21, c22};
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const f32x2_t xv = pack2f(win[tap], win[tap]);
#pragma unroll
                for (int j = 0; j < 4; ++j) fma2f(acc[j], xv, wp[tap][j]);
            }
            float v[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) unpack2f(acc[j], v[2 * j], v[2 * j + 1]);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                v[j] += bj[j];
                if (relu) v[j] = fmaxf(v[j], 0.0f);
            }
            uint4 h, l;
            split8(v, &h, &l);
            *(uint4*)dst = h;
            *(uint4*)(dst + 32) = l;
            dst += (size_t)CBo * 64;
            c00 = c01; c10 = c11; c20 = c21;
            c01 = c02; c11 = c12; c21 = c22;
            c02 = n0; c12 = n1; c22 = n2;
        }
    }
}

// ---- global average pool of the stack's output fused with the count of positive units per plane ------------------------------
// mean[n, c] = sum_p y[n, c, p] / hw  (AdaptiveAvgPool2d::global -> avg_pool2d, src/nn.rs:670-686, src/tensor.rs:1524-1590);
// cnt[n, c] = #{p : y[n, c, p] > 0} is everything the backward of the pool + ReLU + bias chain needs (see gap_bias_grad below).
// One warp per (n, c) plane, two planes in flight per warp.
__global__ void __launch_bounds__(256)
gap_count_kernel(const float* __restrict__ y, float* __restrict__ mean, float* __restrict__ cnt, int planes, int hw) {
    pdl_wait();
    pdl_launch_dependents();
    const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * 256) >> 5;
    const float inv = 1.0f;
    (void)inv;
    for (int p0 = warp * 2; p0 < planes; p0 += nwarps * 2) {
        float s[2] = {0.0f, 0.0f}, c[2] = {0.0f, 0.0f};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int p = p0 + u;
            if (p < planes)
                for (int i = lane; i < hw; i += 32) {
                    const float v = __ldg(y + (size_t)p * hw + i);
                    s[u] += v;
                    c[u] += v > 0.0f ? 1.0f : 0.0f;
                }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
                c[u] += __shfl_xor_sync(0xffffffffu, c[u], o);
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int u = 0; u < 2; ++u)
                if (p0 + u < planes) {
                    mean[p0 + u] = s[u] / (float)hw;
                    if (cnt) cnt[p0 + u] = c[u];
                }
        }
    }
}

// gb[c] (+)= sum_n (g[n, c] / hw) * cnt[n, c]: the gradient of the last conv's bias through avg_pool2d backward (every unit
// of the plane receives g / hw, src/tensor.rs:1600-1655), the ReLU gate (src/ops.rs:358-370) and add_bias_4d's sum over
// n, h, w (src/tensor.rs:2003-2027), without materialising the [N, C, H, W] gradient.  cnt == NULL: no ReLU (every unit passes).
// Stage 1: grid (C/32, S) partial sums over an n-slice in a fixed order; stage 2 folds the S partials in order (deterministic).
__global__ void __launch_bounds__(256)
gap_bias_grad_stage1(const float* __restrict__ g, const float* __restrict__ cnt, float* __restrict__ part, int N, int C, int hw) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float red[8][33];
    const int cl = threadIdx.x & 31, nr = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    const int S = gridDim.y;
    const int n0 = (int)(((long long)blockIdx.y * N) / S), n1 = (int)(((long long)(blockIdx.y + 1) * N) / S);
    float acc = 0.0f;
    if (c < C)
        for (int n = n0 + nr; n < n1; n += 8) {
            const float t = __ldg(g + (size_t)n * C + c) / (float)hw;
            acc += cnt ? t * __ldg(cnt + (size_t)n * C + c) : t * (float)hw;
        }
    red[nr][cl] = acc;
    __syncthreads();
    if (nr == 0 && c < C) {
        float s = 0.0f;
#pragma unroll
        for (int r = 0; r < 8; ++r) s += red[r][cl];
        part[(size_t)blockIdx.y * C + c] = s;
    }
}
__global__ void __launch_bounds__(256)
gap_bias_grad_stage2(const float* __restrict__ part, float* __restrict__ gb, int S, int C, int accumulate) {
    pdl_wait();
    pdl_launch_dependents();
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= C) return;
    float s = 0.0f;
    for (int i = 0; i < S; ++i) s += part[(size_t)i * C + c];
    gb[c] = accumulate ? gb[c] + s : s;
}

// ---- host side -------------------------------------------------------------------------------------------------------------
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

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

constexpr int kSmemBudget = 227 * 1024 - 1024 /*alignment*/ - 512 /*barriers, TMEM slot*/;

struct TmpBuf {
    tp_buf* b = nullptr;
    ~TmpBuf() { if (b) tp_buf_release(b); }
};

// geometry + shared-memory plan of one layer; false: this shape does not go through the kernel
bool plan_layer(int N, int H, int W, int Cin, int Cout, bool pool, int out_mode, int shift_mode, ConvP* p, int* smem_bytes) {
    if (N <= 0 || H <= 0 || W <= 0 || Cin % 32 || Cin <= 0) return false;
    if (Cout != 32 && Cout != 64 && Cout != 128) return false;
    if (W >= 32) return false;
    if (pool && (H < 2 || W < 2)) return false;
    ConvP q{};
    q.N = N; q.H = H; q.W = W; q.CB = Cin / 32; q.Cout = Cout;
    q.Wp = W < 8 ? 8 : W < 16 ? 16 : 32;
    q.R = 128 / q.Wp;
    if (H <= q.R) {
        q.Hp = H + 2;
        q.G = 1 + (q.R - H) / (H + 2);
        q.row_blocks = 1;
        if (pool && q.G > 1 && (q.Hp & 1)) {
            // the pooled read-out pairs tile rows (2k, 2k+1) of the stacked images: every image must start on an even row
            q.Hp += 1;
            q.G = 1 + (q.R - H) / q.Hp;
        }
    } else {
        q.Hp = q.R + 2;
        q.G = 1;
        q.row_blocks = (H + q.R - 1) / q.R;
    }
    if (q.Hp > 256 || q.G > 256) return false;
    q.img_rows = q.Hp * q.Wp;
    q.patch_bytes = q.G * q.img_rows * 128;
    const int need_rows = 128 + 2 * q.Wp + 8;                         // what the nine shifted 128-row reads can touch
    const int rows = q.G * q.img_rows > need_rows ? q.G * q.img_rows : need_rows;
    q.patch_alloc = ((rows * 128 + 1023) / 1024) * 1024;
    q.tiles = ((N + q.G - 1) / q.G) * q.row_blocks;
    q.shift_mode = shift_mode;
    q.a_stage_bytes = shift_mode ? 1024 + q.patch_alloc : 3 * q.patch_alloc;
    const int stage_bytes = pool ? 128 * (Cout + 4) * 4 : out_mode == 2 ? 8 * Cout * 4 : 0;
    const int pair_bytes = 2 * Cout * 128;                              // one weight stage (two taps, hi and lo rows)
    const int w_bytes = 5 * q.CB * pair_bytes;
    int left = kSmemBudget - stage_bytes;
    // patch stages come in pairs (one half of the ring per MMA issuer)
    if (w_bytes <= 96 * 1024 && left - w_bytes >= 2 * q.a_stage_bytes) {
        q.w_resident = 1;
        q.b_stage_bytes = 0; q.nB = 0;
        int nA = (left - w_bytes) / q.a_stage_bytes;
        q.nA = nA >= 6 ? 6 : nA >= 4 ? 4 : 2;
    } else {
        q.w_resident = 0;
        q.b_stage_bytes = pair_bytes;
        q.nA = 4;
        if (left - q.nA * q.a_stage_bytes < 2 * q.b_stage_bytes) q.nA = 2;
        int nB = (left - q.nA * q.a_stage_bytes) / q.b_stage_bytes;
        if (nB < 2) return false;
        q.nB = nB > 8 ? 8 : nB;
    }
    q.out_mode = out_mode;
    if (out_mode == 2) {
        // the pooled epilogue folds whole images inside a tile: every image in one tile, every epilogue warp's rows in one image
        if (pool || q.row_blocks != 1) return false;
        for (int w = 0; w < 4; ++w) {
            q.warp_img[w] = -1;
            for (int r = 32 * w; r < 32 * w + 32; ++r) {
                const int img = r / q.img_rows, rem = r - img * q.img_rows;
                const int ri = rem / q.Wp, rj = rem - ri * q.Wp;
                if (img >= q.G || ri >= H || rj >= W) continue;
                if (q.warp_img[w] >= 0 && q.warp_img[w] != img) return false;
                q.warp_img[w] = img;
            }
        }
    }
    q.pool = pool ? 1 : 0;
    q.relu = 1;
    q.Ho = pool ? H / 2 : H;
    q.Wo = pool ? W / 2 : W;
    *p = q;
    *smem_bytes = q.nA * q.a_stage_bytes + (q.w_resident ? w_bytes : q.nB * q.b_stage_bytes) + stage_bytes + 1024 + 512;
    return true;
}

bool make_map_x(EncodeTiledFn enc, CUtensorMap* map, const uint16_t* ptr, const ConvP& p) {
    cuuint64_t gdim[5] = {64, (cuuint64_t)p.CB, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.N};
    cuuint64_t gstr[4] = {128, (cuuint64_t)p.CB * 128, (cuuint64_t)p.W * p.CB * 128, (cuuint64_t)p.H * p.W * p.CB * 128};
    cuuint32_t box[5] = {64, 1, (cuuint32_t)p.Wp, (cuuint32_t)p.Hp, (cuuint32_t)p.G};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool make_map_w(EncodeTiledFn enc, CUtensorMap* map, const uint16_t* ptr, int CB, int Cout) {
    cuuint64_t gdim[3] = {64, (cuuint64_t)2 * Cout, (cuuint64_t)5 * CB};
    cuuint64_t gstr[2] = {128, (cuuint64_t)2 * Cout * 128};
    cuuint32_t box[3] = {64, (cuuint32_t)(2 * Cout), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename... Args>
int launch_pdl(tp_ctx* ctx, void (*kern)(Args...), dim3 grid, dim3 block, size_t smem, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    TP_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
    TP_LAUNCH_OK(ctx);
    return TP_OK;
}

template <int BN>
int launch_conv(tp_ctx* ctx, const CUtensorMap& mx, const CUtensorMap& mw, const ConvP& p, int smem, bool pdl) {
    auto kern = conv3x3_bx3_kernel<BN>;
    // the opt-in is per device and monotone: keep the largest value ever requested
    static int attr_smem[16] = {};
    const int dev = ctx->device < 16 ? ctx->device : 15;
    if (ctx->device >= 16 || attr_smem[dev] < smem) {
        TP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_smem[dev] = smem;
    }
    const int grid = p.tiles < ctx->sm_count ? p.tiles : ctx->sm_count;
    return launch_pdl(ctx, kern, dim3(grid), dim3(kConvThreads), (size_t)smem, pdl, mx, mw, p);
}

int launch_conv_any(tp_ctx* ctx, const CUtensorMap& mx, const CUtensorMap& mw, const ConvP& p, int smem, bool pdl) {
    if (p.Cout == 32) return launch_conv<32>(ctx, mx, mw, p, smem, pdl);
    if (p.Cout == 64) return launch_conv<64>(ctx, mx, mw, p, smem, pdl);
    return launch_conv<128>(ctx, mx, mw, p, smem, pdl);
}

int default_shift_mode() {
    static const int m = env_int("TAPER_CONV_SHIFT", 2);
    return m;
}

thread_local int g_shift_override = -1;       // tests: force a shift mode for the calls of this thread
thread_local int g_dbg_flags = 0;

}  // namespace

namespace tp {

size_t conv_planes_floats(size_t n, size_t h, size_t w, size_t c) { return n * h * w * c; }     // 2 bf16 per element = 1 float

// One stack of 3x3 / s1 / p1 Conv(+bias)+ReLU layers, each optionally followed by a 2x2 / s2 max-pool, NCHW fp32 in and out.
// TP_ERR_UNSUPPORTED: some layer's shape does not go this way (nothing has been launched).
int conv_stack_fwd(tp_ctx* ctx, const float* x, int N, int C0, int H, int W, int n_layers, const float* const* w2,
                   const float* const* bias, const int* cout, const int* pool, const int* relu, float* y, float* gap_mean, float* gap_cnt) {
    if (n_layers < 1 || n_layers > kMaxPrep) return TP_ERR_UNSUPPORTED;
    EncodeTiledFn enc = get_encode();
    if (!enc) return TP_ERR_UNSUPPORTED;
    const int shift_mode = g_shift_override >= 0 ? g_shift_override : default_shift_mode();
    cudaSetDevice(ctx->device);
    // programmatic dependent launch, also while the step is being captured (the graph keeps the programmatic edges)
    static const bool pdl_in_capture = env_int("TAPER_CONV_PDL_CAPTURE", 1) != 0;
    const bool pdl = !ctx->capturing || pdl_in_capture;
    // ---- plan every layer first (no launches before the whole stack is known to fit) ----
    const bool first_direct = C0 * 9 <= 36;
    if ((size_t)N * H * W * (size_t)(C0 > 32 ? C0 : 32) >= ((size_t)1 << 31)) return TP_ERR_UNSUPPORTED;      // 32-bit index math in the layout kernels
    if (first_direct && (n_layers < 2 || cout[0] % 32 || cout[0] > 256)) return TP_ERR_UNSUPPORTED;
    if (!first_direct && C0 % 32) return TP_ERR_UNSUPPORTED;
    ConvP P[kMaxPrep];
    int smem[kMaxPrep];
    bool gap_in_epilogue = false;
    int h = H, w = W, c = C0;
    for (int l = 0; l < n_layers; ++l) {
        if (!(l == 0 && first_direct)) {
            const bool last = l == n_layers - 1;
            // a trailing global average pool is folded into the last layer's epilogue when its tiles hold whole images
            if (last && gap_mean && plan_layer(N, h, w, c, cout[l], pool[l] != 0, 2, shift_mode, &P[l], &smem[l])) {
                gap_in_epilogue = true;
            } else if (!plan_layer(N, h, w, c, cout[l], pool[l] != 0, last ? 1 : 0, shift_mode, &P[l], &smem[l])) {
                return TP_ERR_UNSUPPORTED;
            }
        }
        if (pool[l]) { h /= 2; w /= 2; }
        c = cout[l];
        if (h < 1 || w < 1) return TP_ERR_UNSUPPORTED;
    }
    // ---- workspaces: weight planes of the tensor-core layers, two ping-pong activation buffers ----
    int rc;
    size_t wtot = 0, woff[kMaxPrep];
    {
        int ci = C0;
        for (int l = 0; l < n_layers; ++l) {
            woff[l] = wtot;
            if (!(l == 0 && first_direct)) wtot += (size_t)ci * 10 * cout[l];          // bf16 pairs: 5 tap pairs x (hi + lo rows) x 64 per 32 channels
            ci = cout[l];
        }
    }
    TmpBuf wplanes, act[2];
    if ((rc = tp_buf_alloc(ctx, wtot ? wtot : 1, &wplanes.b))) return rc;
    size_t amax = first_direct ? 0 : (size_t)N * H * W * C0;
    {
        int hh = H, ww = W;
        for (int l = 0; l + 1 < n_layers; ++l) {
            if (pool[l]) { hh /= 2; ww /= 2; }
            const size_t a = (size_t)N * hh * ww * cout[l];
            if (a > amax) amax = a;
        }
    }
    if ((rc = tp_buf_alloc(ctx, amax ? amax : 1, &act[0].b))) return rc;
    if ((rc = tp_buf_alloc(ctx, amax ? amax : 1, &act[1].b))) return rc;
    // ---- weight planes of all layers in one launch ----
    WPrep wp{};
    {
        int ci = C0;
        for (int l = 0; l < n_layers; ++l) {
            if (!(l == 0 && first_direct)) {
                const int i = wp.count++;
                wp.w2[i] = w2[l];
                wp.dst[i] = (uint16_t*)(wplanes.b->ptr + woff[l]);
                wp.cin[i] = ci; wp.cout[i] = cout[l];
            }
            ci = cout[l];
        }
    }
    if (wp.count) {
        if ((rc = launch_pdl(ctx, conv_w_planes_kernel, dim3(32, wp.count), dim3(256), 0, false, wp))) return rc;
    }
    // ---- layers ----
    // a trailing global average pool that the last epilogue cannot absorb reads the NCHW output from a scratch buffer
    int hl = H, wl = W;
    for (int l = 0; l < n_layers; ++l) if (pool[l]) { hl /= 2; wl /= 2; }
    TmpBuf ytmp;
    float* y_last = y;
    if (gap_mean && !gap_in_epilogue) {
        if ((rc = tp_buf_alloc(ctx, (size_t)N * cout[n_layers - 1] * hl * wl, &ytmp.b))) return rc;
        y_last = ytmp.b->ptr;
    }
    int cur = 0;
    h = H; w = W; c = C0;
    int l0 = 0;
    if (first_direct) {
        const int ho = pool[0] ? h / 2 : h, wo = pool[0] ? w / 2 : w;
        const size_t items = (size_t)N * ho * wo * (cout[0] / 8);
        const size_t sm = (size_t)(C0 * 9 + 1) * cout[0] * sizeof(float);
        {
            // grid: a multiple of the channel-group count in threads (the C1 kernels keep a thread on one channel group)
            const dim3 grid(grid_for(ctx, items, 256, 8)), block(256);
            const bool c1 = C0 == 1 && (256 % (cout[0] / 8)) == 0;
            uint16_t* dst = (uint16_t*)act[cur].b->ptr;
            const int rl = relu[0] ? 1 : 0;
            if (c1 && !pool[0] && h * w >= 64 && !getenv("TAPER_CONV_FIRST_V1"))
                rc = launch_pdl(ctx, conv_first_rows_kernel, dim3(grid_for(ctx, (size_t)N * h * (cout[0] / 8), 256, 8)), block, 0, pdl, x, w2[0], bias[0], dst,
                                N, h, w, cout[0], rl);
            else if (pool[0])
                rc = c1 ? launch_pdl(ctx, conv_first_planes_kernel<true, true>, grid, block, sm, pdl, x, w2[0], bias[0], dst, N, C0, h, w, cout[0], rl)
                        : launch_pdl(ctx, conv_first_planes_kernel<true, false>, grid, block, sm, pdl, x, w2[0], bias[0], dst, N, C0, h, w, cout[0], rl);
            else
                rc = c1 ? launch_pdl(ctx, conv_first_planes_kernel<false, true>, grid, block, sm, pdl, x, w2[0], bias[0], dst, N, C0, h, w, cout[0], rl)
                        : launch_pdl(ctx, conv_first_planes_kernel<false, false>, grid, block, sm, pdl, x, w2[0], bias[0], dst, N, C0, h, w, cout[0], rl);
        }
        if (rc) return rc;
        h = ho; w = wo; c = cout[0];
        l0 = 1;
    } else {
        const size_t items = (size_t)N * h * w * (c / 8);
        if ((rc = launch_pdl(ctx, nchw_to_planes_kernel, dim3(grid_for(ctx, items, 256, 8)), dim3(256), 0, pdl, x, (const float*)nullptr, (uint16_t*)act[cur].b->ptr, N, c, h, w)))
            return rc;
    }
    for (int l = l0; l < n_layers; ++l) {
        ConvP& p = P[l];
        p.bias = bias[l];
        p.relu = relu[l] ? 1 : 0;
        p.accumulate = 0;
        p.dbg = g_dbg_flags;
        const bool last = l == n_layers - 1;
        p.out_planes = last ? nullptr : (uint16_t*)act[cur ^ 1].b->ptr;
        p.out_nchw = last ? y_last : nullptr;
        p.out_mean = gap_mean;
        p.out_cnt = gap_cnt;
        CUtensorMap mx, mw;
        if (!make_map_x(enc, &mx, (const uint16_t*)act[cur].b->ptr, p)) { set_error("conv_stack_fwd: cuTensorMapEncodeTiled (activations) failed"); return TP_ERR_CUDA; }
        if (!make_map_w(enc, &mw, (const uint16_t*)(wplanes.b->ptr + woff[l]), p.CB, p.Cout)) { set_error("conv_stack_fwd: cuTensorMapEncodeTiled (weights) failed"); return TP_ERR_CUDA; }
        if ((rc = launch_conv_any(ctx, mx, mw, p, smem[l], pdl))) return rc;
        cur ^= 1;
    }
    if (gap_mean && !gap_in_epilogue) {
        const size_t planes = (size_t)N * cout[n_layers - 1];
        if ((rc = launch_pdl(ctx, gap_count_kernel, dim3(grid_for(ctx, planes * 16, 256, 8)), dim3(256), 0, pdl, (const float*)y_last, gap_mean,
                             gap_cnt, (int)planes, hl * wl)))
            return rc;
    }
    return TP_OK;
}

// Input gradient of a 3x3 / s1 / p1 convolution as the SAME implicit GEMM (north_star's dY . W^T on the conv path; the reference
// drops this link, SURVEY A1 — full-adjoint mode only):  dX = conv(dZ, W') with dZ = gy * [y > 0] (ReLU backward, src/ops.rs:358-370,
// applied while gy is re-laid as planes) and W'[co*9 + (8 - tap), ci] = W[ci*9 + tap, co] (channels swapped, taps mirrored).
// Replaces gcol = gy_nhwc . W^T  [M, K]  +  col2im  (src/ops.rs:254-265 on the im2col matrix): no [M, 9*C_in] buffer.
// gy, relu_mask_y: NCHW [N, Cout, H, W]; dx: NCHW [N, Cin, H, W], overwritten or accumulated into.
int conv_bx3_dx(tp_ctx* ctx, const float* gy, const float* relu_mask_y, const float* w2, float* dx, int N, int Cin, int H, int W,
                int Cout, int accumulate) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return TP_ERR_UNSUPPORTED;
    if (Cout % 32 || (size_t)N * H * W * (size_t)Cout >= ((size_t)1 << 31)) return TP_ERR_UNSUPPORTED;
    const int shift_mode = g_shift_override >= 0 ? g_shift_override : default_shift_mode();
    cudaSetDevice(ctx->device);
    ConvP p;
    int smem;
    if (!plan_layer(N, H, W, /*Cin of this conv*/ Cout, /*Cout of this conv*/ Cin, false, 1, shift_mode, &p, &smem)) return TP_ERR_UNSUPPORTED;
    int rc;
    TmpBuf wplanes, planes;
    if ((rc = tp_buf_alloc(ctx, (size_t)Cout * 10 * Cin, &wplanes.b))) return rc;
    if ((rc = tp_buf_alloc(ctx, (size_t)N * H * W * Cout, &planes.b))) return rc;
    WPrep wp{};
    wp.count = 1;
    wp.w2[0] = w2;
    wp.dst[0] = (uint16_t*)wplanes.b->ptr;
    wp.cin[0] = Cout; wp.cout[0] = Cin;
    wp.adjoint[0] = 1;
    if ((rc = launch_pdl(ctx, conv_w_planes_kernel, dim3(32, 1), dim3(256), 0, false, wp))) return rc;
    const size_t items = (size_t)N * H * W * (Cout / 8);
    if ((rc = launch_pdl(ctx, nchw_to_planes_kernel, dim3(grid_for(ctx, items, 256, 8)), dim3(256), 0, true, gy, relu_mask_y,
                         (uint16_t*)planes.b->ptr, N, Cout, H, W)))
        return rc;
    p.bias = nullptr;
    p.relu = 0;
    p.accumulate = accumulate ? 1 : 0;
    p.dbg = 0;
    p.out_planes = nullptr;
    p.out_nchw = dx;
    CUtensorMap mx, mw;
    if (!make_map_x(enc, &mx, (const uint16_t*)planes.b->ptr, p)) { set_error("conv_bx3_dx: cuTensorMapEncodeTiled (gradient planes) failed"); return TP_ERR_CUDA; }
    if (!make_map_w(enc, &mw, (const uint16_t*)wplanes.b->ptr, p.CB, p.Cout)) { set_error("conv_bx3_dx: cuTensorMapEncodeTiled (weights) failed"); return TP_ERR_CUDA; }
    return launch_conv_any(ctx, mx, mw, p, smem, true);
}

// Weight gradient of a 3x3 / s1 / p1 convolution as an implicit GEMM over pixels (conv_dw_kernel above): x NCHW [N, Cin, H, W],
// gy / relu_mask_y NCHW [N, Cout, H, W], dw [Cin*9, Cout] (the reference's [K, C_out] view), overwritten or accumulated.
int conv_bx3_dw(tp_ctx* ctx, const float* x, const float* gy, const float* relu_mask_y, float* dw, int N, int Cin, int H, int W, int Cout,
                int accumulate) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return TP_ERR_UNSUPPORTED;
    if (Cin % 32 || Cout % 32 || Cin <= 0 || Cout <= 0 || W >= 32) return TP_ERR_UNSUPPORTED;
    if ((size_t)N * H * W * (size_t)(Cin > Cout ? Cin : Cout) >= ((size_t)1 << 31)) return TP_ERR_UNSUPPORTED;
    cudaSetDevice(ctx->device);
    ConvP g;
    int unused;
    if (!plan_layer(N, H, W, Cin, 32, false, 0, /*three aligned patches*/ 0, &g, &unused)) return TP_ERR_UNSUPPORTED;
    DwP p{};
    p.N = N; p.H = H; p.W = W; p.CBx = Cin / 32; p.CBz = Cout / 32;
    p.Wp = g.Wp; p.R = g.R; p.Hp = g.Hp; p.G = g.G; p.row_blocks = g.row_blocks; p.tiles = g.tiles;
    p.patch_bytes = g.patch_bytes; p.patch_alloc = g.patch_alloc;
    const int hz = g.G > 1 ? g.Hp : g.R;                        // gradient tile rows per image: the patch's image pitch, no halo
    p.z_bytes = g.G * hz * g.Wp * 128;
    const int z_alloc = ((p.z_bytes > 128 * 128 ? p.z_bytes : 128 * 128) + 1023) / 1024 * 1024;
    p.stage_bytes = 3 * p.patch_alloc + z_alloc;
    const int smem = 2 * p.stage_bytes + 1024 + 256;
    if (smem > 227 * 1024) return TP_ERR_UNSUPPORTED;
    const int pairs = p.CBx * p.CBz;
    int slices = ctx->sm_count / pairs;
    if (slices < 1) slices = 1;
    if (slices > p.tiles) slices = p.tiles;
    p.slices = slices;
    int rc;
    TmpBuf xpl, zpl, part;
    if ((rc = tp_buf_alloc(ctx, (size_t)N * H * W * Cin, &xpl.b))) return rc;
    if ((rc = tp_buf_alloc(ctx, (size_t)N * H * W * Cout, &zpl.b))) return rc;
    if ((rc = tp_buf_alloc(ctx, (size_t)pairs * slices * 9216, &part.b))) return rc;
    p.partial = part.b->ptr;
    size_t items = (size_t)N * H * W * (Cin / 8);
    if ((rc = launch_pdl(ctx, nchw_to_planes_kernel, dim3(grid_for(ctx, items, 256, 8)), dim3(256), 0, false, x, (const float*)nullptr,
                         (uint16_t*)xpl.b->ptr, N, Cin, H, W)))
        return rc;
    items = (size_t)N * H * W * (Cout / 8);
    if ((rc = launch_pdl(ctx, nchw_to_planes_kernel, dim3(grid_for(ctx, items, 256, 8)), dim3(256), 0, true, gy, relu_mask_y,
                         (uint16_t*)zpl.b->ptr, N, Cout, H, W)))
        return rc;
    CUtensorMap mx, mz;
    if (!make_map_x(enc, &mx, (const uint16_t*)xpl.b->ptr, g)) { set_error("conv_bx3_dw: cuTensorMapEncodeTiled (activation planes) failed"); return TP_ERR_CUDA; }
    ConvP gz = g;
    gz.CB = p.CBz;
    gz.Hp = hz;
    if (!make_map_x(enc, &mz, (const uint16_t*)zpl.b->ptr, gz)) { set_error("conv_bx3_dw: cuTensorMapEncodeTiled (gradient planes) failed"); return TP_ERR_CUDA; }
    static int attr_smem[16] = {};
    const int dev = ctx->device < 16 ? ctx->device : 15;
    if (ctx->device >= 16 || attr_smem[dev] < smem) {
        TP_CUDA(cudaFuncSetAttribute(conv_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_smem[dev] = smem;
    }
    if ((rc = launch_pdl(ctx, conv_dw_kernel, dim3(slices, pairs), dim3(192), (size_t)smem, true, mx, mz, p))) return rc;
    return launch_pdl(ctx, conv_dw_fold_kernel, dim3(grid_for(ctx, (size_t)pairs * 9216, 256, 4)), dim3(256), 0, true, (const float*)part.b->ptr, dw,
                      p.CBx, p.CBz, slices, Cout, accumulate ? 1 : 0);
}

}  // namespace tp

extern "C" {

int tpdbg_conv_shift_mode(int mode) {
    g_shift_override = mode;
    return 0;
}
int tpdbg_conv_times(long long* out384) {
    return cudaMemcpyFromSymbol(out384, g_conv_t, sizeof(long long) * 6 * 64) == cudaSuccess ? 0 : 1;
}
int tpdbg_conv_flags(int flags) {
    g_dbg_flags = flags;
    return 0;
}

int tp_conv_stack_fwd(tp_ctx* ctx, const tp_buf* x, int n, int c_in, int h, int w, int n_layers, const tp_buf* const* weights,
                      const tp_buf* const* biases, const int* c_out, const int* pool, const int* relu, tp_buf* y) {
    TP_CHECK_ARG(ctx && x && weights && biases && c_out && pool && relu && y, "tp_conv_stack_fwd: NULL argument");
    TP_CHECK_ARG(n_layers >= 1 && n_layers <= kMaxPrep, "tp_conv_stack_fwd: 1..%d layers", kMaxPrep);
    TP_CHECK_ARG(n > 0 && c_in > 0 && h > 0 && w > 0, "tp_conv_stack_fwd: empty input");
    TP_NEED(x, (size_t)n * c_in * h * w, "x");
    const float* wp[kMaxPrep];
    const float* bp[kMaxPrep];
    int ci = c_in, hh = h, ww = w;
    for (int l = 0; l < n_layers; ++l) {
        TP_CHECK_ARG(c_out[l] > 0, "tp_conv_stack_fwd: layer %d has no output channels", l);
        TP_NEED(weights[l], (size_t)ci * 9 * c_out[l], "weight");
        if (biases[l]) TP_NEED(biases[l], (size_t)c_out[l], "bias");
        wp[l] = weights[l]->ptr;
        bp[l] = biases[l] ? biases[l]->ptr : nullptr;
        if (pool[l]) { hh /= 2; ww /= 2; }
        ci = c_out[l];
    }
    TP_CHECK_ARG(hh > 0 && ww > 0, "tp_conv_stack_fwd: the pools leave no output");
    TP_NEED(y, (size_t)n * ci * hh * ww, "y");
    int rc = tp::conv_stack_fwd(ctx, x->ptr, n, c_in, h, w, n_layers, wp, bp, c_out, pool, relu, y->ptr, nullptr, nullptr);
    if (rc == TP_ERR_UNSUPPORTED) tp::set_error("tp_conv_stack_fwd: a layer's shape is outside the tensor-core stack (3x3/s1/p1, C_in %% 32 == 0 "
                                                 "(or C_in*9 <= 36 for the first layer), C_out in {32, 64, 128}, W < 32)");
    return rc;
}

int tp_conv_stack_gap_fwd(tp_ctx* ctx, const tp_buf* x, int n, int c_in, int h, int w, int n_layers, const tp_buf* const* weights,
                          const tp_buf* const* biases, const int* c_out, const int* pool, const int* relu, tp_buf* mean, tp_buf* cnt) {
    TP_CHECK_ARG(ctx && x && weights && biases && c_out && pool && relu && mean, "tp_conv_stack_gap_fwd: NULL argument");
    TP_CHECK_ARG(n_layers >= 1 && n_layers <= kMaxPrep, "tp_conv_stack_gap_fwd: 1..%d layers", kMaxPrep);
    TP_CHECK_ARG(n > 0 && c_in > 0 && h > 0 && w > 0, "tp_conv_stack_gap_fwd: empty input");
    TP_NEED(x, (size_t)n * c_in * h * w, "x");
    const float* wp[kMaxPrep];
    const float* bp[kMaxPrep];
    int ci = c_in;
    for (int l = 0; l < n_layers; ++l) {
        TP_CHECK_ARG(c_out[l] > 0, "tp_conv_stack_gap_fwd: layer %d has no output channels", l);
        TP_NEED(weights[l], (size_t)ci * 9 * c_out[l], "weight");
        if (biases[l]) TP_NEED(biases[l], (size_t)c_out[l], "bias");
        wp[l] = weights[l]->ptr;
        bp[l] = biases[l] ? biases[l]->ptr : nullptr;
        ci = c_out[l];
    }
    TP_NEED(mean, (size_t)n * ci, "mean");
    if (cnt) TP_NEED(cnt, (size_t)n * ci, "cnt");
    int rc = tp::conv_stack_fwd(ctx, x->ptr, n, c_in, h, w, n_layers, wp, bp, c_out, pool, relu, nullptr, mean->ptr, cnt ? cnt->ptr : nullptr);
    if (rc == TP_ERR_UNSUPPORTED) tp::set_error("tp_conv_stack_gap_fwd: a layer's shape is outside the tensor-core stack (see tp_conv_stack_fwd)");
    return rc;
}

int tp_gap_count_fwd(tp_ctx* ctx, const tp_buf* y, tp_buf* mean, tp_buf* cnt, int n, int c, int hw) {
    TP_CHECK_ARG(ctx && n >= 0 && c >= 0 && hw > 0, "tp_gap_count_fwd: bad argument");
    const size_t planes = (size_t)n * c;
    TP_NEED(y, planes * hw, "y"); TP_NEED(mean, planes, "mean");
    if (cnt) TP_NEED(cnt, planes, "cnt");
    if (!planes) return TP_OK;
    TP_CHECK_ARG(planes <= 0x7fffffff, "tp_gap_count_fwd: N*C = %zu exceeds int range", planes);
    cudaSetDevice(ctx->device);
    return launch_pdl(ctx, gap_count_kernel, dim3(tp::grid_for(ctx, planes * 16, 256, 8)), dim3(256), 0, true,
                      (const float*)y->ptr, mean->ptr, cnt ? cnt->ptr : (float*)nullptr, (int)planes, hw);
}

int tp_gap_relu_bias_grad(tp_ctx* ctx, const tp_buf* g, const tp_buf* cnt, tp_buf* gb, int n, int c, int hw, int accumulate) {
    TP_CHECK_ARG(ctx && n >= 0 && c >= 0 && hw > 0, "tp_gap_relu_bias_grad: bad argument");
    TP_NEED(g, (size_t)n * c, "g"); TP_NEED(gb, (size_t)c, "gb");
    if (cnt) TP_NEED(cnt, (size_t)n * c, "cnt");
    if (!c) return TP_OK;
    cudaSetDevice(ctx->device);
    const int S = n >= 512 ? 16 : n >= 64 ? 4 : 1;
    TmpBuf part;
    int rc = tp_buf_alloc(ctx, (size_t)S * c, &part.b);
    if (rc) return rc;
    if ((rc = launch_pdl(ctx, gap_bias_grad_stage1, dim3((c + 31) / 32, S), dim3(256), 0, true, (const float*)g->ptr,
                         cnt ? (const float*)cnt->ptr : (const float*)nullptr, part.b->ptr, n, c, hw)))
        return rc;
    return launch_pdl(ctx, gap_bias_grad_stage2, dim3((c + 255) / 256), dim3(256), 0, true, (const float*)part.b->ptr, gb->ptr, S, c, accumulate);
}

}  // extern "C"

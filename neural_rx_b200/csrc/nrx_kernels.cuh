// nrx_kernels.cuh — sm_100a kernels of the neural-receiver hot path (CGNN forward,
// reference utils/neural_rx.py:544-595 and its pre/post-processing :813-881, :1462-1514).
//
// Data layout in HBM (all activations channels-last fp16, one row per user resource element,
// row index  p = ((b*U + u)*F + f)*14 + t ):
//   z0    [P][32]   StateInit input  [y_re(N) y_im(N) pe_t pe_f h_re(N) h_im(N) 0..]   (normalised)
//   H1/H2 [P][128]  hidden activations of a sep-conv stack (ping-pong)
//   Abuf  [P][64]   aggregated messages a (56 valid, rest 0)
//   Sbuf  [P][64]   state s (56 valid) | pe_t pe_f | 0 x 6   — the pe channels ride along so that
//                   [Abuf | Sbuf] is the 128-wide UpdateState input [a, s, pe] (weights permuted)
// GEMM operands are fp16 K-major SWIZZLE_128B slabs (sm100_prims.cuh), accumulators fp32 in TMEM.
//
// Tiling: one tile = 9 subcarriers x 14 symbols = 126 rows of one (slot, user) -> one M=128 UMMA
// tile; F = 1584 = 9 * 176 so the evaluation grid has no ragged tile.  Depthwise 3x3 needs a
// +-1 subcarrier halo (11 x 14 rows staged in shared memory by one 1-D bulk copy); the symbol
// axis is entirely inside the tile.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "sm100_prims.cuh"

namespace nrx {

constexpr int kT = 14;            // OFDM symbols per slot (symbol_allocation = [0, 14])
constexpr int kTileF = 9;         // subcarriers per tile
constexpr int kTileRows = kTileF * kT;          // 126
constexpr int kHaloRows = (kTileF + 2) * kT;    // 154
constexpr int kThreads = 256;
constexpr int kPowerParts = 16;   // partial sums per slot for the input-power reduction

enum SepMode : int { kHidden = 0, kInitOut = 1, kUpdateOut = 2 };

__device__ __forceinline__ uint4 ld_shared_v4(const void* p) {
    return *reinterpret_cast<const uint4*>(p);
}
__device__ __forceinline__ void st_shared_v4(void* p, uint4 v) { *reinterpret_cast<uint4*>(p) = v; }

__device__ __forceinline__ __half2 u2h(uint32_t v) { return *reinterpret_cast<__half2*>(&v); }
__device__ __forceinline__ uint32_t h2u(__half2 v) { return *reinterpret_cast<uint32_t*>(&v); }

// =============================================================================================
// 1. input power  (CGNN.forward normalisation, utils/neural_rx.py:551-553): partial sums of y^2
// =============================================================================================
// grid (kPowerParts, B), block 256.  y: [B][N*T*F] complex64 viewed as float2.
__global__ void __launch_bounds__(256) nrx_power_kernel(const float2* __restrict__ y, float* __restrict__ partial,
                                                        int n_complex) {
    const int b = blockIdx.y, part = blockIdx.x;
    const int per = (n_complex + kPowerParts - 1) / kPowerParts;
    const int lo = part * per, hi = min(lo + per, n_complex);
    const float2* yb = y + size_t(b) * n_complex;
    float acc = 0.f;
    for (int i = lo + threadIdx.x; i < hi; i += 256) {
        const float2 v = __ldg(yb + i);
        acc = fmaf(v.x, v.x, acc);
        acc = fmaf(v.y, v.y, acc);
    }
    __shared__ float red[256];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {          // fixed-order tree: deterministic
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[b * kPowerParts + part] = red[0];
}

__device__ __forceinline__ float slot_gain(const float* __restrict__ partial, int b, int n_real) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kPowerParts; ++i) s += partial[b * kPowerParts + i];
    const float g = rsqrtf(s / float(n_real));
    return isfinite(g) ? g : 0.f;                // divide_no_nan of the TF original
}

// =============================================================================================
// 1b. inactive-user skipping: ordered list of the (slot, user) planes whose user is active.  An inactive
//     user's state never reaches an active user (its messages are masked, utils/neural_rx.py:192-193), so
//     its planes need not be computed at all; the reference computes them and ignores the result
//     (notebooks/nrx_architecture.ipynb:537).  One block; list[0 .. count) ascending, count in list[-1 + ...]:
//     out[0] = count, out[1 + i] = i-th active plane.
// =============================================================================================
__global__ void __launch_bounds__(1024) nrx_planes_kernel(const float* __restrict__ active_tx, int n_planes, int32_t* __restrict__ out) {
    __shared__ int warp_cnt[32];
    __shared__ int base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n_planes; i0 += 1024) {
        const int i = i0 + tid;
        const bool on = i < n_planes && active_tx[i] != 0.f;
        const unsigned bal = __ballot_sync(0xffffffffu, on);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int before = base;
        for (int w = 0; w < warp; ++w) before += warp_cnt[w];
        if (on) out[1 + before + __popc(bal & ((1u << lane) - 1u))] = i;
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int w = 0; w < 32; ++w) t += warp_cnt[w];
            base += t;
        }
        __syncthreads();
    }
    if (tid == 0) out[0] = base;
}

// =============================================================================================
// 2. pre-processing: LS + FOCC + nearest-pilot broadcast, normalisation, positional encoding
//    (NeuralPUSCHReceiver.estimate_channel utils/neural_rx.py:1462-1514, copy_pytorch.py:899-911;
//     CGNNOFDM.forward :832-839; StateInit concat :112-123)
// =============================================================================================
struct alignas(16) FoccEntry {   // LS estimate of pilot slot k of user u = sum_m y[src[m]] * w[m]
    int32_t src[2];    // t << 16 | f of the contributing pilot REs
    float2 w[2];       // 0.5 / pilot  (0 for a missing member)
    int32_t pad[2];    // 32 bytes: two 16-byte loads
};

struct PrepParams {
    const float2* y;          // [B][N][T][F]
    const float* partial;     // [B][kPowerParts]
    const FoccEntry* focc_re; // [U][T][F]: the LS estimate that fills RE (t, f) of user u — nearest-pilot gather
                              //   (utils/neural_rx.py:973-992) already resolved, rows in the network's (f, t) order
    const float* pos_enc;     // [U][F][T][2]
    __half* z0;               // [Bp*U*F*T][32]
    float* h_ls;              // [B][U][F][T][2N] or null
    int F, U, N, n_pilot_slots, b0, bp;
};

// A block handles kPrepF subcarriers x 14 symbols of one slot.  The received grid is [N][T][F] (subcarrier
// fastest) while the network's rows are subcarrier-major / symbol-fastest: the block first stages its
// [N][T][kPrepF] patch of y in shared memory with coalesced loads along F, then every thread builds whole
// output rows in (f, t) order from the patch (the LS gathers of a pilot slot hit its own FOCC block, which
// lies inside the patch; anything outside falls back to a global load).
constexpr int kPrepF = 16;

// NT = compile-time number of receive antennas (0: taken from the parameters at run time).  With a run-time N the
// indices of the 32-channel row (row[N + a], row[2N + 2 + a], ...) are dynamic and the row lives in local memory:
// ~800 instructions per resource element; with NT known it is built in registers.
constexpr int kPrepRowU4 = kT * 4 + 1;     // one subcarrier's 14 staged z0 rows (64 B each) + 16 B: threads of consecutive
                                           //   subcarriers then write to consecutive 16-byte bank groups
// Thread = one RE of the block's [14 symbols][16 subcarriers] patch, SUBCARRIER-fastest: the reads of the staged y patch,
// of the per-RE table (stored (t, f)-major) and the staging writes are all bank-conflict-free and coalesced.  (With the
// symbol-fastest mapping every read of the patch was a 14-way bank conflict — rows are 128 B apart — and the kernel ran
// at 58 % shared-memory wavefront utilisation.)
template <int NT>
__global__ void __launch_bounds__(256) nrx_prep_kernel(PrepParams p) {
    __shared__ float2 sy[7 * kT * kPrepF];                     // [a][t][fi]   (N <= 7)
    __shared__ uint4 srow[kPrepF * kPrepRowU4];                // one user's z0 rows of the patch: contiguous in global memory
    const int tiles_f = (p.F + kPrepF - 1) / kPrepF;
    const int bl = blockIdx.x / tiles_f, f0 = (blockIdx.x - bl * tiles_f) * kPrepF;
    const int nf = min(kPrepF, p.F - f0);
    const int per_slot = p.F * kT;
    const int b = p.b0 + bl;
    const int N = NT ? NT : p.N;
    const float2* yb = p.y + size_t(b) * N * per_slot;
    for (int i = threadIdx.x; i < N * kT * kPrepF; i += 256) {
        const int fi = i % kPrepF, at = i / kPrepF;            // at = a * T + t
        if (fi < nf) sy[i] = __ldg(yb + size_t(at) * p.F + f0 + fi);
    }
    __syncthreads();
    const float g = slot_gain(p.partial, b, 2 * N * per_slot);
    auto y_at = [&](int a, int tf) -> float2 {                 // tf = t << 16 | f
        const int t = tf >> 16, f = tf & 0xFFFF;
        if (f >= f0 && f < f0 + nf) return sy[(a * kT + t) * kPrepF + (f - f0)];
        return __ldg(yb + (size_t(a) * kT + t) * p.F + f);
    };
    const int t = threadIdx.x / kPrepF, fi = threadIdx.x % kPrepF;
    const bool on = t < kT && fi < nf;
    const int f = f0 + fi;
    for (int u = 0; u < p.U; ++u) {
        if (on) {
            const int rem = f * kT + t;
            FoccEntry e;
            {
                const uint4* ep = reinterpret_cast<const uint4*>(p.focc_re + size_t(u) * per_slot + t * p.F + f);
                const uint4 e0 = __ldg(ep), e1 = __ldg(ep + 1);
                e.src[0] = int(e0.x); e.src[1] = int(e0.y);
                e.w[0] = make_float2(__uint_as_float(e0.z), __uint_as_float(e0.w));
                e.w[1] = make_float2(__uint_as_float(e1.x), __uint_as_float(e1.y));
            }
            float hre[8], him[8];
#pragma unroll
            for (int a = 0; a < 8; ++a)
                if (a < N) {
                    const float2 y0 = y_at(a, e.src[0]);
                    const float2 y1 = y_at(a, e.src[1]);
                    hre[a] = y0.x * e.w[0].x - y0.y * e.w[0].y + (y1.x * e.w[1].x - y1.y * e.w[1].y);
                    him[a] = y0.x * e.w[0].y + y0.y * e.w[0].x + (y1.x * e.w[1].y + y1.y * e.w[1].x);
                }
            const float2 pe = *reinterpret_cast<const float2*>(p.pos_enc + ((size_t(u) * p.F + f) * kT + t) * 2);
            __align__(16) __half row[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) row[c] = __float2half(0.f);
#pragma unroll
            for (int a = 0; a < 8; ++a)
                if (a < N) {
                    const float2 v = sy[(a * kT + t) * kPrepF + fi];
                    row[a] = __float2half(v.x * g);
                    row[N + a] = __float2half(v.y * g);
                    row[2 * N + 2 + a] = __float2half(hre[a] * g);
                    row[3 * N + 2 + a] = __float2half(him[a] * g);
                }
            row[2 * N] = __float2half(pe.x);
            row[2 * N + 1] = __float2half(pe.y);
            const uint4* srcv = reinterpret_cast<const uint4*>(row);
#pragma unroll
            for (int c = 0; c < 4; ++c) srow[fi * kPrepRowU4 + t * 4 + c] = srcv[c];
            if (p.h_ls) {
                float* ho = p.h_ls + ((size_t(b) * p.U + u) * per_slot + rem) * (2 * N);
#pragma unroll
                for (int a = 0; a < 8; ++a)
                    if (a < N) {
                        ho[a] = hre[a];
                        ho[N + a] = him[a];
                    }
            }
        }
        __syncthreads();
        // the patch's rows of this user are one contiguous range of z0: coalesced 16-byte stores
        uint4* dst = reinterpret_cast<uint4*>(p.z0 + ((size_t(bl) * p.U + u) * per_slot + size_t(f0) * kT) * 32);
        for (int j = threadIdx.x; j < nf * kT * 4; j += 256) dst[j] = srow[j + j / (kT * 4)];
        __syncthreads();
    }
}
// =============================================================================================
// 2b. Aerial / TensorRT-shaped pre-processing (NRPreprocessing, utils/neural_rx.py:1614-1713):
//     rx_slot_{real,imag} [B][F][T][N], LS estimates at the non-zero pilots h_hat_{real,imag}
//     [B][n_pilots][U][N] (DMRS-symbol major), FOCC removal = mean of adjacent pilot pairs,
//     per-PRB nearest-pilot gather, normalisation, positional encoding -> the same z0 rows.
// =============================================================================================
__global__ void __launch_bounds__(256) nrx_power_planar_kernel(const float* __restrict__ re, const float* __restrict__ im,
                                                               float* __restrict__ partial, int n_real) {
    const int b = blockIdx.y, part = blockIdx.x;
    const int per = (n_real + kPowerParts - 1) / kPowerParts;
    const int lo = part * per, hi = min(lo + per, n_real);
    const float* rb = re + size_t(b) * n_real;
    const float* ib = im + size_t(b) * n_real;
    float acc = 0.f;
    for (int i = lo + threadIdx.x; i < hi; i += 256) {
        const float a = __ldg(rb + i), c = __ldg(ib + i);
        acc = fmaf(a, a, acc);
        acc = fmaf(c, c, acc);
    }
    __shared__ float red[256];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[b * kPowerParts + part] = red[0];
}

struct PrepAerialParams {
    const float* y_re;        // [B][F][T][N]
    const float* y_im;
    const float* h_re;        // [B][n_pilots][U][N]
    const float* h_im;
    const float* partial;     // [B][kPowerParts]
    const int32_t* nn_prb;    // [U][F*T]  pilot ordinal per RE, (f, t) order
    const float* pos_enc;     // [U][F][T][2]
    __half* z0;               // [Bp*U*F*T][32]
    int F, U, N, n_pilots, b0, bp;
};

template <int NT>
__global__ void __launch_bounds__(256) nrx_prep_aerial_kernel(PrepAerialParams p) {
    const int per_slot = p.F * kT;
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= p.bp * per_slot) return;
    const int bl = idx / per_slot, rem = idx - bl * per_slot;    // rem = f * T + t
    const int b = p.b0 + bl;
    const int N = NT ? NT : p.N;
    const float g = slot_gain(p.partial, b, 2 * N * per_slot);
    const float* yr = p.y_re + (size_t(b) * per_slot + rem) * N;
    const float* yi = p.y_im + (size_t(b) * per_slot + rem) * N;
    for (int u = 0; u < p.U; ++u) {
        const int k = __ldg(p.nn_prb + size_t(u) * per_slot + rem);
        const size_t h0 = ((size_t(b) * p.n_pilots + k) * p.U + u) * N;
        const size_t h1 = ((size_t(b) * p.n_pilots + (k ^ 1)) * p.U + u) * N;   // FOCC partner (:1620-1629)
        const float2 pe = *reinterpret_cast<const float2*>(p.pos_enc + (size_t(u) * per_slot + rem) * 2);
        __align__(16) __half row[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) row[c] = __float2half(0.f);
#pragma unroll
        for (int a = 0; a < 8; ++a)
            if (a < N) {
                row[a] = __float2half(__ldg(yr + a) * g);
                row[N + a] = __float2half(__ldg(yi + a) * g);
                row[2 * N + 2 + a] = __float2half(0.5f * (__ldg(p.h_re + h0 + a) + __ldg(p.h_re + h1 + a)) * g);
                row[3 * N + 2 + a] = __float2half(0.5f * (__ldg(p.h_im + h0 + a) + __ldg(p.h_im + h1 + a)) * g);
            }
        row[2 * N] = __float2half(pe.x);
        row[2 * N + 1] = __float2half(pe.y);
        const size_t prow = (size_t(bl) * p.U + u) * per_slot + rem;
        uint4* dst = reinterpret_cast<uint4*>(p.z0 + prow * 32);
        const uint4* srcv = reinterpret_cast<const uint4*>(row);
#pragma unroll
        for (int c = 0; c < 4; ++c) dst[c] = srcv[c];
    }
}

// =============================================================================================
// 3. separable-conv layer: depthwise 3x3 (CUDA cores, HFMA2) -> swizzled A operand in smem ->
//    pointwise GEMM on tcgen05 (fp32 accumulators in TMEM) -> bias / ReLU / residual epilogue
//    (Keras SeparableConv2D of StateInit utils/neural_rx.py:61-132 and UpdateState :210-270)
// =============================================================================================
struct SepParams {
    const __half* src0;        // [P][C0]
    const __half* src1;        // [P][C1] or null        (C0 + C1 == KPAD)
    const uint8_t* wblob;      // per stack: [pw image | dw taps [9][KPAD] fp16 | bias [NPAD] fp32]
    const int32_t* stack_index;// [BU] (already offset to the pass) or null
    __half* out;               // kHidden: [P][NPAD]; kInitOut / kUpdateOut: Sbuf [P][64]
    const float* pos_enc;      // [U][F][T][2]   (kInitOut)
    int C0, C1;
    int F, U, d_s;
    int tiles_per_bu, num_tiles;
    int default_stack;
    int n_stacks;              // stacks in wblob: stack_index values are clamped to [0, n_stacks)
    uint32_t blob_bytes;
};

template <int KPAD, int NPAD>
struct SepSmem {
    static constexpr int KS = (KPAD + 63) / 64;
    static constexpr int kA = 32768;                         // A operand, later the output staging
    static constexpr int kWpw = KS * NPAD * 128;
    static constexpr int kDw = 9 * KPAD * 2;
    static constexpr int kBias = NPAD * 4;
    static constexpr int kBlob = kWpw + kDw + kBias;
    static constexpr int kIn = kHaloRows * KPAD * 2;
    static constexpr int offA = 0;
    static constexpr int offW = kA;
    static constexpr int offIn = ((offW + kBlob + 127) / 128) * 128;
    static constexpr int kTotal = offIn + kIn + 1024;        // + alignment slack
};

template <int KPAD, int NPAD, int MODE>
__global__ void __launch_bounds__(kThreads, 2) nrx_sepconv_kernel(SepParams p) {
    using L = SepSmem<KPAD, NPAD>;
    constexpr int KS = L::KS;
    constexpr int NCV = KPAD / 8;                  // 16-byte channel vectors per row
    constexpr int NSEG = (KPAD >= 128) ? 1 : 3;    // subcarrier segments per thread column
    constexpr int SEGF = kTileF / NSEG;
    constexpr int NTASK = NSEG * kT * NCV;
    static_assert(NTASK <= kThreads, "depthwise task mapping");
    static_assert(MODE == kHidden || NPAD == 64, "state output is 64 wide");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
    uint8_t* sA = smem + L::offA;
    uint8_t* sW = smem + L::offW;
    const __half* sDw = reinterpret_cast<const __half*>(sW + L::kWpw);
    const float* sBias = reinterpret_cast<const float*>(sW + L::kWpw + L::kDw);
    uint8_t* sIn = smem + L::offIn;
    __shared__ uint64_t bar_in, bar_w, bar_mma;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&tmem_slot, 128);
    if (tid == 0) {
        mbar_init(&bar_in, 1);
        mbar_init(&bar_w, 1);
        mbar_init(&bar_mma, 1);
        fence_mbar_init();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = tmem_slot;
    uint32_t ph_in = 0, ph_w = 0, ph_mma = 0;
    int loaded_stack = -1;

    const int rowB0 = p.C0 * 2, rowB1 = p.C1 * 2;
    uint8_t* sIn0 = sIn;
    uint8_t* sIn1 = sIn + kHaloRows * rowB0;

    // issue the halo-tile load of `tile` (thread 0) and zero the out-of-grid halo rows (all threads)
    auto stage_input = [&](int tile) {
        const int bu = tile / p.tiles_per_bu, ft = tile - bu * p.tiles_per_bu;
        const int f0 = ft * kTileF;
        const int flo = max(f0 - 1, 0), fhi = min(f0 + kTileF + 1, p.F);     // valid [flo, fhi)
        const int fi_lo = flo - (f0 - 1), fi_hi = fhi - (f0 - 1);            // halo-row indices
        if (tid == 0) {
            const uint32_t nrow = uint32_t(fhi - flo) * kT;
            mbar_arrive_expect_tx(&bar_in, nrow * uint32_t(rowB0 + rowB1));
            const size_t grow = (size_t(bu) * p.F + flo) * kT;
            bulk_g2s(sIn0 + fi_lo * kT * rowB0, reinterpret_cast<const uint8_t*>(p.src0) + grow * rowB0,
                     nrow * rowB0, &bar_in);
            if (rowB1)
                bulk_g2s(sIn1 + fi_lo * kT * rowB1, reinterpret_cast<const uint8_t*>(p.src1) + grow * rowB1,
                         nrow * rowB1, &bar_in);
        }
        // rows [0, fi_lo) and [fi_hi, 11) are outside the grid: zero ('same' padding)
        const uint4 z = make_uint4(0, 0, 0, 0);
        const int v0 = rowB0 / 16, v1 = rowB1 / 16;
        for (int fi = 0; fi < kTileF + 2; ++fi) {
            if (fi >= fi_lo && fi < fi_hi) continue;
            for (int i = tid; i < kT * (v0 + v1); i += kThreads) {
                if (i < kT * v0) st_shared_v4(sIn0 + fi * kT * rowB0 + i * 16, z);
                else st_shared_v4(sIn1 + fi * kT * rowB1 + (i - kT * v0) * 16, z);
            }
        }
    };

    int tile = blockIdx.x;
    if (tile < p.num_tiles) stage_input(tile);

    while (tile < p.num_tiles) {
        const int bu = tile / p.tiles_per_bu, ft = tile - bu * p.tiles_per_bu;
        const int f0 = ft * kTileF;
        const int valid_rows = min(kTileF, p.F - f0) * kT;
        const int stack = min(max(p.stack_index ? p.stack_index[bu] : p.default_stack, 0), p.n_stacks - 1);
        if (stack != loaded_stack) {                 // block-uniform; first tile or Var-IO switch
            __syncthreads();                         // nobody still reads the old taps / bias
            if (tid == 0) {
                mbar_arrive_expect_tx(&bar_w, p.blob_bytes);
                bulk_g2s(sW, p.wblob + size_t(stack) * p.blob_bytes, p.blob_bytes, &bar_w);
            }
            mbar_wait(&bar_w, ph_w);
            ph_w ^= 1;
            loaded_stack = stack;
        }
        mbar_wait(&bar_in, ph_in);
        ph_in ^= 1;
        __syncthreads();                             // zero-filled halo rows of this tile visible

        // ---------------- depthwise 3x3: thread = (segment, symbol t, channel vector cv) ------
        if (tid < NTASK) {
            const int cv = tid % NCV;
            const int t = (tid / NCV) % kT;
            const int seg = tid / (NCV * kT);
            const uint8_t* base;
            int rowB;
            if (cv * 8 < p.C0) { base = sIn0 + cv * 16; rowB = rowB0; }
            else { base = sIn1 + (cv * 8 - p.C0) * 2; rowB = rowB1; }
            uint4 kk[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) kk[i] = ld_shared_v4(reinterpret_cast<const uint8_t*>(sDw) + (i * KPAD + cv * 8) * 2);
            const bool has_l = t > 0, has_r = t < kT - 1;
            const uint4 z = make_uint4(0, 0, 0, 0);
            uint4 win[3][3];
            auto load_row = [&](int fi, uint4 (&r)[3]) {
                const uint8_t* q = base + (fi * kT + t) * rowB;
                r[0] = has_l ? ld_shared_v4(q - rowB) : z;
                r[1] = ld_shared_v4(q);
                r[2] = has_r ? ld_shared_v4(q + rowB) : z;
            };
            const int fs = seg * SEGF;               // first output subcarrier (tile-local)
            load_row(fs, win[0]);                    // halo index fi = f_local + 1 -> f_local - 1 is fi = f_local
            load_row(fs + 1, win[1]);
#pragma unroll
            for (int s = 0; s < SEGF; ++s) {
                const int fl = fs + s;
                load_row(fl + 2, win[(s + 2) % 3]);
                __half2 acc[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[c] = __float2half2_rn(0.f);
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const uint4(&r)[3] = win[(s + i) % 3];
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const uint4 x = r[j], w = kk[i * 3 + j];
                        acc[0] = __hfma2(u2h(x.x), u2h(w.x), acc[0]);
                        acc[1] = __hfma2(u2h(x.y), u2h(w.y), acc[1]);
                        acc[2] = __hfma2(u2h(x.z), u2h(w.z), acc[2]);
                        acc[3] = __hfma2(u2h(x.w), u2h(w.w), acc[3]);
                    }
                }
                const int r = fl * kT + t;
                st_shared_v4(sA + (cv >> 3) * 16384 + r * 128 + (((cv & 7) ^ (r & 7)) << 4),
                             make_uint4(h2u(acc[0]), h2u(acc[1]), h2u(acc[2]), h2u(acc[3])));
            }
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();                             // A operand complete, sIn free

        if (tid == 0) {
            tc_fence_after_sync();
            umma_gemm_k(tbase, smem_u32(sA), 16384, smem_u32(sW), NPAD * 128, KPAD, umma_idesc_f16(128, NPAD), false);
            umma_commit(&bar_mma);
        }
        const int next = tile + gridDim.x;
        if (next < p.num_tiles) stage_input(next);   // overlaps the GEMM and the epilogue

        mbar_wait(&bar_mma, ph_mma);
        ph_mma ^= 1;
        tc_fence_after_sync();

        // ---------------- epilogue: TMEM -> registers -> staging (A buffer, now free) ----------
        const int q = warp & 3, hcol = warp >> 2;
        const int r = q * 32 + lane;
        if constexpr (MODE == kHidden) {
            constexpr int COLS = NPAD / 2;           // columns per warp
#pragma unroll
            for (int c0 = 0; c0 < COLS; c0 += 32) {
                float v[32];
                const int col = hcol * COLS + c0;
                tmem_ld32(tmem_addr(tbase, q * 32, col), v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    uint32_t pk[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float a = fmaxf(v[j + 2 * e] + sBias[col + j + 2 * e], 0.f);
                        const float b = fmaxf(v[j + 2 * e + 1] + sBias[col + j + 2 * e + 1], 0.f);
                        pk[e] = pack_half2(a, b);
                    }
                    const int cc = (col + j) >> 3;
                    st_shared_v4(sA + (cc >> 3) * 16384 + r * 128 + (((cc & 7) ^ (r & 7)) << 4),
                                 make_uint4(pk[0], pk[1], pk[2], pk[3]));
                }
            }
        } else {
            // fp32 staging [128][64]: 16 chunks of 4 floats per row, chunk index XOR (row & 7)
            float v[32];
            const int col = hcol * 32;
            tmem_ld32(tmem_addr(tbase, q * 32, col), v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const int c4 = (col + j) >> 2;
                float4 o;
                o.x = v[j] + sBias[col + j];
                o.y = v[j + 1] + sBias[col + j + 1];
                o.z = v[j + 2] + sBias[col + j + 2];
                o.w = v[j + 3] + sBias[col + j + 3];
                *reinterpret_cast<float4*>(sA + r * 256 + ((c4 ^ (r & 7)) << 4)) = o;
            }
        }
        tc_fence_before_sync();
        __syncthreads();                             // staging complete; TMEM drained

        // ---------------- coalesced copy-out ---------------------------------------------------
        const size_t row0 = (size_t(bu) * p.F + f0) * kT;
        if constexpr (MODE == kHidden) {
            constexpr int CPR = NPAD / 8;            // 16-byte chunks per row
            for (int i = tid; i < valid_rows * CPR; i += kThreads) {
                const int rr = i / CPR, cc = i - rr * CPR;
                const uint4 val = ld_shared_v4(sA + (cc >> 3) * 16384 + rr * 128 + (((cc & 7) ^ (rr & 7)) << 4));
                *reinterpret_cast<uint4*>(p.out + (row0 + rr) * NPAD + cc * 8) = val;
            }
        } else {
            const int u = bu % p.U;
            for (int i = tid; i < valid_rows * 16; i += kThreads) {
                const int rr = i >> 4, c4 = i & 15;
                const float4 o = *reinterpret_cast<const float4*>(sA + rr * 256 + ((c4 ^ (rr & 7)) << 4));
                __half* dst = p.out + (row0 + rr) * 64 + c4 * 4;
                float a0 = o.x, a1 = o.y, a2 = o.z, a3 = o.w;
                if constexpr (MODE == kUpdateOut) {  // residual: s <- s + update  (:266)
                    const uint2 old = *reinterpret_cast<const uint2*>(dst);
                    const float2 o01 = __half22float2(u2h(old.x)), o23 = __half22float2(u2h(old.y));
                    a0 += o01.x; a1 += o01.y; a2 += o23.x; a3 += o23.y;
                } else {                             // state init: append the positional encoding
                    const int pe_chunk = p.d_s >> 2;
                    if (c4 == pe_chunk) {
                        const int fl = rr / kT, tt = rr - fl * kT;
                        const float2 pe = *reinterpret_cast<const float2*>(
                            p.pos_enc + ((size_t(u) * p.F + f0 + fl) * kT + tt) * 2);
                        a0 = pe.x; a1 = pe.y; a2 = 0.f; a3 = 0.f;
                    } else if (c4 > pe_chunk) {
                        a0 = a1 = a2 = a3 = 0.f;
                    }
                }
                uint2 pk;
                pk.x = pack_half2(a0, a1);
                pk.y = pack_half2(a2, a3);
                *reinterpret_cast<uint2*>(dst) = pk;
            }
        }
        __syncthreads();                             // staging free before the next depthwise pass
        tile = next;
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 128);
}

// =============================================================================================
// 4. AggregateUserStates (utils/neural_rx.py:135-207): per-RE MLP d_s -> units_agg -> d_s on every
//    user's state, masked sum over the *other* users, 1/(n_active-1) scaling.  One CTA holds the
//    same 128 resource elements of all U users, so the cross-user reduction is CTA-local.
// =============================================================================================
struct alignas(64) AggParams {
    CUtensorMap map_s;         // state tensor [Bp*U][F*T][64] fp16, box = 128 rows x 64 channels, 128-byte swizzle:
                               //   a box lands in shared memory as the K-major UMMA A tile (rows past F*T read zeros)
    const __half* sbuf;        // [Bp*U*F*T][64]
    __half* abuf;              // [Bp*U*F*T][64]
    const uint8_t* wblob;      // [W1 image 64x64 | W2 image 64x64 | b1[64] | b2[64]]
    const float* active_tx;    // [Bp][U] (already offset to the pass)
    int U, rows_per_bu, tiles_per_b, num_tiles;
    int skip_idle;             // inactive-user skipping: a slot with at most one active user has no messages to
                               //   compute (every a row of an active user is exactly zero): no GEMMs, zeros stored
};

constexpr int kAggBlob = 8192 + 8192 + 256 + 256;
constexpr int kAggMaxU = 4;
constexpr int kAggThreads = 128;          // four warps = the four TMEM lane quadrants of one 128-row tile

__host__ __device__ constexpr int agg_smem_bytes(int U) { return U * 16384 + ((kAggBlob + 127) / 128) * 128 + 1024; }
__host__ __device__ constexpr int agg_ctas_per_sm(int U) { return U <= 2 ? 4 : U == 3 ? 3 : 2; }

// Small CTAs, many per SM: a tile is a chain of two GEMM round trips (issue -> commit -> wake, ~1.3 k cycles each)
// with little arithmetic in between, so what pays is the number of tiles in flight per SM.  One CTA = 128 threads,
// ONE shared-memory stage of U x 16 KB that holds, in turn, the state tiles (TMA, A of the first GEMM), the hidden
// activations (A of the second GEMM) and the output staging, U x 64 TMEM columns: four CTAs per SM at U <= 2
// (49 KB, 128 columns each).  Measured on B200 (nrx_large, batch 30): 0.656 ms per step against 0.678 ms for two
// 256-thread CTAs with two stages; with 340 MB moved per launch (state in, messages out) that is 4.2 TB/s.
template <int U>
__global__ void __launch_bounds__(kAggThreads, agg_ctas_per_sm(U)) nrx_agg_kernel(const __grid_constant__ AggParams p) {
    constexpr uint32_t TM_COLS = (U <= 1) ? 64 : (U == 2) ? 128 : 256;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
    uint8_t* sA = smem;
    uint8_t* sH = sA;
    uint8_t* sW = smem + U * 16384;
    const float* sB1 = reinterpret_cast<const float*>(sW + 16384);
    const float* sB2 = sB1 + 64;
    __shared__ uint64_t bar_w, bar_mma, bar_ld;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&tmem_slot, TM_COLS);
    auto load_tile = [&](int tile) {                      // thread 0
        const int b = tile / p.tiles_per_b, rt = tile - b * p.tiles_per_b;
        mbar_arrive_expect_tx(&bar_ld, U * 16384);
#pragma unroll
        for (int u = 0; u < U; ++u) tma_load_3d(sA + u * 16384, &p.map_s, 0, rt * 128, b * U + u, &bar_ld);
    };
    if (tid == 0) {
        mbar_init(&bar_w, 1);
        mbar_init(&bar_mma, 1);
        mbar_init(&bar_ld, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(&bar_w, kAggBlob);
        bulk_g2s(sW, p.wblob, kAggBlob, &bar_w);
        if (int(blockIdx.x) < p.num_tiles) load_tile(blockIdx.x);
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = tmem_slot;
    mbar_wait(&bar_w, 0);
    uint32_t ph_mma = 0, ph_ld = 0;
    const int r = warp * 32 + lane;                       // accumulator row of this thread

    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int b = tile / p.tiles_per_b, rt = tile - b * p.tiles_per_b;
        const int r0 = rt * 128;
        const int valid_rows = min(128, p.rows_per_bu - r0);
        const int next = tile + int(gridDim.x);
        // activity flags of the slot: fetched now, used in the output epilogue
        float m[U];
#pragma unroll
        for (int u = 0; u < U; ++u) m[u] = __ldg(p.active_tx + b * U + u);
        mbar_wait(&bar_ld, ph_ld);
        ph_ld ^= 1;
        bool idle = false;
        if (p.skip_idle) {
            float n_on = 0.f;
#pragma unroll
            for (int u = 0; u < U; ++u) n_on += m[u];
            idle = n_on <= 1.f;                           // block-uniform: the sum over the OTHER active users is empty
        }
        if (idle) {
            __syncthreads();                              // every thread has passed the load barrier
            if (tid == 0 && next < p.num_tiles) load_tile(next);
            for (int i = tid; i < U * 128 * 8; i += kAggThreads) {
                const int u = i >> 10, rr = (i >> 3) & 127, cc = i & 7;
                if (rr < valid_rows && m[u] != 0.f)
                    *reinterpret_cast<uint4*>(p.abuf + ((size_t(b) * U + u) * p.rows_per_bu + r0 + rr) * 64 + cc * 8) =
                        make_uint4(0, 0, 0, 0);
            }
            continue;
        }
        tc_fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after_sync();
#pragma unroll
            for (int u = 0; u < U; ++u)
                umma_gemm_k(tbase + u * 64, smem_u32(sA + u * 16384), 16384, smem_u32(sW), 8192, 64,
                            umma_idesc_f16(128, 64), false);
            umma_commit(&bar_mma);
        }
        mbar_wait(&bar_mma, ph_mma);
        ph_mma ^= 1;
        tc_fence_after_sync();
        // ---- hidden layer epilogue: ReLU(acc + b1) -> fp16 -> sH (the state tiles have been consumed) ------
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int col = 0; col < 64; col += 32) {
                float v[32];
                tmem_ld32(tmem_addr(tbase + u * 64, warp * 32, col), v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    uint32_t pk[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        pk[e] = pack_half2(fmaxf(v[j + 2 * e] + sB1[col + j + 2 * e], 0.f),
                                           fmaxf(v[j + 2 * e + 1] + sB1[col + j + 2 * e + 1], 0.f));
                    const int cc = (col + j) >> 3;
                    st_shared_v4(sH + u * 16384 + r * 128 + ((cc ^ (r & 7)) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
                }
            }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after_sync();
#pragma unroll
            for (int u = 0; u < U; ++u)
                umma_gemm_k(tbase + u * 64, smem_u32(sH + u * 16384), 16384, smem_u32(sW + 8192), 8192, 64,
                            umma_idesc_f16(128, 64), false);
            umma_commit(&bar_mma);
        }
        mbar_wait(&bar_mma, ph_mma);
        ph_mma ^= 1;
        tc_fence_after_sync();
        // ---- output epilogue: masked sum over the other users (:192-204) -> staging in sA ------
        {
            float n_act = 0.f;
#pragma unroll
            for (int u = 0; u < U; ++u) n_act += m[u];
            const float pm = fmaxf(n_act - 1.f, 0.f);
            const float scale = (pm == 0.f) ? 1.f : 1.f / pm;
#pragma unroll
            for (int col = 0; col < 64; col += 32) {
                float sp[U][32];
#pragma unroll
                for (int u = 0; u < U; ++u) tmem_ld32(tmem_addr(tbase + u * 64, warp * 32, col), sp[u]);
                tmem_ld_wait();
                if (p.skip_idle) {                        // planes of skipped users hold stale data: select, never 0 * x
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (m[u] == 0.f) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) sp[u][j] = -sB2[col + j];
                        }
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float tot = 0.f;
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        sp[u][j] = (sp[u][j] + sB2[col + j]) * m[u];
                        tot += sp[u][j];
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) sp[u][j] = (tot - sp[u][j]) * scale;
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        const int cc = (col + j) >> 3;
                        st_shared_v4(sA + u * 16384 + r * 128 + ((cc ^ (r & 7)) << 4),
                                     make_uint4(pack_half2(sp[u][j], sp[u][j + 1]), pack_half2(sp[u][j + 2], sp[u][j + 3]),
                                                pack_half2(sp[u][j + 4], sp[u][j + 5]), pack_half2(sp[u][j + 6], sp[u][j + 7])));
                    }
            }
        }
        tc_fence_before_sync();
        __syncthreads();
        for (int i = tid; i < U * 128 * 8; i += kAggThreads) {
            const int u = i >> 10, rr = (i >> 3) & 127, cc = i & 7;
            if (rr < valid_rows)
                *reinterpret_cast<uint4*>(p.abuf + ((size_t(b) * U + u) * p.rows_per_bu + r0 + rr) * 64 + cc * 8) =
                    ld_shared_v4(sA + u * 16384 + rr * 128 + ((cc ^ (rr & 7)) << 4));
        }
        fence_proxy_async_smem();                // the stage is refilled by TMA (async proxy): generic accesses first
        __syncthreads();
        if (tid == 0 && next < p.num_tiles) load_tile(next);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, TM_COLS);
}

// =============================================================================================
// 5. read-outs (ReadoutLLRs utils/neural_rx.py:309-355, ReadoutChEst :358-404) fused with the
//    resource-grid demapping of CGNNOFDM.forward (:843-858; ordering utils/onnx_utils.py:486-514).
//    Both heads share one GEMM chain: hidden [128 x 256] = s . [W_llr1 | W_h1], then
//    [128 x 32] = hidden . blockdiag(W_llr2, W_h2).
// =============================================================================================
struct ReadoutParams {
    const __half* sbuf;          // [Bp*U*F*T][64]
    const uint8_t* wblob;        // per LLR head: [W1 image 256x64 | W2 image 32x256 | b1[256] | b2[32]]
    const int32_t* head_index;   // [Bp*U] or null
    const int32_t* data_index;   // [T*F]: ordinal among data REs or -1
    float* llr;                  // [Bp][U][n_data*out_bits] or null
    float* llr_grid;             // [Bp][U][F][T][out_bits]  or null
    float* h_ref;                // [Bp][U][F][T][2N]        or null
    float* llr_aerial;           // [Bp][out_bits][U][F][T] = -LLR (NeuralReceiverONNX, :1809-1810) or null
    int F, U, N2, out_bits, n_data;
    int rows_per_bu, tiles_per_bu, num_tiles, default_head;
    int n_heads;                 // heads in wblob: head_index values are clamped to [0, n_heads)
    const int32_t* plane_list;   // inactive-user skipping: [0] = number of active planes, [1 + i] = i-th one; or null
    int vec;                     // output pointers are 16-byte aligned: vector stores
};

constexpr int kRoW1 = 256 * 128, kRoW2 = 4 * 32 * 128;
// The first layer's bias rides in the GEMM: state channels 62 and 63 (always padding, d_s <= 60) are set to 1.0 when a
// tile is staged and rows 62 / 63 of the W1 image hold the fp16 high and low halves of the bias (about 22 mantissa bits).
constexpr int kRoBiasK = 62;
// TMEM columns: first accumulator [0, 256) fp32, hidden activations [256, 384) as fp16 pairs (the A operand of the second
// GEMM is read from tensor memory: no 64 KB hidden tile in shared memory, no operand fetch for it), output [384, 416)
constexpr uint32_t kRoHidCol = 256, kRoOutCol = 384;
#ifndef NRX_RO_AHEAD
#define NRX_RO_AHEAD 2
#endif
constexpr int kRoAhead = NRX_RO_AHEAD;   // state tiles in flight per CTA (registers); measured 2: 139 us, 3: 146, 4: 157
constexpr int kRoBlob = kRoW1 + kRoW2 + 1024 + 128;
constexpr int kRoSmem = 32768 + ((kRoBlob + 127) / 128) * 128 + 1024;

// One CTA per SM.  Two per SM (state tile aliased onto the hidden tile, second accumulator onto the first: 106 KB,
// 256 TMEM columns) were measured equal (0.220 vs 0.213 ms per 30-slot step): a tile is bound by the 256-column
// accumulator read + bias/ReLU/pack + 64 KB hidden-tile write, not by the tensor pipe.
// The tile loop is software-pipelined over two state tiles: tile t+1's first GEMM is issued right behind tile t's second
// one (its accumulator has just been drained), so it runs under tile t's output epilogue and only the second GEMM's round
// trip stays on the per-tile chain.  A change of LLR head between two tiles (Var-IO) falls back to the serial start.
__global__ void __launch_bounds__(kThreads, 1) nrx_readout_kernel(ReadoutParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
    uint8_t* sA = smem;                   // 2 x [128][128 B]
    uint8_t* sW = smem + 32768;
    const float* sB2 = reinterpret_cast<const float*>(sW + kRoW1 + kRoW2) + 256;
    __shared__ uint64_t bar_w, bar_m1, bar_m2;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    if (tid == 0) {
        mbar_init(&bar_w, 1);
        mbar_init(&bar_m1, 1);
        mbar_init(&bar_m2, 1);
        fence_mbar_init();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = tmem_slot;
    uint32_t ph_w = 0, ph_m1 = 0, ph_m2 = 0;
    int loaded_head = -1;
    const int q = warp & 3, hcol = warp >> 2;
    const int r = q * 32 + lane;

    const int num_tiles = p.plane_list ? p.plane_list[0] * p.tiles_per_bu : p.num_tiles;
    auto plane_of = [&](int tile) {
        const int pl = tile / p.tiles_per_bu;
        return p.plane_list ? p.plane_list[1 + pl] : pl;
    };
    auto head_of = [&](int bu) { return min(max(p.head_index ? p.head_index[bu] : p.default_head, 0), p.n_heads - 1); };

    // state rows of the next kRoAhead tiles travel in registers: with one tile (16 KB per SM) in flight the kernel was
    // bound by the latency of these loads (0.18 -> 0.12 ms per 30-slot step with the loads stubbed out)
    constexpr int NV = 128 * 8 / kThreads;
    uint4 pre[kRoAhead][NV];
    auto fetch = [&](int tile, uint4 (&dst)[NV]) {
#pragma unroll
        for (int v = 0; v < NV; ++v) dst[v] = make_uint4(0, 0, 0, 0);
        if (tile >= num_tiles) return;
        const int pl = tile / p.tiles_per_bu, rt = tile - pl * p.tiles_per_bu;
        const int bu = p.plane_list ? p.plane_list[1 + pl] : pl;
        const int r0 = rt * 128, valid_rows = min(128, p.rows_per_bu - r0);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int i = tid + v * kThreads;
            const int rr = i >> 3, cc = i & 7;
#ifndef NRX_RO_NOFETCH
            if (rr < valid_rows)
                dst[v] = __ldg(reinterpret_cast<const uint4*>(p.sbuf + (size_t(bu) * p.rows_per_bu + r0 + rr) * 64 + cc * 8));
#endif
        }
    };
    // writes the tile held in pre[0] to a state buffer, moves the others up and starts the loads of the tile after them
    auto stage = [&](uint8_t* dst, int tile_after) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int i = tid + v * kThreads;
            const int rr = i >> 3, cc = i & 7;
            uint4 x = pre[0][v];
            if (cc == 7) x.w = 0x3C003C00u;   // channels 62, 63 = 1.0: they carry the first layer's bias (kRoBiasK)
            st_shared_v4(dst + rr * 128 + ((cc ^ (rr & 7)) << 4), x);
#pragma unroll
            for (int d = 0; d + 1 < kRoAhead; ++d) pre[d][v] = pre[d + 1][v];
        }
        fetch(tile_after, pre[kRoAhead - 1]);
    };
    auto issue_first = [&](int buf) {
#ifndef NRX_RO_NOMMA1
        umma_gemm_k(tbase, smem_u32(sA + buf * 16384), 16384, smem_u32(sW), 256 * 128, 64, umma_idesc_f16(128, 256), false);
#endif
        umma_commit(&bar_m1);
    };
#pragma unroll
    for (int d = 0; d < kRoAhead; ++d) fetch(blockIdx.x + d * gridDim.x, pre[d]);
    bool first_issued = false;            // this tile's first GEMM was issued during the previous tile
    int buf = 0;

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int pl = tile / p.tiles_per_bu, rt = tile - pl * p.tiles_per_bu;
        const int bu = p.plane_list ? p.plane_list[1 + pl] : pl;
        const int r0 = rt * 128;
        const int valid_rows = min(128, p.rows_per_bu - r0);
        if (!first_issued) {
            const int head = head_of(bu);
            if (head != loaded_head) {
                __syncthreads();          // the previous tile's output epilogue still reads the biases
                if (tid == 0) {
                    mbar_arrive_expect_tx(&bar_w, kRoBlob);
                    bulk_g2s(sW, p.wblob + size_t(head) * kRoBlob, kRoBlob, &bar_w);
                }
                mbar_wait(&bar_w, ph_w);
                ph_w ^= 1;
                loaded_head = head;
            }
            stage(sA + buf * 16384, tile + kRoAhead * int(gridDim.x));
            fence_proxy_async_smem();
            tc_fence_before_sync();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after_sync();
                issue_first(buf);
            }
        }
        const int next = tile + int(gridDim.x);
        const bool has_next = next < num_tiles;
        // demapped position of this thread's row: fetched now, used after the second GEMM (the load used to sit,
        // with its full latency, on the four warps of the output epilogue)
        int d_row = -1;
        if (warp < 4 && p.llr && r < valid_rows) {
            const int prow = r0 + r, f = prow / kT, t = prow - f * kT;
            d_row = __ldg(p.data_index + t * p.F + f);
        }
        const bool pipe = has_next && head_of(plane_of(next)) == loaded_head;
        mbar_wait(&bar_m1, ph_m1);
        ph_m1 ^= 1;
        tc_fence_after_sync();
        {   // hidden epilogue: 128 columns per thread in chunks of 32, the next chunk's TMEM load in flight during the
            // conversion of the current one; the bias arrived through the GEMM, ReLU is fused into the fp16x2 conversion
            float v[2][32];
#ifdef NRX_RO_NOLDTM
            for (int j = 0; j < 32; ++j) v[0][j] = v[1][j] = float(tid + j);
#else
            tmem_ld32(tmem_addr(tbase, q * 32, hcol * 128), v[0]);
            tmem_ld_wait();
#endif
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int col = hcol * 128 + c * 32;
#ifndef NRX_RO_NOLDTM
                if (c < 3) tmem_ld32(tmem_addr(tbase, q * 32, col + 32), v[(c + 1) & 1]);
#endif
                const float(&x)[32] = v[c & 1];
                uint32_t hw[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) hw[j] = pack_relu_half2(x[2 * j], x[2 * j + 1]);
#ifndef NRX_RO_NOHST
                tmem_st16(tmem_addr(tbase + kRoHidCol, q * 32, hcol * 64 + c * 16), hw);
#endif
#ifndef NRX_RO_NOLDTM
                if (c < 3) tmem_ld_wait();
#endif
            }
        }
        tmem_st_wait();
        if (pipe) stage(sA + (buf ^ 1) * 16384, tile + (kRoAhead + 1) * int(gridDim.x));   // its last reader was the first GEMM of the tile before this one
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after_sync();
#ifndef NRX_RO_NOMMA2
            umma_gemm_k_ts(tbase + kRoOutCol, tbase + kRoHidCol, smem_u32(sW + kRoW1), 32 * 128, 256, umma_idesc_f16(128, 32));
#endif
            umma_commit(&bar_m2);
            if (pipe) issue_first(buf ^ 1);        // accumulator columns 0..255 were drained before the barrier above
        }
        first_issued = pipe;
        if (pipe) buf ^= 1;
        mbar_wait(&bar_m2, ph_m2);
        ph_m2 ^= 1;
        tc_fence_after_sync();
        {   // output epilogue: warps 0..3 own the LLR columns 0..15, warps 4..7 the channel-estimate columns 16..31
            float v[16];
            tmem_ld16(tmem_addr(tbase + kRoOutCol, q * 32, hcol * 16), v);
            tmem_ld_wait();
#ifdef NRX_RO_NOSTG
            if (r < valid_rows && v[0] == 1234.5f) {
#else
            if (r < valid_rows) {
#endif
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] += sB2[hcol * 16 + j];
                const int prow = r0 + r;                      // row inside the (slot, user) grid
                const size_t grow = size_t(bu) * p.rows_per_bu + prow;
                if (hcol == 0) {
                    // vector stores for the shipped widths (2 / 4 / 6 bits per symbol): a row's values are contiguous
                    auto store_bits = [&](float* o) {
                        if (p.vec && p.out_bits == 4) {
                            *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
                        } else if (p.vec && p.out_bits == 2) {
                            *reinterpret_cast<float2*>(o) = make_float2(v[0], v[1]);
                        } else if (p.vec && p.out_bits == 6) {
                            *reinterpret_cast<float2*>(o) = make_float2(v[0], v[1]);
                            *reinterpret_cast<float2*>(o + 2) = make_float2(v[2], v[3]);
                            *reinterpret_cast<float2*>(o + 4) = make_float2(v[4], v[5]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (j < p.out_bits) o[j] = v[j];
                        }
                    };
                    if (p.llr_grid) store_bits(p.llr_grid + grow * p.out_bits);
                    if (p.llr && d_row >= 0) store_bits(p.llr + (size_t(bu) * p.n_data + d_row) * p.out_bits);
                    if (p.llr_aerial) {
                        const int bb = bu / p.U, uu = bu - bb * p.U;
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (j < p.out_bits)
                                p.llr_aerial[((size_t(bb) * p.out_bits + j) * p.U + uu) * p.rows_per_bu + prow] = -v[j];
                    }
                } else if (p.h_ref) {
                    float* o = p.h_ref + grow * p.N2;
                    if (p.vec && p.N2 == 8) {
                        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
                        *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (j < p.N2) o[j] = v[j];
                    }
                }
            }
        }
        // no barrier here: the next hidden tile is written only after every thread has seen this tile's second GEMM
        // complete, and the next second GEMM is issued behind the barrier that follows the hidden epilogue
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

}  // namespace nrx

// nrx_stack_tm.cuh — TMEM-resident UpdateState stack kernel (sm_100a), execution plan 4.
//
// Same arithmetic as nrx_stack_kernel<kStackUpdate> (three SeparableConv2D layers of UpdateState,
// utils/neural_rx.py:210-270, residual of :266), different data flow: the hidden activations never
// touch shared memory.
//
//   * An MMA row (TMEM lane) is one resource element (f, t).  A warp owns the 32 lanes of its TMEM
//     quadrant = two *sequences* (16 lanes each: lanes 0-6 the even, 8-14 the odd OFDM symbols of one
//     subcarrier; lanes 7 and 15 idle).  A CTA therefore advances 8 sequences in lock step; a
//     sequence walks down its own range of subcarriers (a *job*) one subcarrier per step.
//   * With the 16x256b fragment of tcgen05.ld / tcgen05.st (tools/tmem_frag_probe.cu) thread T owns,
//     in both sequences of its warp, the adjacent symbols (2i, 2i+1), i = T/4, and per 16-channel
//     group the four consecutive channels 4g .. 4g+3, g = T%4.
//   * Depthwise 3x3 in "scatter" form: the thread that holds input row f of its pixels adds the
//     row's contribution to the pending outputs f-1 (which completes), f and f+1, so the +-1
//     subcarrier halo lives in registers (2 x 64 half2 pending sums per thread) and the +-1 symbol
//     halo is one shuffle per pixel-column from lane T-4 / T+4.  Taps and biases sit in shared
//     memory (copied from the kernel parameters) and are read as 8-byte broadcast loads.
//   * The completed depthwise row is written to tensor memory (tcgen05.st) and is the A operand
//     of tcgen05.mma (A in TMEM, B = pointwise weights resident in shared memory, D in TMEM).
//     The next layer's warps read D with tcgen05.ld in exactly the fragment they need (output
//     channels are stored in fragment order, tm_phys_col): bias + ReLU + fp16 in registers, straight
//     into their own depthwise pass.
//   * Warp roles: warps 0-3 layer 1 (input rows by tensor-map TMA, 128B swizzle -> conflict-free
//     fragment reads); warps 4-7 layer 2; warps 8-11 layer 3 + output epilogue (bias, residual, TMA
//     store).  The three layers work on consecutive steps concurrently; mbarriers (A full / D full /
//     D empty per layer) are the only cross-warp synchronisation.  All roles execute one
//     branch-free pass body (predicated loads, masks): three bodies do not fit the instruction cache.
//
// Tensor memory (512 columns): D1 [0,128) D2 [128,256) D3 [256,320) A1 [320,384) A2 [384,448)
// A3 [448,512).  Shared memory: pointwise weights 80 KB | input ring 3 x 32 KB | residual 16 KB |
// output staging 16 KB.
//
// The accumulation order of every fp16 / fp32 sum equals the one of nrx_stack_kernel, so the two
// plans agree bit for bit (tests/test_gpu_parity.py::test_fused_equals_layerwise, test_tm_plan_job_split).
// Measured slower than plan 1 (0.545 vs 0.45 ms per launch): DESIGN.md 4.8, ROADMAP.md 6.
#pragma once
#include <cuda.h>

#include "nrx_stack.cuh"

namespace nrx {

constexpr int kTmThreads = 384;
constexpr int kTmStages = 3;
constexpr int kTmSlot = 2048;        // shared-memory slot of one 14 x 128 B box (rows 14, 15 stay zero)
constexpr int kTmFill = 6;           // pipeline fill: steps before a sequence emits its first output row
constexpr int kTmSeqs = 8;           // sequences per CTA
constexpr int kTmBoxBytes = kT * 128;

struct alignas(64) TmParams {
    CUtensorMap map_a, map_s, map_o;  // [planes][F*14][64] fp16: aggregated messages, state in, state out
    uint32_t tap[3][64][9];           // depthwise taps as half2 per channel pair: [layer][K/2][3x3]
    float bias[320];                  // [128 | 128 | 64]
    const uint8_t* wblob;             // StackSmem<kStackUpdate> image (pointwise B images first)
    int F, jobs_per_plane, num_jobs, num_items, steps_per_item;
};

struct TmSmem {
    static constexpr int kW = 81920;                       // pw1 32 KB | pw2 32 KB | pw3 16 KB
    static constexpr int oZ = kW;
    static constexpr int kStage = 4 * 2 * 2 * kTmSlot;     // [warp][sequence half][a | s]
    static constexpr int oRes = oZ + kTmStages * kStage;
    static constexpr int oOut = oRes + 4 * 2 * kTmSlot;
    static constexpr int kUsed = oOut + 4 * 2 * kTmSlot;
    static constexpr int kTotal = kUsed + 1024;            // + base alignment slack
};

struct TmSeq {
    int plane, f0, f1, j;
    bool valid;
};

// sequence `sigma` of the CTA's pass n: which job, which step inside the job
__device__ __forceinline__ TmSeq tm_seq_at(const TmParams& P, int n, int sigma) {
    TmSeq s;
    const int il = n / P.steps_per_item;
    s.j = n - il * P.steps_per_item;
    const int item = int(blockIdx.x) + il * int(gridDim.x);
    const int job = item * kTmSeqs + sigma;
    s.valid = job < P.num_jobs;
    s.plane = job / P.jobs_per_plane;
    const int i = job - s.plane * P.jobs_per_plane;
    s.f0 = (i * P.F) / P.jobs_per_plane;
    s.f1 = ((i + 1) * P.F) / P.jobs_per_plane;
    return s;
}

// Channel order.  A thread of the 16x256b fragment owns, per 16-column group c of an accumulator,
// the columns 16c + 8e + 2g + d (e, d in {0,1}; g = T%4).  The pointwise weight images and biases of
// plan 4 store output channel  16c + 4g + 2e + d  in that physical column (tm_phys_col), so the
// four values a thread holds are four CONSECUTIVE channels = the A-operand columns 8c + 2g + e it
// may write with the same fragment shape: every layer keeps the natural K order.
__host__ __device__ constexpr int tm_phys_col(int n) { return (n & ~15) | ((n & 2) << 2) | ((n >> 1) & 6) | (n & 1); }

// Two channel pairs (e = 0, 1) x two symbols (px) x two sequences (h) of one input row: contribution to
// the three pending outputs of each.  pa = output row r (completes: returned in o), pb = output row
// r+1; afterwards pa/pb are the pending sums of rows r+1/r+2.  tp = the thread's taps of the
// chunk, [tap 0..8][g][e] (tp already points at g).  A thread's symbols are adjacent (2i, 2i+1): only
// the outer neighbours come from other lanes (pair i-1 / i+1 = lane T-4 / T+4).  Tap order and
// accumulation order per output are those of dw_slide (nrx_stack.cuh).
struct TmIn {            // the three inputs of every output of a chunk: symbols t-1, t, t+1; index = (h * 2 + px) * 2 + e
    __half2 v[3][8];
};
// phase 1: neighbour exchange + the three taps that complete output row r (returned in o)
__device__ __forceinline__ void tm_chunk_complete(const uint32_t (&x)[8], const uint32_t* pa, const uint32_t* tp, int lane_l,
                                                  int lane_r, TmIn& in, uint32_t (&o)[8]) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int ee = 0; ee < 2; ++ee) {
            const int i0 = (h * 2 + 0) * 2 + ee, i1 = (h * 2 + 1) * 2 + ee;
            const __half2 x0 = u2h(x[i0]), x1 = u2h(x[i1]);
            in.v[0][i0] = u2h(__shfl_sync(0xffffffffu, x[i1], lane_l));    // symbol 2i-1
            in.v[1][i0] = x0;
            in.v[2][i0] = x1;
            in.v[0][i1] = x0;
            in.v[1][i1] = x1;
            in.v[2][i1] = u2h(__shfl_sync(0xffffffffu, x[i0], lane_r));    // symbol 2i+2
        }
    __half2 e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) e[i] = u2h(pa[i]);
#pragma unroll
    for (int tap = 6; tap < 9; ++tap) {
        const uint2 tv = *reinterpret_cast<const uint2*>(tp + 8 * tap);
#pragma unroll
        for (int i = 0; i < 8; ++i) e[i] = __hfma2(in.v[tap % 3][i], u2h((i & 1) ? tv.y : tv.x), e[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = h2u(e[i]);
}
// phase 2: the pending sums of rows r+1 (taps 3, 4, 5 on top of pb) and r+2 (taps 0, 1, 2 from zero)
__device__ __forceinline__ void tm_chunk_pending(const TmIn& in, uint32_t* pa, uint32_t* pb, const uint32_t* tp) {
    __half2 a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = u2h(pb[i]);
        b[i] = __float2half2_rn(0.f);
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const int tap = k < 3 ? 3 + k : k - 3;
        const uint2 tv = *reinterpret_cast<const uint2*>(tp + 8 * tap);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const __half2 w = u2h((i & 1) ? tv.y : tv.x);
            if (k < 3) a[i] = __hfma2(in.v[tap % 3][i], w, a[i]);
            else b[i] = __hfma2(in.v[tap % 3][i], w, b[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        pa[i] = h2u(a[i]);
        pb[i] = h2u(b[i]);
    }
}

#ifdef NRX_PHASE_TIMING
// debug build (tools/tm_phase_timing.py): cycle accounting of CTA 0, lane 0 of the first warp of each role:
// [3r] input waits, [3r+2] depthwise pass (incl. GEMM issue), [9+r] GEMM issue -> complete, [12+r] wait for
// the other warps / the consumer before the GEMM issue, [15+r] epilogue, [20] kernel total, [21] steps
__device__ unsigned long long g_tm_cycles[32];
#define NRX_TM_TICK(i)                                                     \
    do {                                                                   \
        if (tm_timer) {                                                    \
            const long long now_ = clock64();                              \
            tm_acc[i] += (unsigned long long)(now_ - tm_last);             \
            tm_last = now_;                                                \
        }                                                                  \
    } while (0)
#ifdef NRX_TM_FINE                                         // finer: inside the pass body (perturbs the schedule)
#define NRX_TM_TICK2(i)                                                    \
    do {                                                                   \
        if (tm_timer) {                                                    \
            const long long now_ = clock64();                              \
            tm_fine[i] += (unsigned long long)(now_ - tm_last2);           \
            tm_last2 = now_;                                               \
        }                                                                  \
    } while (0)
#else
#define NRX_TM_TICK2(i) do { } while (0)
#endif
#else
#define NRX_TM_TICK(i) do { } while (0)
#define NRX_TM_TICK2(i) do { } while (0)
#endif

// predicated 8-byte shared-memory load (role 0 only reads the input rows)
__device__ __forceinline__ uint2 lds64_if(const void* p, uint32_t pred) {
    uint2 z = make_uint2(0u, 0u);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t@p ld.shared.v2.b32 {%0, %1}, [%2];\n\t}"
                 : "+r"(z.x), "+r"(z.y)
                 : "r"(smem_u32(p)), "r"(pred));
    return z;
}

__global__ void __launch_bounds__(kTmThreads, 1) nrx_stack_tm_kernel(const __grid_constant__ TmParams P) {
    using S = TmSmem;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sW = smem;
    uint8_t* sZ = smem + S::oZ;
    uint8_t* sRes = smem + S::oRes;
    uint8_t* sOut = smem + S::oOut;
    __shared__ uint64_t bar_w, bar_afull[3], bar_dfull[3], bar_dempty[3], bar_z[4][kTmStages], bar_res[4];
    __shared__ uint32_t tmem_slot;
    // taps [layer][chunk of 8 channel pairs][tap][g][e] and biases (physical column order)
    __shared__ __align__(16) uint32_t sTap[3 * 64 * 9];
    __shared__ __align__(16) float sBias[320];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int role = warp >> 2, q = warp & 3;
    const int pi = lane >> 2, g = lane & 3;                // symbol pair (2 pi, 2 pi + 1; pair 7 idle), column group
    const int lane_l = (lane + 28) & 31, lane_r = (lane + 4) & 31;

    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    if (tid == 0) {
        mbar_init(&bar_w, 1);
        for (int i = 0; i < 3; ++i) {
            mbar_init(&bar_afull[i], 4);
            mbar_init(&bar_dfull[i], 1);
            mbar_init(&bar_dempty[i], 4);
        }
        for (int i = 0; i < 4; ++i) {
            for (int s = 0; s < kTmStages; ++s) mbar_init(&bar_z[i][s], 2);
            mbar_init(&bar_res[i], 2);
        }
        fence_mbar_init();
    }
    for (int i = tid; i < 3 * 64 * 9; i += kTmThreads) {     // [layer][chunk][tap][g][e] <- tap[layer][8 c + 2 g + e][tap]
        const int ge = i & 7, k = (i >> 3) % 9, lc = i / 72;
        sTap[i] = P.tap[lc >> 3][8 * (lc & 7) + ge][k];
    }
    for (int i = tid; i < 320; i += kTmThreads) sBias[i] = P.bias[i];
    // rows 14 and 15 of every slot are never written by the TMA: the idle symbol pair reads zeros there
    for (int i = tid; i < (S::kUsed - S::oZ) / 16; i += kTmThreads) st_shared_v4(smem + S::oZ + i * 16, make_uint4(0, 0, 0, 0));
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = tmem_slot;
    const uint32_t tlane = tbase + (uint32_t(32 * q) << 16);         // + (16 h << 16) for the second sequence
    if (tid == 0) {
        mbar_arrive_expect_tx(&bar_w, S::kW);
        bulk_g2s(sW, P.wblob, S::kW, &bar_w);
    }

    const int n_my = int(blockIdx.x) < P.num_items ? (P.num_items - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x) : 0;
    const int n_total = n_my * P.steps_per_item;
    // byte offsets of this thread's two symbols / four channels inside a 14 x 128 B box (128-byte swizzle):
    // channels 16 c + 4 g + {0..3}  ->  16-byte chunk 2 c + g/2, half g%2
    const int row0 = (2 * pi) * 128, row1 = (2 * pi + 1) * 128;
    const int sw0 = (2 * pi) & 7, sw1 = (2 * pi + 1) & 7;
    auto box_off = [&](int px, int c) {
        return (px ? row1 : row0) + ((((2 * c) + (g >> 1)) ^ (px ? sw1 : sw0)) << 4) + ((g & 1) << 3);
    };

#ifdef NRX_PHASE_TIMING
    const bool tm_timer = blockIdx.x == 0 && q == 0 && lane == 0;
    unsigned long long tm_acc[4] = {0, 0, 0, 0};          // waits, -, pass, epilogue
    unsigned long long tm_mma = 0, tm_mma_wait = 0;
    unsigned long long tm_fine[4] = {0, 0, 0, 0};         // accumulator load + wait, convert, shuffles + depthwise, A store
    long long tm_last2 = 0;
    long long tm_last = clock64();
    const long long tm_start = tm_last;
#endif
    // ---- input fetch (role 0), residual fetch and output epilogue (role 2); lanes 0 and 1 drive the
    //      TMA of the warp's two sequences ---------------------------------------------------------
    auto issue_z = [&](int m) {
        if (lane < 2) {
            const TmSeq s = tm_seq_at(P, m, 2 * q + lane);
            const int fz = s.f0 - 3 + s.j;
            const int st = m % kTmStages;
            uint8_t* slot = sZ + st * S::kStage + q * (4 * kTmSlot) + lane * (2 * kTmSlot);
            mbar_arrive_expect_tx(&bar_z[q][st], 2 * kTmBoxBytes);
            tma_load_3d(slot, &P.map_a, 0, fz * kT, s.valid ? s.plane : -1, &bar_z[q][st]);
            tma_load_3d(slot + kTmSlot, &P.map_s, 0, fz * kT, s.valid ? s.plane : -1, &bar_z[q][st]);
        }
    };
    auto issue_res = [&](int m) {
        if (lane < 2) {
            const TmSeq s = tm_seq_at(P, m, 2 * q + lane);
            const int fo = s.f0 - kTmFill + s.j;
            mbar_arrive_expect_tx(&bar_res[q], kTmBoxBytes);
            tma_load_3d(sRes + q * (2 * kTmSlot) + lane * kTmSlot, &P.map_s, 0, fo * kT, s.valid ? s.plane : -1, &bar_res[q]);
        }
    };
    // bias + residual + fp16 + TMA store of output pass m (accumulator D3)
    auto epilogue = [&](int m) {
        mbar_wait(&bar_dfull[2], m & 1);
        tc_fence_after_sync();
        mbar_wait(&bar_res[q], m & 1);
        if (lane < 2) bulk_wait_read_all();            // the previous store has read the staging rows
        __syncwarp();
#pragma unroll 1
        for (int hc = 0; hc < 4; ++hc) {                   // two 16-column groups per iteration: both TMEM loads in flight
            const int h = hc >> 1, c0 = (hc & 1) * 2;
            uint32_t r0[8], r1[8];
            tmem_ld_16x256b_x2(tlane + (uint32_t(16 * h) << 16) + 256 + 16 * c0, r0);
            tmem_ld_16x256b_x2(tlane + (uint32_t(16 * h) << 16) + 256 + 16 * (c0 + 1), r1);
            const uint8_t* rs = sRes + q * (2 * kTmSlot) + h * kTmSlot;
            uint8_t* os = sOut + q * (2 * kTmSlot) + h * kTmSlot;
            uint2 old[2][2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                old[j][0] = lds64(rs + box_off(0, c0 + j));
                old[j][1] = lds64(rs + box_off(1, c0 + j));
            }
            tmem_ld_wait16(r0, r1);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint32_t(&r)[8] = j ? r1 : r0;
                const int c = c0 + j;
                const float2 bA = *reinterpret_cast<const float2*>(sBias + 256 + 16 * c + 2 * g);
                const float2 bB = *reinterpret_cast<const float2*>(sBias + 256 + 16 * c + 8 + 2 * g);
                const float2 o00 = __half22float2(u2h(old[j][0].x)), o01 = __half22float2(u2h(old[j][0].y));
                const float2 o10 = __half22float2(u2h(old[j][1].x)), o11 = __half22float2(u2h(old[j][1].y));
                uint2 w0, w1;                              // s <- s + update (:266)
                w0.x = pack_half2((__uint_as_float(r[0]) + bA.x) + o00.x, (__uint_as_float(r[1]) + bA.y) + o00.y);
                w0.y = pack_half2((__uint_as_float(r[4]) + bB.x) + o01.x, (__uint_as_float(r[5]) + bB.y) + o01.y);
                w1.x = pack_half2((__uint_as_float(r[2]) + bA.x) + o10.x, (__uint_as_float(r[3]) + bA.y) + o10.y);
                w1.y = pack_half2((__uint_as_float(r[6]) + bB.x) + o11.x, (__uint_as_float(r[7]) + bB.y) + o11.y);
                sts64(os + box_off(0, c), w0);
                sts64(os + box_off(1, c), w1);
            }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane < 2) {
            const TmSeq s = tm_seq_at(P, m, 2 * q + lane);
            const int fo = s.f0 - kTmFill + s.j;
            if (s.valid && s.j >= kTmFill && fo < s.f1) tma_store_3d(&P.map_o, 0, fo * kT, s.plane, sOut + q * (2 * kTmSlot) + lane * kTmSlot);
            bulk_commit();
        }
    };

    // the elected lane of a layer's first warp: all four warps have stored their A rows and the
    // consumer has drained the previous accumulator -> issue the layer's GEMM
    auto issue_mma = [&](int L, int n) {
#ifdef NRX_PHASE_TIMING
        const long long t_a = clock64();
#endif
        mbar_wait(&bar_afull[L], n & 1);
        if (n >= 1 && L < 2) mbar_wait(&bar_dempty[L], (n - 1) & 1);   // D3 is drained by this role's own epilogue
        tc_fence_after_sync();
#ifdef NRX_PHASE_TIMING
        const long long t_b = clock64();
        if (blockIdx.x == 0) tm_mma_wait += (unsigned long long)(t_b - t_a);
#endif
        const uint32_t dcol = L == 0 ? 0u : L == 1 ? 128u : 256u;
        const uint32_t acol = 320u + 64u * uint32_t(L);
        const uint32_t woff = L == 0 ? 0u : L == 1 ? 32768u : 65536u;
        const uint32_t nn = L == 2 ? 64u : 128u;
        umma_gemm_k_ts(tbase + dcol, tbase + acol, smem_u32(sW + woff), nn * 128u, 128, umma_idesc_f16(128, nn));
        umma_commit(&bar_dfull[L]);
#ifdef NRX_PHASE_TIMING
        if (blockIdx.x == 0) {                             // perturbs the schedule: the issuing warp waits for its GEMM
            mbar_wait(&bar_dfull[L], n & 1);
            tm_mma += (unsigned long long)(clock64() - t_b);
        }
#endif
    };

    uint32_t pa[64], pb[64];                               // [chunk 8][h 2][px 2][e 2]
#pragma unroll
    for (int c = 0; c < 64; ++c) pa[c] = pb[c] = 0u;

    // per-role constants of the common pass body (L = role: layer whose depthwise this warp runs)
    const int L = role;
    const uint32_t* tapL = sTap + L * (64 * 9) + 2 * g;
    const float* biasL = sBias + (L == 2 ? 128 : 0) + 2 * g;   // bias of the layer that produced this warp's input
    const uint32_t dsrc = tlane + (L == 2 ? 128u : 0u);    // accumulator this warp reads (role 0: values unused)
    const uint32_t adst = tlane + 320u + 64u * uint32_t(L);
    uint64_t* const bar_in = &bar_dfull[L == 0 ? 0 : L - 1];
    uint64_t* const bar_in_empty = &bar_dempty[L == 0 ? 0 : L - 1];
    const uint32_t m_lds = role == 0 ? 1u : 0u;            // role 0 takes its input row from shared memory,
    const int drain_lane = role == 0 ? 32 : 0;             // the others from the previous layer's accumulator

    if (role == 0) {
        if (n_total > 0) issue_z(0);
        if (n_total > 1) issue_z(1);
    }
    mbar_wait(&bar_w, 0);

    // Every warp runs the SAME branch-free instruction stream for the pass body (three specialised
    // bodies exceed the instruction cache): both input paths are executed and the role selects by
    // masks / predicates; the role-dependent barriers sit outside the unrolled body.
    for (int n = 0; n < n_total; ++n) {
        const uint8_t* zb = sZ;
        uint32_t m_acc[2] = {0u, 0u};
        uint64_t* wait_bar = bar_in;
        uint32_t wait_parity = n & 1;
        if (role == 0) {
            if (n + 2 < n_total) issue_z(n + 2);
            const int st = n % kTmStages;
            wait_bar = &bar_z[q][st];
            wait_parity = (n / kTmStages) & 1;
            zb = sZ + st * S::kStage + q * (4 * kTmSlot);
        } else {
            if (role == 2 && n >= 1) issue_res(n - 1);     // residual rows of the output pass this step finishes
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const TmSeq s = tm_seq_at(P, n, 2 * q + h);
                const int f_in = s.f0 - 3 - L + s.j;       // layer 2 reads H1[f0-4+j], layer 3 reads H2[f0-5+j]
                // rows outside the grid are the zero padding of the next layer; the idle pair holds zeros
                m_acc[h] = (pi < 7 && s.valid && f_in >= 0 && f_in < P.F) ? 0xffffffffu : 0u;
            }
        }
        NRX_TM_TICK(2);                                    // bookkeeping between passes counts as pass time
        mbar_wait_sleep(wait_bar, wait_parity);
        tc_fence_after_sync();
        NRX_TM_TICK(0);
#ifdef NRX_TM_FINE
        tm_last2 = clock64();
#endif
        if (n >= 1) {                                      // the previous GEMM has consumed this layer's A tile
            mbar_wait(&bar_dfull[L], (n - 1) & 1);
            tc_fence_after_sync();
        }
        uint32_t d0[8], d1[8];                             // accumulator fragments of the two sequences
        tmem_ld_16x256b_x2(dsrc, d0);
        tmem_ld_16x256b_x2(dsrc + (16u << 16), d1);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float2 bA = *reinterpret_cast<const float2*>(biasL + 16 * c);
            const float2 bB = *reinterpret_cast<const float2*>(biasL + 16 * c + 8);
            // input row from shared memory (role 0): channels 16 c + 4 g + {0..3} of `a` (c < 4) or `s | pe`
            const uint8_t* zc = zb + (c < 4 ? 0 : kTmSlot);
            const uint2 z00 = lds64_if(zc + box_off(0, c & 3), m_lds), z01 = lds64_if(zc + box_off(1, c & 3), m_lds);
            const uint2 z10 = lds64_if(zc + 2 * kTmSlot + box_off(0, c & 3), m_lds), z11 = lds64_if(zc + 2 * kTmSlot + box_off(1, c & 3), m_lds);
            tmem_ld_wait16(d0, d1);
            NRX_TM_TICK2(0);
            if (c == 7) {                                  // accumulator drained
                tc_fence_before_sync();
                __syncwarp();
                if (lane == drain_lane) mbar_arrive(bar_in_empty);
            }
            uint32_t x[8];                                 // [(h * 2 + px) * 2 + e]: bias + ReLU + fp16 | input row
            x[0] = (pack_relu_half2(__uint_as_float(d0[0]) + bA.x, __uint_as_float(d0[1]) + bA.y) & m_acc[0]) | z00.x;
            x[1] = (pack_relu_half2(__uint_as_float(d0[4]) + bB.x, __uint_as_float(d0[5]) + bB.y) & m_acc[0]) | z00.y;
            x[2] = (pack_relu_half2(__uint_as_float(d0[2]) + bA.x, __uint_as_float(d0[3]) + bA.y) & m_acc[0]) | z01.x;
            x[3] = (pack_relu_half2(__uint_as_float(d0[6]) + bB.x, __uint_as_float(d0[7]) + bB.y) & m_acc[0]) | z01.y;
            x[4] = (pack_relu_half2(__uint_as_float(d1[0]) + bA.x, __uint_as_float(d1[1]) + bA.y) & m_acc[1]) | z10.x;
            x[5] = (pack_relu_half2(__uint_as_float(d1[4]) + bB.x, __uint_as_float(d1[5]) + bB.y) & m_acc[1]) | z10.y;
            x[6] = (pack_relu_half2(__uint_as_float(d1[2]) + bA.x, __uint_as_float(d1[3]) + bA.y) & m_acc[1]) | z11.x;
            x[7] = (pack_relu_half2(__uint_as_float(d1[6]) + bB.x, __uint_as_float(d1[7]) + bB.y) & m_acc[1]) | z11.y;
            NRX_TM_TICK2(1);
            TmIn in;
            uint32_t o[8];
            tm_chunk_complete(x, pa + 8 * c, tapL + c * 72, lane_l, lane_r, in, o);
            tmem_st_16x256b_x1(adst + 8 * c, o[0], o[1], o[2], o[3]);
            tmem_st_16x256b_x1(adst + (16u << 16) + 8 * c, o[4], o[5], o[6], o[7]);
            if (c < 7) {                                   // next chunk's accumulator fragments arrive during phase 2
                tmem_ld_16x256b_x2(dsrc + 16 * (c + 1), d0);
                tmem_ld_16x256b_x2(dsrc + (16u << 16) + 16 * (c + 1), d1);
            }
            tm_chunk_pending(in, pa + 8 * c, pb + 8 * c, tapL + c * 72);
            NRX_TM_TICK2(2);
        }
        NRX_TM_TICK(2);
        if (role == 2 && n >= 1) epilogue(n - 1);
        NRX_TM_TICK(3);
        tmem_st_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&bar_afull[L]);
            if (q == 0) issue_mma(L, n);
        }
        __syncwarp();
    }
    if (role == 2 && n_total > 0) {
        issue_res(n_total - 1);
        epilogue(n_total - 1);
        if (lane < 2) bulk_wait_all();
    }
#ifdef NRX_PHASE_TIMING
    if (tm_timer) {
        for (int i = 0; i < 3; ++i) atomicAdd(&g_tm_cycles[3 * role + i], tm_acc[i]);
        atomicAdd(&g_tm_cycles[9 + role], tm_mma);
        atomicAdd(&g_tm_cycles[12 + role], tm_mma_wait);
        atomicAdd(&g_tm_cycles[15 + role], tm_acc[3]);
        if (role == 0) atomicAdd(&g_tm_cycles[20], (unsigned long long)(clock64() - tm_start));
        if (role == 0) atomicAdd(&g_tm_cycles[21], (unsigned long long)n_total);
        if (role == 1) for (int i = 0; i < 4; ++i) atomicAdd(&g_tm_cycles[22 + i], tm_fine[i]);
    }
#endif
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

// Jobs per plane that minimise  waves x steps-per-item  on `num_sms` persistent CTAs.
inline int tm_choose_jobs(int planes, int F, int num_sms) {
    int best_j = 1;
    long long best = -1;
    for (int j = 1; j <= F && j <= 1024; ++j) {
        const long long items = ((long long)planes * j + kTmSeqs - 1) / kTmSeqs;
        const long long waves = (items + num_sms - 1) / num_sms;
        const long long steps = (F + j - 1) / j + kTmFill;
        const long long cost = waves * steps;
        if (best < 0 || cost < best) { best = cost; best_j = j; }
    }
    return best_j;
}

}  // namespace nrx

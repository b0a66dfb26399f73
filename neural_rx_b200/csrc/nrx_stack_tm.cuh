// nrx_stack_tm.cuh — TMEM-resident UpdateState stack kernel (sm_100a), execution plan 4.
//
// Same arithmetic as nrx_stack_kernel<kStackUpdate> (three SeparableConv2D layers of UpdateState,
// utils/neural_rx.py:210-270, residual of :266), different data flow: the hidden activations never
// touch shared memory.
//
//   * An MMA row (TMEM lane) is one resource element (f, t).  A warp owns the 32 lanes of its TMEM
//     quadrant = two *sequences* (lanes 0-13 and 16-29: the 14 OFDM symbols of one subcarrier each;
//     lanes 14, 15, 30, 31 idle).  A CTA therefore advances 8 sequences in lock step; a sequence
//     walks down its own range of subcarriers (a *job*) one subcarrier per step.
//   * Depthwise 3x3 in "scatter" form: the thread that holds input row f of its pixel adds the
//     row's contribution to the pending outputs f-1 (which completes), f and f+1, so the +-1
//     subcarrier halo lives in registers (2 x 64 half2 pending sums per thread) and the +-1 symbol
//     halo comes from the neighbouring lanes by warp shuffles.  The nine taps and the biases are
//     kernel parameters (constant bank): every lane of a warp works on the same channel.
//   * The completed depthwise row is written to tensor memory (tcgen05.st) and is the A operand
//     of tcgen05.mma (A in TMEM, B = pointwise weights resident in shared memory, D in TMEM).
//     The next layer's warps read D with tcgen05.ld in exactly the row-per-lane mapping they need:
//     bias + ReLU + fp16 in registers, straight into their own depthwise pass.
//   * Warp roles: warps 0-3 layer 1 (input rows by tensor-map TMA, 128B swizzle -> conflict-free
//     row-per-lane reads) + output epilogue (bias, residual, TMA store); warps 4-7 layer 2;
//     warps 8-11 layer 3.  The three layers work on consecutive steps concurrently; mbarriers
//     (A full / D full / D empty per layer) are the only cross-warp synchronisation.
//
// Tensor memory (512 columns): D1 [0,128) D2 [128,256) D3 [256,320) A1 [320,384) A2 [384,448)
// A3 [448,512).  Shared memory: pointwise weights 80 KB | input ring 3 x 32 KB | residual 16 KB |
// output staging 16 KB.
//
// The accumulation order of every fp16 / fp32 sum equals the one of nrx_stack_kernel, so the two
// plans are expected to agree bit for bit (tests/test_gpu_parity.py::test_tm_plan_equals_fused).
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

// Four channel pairs (one 16-byte chunk) of one input row: contribution to the three pending outputs
// of each pair.  pa = output row r (completes: returned in o), pb = output row r+1; afterwards
// pa/pb are the pending sums of rows r+1/r+2.  tp = taps of the chunk, [tap 0..8][pair 0..3].
// Tap order and accumulation order per output are those of dw_slide (nrx_stack.cuh).
__device__ __forceinline__ void tm_chunk(const uint32_t (&x)[4], uint32_t* pa, uint32_t* pb, const uint32_t* tp, int lane_l,
                                         int lane_r, uint32_t (&o)[4]) {
    __half2 hx[4], hl[4], hr[4], e[4], a[4], b[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        hx[w] = u2h(x[w]);
        hl[w] = u2h(__shfl_sync(0xffffffffu, x[w], lane_l));
        hr[w] = u2h(__shfl_sync(0xffffffffu, x[w], lane_r));
        e[w] = u2h(pa[w]);
        a[w] = u2h(pb[w]);
        b[w] = __float2half2_rn(0.f);
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int tap = k < 3 ? 6 + k : k < 6 ? k : k - 6;       // 6,7,8 (row r) | 3,4,5 (row r+1) | 0,1,2 (row r+2)
        const uint4 tv = *reinterpret_cast<const uint4*>(tp + 4 * tap);
        const uint32_t tw[4] = {tv.x, tv.y, tv.z, tv.w};
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const __half2 in = (tap % 3) == 0 ? hl[w] : (tap % 3) == 1 ? hx[w] : hr[w];
            if (k < 3) e[w] = __hfma2(in, u2h(tw[w]), e[w]);
            else if (k < 6) a[w] = __hfma2(in, u2h(tw[w]), a[w]);
            else b[w] = __hfma2(in, u2h(tw[w]), b[w]);
        }
    }
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        o[w] = h2u(e[w]);
        pa[w] = h2u(a[w]);
        pb[w] = h2u(b[w]);
    }
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
    // taps [layer][chunk of 4 channel pairs][9][4] and biases: read by every lane at the same address (broadcast)
    __shared__ __align__(16) uint32_t sTap[3 * 64 * 9];
    __shared__ __align__(16) float sBias[320];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int role = warp >> 2, q = warp & 3;
    const int h = lane >> 4, t = lane & 15;
    const int sigma = 2 * q + h;
    const int lane_l = (lane + 31) & 31, lane_r = (lane + 1) & 31;

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
    for (int i = tid; i < 3 * 64 * 9; i += kTmThreads) {     // [layer][chunk][tap][pair in chunk]
        const int w = i & 3, k = (i >> 2) % 9, lc = i / 36;
        sTap[i] = P.tap[lc >> 4][4 * (lc & 15) + w][k];
    }
    for (int i = tid; i < 320; i += kTmThreads) sBias[i] = P.bias[i];
    // rows 14 and 15 of every slot are never written by the TMA: the idle lanes read zeros there
    for (int i = tid; i < (S::kUsed - S::oZ) / 16; i += kTmThreads) st_shared_v4(smem + S::oZ + i * 16, make_uint4(0, 0, 0, 0));
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = tmem_slot;
    const uint32_t tlane = tbase + (uint32_t(32 * q) << 16);
    if (tid == 0) {
        mbar_arrive_expect_tx(&bar_w, S::kW);
        bulk_g2s(sW, P.wblob, S::kW, &bar_w);
    }

    const int n_my = int(blockIdx.x) < P.num_items ? (P.num_items - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x) : 0;
    const int n_total = n_my * P.steps_per_item;
    const int slot_off = q * (2 * kTmSlot) + h * kTmSlot;    // this sequence's slot inside sRes / sOut

    // ---- role 0 only: input / residual fetch and the output epilogue -----------------------------
    auto issue_z = [&](int m) {
        if (t == 0) {
            const TmSeq s = tm_seq_at(P, m, sigma);
            const int fz = s.f0 - 3 + s.j;
            const int st = m % kTmStages;
            uint8_t* slot = sZ + st * S::kStage + q * (4 * kTmSlot) + h * (2 * kTmSlot);
            mbar_arrive_expect_tx(&bar_z[q][st], 2 * kTmBoxBytes);
            tma_load_3d(slot, &P.map_a, 0, fz * kT, s.valid ? s.plane : -1, &bar_z[q][st]);
            tma_load_3d(slot + kTmSlot, &P.map_s, 0, fz * kT, s.valid ? s.plane : -1, &bar_z[q][st]);
        }
    };
    auto issue_res = [&](int m) {
        if (t == 0) {
            const TmSeq s = tm_seq_at(P, m, sigma);
            const int fo = s.f0 - kTmFill + s.j;
            mbar_arrive_expect_tx(&bar_res[q], kTmBoxBytes);
            tma_load_3d(sRes + slot_off, &P.map_s, 0, fo * kT, s.valid ? s.plane : -1, &bar_res[q]);
        }
    };
    // bias + residual + fp16 + TMA store of output pass m (accumulator D3)
    auto epilogue = [&](int m) {
        const TmSeq s = tm_seq_at(P, m, sigma);
        const int fo = s.f0 - kTmFill + s.j;
        const bool ok = s.valid && s.j >= kTmFill && fo < s.f1;
        mbar_wait(&bar_dfull[2], m & 1);
        tc_fence_after_sync();
        mbar_wait(&bar_res[q], m & 1);
        if (t == 0) bulk_wait_read_all();              // the previous store has read the staging rows
        __syncwarp();
        const uint8_t* rs = sRes + slot_off + t * 128;
        uint8_t* os = sOut + slot_off + t * 128;
        const int swz = t & 7;
#pragma unroll 1
        for (int c4 = 0; c4 < 4; ++c4) {
            float v[16];
            tmem_ld16(tlane + 256 + 16 * c4, v);
            uint4 old[2];
            old[0] = ld_shared_v4(rs + (((2 * c4) ^ swz) << 4));
            old[1] = ld_shared_v4(rs + (((2 * c4 + 1) ^ swz) << 4));
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const uint32_t ow[4] = {old[e].x, old[e].y, old[e].z, old[e].w};
                const float4 b0 = *reinterpret_cast<const float4*>(sBias + 256 + 16 * c4 + 8 * e);
                const float4 b1 = *reinterpret_cast<const float4*>(sBias + 256 + 16 * c4 + 8 * e + 4);
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                uint32_t pk[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 of = __half22float2(u2h(ow[i]));
                    const float o0 = v[8 * e + 2 * i] + bb[2 * i];
                    const float o1 = v[8 * e + 2 * i + 1] + bb[2 * i + 1];
                    pk[i] = pack_half2(o0 + of.x, o1 + of.y);     // s <- s + update (:266)
                }
                st_shared_v4(os + (((2 * c4 + e) ^ swz) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
            }
        }
        tc_fence_before_sync();                            // accumulator drained
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_dempty[2]);
        fence_proxy_async_smem();
        __syncwarp();
        if (t == 0) {
            if (ok) tma_store_3d(&P.map_o, 0, fo * kT, s.plane, sOut + slot_off);
            bulk_commit();
        }
    };

    // the elected lane of a layer's first warp: all four warps have stored their A rows and the
    // consumer has drained the previous accumulator -> issue the layer's GEMM
    auto issue_mma = [&](int L, int n) {
        mbar_wait(&bar_afull[L], n & 1);
        if (n >= 1) mbar_wait(&bar_dempty[L], (n - 1) & 1);
        tc_fence_after_sync();
        const uint32_t dcol = L == 0 ? 0u : L == 1 ? 128u : 256u;
        const uint32_t acol = 320u + 64u * uint32_t(L);
        const uint32_t woff = L == 0 ? 0u : L == 1 ? 32768u : 65536u;
        const uint32_t nn = L == 2 ? 64u : 128u;
        umma_gemm_k_ts(tbase + dcol, tbase + acol, smem_u32(sW + woff), nn * 128u, 128, umma_idesc_f16(128, nn));
        umma_commit(&bar_dfull[L]);
    };

    uint32_t pa[64], pb[64];
#pragma unroll
    for (int c = 0; c < 64; ++c) pa[c] = pb[c] = 0u;

    // per-role constants of the common pass body (L = role: layer whose depthwise this warp runs)
    const int L = role;
    const uint32_t* tapL = sTap + L * (64 * 9);
    const float* biasL = sBias + (L == 2 ? 128 : 0);        // bias of the layer that produced this warp's input
    const uint32_t dsrc = tlane + (L == 2 ? 128u : 0u);    // accumulator this warp reads (role 0: values unused)
    const uint32_t adst = tlane + 320u + 64u * uint32_t(L);
    uint64_t* const bar_in = &bar_dfull[L == 0 ? 0 : L - 1];
    uint64_t* const bar_in_empty = &bar_dempty[L == 0 ? 0 : L - 1];
    const uint32_t m_lds = role == 0 ? 0xffffffffu : 0u;   // role 0 takes its input row from shared memory,
    const int drain_lane = role == 0 ? 32 : 0;             // the others from the previous layer's accumulator
    const int swz = t & 7;

    if (role == 0) {
        if (n_total > 0) issue_z(0);
        if (n_total > 1) issue_z(1);
#pragma unroll
        for (int c = 0; c < 16; ++c) tmem_st4(adst + 4 * c, 0u, 0u, 0u, 0u);
        tmem_st_wait();
    }
    mbar_wait(&bar_w, 0);

    // Every warp runs the SAME branch-free instruction stream for the pass body (three specialised
    // bodies exceed the instruction cache): both input paths are executed and the role selects by
    // masks; the role-dependent barriers sit outside the unrolled body.
    const int n_end = role == 0 ? n_total + 3 : n_total;
    for (int n = 0; n < n_end; ++n) {
        const uint8_t* za = sZ + t * 128;
        uint32_t m_acc = 0u;
        uint64_t* wait_bar = bar_in;
        uint32_t wait_parity = n & 1;
        if (role == 0) {
            if (n >= 3) epilogue(n - 3);
            if (n >= 2 && n - 2 < n_total) issue_res(n - 2);
            if (n >= n_total) continue;
            if (n + 2 < n_total) issue_z(n + 2);
            const int st = n % kTmStages;
            wait_bar = &bar_z[q][st];
            wait_parity = (n / kTmStages) & 1;
            za = sZ + st * S::kStage + q * (4 * kTmSlot) + h * (2 * kTmSlot) + t * 128;
        } else {
            const TmSeq s = tm_seq_at(P, n, sigma);
            const int f_in = s.f0 - 3 - L + s.j;           // layer 2 reads H1[f0-4+j], layer 3 reads H2[f0-5+j]
            // rows outside the grid are the zero padding of the next layer; idle lanes hold zeros
            m_acc = (t < kT && s.valid && f_in >= 0 && f_in < P.F) ? 0xffffffffu : 0u;
        }
        mbar_wait(wait_bar, wait_parity);
        tc_fence_after_sync();
        uint32_t v[2][8];
        tmem_ld8(dsrc, v[0]);
#pragma unroll
        for (int ch = 0; ch < 16; ++ch) {
            const uint4 z = ld_shared_v4((ch < 8 ? za : za + kTmSlot) + (((ch & 7) ^ swz) << 4));
            tmem_ld_wait8(v[ch & 1]);
            if (ch < 15) {
                tmem_ld8(dsrc + 8 * (ch + 1), v[(ch + 1) & 1]);
            } else {                                       // accumulator drained
                tc_fence_before_sync();
                __syncwarp();
                if (lane == drain_lane) mbar_arrive(bar_in_empty);
            }
            const float4 b0 = *reinterpret_cast<const float4*>(biasL + 8 * ch);
            const float4 b1 = *reinterpret_cast<const float4*>(biasL + 8 * ch + 4);
            uint32_t xw[4];                                // bias + ReLU + fp16 | input row from shared memory
            xw[0] = (pack_relu_half2(__uint_as_float(v[ch & 1][0]) + b0.x, __uint_as_float(v[ch & 1][1]) + b0.y) & m_acc) | (z.x & m_lds);
            xw[1] = (pack_relu_half2(__uint_as_float(v[ch & 1][2]) + b0.z, __uint_as_float(v[ch & 1][3]) + b0.w) & m_acc) | (z.y & m_lds);
            xw[2] = (pack_relu_half2(__uint_as_float(v[ch & 1][4]) + b1.x, __uint_as_float(v[ch & 1][5]) + b1.y) & m_acc) | (z.z & m_lds);
            xw[3] = (pack_relu_half2(__uint_as_float(v[ch & 1][6]) + b1.z, __uint_as_float(v[ch & 1][7]) + b1.w) & m_acc) | (z.w & m_lds);
            uint32_t o[4];
            tm_chunk(xw, pa + 4 * ch, pb + 4 * ch, tapL + ch * 36, lane_l, lane_r, o);
            if (ch == 0 && n >= 1) {                       // the previous GEMM has consumed this layer's A tile
                mbar_wait(&bar_dfull[L], (n - 1) & 1);
                tc_fence_after_sync();
            }
            tmem_st4(adst + 4 * ch, o[0], o[1], o[2], o[3]);
        }
        tmem_st_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&bar_afull[L]);
            if (q == 0) issue_mma(L, n);
        }
        __syncwarp();
    }
    if (role == 0 && t == 0) bulk_wait_all();
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

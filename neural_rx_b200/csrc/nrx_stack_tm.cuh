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

// One channel pair of one input row: contribution to the three pending outputs.  pa = output row
// r (completes: returned), pb = output row r+1; afterwards pa/pb are the pending sums of rows r+1/r+2.
// Tap order and accumulation order are those of dw_slide (nrx_stack.cuh).
__device__ __forceinline__ uint32_t tm_col(const TmParams& P, const int L, const int C, uint32_t x, uint32_t& pa, uint32_t& pb, int lane_l, int lane_r) {
    const __half2 hx = u2h(x);
    const __half2 hl = u2h(__shfl_sync(0xffffffffu, x, lane_l));
    const __half2 hr = u2h(__shfl_sync(0xffffffffu, x, lane_r));
    __half2 e = __hfma2(hl, u2h(P.tap[L][C][6]), u2h(pa));
    e = __hfma2(hx, u2h(P.tap[L][C][7]), e);
    e = __hfma2(hr, u2h(P.tap[L][C][8]), e);
    __half2 a = __hfma2(hl, u2h(P.tap[L][C][3]), u2h(pb));
    a = __hfma2(hx, u2h(P.tap[L][C][4]), a);
    a = __hfma2(hr, u2h(P.tap[L][C][5]), a);
    __half2 b = __hfma2(hl, u2h(P.tap[L][C][0]), __float2half2_rn(0.f));
    b = __hfma2(hx, u2h(P.tap[L][C][1]), b);
    b = __hfma2(hr, u2h(P.tap[L][C][2]), b);
    pa = h2u(a);
    pb = h2u(b);
    return h2u(e);
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

    if (role == 0) {
        // =====================================================================================
        // layer 1 (input rows from shared memory) + output epilogue of layer 3
        // =====================================================================================
#pragma unroll
        for (int c = 0; c < 16; ++c) tmem_st4(tlane + 320 + 4 * c, 0u, 0u, 0u, 0u);   // K padding columns of A1 stay zero
        tmem_st_wait();
        const int slot_off = q * (2 * kTmSlot) + h * kTmSlot;    // this sequence's slot inside sRes / sOut

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
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                float v[16];
                tmem_ld16(tlane + 256 + 16 * c4, v);
                uint4 old[2];
                old[0] = ld_shared_v4(rs + (((2 * c4) ^ swz) << 4));
                old[1] = ld_shared_v4(rs + (((2 * c4 + 1) ^ swz) << 4));
                tmem_ld_wait();
                if (c4 == 3) {                             // accumulator drained
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_dempty[2]);
                }
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const uint32_t ow[4] = {old[e].x, old[e].y, old[e].z, old[e].w};
                    uint32_t pk[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 of = __half22float2(u2h(ow[i]));
                        const int col = 16 * c4 + 8 * e + 2 * i;
                        const float o0 = v[8 * e + 2 * i] + P.bias[256 + col];
                        const float o1 = v[8 * e + 2 * i + 1] + P.bias[256 + col + 1];
                        pk[i] = pack_half2(o0 + of.x, o1 + of.y);     // s <- s + update (:266)
                    }
                    st_shared_v4(os + (((2 * c4 + e) ^ swz) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (t == 0) {
                if (ok) tma_store_3d(&P.map_o, 0, fo * kT, s.plane, sOut + slot_off);
                bulk_commit();
            }
        };

        if (n_total > 0) issue_z(0);
        if (n_total > 1) issue_z(1);
        mbar_wait(&bar_w, 0);
        for (int n = 0; n < n_total + 3; ++n) {
            if (n >= 3) epilogue(n - 3);
            if (n >= 2 && n - 2 < n_total) issue_res(n - 2);
            if (n >= n_total) continue;
            if (n + 2 < n_total) issue_z(n + 2);
            // ---- layer-1 pass n ----
            const int st = n % kTmStages;
            mbar_wait(&bar_z[q][st], (n / kTmStages) & 1);
            const uint8_t* za = sZ + st * S::kStage + q * (4 * kTmSlot) + h * (2 * kTmSlot) + t * 128;
            const int swz = t & 7;
#pragma unroll
            for (int ch = 0; ch < 16; ++ch) {
                if (ch == 7) continue;                     // channels 56..63 of `a` are padding
                const uint8_t* src = (ch < 8 ? za : za + kTmSlot) + (((ch & 7) ^ swz) << 4);
                const uint4 v = ld_shared_v4(src);
                uint32_t o0, o1 = 0u, o2 = 0u, o3 = 0u;
                o0 = tm_col(P, 0, 4 * ch, v.x, pa[4 * ch], pb[4 * ch], lane_l, lane_r);
                if (ch != 15) {                            // channels 58..63 of `s | pe` are padding
                    o1 = tm_col(P, 0, 4 * ch + 1, v.y, pa[4 * ch + 1], pb[4 * ch + 1], lane_l, lane_r);
                    o2 = tm_col(P, 0, 4 * ch + 2, v.z, pa[4 * ch + 2], pb[4 * ch + 2], lane_l, lane_r);
                    o3 = tm_col(P, 0, 4 * ch + 3, v.w, pa[4 * ch + 3], pb[4 * ch + 3], lane_l, lane_r);
                }
                if (ch == 0 && n >= 1) {                   // the previous GEMM has consumed A1
                    mbar_wait(&bar_dfull[0], (n - 1) & 1);
                    tc_fence_after_sync();
                }
                tmem_st4(tlane + 320 + 4 * ch, o0, o1, o2, o3);
            }
            tmem_st_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&bar_afull[0]);
                if (q == 0) issue_mma(0, n);
            }
            __syncwarp();
        }
        if (t == 0) bulk_wait_all();
    } else {
        // =====================================================================================
        // layers 2 and 3: input = previous layer's accumulator (bias + ReLU + fp16 in registers)
        // =====================================================================================
        auto run = [&](auto ltag) {
            constexpr int L = decltype(ltag)::value;       // 1: consumes D1 -> A2, 2: consumes D2 -> A3
            constexpr uint32_t dsrc = L == 1 ? 0u : 128u, adst = 320u + 64u * L;
            constexpr int boff = L == 1 ? 0 : 128;
            for (int n = 0; n < n_total; ++n) {
                const TmSeq s = tm_seq_at(P, n, sigma);
                const int f_in = s.f0 - 3 - L + s.j;
                const uint32_t mask = (t < kT && s.valid && f_in >= 0 && f_in < P.F) ? 0xffffffffu : 0u;
                mbar_wait(&bar_dfull[L - 1], n & 1);
                tc_fence_after_sync();
                uint32_t v[2][8];
                tmem_ld8(tlane + dsrc, v[0]);
#pragma unroll
                for (int ch = 0; ch < 16; ++ch) {
                    tmem_ld_wait8(v[ch & 1]);
                    if (ch < 15) {
                        tmem_ld8(tlane + dsrc + 8 * (ch + 1), v[(ch + 1) & 1]);
                    } else {                               // accumulator drained
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bar_dempty[L - 1]);
                    }
                    uint32_t o[4];
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        const float v0 = __uint_as_float(v[ch & 1][2 * w]) + P.bias[boff + 8 * ch + 2 * w];
                        const float v1 = __uint_as_float(v[ch & 1][2 * w + 1]) + P.bias[boff + 8 * ch + 2 * w + 1];
                        const uint32_t x = pack_relu_half2(v0, v1) & mask;   // rows outside the grid are zero padding
                        if (w == 0) o[0] = tm_col(P, L, 4 * ch + 0, x, pa[4 * ch + 0], pb[4 * ch + 0], lane_l, lane_r);
                        if (w == 1) o[1] = tm_col(P, L, 4 * ch + 1, x, pa[4 * ch + 1], pb[4 * ch + 1], lane_l, lane_r);
                        if (w == 2) o[2] = tm_col(P, L, 4 * ch + 2, x, pa[4 * ch + 2], pb[4 * ch + 2], lane_l, lane_r);
                        if (w == 3) o[3] = tm_col(P, L, 4 * ch + 3, x, pa[4 * ch + 3], pb[4 * ch + 3], lane_l, lane_r);
                    }
                    if (ch == 0 && n >= 1) {               // the previous GEMM has consumed A_L
                        mbar_wait(&bar_dfull[L], (n - 1) & 1);
                        tc_fence_after_sync();
                    }
                    tmem_st4(tlane + adst + 4 * ch, o[0], o[1], o[2], o[3]);
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
        };
        if (role == 1) run(std::integral_constant<int, 1>{});
        else run(std::integral_constant<int, 2>{});
    }
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

// AggregateUserStates for two users (utils/neural_rx.py:135-207; every shipped configuration has max_num_tx = 2):
// persistent, warp-specialised version of nrx_agg_kernel<2>.  Same arithmetic in the same order — the outputs are
// bit-identical (tests/test_gpu_kernels.py) — but the roles run concurrently on different tiles:
//
//   producer   (1 warp)  TMA loads of the two users' [128 rows x 64 ch] state tiles into a ring of kAggStages stages.
//                        A tile keeps its stage for its whole life: state tiles -> hidden tiles -> staged messages.
//   tensor     (1 warp)  first GEMM (state -> hidden, K = 64, N = 64 per user) into accumulator pair b, second GEMM
//                        (hidden -> messages) into the second pair.  Operand descriptors are base + constant.  An
//                        M128 N64 K16 MMA takes ~95 cycles (operand fetch), 1.5 k cycles per tile: not the limiter
//   epilogue   (4 G warps) warp w owns TMEM lane quadrant w % 4 (32 rows) and column group w / 4 (64 / G columns);
//              E1: +b1, ReLU, fp16 -> hidden tile;  E2: (+b2) * active, sum over the other user, scale, fp16 -> stage.
//              Order E1(i), E2(i-1), E1(i+1), ...: every GEMM has a full epilogue of the other kind to complete in.
//   store      (1 warp)  tensor-map TMA store of the staged messages (rows past the plane are clipped), then releases
//                        the stage to the producer
// Biases are kernel parameters (constant-bank operands).  TMEM: 2 x 128 columns (first GEMM) + 2 x 128 (second).
// Measured (nrx_large, 30 slots, us per launch): one-tile-per-CTA kernel 79; this kernel with the messages copied out
// by the epilogue warps 72 (8 warps) / 77 (16 warps); messages stored straight from registers 97; see DESIGN.md 4.3.
#pragma once
#include "nrx_kernels.cuh"
#include "nrx_stack_ws.cuh"   // WS_TICK / g_ws_cycles (debug builds with -DNRX_PHASE_TIMING; tools/agg_timing.py)

namespace nrx {

constexpr int kAggStages = 6;
#ifndef NRX_AGG_GROUPS
#define NRX_AGG_GROUPS 2
#endif
constexpr int kAggGroups = NRX_AGG_GROUPS;                    // column groups = epilogue warps per lane quadrant
constexpr int kAggEWarps = 4 * kAggGroups;
constexpr int kAggWsThreads = (kAggEWarps + 3) * 32;
constexpr int kAggWsSmem = 1024 + ((kAggBlob + 1023) / 1024) * 1024 + kAggStages * 32768;

struct alignas(64) AggWsParams {
    AggParams a;
    CUtensorMap map_a;         // message tensor, same shape and box as a.map_s
    float b1[64], b2[64];
};

__device__ __forceinline__ void agg_ld_acc(uint32_t taddr, float (&v)[32]) { tmem_ld32(taddr, v); }
__device__ __forceinline__ void agg_ld_acc(uint32_t taddr, float (&v)[16]) { tmem_ld16(taddr, v); }

// the tile sequence of one CTA (tile, tile + grid, ...) as (slot, row tile) without a division per step
struct AggTileIter {
    int tile, sl, rt, step_sl, step_rt, per_b, step;
    __device__ AggTileIter(int first, int stride, int tiles_per_b)
        : tile(first), sl(first / tiles_per_b), rt(first % tiles_per_b), step_sl(stride / tiles_per_b),
          step_rt(stride % tiles_per_b), per_b(tiles_per_b), step(stride) {}
    __device__ void next() {
        tile += step;
        sl += step_sl;
        rt += step_rt;
        if (rt >= per_b) {
            rt -= per_b;
            ++sl;
        }
    }
};

__global__ void __launch_bounds__(kAggWsThreads, 1) nrx_agg_ws_kernel(const __grid_constant__ AggWsParams wp) {
    const AggParams& p = wp.a;
    constexpr int U = 2;
    constexpr int CW = 64 / kAggGroups;                            // columns per epilogue warp
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
    uint8_t* sW = smem;
    uint8_t* sIn = smem + ((kAggBlob + 1023) / 1024) * 1024;      // kAggStages x [2 users][128][128 B]
    __shared__ uint64_t bar_w, bar_full[kAggStages], bar_empty[kAggStages], bar_out[kAggStages];
    __shared__ uint64_t bar_acc1[2], bar_hid[2], bar_acc2[2], bar_acc2_free[2];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int kTensorWarp = kAggEWarps, kProducerWarp = kAggEWarps + 1, kStoreWarp = kAggEWarps + 2;
    if (warp == kTensorWarp) tmem_alloc(&tmem_slot, 512);
    if (tid == 0) {
        mbar_init(&bar_w, 1);
        for (int s = 0; s < kAggStages; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], 1);
            mbar_init(&bar_out[s], kAggEWarps);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bar_acc1[b], 1);
            mbar_init(&bar_hid[b], kAggEWarps);
            mbar_init(&bar_acc2[b], 1);
            mbar_init(&bar_acc2_free[b], kAggEWarps);
        }
        fence_mbar_init();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = tmem_slot;
#ifdef NRX_PHASE_TIMING
    const bool ws_timed = blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == kTensorWarp || warp == kProducerWarp);
    unsigned long long ws_cyc[16];
    for (int i = 0; i < 16; ++i) ws_cyc[i] = 0;
    long long ws_last = clock64();
#endif

    // every role walks the same tile sequence and counts the tiles that carry work the same way
    auto idle_slot = [&](int sl) {
        return p.skip_idle && (__ldg(p.active_tx + sl * U) + __ldg(p.active_tx + sl * U + 1) <= 1.f);
    };

    if (warp == kProducerWarp) {
        // ---------------- producer ----------------
        if (lane == 0) {
            mbar_arrive_expect_tx(&bar_w, kAggBlob);
            bulk_g2s(sW, p.wblob, kAggBlob, &bar_w);
            int j = 0;
            for (AggTileIter t(blockIdx.x, gridDim.x, p.tiles_per_b); t.tile < p.num_tiles; t.next()) {
                if (idle_slot(t.sl)) continue;
                const int s = j % kAggStages;
                WS_TICK(0);
                if (j >= kAggStages) mbar_wait_sleep(&bar_empty[s], ((j / kAggStages) - 1) & 1);
                WS_TICK(1);
#ifdef NRX_AGW_NOLOAD
                mbar_arrive(&bar_full[s]);
#else
                mbar_arrive_expect_tx(&bar_full[s], U * 16384);
#pragma unroll
                for (int u = 0; u < U; ++u)
                    tma_load_3d(sIn + s * 32768 + u * 16384, &p.map_s, 0, t.rt * 128, t.sl * U + u, &bar_full[s]);
#endif
                ++j;
            }
        }
    } else if (warp == kTensorWarp) {
        // ---------------- tensor warp ----------------
        if (lane == 0) {
            mbar_wait_sleep(&bar_w, 0);
            int n = 0;
            for (AggTileIter t(blockIdx.x, gridDim.x, p.tiles_per_b); t.tile < p.num_tiles; t.next()) n += idle_slot(t.sl) ? 0 : 1;
            constexpr uint32_t idesc = umma_idesc_f16(128, 64);
            // descriptor = base descriptor + (byte offset >> 4): the start-address field never overflows below 256 KB
            const uint64_t d_in = umma_smem_desc(smem_u32(sIn));
            const uint64_t d_w1 = umma_smem_desc(smem_u32(sW)), d_w2 = umma_smem_desc(smem_u32(sW + 8192));
            for (int j = 0; j <= n; ++j) {
                if (j < n) {
                    // first GEMM of tile j: its accumulator pair was drained by E1(j-2), observed below in round j-1
                    const int s = j % kAggStages, b = j & 1;
                    WS_TICK(0);
                    mbar_wait_sleep(&bar_full[s], (j / kAggStages) & 1);
                    WS_TICK(1);
                    tc_fence_after_sync();
                    const uint64_t da = d_in + uint64_t(uint32_t(s) * (32768u >> 4));
                    const uint32_t acc = tbase + b * 128;
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_f16(acc + u * 64, da + uint64_t((u * 16384 + k * 32) >> 4), d_w1 + uint64_t((k * 32) >> 4), idesc, k > 0);
                    umma_commit(&bar_acc1[b]);
                    WS_TICK(2);
                }
                if (j >= 1) {
                    // second GEMM of tile j-1: hidden tile written by E1(j-1); its accumulator pair drained by E2(j-3)
                    const int i = j - 1, b = i & 1;
                    mbar_wait_sleep(&bar_hid[b], (i >> 1) & 1);
                    WS_TICK(3);
                    if (i >= 2) mbar_wait_sleep(&bar_acc2_free[b], ((i >> 1) - 1) & 1);
                    WS_TICK(4);
                    tc_fence_after_sync();
                    const uint64_t da = d_in + uint64_t(uint32_t(i % kAggStages) * (32768u >> 4));
                    const uint32_t acc = tbase + 256 + b * 128;
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_f16(acc + u * 64, da + uint64_t((u * 16384 + k * 32) >> 4), d_w2 + uint64_t((k * 32) >> 4), idesc, k > 0);
                    umma_commit(&bar_acc2[b]);
                    WS_TICK(5);
                }
            }
        }
    } else if (warp == kStoreWarp) {
        // ---------------- store warp ----------------
        if (lane == 0) {
            int j = 0;
            for (AggTileIter t(blockIdx.x, gridDim.x, p.tiles_per_b); t.tile < p.num_tiles; t.next()) {
                if (idle_slot(t.sl)) continue;
                const int s = j % kAggStages;
                mbar_wait_sleep(&bar_out[s], (j / kAggStages) & 1);
#ifndef NRX_AGW_NOSTG
#pragma unroll
                for (int u = 0; u < U; ++u) tma_store_3d(&wp.map_a, 0, t.rt * 128, t.sl * U + u, sIn + s * 32768 + u * 16384);
                bulk_commit();
                bulk_wait_read_all();
#endif
                mbar_arrive(&bar_empty[s]);
                ++j;
            }
            bulk_wait_all();
        }
    } else {
        // ---------------- epilogue warps ----------------
        const int q = warp & 3, g = warp >> 2;
        const int r = q * 32 + lane;                       // accumulator row of this thread
        const int c0 = g * CW;                             // its columns
        constexpr int LPR = CW / 8;                        // copy-out: lanes per row (16-byte pieces of the warp's columns)
        constexpr int RPI = 32 / LPR;                      // rows per store instruction
        const int co_row = q * 32 + lane / LPR, co_cc = g * LPR + lane % LPR;
        // E2 of work tile i: messages of both users for this thread's row and columns
        auto e2 = [&](int i, float m0, float m1) {
            const int b = i & 1;
            const float pm = fmaxf(m0 + m1 - 1.f, 0.f);
            const float scale = (pm == 0.f) ? 1.f : 1.f / pm;
            WS_TICK(5);
            mbar_wait_sleep(&bar_acc2[b], (i >> 1) & 1);
            WS_TICK(6);
            tc_fence_after_sync();
            float sp[U][CW];
#pragma unroll
            for (int u = 0; u < U; ++u) agg_ld_acc(tmem_addr(tbase + 256 + b * 128 + u * 64, q * 32, c0), sp[u]);
            tmem_ld_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_acc2_free[b]);
            WS_TICK(7);
            uint8_t* stg = sIn + (i % kAggStages) * 32768;   // the tile's own stage: its hidden tiles have been consumed
            const bool both = m0 == 1.f && m1 == 1.f;      // x * 1 == x: the multiplications drop out bit-exactly
#pragma unroll
            for (int j = 0; j < CW; j += 8) {
                uint32_t pk[U][4];
#pragma unroll
                for (int e = 0; e < 8; e += 2) {
                    float o[2][U];
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
#ifdef NRX_AGW_NOE2
                        o[k][0] = sp[0][j + e + k];
                        o[k][1] = sp[1][j + e + k];
                        continue;
#endif
                        float x0 = sp[0][j + e + k] + wp.b2[c0 + j + e + k];
                        float x1 = sp[1][j + e + k] + wp.b2[c0 + j + e + k];
                        if (!both) {
                            x0 *= m0;
                            x1 *= m1;
                        }
                        const float tot = (0.f + x0) + x1;
                        o[k][0] = tot - x0;
                        o[k][1] = tot - x1;
                        if (!both) {
                            o[k][0] *= scale;
                            o[k][1] *= scale;
                        }
                    }
                    pk[0][e >> 1] = pack_half2(o[0][0], o[1][0]);
                    pk[1][e >> 1] = pack_half2(o[0][1], o[1][1]);
                }
                const int cc = (c0 + j) >> 3;
#pragma unroll
                for (int u = 0; u < U; ++u)
                    st_shared_v4(stg + u * 16384 + r * 128 + ((cc ^ (r & 7)) << 4), make_uint4(pk[u][0], pk[u][1], pk[u][2], pk[u][3]));
            }
            WS_TICK(8);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_out[i % kAggStages]);
            WS_TICK(9);
        };
        int i = 0;
        float prev_m0 = 0.f, prev_m1 = 0.f;
        AggTileIter t(blockIdx.x, gridDim.x, p.tiles_per_b);
        float m0 = 0.f, m1 = 0.f;
        if (t.tile < p.num_tiles) {
            m0 = __ldg(p.active_tx + t.sl * U);
            m1 = __ldg(p.active_tx + t.sl * U + 1);
        }
        while (t.tile < p.num_tiles) {
            const int sl = t.sl, rt = t.rt;
            t.next();
            float n0 = 0.f, n1 = 0.f;                      // activity flags of the next tile: in flight during this one
            if (t.tile < p.num_tiles) {
                n0 = __ldg(p.active_tx + t.sl * U);
                n1 = __ldg(p.active_tx + t.sl * U + 1);
            }
            if (p.skip_idle && m0 + m1 <= 1.f) {
                // at most one active user: its messages are exactly zero, nothing is read or multiplied
                const int r0 = rt * 128, valid_rows = min(128, p.rows_per_bu - r0);
                __half* dst = p.abuf + ((size_t(sl) * U * p.rows_per_bu + r0) * 64 + co_cc * 8);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if ((u == 0 ? m0 : m1) == 0.f) continue;
#pragma unroll
                    for (int k = 0; k < 32 / RPI; ++k) {
                        const int rr = co_row + k * RPI;
                        if (rr < valid_rows) *reinterpret_cast<uint4*>(dst + (size_t(u) * p.rows_per_bu + rr) * 64) = make_uint4(0, 0, 0, 0);
                    }
                }
            } else {
                // ---- E1 of work tile i ----
                const int b = i & 1;
                WS_TICK(0);
                mbar_wait_sleep(&bar_acc1[b], (i >> 1) & 1);
                WS_TICK(1);
                tc_fence_after_sync();
                {
                    float v[U][CW];
#pragma unroll
                    for (int u = 0; u < U; ++u) agg_ld_acc(tmem_addr(tbase + b * 128 + u * 64, q * 32, c0), v[u]);
                    tmem_ld_wait();
                    WS_TICK(2);
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int j = 0; j < CW; j += 8) {
                            uint32_t pk[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e)
#ifdef NRX_AGW_NOE1
                                pk[e] = pack_half2(v[u][j + 2 * e], v[u][j + 2 * e + 1]);
#else
                                pk[e] = pack_half2(fmaxf(v[u][j + 2 * e] + wp.b1[c0 + j + 2 * e], 0.f),
                                                   fmaxf(v[u][j + 2 * e + 1] + wp.b1[c0 + j + 2 * e + 1], 0.f));
#endif
                            const int cc = (c0 + j) >> 3;
                            st_shared_v4(sIn + (i % kAggStages) * 32768 + u * 16384 + r * 128 + ((cc ^ (r & 7)) << 4),
                                         make_uint4(pk[0], pk[1], pk[2], pk[3]));
                        }
                }
                WS_TICK(3);
                fence_proxy_async_smem();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_hid[b]);
                WS_TICK(4);
                // ---- E2 of the previous work tile ----
                if (i >= 1) e2(i - 1, prev_m0, prev_m1);
                prev_m0 = m0; prev_m1 = m1;
                ++i;
            }
            m0 = n0;
            m1 = n1;
        }
        if (i >= 1) e2(i - 1, prev_m0, prev_m1);
    }
#ifdef NRX_PHASE_TIMING
    if (ws_timed) {
        const int base = warp == 0 ? 0 : warp == kTensorWarp ? 16 : 32;
        for (int i = 0; i < 16; ++i) g_ws_cycles[base + i] += ws_cyc[i];
    }
#endif
    tc_fence_before_sync();
    __syncthreads();
    if (warp == kTensorWarp) tmem_dealloc(tbase, 512);
}

}  // namespace nrx

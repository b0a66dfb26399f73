// nrx_stack_ws.cuh — warp-specialised, software-pipelined sep-conv STACK kernel (sm_100a).
//
// Same function, same arithmetic in the same order as nrx_stack_kernel (nrx_stack.cuh): the three
// SeparableConv2D layers of StateInit (utils/neural_rx.py:61-132) or UpdateState (:210-270, residual
// :266) fused by line buffers along the subcarrier axis — bit-identical outputs.  What changes is the
// schedule.  nrx_stack_kernel runs a step as ONE dependent chain on all warps (depthwise -> barrier ->
// MMA issue -> everybody waits -> epilogue -> barrier, three times); here three roles run concurrently
// and hand tiles to each other through mbarriers, with two tiles in flight:
//
//   D  warps 4-7   depthwise 3x3 (HFMA2) of one tile-layer after the other, never waiting for an MMA:
//                  ... dw1(k), dw3(k-1), dw2(k), dw1(k+1), dw3(k), dw2(k+1) ...
//                  Each pass is done K-slab by K-slab (channels 0-63, then 64-127): the four UMMAs of
//                  slab 0 run on the tensor core while slab 1 is still being convolved, so one 32 KB
//                  A operand buffer suffices.
//   P  warp 8      one elected thread: TMA loads of the layer-1 input window (tensor map, rows outside
//                  the grid arrive as zeros), tcgen05.mma issue per slab, tcgen05.commit -> mbarriers.
//   E  warps 0-3   epilogues, one op behind D: TMEM -> bias -> ReLU -> fp16 -> hidden-tile ring in
//                  shared memory (epi1/epi2), or bias -> fp32 staging -> residual / pos. encoding ->
//                  coalesced global store (epi3).  Two TMEM accumulators alternate op by op.
//
// A step covers 8 subcarriers (112 rows of the M=128 tile).  The hidden tiles H1 / H2 are RINGS of 10
// subcarrier slots (8 fresh + the 2 carried from the previous step: no carry copies).  Work per thread
// in a depthwise pass: warp = output subcarrier pair, half-warp = symbol half, lane = channel quad of the
// current slab; 7 symbols x 4 channels x 2 subcarriers, sliding window of 9 symbol positions in
// registers: 1.83 shared-memory loads per tile byte instead of 2.44 and 0.2 instead of 0.5 load / store
// instructions per HFMA2, same HFMA2 count and order.  One D warp and one E warp per SM sub-partition (eight E warps, two per
// sub-partition, were measured 7 % slower: they slow the depthwise passes down by more than they gain).
//
// Hand-offs between warps use the 16 hardware named barriers (bar.arrive by the producer group, bar.sync by
// the consumer group: a waiting warp is parked by the hardware and issues nothing); mbarriers are only used
// where the hardware signals completion itself (TMA bytes landed, tcgen05.commit).  Polling loops on
// mbarriers were measured to cost 38 % of all issued instructions in the first version of this kernel.
// The warp arbiter of an SM sub-partition prefers the highest warp id: the issuing warp P (few instructions, all of
// them on the critical path) gets the top id, then the depthwise warps, whose HFMA2 stream bounds the kernel.
//
// Shared memory (update stack): pointwise B images 80 KB | A operand 32 KB (its unused rows 112-127
// hold the mbarriers) | H1 ring 37.2 KB | H2 ring 37.2 KB | layer-1 window 35 KB = 221.4 KB.
// Depthwise taps are read from global memory (L1-resident, 6.8 KB) one pass ahead.
#pragma once
#include <cuda.h>

#include "nrx_stack.cuh"

namespace nrx {

constexpr int kWsStepF = 8;                         // subcarriers per step
constexpr int kWsRows = kWsStepF * kT;              // 112 rows of the M=128 tile
constexpr int kWsWin = kWsStepF + 2;                // window / ring length in subcarriers
constexpr int kWsSlot = kT * kHRow;                 // one subcarrier of a hidden ring: 14 rows x 272 B
constexpr int kWsRing = kWsWin * kWsSlot;           // 38 080 B
constexpr int kWsDWarps = 4, kWsEWarps = 4;
constexpr int kWsThreads = (kWsDWarps + kWsEWarps + 1) * 32;   // 288
constexpr int kWsRunIn = 4;                         // subcarriers computed before c0 (as in nrx_stack.cuh)

struct alignas(64) WsParams {
    CUtensorMap map_a;     // update: a [planes][F*T][64]; init: z0 [planes][F*T][32]   (box = 140 rows, no swizzle)
    CUtensorMap map_s;     // update: s [planes][F*T][64]
    StackParams p;
    float bias[320];       // biases of the launch's stack [128 | 128 | 64]: constant-bank operands of the epilogue adds
};

template <int MODE>
struct WsSmem {
    using G = StackSmem<MODE>;                                   // layout of the global weight image
    static constexpr int kPw = G::oTap1;                         // Pw1 | Pw2 | Pw3 (contiguous in the image)
    static constexpr int offA = align_up_c(kPw, 1024);
    static constexpr int offBar = offA + kWsRows * 128;          // rows 112-127 of A slab 0 (2 048 B, never written by D)
    static constexpr int offH1 = offA + 32768;
    static constexpr int offH2 = offH1 + kWsRing;
    static constexpr int offZ = align_up_c(offH2 + kWsRing, 128);
    static constexpr int kZRow = MODE == kStackInit ? 64 : 128;
    static constexpr int kZArr = kWsWin * kT * kZRow;
    static constexpr int kZ = MODE == kStackInit ? kZArr : 2 * kZArr;
    static constexpr int kTotal = offZ + kZ + 1024;              // + base alignment slack
    static_assert(kTotal <= 232448, "shared memory budget");
};

enum WsBar : int {     // mbarriers: completion is signalled by hardware (bulk copy / TMA bytes, tcgen05.commit)
    kBarW = 0, kBarZFull, kBarAFree0, kBarAFree1, kBarDFull0, kBarDFull1, kBarCount
};
enum WsNamed : int {   // hardware named barriers (id 0 is __syncthreads)
    kNbAFull0 = 1, kNbAFull1, kNbDFree0, kNbDFree1, kNbDFullE0, kNbDFullE1, kNbH1Full, kNbH2Full, kNbH1Free, kNbH2Free,
    kNbZFree, kNbE
};
constexpr int kWsNbDP = (kWsDWarps + 1) * 32;                // D group + P warp
constexpr int kWsNbEP = (kWsEWarps + 1) * 32;                // E group + P warp
constexpr int kWsNbDE = (kWsDWarps + kWsEWarps) * 32;        // D group + E group

// two IEEE fp32 additions in one instruction (packed f32x2 datapath of sm_100)
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    unsigned long long ua, ub, ud;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(ub) : "f"(b.x), "f"(b.y));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(ud) : "l"(ua), "l"(ub));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(ud));
    return d;
}
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ void tma_load_3d_plain(void* smem_dst, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
    tma_load_3d(smem_dst, tmap, c0, c1, c2, bar);
}

// One depthwise pass of one K-slab for this thread: output subcarriers 2*fq, 2*fq+1 of the tile, symbols
// 7*h .. 7*h+6, one channel quad.  `rowp[i]` = address of symbol 7*h - 1 of window subcarrier 2*fq + i (lane
// and slab offsets included; for h == 0 it points one row below the subcarrier and is never read);
// RS = row stride.  Window position jj <-> symbol 7*h - 1 + jj; the position outside the slot (symbol -1 or
// 14) is the zero padding.  Tap order (i outer, j inner) as in dw_slide: bit-identical sums.
template <int RS>
__device__ __forceinline__ void ws_dw_pass(const uint8_t* const (&rowp)[4], bool h, const uint2 (&tap)[9], uint8_t* a_slab,
                                           const uint32_t (&a_off)[14]) {
    const uint2 zero = make_uint2(0u, 0u);
    uint2 win[3][9];
    auto load_row = [&](int i, uint2(&x)[9]) {
        const uint8_t* q = rowp[i];
        x[0] = h ? lds64(q) : zero;
#pragma unroll
        for (int j = 1; j < 8; ++j) x[j] = lds64(q + j * RS);
        x[8] = h ? zero : lds64(q + 8 * RS);
    };
    load_row(0, win[0]);
    load_row(1, win[1]);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        load_row(s + 2, win[(s + 2) % 3]);
        __half2 a0[7], a1[7];
#pragma unroll
        for (int r = 0; r < 7; ++r) a0[r] = a1[r] = __float2half2_rn(0.f);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const uint2(&x)[9] = win[(s + i) % 3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const __half2 w0 = u2h(tap[i * 3 + j].x), w1 = u2h(tap[i * 3 + j].y);
#pragma unroll
                for (int r = 0; r < 7; ++r) {
                    a0[r] = __hfma2(u2h(x[r + j].x), w0, a0[r]);
                    a1[r] = __hfma2(u2h(x[r + j].y), w1, a1[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 7; ++r) sts64(a_slab + a_off[s * 7 + r], make_uint2(h2u(a0[r]), h2u(a1[r])));
    }
}

// The op sequence of an item with K tiles (identical in all roles): L1(0), L2(0), then L1(k), L3(k-1), L2(k)
// for k = 1 .. K-1, then L3(K-1): 3K ops, every layer one op after the epilogue that feeds it.
__device__ __forceinline__ void ws_op(int o, int K, int& layer, int& k) {
    if (o < 2) { layer = o + 1; k = 0; return; }
    if (o == 3 * K - 1) { layer = 3; k = K - 1; return; }
    const int q = (o - 2) / 3, r = (o - 2) - 3 * q;
    layer = r == 0 ? 1 : r == 1 ? 3 : 2;
    k = r == 1 ? q : q + 1;
}

// Optional per-role cycle accounting (build with -DNRX_PHASE_TIMING; tools/ws_timing.py): lane 0 of the first D warp,
// of the first E warp and the P thread of CTA 0 accumulate clock64 deltas between their phase boundaries.
#ifdef NRX_PHASE_TIMING
__device__ unsigned long long g_ws_cycles[48];
#define WS_TICK(i)                                                               \
    do {                                                                         \
        if (ws_timed) {                                                          \
            const long long now_ = clock64();                                    \
            ws_cyc[i] += (unsigned long long)(now_ - ws_last);                   \
            ws_last = now_;                                                      \
        }                                                                        \
    } while (0)
#else
#define WS_TICK(i) do { } while (0)
#endif

template <int MODE>
__global__ void __launch_bounds__(kWsThreads, 1) nrx_stack_ws_kernel(const __grid_constant__ WsParams wp) {
    using G = StackSmem<MODE>;
    using L = WsSmem<MODE>;
    const StackParams& p = wp.p;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
    uint8_t* sW = smem;
    uint8_t* sA = smem + L::offA;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::offBar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::offBar + kBarCount * 8);
    uint8_t* sH1 = smem + L::offH1;
    uint8_t* sH2 = smem + L::offH2;
    uint8_t* sZ = smem + L::offZ;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool role_e = warp < kWsEWarps, role_d = !role_e && warp < kWsEWarps + kWsDWarps, role_p = !role_e && !role_d;

    if (tid == 0) {
        mbar_init(&bars[kBarW], 1);
        mbar_init(&bars[kBarZFull], 1);
        mbar_init(&bars[kBarAFree0], 1);
        mbar_init(&bars[kBarAFree1], 1);
        mbar_init(&bars[kBarDFull0], 1);
        mbar_init(&bars[kBarDFull1], 1);
        fence_mbar_init();
    }
    if (role_p) tmem_alloc(tmem_slot, 256);             // two accumulators of 128 columns
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = *tmem_slot;

    // ---- per-thread constants of the roles ----------------------------------------------------------
    // D: warp = output subcarrier pair, half-warp = symbol half, lane & 15 = channel quad of the slab
    const int dfq = warp & 3, dh = lane >> 4, dq = lane & 15;
    uint32_t a_off[14];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int r = 0; r < 7; ++r) {
            const int row = (2 * dfq + s) * kT + 7 * dh + r;
            a_off[s * 7 + r] = uint32_t(row * 128 + ((((dq >> 1) ^ (row & 7)) << 4) | ((dq & 1) << 3)));
        }
    // E: warp = TMEM lane quadrant; thread = accumulator row
    const int q4 = warp & 3;
    const int erow = q4 * 32 + lane;
    const int efl = erow / kT, et = erow - efl * kT;
    const bool erow_ok = erow < kWsRows;
    // E copy-out tasks: thread = (row co_rr + 16*it, 8-channel group co_g), it = 0..6
    constexpr int kCo = kWsRows / 16;
    const int etid = tid & 127;
    const int co_g = etid & 7, co_rr = etid >> 3;
    int co_fl[kCo], co_t[kCo];
#pragma unroll
    for (int it = 0; it < kCo; ++it) {
        co_fl[it] = (co_rr + 16 * it) / kT;
        co_t[it] = (co_rr + 16 * it) - co_fl[it] * kT;
    }

#ifdef NRX_PHASE_TIMING
    const bool ws_timed = blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == kWsEWarps || warp == kWsEWarps + kWsDWarps);
    unsigned long long ws_cyc[16];
    for (int i = 0; i < 16; ++i) ws_cyc[i] = 0;
    long long ws_last = clock64();
#endif
    const uint8_t* gtap_base = nullptr;                 // taps of the loaded stack in global memory
    uint32_t n_op = 0, n_tile = 0, n_w = 0;             // global op / tile / weight-load counters (identical in all roles)
    uint32_t n_dfull[2] = {0, 0};                       // P: completed phases of the two "accumulator full" mbarriers
    int loaded_stack = -1;

    StackCursor cur = stack_begin(p);                   // balanced CTA ranges or equal chunks per plane (nrx_stack.cuh)
    int bu, c0, c1;
    while (stack_next(p, cur, bu, c0, c1)) {
        const int K = (c1 - c0 + kWsRunIn + kWsStepF - 1) / kWsStepF;
        const int stack = stack_of(p, bu);
        __syncthreads();                                // B1: every role has finished the previous item
        if (stack != loaded_stack) {
            const uint8_t* blob = p.wblob + size_t(stack) * G::kBlob;
            if (tid == kWsThreads - 32) {
                mbar_arrive_expect_tx(&bars[kBarW], uint32_t(L::kPw));
                bulk_g2s(sW, blob, L::kPw, &bars[kBarW]);
            }
            mbar_wait(&bars[kBarW], n_w & 1);
            ++n_w;
            gtap_base = blob;
            loaded_stack = stack;
        }
        {   // the two carried subcarriers of both rings start as zeros (run-in rows / f < 0 padding)
            const uint4 z = make_uint4(0, 0, 0, 0);
            for (int i = tid; i < 2 * kWsSlot / 16; i += kWsThreads) {
                st_shared_v4(sH1 + i * 16, z);
                st_shared_v4(sH2 + i * 16, z);
            }
        }
        __syncthreads();                                // B2

        if (role_d) {
            // =====================================================================================
            // D: depthwise passes
            // =====================================================================================
            // depthwise taps of the pass after the current one are requested before its arithmetic starts (global
            // memory, L1-resident: 6.8 KB per stack): `tap` holds the current pass, `tapn` the next
            uint2 tap[9], tapn[9];
            auto fetch_taps = [&](int layer, int slab) {
                const int tap_off = layer == 1 ? G::oTap1 : layer == 2 ? G::oTap2 : G::oTap3;
                const int kp = layer == 1 ? G::KP1 : 128;
                const int ch = slab * 64 + 4 * dq;
                const bool ok = ch < kp;                        // (StateInit layer 1 has 32 channels)
#pragma unroll
                for (int i = 0; i < 9; ++i)
                    tapn[i] = ok ? __ldg(reinterpret_cast<const uint2*>(gtap_base + tap_off + (i * kp + ch) * 2)) : make_uint2(0u, 0u);
            };
            auto rotate_taps = [&]() {
#pragma unroll
                for (int i = 0; i < 9; ++i) tap[i] = tapn[i];
            };
            fetch_taps(1, 0);                                   // first pass of the item: layer 1, slab 0
            auto slab_done = [&](int slab) {             // this thread's A-operand stores -> visible to the tensor core
                fence_proxy_async_smem();
                __syncwarp();
                named_bar_arrive(kNbAFull0 + slab, kWsNbDP);
            };
#pragma unroll 1
            for (int o = 0; o < 3 * K; ++o) {
                int layer, k;
                ws_op(o, K, layer, k);
                const uint32_t n = n_op + uint32_t(o), g = n_tile + uint32_t(k);
                WS_TICK(0);
                if (layer == 1) {
                    constexpr int RS = L::kZRow;
                    mbar_wait(&bars[kBarZFull], g & 1);
                    WS_TICK(1);
#pragma unroll 1
                    for (int slab = 0; slab < 2; ++slab) {
                        mbar_wait(&bars[kBarAFree0 + slab], (n & 1) ^ 1);
                        WS_TICK(2);
                        rotate_taps();
                        {   // next pass: slab 1 of this op, or slab 0 of the next op
                            int nl = layer, nk;
                            if (slab == 1 && o + 1 < 3 * K) ws_op(o + 1, K, nl, nk);
                            fetch_taps(nl, slab ^ 1);
                        }
                        if (MODE == kStackUpdate || (slab == 0 && dq < 8)) {
                            const uint8_t* zb = sZ + (MODE == kStackUpdate ? slab * L::kZArr : 0) +
                                                (2 * dfq * kT + 7 * dh - 1) * RS + dq * 8;
                            const uint8_t* const rq[4] = {zb, zb + kT * RS, zb + 2 * kT * RS, zb + 3 * kT * RS};
                            ws_dw_pass<RS>(rq, dh != 0, tap, sA + slab * 16384, a_off);
                        }
                        WS_TICK(3);
                        slab_done(slab);
                        WS_TICK(4);
                    }
                } else {
                    const uint8_t* ring = layer == 2 ? sH1 : sH2;
                    int slot = (kWsStepF * k) % kWsWin + 2 * dfq;
                    slot = slot >= kWsWin ? slot - kWsWin : slot;
                    const uint8_t* rp[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        rp[i] = ring + slot * kWsSlot + (7 * dh - 1) * kHRow + dq * 8;
                        slot = slot + 1 == kWsWin ? 0 : slot + 1;
                    }
                    named_bar_sync(layer == 2 ? kNbH1Full : kNbH2Full, kWsNbDE);
                    WS_TICK(5);
#pragma unroll 1
                    for (int slab = 0; slab < 2; ++slab) {
                        rotate_taps();
                        {
                            int nl = layer, nk;
                            if (slab == 1 && o + 1 < 3 * K) ws_op(o + 1, K, nl, nk);
                            fetch_taps(nl, slab ^ 1);
                        }
                        const uint8_t* const rq[4] = {rp[0] + slab * 128, rp[1] + slab * 128, rp[2] + slab * 128, rp[3] + slab * 128};
                        mbar_wait(&bars[kBarAFree0 + slab], (n & 1) ^ 1);
                        WS_TICK(2);
                        ws_dw_pass<kHRow>(rq, dh != 0, tap, sA + slab * 16384, a_off);
                        WS_TICK(6);
                        slab_done(slab);
                        WS_TICK(4);
                    }
                }
                // every lane's reads of the window are complete: hand it back to its producer (if the item has
                // another tile that needs it)
                __syncwarp();
                if (k + 1 < K) {
                    if (layer == 1) named_bar_arrive(kNbZFree, kWsNbDP);
                    else named_bar_arrive(layer == 2 ? kNbH1Free : kNbH2Free, kWsNbDE);
                }
            }
        } else if (role_e) {
            // =====================================================================================
            // E: epilogues
            // =====================================================================================
            auto drained = [&](int acc, bool reused) {  // this warp's TMEM reads of the accumulator are complete
                tc_fence_before_sync();
                __syncwarp();
                if (reused) named_bar_arrive(kNbDFree0 + acc, kWsNbEP);
            };
            // hidden layers: accumulator -> + bias -> ReLU -> fp16 -> ring rows.  BOFF selects the layer's biases in the
            // kernel parameters (constant-bank operands: no loads, no registers).
            auto epi_hidden = [&](int o, auto boff_tag, int f_out0, uint8_t* ring, int s0, int nb_free, bool first, int nb_full) {
                constexpr int BOFF = decltype(boff_tag)::value;
                const int acc = o & 1;
                const int f = f_out0 + efl;
                const bool in_grid = f >= 0 && f < p.F;
                int slot = s0 + 2 + efl;
                slot = slot >= kWsWin ? slot - kWsWin : slot;
                uint8_t* dst = ring + slot * kWsSlot + et * kHRow;
                const uint32_t tsrc = tmem_addr(tbase + acc * 128, q4 * 32, 0);
                WS_TICK(0);
                named_bar_sync(kNbDFullE0 + acc, kWsNbEP);          // P has seen the commit of this op's MMAs
                tc_fence_after_sync();
                float v[2][32];
                tmem_ld32(tsrc, v[0]);
                tmem_ld_wait();
                WS_TICK(1);
                if (!first) named_bar_sync(nb_free, kWsNbDE);       // D has consumed the slots about to be overwritten
                WS_TICK(2);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c < 3) tmem_ld32(tsrc + (c + 1) * 32, v[(c + 1) & 1]);     // next chunk in flight during the math
                    if (erow_ok) {
                        const float(&x)[32] = v[c & 1];
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            uint4 ov;
                            ov.x = pack_relu_half2(x[j] + wp.bias[BOFF + c * 32 + j], x[j + 1] + wp.bias[BOFF + c * 32 + j + 1]);
                            ov.y = pack_relu_half2(x[j + 2] + wp.bias[BOFF + c * 32 + j + 2], x[j + 3] + wp.bias[BOFF + c * 32 + j + 3]);
                            ov.z = pack_relu_half2(x[j + 4] + wp.bias[BOFF + c * 32 + j + 4], x[j + 5] + wp.bias[BOFF + c * 32 + j + 5]);
                            ov.w = pack_relu_half2(x[j + 6] + wp.bias[BOFF + c * 32 + j + 6], x[j + 7] + wp.bias[BOFF + c * 32 + j + 7]);
                            if (!in_grid) ov = make_uint4(0, 0, 0, 0);      // zero padding of the next layer (grid edges only)
                            st_shared_v4(dst + (c * 32 + j) * 2, ov);
                        }
                    }
                    if (c < 3) tmem_ld_wait();
                }
                WS_TICK(3);
                drained(acc, o + 2 < 3 * K);
                __syncwarp();
                named_bar_arrive(nb_full, kWsNbDE);
                WS_TICK(4);
            };
            auto epi_out = [&](int o, int k) {
                const int acc = o & 1;
                const int b = c0 - kWsRunIn + kWsStepF * k;
                const int u = bu % p.U;
                // fp32 staging of subcarrier j of the tile: the H2 ring slot that epi2 of the next tile overwrites
                const int s0n = (kWsStepF * (k + 1)) % kWsWin;
                auto stage_row = [&](int fl, int t) -> uint8_t* {
                    int slot = s0n + 2 + fl;
                    slot = slot >= kWsWin ? slot - kWsWin : slot;
                    return sH2 + slot * kWsSlot + t * 256;
                };
                // what the copy-out needs from global memory — the old state (residual) or the positional encoding of its
                // 7 rows — is requested before the wait for the accumulator: the L2 latency hides behind the MMA
                uint4 co_old[kCo];
                float2 co_pe[kCo];
                bool co_ok[kCo];
#pragma unroll
                for (int it = 0; it < kCo; ++it) {
                    const int f = b + co_fl[it];
                    co_ok[it] = f >= c0 && f < c1;
                    co_old[it] = make_uint4(0, 0, 0, 0);
                    co_pe[it] = make_float2(0.f, 0.f);
                    if (co_ok[it]) {
                        if constexpr (MODE == kStackUpdate) {
                            co_old[it] = __ldg(reinterpret_cast<const uint4*>(p.s_in + ((size_t(bu) * p.F + f) * kT + co_t[it]) * 64 + co_g * 8));
                        } else if (8 * co_g + 8 > p.d_s) {
                            co_pe[it] = __ldg(reinterpret_cast<const float2*>(p.pos_enc + ((size_t(u) * p.F + f) * kT + co_t[it]) * 2));
                        }
                    }
                }
                WS_TICK(5);
                named_bar_sync(kNbDFullE0 + acc, kWsNbEP);
                tc_fence_after_sync();
                float v[2][32];
                tmem_ld32(tmem_addr(tbase + acc * 128, q4 * 32, 0), v[0]);
                tmem_ld32(tmem_addr(tbase + acc * 128, q4 * 32, 32), v[1]);
                tmem_ld_wait();
                drained(acc, o + 2 < 3 * K);
                WS_TICK(6);
                if (erow_ok) {
                    uint8_t* dst = stage_row(efl, et);
#pragma unroll
                    for (int c = 0; c < 2; ++c)
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float4 ov;
                            ov.x = v[c][j] + wp.bias[256 + c * 32 + j];
                            ov.y = v[c][j + 1] + wp.bias[256 + c * 32 + j + 1];
                            ov.z = v[c][j + 2] + wp.bias[256 + c * 32 + j + 2];
                            ov.w = v[c][j + 3] + wp.bias[256 + c * 32 + j + 3];
                            const int c4 = c * 8 + (j >> 2);
                            *reinterpret_cast<float4*>(dst + ((c4 ^ (et & 7)) << 4)) = ov;
                        }
                }
                WS_TICK(7);
                named_bar_sync(kNbE, kWsEWarps * 32);                 // staging complete
                WS_TICK(8);
#pragma unroll
                for (int it = 0; it < kCo; ++it) {
                    if (!co_ok[it]) continue;
                    const int fl = co_fl[it], t = co_t[it];
                    const uint8_t* src = stage_row(fl, t);
                    const int sw = t & 7;
                    const float4 o0 = *reinterpret_cast<const float4*>(src + (((2 * co_g) ^ sw) << 4));
                    const float4 o1 = *reinterpret_cast<const float4*>(src + (((2 * co_g + 1) ^ sw) << 4));
                    float a[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
                    if constexpr (MODE == kStackUpdate) {   // s <- s + update (:266); pe channels ride along (update = 0 there)
                        const uint32_t ow[4] = {co_old[it].x, co_old[it].y, co_old[it].z, co_old[it].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 of = __half22float2(u2h(ow[e]));
                            a[2 * e] += of.x;
                            a[2 * e + 1] += of.y;
                        }
                    } else if (8 * co_g + 8 > p.d_s) {      // append the positional encoding after the d_s state channels
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int ch = 8 * co_g + e;
                            if (ch >= p.d_s) a[e] = ch == p.d_s ? co_pe[it].x : ch == p.d_s + 1 ? co_pe[it].y : 0.f;
                        }
                    }
                    uint4 pk;
                    pk.x = pack_half2(a[0], a[1]);
                    pk.y = pack_half2(a[2], a[3]);
                    pk.z = pack_half2(a[4], a[5]);
                    pk.w = pack_half2(a[6], a[7]);
                    *reinterpret_cast<uint4*>(p.s_out + ((size_t(bu) * p.F + (b + fl)) * kT + t) * 64 + co_g * 8) = pk;
                }
                WS_TICK(9);
                named_bar_sync(kNbE, kWsEWarps * 32);                 // staging slots free again (epi2 of the next tile)
                WS_TICK(10);
            };
#pragma unroll 1
            for (int o = 0; o < 3 * K; ++o) {
                int layer, k;
                ws_op(o, K, layer, k);
                if (layer == 3) {
                    epi_out(o, k);
                } else {
                    const int b = c0 - kWsRunIn + kWsStepF * k;
                    const int s0 = (kWsStepF * k) % kWsWin;
                    if (layer == 1) epi_hidden(o, std::integral_constant<int, 0>{}, b + 2, sH1, s0, kNbH1Free, k == 0, kNbH1Full);
                    else epi_hidden(o, std::integral_constant<int, 128>{}, b + 1, sH2, s0, kNbH2Free, k == 0, kNbH2Full);
                }
            }
        } else {
            // =====================================================================================
            // P: TMA loads of the layer-1 window, MMA issue.  The whole warp takes part in the named barriers,
            //    lane 0 issues.
            // =====================================================================================
            auto load_z = [&](int k, uint32_t g) {       // lane 0
                const int row0 = (c0 - kWsRunIn + kWsStepF * k + 1) * kT;
                mbar_arrive_expect_tx(&bars[kBarZFull], uint32_t(L::kZ));
                tma_load_3d(sZ, &wp.map_a, 0, row0, bu, &bars[kBarZFull]);
                if constexpr (MODE == kStackUpdate) tma_load_3d(sZ + L::kZArr, &wp.map_s, 0, row0, bu, &bars[kBarZFull]);
            };
            if (lane == 0) load_z(0, n_tile);
#pragma unroll 1
            for (int o = 0; o < 3 * K; ++o) {
                int layer, k;
                ws_op(o, K, layer, k);
                const uint32_t n = n_op + uint32_t(o);
                const int acc = o & 1;
                const int w_off = layer == 1 ? G::oPw1 : layer == 2 ? G::oPw2 : G::oPw3;
                const int N = layer == 3 ? 64 : 128;
                const int k_slab0 = layer == 1 && G::KP1 < 64 ? G::KP1 : 64;
                const bool second = layer != 1 || G::KP1 > 64;
                const uint32_t idesc = umma_idesc_f16(128, N);
                const uint32_t d = tbase + acc * 128;
                WS_TICK(0);
                if (o >= 2) named_bar_sync(kNbDFree0 + acc, kWsNbEP);        // E has drained this accumulator
                WS_TICK(1);
                named_bar_sync(kNbAFull0, kWsNbDP);
                WS_TICK(2);
                if (lane == 0) {
                    tc_fence_after_sync();
                    umma_gemm_k(d, smem_u32(sA), 16384, smem_u32(sW + w_off), N * 128, k_slab0, idesc, false);
                    umma_commit(&bars[kBarAFree0]);
                }
                WS_TICK(3);
                named_bar_sync(kNbAFull1, kWsNbDP);
                WS_TICK(4);
                if (lane == 0) {
                    tc_fence_after_sync();
                    if (second)
                        umma_gemm_k(d, smem_u32(sA + 16384), 16384, smem_u32(sW + w_off + N * 128), N * 128, 64, idesc, true);
                    umma_commit(&bars[kBarAFree1]);
                    umma_commit(&bars[kBarDFull0 + acc]);
                }
                if (layer == 1 && k + 1 < K) {                               // next tile's window: D has finished with this one
                    named_bar_sync(kNbZFree, kWsNbDP);
                    if (lane == 0) load_z(k + 1, n_tile + uint32_t(k) + 1);
                }
                WS_TICK(5);
                if (lane == 0) {                                             // the only poller of the MMA completion
                    mbar_wait(&bars[kBarDFull0 + acc], n_dfull[acc] & 1);
                    ++n_dfull[acc];
                }
                __syncwarp();
                named_bar_arrive(kNbDFullE0 + acc, kWsNbEP);                 // release the E group
                WS_TICK(6);
            }
        }
        __syncwarp();
        WS_TICK(15);                                    // item switch: barriers, weight load, carry zeroing, drain
        n_op += 3u * uint32_t(K);
        n_tile += uint32_t(K);
    }
#ifdef NRX_PHASE_TIMING
    if (ws_timed) {
        const int base = warp == kWsEWarps ? 0 : warp == 0 ? 16 : 32;
        for (int i = 0; i < 16; ++i) g_ws_cycles[base + i] += ws_cyc[i];
    }
#endif
    tc_fence_before_sync();
    __syncthreads();
    if (role_p) tmem_dealloc(tbase, 256);
}

// Chunks per (slot, user) plane for the pipelined kernel: minimise  waves x (ops per item + drain)  on
// `num_sms` persistent CTAs (an item costs three ops per 8-subcarrier step plus about three op slots of
// pipeline fill / drain).
inline int ws_choose_chunks(int planes, int F, int num_sms) {
    int best_n = 1;
    long long best = -1;
    const int n_max = F / 5 > 0 ? F / 5 : 1;
    for (int n = 1; n <= n_max && n <= 512; ++n) {
        const int lmax = (F + n - 1) / n;
        const long long steps = (lmax + kWsRunIn + kWsStepF - 1) / kWsStepF;
        const long long waves = ((long long)planes * n + num_sms - 1) / num_sms;
        const long long cost = waves * (3 * steps + 3);
        if (best < 0 || cost < best) { best = cost; best_n = n; }
    }
    return best_n;
}

}  // namespace nrx

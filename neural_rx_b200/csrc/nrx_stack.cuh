// nrx_stack.cuh — fused separable-conv STACK kernel (sm_100a): the three SeparableConv2D layers of
// StateInit (utils/neural_rx.py:61-132) or of UpdateState (:210-270, incl. the residual of :266)
// run inside ONE kernel; the two 128-channel hidden activations never leave the SM.
//
// Work decomposition.  An *item* is one (slot, user) plane restricted to a chunk [c0, c1) of
// subcarriers.  A persistent CTA walks its chunk in steps of 9 subcarriers (126 rows = one M=128
// UMMA tile).  At step k (b = c0 - 4 + 9k) it computes, one after the other,
//     layer 1:  H1[f]  for f in [b+2, b+11)   from  Z [b+1, b+12)   (Z fetched by bulk copy)
//     layer 2:  H2[f]  for f in [b+1, b+10)   from  H1[b,   b+11)   (2 carried + 9 fresh subcarriers)
//     layer 3:  out[f] for f in [b,   b+9 )   from  H2[b-1, b+10)   (2 carried + 9 fresh subcarriers)
// i.e. every layer lags the previous one by one subcarrier, so the +-1 subcarrier halo of the 3x3
// depthwise convolution is always already on chip ("line buffer" fusion).  The only recomputation
// is the 4-subcarrier run-in at the start of a chunk.  Rows outside the grid are forced to zero
// at every layer input (Keras 'same' padding applies per layer).
//
// Per layer:  depthwise 3x3 on CUDA cores (HFMA2, register sliding window over subcarriers, two
// OFDM symbols x four channels per thread) -> fp16 A operand in the 128B-swizzled UMMA layout ->
// tcgen05.mma (fp32 accumulators in TMEM) -> epilogue (bias, ReLU, fp16) back into shared memory.
//
// Shared memory (update stack): weights of all three layers 88 KB | A operand 32 KB |
// hidden tile 9 sc + two 2-sc carries 48 KB | Z window (a | s) 38.5 KB  = 207 KB, one CTA per SM.
#pragma once
#include <type_traits>

#include "nrx_kernels.cuh"

namespace nrx {

constexpr int kStackThreads = 256;
constexpr int kHRow = 272;                  // hidden-tile row stride in bytes (256 + 16: conflict-free row-per-lane stores)
constexpr int kCarryRows = 2 * kT;          // 28
constexpr int kStepF = kTileF;              // 9 subcarriers per step
constexpr int kRunIn = 4;                   // subcarriers computed before c0 (see header comment)

enum StackMode : int { kStackInit = 0, kStackUpdate = 1 };

struct StackParams {
    const __half* z0;            // init:   [BU*F*T][32]  normalised [y, pe, h_ls] (nrx_prep_kernel)
    const __half* a_in;          // update: [BU*F*T][64]  aggregated messages a
    const __half* s_in;          // update: [BU*F*T][64]  state s | pe | 0   (also the residual)
    __half* s_out;               //         [BU*F*T][64]  new state (never aliases s_in: halo rows of
                                 //                       neighbouring chunks are read while others are written)
    __half* sp_out;              // [BU*F*T][64] or null: message MLP of the NEXT iteration's AggregateUserStates
                                 //   applied to the new state, sp = Dense2(relu(Dense1(s_new)))  (:184-188)
    const float* active_tx;      // [B][U] (pair mode)
    int pair_agg;                // U == 2 fast path: a_in is the sp tensor written by the previous stack; user u
                                 //   reads the OTHER user's plane, masked by that user's active flag (:192-204 with
                                 //   two users: a_u = sp_v * active_v, scale 1)
    const uint8_t* wblob;        // per stack: StackSmem<MODE> weight image
    const int32_t* stack_index;  // [BU] or null
    const float* pos_enc;        // [U][F][T][2]   (init)
    int F, U, d_s;
    int n_chunks, num_items;     // n_chunks == 0: balanced ranges (nrx_stack_kernel, nrx_stack_ws_kernel), see stack_begin
    int num_planes;              // planes of the launch (all of them; with plane_list the active count is plane_list[0])
    int default_stack;
    int n_stacks;                // stacks in wblob: stack_index values are clamped to [0, n_stacks)
    const int32_t* plane_list;   // inactive-user skipping: [0] = number of active planes, [1 + i] = i-th one; or null
};

// work items of a launch and the (slot, user) plane of item `item` (all planes, or only the active ones)
__device__ __forceinline__ int stack_num_items(const StackParams& p) {
    return p.plane_list ? p.plane_list[0] * p.n_chunks : p.num_items;
}
__device__ __forceinline__ int stack_plane(const StackParams& p, int item, int& cj) {
    const int pl = item / p.n_chunks;
    cj = item - pl * p.n_chunks;
    return p.plane_list ? p.plane_list[1 + pl] : pl;
}

// Work distribution of nrx_stack_kernel / nrx_stack_ws_kernel.  n_chunks > 0: every plane is cut into n_chunks equal chunks, CTA b takes the
// items b, b + grid, ... (choose_chunks picks n_chunks; the makespan is waves x steps of the longest chunk).
// n_chunks == 0 ("balanced"): the launch's planes are laid end to end (W = planes x F subcarriers) and CTA b owns the
// subcarriers [b W / G, (b + 1) W / G) of that line, G = min(grid, W / kMinSeg); where the range crosses a plane boundary
// it becomes two (or more) items.  Every CTA then walks ~W / G + 4 subcarriers whatever the batch: at 30 slots x 2 users
// x 1584 subcarriers on 148 CTAs 73-74 steps instead of 7 waves x 11 = 77.  Outputs do not depend on the cut (each item
// recomputes its own 4-subcarrier run-in exactly).
constexpr int kMinSeg = 5;
struct StackCursor {
    long long g, g_end;
};
__host__ __device__ __forceinline__ void stack_balanced_range(long long W, int grid, int cta, long long& g0, long long& g1) {
    long long G = W / kMinSeg;
    G = G < 1 ? 1 : G > grid ? grid : G;
    if (cta >= G) { g0 = g1 = 0; return; }
    g0 = cta * W / G;
    g1 = (cta + 1) * W / G;
}
__device__ __forceinline__ StackCursor stack_begin(const StackParams& p) {
    StackCursor c;
    const int planes = p.plane_list ? p.plane_list[0] : p.num_planes;
    if (p.n_chunks > 0) {
        c.g = blockIdx.x;
        c.g_end = (long long)planes * p.n_chunks;
    } else {
        stack_balanced_range((long long)planes * p.F, int(gridDim.x), int(blockIdx.x), c.g, c.g_end);
    }
    return c;
}
// next item of this CTA: plane bu, subcarriers [c0, c1); false when the CTA's work is done
__device__ __forceinline__ bool stack_next(const StackParams& p, StackCursor& c, int& bu, int& c0, int& c1) {
    if (c.g >= c.g_end) return false;
    int pl;
    if (p.n_chunks > 0) {
        pl = int(c.g / p.n_chunks);
        const int cj = int(c.g - (long long)pl * p.n_chunks);
        c0 = int((long long)cj * p.F / p.n_chunks);
        c1 = int((long long)(cj + 1) * p.F / p.n_chunks);
        c.g += gridDim.x;
    } else {
        pl = int(c.g / p.F);
        c0 = int(c.g - (long long)pl * p.F);
        const long long left = c.g_end - c.g;
        c1 = left < p.F - c0 ? c0 + int(left) : p.F;
        c.g += c1 - c0;
    }
    bu = p.plane_list ? p.plane_list[1 + pl] : pl;
    return true;
}

// per-plane stack (Var-IO) with the index clamped into the weight image: the index arrays come straight from the caller
__device__ __forceinline__ int stack_of(const StackParams& p, int bu) {
    const int s = p.stack_index ? p.stack_index[bu] : p.default_stack;
    return min(max(s, 0), p.n_stacks - 1);
}

__host__ __device__ constexpr int align_up_c(int v, int a) { return (v + a - 1) / a * a; }

template <int MODE>
struct StackSmem {
    static constexpr int KP1 = MODE == kStackInit ? 32 : 128;          // padded K of layer 1
    static constexpr int kPw1 = ((KP1 + 63) / 64) * 128 * 128;         // B images: [K slabs][N rows][128 B]
    static constexpr int kPw2 = 2 * 128 * 128, kPw3 = 2 * 64 * 128;
    static constexpr int kTap1 = 9 * KP1 * 2, kTap = 9 * 128 * 2;      // depthwise taps [9][K] fp16
    static constexpr int oPw1 = 0, oPw2 = kPw1, oPw3 = oPw2 + kPw2;
    static constexpr int oTap1 = oPw3 + kPw3, oTap2 = oTap1 + kTap1, oTap3 = oTap2 + kTap;
    static constexpr int oBias = oTap3 + kTap;                         // fp32 [128 | 128 | 64]
    static constexpr int oAggW1 = align_up_c(oBias + (128 + 128 + 64) * 4, 1024);   // message MLP: W1 image [64 rows][128 B]
    static constexpr int oAggW2 = oAggW1 + 8192;                       //              W2 image [64 rows][128 B]
    static constexpr int oAggB = oAggW2 + 8192;                        //              b1[64] | b2[64] fp32
    static constexpr int kBlob = oAggB + 512;
    static constexpr int offA = align_up_c(kBlob, 1024);               // A operand / fp32 output staging
    static constexpr int offH = offA + 32768;                          // fresh hidden tile (9 subcarriers); 1024-aligned:
                                                                       //   doubles as the two A slabs of the message MLP
    static constexpr int offC1 = offH + align_up_c(kTileRows * kHRow, 1024);   // carry of H1 (2 subcarriers)
    static constexpr int offC2 = offC1 + kCarryRows * kHRow;           // carry of H2
    static constexpr int offZ = offC2 + kCarryRows * kHRow;            // layer-1 input window (11 subcarriers)
    static constexpr int kZRow = MODE == kStackInit ? 64 : 128;
    static constexpr int kZArr = kHaloRows * kZRow;
    static constexpr int kZ = MODE == kStackInit ? kZArr : 2 * kZArr;
    static constexpr int kTotal = offZ + kZ + 1024;                    // + base alignment slack
};

__device__ __forceinline__ uint2 lds64(const void* p) { return *reinterpret_cast<const uint2*>(p); }
__device__ __forceinline__ void sts64(void* p, uint2 v) { *reinterpret_cast<uint2*>(p) = v; }


// Depthwise 3x3 of NFOUT consecutive subcarriers for two OFDM symbols (t0, t0+1) and four
// channels.  Input rows (fi, t) live at  (fi < 2 ? carry : fresh - 2 rows) + (fi*14 + t)*RS  (the
// pointers already include the channel offset); the result goes to the swizzled A operand.
// tail[][] returns the thread's own columns of the last two input subcarriers (the next carry).
template <int RS, int NFOUT>
__device__ __forceinline__ void dw_slide(const uint8_t* carry, const uint8_t* fresh, int t0, bool has_l, bool has_r,
                                         const uint2 (&kk)[9], uint8_t* a_thr, int chunk, int r_base,
                                         uint2 (&tail)[2][2]) {
    const uint2 zero = make_uint2(0u, 0u);
    uint2 win[3][4];
    auto load_row = [&](int fi, uint2(&r)[4]) {
        const uint8_t* q = (fi < 2 ? carry + fi * (kT * RS) : fresh + (fi - 2) * (kT * RS)) + t0 * RS;
        r[0] = has_l ? lds64(q - RS) : zero;
        r[1] = lds64(q);
        r[2] = lds64(q + RS);
        r[3] = has_r ? lds64(q + 2 * RS) : zero;
    };
    load_row(0, win[0]);
    load_row(1, win[1]);
#pragma unroll
    for (int fl = 0; fl < NFOUT; ++fl) {
        load_row(fl + 2, win[(fl + 2) % 3]);
        __half2 a00 = __float2half2_rn(0.f), a01 = a00, a10 = a00, a11 = a00;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const uint2(&r)[4] = win[(fl + i) % 3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const uint2 w = kk[i * 3 + j];
                a00 = __hfma2(u2h(r[j].x), u2h(w.x), a00);
                a01 = __hfma2(u2h(r[j].y), u2h(w.y), a01);
                a10 = __hfma2(u2h(r[j + 1].x), u2h(w.x), a10);
                a11 = __hfma2(u2h(r[j + 1].y), u2h(w.y), a11);
            }
        }
        const int r0 = r_base + fl * kT + t0, r1 = r0 + 1;
        sts64(a_thr + r0 * 128 + ((chunk ^ (r0 & 7)) << 4), make_uint2(h2u(a00), h2u(a01)));
        sts64(a_thr + r1 * 128 + ((chunk ^ (r1 & 7)) << 4), make_uint2(h2u(a10), h2u(a11)));
    }
    tail[0][0] = win[NFOUT % 3][1];
    tail[0][1] = win[NFOUT % 3][2];
    tail[1][0] = win[(NFOUT + 1) % 3][1];
    tail[1][1] = win[(NFOUT + 1) % 3][2];
}

// Optional per-phase cycle accounting (build with -DNRX_PHASE_TIMING; tools/phase_timing.py):
// one thread of CTA 0 accumulates clock64 deltas between the phase boundaries of every step.
#ifdef NRX_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[32];
#define NRX_TICK(i)                                                              \
    do {                                                                         \
        if (blockIdx.x == 0 && threadIdx.x == 32) {                              \
            const long long now_ = clock64();                                    \
            s_phase[i] += (unsigned long long)(now_ - tick_last);                \
            tick_last = now_;                                                    \
        }                                                                        \
    } while (0)
#else
#define NRX_TICK(i) do { } while (0)
#endif

template <int MODE, bool MLP>
__global__ void __launch_bounds__(kStackThreads, 1) nrx_stack_kernel(StackParams p) {
    using L = StackSmem<MODE>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
    uint8_t* sW = smem;
    uint8_t* sA = smem + L::offA;
    uint8_t* sC1 = smem + L::offC1;
    uint8_t* sH = smem + L::offH;
    uint8_t* sC2 = smem + L::offC2;
    uint8_t* sZ = smem + L::offZ;
    const float* sBias = reinterpret_cast<const float*>(sW + L::oBias);
    __shared__ uint64_t bar_z, bar_w, bar_mma, bar_mlp;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&tmem_slot, MLP ? 256 : 128);   // [0,128) layer accumulator, [128,192) / [192,256) message MLP
    if (tid == 0) {
        mbar_init(&bar_z, 1);
        mbar_init(&bar_w, 1);
        mbar_init(&bar_mma, 1);
        mbar_init(&bar_mlp, 1);
        fence_mbar_init();
    }
    {   // rows 126/127 of the A operand are never produced by the depthwise pass: keep them finite
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (int i = tid; i < 32768 / 16; i += kStackThreads) st_shared_v4(sA + i * 16, z);
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = tmem_slot;
    uint32_t ph_z = 0, ph_w = 0, ph_mma = 0;
    int loaded_stack = -1;
#ifdef NRX_PHASE_TIMING
    __shared__ unsigned long long s_phase[32];
    if (tid == 32)
        for (int i = 0; i < 32; ++i) s_phase[i] = 0;
    long long tick_last = clock64();
#endif

    // ---- depthwise task mapping -------------------------------------------------------------
    // 128-channel layers: thread = (symbol pair ph in 0..6, channel quad qh in 0..31); warp-uniform ph
    const bool act_h = tid < 7 * 32;
    const int ph = tid >> 5, qh = tid & 31;
    // 32-channel first layer of StateInit: thread = (subcarrier segment, symbol pair, channel quad 0..7)
    const bool act_1 = MODE == kStackInit ? tid < 3 * 56 : act_h;
    const int seg1 = tid / 56, p1 = (tid % 56) >> 3, q1 = tid & 7;

    // copy-out task mapping: thread = (row co_rr + 32*it, 8-channel group co_g), it = 0..3
    const int co_g = tid & 7, co_rr = tid >> 3;
    int co_fl[4], co_t[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        co_fl[it] = (co_rr + 32 * it) / kT;
        co_t[it] = (co_rr + 32 * it) - co_fl[it] * kT;
    }

    const int q4 = warp & 3, hc = warp >> 2;      // epilogue: TMEM lane quadrant, column half
    const int erow = q4 * 32 + lane;              // accumulator row of this thread
    const int efl = erow / kT, et = erow - (erow / kT) * kT;

    // ---- deferred message-MLP chain of the previous tile (see the end of the step loop) ------------
    bool pend = false;
    size_t pend_row0 = 0;
    int pend_lo = 0, pend_hi = 0;
    uint32_t ph_mlp = 0;
    const float* sAggB = reinterpret_cast<const float*>(sW + L::oAggB);
    // hidden layer of the MLP: TMEM[128..191] -> relu(. + b1) -> fp16 A slab, then issue the second GEMM
    auto mlp_hidden = [&]() {
        const int col = hc * 32;
        float4 bq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) bq[j] = *reinterpret_cast<const float4*>(sAggB + col + j * 4);
        mbar_wait(&bar_mlp, ph_mlp);
        ph_mlp ^= 1;
        tc_fence_after_sync();
        float v[32];
        tmem_ld32(tmem_addr(tbase + 128, q4 * 32, col), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            const float4 b0 = bq[j >> 2], b1 = bq[(j >> 2) + 1];
            uint4 o;
            o.x = pack_relu_half2(v[j] + b0.x, v[j + 1] + b0.y);
            o.y = pack_relu_half2(v[j + 2] + b0.z, v[j + 3] + b0.w);
            o.z = pack_relu_half2(v[j + 4] + b1.x, v[j + 5] + b1.y);
            o.w = pack_relu_half2(v[j + 6] + b1.z, v[j + 7] + b1.w);
            const int cc = (col + j) >> 3;
            st_shared_v4(sH + 16384 + erow * 128 + ((cc ^ (erow & 7)) << 4), o);
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();                                // hidden slab complete
        if (tid == 0) {
            tc_fence_after_sync();
            umma_gemm_k(tbase + 192, smem_u32(sH + 16384), 16384, smem_u32(sW + L::oAggW2), 8192, 64,
                        umma_idesc_f16(128, 64), false);
            umma_commit(&bar_mlp);
        }
    };
    // wait for the second GEMM (its operands live in the hidden-tile region, which the next hidden
    // epilogue overwrites)
    auto mlp_wait2 = [&]() {
        mbar_wait(&bar_mlp, ph_mlp);
        ph_mlp ^= 1;
        tc_fence_after_sync();
    };
    // output of the MLP: TMEM[192..255] + b2 -> fp16 -> global sp tensor (row per lane)
    auto mlp_store = [&]() {
        const int col = hc * 32;
        float4 bq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) bq[j] = *reinterpret_cast<const float4*>(sAggB + 64 + col + j * 4);
        float v[32];
        tmem_ld32(tmem_addr(tbase + 192, q4 * 32, col), v);
        tmem_ld_wait();
        if (erow < kTileRows && erow >= pend_lo && erow < pend_hi) {
            __half* dst = p.sp_out + (pend_row0 + erow) * 64 + col;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                const float4 b0 = bq[j >> 2], b1 = bq[(j >> 2) + 1];
                uint4 o;
                o.x = pack_half2(v[j] + b0.x, v[j + 1] + b0.y);
                o.y = pack_half2(v[j + 2] + b0.z, v[j + 3] + b0.w);
                o.z = pack_half2(v[j + 4] + b1.x, v[j + 5] + b1.y);
                o.w = pack_half2(v[j + 6] + b1.z, v[j + 7] + b1.w);
                *reinterpret_cast<uint4*>(dst + j) = o;
            }
        }
        tc_fence_before_sync();
        pend = false;
    };

    bool z_prefetched = false;                          // the next item's first window is already in flight
    bool w_pending = false;                             // a weight blob load has been issued and not yet waited for
    StackCursor cur = stack_begin(p);
    int bu, c0, c1;
    while (stack_next(p, cur, bu, c0, c1)) {
        const int nsteps = (c1 - c0 + kRunIn + kStepF - 1) / kStepF;
        const int stack = stack_of(p, bu);
        if (stack != loaded_stack) {                   // block-uniform: first item or Var-IO switch
            if (MLP && pend) {                         // the pending MLP still reads the resident weights
                mlp_hidden();
                mlp_wait2();
                mlp_store();
            }
            __syncthreads();
            if (tid == 0) {
                mbar_arrive_expect_tx(&bar_w, L::kBlob);
                bulk_g2s(sW, p.wblob + size_t(stack) * L::kBlob, L::kBlob, &bar_w);
            }
            w_pending = true;                          // waited for after the first input window has been requested
            loaded_stack = stack;
        }
        // message source plane: the aggregated tensor of this user, or (two-user fast path) the
        // other user's sp tensor gated by that user's activity flag
        const int bu_a = p.pair_agg ? (bu ^ 1) : bu;
        const bool a_live = !p.pair_agg || p.active_tx[bu_a] != 0.f;
        {   // carries of the hidden layers start as zeros (run-in rows; also the f < 0 padding)
            const uint4 z = make_uint4(0, 0, 0, 0);
            for (int i = tid; i < kCarryRows * kHRow / 16; i += kStackThreads) {
                st_shared_v4(sC1 + i * 16, z);
                st_shared_v4(sC2 + i * 16, z);
            }
        }

        // fetch the 11-subcarrier layer-1 window starting at subcarrier zf0 (thread 0) and zero the
        // rows outside the grid (all threads); the window buffer must be free when this is called
        auto stage_z_of = [&](int zbu, int zbu_a, bool za_live, int zf0) {
            const int flo = max(zf0, 0), fhi = min(zf0 + kStepF + 2, p.F);
            const int nrow = max(fhi - flo, 0) * kT;
            if (tid == 0) {
                const size_t grow = (size_t(zbu) * p.F + flo) * kT;
                const size_t grow_a = (size_t(zbu_a) * p.F + flo) * kT;
                const int so = (flo - zf0) * kT * L::kZRow;
                if constexpr (MODE == kStackInit) {
                    mbar_arrive_expect_tx(&bar_z, uint32_t(nrow) * 64u);
                    if (nrow) bulk_g2s(sZ + so, reinterpret_cast<const uint8_t*>(p.z0) + grow * 64, uint32_t(nrow) * 64u, &bar_z);
                } else {
                    mbar_arrive_expect_tx(&bar_z, uint32_t(nrow) * (za_live ? 256u : 128u));
                    if (nrow) {
                        if (za_live)
                            bulk_g2s(sZ + so, reinterpret_cast<const uint8_t*>(p.a_in) + grow_a * 128, uint32_t(nrow) * 128u, &bar_z);
                        bulk_g2s(sZ + L::kZArr + so, reinterpret_cast<const uint8_t*>(p.s_in) + grow * 128, uint32_t(nrow) * 128u, &bar_z);
                    }
                }
            }
            const uint4 z = make_uint4(0, 0, 0, 0);
            constexpr int V = kT * L::kZRow / 16;       // 16-byte vectors per subcarrier per array
            for (int fi = 0; fi < kStepF + 2; ++fi) {
                const int f = zf0 + fi;
                if (f >= 0 && f < p.F) continue;
                for (int i = tid; i < V; i += kStackThreads) {
                    st_shared_v4(sZ + fi * kT * L::kZRow + i * 16, z);
                    if constexpr (MODE == kStackUpdate) st_shared_v4(sZ + L::kZArr + fi * kT * L::kZRow + i * 16, z);
                }
            }
        };
        auto stage_z = [&](int zf0) { stage_z_of(bu, bu_a, a_live, zf0); };
        if constexpr (MODE == kStackUpdate) {
            if (!a_live)       // the only other user is inactive: its message is masked to zero (:192-193)
                for (int i = tid; i < L::kZArr / 16; i += kStackThreads) st_shared_v4(sZ + i * 16, make_uint4(0, 0, 0, 0));
        }

        // bias + ReLU epilogue of a hidden layer: TMEM -> fp16 rows of the fresh hidden tile;
        // rows whose subcarrier is outside the grid become zeros (padding of the next layer)
        auto epi_hidden = [&](const float* bias, int f_out0) {
            const int f = f_out0 + efl;
            const bool in_grid = f >= 0 && f < p.F;
            const bool warp_oob = !__all_sync(0xffffffffu, in_grid || erow >= kTileRows);   // grid edges only
            // both 32-column TMEM loads in flight, bias rows fetched while they complete
            // (splitting the GEMM into two N=64 halves to overlap the second half with the first
            // half's conversion was measured slower: 3.89 vs 3.60 ms per 30-slot step)
            float v[2][32];
            tmem_ld32(tmem_addr(tbase, q4 * 32, hc * 64), v[0]);
            tmem_ld32(tmem_addr(tbase, q4 * 32, hc * 64 + 32), v[1]);
            float4 bq[2][8];
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int j = 0; j < 8; ++j) bq[c][j] = *reinterpret_cast<const float4*>(bias + hc * 64 + c * 32 + j * 4);
            tmem_ld_wait();
            if (erow < kTileRows) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int col = hc * 64 + c * 32;
                    uint4 o[4];
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        const float4 b0 = bq[c][j >> 2], b1 = bq[c][(j >> 2) + 1];
                        o[j >> 3].x = pack_relu_half2(v[c][j] + b0.x, v[c][j + 1] + b0.y);
                        o[j >> 3].y = pack_relu_half2(v[c][j + 2] + b0.z, v[c][j + 3] + b0.w);
                        o[j >> 3].z = pack_relu_half2(v[c][j + 4] + b1.x, v[c][j + 5] + b1.y);
                        o[j >> 3].w = pack_relu_half2(v[c][j + 6] + b1.z, v[c][j + 7] + b1.w);
                    }
                    if (warp_oob && !in_grid) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) o[j] = make_uint4(0, 0, 0, 0);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) st_shared_v4(sH + erow * kHRow + (col + j * 8) * 2, o[j]);
                }
            }
        };

        // one 128-channel depthwise pass (layers 2, 3 and the update stack's layer 1 use the same mapping)
        uint2 tail[2][2];
        auto load_taps = [&](uint2(&kk)[9], int tap_off, int kp, int q) {
#pragma unroll
            for (int i = 0; i < 9; ++i) kk[i] = lds64(sW + tap_off + (i * kp + q * 4) * 2);
        };
        // One 128-channel depthwise pass, balanced over the four SM sub-partitions: warps 0-6 own one
        // symbol pair each and slide over output subcarriers 0..7; warp 7 computes subcarrier 8 for all
        // seven symbol pairs (it is the only one that sees the last two input subcarriers, so it also
        // keeps the next carry).  Busiest sub-partition: 576 HFMA2 instead of 648.
        uint2 tail7[7][2][2];
        auto dw128 = [&](auto rs_tag, const uint8_t* carry, const uint8_t* fresh, int tap_off) {
            constexpr int RS = decltype(rs_tag)::value;
            uint2 kk[9];
            load_taps(kk, tap_off, 128, qh);
            uint8_t* a_thr = sA + (qh >> 4) * 16384 + (qh & 1) * 8;
            if (warp < 7) {
                uint2 unused[2][2];
                dw_slide<RS, kStepF - 1>(carry, fresh, 2 * ph, ph > 0, ph < 6, kk, a_thr, (qh >> 1) & 7, 0, unused);
            } else {
                const uint8_t* c8 = fresh + (kStepF - 3) * (kT * RS);      // input subcarriers 8, 9, 10
#pragma unroll
                for (int pp = 0; pp < 7; ++pp)
                    dw_slide<RS, 1>(c8, c8 + 2 * (kT * RS), 2 * pp, pp > 0, pp < 6, kk, a_thr, (qh >> 1) & 7,
                                    (kStepF - 1) * kT, tail7[pp]);
            }
        };
        auto dw_hidden = [&](const uint8_t* carry, int tap_off) {
            dw128(std::integral_constant<int, kHRow>{}, carry + qh * 8, sH + qh * 8, tap_off);
        };
        auto save_carry = [&](uint8_t* carry) {         // after the post-depthwise barrier
            if (warp == 7) {
#pragma unroll
                for (int pp = 0; pp < 7; ++pp)
#pragma unroll
                    for (int ci = 0; ci < 2; ++ci)
#pragma unroll
                        for (int e = 0; e < 2; ++e)
                            sts64(carry + ((ci * kT) + 2 * pp + e) * kHRow + qh * 8, tail7[pp][ci][e]);
            }
        };
        auto issue_mma = [&](int w_off, int b_slab_bytes, int K, int N) {
            if (tid == 0) {
                tc_fence_after_sync();
                umma_gemm_k(tbase, smem_u32(sA), 16384, smem_u32(sW + w_off), b_slab_bytes, K, umma_idesc_f16(128, N), false);
                umma_commit(&bar_mma);
            }
        };
        auto wait_mma = [&]() {
            mbar_wait(&bar_mma, ph_mma);
            ph_mma ^= 1;
            tc_fence_after_sync();
        };

        if (!z_prefetched) stage_z(c0 - kRunIn + 1);
        z_prefetched = false;
        if (w_pending) {                               // weights land while the window is in flight
            mbar_wait(&bar_w, ph_w);
            ph_w ^= 1;
            w_pending = false;
        }
        for (int k = 0; k < nsteps; ++k) {
            const int b = c0 - kRunIn + kStepF * k;
            NRX_TICK(21);                               // item set-up / loop overhead
            mbar_wait(&bar_z, ph_z);
            ph_z ^= 1;
            __syncthreads();                            // zero-filled rows / carries visible
            NRX_TICK(0);

            // ================= layer 1: Z[b+1, b+12) -> H1[b+2, b+11) =================
            if constexpr (MODE == kStackInit) {
                if (act_1) {
                    uint2 kk[9];
                    load_taps(kk, L::oTap1, 32, q1);
                    const uint8_t* zc = sZ + seg1 * 3 * kT * 64 + q1 * 8;
                    dw_slide<64, 3>(zc, zc + 2 * kT * 64, 2 * p1, p1 > 0, p1 < 6, kk, sA + (q1 & 1) * 8, q1 >> 1,
                                    seg1 * 3 * kT, tail);
                }
            } else {
                const uint8_t* zc = sZ + (qh >> 4) * L::kZArr + (qh & 15) * 8;
                dw128(std::integral_constant<int, 128>{}, zc, zc + 2 * kT * 128, L::oTap1);
            }
            NRX_TICK(1);
            fence_proxy_async_smem();
            tc_fence_before_sync();
            __syncthreads();                            // A complete; Z window free
            NRX_TICK(2);
            issue_mma(L::oPw1, 128 * 128, L::KP1, 128);
            if (k + 1 < nsteps) {
                stage_z(b + kStepF + 1);                // prefetch overlaps layers 1-3 of this step
            } else {                                    // last step: first window of the CTA's next item
                StackCursor nx = cur;
                int nbu, nc0, nc1;
                if (stack_next(p, nx, nbu, nc0, nc1) && stack_of(p, nbu) == loaded_stack) {
                    const int nbu_a = p.pair_agg ? (nbu ^ 1) : nbu;
                    const bool na_live = !p.pair_agg || p.active_tx[nbu_a] != 0.f;
                    stage_z_of(nbu, nbu_a, na_live, nc0 - kRunIn + 1);
                    z_prefetched = true;
                }
            }
            if (MLP && pend) mlp_hidden();              // previous tile's MLP, inside this GEMM's shadow
            NRX_TICK(3);
            wait_mma();
            if (MLP && pend) mlp_wait2();               // its slabs sit where the hidden tile is about to land
            NRX_TICK(4);
            epi_hidden(sBias, b + 2);
            NRX_TICK(5);
            tc_fence_before_sync();
            __syncthreads();                            // H1 fresh tile complete, TMEM drained
            NRX_TICK(6);

            // ================= layer 2: H1[b, b+11) -> H2[b+1, b+10) =================
            dw_hidden(sC1, L::oTap2);
            NRX_TICK(7);
            fence_proxy_async_smem();
            tc_fence_before_sync();
            __syncthreads();                            // A complete; H1 tile + carry fully consumed
            NRX_TICK(8);
            issue_mma(L::oPw2, 128 * 128, 128, 128);
            save_carry(sC1);
            if (MLP && pend) mlp_store();
            // prefetch what the copy-out needs from global memory (issued two GEMMs ahead of its use):
            // the old state (residual, update stack) or the positional encoding (StateInit)
            const int u = bu % p.U;
            uint4 co_old[4];
            float2 co_pe[4];
            bool co_ok[4];
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int f = b + co_fl[it];
                co_ok[it] = (co_rr + 32 * it) < kTileRows && f >= c0 && f < c1;
                co_old[it] = make_uint4(0, 0, 0, 0);
                co_pe[it] = make_float2(0.f, 0.f);
                if (co_ok[it]) {
                    if constexpr (MODE == kStackUpdate) {
                        const size_t grow = (size_t(bu) * p.F + f) * kT + co_t[it];
                        co_old[it] = __ldg(reinterpret_cast<const uint4*>(p.s_in + grow * 64 + co_g * 8));
                    } else if (8 * co_g + 8 > p.d_s) {
                        co_pe[it] = __ldg(reinterpret_cast<const float2*>(p.pos_enc + ((size_t(u) * p.F + f) * kT + co_t[it]) * 2));
                    }
                }
            }
            NRX_TICK(9);
            wait_mma();
            NRX_TICK(10);
            epi_hidden(sBias + 128, b + 1);
            NRX_TICK(11);
            tc_fence_before_sync();
            __syncthreads();
            NRX_TICK(12);

            // ================= layer 3: H2[b-1, b+10) -> out[b, b+9) =================
            dw_hidden(sC2, L::oTap3);
            NRX_TICK(13);
            fence_proxy_async_smem();
            tc_fence_before_sync();
            __syncthreads();
            NRX_TICK(14);
            issue_mma(L::oPw3, 64 * 128, 128, 64);
            save_carry(sC2);
            NRX_TICK(15);
            wait_mma();
            NRX_TICK(16);
            {   // fp32 staging [128][64] in the A buffer (free now): 16 chunks of 4 floats, chunk ^ (row & 7)
                float v[32];
                const int col = hc * 32;
                tmem_ld32(tmem_addr(tbase, q4 * 32, col), v);
                const float* b3 = sBias + 256;
                float4 b3q[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) b3q[j] = *reinterpret_cast<const float4*>(b3 + col + j * 4);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const int c4 = (col + j) >> 2;
                    const float4 bb = b3q[j >> 2];
                    float4 o;
                    o.x = v[j] + bb.x;
                    o.y = v[j + 1] + bb.y;
                    o.z = v[j + 2] + bb.z;
                    o.w = v[j + 3] + bb.w;
                    *reinterpret_cast<float4*>(sA + erow * 256 + ((c4 ^ (erow & 7)) << 4)) = o;
                }
            }
            NRX_TICK(17);
            tc_fence_before_sync();
            __syncthreads();                            // staging complete; TMEM drained
            NRX_TICK(18);

            // ---- coalesced copy-out of the chunk's own rows: residual (update) / pe append (init) ----
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                if (!co_ok[it]) continue;
                const int rr = co_rr + 32 * it;
                const float4 o0 = *reinterpret_cast<const float4*>(sA + rr * 256 + (((2 * co_g) ^ (rr & 7)) << 4));
                const float4 o1 = *reinterpret_cast<const float4*>(sA + rr * 256 + (((2 * co_g + 1) ^ (rr & 7)) << 4));
                float a[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
                if constexpr (MODE == kStackUpdate) {   // s <- s + update (:266); pe channels ride along (update = 0 there)
                    const uint32_t ow[4] = {co_old[it].x, co_old[it].y, co_old[it].z, co_old[it].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 of = __half22float2(u2h(ow[e]));
                        a[2 * e] += of.x;
                        a[2 * e + 1] += of.y;
                    }
                } else {                                // append the positional encoding after the d_s state channels
                    if (8 * co_g + 8 > p.d_s) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int ch = 8 * co_g + e;
                            if (ch >= p.d_s) a[e] = ch == p.d_s ? co_pe[it].x : ch == p.d_s + 1 ? co_pe[it].y : 0.f;
                        }
                    }
                }
                uint4 pk;
                pk.x = pack_half2(a[0], a[1]);
                pk.y = pack_half2(a[2], a[3]);
                pk.z = pack_half2(a[4], a[5]);
                pk.w = pack_half2(a[6], a[7]);
                const size_t grow = (size_t(bu) * p.F + (b + co_fl[it])) * kT + co_t[it];
                *reinterpret_cast<uint4*>(p.s_out + grow * 64 + co_g * 8) = pk;
                if (MLP && p.sp_out) st_shared_v4(sH + rr * 128 + ((co_g ^ (rr & 7)) << 4), pk);   // A slab of the message MLP
            }
            NRX_TICK(19);
            if (MLP && p.sp_out) {
                // ---- message MLP of the next AggregateUserStates on the fresh state tile (:184-188),
                //      sp = Dense2(relu(Dense1(s_new))): the first GEMM is issued here, the rest of the
                //      chain (mlp_hidden / mlp_store) runs inside the MMA bubbles of the next step
                fence_proxy_async_smem();
                tc_fence_before_sync();
                __syncthreads();                        // state slab complete; fp32 staging free
                if (tid == 0) {
                    tc_fence_after_sync();
                    umma_gemm_k(tbase + 128, smem_u32(sH), 16384, smem_u32(sW + L::oAggW1), 8192, 64,
                                umma_idesc_f16(128, 64), false);
                    umma_commit(&bar_mlp);
                }
                pend = true;
                pend_row0 = (size_t(bu) * p.F + b) * kT;     // global row of tile row 0
                pend_lo = (c0 - b) * kT;                     // tile rows [pend_lo, pend_hi) belong to the chunk
                pend_hi = (c1 - b) * kT;
            } else
            __syncthreads();                            // staging / slabs free before the next depthwise pass
            NRX_TICK(20);
        }
    }
    if (MLP && pend) {                                  // flush the last tile's MLP
        mlp_hidden();
        mlp_wait2();
        mlp_store();
    }
#ifdef NRX_PHASE_TIMING
    if (blockIdx.x == 0 && tid == 32)
        for (int i = 0; i < 32; ++i) g_phase_cycles[i] += s_phase[i];
#endif
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, MLP ? 256 : 128);
}

// Number of chunks per (slot, user) plane that minimises the makespan  waves x steps-per-item  on
// `num_sms` persistent CTAs (ties: fewer chunks = less run-in recomputation).
inline int choose_chunks(int planes, int F, int num_sms) {
    int best_n = 1;
    long long best = -1;
    const int n_max = F / 5 > 0 ? F / 5 : 1;
    for (int n = 1; n <= n_max && n <= 512; ++n) {
        const int lmax = (F + n - 1) / n;
        const long long steps = (lmax + kRunIn + kStepF - 1) / kStepF;
        const long long waves = ((long long)planes * n + num_sms - 1) / num_sms;
        const long long cost = waves * steps;
        if (best < 0 || cost < best) { best = cost; best_n = n; }
    }
    return best_n;
}

}  // namespace nrx

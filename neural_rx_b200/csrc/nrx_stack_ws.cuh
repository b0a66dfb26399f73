// nrx_stack_ws.cuh — warp-specialised variant of the fused sep-conv stack kernel (nrx_stack.cuh).
//
// Same algorithm, same shared-memory layout, same arithmetic in the same order; the difference is
// who does the output side of layer 3.  The CTA has 8 MAIN warps (depthwise passes, GEMM issue,
// hidden-layer epilogues — the critical chain) and 4 HELPER warps.  After the main warps have
// issued the layer-3 GEMM of step k they only wait until its operand buffer is free and go straight
// on to layer 1 of step k+1; the helper warps pick the accumulator up from its own TMEM columns,
// add bias, stage fp32 rows in the (now dead) hidden-tile region, add the residual / append the
// positional encoding and store the new state — all in the shadow of the main warps' next
// depthwise pass.  Hand-offs:
//   bar_mma3  (tcgen05.commit)  main -> helpers: layer-3 accumulator of step k complete
//                               (implies the depthwise pass no longer reads the hidden tile);
//                               the main warps wait on it too before overwriting the A operand
//   bar_hfree (mbarrier.arrive) helpers -> main: staging in the hidden-tile region consumed and
//                               TMEM columns read — the next hidden-layer epilogue may overwrite
// Block-wide barriers inside the step loop are named barriers over the 256 main threads
// (bar.sync 1) or the 128 helper threads (bar.sync 2).
#pragma once
#include "nrx_stack.cuh"

namespace nrx {

constexpr int kWsMain = 256, kWsHelp = 128, kWsThreads = kWsMain + kWsHelp;

__device__ __forceinline__ void bar_main() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void bar_help() { asm volatile("bar.sync 2, 128;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(kWsThreads, 1) nrx_stack_ws_kernel(StackParams p) {
    using L = StackSmem<MODE>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sW = smem;
    uint8_t* sA = smem + L::offA;
    uint8_t* sC1 = smem + L::offC1;
    uint8_t* sH = smem + L::offH;
    uint8_t* sC2 = smem + L::offC2;
    uint8_t* sZ = smem + L::offZ;
    const float* sBias = reinterpret_cast<const float*>(sW + L::oBias);
    __shared__ uint64_t bar_z, bar_w, bar_mma, bar_mma3, bar_hfree;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_main = tid < kWsMain;
    if (warp == 0) tmem_alloc(&tmem_slot, 256);        // [0,128) layers 1-2, [128,192) layer 3
    if (tid == 0) {
        mbar_init(&bar_z, 1);
        mbar_init(&bar_w, 1);
        mbar_init(&bar_mma, 1);
        mbar_init(&bar_mma3, 1);
        mbar_init(&bar_hfree, 1);
        fence_mbar_init();
    }
    {
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (int i = tid; i < 32768 / 16; i += kWsThreads) st_shared_v4(sA + i * 16, z);
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = tmem_slot;
    uint32_t ph_z = 0, ph_w = 0, ph_mma = 0, ph_mma3 = 0, ph_hfree = 0;
    int loaded_stack = -1;
    bool z_prefetched = false;
    int gstep = 0;                                      // steps processed so far by this CTA
#ifdef NRX_PHASE_TIMING
    __shared__ unsigned long long s_phase[32];
    if (tid == 32)
        for (int i = 0; i < 32; ++i) s_phase[i] = 0;
    long long tick_last = clock64();
#endif

    // ---- main-warp task mappings (as in nrx_stack_kernel) -----------------------------------------
    const int ph = tid >> 5, qh = tid & 31;
    const bool act_1 = tid < 3 * 56;
    const int seg1 = tid / 56, p1 = (tid % 56) >> 3, q1 = tid & 7;
    const int q4 = warp & 3, hc = (warp >> 2) & 1;
    const int erow = q4 * 32 + lane;
    const int efl = erow / kT;
    // ---- helper-warp mappings -----------------------------------------------------------------------
    const int htid = tid - kWsMain;                     // 0..127 for helpers
    const int co_g = htid & 7, co_rr = htid >> 3;       // copy-out: (row co_rr + 16*it, 8-channel group), it = 0..7

    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const int bu = item / p.n_chunks, cj = item - bu * p.n_chunks;
        const int c0 = int((long long)cj * p.F / p.n_chunks), c1 = int((long long)(cj + 1) * p.F / p.n_chunks);
        const int nsteps = (c1 - c0 + kRunIn + kStepF - 1) / kStepF;
        const int stack = p.stack_index ? p.stack_index[bu] : p.default_stack;
        if (stack != loaded_stack) {                    // block-uniform; helpers arrive once they have drained
            __syncthreads();
            if (tid == 0) {
                mbar_arrive_expect_tx(&bar_w, L::kBlob);
                bulk_g2s(sW, p.wblob + size_t(stack) * L::kBlob, L::kBlob, &bar_w);
            }
            mbar_wait(&bar_w, ph_w);
            ph_w ^= 1;
            loaded_stack = stack;
        }

        if (!is_main) {
            // =============================== HELPER WARPS ===============================
            const int u = bu % p.U;
            for (int k = 0; k < nsteps; ++k) {
                const int b = c0 - kRunIn + kStepF * k;
                // what the copy-out needs from global memory: issued before waiting for the GEMM
                uint4 co_old[8];
                float2 co_pe[8];
                bool co_ok[8];
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rr = co_rr + 16 * it, fl = rr / kT, t = rr - fl * kT, f = b + fl;
                    co_ok[it] = rr < kTileRows && f >= c0 && f < c1;
                    co_old[it] = make_uint4(0, 0, 0, 0);
                    co_pe[it] = make_float2(0.f, 0.f);
                    if (co_ok[it]) {
                        if constexpr (MODE == kStackUpdate) {
                            const size_t grow = (size_t(bu) * p.F + f) * kT + t;
                            co_old[it] = __ldg(reinterpret_cast<const uint4*>(p.s_in + grow * 64 + co_g * 8));
                        } else if (8 * co_g + 8 > p.d_s) {
                            co_pe[it] = __ldg(reinterpret_cast<const float2*>(p.pos_enc + ((size_t(u) * p.F + f) * kT + t) * 2));
                        }
                    }
                }
                mbar_wait(&bar_mma3, ph_mma3);
                ph_mma3 ^= 1;
                tc_fence_after_sync();
                {   // accumulator [128 x 64] + bias -> fp32 staging in the hidden-tile region
                    //   (16 chunks of 4 floats per row, chunk ^ (row & 7))
                    float v[2][32];
                    tmem_ld32(tmem_addr(tbase + 128, q4 * 32, 0), v[0]);
                    tmem_ld32(tmem_addr(tbase + 128, q4 * 32, 32), v[1]);
                    const float* b3 = sBias + 256;
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 2; ++c)
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const int c4 = (c * 32 + j) >> 2;
                            const float4 bb = *reinterpret_cast<const float4*>(b3 + c * 32 + j);
                            float4 o;
                            o.x = v[c][j] + bb.x;
                            o.y = v[c][j + 1] + bb.y;
                            o.z = v[c][j + 2] + bb.z;
                            o.w = v[c][j + 3] + bb.w;
                            *reinterpret_cast<float4*>(sH + erow * 256 + ((c4 ^ (erow & 7)) << 4)) = o;
                        }
                }
                tc_fence_before_sync();
                bar_help();                             // staging complete, TMEM columns read
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    if (!co_ok[it]) continue;
                    const int rr = co_rr + 16 * it, fl = rr / kT, t = rr - fl * kT;
                    const float4 o0 = *reinterpret_cast<const float4*>(sH + rr * 256 + (((2 * co_g) ^ (rr & 7)) << 4));
                    const float4 o1 = *reinterpret_cast<const float4*>(sH + rr * 256 + (((2 * co_g + 1) ^ (rr & 7)) << 4));
                    float a[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
                    if constexpr (MODE == kStackUpdate) {       // s <- s + update (:266)
                        const uint32_t ow[4] = {co_old[it].x, co_old[it].y, co_old[it].z, co_old[it].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 of = __half22float2(u2h(ow[e]));
                            a[2 * e] += of.x;
                            a[2 * e + 1] += of.y;
                        }
                    } else if (8 * co_g + 8 > p.d_s) {          // append the positional encoding
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
                    const size_t grow = (size_t(bu) * p.F + (b + fl)) * kT + t;
                    *reinterpret_cast<uint4*>(p.s_out + grow * 64 + co_g * 8) = pk;
                }
                bar_help();                             // every helper has finished reading the staging
                if (htid == 0) mbar_arrive(&bar_hfree);
            }
            continue;
        }

        // ================================= MAIN WARPS =================================
        const int bu_a = p.pair_agg ? (bu ^ 1) : bu;
        const bool a_live = !p.pair_agg || p.active_tx[bu_a] != 0.f;
        {
            const uint4 z = make_uint4(0, 0, 0, 0);
            for (int i = tid; i < kCarryRows * kHRow / 16; i += kWsMain) {
                st_shared_v4(sC1 + i * 16, z);
                st_shared_v4(sC2 + i * 16, z);
            }
        }
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
            constexpr int V = kT * L::kZRow / 16;
            for (int fi = 0; fi < kStepF + 2; ++fi) {
                const int f = zf0 + fi;
                if (f >= 0 && f < p.F) continue;
                for (int i = tid; i < V; i += kWsMain) {
                    st_shared_v4(sZ + fi * kT * L::kZRow + i * 16, z);
                    if constexpr (MODE == kStackUpdate) st_shared_v4(sZ + L::kZArr + fi * kT * L::kZRow + i * 16, z);
                }
            }
        };
        auto stage_z = [&](int zf0) { stage_z_of(bu, bu_a, a_live, zf0); };
        if constexpr (MODE == kStackUpdate) {
            if (!a_live)
                for (int i = tid; i < L::kZArr / 16; i += kWsMain) st_shared_v4(sZ + i * 16, make_uint4(0, 0, 0, 0));
        }

        auto epi_hidden = [&](const float* bias, int f_out0) {
            const int f = f_out0 + efl;
            const bool in_grid = f >= 0 && f < p.F;
            const bool warp_oob = !__all_sync(0xffffffffu, in_grid || erow >= kTileRows);
            float v[2][32];
            tmem_ld32(tmem_addr(tbase, q4 * 32, hc * 64), v[0]);
            tmem_ld32(tmem_addr(tbase, q4 * 32, hc * 64 + 32), v[1]);
            tmem_ld_wait();
            if (erow < kTileRows) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int col = hc * 64 + c * 32;
                    uint4 o[4];
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        const float4 b0 = *reinterpret_cast<const float4*>(bias + col + j);
                        const float4 b1 = *reinterpret_cast<const float4*>(bias + col + j + 4);
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

        uint2 tail[2][2];
        auto load_taps = [&](uint2(&kk)[9], int tap_off, int kp, int q) {
#pragma unroll
            for (int i = 0; i < 9; ++i) kk[i] = lds64(sW + tap_off + (i * kp + q * 4) * 2);
        };
        // 128-channel depthwise pass: warps 0-6 = one symbol pair each, 9 output subcarriers
        auto dw128 = [&](auto rs_tag, const uint8_t* carry, const uint8_t* fresh, int tap_off) {
            constexpr int RS = decltype(rs_tag)::value;
            if (warp < 7) {
                uint2 kk[9];
                load_taps(kk, tap_off, 128, qh);
                dw_slide<RS, kStepF>(carry, fresh, 2 * ph, ph > 0, ph < 6, kk, sA + (qh >> 4) * 16384 + (qh & 1) * 8,
                                     (qh >> 1) & 7, 0, tail);
            }
        };
        auto dw_hidden = [&](const uint8_t* carry, int tap_off) {
            dw128(std::integral_constant<int, kHRow>{}, carry + qh * 8, sH + qh * 8, tap_off);
        };
        auto save_carry = [&](uint8_t* carry) {
            if (warp < 7) {
#pragma unroll
                for (int ci = 0; ci < 2; ++ci)
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        sts64(carry + ((ci * kT) + 2 * ph + e) * kHRow + qh * 8, tail[ci][e]);
            }
        };
        auto issue_mma = [&](int w_off, int K) {
            if (tid == 0) {
                tc_fence_after_sync();
                umma_gemm_k(tbase, smem_u32(sA), 16384, smem_u32(sW + w_off), 128 * 128, K, umma_idesc_f16(128, 128), false);
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
        for (int k = 0; k < nsteps; ++k, ++gstep) {
            const int b = c0 - kRunIn + kStepF * k;
            NRX_TICK(21);
            mbar_wait(&bar_z, ph_z);
            ph_z ^= 1;
            bar_main();                                 // zero-filled rows / carries visible
            NRX_TICK(0);

            // ================= layer 1 =================
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
            bar_main();                                 // A complete; Z window free
            NRX_TICK(2);
            issue_mma(L::oPw1, L::KP1);
            if (k + 1 < nsteps) {
                stage_z(b + kStepF + 1);
            } else if (item + int(gridDim.x) < p.num_items) {
                const int nitem = item + int(gridDim.x);
                const int nbu = nitem / p.n_chunks, ncj = nitem - nbu * p.n_chunks;
                const int nstack = p.stack_index ? p.stack_index[nbu] : p.default_stack;
                if (nstack == loaded_stack) {
                    const int nbu_a = p.pair_agg ? (nbu ^ 1) : nbu;
                    const bool na_live = !p.pair_agg || p.active_tx[nbu_a] != 0.f;
                    stage_z_of(nbu, nbu_a, na_live, int((long long)ncj * p.F / p.n_chunks) - kRunIn + 1);
                    z_prefetched = true;
                }
            }
            NRX_TICK(3);
            wait_mma();
            NRX_TICK(4);
            if (gstep > 0) {                            // helpers done with the previous step's staging (hidden-tile region)
                mbar_wait(&bar_hfree, ph_hfree);
                ph_hfree ^= 1;
            }
            NRX_TICK(22);
            epi_hidden(sBias, b + 2);
            NRX_TICK(5);
            tc_fence_before_sync();
            bar_main();
            NRX_TICK(6);

            // ================= layer 2 =================
            dw_hidden(sC1, L::oTap2);
            NRX_TICK(7);
            fence_proxy_async_smem();
            tc_fence_before_sync();
            bar_main();
            NRX_TICK(8);
            issue_mma(L::oPw2, 128);
            save_carry(sC1);
            NRX_TICK(9);
            wait_mma();
            NRX_TICK(10);
            epi_hidden(sBias + 128, b + 1);
            NRX_TICK(11);
            tc_fence_before_sync();
            bar_main();
            NRX_TICK(12);

            // ================= layer 3 (output side handled by the helper warps) =================
            dw_hidden(sC2, L::oTap3);
            NRX_TICK(13);
            fence_proxy_async_smem();
            tc_fence_before_sync();
            bar_main();
            NRX_TICK(14);
            if (tid == 0) {
                tc_fence_after_sync();
                umma_gemm_k(tbase + 128, smem_u32(sA), 16384, smem_u32(sW + L::oPw3), 64 * 128, 128, umma_idesc_f16(128, 64), false);
                umma_commit(&bar_mma3);
            }
            save_carry(sC2);
            NRX_TICK(15);
            mbar_wait(&bar_mma3, ph_mma3);              // the A operand may be overwritten by the next depthwise pass
            ph_mma3 ^= 1;
            NRX_TICK(16);
        }
    }
#ifdef NRX_PHASE_TIMING
    if (blockIdx.x == 0 && tid == 32)
        for (int i = 0; i < 32; ++i) g_phase_cycles[i] += s_phase[i];
#endif
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 256);
}

}  // namespace nrx

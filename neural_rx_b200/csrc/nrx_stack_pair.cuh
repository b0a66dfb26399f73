// nrx_stack_pair.cuh — CTA-pair variant of the fused sep-conv stack kernel (two users only).
//
// The two users of a slot run the same chunk of subcarriers in the two CTAs of a cluster
// (rank = user).  Each CTA keeps its own window / hidden tile / A operand exactly as
// nrx_stack_kernel does, but the pointwise GEMMs are issued by the leader CTA for BOTH CTAs with
// tcgen05.mma.cta_group::2 (M = 256): every CTA only holds HALF of each weight matrix (rank r has
// output channels [r*N/2, (r+1)*N/2) of every layer), the MMA takes half the time and each SM reads
// half of the B operand from its shared memory.  Same arithmetic, bit-identical results.
//
// Protocol per GEMM (validated stand-alone in tools/umma_pair_probe.cu):
//   both CTAs : depthwise pass -> fence.proxy.async -> __syncthreads -> thread 0 arrives on the
//               LEADER's bar_ready (count 2, mbarrier.arrive.release.cluster)
//   leader    : thread 0 waits bar_ready, issues K/16 MMAs, tcgen05.commit multicast to bar_mma of
//               both CTAs
//   both CTAs : wait their own bar_mma, run their epilogue from their own TMEM
// A CTA cannot run ahead of its partner by more than one GEMM: it needs the commit to go on, and
// the leader needs the partner's arrival to issue.
#pragma once
#include "nrx_stack.cuh"

namespace nrx {

template <int MODE>
struct StackPairSmem {
    static constexpr int KP1 = MODE == kStackInit ? 32 : 128;
    static constexpr int kPw1 = ((KP1 + 63) / 64) * 64 * 128;          // per-rank halves: [K slabs][N/2 rows][128 B]
    static constexpr int kPw2 = 2 * 64 * 128, kPw3 = 2 * 32 * 128;
    static constexpr int kTap1 = 9 * KP1 * 2, kTap = 9 * 128 * 2;
    static constexpr int oPw1 = 0, oPw2 = kPw1, oPw3 = oPw2 + kPw2;
    static constexpr int oTap1 = oPw3 + kPw3, oTap2 = oTap1 + kTap1, oTap3 = oTap2 + kTap;
    static constexpr int oBias = oTap3 + kTap;                         // fp32 [128 | 128 | 64] (all channels)
    static constexpr int kBlob = oBias + (128 + 128 + 64) * 4;         // per (stack, rank)
    static constexpr int offA = align_up_c(kBlob, 1024);
    static constexpr int offH = offA + 32768;
    static constexpr int offC1 = offH + align_up_c(kTileRows * kHRow, 1024);
    static constexpr int offC2 = offC1 + kCarryRows * kHRow;
    static constexpr int offZ = offC2 + kCarryRows * kHRow;
    static constexpr int kZRow = MODE == kStackInit ? 64 : 128;
    static constexpr int kZArr = kHaloRows * kZRow;
    static constexpr int kZ = MODE == kStackInit ? kZArr : 2 * kZArr;
    static constexpr int kTotal = offZ + kZ + 1024;
};

// StackParams as for nrx_stack_kernel, with  num_items = pairs (slot, chunk)  and  wblob = [stack][rank].
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kStackThreads, 1) nrx_stack_pair_kernel(StackParams p) {
    using L = StackPairSmem<MODE>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sW = smem;
    uint8_t* sA = smem + L::offA;
    uint8_t* sC1 = smem + L::offC1;
    uint8_t* sH = smem + L::offH;
    uint8_t* sC2 = smem + L::offC2;
    uint8_t* sZ = smem + L::offZ;
    const float* sBias = reinterpret_cast<const float*>(sW + L::oBias);
    __shared__ uint64_t bar_z, bar_w, bar_mma, bar_ready;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();            // = user index inside the slot
    if (tid == 0) {
        mbar_init(&bar_z, 1);
        mbar_init(&bar_w, 1);
        mbar_init(&bar_mma, 1);
        mbar_init(&bar_ready, 2);                       // one arrival per CTA of the pair (the leader's copy is used)
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc_pair(&tmem_slot, 128);
    {
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (int i = tid; i < 32768 / 16; i += kStackThreads) st_shared_v4(sA + i * 16, z);
    }
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();                                 // barriers initialised and TMEM allocated in both CTAs
    tc_fence_after_sync();
    const uint32_t tbase = tmem_slot;
    const uint32_t ready_leader = mapa_shared(smem_u32(&bar_ready), 0);
    uint32_t ph_z = 0, ph_w = 0, ph_mma = 0, ph_ready = 0;
    int loaded_stack = -1;
    bool z_prefetched = false, w_pending = false;

    const int ph = tid >> 5, qh = tid & 31;
    const bool act_1 = tid < 3 * 56;
    const int seg1 = tid / 56, p1 = (tid % 56) >> 3, q1 = tid & 7;
    const int co_g = tid & 7, co_rr = tid >> 3;
    int co_fl[4], co_t[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        co_fl[it] = (co_rr + 32 * it) / kT;
        co_t[it] = (co_rr + 32 * it) - co_fl[it] * kT;
    }
    const int q4 = warp & 3, hc = warp >> 2;
    const int erow = q4 * 32 + lane;
    const int efl = erow / kT;

    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    for (int item = cluster_id; item < p.num_items; item += n_clusters) {
        const int slot = item / p.n_chunks, cj = item - slot * p.n_chunks;
        const int bu = slot * 2 + int(rank);
        const int c0 = int((long long)cj * p.F / p.n_chunks), c1 = int((long long)(cj + 1) * p.F / p.n_chunks);
        const int nsteps = (c1 - c0 + kRunIn + kStepF - 1) / kStepF;
        const int stack = stack_of(p, bu);   // host guarantees: same for both users
        if (stack != loaded_stack) {
            __syncthreads();
            if (tid == 0) {
                mbar_arrive_expect_tx(&bar_w, L::kBlob);
                bulk_g2s(sW, p.wblob + (size_t(stack) * 2 + rank) * L::kBlob, L::kBlob, &bar_w);
            }
            w_pending = true;
            loaded_stack = stack;
        }
        {
            const uint4 z = make_uint4(0, 0, 0, 0);
            for (int i = tid; i < kCarryRows * kHRow / 16; i += kStackThreads) {
                st_shared_v4(sC1 + i * 16, z);
                st_shared_v4(sC2 + i * 16, z);
            }
        }
        auto stage_z_of = [&](int zbu, int zf0) {
            const int flo = max(zf0, 0), fhi = min(zf0 + kStepF + 2, p.F);
            const int nrow = max(fhi - flo, 0) * kT;
            if (tid == 0) {
                const size_t grow = (size_t(zbu) * p.F + flo) * kT;
                const int so = (flo - zf0) * kT * L::kZRow;
                if constexpr (MODE == kStackInit) {
                    mbar_arrive_expect_tx(&bar_z, uint32_t(nrow) * 64u);
                    if (nrow) bulk_g2s(sZ + so, reinterpret_cast<const uint8_t*>(p.z0) + grow * 64, uint32_t(nrow) * 64u, &bar_z);
                } else {
                    mbar_arrive_expect_tx(&bar_z, uint32_t(nrow) * 256u);
                    if (nrow) {
                        bulk_g2s(sZ + so, reinterpret_cast<const uint8_t*>(p.a_in) + grow * 128, uint32_t(nrow) * 128u, &bar_z);
                        bulk_g2s(sZ + L::kZArr + so, reinterpret_cast<const uint8_t*>(p.s_in) + grow * 128, uint32_t(nrow) * 128u, &bar_z);
                    }
                }
            }
            const uint4 z = make_uint4(0, 0, 0, 0);
            constexpr int V = kT * L::kZRow / 16;
            for (int fi = 0; fi < kStepF + 2; ++fi) {
                const int f = zf0 + fi;
                if (f >= 0 && f < p.F) continue;
                for (int i = tid; i < V; i += kStackThreads) {
                    st_shared_v4(sZ + fi * kT * L::kZRow + i * 16, z);
                    if constexpr (MODE == kStackUpdate) st_shared_v4(sZ + L::kZArr + fi * kT * L::kZRow + i * 16, z);
                }
            }
        };

        auto epi_hidden = [&](const float* bias, int f_out0) {
            const int f = f_out0 + efl;
            const bool in_grid = f >= 0 && f < p.F;
            const bool warp_oob = !__all_sync(0xffffffffu, in_grid || erow >= kTileRows);
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
        auto load_taps = [&](uint2(&kk)[9], int tap_off, int kp, int q) {
#pragma unroll
            for (int i = 0; i < 9; ++i) kk[i] = lds64(sW + tap_off + (i * kp + q * 4) * 2);
        };
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
                const uint8_t* c8 = fresh + (kStepF - 3) * (kT * RS);
#pragma unroll
                for (int pp = 0; pp < 7; ++pp)
                    dw_slide<RS, 1>(c8, c8 + 2 * (kT * RS), 2 * pp, pp > 0, pp < 6, kk, a_thr, (qh >> 1) & 7,
                                    (kStepF - 1) * kT, tail7[pp]);
            }
        };
        auto dw_hidden = [&](const uint8_t* carry, int tap_off) {
            dw128(std::integral_constant<int, kHRow>{}, carry + qh * 8, sH + qh * 8, tap_off);
        };
        auto save_carry = [&](uint8_t* carry) {
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
        // after the local post-depthwise barrier: report "A operand ready", the leader issues for the pair
        auto issue_pair = [&](int w_off, int b_slab_bytes, int K, int N) {
            if (tid == 0) {
                mbar_arrive_cluster(ready_leader);
                if (rank == 0) {
                    mbar_wait_cluster(&bar_ready, ph_ready);
                    tc_fence_after_sync();
                    umma_gemm_k_pair(tbase, smem_u32(sA), 16384, smem_u32(sW + w_off), b_slab_bytes, K, umma_idesc_f16(256, N));
                    umma_commit_pair(&bar_mma, 0b11);
                }
            }
            ph_ready ^= 1;
        };
        auto wait_mma = [&]() {
            mbar_wait_cluster(&bar_mma, ph_mma);
            ph_mma ^= 1;
            tc_fence_after_sync();
        };

        if (!z_prefetched) stage_z_of(bu, c0 - kRunIn + 1);
        z_prefetched = false;
        if (w_pending) {
            mbar_wait(&bar_w, ph_w);
            ph_w ^= 1;
            w_pending = false;
        }
        for (int k = 0; k < nsteps; ++k) {
            const int b = c0 - kRunIn + kStepF * k;
            mbar_wait(&bar_z, ph_z);
            ph_z ^= 1;
            __syncthreads();

            // ================= layer 1 =================
            if constexpr (MODE == kStackInit) {
                if (act_1) {
                    uint2 kk[9];
                    load_taps(kk, L::oTap1, 32, q1);
                    const uint8_t* zc = sZ + seg1 * 3 * kT * 64 + q1 * 8;
                    uint2 unused[2][2];
                    dw_slide<64, 3>(zc, zc + 2 * kT * 64, 2 * p1, p1 > 0, p1 < 6, kk, sA + (q1 & 1) * 8, q1 >> 1,
                                    seg1 * 3 * kT, unused);
                }
            } else {
                const uint8_t* zc = sZ + (qh >> 4) * L::kZArr + (qh & 15) * 8;
                dw128(std::integral_constant<int, 128>{}, zc, zc + 2 * kT * 128, L::oTap1);
            }
            fence_proxy_async_smem();
            tc_fence_before_sync();
            __syncthreads();
            issue_pair(L::oPw1, 64 * 128, L::KP1, 128);
            if (k + 1 < nsteps) {
                stage_z_of(bu, b + kStepF + 1);
            } else if (item + n_clusters < p.num_items) {
                const int nitem = item + n_clusters;
                const int nslot = nitem / p.n_chunks, ncj = nitem - nslot * p.n_chunks;
                const int nbu = nslot * 2 + int(rank);
                const int nstack = stack_of(p, nbu);
                if (nstack == loaded_stack) {
                    stage_z_of(nbu, int((long long)ncj * p.F / p.n_chunks) - kRunIn + 1);
                    z_prefetched = true;
                }
            }
            wait_mma();
            epi_hidden(sBias, b + 2);
            tc_fence_before_sync();
            __syncthreads();

            // ================= layer 2 =================
            dw_hidden(sC1, L::oTap2);
            fence_proxy_async_smem();
            tc_fence_before_sync();
            __syncthreads();
            issue_pair(L::oPw2, 64 * 128, 128, 128);
            save_carry(sC1);
            const int u = int(rank);
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
            wait_mma();
            epi_hidden(sBias + 128, b + 1);
            tc_fence_before_sync();
            __syncthreads();

            // ================= layer 3 =================
            dw_hidden(sC2, L::oTap3);
            fence_proxy_async_smem();
            tc_fence_before_sync();
            __syncthreads();
            issue_pair(L::oPw3, 32 * 128, 128, 64);
            save_carry(sC2);
            wait_mma();
            {
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
            tc_fence_before_sync();
            __syncthreads();
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                if (!co_ok[it]) continue;
                const int rr = co_rr + 32 * it;
                const float4 o0 = *reinterpret_cast<const float4*>(sA + rr * 256 + (((2 * co_g) ^ (rr & 7)) << 4));
                const float4 o1 = *reinterpret_cast<const float4*>(sA + rr * 256 + (((2 * co_g + 1) ^ (rr & 7)) << 4));
                float a[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
                if constexpr (MODE == kStackUpdate) {
                    const uint32_t ow[4] = {co_old[it].x, co_old[it].y, co_old[it].z, co_old[it].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 of = __half22float2(u2h(ow[e]));
                        a[2 * e] += of.x;
                        a[2 * e + 1] += of.y;
                    }
                } else if (8 * co_g + 8 > p.d_s) {
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
                const size_t grow = (size_t(bu) * p.F + (b + co_fl[it])) * kT + co_t[it];
                *reinterpret_cast<uint4*>(p.s_out + grow * 64 + co_g * 8) = pk;
            }
            __syncthreads();
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();                                 // the partner may still be reading its accumulator
    if (warp == 0) tmem_dealloc_pair(tbase, 128);
}

}  // namespace nrx

// umma_pair_probe.cu — known-answer test of the CTA-pair (tcgen05 cta_group::2) building block that
// the next stack-kernel generation needs (DESIGN.md §4.2 "Next step"): two CTAs of a cluster, each
// with its OWN thread-written A tile (128 rows) and HALF of the B operand (N/2 weight rows) in
// shared memory; the leader CTA issues one M=256 MMA sequence for the pair, the commit is multicast
// to a barrier in both CTAs, and each CTA reads its 128 x N accumulator from its own TMEM.
// What it establishes: (1) the allocation / commit / remote-arrive protocol, (2) which CTA's B half
// lands in which accumulator columns, (3) that per-CTA smem only needs N/2 rows of the weights.
//
// build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/umma_pair_probe tools/umma_pair_probe.cu
// run  :  tools/umma_pair_probe
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../neural_rx_b200/csrc/sm100_prims.cuh"

using namespace nrx;

// the pair primitives under test live in sm100_prims.cuh (shared with nrx_stack_pair.cuh)
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(99); } } while (0)

// cluster of 2 CTAs x 128 threads.  A: [256][K] fp16 (rows 128*rank.. belong to CTA rank);
// Bimg: per-CTA image [K/64 slabs][N/2 rows][128 B] (rank-major); D: [256][N] fp32
template <int K, int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_kernel(const __half* __restrict__ A, const uint8_t* __restrict__ Bimg, float* __restrict__ D) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int KS = K / 64, NH = N / 2;
    constexpr int A_BYTES = KS * 128 * 128, B_BYTES = KS * NH * 128;
    uint8_t* sA = smem;
    uint8_t* sB = smem + A_BYTES;
    __shared__ uint64_t bar_load, bar_ready, bar_mma;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    if (tid == 0) {
        mbar_init(&bar_load, 1);
        mbar_init(&bar_ready, 2);        // one arrival per CTA of the pair (leader's copy is the one used)
        mbar_init(&bar_mma, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc_pair(&tmem_slot, 256);
    tc_fence_before_sync();
    cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tbase = tmem_slot;

    if (tid == 0) {
        mbar_arrive_expect_tx(&bar_load, B_BYTES);
        bulk_g2s(sB, Bimg + size_t(rank) * B_BYTES, B_BYTES, &bar_load);
    }
    for (int idx = tid; idx < 128 * K; idx += 128) {
        const int r = idx / K, k = idx % K;
        *reinterpret_cast<__half*>(sA + (k / 64) * (128 * 128) + sw128_offset(r, k % 64)) = A[(size_t(rank) * 128 + r) * K + k];
    }
    fence_proxy_async_smem();
    mbar_wait(&bar_load, 0);
    __syncthreads();
    // both CTAs report "operands in place" to the LEADER's barrier
    if (tid == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&bar_ready), 0));
    if (rank == 0 && tid == 0) {
        mbar_wait_cluster(&bar_ready, 0);
        tc_fence_after_sync();
        umma_gemm_k_pair(tbase, smem_u32(sA), 128 * 128, smem_u32(sB), NH * 128, K, umma_idesc_f16(256, N));
        umma_commit_pair(&bar_mma, 0b11);
    }
    mbar_wait_cluster(&bar_mma, 0);
    tc_fence_after_sync();
    const int row = warp * 32 + lane;
    for (int c = 0; c < N; c += 32) {
        float v[32];
        tmem_ld32(tmem_addr(tbase, warp * 32, c), v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) D[(size_t(rank) * 128 + row) * N + c + j] = v[j];
    }
    tc_fence_before_sync();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc_pair(tbase, 256);
}

static float h2f(__half h) { return __half2float(h); }

template <int K, int N>
static int run_case(const char* name) {
    srand(4321 + K + N);
    std::vector<__half> A(256 * K), W(size_t(N) * K);
    for (auto& v : A) v = __float2half((rand() % 2001 - 1000) / 500.f);
    for (auto& v : W) v = __float2half((rand() % 2001 - 1000) / 1000.f);
    constexpr int KS = K / 64, NH = N / 2;
    // hypothesis: CTA rank r supplies weight rows [r*N/2, (r+1)*N/2)
    std::vector<uint8_t> img(size_t(2) * KS * NH * 128, 0);
    for (int r = 0; r < 2; ++r)
        for (int n = 0; n < NH; ++n)
            for (int k = 0; k < K; ++k)
                *reinterpret_cast<__half*>(&img[size_t(r) * KS * NH * 128 + size_t(k / 64) * NH * 128 + sw128_offset(n, k % 64)]) =
                    W[size_t(r * NH + n) * K + k];
    __half* dA; uint8_t* dB; float* dD;
    CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, img.size())); CK(cudaMalloc(&dD, size_t(256) * N * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, img.data(), img.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xFF, size_t(256) * N * 4));
    const int smem = 1024 + KS * 128 * 128 + KS * NH * 128;
    auto kern = pair_kernel<K, N>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<2, 128, smem>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-24s CUDA ERROR %s\n", name, cudaGetErrorString(e)); return 1; }
    std::vector<float> D(size_t(256) * N);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double err_direct = 0, err_swapped = 0, mx = 0;
    for (int r = 0; r < 256; ++r)
        for (int n = 0; n < N; ++n) {
            double s = 0, s2 = 0;
            const int n2 = (n + NH) % N;
            for (int k = 0; k < K; ++k) {
                s += double(h2f(A[size_t(r) * K + k])) * h2f(W[size_t(n) * K + k]);
                s2 += double(h2f(A[size_t(r) * K + k])) * h2f(W[size_t(n2) * K + k]);
            }
            err_direct = fmax(err_direct, fabs(s - D[size_t(r) * N + n]));
            err_swapped = fmax(err_swapped, fabs(s2 - D[size_t(r) * N + n]));
            mx = fmax(mx, fabs(s));
        }
    const bool ok = err_direct <= 2e-3 * fmax(mx, 1.0);
    printf("%-24s K=%3d N=%3d %s  max|err| columns-as-ranked %.3e, halves-swapped %.3e (max|ref| %.2f)\n", name, K, N,
           ok ? "PASS" : "FAIL", err_direct, err_swapped, mx);
    if (!ok) printf("   D[0][0..3] = %.4f %.4f %.4f %.4f, D[128][0] = %.4f, D[0][%d] = %.4f\n", D[0], D[1], D[2], D[3],
                    D[size_t(128) * N], NH, D[NH]);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return ok ? 0 : 1;
}

int main() {
    int fails = 0;
    fails += run_case<64, 128>("pair gemm 64 -> 128");
    fails += run_case<128, 128>("pair gemm 128 -> 128");
    fails += run_case<128, 64>("pair gemm 128 -> 64");
    printf("umma_pair_probe: %d failure(s)\n", fails);
    return fails;
}

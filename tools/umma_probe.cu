// umma_probe.cu — stand-alone known-answer test of the tcgen05 building block used by every
// GEMM-bearing kernel in this repo (sm100_prims.cuh): thread-written 128B-swizzled K-major A
// tile, bulk-copied pre-swizzled B image, K/16 tcgen05.mma into TMEM, tcgen05.commit ->
// mbarrier, tcgen05.ld epilogue; plus a chained second GEMM whose A operand is produced by the
// first epilogue (the pattern of the aggregation MLP / readout heads).
//
// build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/umma_probe tools/umma_probe.cu
// run  :  tools/umma_probe            (prints PASS/FAIL per case, exit code = #failures)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../neural_rx_b200/csrc/sm100_prims.cuh"

using namespace nrx;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(99); } } while (0)

// one CTA, 128 threads.  D1 = A*B1^T ; if CHAIN: A2 = fp16(relu(D1)) (N1 columns), D2 = A2*B2^T
template <int K, int N1, bool CHAIN, int N2>
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __half* __restrict__ A, const uint8_t* __restrict__ B1img,
             const uint8_t* __restrict__ B2img, float* __restrict__ D1, float* __restrict__ D2) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int KS = (K + 63) / 64;                 // A slabs
    constexpr int A_BYTES = KS * 128 * 128;
    constexpr int B1_BYTES = KS * N1 * 128;
    constexpr int K2S = (N1 + 63) / 64;
    constexpr int A2_BYTES = CHAIN ? K2S * 128 * 128 : 0;
    constexpr int B2_BYTES = CHAIN ? K2S * N2 * 128 : 0;
    uint8_t* sA = smem;
    uint8_t* sB1 = sA + A_BYTES;
    uint8_t* sA2 = sB1 + ((B1_BYTES + 1023) / 1024) * 1024;
    uint8_t* sB2 = sA2 + A2_BYTES;
    __shared__ uint64_t bar_load, bar_mma;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr uint32_t TM_COLS = 512;
    if (warp == 0) tmem_alloc(&tmem_slot, TM_COLS);
    if (tid == 0) {
        mbar_init(&bar_load, 1);
        mbar_init(&bar_mma, 1);
        fence_mbar_init();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = tmem_slot;

    if (tid == 0) {
        mbar_arrive_expect_tx(&bar_load, B1_BYTES + B2_BYTES);
        bulk_g2s(sB1, B1img, B1_BYTES, &bar_load);
        if (CHAIN) bulk_g2s(sB2, B2img, B2_BYTES, &bar_load);
    }
    // thread-written A tile (zero-fill K padding up to the slab)
    for (int idx = tid; idx < 128 * KS * 64; idx += 128) {
        const int r = idx / (KS * 64), k = idx % (KS * 64);
        const __half v = (k < K) ? A[r * K + k] : __float2half(0.f);
        *reinterpret_cast<__half*>(sA + (k / 64) * (128 * 128) + sw128_offset(r, k % 64)) = v;
    }
    fence_proxy_async_smem();
    mbar_wait(&bar_load, 0);
    __syncthreads();

    if (tid == 0) {
        tc_fence_after_sync();
        umma_gemm_k(tbase, smem_u32(sA), 128 * 128, smem_u32(sB1), N1 * 128, K, umma_idesc_f16(128, N1), false);
        umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    tc_fence_after_sync();

    const int row = warp * 32 + lane;
    if constexpr (N1 % 32 == 0) {
        for (int c = 0; c < N1; c += 32) {
            float v[32];
            tmem_ld32(tmem_addr(tbase, warp * 32, c), v);
            tmem_ld_wait();
            for (int j = 0; j < 32; ++j) {
                D1[row * N1 + c + j] = v[j];
                if (CHAIN) {
                    const int k = c + j;
                    *reinterpret_cast<__half*>(sA2 + (k / 64) * (128 * 128) + sw128_offset(row, k % 64)) =
                        __float2half(fmaxf(v[j], 0.f));
                }
            }
        }
    } else {
        for (int c = 0; c < N1; c += 16) {
            float v[16];
            tmem_ld16(tmem_addr(tbase, warp * 32, c), v);
            tmem_ld_wait();
            for (int j = 0; j < 16; ++j) D1[row * N1 + c + j] = v[j];
        }
    }
    if constexpr (CHAIN) {
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after_sync();
            umma_gemm_k(tbase + 256, smem_u32(sA2), 128 * 128, smem_u32(sB2), N2 * 128, N1, umma_idesc_f16(128, N2), false);
            umma_commit(&bar_mma);
        }
        mbar_wait(&bar_mma, 1);
        tc_fence_after_sync();
        for (int c = 0; c < N2; c += 16) {
            float v[16];
            tmem_ld16(tmem_addr(tbase + 256, warp * 32, c), v);
            tmem_ld_wait();
            for (int j = 0; j < 16; ++j) D2[row * N2 + c + j] = v[j];
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, TM_COLS);
}

static float h2f(__half h) { return __half2float(h); }

// host: pre-swizzled image of W^T, i.e. rows n, K-major, slabs of 64 k
static std::vector<uint8_t> make_b_image(const std::vector<__half>& W /*[N][K]*/, int N, int K) {
    const int KS = (K + 63) / 64;
    std::vector<uint8_t> img(size_t(KS) * N * 128, 0);
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k)
            *reinterpret_cast<__half*>(&img[size_t(k / 64) * N * 128 + sw128_offset(n, k % 64)]) = W[size_t(n) * K + k];
    return img;
}

template <int K, int N1, bool CHAIN, int N2>
static int run_case(const char* name) {
    srand(1234 + K * 7 + N1);
    std::vector<__half> A(128 * K), W1(size_t(N1) * K), W2(size_t(N2) * N1);
    for (auto& v : A) v = __float2half((rand() % 2001 - 1000) / 500.f);
    for (auto& v : W1) v = __float2half((rand() % 2001 - 1000) / 1000.f);
    for (auto& v : W2) v = __float2half((rand() % 2001 - 1000) / 1000.f);
    auto b1 = make_b_image(W1, N1, K);
    auto b2 = make_b_image(W2, N2, N1);
    __half* dA; uint8_t *dB1, *dB2; float *dD1, *dD2;
    CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB1, b1.size())); CK(cudaMalloc(&dB2, b2.size()));
    CK(cudaMalloc(&dD1, 128 * N1 * 4)); CK(cudaMalloc(&dD2, 128 * N2 * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB1, b1.data(), b1.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB2, b2.data(), b2.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dD1, 0xFF, 128 * N1 * 4)); CK(cudaMemset(dD2, 0xFF, 128 * N2 * 4));
    const int smem = 1024 + ((K + 63) / 64) * 16384 + (((K + 63) / 64) * N1 * 128 + 1023) / 1024 * 1024 +
                     (CHAIN ? ((N1 + 63) / 64) * (16384 + N2 * 128) : 0);
    auto kern = probe_kernel<K, N1, CHAIN, N2>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<1, 128, smem>>>(dA, dB1, dB2, dD1, dD2);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-28s CUDA ERROR %s\n", name, cudaGetErrorString(e)); exit(98); }
    std::vector<float> D1(128 * N1), D2(128 * N2);
    CK(cudaMemcpy(D1.data(), dD1, D1.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(D2.data(), dD2, D2.size() * 4, cudaMemcpyDeviceToHost));
    double e1 = 0, e2 = 0, m1 = 0, m2 = 0;
    std::vector<float> R1(128 * N1);
    for (int r = 0; r < 128; ++r)
        for (int n = 0; n < N1; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += double(h2f(A[r * K + k])) * h2f(W1[size_t(n) * K + k]);
            R1[r * N1 + n] = float(s);
            e1 = fmax(e1, fabs(s - D1[r * N1 + n])); m1 = fmax(m1, fabs(s));
        }
    if (CHAIN)
        for (int r = 0; r < 128; ++r)
            for (int n = 0; n < N2; ++n) {
                double s = 0;
                for (int k = 0; k < N1; ++k)   // device rounds relu(D1) (its own D1) to fp16
                    s += double(h2f(__float2half(fmaxf(D1[r * N1 + k], 0.f)))) * h2f(W2[size_t(n) * N1 + k]);
                e2 = fmax(e2, fabs(s - D2[r * N2 + n])); m2 = fmax(m2, fabs(s));
            }
    const bool ok = (e1 <= 2e-3 * fmax(m1, 1.0)) && (!CHAIN || e2 <= 2e-3 * fmax(m2, 1.0));
    printf("%-28s K=%3d N1=%3d %s max|err1|=%.3e (max|ref|=%.2f)", name, K, N1, ok ? "PASS" : "FAIL", e1, m1);
    if (CHAIN) printf("  N2=%3d max|err2|=%.3e (max|ref|=%.2f)", N2, e2, m2);
    printf("\n");
    if (!ok) {
        printf("   first rows: got %.4f %.4f %.4f %.4f | ref %.4f %.4f %.4f %.4f | row1 got %.4f ref %.4f | row8 got %.4f ref %.4f | row64 got %.4f ref %.4f\n",
               D1[0], D1[1], D1[2], D1[3], R1[0], R1[1], R1[2], R1[3], D1[N1], R1[N1], D1[8 * N1], R1[8 * N1], D1[64 * N1], R1[64 * N1]);
    }
    cudaFree(dA); cudaFree(dB1); cudaFree(dB2); cudaFree(dD1); cudaFree(dD2);
    return ok ? 0 : 1;
}

int main() {
    int fails = 0;
    fails += run_case<64, 64, false, 16>("gemm 64x64");
    fails += run_case<32, 128, false, 16>("gemm K32 (18->128 layer)");
    fails += run_case<128, 128, false, 16>("gemm 128x128");
    fails += run_case<128, 64, false, 16>("gemm 128->64");
    fails += run_case<64, 256, false, 16>("gemm 64->256 (readout hid)");
    fails += run_case<256, 16, false, 16>("gemm 256->16 (readout out)");
    fails += run_case<64, 64, true, 64>("chain 64->64->64 (agg MLP)");
    fails += run_case<64, 128, true, 16>("chain 64->128->16 (readout)");
    printf("umma_probe: %d failure(s)\n", fails);
    return fails;
}

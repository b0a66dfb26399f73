// Issue-rate probe for the depthwise inner loop: HFMA2 with three register operands against HFMA2 whose multiplier
// comes from the constant bank (kernel parameter or __constant__), and FFMA likewise.  Prints cycles per warp
// instruction per SM sub-partition at 1, 2 and 4 warps per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fma_probe tools/fma_probe.cu && /tmp/fma_probe
// Measured on B200 (cycles per warp instruction per sub-partition, 2-4 warps): HFMA2 2.00 in every form (nvcc never
// folds a constant operand into HFMA2), FFMA with three registers 1.66, FFMA with a uniform-register / constant multiplier
// 1.05, HFMA2 + FFMA mixed 2:1 1.87.  So the pipe does ~32 FMA per clock per sub-partition whether they are packed
// fp16 pairs or fp32: fp16x2 buys registers and shared-memory bytes, not FMA rate, and the two forms do not add up.
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>

__constant__ __half2 c_w[16];
struct P { __half2 w[16]; float f[16]; };

template <int MODE>
__global__ void probe(const __grid_constant__ P p, const __half2* gw, __half2* out, long long* cyc, int iters) {
    __half2 acc[8], x[8], w[8];
    float facc[8], fx[8], fw[8];
    for (int i = 0; i < 8; ++i) {
        acc[i] = __float2half2_rn(0.f);
        x[i] = __float2half2_rn(1.f + 0.001f * (threadIdx.x + i));
        w[i] = gw[(threadIdx.x + i) & 15];
        facc[i] = 0.f; fx[i] = 1.f + 0.001f * (threadIdx.x + i); fw[i] = __half2float(__low2half(w[i]));
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) acc[i] = __hfma2(x[i], w[(i + r) & 7], acc[i]);          // 3 registers
                if (MODE == 1) acc[i] = __hfma2(x[i], p.w[(i + r) & 15], acc[i]);       // kernel-parameter constant
                if (MODE == 2) acc[i] = __hfma2(x[i], c_w[(i + 2 * r) & 15], acc[i]);   // __constant__
                if (MODE == 3) facc[i] = fmaf(fx[i], fw[(i + r) & 7], facc[i]);         // FFMA 3 registers
                if (MODE == 4) facc[i] = fmaf(fx[i], p.f[(i + r) & 15], facc[i]);       // FFMA constant
                if (MODE == 5) {                                                         // HFMA2 const + FFMA const mixed 2:1
                    acc[i] = __hfma2(x[i], p.w[(i + r) & 15], acc[i]);
                    if (i & 1) facc[i] = fmaf(fx[i], p.f[(i + r) & 15], facc[i]);
                }
            }
    }
    const long long t1 = clock64();
    __half2 s = acc[0];
    float fs = facc[0];
    for (int i = 1; i < 8; ++i) { s = __hadd2(s, acc[i]); fs += facc[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = __hadd2(s, __float2half2_rn(fs));
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    P p;
    __half2 hw[16];
    for (int i = 0; i < 16; ++i) { hw[i] = __float2half2_rn(0.5f + 0.01f * i); p.w[i] = hw[i]; p.f[i] = 0.5f + 0.01f * i; }
    cudaMemcpyToSymbol(c_w, hw, sizeof hw);
    __half2 *gw, *out; long long* cyc;
    cudaMalloc(&gw, sizeof hw); cudaMemcpy(gw, hw, sizeof hw, cudaMemcpyHostToDevice);
    cudaMalloc(&out, 1024 * sizeof(__half2)); cudaMalloc(&cyc, 8 * sizeof(long long));
    const int iters = 2000;
    const char* names[] = {"HFMA2 3-reg", "HFMA2 x c[param]", "HFMA2 x c[__constant__]", "FFMA 3-reg", "FFMA x c[param]", "HFMA2 c + 0.5 FFMA c"};
    for (int mode = 0; mode < 6; ++mode)
        for (int warps = 4; warps <= 16; warps *= 2) {
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                switch (mode) {
                    case 0: probe<0><<<1, warps * 32>>>(p, gw, out, cyc, iters); break;
                    case 1: probe<1><<<1, warps * 32>>>(p, gw, out, cyc, iters); break;
                    case 2: probe<2><<<1, warps * 32>>>(p, gw, out, cyc, iters); break;
                    case 3: probe<3><<<1, warps * 32>>>(p, gw, out, cyc, iters); break;
                    case 4: probe<4><<<1, warps * 32>>>(p, gw, out, cyc, iters); break;
                    default: probe<5><<<1, warps * 32>>>(p, gw, out, cyc, iters); break;
                }
                cudaDeviceSynchronize();
                cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
            }
            const double instr_per_smsp = double(iters) * 32 * (mode == 5 ? 1.5 : 1.0) * (warps / 4);
            printf("%-26s %2d warps/SMSP: %8lld cycles, %.3f cycles per warp-instruction per SMSP\n", names[mode], warps / 4, h,
                   double(h) / instr_per_smsp);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

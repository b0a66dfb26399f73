// tmem_frag_probe.cu — prints the register <-> (TMEM lane, column) mapping of the 16x256b shape of
// tcgen05.ld / tcgen05.st against the known 32x32b mapping (thread = lane, register = column).
// build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_frag_probe tools/tmem_frag_probe.cu
// run  :  tools/tmem_frag_probe     (exit code = number of mismatches against the expected layout)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../neural_rx_b200/csrc/sm100_prims.cuh"
using namespace nrx;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(99); } } while (0)

// out_ld [32 threads][8 regs]: values seen by ld.16x256b.x2 when TMEM[lane][col] = lane*256 + col (lanes 16..31, cols 0..15)
// out_st [32 lanes][8 cols]  : TMEM contents (read back 32x32b) after st.16x256b.x1 of value T*16 + reg at lane base 0 and 16
__global__ void __launch_bounds__(128, 1) probe(uint32_t* out_ld, uint32_t* out_st) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc(&tmem_slot, 64);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = tmem_slot;
    if (warp == 1) {                                 // quadrant 1: lanes 32..63
        const uint32_t tl = tbase + (32u << 16);
        // ---- fill with 32x32b, read with 16x256b.x2 from the upper 16 lanes of the quadrant ----
        uint32_t f[16];
        for (int c = 0; c < 16; ++c) f[c] = lane * 256 + c;
        for (int c = 0; c < 16; c += 4) tmem_st4(tl + c, f[c], f[c + 1], f[c + 2], f[c + 3]);
        tmem_st_wait();
        __syncwarp();
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(tbase + (48u << 16)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int k = 0; k < 8; ++k) out_ld[lane * 8 + k] = r[k];
        __syncwarp();
        // ---- write with 16x256b.x1 (both 16-lane halves), read back with 32x32b ----
        for (int hh = 0; hh < 2; ++hh) {
            const uint32_t v0 = (hh * 32 + lane) * 16;
            asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1, %2, %3, %4};"
                         ::"r"(tbase + (uint32_t(32 + 16 * hh) << 16) + 32), "r"(v0), "r"(v0 + 1), "r"(v0 + 2), "r"(v0 + 3) : "memory");
        }
        tmem_st_wait();
        __syncwarp();
        uint32_t g[8];
        tmem_ld8(tl + 32, g);
        tmem_ld_wait8(g);
        for (int k = 0; k < 8; ++k) out_st[lane * 8 + k] = g[k];
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 64);
}

int main() {
    uint32_t *d_ld, *d_st, h_ld[256], h_st[256];
    CK(cudaMalloc(&d_ld, sizeof h_ld));
    CK(cudaMalloc(&d_st, sizeof h_st));
    probe<<<1, 128>>>(d_ld, d_st);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h_ld, d_ld, sizeof h_ld, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h_st, d_st, sizeof h_st, cudaMemcpyDeviceToHost));
    int bad = 0;
    // expected: thread T, reg k: rep j = k/4, row = T/4 + 8*((k%4)/2), col = 8*j + 2*(T%4) + (k%2); lanes 16.. of the quadrant
    printf("ld.16x256b.x2: thread reg -> (lane-in-quadrant, col)\n");
    for (int T = 0; T < 32; ++T) {
        printf("T%2d:", T);
        for (int k = 0; k < 8; ++k) {
            const int ln = h_ld[T * 8 + k] / 256, col = h_ld[T * 8 + k] % 256;
            const int eln = 16 + T / 4 + 8 * ((k % 4) / 2), ecol = 8 * (k / 4) + 2 * (T % 4) + (k % 2);
            if (ln != eln || col != ecol) ++bad;
            printf(" (%2d,%2d)", ln, col);
        }
        printf("\n");
    }
    printf("st.16x256b.x1: TMEM lane, col -> (thread, reg) [value = (hh*32+T)*16+reg]\n");
    for (int ln = 0; ln < 32; ++ln) {
        printf("lane %2d:", ln);
        for (int c = 0; c < 8; ++c) {
            const int v = h_st[ln * 8 + c], T = (v / 16) % 32, hh = v / 16 / 32, k = v % 16;
            const int l16 = ln % 16, eT = (l16 % 8) * 4 + c / 2, ek = 2 * (l16 / 8) + (c % 2), ehh = ln / 16;
            if (T != eT || k != ek || hh != ehh) ++bad;
            printf(" (h%d T%2d r%d)", hh, T, k);
        }
        printf("\n");
    }
    printf(bad ? "MISMATCH vs expected layout: %d\n" : "PASS (layout as expected) %d\n", bad);
    return bad;
}

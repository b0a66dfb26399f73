// sm100_prims.cuh — thin inline-PTX wrappers for the Blackwell (sm_100a) primitives the
// neural-receiver kernels are built from: mbarrier, 1-D bulk async copy (UBLKCP), tensor-memory
// allocation, tcgen05.mma (UMMA) with hand-built shared-memory / instruction descriptors,
// tcgen05.ld, and the proxy fences that order generic-proxy smem writes before async-proxy reads.
//
// Conventions used everywhere in this repo
//   * GEMM operands are fp16, K-major, in the canonical 128-byte-swizzle layout: a "slab" holds
//     64 K-elements (128 B) for `rows` rows; row r lives at r*128 B and its eight 16-byte chunks
//     are XOR-permuted by (r & 7).  K > 64 uses several slabs, `rows*128` bytes apart.
//   * accumulators are fp32 in TMEM, M = 128 rows <-> TMEM lanes 0..127, one column per n.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace nrx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrive_count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrive_count)
                 : "memory");
}
// make barrier inits visible to the async proxy (bulk copies / tcgen05.commit arrive on them)
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t tx_bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(tx_bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a lost arrive must never hang the GPU (a hang costs the whole box); after
// ~2^26 probes (seconds) the kernel traps, which surfaces as a CUDA error on the host.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}

// ------------------------------------------------------------------------------------------
// fences between proxies / tcgen05 ordering
// ------------------------------------------------------------------------------------------
// generic-proxy st.shared -> visible to async proxy (UMMA operand reads, bulk stores)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// 1-D bulk async copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
// size and both addresses must be multiples of 16 bytes.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                         uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// shared -> global bulk store (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_all() {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// Tensor memory
// ------------------------------------------------------------------------------------------
// executed by ONE full warp; ncols power of two in [32, 512]; base address lands in *smem_slot
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(smem_slot)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread `lane` of warp w gets TMEM lane 32*(w%4)+lane.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// TMEM address: lane in bits [31:16], column in bits [15:0]
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, uint32_t lane, uint32_t col) {
    return base + (lane << 16) + col;
}

// ------------------------------------------------------------------------------------------
// UMMA descriptors
// ------------------------------------------------------------------------------------------
constexpr uint32_t kSlabRowBytes = 128;   // 64 fp16 along K
constexpr uint32_t kSlabK = 64;

// byte offset of (row, k) inside one 128B-swizzled K-major slab
__host__ __device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t k_in_slab) {
    const uint32_t chunk = (k_in_slab >> 3) ^ (row & 7u);
    return row * kSlabRowBytes + (chunk << 4) + ((k_in_slab & 7u) << 1);
}

// Shared-memory matrix descriptor, K-major, SWIZZLE_128B, 8-row groups 1024 B apart.
// (field layout: cute/arch/mma_sm100_desc.hpp SmemDescriptor — start[0,14) lbo[16,30)
//  sbo[32,46) version[46,48)=1 base_offset[49,52) layout[61,64)=2)
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr_bytes & 0x3FFFFu) >> 4);   // start address
    d |= static_cast<uint64_t>(1) << 16;                              // LBO (unused with swizzle)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;                      // SBO: 8 rows * 128 B
    d |= static_cast<uint64_t>(1) << 46;                              // descriptor version (sm_100)
    d |= static_cast<uint64_t>(2) << 61;                              // SWIZZLE_128B
    return d;
}

// Instruction descriptor for kind::f16, fp16 A/B (K-major both), fp32 accumulate.
// (cute InstrDescriptor: c_format[4,6)=1 a_format[7,10)=0 b_format[10,13)=0 a_major[15]=0
//  b_major[16]=0 n_dim[17,23)=N>>3 m_dim[24,29)=M>>4)
__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t M, uint32_t N) {
    return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(static_cast<uint32_t>(accumulate))
        : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}

// Issue a full [128 x K] x [N x K]^T product as K/16 UMMAs.  a_base / b_base are the byte
// addresses of slab 0 of each operand; slabs are a_slab_bytes / b_slab_bytes apart.
__device__ __forceinline__ void umma_gemm_k(uint32_t tmem_d, uint32_t a_base, uint32_t a_slab_bytes,
                                            uint32_t b_base, uint32_t b_slab_bytes, uint32_t K,
                                            uint32_t idesc, bool accumulate_first) {
    for (uint32_t k = 0; k < K; k += 16) {
        const uint32_t slab = k / kSlabK;
        const uint32_t koff = (k % kSlabK) * 2;   // bytes inside the 128 B row
        const uint64_t ad = umma_smem_desc(a_base + slab * a_slab_bytes + koff);
        const uint64_t bd = umma_smem_desc(b_base + slab * b_slab_bytes + koff);
        umma_f16(tmem_d, ad, bd, idesc, accumulate_first || k > 0);
    }
}


// ------------------------------------------------------------------------------------------
// CTA pairs (cluster of 2, tcgen05 cta_group::2) — validated by tools/umma_pair_probe.cu:
// one MMA of M = 256 spans both CTAs; each CTA supplies its own 128-row A tile (same smem offset
// in both) and HALF of the B operand (rank r holds weight rows [r*N/2, (r+1)*N/2)); each CTA's
// TMEM receives D for its 128 rows and all N columns.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// cluster-wide address of a shared-memory location in CTA `rank`
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// bounded wait with cluster-scope acquire (pairs with remote arrives and multicast commits)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0, ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (!ok && ++spins > (1u << 26)) __trap();
    }
}
// executed by ONE full warp in EACH CTA of the pair (same warp id, same smem slot offset)
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// issued by ONE thread of the leader CTA (rank 0)
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(static_cast<uint32_t>(accumulate)), "r"(0u)
        : "memory");
}
// arrive on the barrier at this smem offset in every CTA of `cta_mask` when the pair's MMAs are done
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask)
                 : "memory");
}
// [256 x K] x [N x K]^T over the pair: a_base / b_base are the (identical) smem offsets in both CTAs
__device__ __forceinline__ void umma_gemm_k_pair(uint32_t tmem_d, uint32_t a_base, uint32_t a_slab_bytes, uint32_t b_base,
                                                 uint32_t b_slab_bytes, uint32_t K, uint32_t idesc) {
    for (uint32_t k = 0; k < K; k += 16) {
        const uint32_t slab = k / kSlabK, koff = (k % kSlabK) * 2;
        umma_f16_pair(tmem_d, umma_smem_desc(a_base + slab * a_slab_bytes + koff),
                      umma_smem_desc(b_base + slab * b_slab_bytes + koff), idesc, k > 0);
    }
}

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
// fp32 pair -> fp16x2 with ReLU in one instruction (lo in the low half)
__device__ __forceinline__ uint32_t pack_relu_half2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

}  // namespace nrx

// ------------------------------------------------------------------------------------------
// Additions for the TMEM-resident stack kernel (nrx_stack_tm.cuh): A operand of the MMA in tensor
// memory (written by tcgen05.st from registers), tensor-map TMA with the 128-byte swizzle.
// Validated by tools/umma_ts_probe.cu.
// ------------------------------------------------------------------------------------------
namespace nrx {

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 8 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
// wait for the thread's outstanding tcgen05.ld; the registers are operands so that no use of them
// can be scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait8(uint32_t (&r)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
                 :
                 : "memory");
}
// thread `lane` of warp w writes 4 consecutive 32-bit columns of TMEM lane 32*(w%4)+lane
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d)
                 : "memory");
}
// same, 16 consecutive columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]^T : A is fp16 [128 lanes][K/2 columns], two consecutive K elements
// per 32-bit column (even k in the low half); one instruction consumes K = 16 = 8 columns.
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(static_cast<uint32_t>(accumulate))
        : "memory");
}
__device__ __forceinline__ void umma_gemm_k_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_base, uint32_t b_slab_bytes,
                                               uint32_t K, uint32_t idesc) {
    for (uint32_t k = 0; k < K; k += 16) {
        const uint32_t slab = k / kSlabK, koff = (k % kSlabK) * 2;
        umma_f16_ts(tmem_d, tmem_a + k / 2, umma_smem_desc(b_base + slab * b_slab_bytes + koff), idesc, k > 0);
    }
}

// 16x256b fragments (tools/tmem_frag_probe.cu): thread T holds TMEM lanes base + T/4 (r0, r1 / r4, r5) and
// base + T/4 + 8 (r2, r3 / r6, r7), columns col + 2*(T%4) + {0, 1} (r0..r3) and col + 8 + 2*(T%4) + {0, 1} (r4..r7);
// `taddr` carries the first of the 16 lanes (a multiple of 16 inside the warp's quadrant).
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&a)[8], uint32_t (&b)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]),
                   "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7])
                 :
                 : "memory");
}
// lanes base + T/4: columns col + 2*(T%4) + {0,1} <- a, b;  lanes base + T/4 + 8: same columns <- c, d
__device__ __forceinline__ void tmem_st_16x256b_x1(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d)
                 : "memory");
}

// wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (or the hint
// expires) instead of probing in a tight loop that steals issue slots from the warps that work
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0, ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
            : "memory");
        if (!ok && ++spins > (1u << 22)) __trap();        // a probe returns after tens of cycles (measured: ~85 probes per
    }                                                     // 3 k-cycle wait), so the bound is ~0.1 s: a lost arrive surfaces as a
}                                                         // CUDA error, never as a hang

// 3-D tensor-map TMA (box lands in the map's swizzle pattern; out-of-range coordinates read zeros
// and are clipped on stores)
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* tmap, int c0, int c1, int c2, const void* smem_src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(smem_src))
                 : "memory");
}

}  // namespace nrx

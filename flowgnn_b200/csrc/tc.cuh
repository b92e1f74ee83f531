// tcgen05 / TMEM building blocks (sm_100a inline PTX) used by the fused layer kernels.
//
// Conventions used throughout:
//   * one CTA per SM, cta_group::1, UMMA M = 128: accumulator row i lives in TMEM lane i, column j in
//     TMEM column j (fp32);
//   * the A operand is read from TMEM (".ts" form): row i in lane i, two consecutive bf16 K-elements per
//     32-bit column, so one K = 16 MMA consumes 8 columns;
//   * the B operand (weights, [N][K] "out x in", i.e. K-major) is stationary in shared memory in the
//     no-swizzle canonical layout: 8-row x 16-byte core matrices, K-chunk-major
//         byte_offset(n, k) = (k / 8) * (N * 16) + n * 16 + (k % 8) * 2
//     -> descriptor LBO (K direction) = N * 16 bytes, SBO (N direction, 8-row groups) = 128 bytes.
//   * warp w may only touch TMEM lanes [32 * (w % 4), 32 * (w % 4) + 32).
#pragma once

#include <cuda_bf16.h>
#include <stdint.h>

#include "common.cuh"

namespace fg {
namespace tc {

#ifdef __CUDACC__

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish()
{
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// mbarrier arrive once all previously issued MMAs of this thread have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, both operands K-major (cute::UMMA::InstrDescriptor bit layout)
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor bit layout, version 1)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           (1ull << 46);
}

// D[tmem] (+)= A[tmem] * B[smem]^T, one K = 16 step; issued by ONE thread
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}

// 32 lanes x 32 bit: thread = TMEM lane (row), registers = consecutive columns
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t (&r)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

// x = hi + lo (+ O(2^-16 |x|)), both bf16 (round-to-nearest); returns hi in the low half, lo in the high half
__device__ __forceinline__ uint32_t split_bf16(float x)
{
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
    return (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
}
// two split words (x0, x1) -> packed hi pair (x0.hi | x1.hi << 16) and lo pair
__device__ __forceinline__ uint32_t pack_hi(uint32_t s0, uint32_t s1) { return __byte_perm(s0, s1, 0x5410); }
__device__ __forceinline__ uint32_t pack_lo(uint32_t s0, uint32_t s1) { return __byte_perm(s0, s1, 0x7632); }

#endif  // __CUDACC__

// byte offset of weight element (n, k) in the stationary B layout for an [N][K] matrix (host side repacks with this)
__host__ __device__ constexpr size_t b_offset_bytes(int n, int k, int N) { return (size_t)(k / 8) * ((size_t)N * 16) + (size_t)n * 16 + (size_t)(k % 8) * 2; }

}  // namespace tc
}  // namespace fg

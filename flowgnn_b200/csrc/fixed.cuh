// Helpers of the ap_fixed<16,I> kernels (gin_fixed.cu, dgn_fixed.cu): 16-bit two's-complement values with F = 16 - I
// fraction bits, Vitis' default modes AP_TRN (floor on assignment) and AP_WRAP (keep the low 16 bits).
#pragma once

#include "internal.cuh"

namespace fg {

#ifdef __CUDACC__

// floor(a * w / 2^F) + c for a held as raw << 16 and w as raw << (16 - F): the high word of the 64-bit product
__device__ __forceinline__ int mad_hi(int a, int b, int c)
{
    int d;
    asm("mad.hi.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// acc + floor(a * w / 1024) (mod 2^16) for RAW 16-bit operands: a * w + 2^30 is non-negative, its logical shift adds 2^20 (= 0 mod 2^16)
__device__ __forceinline__ int mac_floor(int a, int w, int acc)
{
    return (int)((unsigned)acc + ((unsigned)(a * w + 0x40000000) >> 10));
}
__device__ __forceinline__ long long mul_wide(int a, int b)
{
    long long d;
    asm("mul.wide.s32 %0, %1, %2;" : "=l"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ long long mad_wide(int a, int b, long long c)
{
    long long d;
    asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
    return d;
}
__device__ __forceinline__ int wrap16(int x) { return (int)(int16_t)x; }
__device__ __forceinline__ int relu16(int x) { x = wrap16(x); return x < 0 ? 0 : x; }      // ap_fixed_relu on the wrapped value (*/src/util.h:20-25)
__device__ __forceinline__ int abs16(int x) { x = wrap16(x); return x < 0 ? wrap16(-x) : x; }   // hls::abs returns the operand's type: -(-2^15) wraps

#endif

struct FixedEmbedOffsets { int off[ND_FEATURE]; };

// h0_v = sum over the nine categorical features of Table[off_f + x_vf], mod 2^16 (GIN/src/load_inputs.cc:174-220, DGN :114-168); dim 100
int fixed_embed_launch(const int* feat, const int16_t* table, const FixedEmbedOffsets& off, int16_t* h, long num_nodes, int sm_count, cudaStream_t s);

}  // namespace fg

// Generic "aggregate -> tensor-core node transform" GEMM, the structure proven by pna_tc.cu, for the models whose node
// transform is ONE dense layer per message-passing step (GCN: 100 -> 100, DGN: 200 -> 100).
//
//   * the model's aggregate kernel (CUDA cores) writes the A operand to HBM as bf16 hi + bf16 lo blocks in the tcgen05
//     no-swizzle K-major canonical layout, per tile of 128 nodes and K chunk of 64:
//         block(t, c) = [hi 16 KB | lo 16 KB],   byte(r, k) = (k / 8) * 2048 + r * 16 + (k % 8) * 2
//     (lane = row, eight consecutive k per thread = one 16-byte store; K is padded with ZEROS: garbage times a zero
//     weight could be NaN);
//   * gemm_kernel: persistent CTA per SM, 192 threads.  Warp 0: producer (one bulk-TMA copy of an A block and one of the
//     matching weight chunk per stage); warp 1: tcgen05.mma issuer, hi*hi + lo*hi + hi*lo, M = 128, N = NPAD, SS operands,
//     two accumulators in tensor memory; warps 2..5: epilogue, thread = row, 16 columns per tcgen05.ld, the model's
//     functor turns accumulator columns into stores.
// The 3-product split keeps the fp32 contract (1e-4) -- error budget in gin_tc2.cu / DESIGN.md.
#pragma once

#include "internal.cuh"
#include "tc.cuh"

namespace fg {
namespace tcg {

constexpr int TM = 128;                      // nodes per tile (UMMA M)
constexpr int KC = 64;                       // K per chunk
constexpr int A_HALF = TM * KC * 2;          // 16,384
constexpr int A_BLOCK = 2 * A_HALF;          // 32,768
constexpr int LBO_A = TM * 16;
constexpr int NT = 192;

template <int NPAD>
struct Cfg {
    static_assert(NPAD % 16 == 0 && NPAD >= 16 && NPAD <= 256, "UMMA N for M = 128");
    static constexpr int B_HALF = NPAD * KC * 2;
    static constexpr int B_BLOCK = 2 * B_HALF;
    static constexpr int LBO_B = NPAD * 16;
    static constexpr int STAGE_BYTES = A_BLOCK + B_BLOCK;
    static constexpr int STAGES = (3 * STAGE_BYTES + 256 <= 232448) ? 3 : 2;
    static constexpr int ACC_COLS = NPAD <= 128 ? 128 : 256;
    static constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;
    static constexpr int BAR = STAGES * STAGE_BYTES;              // full[STAGES], empty[STAGES], acc_full[2], acc_empty[2]
    static constexpr int TMEM_PTR = BAR + (2 * STAGES + 4) * 8;
    static constexpr int BYTES = TMEM_PTR + 16;
    static_assert(BYTES <= 232448, "shared memory budget");
};

#ifdef __CUDACC__

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_park(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
            : "memory");
    } while (!done);
}
// D[tmem] (+)= A[smem] * B[smem]^T, one K = 16 step, both operands K-major; issued by ONE thread
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// (x0, x1) -> packed bf16 pairs: hi = rn(x), lo = rn(x - hi); element 0 in the low half
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo)
{
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    const float r0 = x0 - __uint_as_float(hi << 16);
    const float r1 = x1 - __uint_as_float(hi & 0xFFFF0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}
// eight consecutive k (k0 % 8 == 0) of row v -> the hi and lo blocks of its tile
template <int NCHUNK>
__device__ __forceinline__ void put8(unsigned char* apack, long v, int k0, const float (&x)[8])
{
    const long t = v / TM;
    const int r = (int)(v - t * TM), c = k0 / KC, kk = k0 % KC;
    unsigned char* blk = apack + ((size_t)t * NCHUNK + c) * A_BLOCK + (kk / 8) * LBO_A + r * 16;
    uint4 hi, lo;
    split2(x[0], x[1], hi.x, lo.x);
    split2(x[2], x[3], hi.y, lo.y);
    split2(x[4], x[5], hi.z, lo.z);
    split2(x[6], x[7], hi.w, lo.w);
    *reinterpret_cast<uint4*>(blk) = hi;
    *reinterpret_cast<uint4*>(blk + A_HALF) = lo;
}
template <int NCHUNK>
__device__ __forceinline__ void put8_zero(unsigned char* apack, long v, int k0)
{
    const long t = v / TM;
    const int r = (int)(v - t * TM), c = k0 / KC, kk = k0 % KC;
    unsigned char* blk = apack + ((size_t)t * NCHUNK + c) * A_BLOCK + (kk / 8) * LBO_A + r * 16;
    *reinterpret_cast<uint4*>(blk) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(blk + A_HALF) = make_uint4(0, 0, 0, 0);
}

struct GemmArgs {
    const unsigned char* apack;      // [tiles][NCHUNK][32768]
    const unsigned char* wpack;      // [NCHUNK][Cfg<NPAD>::B_BLOCK] this layer
    int num_nodes; int num_tiles;
};

// Epi: struct with   __device__ State begin(int v, bool live) const;
//                    __device__ void store(const State&, int v, int d0, const uint32_t (&acc)[16]) const;   (columns d0 .. d0+15)
template <int NCHUNK, int NPAD, class Epi>
__global__ void __launch_bounds__(NT, 1) gemm_kernel(GemmArgs g, Epi epi)
{
    using C = Cfg<NPAD>;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + C::BAR);
    uint64_t* bar_full = bar;
    uint64_t* bar_empty = bar + C::STAGES;
    uint64_t* bar_acc_full = bar + 2 * C::STAGES;
    uint64_t* bar_acc_empty = bar + 2 * C::STAGES + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + C::TMEM_PTR);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0)
    {
        for (int i = 0; i < C::STAGES; i++) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&bar_acc_full[i], 1); mbar_init(&bar_acc_empty[i], 128); }
        fence_mbar_init();
    }
    if (warp == 1)
    {
        tc::tmem_alloc(tmem_ptr, C::TMEM_COLS);
        tc::tmem_relinquish();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = *tmem_ptr;

    if (warp == 0)
    {
        if (lane == 0)
        {
            uint32_t n = 0;
            for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x)
                for (int c = 0; c < NCHUNK; c++, n++)
                {
                    const uint32_t s = n % C::STAGES;
                    mbar_wait_park(&bar_empty[s], ((n / C::STAGES) & 1) ^ 1);
                    unsigned char* st = smem + s * C::STAGE_BYTES;
                    mbar_arrive_expect_tx(&bar_full[s], C::STAGE_BYTES);
                    tma_load_1d(st, g.apack + ((size_t)tile * NCHUNK + c) * A_BLOCK, A_BLOCK, &bar_full[s]);
                    tma_load_1d(st + A_BLOCK, g.wpack + (size_t)c * C::B_BLOCK, C::B_BLOCK, &bar_full[s]);
                }
        }
    }
    else if (warp == 1)
    {
        if (lane == 0)
        {
            const uint32_t idesc = tc::idesc_bf16(TM, NPAD);
            uint32_t n = 0, it = 0;
            for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, it++)
            {
                const uint32_t a = it & 1;
                mbar_wait_park(&bar_acc_empty[a], ((it >> 1) & 1) ^ 1);
                tc::fence_after_sync();
                const uint32_t d_tmem = tbase + a * C::ACC_COLS;
                for (int c = 0; c < NCHUNK; c++, n++)
                {
                    const uint32_t s = n % C::STAGES;
                    mbar_wait_park(&bar_full[s], (n / C::STAGES) & 1);
                    tc::fence_after_sync();
                    const uint32_t a_addr = smem_u32(smem + s * C::STAGE_BYTES), b_addr = a_addr + A_BLOCK;
#pragma unroll
                    for (int j = 0; j < KC / 16; j++)
                    {
                        const uint64_t a_hi = tc::smem_desc(a_addr + 2 * j * LBO_A, LBO_A, 128);
                        const uint64_t a_lo = tc::smem_desc(a_addr + A_HALF + 2 * j * LBO_A, LBO_A, 128);
                        const uint64_t b_hi = tc::smem_desc(b_addr + 2 * j * C::LBO_B, C::LBO_B, 128);
                        const uint64_t b_lo = tc::smem_desc(b_addr + C::B_HALF + 2 * j * C::LBO_B, C::LBO_B, 128);
                        mma_ss(d_tmem, a_hi, b_hi, idesc, !(c == 0 && j == 0));
                        mma_ss(d_tmem, a_lo, b_hi, idesc, true);
                        mma_ss(d_tmem, a_hi, b_lo, idesc, true);
                    }
                    tc::commit(&bar_empty[s]);                    // the stage is free once these MMAs have read it
                }
                tc::commit(&bar_acc_full[a]);
            }
        }
    }
    else
    {
        const int lg = warp & 3;                                  // TMEM lane group this warp may read
        const int row = lg * 32 + lane;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, it++)
        {
            const uint32_t a = it & 1;
            const int v = tile * TM + row;
            const bool live = v < g.num_nodes;
            const auto st = epi.begin(v, live);
            mbar_wait_park(&bar_acc_full[a], (it >> 1) & 1);
            tc::fence_after_sync();
            const uint32_t taddr = tbase + a * C::ACC_COLS + ((uint32_t)(lg * 32) << 16);
#pragma unroll 1
            for (int d0 = 0; d0 < NPAD; d0 += 16)
            {
                uint32_t acc[16];
                tc::ld16(taddr + d0, acc);
                tc::wait_ld();
                if (live) epi.store(st, v, d0, acc);
            }
            tc::fence_before_sync();
            mbar_arrive(&bar_acc_empty[a]);
        }
    }

    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tbase, C::TMEM_COLS);
}

#endif  // __CUDACC__

// W [n_real][k_real] (reference "[out][in]", element (n, k) at w[n * ld + k_map]) -> NCHUNK blocks [NPAD x 64] hi | lo in the
// canonical K-major layout; `kmap(k)` gives the source column of padded k, or -1 for a zero column
template <int NPAD, class KMap>
inline void pack_weights(const float* w, int n_real, int ld, int nchunk, KMap kmap, unsigned char* dst, uint16_t (*bf16_rn)(float),
                         float (*bf16_to_float)(uint16_t))
{
    using C = Cfg<NPAD>;
    for (size_t i = 0; i < (size_t)nchunk * C::B_BLOCK; i++) dst[i] = 0;
    for (int k = 0; k < nchunk * KC; k++)
    {
        const int ks = kmap(k);
        if (ks < 0) continue;
        for (int n = 0; n < n_real; n++)
        {
            const float x = w[(size_t)n * ld + ks];
            const uint16_t hi = bf16_rn(x), lo = bf16_rn(x - bf16_to_float(hi));
            const int c = k / KC, kk = k % KC;
            unsigned char* o = dst + (size_t)c * C::B_BLOCK + (size_t)(kk / 8) * C::LBO_B + (size_t)n * 16 + (size_t)(kk % 8) * 2;
            o[0] = (unsigned char)(hi & 0xFF); o[1] = (unsigned char)(hi >> 8);
            o[C::B_HALF] = (unsigned char)(lo & 0xFF); o[C::B_HALF + 1] = (unsigned char)(lo >> 8);
        }
    }
}

}  // namespace tcg
}  // namespace fg

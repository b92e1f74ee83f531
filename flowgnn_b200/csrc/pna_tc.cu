// PNA layer with the node transform on the B200 tensor cores (option "pna_tc").
//
// Reference work per layer (PNA/src/message_passing.cc:88-147, node_embedding.cc:106-215), as in pna.cu:
//   per (v, d): S = sum h_u, Q = sum h_u^2, min, max over in-edges  ->  mean, min, max, std   (A row of 320 values)
//   acc = b + sum_in [T0 + T1 t + T2 s]  ==  b + G0 + t G1 + s G2  with  G = A [320] x Wcat [320 x 240];  h <- h + relu(acc)
// The FFMA kernel (pna.cu) spends 154 kFLOP per node on the FP32 pipe (56 ms per layer on molpcba, 42 % of the FFMA peak).
// Here a layer is three launches:
//   1. pna_aggregate_kernel   message passing on the CUDA cores (lane = row, eight columns per thread); every A value is split into bf16 hi + bf16 lo
//                             (x = hi + lo + O(2^-17 |x|)) and written to HBM in the tcgen05 no-swizzle K-major canonical
//                             layout, blocked per tile of 128 nodes and K chunk of 64:
//                                 block(t, c) = [hi 16 KB | lo 16 KB],  byte(r, k) = (k / 8) * 2048 + r * 16 + (k % 8) * 2
//   2. pna_gemm_kernel        persistent CTA per SM, warp specialised: a producer thread streams A blocks and the matching
//                             weight chunks ([240 x 64] hi | lo, same layout, LBO 3840) with ONE bulk-TMA copy each through
//                             a two-stage ring; an issuer thread runs hi*hi + lo*hi + hi*lo (3 x 4 k-steps of SS
//                             tcgen05.mma per chunk, M = 128, N = 240) into one of two 256-column accumulators in tensor
//                             memory; four epilogue warps read the other accumulator, combine the three scaler groups
//                             with the node's degree scalers, add the bias, relu and the residual, and store h'.
//   3. pna_exact_rows_kernel  rows the tensor path must not touch: out-degree 0 (the reference's own expression decides
//                             between inf and NaN) and rows whose aggregates are not finite (a bf16 split
//                             of inf is inf + NaN) are evaluated in fp32, a warp per row, exactly as pna.cu does.
// The 3-product split keeps the fp32 contract of BASELINE.json (1e-4); see gin_tc2.cu for the error budget.
#include "internal.cuh"
#include "layers.cuh"
#include "tc.cuh"

#include <algorithm>
#include <vector>

namespace fg {

namespace {

constexpr int D = 80;
constexpr int Q = D / 4;
constexpr int KA = 4 * D;                    // 320
constexpr int NC = 3 * D;                    // 240
constexpr int TM = 128;                      // nodes per tile (UMMA M)
constexpr int KC = 64;                       // K per chunk
constexpr int NCHUNK = KA / KC;              // 5
constexpr int A_HALF = TM * KC * 2;          // 16,384: one hi or lo block of A
constexpr int A_BLOCK = 2 * A_HALF;          // 32,768
constexpr int B_HALF = NC * KC * 2;          // 30,720
constexpr int B_BLOCK = 2 * B_HALF;          // 61,440
constexpr int LBO_A = TM * 16, LBO_B = NC * 16;
constexpr int STAGES = 2;
constexpr int STAGE_BYTES = A_BLOCK + B_BLOCK;                    // 94,208
constexpr int NT = 192;                      // warp 0 producer, warp 1 MMA issuer, warps 2..5 epilogue
constexpr uint32_t TMEM_COLS = 512;          // two accumulators of 256 columns (240 used)

struct Smem {
    static constexpr int STAGE = 0;
    static constexpr int BAR = STAGES * STAGE_BYTES;              // full[2], empty[2], acc_full[2], acc_empty[2]
    static constexpr int TMEM_PTR = BAR + 8 * 8;
    static constexpr int BYTES = TMEM_PTR + 16;
};
static_assert(Smem::BYTES <= 232448, "shared memory budget");
enum { BAR_FULL = 0, BAR_EMPTY = 2, BAR_ACC_FULL = 4, BAR_ACC_EMPTY = 6 };

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

// ap_fixed_max / ap_fixed_min of ap_fixed<16,6> (PNA/src/util.h:34-46)
constexpr float FM_MAX = 32.0f - 0.0009765625f, FM_MIN = -32.0f;

// The four aggregates of columns 4q..4q+3 of node v, in-edges in CSR order (identical to the loop of pna.cu)
__device__ __forceinline__ void aggregate4(const float* __restrict__ h_in, const int* __restrict__ in_ptr, const int* __restrict__ src, int v,
                                           int q, float (&mean)[4], float (&mn)[4], float (&mx)[4], float (&sd)[4])
{
    const int eb = __ldg(in_ptr + v), ee = __ldg(in_ptr + v + 1);
    float s[4] = {0.f, 0.f, 0.f, 0.f}, sq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; j++) { mn[j] = FM_MAX; mx[j] = FM_MIN; }
    for (int e = eb; e < ee; e++)
    {
        const float4 hu = ldg_f4(h_in + (size_t)__ldg(src + e) * D + 4 * q);
        const float x[4] = {hu.x, hu.y, hu.z, hu.w};
#pragma unroll
        for (int j = 0; j < 4; j++)
        {
            s[j] += x[j];
            sq[j] += x[j] * x[j];
            if (x[j] < mn[j]) mn[j] = x[j];
            if (x[j] > mx[j]) mx[j] = x[j];
        }
    }
    int in_deg = ee - eb;
    if (in_deg == 0) in_deg = 1;
    const float fn = (float)in_deg;
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
        mean[j] = s[j] / fn;
        sd[j] = sqrtf(relu_f(sq[j] / fn - mean[j] * mean[j]));
    }
}

__device__ __forceinline__ bool finite4(const float (&x)[4])
{
    return isfinite(x[0]) && isfinite(x[1]) && isfinite(x[2]) && isfinite(x[3]);
}

// ---- 1. message passing -> bf16 hi/lo A blocks ---------------------------------------------------------------------------
// A warp owns 32 consecutive rows (lane = row, the blocks are 128 rows tall) and eight columns: every gathered piece of a
// source row is one full 32-byte sector, and every store instruction writes 32 x 16 bytes = 512 contiguous bytes of a
// block (eight consecutive k of one row are 16 contiguous bytes in the canonical layout).  The first version mapped
// threads to (row, four columns) like pna.cu: its 8-byte stores scattered over eleven sectors per instruction.
constexpr int AG_WARPS = 8;
constexpr int Q2 = D / 8;                    // column groups of eight
__global__ void __launch_bounds__(AG_WARPS * 32) pna_aggregate_kernel(const float* __restrict__ h_in, const int* __restrict__ in_ptr,
                                                                      const int* __restrict__ src, unsigned char* __restrict__ apack,
                                                                      unsigned char* __restrict__ nonfinite, long num_nodes)
{
    const int lane = threadIdx.x & 31;
    const long warp = blockIdx.x * (long)AG_WARPS + (threadIdx.x >> 5), nwarps = (long)gridDim.x * AG_WARPS;
    const long items = ((num_nodes + 31) / 32) * Q2;
    for (long item = warp; item < items; item += nwarps)
    {
        const long rb = item / Q2;
        const int q2 = (int)(item - rb * Q2);
        const long v = rb * 32 + lane;
        if (v >= num_nodes) continue;
        const int eb = __ldg(in_ptr + v), ee = __ldg(in_ptr + v + 1);
        float s[8], sq[8], mn[8], mx[8];
#pragma unroll
        for (int j = 0; j < 8; j++) { s[j] = 0.f; sq[j] = 0.f; mn[j] = FM_MAX; mx[j] = FM_MIN; }
        for (int e = eb; e < ee; e++)
        {
            const float* hu = h_in + (size_t)__ldg(src + e) * D + 8 * q2;
            const float4 x0 = ldg_f4(hu), x1 = ldg_f4(hu + 4);
            const float x[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
            for (int j = 0; j < 8; j++)
            {
                s[j] += x[j];
                sq[j] += x[j] * x[j];
                if (x[j] < mn[j]) mn[j] = x[j];
                if (x[j] > mx[j]) mx[j] = x[j];
            }
        }
        int in_deg = ee - eb;
        if (in_deg == 0) in_deg = 1;
        const float fn = (float)in_deg;
        float mean[8], sd[8];
        bool fin = true;
#pragma unroll
        for (int j = 0; j < 8; j++)
        {
            mean[j] = s[j] / fn;
            sd[j] = sqrtf(relu_f(sq[j] / fn - mean[j] * mean[j]));
            fin = fin && isfinite(mean[j]) && isfinite(mn[j]) && isfinite(mx[j]) && isfinite(sd[j]);
        }
        if (!fin) nonfinite[v] = 1;
        const long t = v / TM;
        const int r = (int)(v - t * TM);
        auto put = [&](int g, const float (&x)[8]) {              // aggregator_t order (PNA/src/dcl.h:29-35): mean, min, max, std
            const int k = g * D + 8 * q2, c = k / KC, kk = k % KC;
            unsigned char* blk = apack + ((size_t)t * NCHUNK + c) * A_BLOCK + (kk / 8) * LBO_A + r * 16;
            uint4 hi, lo;
            split2(x[0], x[1], hi.x, lo.x);
            split2(x[2], x[3], hi.y, lo.y);
            split2(x[4], x[5], hi.z, lo.z);
            split2(x[6], x[7], hi.w, lo.w);
            *reinterpret_cast<uint4*>(blk) = hi;
            *reinterpret_cast<uint4*>(blk + A_HALF) = lo;
        };
        put(0, mean); put(1, mn); put(2, mx); put(3, sd);
    }
}

// ---- 2. G = A x Wcat on tcgen05, fused combine / relu / residual epilogue ----------------------------------------------------
struct PnaGemmParams {
    const unsigned char* apack;      // [tiles][5][32768]
    const unsigned char* wpack;      // [5][61440] this layer
    const float* h_in; float* h_out;
    const float* b;                  // [80]
    const int* out_deg;
    const unsigned char* nonfinite;
    float avg_deg;
    int num_nodes; int num_tiles;
};

__global__ void __launch_bounds__(NT, 1) pna_gemm_kernel(PnaGemmParams p)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + Smem::BAR);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + Smem::TMEM_PTR);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0)
    {
        for (int i = 0; i < 2; i++)
        {
            mbar_init(&bar[BAR_FULL + i], 1);
            mbar_init(&bar[BAR_EMPTY + i], 1);
            mbar_init(&bar[BAR_ACC_FULL + i], 1);
            mbar_init(&bar[BAR_ACC_EMPTY + i], 128);
        }
        fence_mbar_init();
    }
    if (warp == 1)
    {
        tc::tmem_alloc(tmem_ptr, TMEM_COLS);
        tc::tmem_relinquish();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = *tmem_ptr;

    if (warp == 0)
    {
        // ---- producer: one A block + one weight chunk per stage ----
        if (lane == 0)
        {
            uint32_t g = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x)
                for (int c = 0; c < NCHUNK; c++, g++)
                {
                    const uint32_t s = g % STAGES;
                    mbar_wait_park(&bar[BAR_EMPTY + s], ((g / STAGES) & 1) ^ 1);
                    unsigned char* st = smem + Smem::STAGE + s * STAGE_BYTES;
                    mbar_arrive_expect_tx(&bar[BAR_FULL + s], STAGE_BYTES);
                    tma_load_1d(st, p.apack + ((size_t)tile * NCHUNK + c) * A_BLOCK, A_BLOCK, &bar[BAR_FULL + s]);
                    tma_load_1d(st + A_BLOCK, p.wpack + (size_t)c * B_BLOCK, B_BLOCK, &bar[BAR_FULL + s]);
                }
        }
    }
    else if (warp == 1)
    {
        // ---- MMA issuer ----
        if (lane == 0)
        {
            const uint32_t idesc = tc::idesc_bf16(TM, NC);
            uint32_t g = 0, it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, it++)
            {
                const uint32_t a = it & 1;
                mbar_wait_park(&bar[BAR_ACC_EMPTY + a], ((it >> 1) & 1) ^ 1);
                tc::fence_after_sync();
                const uint32_t d_tmem = tbase + a * 256;
                for (int c = 0; c < NCHUNK; c++, g++)
                {
                    const uint32_t s = g % STAGES;
                    mbar_wait_park(&bar[BAR_FULL + s], (g / STAGES) & 1);
                    tc::fence_after_sync();
                    const uint32_t a_addr = smem_u32(smem + Smem::STAGE + s * STAGE_BYTES), b_addr = a_addr + A_BLOCK;
#pragma unroll
                    for (int j = 0; j < KC / 16; j++)
                    {
                        const uint64_t a_hi = tc::smem_desc(a_addr + 2 * j * LBO_A, LBO_A, 128);
                        const uint64_t a_lo = tc::smem_desc(a_addr + A_HALF + 2 * j * LBO_A, LBO_A, 128);
                        const uint64_t b_hi = tc::smem_desc(b_addr + 2 * j * LBO_B, LBO_B, 128);
                        const uint64_t b_lo = tc::smem_desc(b_addr + B_HALF + 2 * j * LBO_B, LBO_B, 128);
                        mma_ss(d_tmem, a_hi, b_hi, idesc, !(c == 0 && j == 0));
                        mma_ss(d_tmem, a_lo, b_hi, idesc, true);
                        mma_ss(d_tmem, a_hi, b_lo, idesc, true);
                    }
                    tc::commit(&bar[BAR_EMPTY + s]);              // the stage is free once these MMAs have read it
                }
                tc::commit(&bar[BAR_ACC_FULL + a]);
            }
        }
    }
    else
    {
        // ---- epilogue: thread = row (TMEM lane), 16 output columns per step ----
        const int lg = warp & 3;                                  // TMEM lane group this warp may read
        const int row = lg * 32 + lane;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, it++)
        {
            const uint32_t a = it & 1;
            const int v = tile * TM + row;
            const bool live = v < p.num_nodes;
            int od = 0;
            bool skip = true;
            if (live)
            {
                od = __ldg(p.out_deg + v);
                skip = od == 0 || __ldg(p.nonfinite + v) != 0;         // written by pna_exact_rows_kernel
            }
            const float log_degree = logf((float)(od + 1));
            const float t = log_degree / p.avg_deg;
            float scale = p.avg_deg / log_degree;
            if (scale == 0) scale = 1;
            mbar_wait_park(&bar[BAR_ACC_FULL + a], (it >> 1) & 1);
            tc::fence_after_sync();
            const uint32_t taddr = tbase + a * 256 + ((uint32_t)(lg * 32) << 16);
#pragma unroll 1
            for (int d0 = 0; d0 < D; d0 += 16)
            {
                uint32_t g0[16], g1[16], g2[16];
                tc::ld16(taddr + d0, g0);
                tc::ld16(taddr + D + d0, g1);
                tc::ld16(taddr + 2 * D + d0, g2);
                tc::wait_ld();
                if (!skip)
                {
                    const float* hin = p.h_in + (size_t)v * D + d0;
                    float* hout = p.h_out + (size_t)v * D + d0;
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                    {
                        const float4 hv = ldg_f4(hin + j);
                        const float4 bb = ldg_f4(p.b + d0 + j);
                        float4 o;
                        o.x = hv.x + relu_f(bb.x + __uint_as_float(g0[j]) + (__uint_as_float(g1[j]) * t + __uint_as_float(g2[j]) * scale));
                        o.y = hv.y + relu_f(bb.y + __uint_as_float(g0[j + 1]) + (__uint_as_float(g1[j + 1]) * t + __uint_as_float(g2[j + 1]) * scale));
                        o.z = hv.z + relu_f(bb.z + __uint_as_float(g0[j + 2]) + (__uint_as_float(g1[j + 2]) * t + __uint_as_float(g2[j + 2]) * scale));
                        o.w = hv.w + relu_f(bb.w + __uint_as_float(g0[j + 3]) + (__uint_as_float(g1[j + 3]) * t + __uint_as_float(g2[j + 3]) * scale));
                        stg_f4_stream(hout + j, o);
                    }
                }
            }
            tc::fence_before_sync();
            mbar_arrive(&bar[BAR_ACC_EMPTY + a]);
        }
    }

    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tbase, TMEM_COLS);
}

// ---- 3. rows evaluated in fp32 (out-degree 0, non-finite aggregates) -----------------------------------------------------------
constexpr int EX_WARPS = 8;
__global__ void __launch_bounds__(EX_WARPS * 32) pna_exact_rows_kernel(const float* __restrict__ h_in, float* __restrict__ h_out,
                                                                       const int* __restrict__ in_ptr, const int* __restrict__ src,
                                                                       const int* __restrict__ out_deg, const unsigned char* __restrict__ nonfinite,
                                                                       const float* __restrict__ wcat, const float* __restrict__ w_ref,
                                                                       const float* __restrict__ b, float avg_deg, int num_nodes)
{
    __shared__ __align__(16) float agg[EX_WARPS][KA];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int warp = blockIdx.x * EX_WARPS + wid, nwarps = gridDim.x * EX_WARPS;
    float* a = agg[wid];
    for (int v0 = warp * 32; v0 < num_nodes; v0 += nwarps * 32)
    {
        const int vl = v0 + lane;
        const int od_l = vl < num_nodes ? __ldg(out_deg + vl) : 1;
        const bool want = vl < num_nodes && (od_l == 0 || __ldg(nonfinite + vl) != 0);
        unsigned todo = __ballot_sync(0xFFFFFFFFu, want);
        while (todo)
        {
            const int i = __ffs(todo) - 1;
            todo &= todo - 1;
            const int v = v0 + i;
            const int od = __shfl_sync(0xFFFFFFFFu, od_l, i);
            __syncwarp();
            if (lane < Q)
            {
                float mean[4], mn[4], mx[4], sd[4];
                aggregate4(h_in, in_ptr, src, v, lane, mean, mn, mx, sd);
                st_f4(a + 4 * lane, make_float4(mean[0], mean[1], mean[2], mean[3]));
                st_f4(a + D + 4 * lane, make_float4(mn[0], mn[1], mn[2], mn[3]));
                st_f4(a + 2 * D + 4 * lane, make_float4(mx[0], mx[1], mx[2], mx[3]));
                st_f4(a + 3 * D + 4 * lane, make_float4(sd[0], sd[1], sd[2], sd[3]));
            }
            __syncwarp();
            const float log_degree = logf((float)(od + 1));
            const float t = log_degree / avg_deg;
            float scale = avg_deg / log_degree;
            if (scale == 0) scale = 1;
            for (int o = lane; o < D; o += 32)
            {
                float res;
                if (od == 0)
                {
                    // the reference's expression, term by term (PNA/src/node_embedding.cc:158-186), as in pna.cu
                    float acc = 0.f;
                    for (int k = 0; k < D; k++)
                    {
                        const float mean = a[k], mnv = a[D + k], mxv = a[2 * D + k], sdv = a[3 * D + k];
                        const float* w = w_ref + (size_t)o * 12 * D + k;       // [scaler][aggr][in], aggr: mean, min, max, std
#define WREF(sc, ag) __ldg(w + ((sc) * 4 + (ag)) * D)
                        const float t0 = __fadd_rn(__fadd_rn(__fmul_rn(mean, WREF(0, 0)), __fmul_rn(sdv, WREF(0, 3))),
                                                   __fadd_rn(__fmul_rn(mnv, WREF(0, 1)), __fmul_rn(mxv, WREF(0, 2))));
                        const float t1 = __fadd_rn(__fadd_rn(__fmul_rn(mean, WREF(1, 0)), __fmul_rn(sdv, WREF(1, 3))),
                                                   __fadd_rn(__fmul_rn(mnv, WREF(1, 1)), __fmul_rn(mxv, WREF(1, 2))));
                        const float t2 = __fadd_rn(__fadd_rn(__fmul_rn(mean, WREF(2, 0)), __fmul_rn(sdv, WREF(2, 3))),
                                                   __fadd_rn(__fmul_rn(mnv, WREF(2, 1)), __fmul_rn(mxv, WREF(2, 2))));
#undef WREF
                        const float addend = __fadd_rn(t0, __fadd_rn(__fmul_rn(t1, t), __fmul_rn(t2, scale)));
                        acc = __fadd_rn(addend, (k == 0) ? __ldg(b + o) : acc);
                    }
                    res = acc;
                }
                else
                {
                    // the FFMA kernel's evaluation (pna.cu): G = A x Wcat with k ascending, then b + G0 + (G1 t + G2 s)
                    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
                    for (int k = 0; k < KA; k++)
                    {
                        const float av = a[k];
                        g0 = fmaf(av, __ldg(wcat + (size_t)k * NC + o), g0);
                        g1 = fmaf(av, __ldg(wcat + (size_t)k * NC + D + o), g1);
                        g2 = fmaf(av, __ldg(wcat + (size_t)k * NC + 2 * D + o), g2);
                    }
                    res = __ldg(b + o) + g0 + (g1 * t + g2 * scale);
                }
                h_out[(size_t)v * D + o] = __ldg(h_in + (size_t)v * D + o) + relu_f(res);
            }
        }
    }
}

}  // namespace

// bytes of one layer's weight image for pna_gemm_kernel
size_t pna_tc_pack_bytes() { return (size_t)NCHUNK * B_BLOCK; }

// wcat [320][240] (k = aggr*80 + in, n = scaler*80 + out) -> per K chunk of 64 the [240 x 64] block as bf16 hi | lo in
// the canonical K-major layout byte(n, k) = (k / 8) * 3840 + n * 16 + (k % 8) * 2
void pna_tc_pack_layer(const float* wcat, unsigned char* dst, uint16_t (*bf16_rn)(float), float (*bf16_to_float)(uint16_t))
{
    for (int k = 0; k < KA; k++)
        for (int n = 0; n < NC; n++)
        {
            const float x = wcat[(size_t)k * NC + n];
            const uint16_t hi = bf16_rn(x), lo = bf16_rn(x - bf16_to_float(hi));
            const int c = k / KC, kk = k % KC;
            unsigned char* o = dst + (size_t)c * B_BLOCK + (size_t)(kk / 8) * LBO_B + (size_t)n * 16 + (size_t)(kk % 8) * 2;
            o[0] = (unsigned char)(hi & 0xFF); o[1] = (unsigned char)(hi >> 8);
            o[B_HALF] = (unsigned char)(lo & 0xFF); o[B_HALF + 1] = (unsigned char)(lo >> 8);
        }
}

int pna_layer_tc_launch(DeviceBatch& b, const PnaWeights& w, int layer, const float* h_in, float* h_out, int sm_count, cudaStream_t s)
{
    FG_TRY(opt_in_smem(reinterpret_cast<const void*>(&pna_gemm_kernel), Smem::BYTES));
    const long N = b.total_nodes;
    const int num_tiles = (int)ceil_div<long>(N, TM);
    FG_TRY(b.apack.reserve((size_t)num_tiles * NCHUNK * A_BLOCK));
    FG_TRY(b.nonfinite.reserve((size_t)N + 16));
    FG_TRY(zero_bytes_launch(b.nonfinite.ptr, (size_t)N, s));
    {
        const int blocks = (int)std::min<long>(ceil_div<long>(ceil_div<long>(N, 32) * Q2, AG_WARPS), (long)sm_count * 8);
        pna_aggregate_kernel<<<blocks, AG_WARPS * 32, 0, s>>>(h_in, b.in_ptr.as<int>(), b.src.as<int>(), b.apack.as<unsigned char>(),
                                                    b.nonfinite.as<unsigned char>(), N);
        FG_CUDA(cudaGetLastError());
    }
    {
        PnaGemmParams p{};
        p.apack = b.apack.as<unsigned char>();
        p.wpack = w.wpack_tc.as<unsigned char>() + (size_t)layer * pna_tc_pack_bytes();
        p.h_in = h_in; p.h_out = h_out;
        p.b = w.b.as<float>() + (size_t)layer * D;
        p.out_deg = b.out_deg.as<int>();
        p.nonfinite = b.nonfinite.as<unsigned char>();
        p.avg_deg = w.avg_deg;
        p.num_nodes = (int)N; p.num_tiles = num_tiles;
        pna_gemm_kernel<<<std::min(num_tiles, sm_count), NT, Smem::BYTES, s>>>(p);
        FG_CUDA(cudaGetLastError());
    }
    return pna_exact_rows_launch(b, w, layer, h_in, h_out, sm_count, s);
}

// rows the tensor paths leave out (out-degree 0, non-finite aggregates), in fp32
int pna_exact_rows_launch(DeviceBatch& b, const PnaWeights& w, int layer, const float* h_in, float* h_out, int sm_count, cudaStream_t s)
{
    const long N = b.total_nodes;
    const int blocks = (int)std::min<long>(ceil_div<long>(N, 32 * EX_WARPS), (long)sm_count * 8);
    pna_exact_rows_kernel<<<blocks, EX_WARPS * 32, 0, s>>>(h_in, h_out, b.in_ptr.as<int>(), b.src.as<int>(), b.out_deg.as<int>(),
                                                           b.nonfinite.as<unsigned char>(), w.wcat.as<float>() + (size_t)layer * KA * NC,
                                                           w.w_ref.as<float>() + (size_t)layer * D * 12 * D, w.b.as<float>() + (size_t)layer * D,
                                                           w.avg_deg, (int)N);
    FG_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace fg

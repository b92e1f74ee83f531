// PNA layer, ONE kernel per layer: TMA-staged graph-aligned tiles, shared-memory aggregation as the A producer of a
// tcgen05 GEMM, fused combine / relu / residual epilogue.
//
// Reference work per layer (PNA/src/message_passing.cc:88-147, node_embedding.cc:106-215):
//   per (v, d): S = sum h_u, Q = sum h_u^2, min, max over in-edges  ->  mean, min, max, std   (A row of 320 values)
//   acc = b + sum_in [T0 + T1 t + T2 s]  ==  b + G0 + t G1 + s G2  with  G = A [320] x Wcat [320 x 240];  h <- h + relu(acc)
//
// pna_tc.cu runs this as aggregate kernel -> 1,280 B per node of bf16 hi/lo A blocks through HBM -> GEMM kernel: 5.3 x the
// algorithmic DRAM traffic, and two kernels that each take as long as the tensor work alone.  Here the aggregation is the
// A producer INSIDE the GEMM kernel (the reference runs NT and MP of a layer in one DATAFLOW region,
// PNA/src/conv_layer.cc:37-104):
//   * tiles are whole graphs packed into <= 128 rows (prep.cu::pack_tiles_kernel); a producer thread lands a tile's
//     feature rows in shared memory as five COLUMN SLICES of 16 features with 2-D tensor-map TMA copies
//     (cp.async.bulk.tensor.2d, box 16 x 128): slice c is all chunk c of the GEMM needs, so the next tile's slice c is
//     requested the moment the gather warps leave chunk c -- the row stage is single-buffered and still never waited for;
//   * K is permuted so that one K chunk of 64 holds ALL FOUR aggregates of 16 feature columns
//     (k' = (d / 16) * 64 + aggregate * 16 + d % 16): 16 gather warps (4 lanes per row, a float4 of columns per lane) walk
//     the in-edges of their rows in CSR order in shared memory, finish mean / min / max / std, split to bf16 hi/lo and
//     write the chunk's [128 x 64] A block in the canonical K-major layout -- one chunk at a time, into a two-stage ring;
//   * a second producer thread streams the weights ([240 x 32] half chunks, bf16 hi | lo, 30 KB, L2-resident) through a
//     FOUR-stage ring of their own (two stages of 61 KB left the tensor pipe waiting for the L2 round trip of every chunk);
//     the MMA thread runs hi*hi + lo*hi + hi*lo (3 x 4 SS tcgen05.mma per chunk, M = 128, N = 240) into one of two
//     256-column accumulators in tensor memory; four epilogue warps combine the three scaler groups of the other one.
// Rows the tensor path must not touch (out-degree 0, non-finite aggregates: SURVEY.md F6) are flagged here and evaluated
// in fp32 by pna_exact_rows_kernel (pna_tc.cu) afterwards, exactly as before.
#include "internal.cuh"
#include "layers.cuh"
#include "tc.cuh"

#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <string>
#include <type_traits>

namespace fg {

namespace {

constexpr int D = 80;
constexpr int KA = 4 * D;                    // 320
constexpr int NC = 3 * D;                    // 240
constexpr int TM = 128;                      // rows per tile (UMMA M)
constexpr int KC = 64;                       // K per chunk = 16 feature columns x 4 aggregates
constexpr int NCHUNK = KA / KC;              // 5
constexpr int A_HALF = TM * KC * 2;          // 16,384: one hi or lo block of A
constexpr int A_BLOCK = 2 * A_HALF;          // 32,768
constexpr int B_HALF = NC * KC * 2;          // 30,720
constexpr int B_BLOCK = 2 * B_HALF;          // 61,440
constexpr int LBO_A = TM * 16, LBO_B = NC * 16;
constexpr int ROW_BYTES = D * 4;             // 320
constexpr int A_STAGES = 2;
constexpr int KH = 32;                       // K per weight stage (half a chunk)
constexpr int W_HALF = NC * KH * 2;          // 15,360: one hi or lo block of a weight half chunk
constexpr int W_BLOCK = 2 * W_HALF;          // 30,720
constexpr int W_STAGES = 4;
constexpr int NSLICE = NCHUNK;               // column slices of the row stage: 16 features each
constexpr int SLICE_ROW_BYTES = 16 * 4;      // 64
constexpr int SLICE_BYTES = TM * SLICE_ROW_BYTES;                 // 8,192
constexpr int DESC_BYTES = TM * 16;          // 2,048
constexpr int GATHER_WARPS = 16;
// warps 0-15 gather, 16 rows producer, 17 weight producer, 18 MMA issuer, 19 idle, 20-23 epilogue (TMEM lane group = warp % 4)
constexpr int ROWS_WARP = GATHER_WARPS, W_WARP = ROWS_WARP + 1, MMA_WARP = W_WARP + 1, EPI_WARP0 = 20;
constexpr int NT = (EPI_WARP0 + 4) * 32;     // 768
constexpr uint32_t TMEM_COLS = 512;          // two accumulators of 256 columns (240 used)
constexpr float FM_MAX = 32.0f - 0.0009765625f, FM_MIN = -32.0f;      // ap_fixed_max / ap_fixed_min of ap_fixed<16,6> (PNA/src/util.h:34-46)

enum { BAR_W_FULL = 0 /* 4 */, BAR_W_EMPTY = 4 /* 4 */, BAR_A_FULL = 8 /* 2 */, BAR_A_EMPTY = 10 /* 2 */, BAR_ACC_FULL = 12 /* 2 */, BAR_ACC_EMPTY = 14 /* 2 */,
       BAR_S_FULL = 16 /* 5 */, BAR_S_FREE = 21 /* 5 */, BAR_D_FULL = 26, BAR_COUNT = 27 };
struct Smem {
    static constexpr int W = 0;                                   // [4][hi 15,360 | lo 15,360] weight half chunks
    static constexpr int A = W + W_STAGES * W_BLOCK;              // [2][hi 16,384 | lo 16,384] A chunks
    static constexpr int ROWS = A + A_STAGES * A_BLOCK;           // [5 slices][128 rows][16] fp32 feature rows of the tile
    static constexpr int DESC = ROWS + NSLICE * SLICE_BYTES;      // [128] int4 row descriptors (degree-ordered)
    static constexpr int TILE = DESC + DESC_BYTES;                // [4] int2 tile records
    static constexpr int BAR = TILE + 4 * 8;
    static constexpr int TMEM_PTR = BAR + BAR_COUNT * 8;
    static constexpr int BYTES = TMEM_PTR + 16;
};
static_assert(Smem::ROWS % 1024 == 0 && SLICE_BYTES % 1024 == 0 && Smem::A % 1024 == 0, "alignment of the (swizzled) TMA / UMMA operands");
static_assert(Smem::BYTES <= 232448, "shared memory budget");

struct PnaFusedParams {
    const float* h_in; float* h_out;
    const int* in_ptr; const int* src;
    const int4* row_desc;            // sorted: per tile ordered by in-degree, row position in .y bits 24..30; else in node order (prep.cu)
    int sorted;
    const int2* tiles; const int* tile_count;
    const unsigned char* wpack;      // [5][61440] this layer, K permuted (pna_fused_pack_layer)
    const float* b;                  // [80]
    const int* out_deg;
    unsigned char* nonfinite;        // [N] zeroed by the caller; set here for rows with a non-finite aggregate
    float avg_deg;
};

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
// The same issued by the elected lane of a CONVERGED warp (elect.sync inside the asm block): the loop around it runs on all 32
// lanes in the uniform datapath and ptxas emits one predicated UTCHMMA per MMA -- issued from a `lane == 0` branch every MMA
// is wrapped in an elect-and-retry loop with R2UR moves (~150 cycles per MMA, more than the tensor pipe needs to execute it)
__device__ __forceinline__ void mma_ss_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xFFFFFFFF;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
__device__ __forceinline__ void commit_elect(uint64_t* bar)
{
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xFFFFFFFF;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
        : "memory");
}
// Shared-memory address of dynamic shared memory in a kernel without static shared memory (the kernel checks it): with it
// every operand descriptor of the MMA issuer is a compile-time constant.
constexpr uint32_t SMEM_BASE = 0x400;
__host__ __device__ constexpr uint64_t desc_imm(uint32_t off, uint32_t lbo)
{
    return (uint64_t)(((SMEM_BASE + off) >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
}
// A chunks use the 128-byte-swizzle K-major layout (row = 128 contiguous bytes, 16-byte unit u of row r at u ^ (r % 8); 8-row groups
// 1,024 bytes apart, a K = 16 step advances the start address by 32 bytes): the eight rows of a gather warp's store then fall into
// eight different bank groups, where the no-swizzle layout (8-byte pieces at a 2,048-byte stride) cost a 2-way conflict per store
__host__ __device__ constexpr uint64_t desc_imm_sw128(uint32_t off)
{
    return (uint64_t)(((SMEM_BASE + off) >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F&& f)
{
    if constexpr (I < N)
    {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo)
{
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    const float r0 = x0 - __uint_as_float(hi << 16);
    const float r1 = x1 - __uint_as_float(hi & 0xFFFF0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ int4 lds_i4(uint32_t addr)
{
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_v2(uint32_t addr, uint32_t a, uint32_t b)
{
    asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ int2 tile_of(const PnaFusedParams& p, int t, int ntiles)
{
    int2 v = make_int2(0, 0);
    if (t < ntiles) asm volatile("ld.global.nc.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p.tiles + t));
    return v;
}

// S, Q, min, max of one source row's float4.  Explicit round-to-nearest multiplies and adds, no FMA contraction: this is
// the reference's statement order in IEEE fp32 (PNA/src/message_passing.cc:127-133; oracle build -ffp-contract=off), so the
// aggregates -- and the cancellation in the variance below -- are the reference's bit for bit.
struct Agg4 { float s[4], q[4], mn[4], mx[4]; };
__device__ __forceinline__ void agg_init(Agg4& a)
{
#pragma unroll
    for (int j = 0; j < 4; j++) { a.s[j] = 0.f; a.q[j] = 0.f; a.mn[j] = FM_MAX; a.mx[j] = FM_MIN; }
}
__device__ __forceinline__ void agg_add(Agg4& a, const float4& h)
{
    const float x[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
        a.s[j] = __fadd_rn(a.s[j], x[j]);
        a.q[j] = __fadd_rn(a.q[j], __fmul_rn(x[j], x[j]));
        if (x[j] < a.mn[j]) a.mn[j] = x[j];
        if (x[j] > a.mx[j]) a.mx[j] = x[j];
    }
}

__global__ void __launch_bounds__(NT, 1) pna_layer_fused_kernel(const __grid_constant__ CUtensorMap tmap_h, PnaFusedParams p)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + Smem::BAR);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + Smem::TMEM_PTR);
    int2* tile_rec = reinterpret_cast<int2*>(smem + Smem::TILE);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0)
    {
        for (int i = 0; i < W_STAGES; i++) { mbar_init(&bar[BAR_W_FULL + i], 1); mbar_init(&bar[BAR_W_EMPTY + i], 1); }
        for (int i = 0; i < 2; i++)
        {
            mbar_init(&bar[BAR_A_FULL + i], GATHER_WARPS);
            mbar_init(&bar[BAR_A_EMPTY + i], 1);
            mbar_init(&bar[BAR_ACC_FULL + i], 1);
            mbar_init(&bar[BAR_ACC_EMPTY + i], 128);
        }
        for (int i = 0; i < NSLICE; i++) { mbar_init(&bar[BAR_S_FULL + i], 1); mbar_init(&bar[BAR_S_FREE + i], GATHER_WARPS); }
        mbar_init(&bar[BAR_D_FULL], 1);
        fence_mbar_init();
    }
    if (warp == MMA_WARP)
    {
        tc::tmem_alloc(tmem_ptr, TMEM_COLS);
        tc::tmem_relinquish();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = *tmem_ptr;
    const int ntiles = __ldg(p.tile_count);
    const int first = blockIdx.x, step = gridDim.x;

    if (warp < GATHER_WARPS)
    {
        // ===== gather warps: thread = (one of the warp's 8 rows, lane q of 4) -> 16 A values per chunk =====
        const int q = lane & 3;
        const uint32_t rows_base = smem_u32(smem + Smem::ROWS), desc_base = smem_u32(smem + Smem::DESC);
        uint32_t it = 0, g = 0;
        for (int t = first; t < ntiles; t += step, it++)
        {
            mbar_wait_park(&bar[BAR_D_FULL], it & 1);
            const int2 ti = tile_rec[it & 3];
            const int start = ti.x, rows = ti.y & 0xFFFF;
            const bool ext = (ti.y >> 30) & 1;
            // slot -> row through the degree-ordered descriptors; lanes 0..15 of a warp take four slots of the lower half (small
            // in-degrees), lanes 16..31 four of the upper half, so that every warp gets the same mix (a chunk waits for the slowest)
            const int slot = p.sorted ? (lane >> 4) * (TM / 2) + warp * 4 + ((lane >> 2) & 3) : warp * 8 + (lane >> 2);
            const bool live = slot < rows;
            const int4 d = live ? lds_i4(desc_base + slot * 16) : make_int4(0, 0, 0, 0);
            const int R = (live && p.sorted) ? ((d.y >> 24) & 0x7F) : slot;
            const int node = start + (live ? R : 0);
            int deg = live ? (int)((unsigned)d.x >> 24) : 0;
            if (deg == 255) deg = __ldg(p.in_ptr + node + 1) - __ldg(p.in_ptr + node);
            const int rel[4] = {(d.x & 0xFFFF) - 32768, (d.y & 0xFFFF) - 32768, (d.z & 0xFFFF) - 32768, (d.w & 0xFFFF) - 32768};
            const int eb = deg > 4 || ext ? __ldg(p.in_ptr + node) : 0;
            const float fn = (float)(deg == 0 ? 1 : deg);
            // x / n as three instructions: q0 = x r, q = q0 + (x - q0 n) r with r = RN(1 / n) (one Newton step on the MUFU
            // approximation).  For a correctly rounded reciprocal this is the correctly rounded quotient (Markstein); checked
            // against IEEE division on 8e7 random operands for n = 1..40 without a mismatch.  The IEEE division routine takes
            // its slow path whenever a numerator is 0 -- and after relu most of them are -- which made the division a third of
            // the gather's instructions.
            float rn = __frcp_rn(fn);
            auto div_n = [&](float x) { const float q0 = __fmul_rn(x, rn); return __fmaf_rn(__fmaf_rn(-q0, fn, x), rn, q0); };
            float finite_probe = 0.f;                     // stays 0 while every sum is finite (x * 0 is NaN for inf and NaN)
#pragma unroll 1
            for (int c = 0; c < NCHUNK; c++, g++)
            {
                const uint32_t s = g & 1;
                Agg4 a;
                agg_init(a);
                mbar_wait_park(&bar[BAR_S_FULL + c], it & 1);                          // column slice c of this tile has landed
                if (!ext)
                {
                    // the slices land with the 64-byte TMA swizzle: 16-byte unit u of row r sits at unit u ^ ((r >> 1) & 3).  With plain
                    // 64-byte rows the eight rows of a load instruction fall into two bank groups (4-way conflicts); swizzled, into eight
                    const uint32_t sl = rows_base + c * SLICE_BYTES;
                    auto row_addr = [&](int r) { return sl + (uint32_t)r * SLICE_ROW_BYTES + (uint32_t)((q ^ ((r >> 1) & 3)) << 4); };
#pragma unroll
                    for (int e = 0; e < 4; e++)
                        if (e < deg) agg_add(a, lds_f4(row_addr(R + rel[e])));
                    for (int e = 4; e < deg; e++)
                        agg_add(a, lds_f4(row_addr(__ldg(p.src + eb + e) - start)));
                }
                else
                {
                    // a graph of more than 128 nodes: its sources may lie outside the stage
                    for (int e = 0; e < deg; e++)
                        agg_add(a, ldg_f4(p.h_in + (size_t)__ldg(p.src + eb + e) * D + 16 * c + 4 * q));
                }
                // this warp is done with slice c: the producer may fetch slice c of the next tile
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar[BAR_S_FREE + c]);
                float mean[4], sd[4];
#pragma unroll
                for (int j = 0; j < 4; j++)
                {
                    // mean = S / n;  std = sqrt(relu(Q / n - mean^2))  (PNA/src/node_embedding.cc:123-150), statement by statement
                    mean[j] = div_n(a.s[j]);
                    const float var = relu_f(__fsub_rn(div_n(a.q[j]), __fmul_rn(mean[j], mean[j])));
                    float sq;
                    asm("sqrt.approx.f32 %0, %1;" : "=f"(sq) : "f"(var));      // <= 1 ulp from the IEEE root; the value goes through a bf16 split anyway
                    sd[j] = var == 0.f ? 0.f : sq;
                    finite_probe = fmaf(a.s[j], 0.f, fmaf(a.q[j], 0.f, finite_probe));
                }
                // the stage's A block must be free (the MMAs of chunk g - 2 have read it)
                if (g >= 2) mbar_wait_park(&bar[BAR_A_EMPTY + s], ((g >> 1) - 1) & 1);
                // aggregator_t order (PNA/src/dcl.h:29-35): mean, min, max, std; kk = aggregate * 16 + 4 q + j inside the chunk
                const uint32_t arow = (uint32_t)(live ? R : slot);
                const uint32_t a_row = smem_u32(smem + Smem::A + s * A_BLOCK) + arow * 128 + (q & 1) * 8;
                auto put = [&](int ag, const float (&x)[4]) {
                    uint32_t h0 = 0, l0 = 0, h1 = 0, l1 = 0;
                    if (live) { split2(x[0], x[1], h0, l0); split2(x[2], x[3], h1, l1); }
                    const uint32_t dst = a_row + ((((uint32_t)(2 * ag + (q >> 1))) ^ (arow & 7)) << 4);
                    sts_v2(dst, h0, h1);
                    sts_v2(dst + A_HALF, l0, l1);
                };
                put(0, mean); put(1, a.mn); put(2, a.mx); put(3, sd);
                // rows with a non-finite aggregate (an inf or NaN among the sources: then S or Q is not finite) are left to the
                // fp32 kernel; the epilogue of this tile reads the flag, so it is written before the arrival that lets the last
                // chunk's MMAs (and with them the accumulator barrier) complete
                if (c == NCHUNK - 1 && live && !(finite_probe == 0.f)) *(volatile unsigned char*)(p.nonfinite + node) = 1;
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar[BAR_A_FULL + s]);
            }
        }
    }
    else if (warp == ROWS_WARP)
    {
        // ===== rows producer: descriptors and the five column slices of the next tile, each as soon as the gather warps have left it =====
        if (lane == 0)
        {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_h) : "memory");
            uint32_t it = 0;
            for (int t = first; t < ntiles; t += step, it++)
            {
                const int2 ti = tile_of(p, t, ntiles);
                const int rows = ti.y & 0xFFFF;
                for (int c = 0; c < NSLICE; c++)
                {
                    if (it >= 1) mbar_wait_park(&bar[BAR_S_FREE + c], (it - 1) & 1);
                    if (c == 0)
                    {
                        // every gather warp is past chunk 0 of the previous tile, so past its descriptor and tile-record reads
                        tile_rec[it & 3] = ti;
                        mbar_arrive_expect_tx(&bar[BAR_D_FULL], (uint32_t)rows * 16);
                        if (rows) tma_load_1d(smem + Smem::DESC, p.row_desc + ti.x, (uint32_t)rows * 16, &bar[BAR_D_FULL]);
                    }
                    mbar_arrive_expect_tx(&bar[BAR_S_FULL + c], SLICE_BYTES);
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                                     smem_u32(smem + Smem::ROWS + c * SLICE_BYTES)),
                                 "l"(&tmap_h), "r"(16 * c), "r"(ti.x), "r"(smem_u32(&bar[BAR_S_FULL + c]))
                                 : "memory");
                }
                const int2 tn = tile_of(p, t + 2 * step, ntiles);
                if (tn.y & 0xFFFF)
                {
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.h_in + (size_t)tn.x * D), "r"((tn.y & 0xFFFF) * ROW_BYTES) : "memory");
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.row_desc + tn.x), "r"((tn.y & 0xFFFF) * 16) : "memory");
                }
            }
        }
    }
    else if (warp == W_WARP)
    {
        // ===== weight producer: [240 x 32] hi | lo half chunks through a four-stage ring =====
        if (lane == 0)
        {
            uint32_t wh = 0;
            for (int t = first; t < ntiles; t += step)
                for (int c2 = 0; c2 < 2 * NCHUNK; c2++, wh++)
                {
                    const uint32_t ws = wh & (W_STAGES - 1);
                    if (wh >= W_STAGES) mbar_wait_park(&bar[BAR_W_EMPTY + ws], ((wh / W_STAGES) - 1) & 1);
                    mbar_arrive_expect_tx(&bar[BAR_W_FULL + ws], W_BLOCK);
                    tma_load_1d(smem + Smem::W + ws * W_BLOCK, p.wpack + (size_t)c2 * W_BLOCK, W_BLOCK, &bar[BAR_W_FULL + ws]);
                }
        }
    }
    else if (warp == MMA_WARP)
    {
        // ===== MMA issuer: the converged warp, tcgen05 instructions predicated on elect.sync, every shared-memory descriptor a
        // compile-time constant (A stage S = chunk parity; its two weight half chunks sit in weight stages 2 S and 2 S + 1) =====
        {
            constexpr uint32_t IDESC = tc::idesc_bf16(TM, NC);
            if ((smem_u32(smem) & 0xFFFFFFu) != SMEM_BASE) asm volatile("trap;");
            auto mma_chunk = [&](auto SSEL, uint32_t d_tmem, bool first, uint32_t g) {
                constexpr int S = decltype(SSEL)::value;
                mbar_wait_park(&bar[BAR_A_FULL + S], (g >> 1) & 1);
                static_for<0, 2>([&](auto H) {
                    constexpr int h = decltype(H)::value, ws = 2 * S + h;
                    // weight half chunk wh = 2 g + h is the (g >> 1)-th use of stage ws
                    mbar_wait_park(&bar[BAR_W_FULL + ws], (g >> 1) & 1);
                    tc::fence_after_sync();
                    static_for<0, KH / 16>([&](auto JJ) {
                        constexpr int jj = decltype(JJ)::value, j = h * (KH / 16) + jj;                  // k-step inside the A chunk
                        constexpr uint64_t a_hi = desc_imm_sw128(Smem::A + S * A_BLOCK + 32 * j);
                        constexpr uint64_t a_lo = desc_imm_sw128(Smem::A + S * A_BLOCK + A_HALF + 32 * j);
                        constexpr uint64_t b_hi = desc_imm(Smem::W + ws * W_BLOCK + 2 * jj * LBO_B, LBO_B);
                        constexpr uint64_t b_lo = desc_imm(Smem::W + ws * W_BLOCK + W_HALF + 2 * jj * LBO_B, LBO_B);
                        mma_ss_elect(d_tmem, a_hi, b_hi, IDESC, !(first && j == 0));
                        mma_ss_elect(d_tmem, a_lo, b_hi, IDESC, true);
                        mma_ss_elect(d_tmem, a_hi, b_lo, IDESC, true);
                    });
                    commit_elect(&bar[BAR_W_EMPTY + ws]);          // the weight stage is free once these MMAs have read it
                });
                commit_elect(&bar[BAR_A_EMPTY + S]);
            };
            uint32_t g = 0, it = 0;
            for (int t = first; t < ntiles; t += step, it++)
            {
                const uint32_t acc = it & 1;
                if (it >= 2) mbar_wait_park(&bar[BAR_ACC_EMPTY + acc], ((it >> 1) - 1) & 1);
                tc::fence_after_sync();
                const uint32_t d_tmem = tbase + acc * 256;
                // five chunks per tile: the chunk parity alternates across tiles, so two unrolled bodies by the parity of g
                for (int c = 0; c < NCHUNK; c++, g++)
                {
                    if (g & 1) mma_chunk(std::integral_constant<int, 1>{}, d_tmem, c == 0, g);
                    else mma_chunk(std::integral_constant<int, 0>{}, d_tmem, c == 0, g);
                }
                commit_elect(&bar[BAR_ACC_FULL + acc]);
            }
        }
    }
    else if (warp >= EPI_WARP0)
    {
        // ===== epilogue: thread = row (TMEM lane), 16 output columns per step =====
        const int lg = warp & 3;                                  // TMEM lane group this warp may read
        const int row = lg * 32 + lane;
        uint32_t it = 0;
        for (int t = first; t < ntiles; t += step, it++)
        {
            const uint32_t acc = it & 1;
            mbar_wait_park(&bar[BAR_ACC_FULL + acc], (it >> 1) & 1);
            tc::fence_after_sync();
            const int2 ti = tile_rec[it & 3];
            const int rows = ti.y & 0xFFFF;
            const bool live = row < rows;
            const int v = ti.x + (live ? row : 0);
            int od = 0;
            bool skip = true;
            if (live)
            {
                od = __ldg(p.out_deg + v);
                // the flag was written (by this CTA's gather warps) before the arrival that completed the accumulator barrier
                skip = od == 0 || *(volatile unsigned char*)(p.nonfinite + v) != 0;
            }
            const float log_degree = logf((float)(od + 1));
            const float tt = log_degree / p.avg_deg;
            float scale = p.avg_deg / log_degree;
            if (scale == 0) scale = 1;
            const uint32_t taddr = tbase + acc * 256 + ((uint32_t)(lg * 32) << 16);
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
                        o.x = hv.x + relu_f(bb.x + __uint_as_float(g0[j]) + (__uint_as_float(g1[j]) * tt + __uint_as_float(g2[j]) * scale));
                        o.y = hv.y + relu_f(bb.y + __uint_as_float(g0[j + 1]) + (__uint_as_float(g1[j + 1]) * tt + __uint_as_float(g2[j + 1]) * scale));
                        o.z = hv.z + relu_f(bb.z + __uint_as_float(g0[j + 2]) + (__uint_as_float(g1[j + 2]) * tt + __uint_as_float(g2[j + 2]) * scale));
                        o.w = hv.w + relu_f(bb.w + __uint_as_float(g0[j + 3]) + (__uint_as_float(g1[j + 3]) * tt + __uint_as_float(g2[j + 3]) * scale));
                        stg_f4_stream(hout + j, o);
                    }
                }
            }
            tc::fence_before_sync();
            mbar_arrive(&bar[BAR_ACC_EMPTY + acc]);
        }
    }

    tc::fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP) tc::tmem_dealloc(tbase, TMEM_COLS);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library links cudart only)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int encode_rows_map(CUtensorMap* map, const float* h, long num_nodes)
{
    static EncodeTiledFn fn = nullptr;
    if (!fn)
    {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        FG_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
        if (!ptr || qres != cudaDriverEntryPointSuccess) { set_last_error("cuTensorMapEncodeTiled is not available in this driver"); return FG_ERR_STATE; }
        fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    // h as a 2-D tensor [num_nodes rows][80 features] fp32; box = 16 features x 128 rows -> a dense [128][16] slice in shared memory;
    // rows beyond the tensor read as zero
    const cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)num_nodes};
    const cuuint64_t gstride[1] = {(cuuint64_t)ROW_BYTES};
    const cuuint32_t box[2] = {16, (cuuint32_t)TM};
    const cuuint32_t estride[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(h), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r)); return FG_ERR_STATE; }
    return 0;
}

}  // namespace

size_t pna_fused_pack_bytes() { return (size_t)2 * NCHUNK * W_BLOCK; }

// wcat [320][240] (k = aggr*80 + in, n = scaler*80 + out) -> K permuted to k' = (in / 16) * 64 + aggr * 16 + in % 16, per K
// half chunk of 32 the [240 x 32] block as bf16 hi | lo in the canonical K-major layout byte(n, k) = (k / 8) * 3840 + n * 16 + (k % 8) * 2
void pna_fused_pack_layer(const float* wcat, unsigned char* dst, uint16_t (*bf16_rn)(float), float (*bf16_to_float)(uint16_t))
{
    for (int ag = 0; ag < 4; ag++)
        for (int i = 0; i < D; i++)
            for (int n = 0; n < NC; n++)
            {
                const float x = wcat[(size_t)(ag * D + i) * NC + n];
                const uint16_t hi = bf16_rn(x), lo = bf16_rn(x - bf16_to_float(hi));
                const int kp = (i / 16) * KC + ag * 16 + i % 16, c2 = kp / KH, kk = kp % KH;
                unsigned char* o = dst + (size_t)c2 * W_BLOCK + (size_t)(kk / 8) * LBO_B + (size_t)n * 16 + (size_t)(kk % 8) * 2;
                o[0] = (unsigned char)(hi & 0xFF); o[1] = (unsigned char)(hi >> 8);
                o[W_HALF] = (unsigned char)(lo & 0xFF); o[W_HALF + 1] = (unsigned char)(lo >> 8);
            }
}

int pna_layer_fused_launch(DeviceBatch& b, const PnaWeights& w, int layer, const float* h_in, float* h_out, int sm_count, cudaStream_t s)
{
    FG_TRY(opt_in_smem(reinterpret_cast<const void*>(&pna_layer_fused_kernel), Smem::BYTES));
    PnaFusedParams p{};
    p.h_in = h_in; p.h_out = h_out;
    p.in_ptr = b.in_ptr.as<int>(); p.src = b.src.as<int>();
    static const int sorted_env = [] { const char* e = std::getenv("FLOWGNN_B200_PNA_SORTED"); return e ? std::atoi(e) : 0; }();
    p.sorted = sorted_env;
    p.row_desc = p.sorted ? b.row_desc_sorted.as<int4>() : b.row_desc.as<int4>();
    p.tiles = b.tiles.as<int2>(); p.tile_count = b.tile_count.as<int>();
    p.wpack = w.wpack_fused.as<unsigned char>() + (size_t)layer * pna_fused_pack_bytes();
    p.b = w.b.as<float>() + (size_t)layer * D;
    p.out_deg = b.out_deg.as<int>();
    p.nonfinite = b.nonfinite.as<unsigned char>();
    p.avg_deg = w.avg_deg;
    alignas(64) CUtensorMap tmap;
    FG_TRY(encode_rows_map(&tmap, h_in, b.total_nodes));
    const int grid = (int)std::max<long>(1, std::min<long>(b.max_tiles, sm_count));
    pna_layer_fused_kernel<<<grid, NT, Smem::BYTES, s>>>(tmap, p);
    FG_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace fg

// GIN / GIN-VN layer on the B200 tensor cores: one persistent, warp-specialised kernel per layer.
//
// Reference work per layer (GIN/src/message_passing.cc:77-150, node_embedding.cc:23-201):
//   m_v = sum_{(u,v)} relu(h_u + EE_l[attr_uv]);  a_v = m_v + h_v  (eps is never loaded: SURVEY.md F4)
//   z = relu(W1 a + b1);  h'_v = W2 z + b2  (+ relu unless last layer)
//
// B200 mapping (one CTA per SM, tiles of 128 consecutive nodes, UMMA M = 128, cta_group::1):
//   * W1 / W2 are STATIONARY in shared memory for the whole launch, each as a bf16 hi + bf16 lo pair
//     (no-swizzle K-major canonical layout, 4 x 46,592 B, loaded once by bulk TMA);
//   * activations flow through TENSOR MEMORY: 8 gather warps build a_v in fp32 (coalesced float4 gathers,
//     CSR order -> deterministic), split it into bf16 hi/lo and hand it over through a small shared-memory
//     transposition buffer into TMEM (tcgen05.st) as the A operand of GEMM1;
//   * GEMM1 = 3 x 7 tcgen05.mma (hi*hi + lo*hi + hi*lo, fp32 accumulate in TMEM, N = 208);
//     4 epilogue warps read z from TMEM, add b1, relu, split to bf16 hi/lo and write it back IN PLACE: the
//     16 fp32 columns of k-step j become its 8 hi + 8 lo operand columns for GEMM2 (3 x 13 MMAs, N = 112);
//   * the second epilogue adds b2 (+ relu) and streams h' to HBM.
// The 3-product bf16 split keeps the fp32 contract (|y - y_ref| <= 1e-4 * max(1,|y_ref|); measured 7e-6 on
// molhiv, 1.7e-5 on GIN-VN, tools/split_precision_probe.py) at 2.25 PFLOP/s-class tensor throughput.
#include "internal.cuh"
#include "layers.cuh"
#include "tc.cuh"

#include <algorithm>

namespace fg {

namespace {

constexpr int D = 100;
constexpr int Q = D / 4;
constexpr int TM = 128;                       // nodes per tile
constexpr int K1 = 112, N1 = 208;             // GEMM1  [TM x K1] * [N1 x K1]^T
constexpr int K2 = 208, N2 = 112;             // GEMM2  [TM x K2] * [N2 x K2]^T
constexpr int WBLOCK = N1 * K1 * 2;           // bytes of one bf16 weight block (N1*K1 == N2*K2)
static_assert(N1 * K1 == N2 * K2, "weight blocks share a size");

constexpr int EPI_WARPS = 4, GATHER_WARPS = 8;
constexpr int MMA_WARP = EPI_WARPS + GATHER_WARPS;
constexpr int NT = (MMA_WARP + 1) * 32;       // 416 threads

// tensor-memory columns
constexpr uint32_t TC_A1_HI = 0, TC_A1_LO = 56, TC_Z = 128, TC_H = 384;
constexpr uint32_t TMEM_COLS = 512;

// gather K-slices (float4 chunks): team T handles slices T and T + 2
constexpr int SLICE_Q = 7;
constexpr int STAGE_LO = 16;                  // per row: 14 hi words at 0, 14 lo words at 16 (16-byte aligned: the 8-word
constexpr int STAGE_WORDS = 36;               // TMEM store operands then come from aligned LDS.128); stride = 4 x odd words
constexpr int STAGE_BYTES = TM * STAGE_WORDS * 4;

struct Smem {
    static constexpr int W = 0;                                   // W1_hi | W1_lo | W2_hi | W2_lo
    static constexpr int EE = W + 4 * WBLOCK;                     // [13][100] fp32
    static constexpr int B1 = EE + ED_FEATURE_PER_LAYER * D * 4;  // [208]
    static constexpr int B2 = B1 + N1 * 4;                        // [112]
    static constexpr int STAGE = B2 + N2 * 4;                     // 2 teams
    static constexpr int BAR = STAGE + 2 * STAGE_BYTES;
    static constexpr int TMEM_PTR = BAR + 8 * 8;
    static constexpr int BYTES = TMEM_PTR + 16;
};
static_assert(Smem::STAGE % 16 == 0 && Smem::BAR % 8 == 0, "alignment");
static_assert(Smem::BYTES <= 232448, "shared memory budget");

enum { BAR_W = 0, BAR_A1_FULL, BAR_G1_DONE, BAR_A2_FULL, BAR_G2_DONE };

struct GinTcParams {
    const float* h_in; float* h_out;
    const int* in_ptr; const int* src; const uint8_t* code;
    const float* ee_raw;             // [13][100] this layer
    const unsigned char* wpack;      // 4 x WBLOCK this layer
    const float* b1; const float* b2;   // [208], [112] zero padded
    int num_nodes; int num_tiles; int relu_out;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void team_sync(int team) { asm volatile("bar.sync %0, 128;" ::"r"(team + 1) : "memory"); }

// (x0, x1) -> packed bf16 pairs: hi = rn(x), lo = rn(x - hi); element 0 in the low half
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo)
{
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    const float r0 = x0 - __uint_as_float(hi << 16);
    const float r1 = x1 - __uint_as_float(hi & 0xFFFF0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}

__device__ __forceinline__ void st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st2(uint32_t taddr, uint32_t a, uint32_t b)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(a), "r"(b) : "memory");
}

// One K-slice of the tile's A operand: NQ float4 chunks starting at chunk Q0 for all TM rows, computed by
// the 128 threads of a team and written to the team's transposition buffer as packed bf16 (hi | lo) words.
template <int NQ>
__device__ __forceinline__ void gather_slice(const GinTcParams& p, const float* ee, uint32_t* stage, int tt, int n0, int rows, int q0)
{
    for (int item = tt; item < TM * NQ; item += 128)
    {
        const int v = item / NQ, qq = item - v * NQ;
        const int q = q0 + qq;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (v < rows)
        {
            const int node = n0 + v;
            const int eb = __ldg(p.in_ptr + node), ee_end = __ldg(p.in_ptr + node + 1);
            float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int e = eb; e < ee_end; e++)
            {
                const int u = __ldg(p.src + e);
                const int c = __ldg(p.code + e);
                const int a0 = c / 12, r = c - a0 * 12;
                const float4 hu = ldg_f4(p.h_in + (size_t)u * D + 4 * q);
                // ((0 + T[a0]) + T[5 + a1]) + T[11 + a2]   (GIN/src/message_passing.cc:136-142)
                const float4 t0 = ld_f4(ee + a0 * D + 4 * q);
                const float4 t1 = ld_f4(ee + (5 + (r >> 1)) * D + 4 * q);
                const float4 t2 = ld_f4(ee + (11 + (r & 1)) * D + 4 * q);
                const float4 t = make_float4((t0.x + t1.x) + t2.x, (t0.y + t1.y) + t2.y, (t0.z + t1.z) + t2.z, (t0.w + t1.w) + t2.w);
                m.x += relu_f(t.x + hu.x); m.y += relu_f(t.y + hu.y); m.z += relu_f(t.z + hu.z); m.w += relu_f(t.w + hu.w);
            }
            const float4 hv = ldg_f4(p.h_in + (size_t)node * D + 4 * q);
            a = make_float4(m.x + hv.x, m.y + hv.y, m.z + hv.z, m.w + hv.w);
        }
        uint32_t h0, l0, h1, l1;
        split2(a.x, a.y, h0, l0);
        split2(a.z, a.w, h1, l1);
        uint32_t* row = stage + v * STAGE_WORDS;
        *reinterpret_cast<uint2*>(row + 2 * qq) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(row + STAGE_LO + 2 * qq) = make_uint2(l0, l1);
    }
}

// row `lane` of the team's buffer -> TMEM columns [col0, col0 + NW) of A1_hi and A1_lo (NW = 14 or 8 words)
template <int NW>
__device__ __forceinline__ void stage_to_tmem(const uint32_t* stage, int row, uint32_t lane_base, uint32_t col0)
{
    const uint4* r4 = reinterpret_cast<const uint4*>(stage + row * STAGE_WORDS);
    uint32_t w[STAGE_WORDS];
#pragma unroll
    for (int i = 0; i < STAGE_WORDS / 4; i++)
    {
        const uint4 x = r4[i];
        w[4 * i] = x.x; w[4 * i + 1] = x.y; w[4 * i + 2] = x.z; w[4 * i + 3] = x.w;
    }
#pragma unroll
    for (int part = 0; part < 2; part++)
    {
        const uint32_t t = lane_base + (part ? TC_A1_LO : TC_A1_HI) + col0;
        const uint32_t* s = w + part * STAGE_LO;
        const uint32_t v8[8] = {s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7]};
        tc::st8(t, v8);
        if (NW == 14)
        {
            st4(t + 8, s[8], s[9], s[10], s[11]);
            st2(t + 12, s[12], s[13]);
        }
    }
}

__global__ void __launch_bounds__(NT, 1) gin_layer_tc_kernel(GinTcParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    float* ee = reinterpret_cast<float*>(smem + Smem::EE);
    float* b1s = reinterpret_cast<float*>(smem + Smem::B1);
    float* b2s = reinterpret_cast<float*>(smem + Smem::B2);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + Smem::BAR);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + Smem::TMEM_PTR);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0)
    {
        mbar_init(&bar[BAR_W], 1);
        mbar_init(&bar[BAR_A1_FULL], GATHER_WARPS);
        mbar_init(&bar[BAR_G1_DONE], 1);
        mbar_init(&bar[BAR_A2_FULL], EPI_WARPS);
        mbar_init(&bar[BAR_G2_DONE], 1);
        fence_mbar_init();
    }
    if (warp == MMA_WARP)
    {
        tc::tmem_alloc(tmem_ptr, TMEM_COLS);
        tc::tmem_relinquish();
    }
    for (int i = tid; i < ED_FEATURE_PER_LAYER * Q; i += NT) st_f4(ee + 4 * i, ldg_f4(p.ee_raw + 4 * i));
    for (int i = tid; i < N1; i += NT) b1s[i] = __ldg(p.b1 + i);
    for (int i = tid; i < N2; i += NT) b2s[i] = __ldg(p.b2 + i);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = *tmem_ptr;

    if (warp == MMA_WARP)
    {
        // ===== weight loader + MMA issuer (one thread) =====
        if (lane == 0)
        {
            mbar_arrive_expect_tx(&bar[BAR_W], 4 * WBLOCK);
#pragma unroll
            for (int i = 0; i < 4; i++) tma_load_1d(smem + Smem::W + i * WBLOCK, p.wpack + (size_t)i * WBLOCK, WBLOCK, &bar[BAR_W]);
            mbar_wait(&bar[BAR_W], 0);
            const uint32_t w_addr = smem_u32(smem + Smem::W);
            const uint32_t idesc1 = tc::idesc_bf16(TM, N1), idesc2 = tc::idesc_bf16(TM, N2);
            int it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, it++)
            {
                const uint32_t ph = it & 1;
                mbar_wait(&bar[BAR_A1_FULL], ph);
                tc::fence_after_sync();
                bool acc = false;
#pragma unroll
                for (int prod = 0; prod < 3; prod++)
                {
                    const uint32_t a_col = tbase + (prod == 1 ? TC_A1_LO : TC_A1_HI);
                    const uint32_t b_addr = w_addr + (prod == 2 ? WBLOCK : 0);
#pragma unroll
                    for (int j = 0; j < K1 / 16; j++)
                    {
                        tc::mma_ts(tbase + TC_Z, a_col + 8 * j, tc::smem_desc(b_addr + 2 * j * N1 * 16, N1 * 16, 128), idesc1, acc);
                        acc = true;
                    }
                }
                tc::commit(&bar[BAR_G1_DONE]);

                mbar_wait(&bar[BAR_A2_FULL], ph);
                tc::fence_after_sync();
                acc = false;
#pragma unroll
                for (int prod = 0; prod < 3; prod++)
                {
                    const uint32_t a_col = tbase + TC_Z + (prod == 1 ? 8 : 0);
                    const uint32_t b_addr = w_addr + 2 * WBLOCK + (prod == 2 ? WBLOCK : 0);
#pragma unroll
                    for (int j = 0; j < K2 / 16; j++)
                    {
                        tc::mma_ts(tbase + TC_H, a_col + 16 * j, tc::smem_desc(b_addr + 2 * j * N2 * 16, N2 * 16, 128), idesc2, acc);
                        acc = true;
                    }
                }
                tc::commit(&bar[BAR_G2_DONE]);
            }
        }
    }
    else if (warp >= EPI_WARPS)
    {
        // ===== gather warps: build the A operand of GEMM1 in tensor memory =====
        const int gw = warp - EPI_WARPS;          // 0..7
        const int team = gw >> 2;                 // 0 / 1
        const int tt = tid - (EPI_WARPS + 4 * team) * 32;   // 0..127 within the team
        const int quad = warp & 3;
        const uint32_t lane_base = tbase + ((uint32_t)(quad * 32) << 16);
        uint32_t* stage = reinterpret_cast<uint32_t*>(smem + Smem::STAGE + team * STAGE_BYTES);
        const int row = quad * 32 + lane;
        if (team == 0)
        {
            // K padding columns (k = 100..111) of A1 stay zero for the whole launch
            const uint32_t z8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            tc::st8(lane_base + TC_A1_HI + 48, z8);
            tc::st8(lane_base + TC_A1_LO + 48, z8);
            tc::wait_st();
        }
        team_sync(team);
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, it++)
        {
            const int n0 = tile * TM;
            const int rows = min(TM, p.num_nodes - n0);
            if (gw == 0 && lane == 0)
            {
                const int next = tile + gridDim.x;
                if (next < p.num_tiles)
                {
                    const int rn = min(TM, p.num_nodes - next * TM);
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.h_in + (size_t)next * TM * D), "r"(rn * D * 4) : "memory");
                }
            }
#pragma unroll
            for (int si = 0; si < 2; si++)
            {
                const int s = team + 2 * si;
                if (s < 3) gather_slice<SLICE_Q>(p, ee, stage, tt, n0, rows, SLICE_Q * s);
                else gather_slice<Q - 3 * SLICE_Q>(p, ee, stage, tt, n0, rows, SLICE_Q * s);
                team_sync(team);
                if (si == 0 && it > 0)
                {
                    // A1 of the previous tile has been consumed once GEMM1 of that tile completed
                    mbar_wait(&bar[BAR_G1_DONE], (it - 1) & 1);
                    tc::fence_after_sync();
                }
                if (s < 3) stage_to_tmem<14>(stage, row, lane_base, 14 * s);
                else stage_to_tmem<8>(stage, row, lane_base, 14 * s);
                team_sync(team);
            }
            tc::wait_st();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar[BAR_A1_FULL]);
        }
    }
    else
    {
        // ===== epilogue warps: thread = tile row (TMEM lane) =====
        const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, it++)
        {
            const uint32_t ph = it & 1;
            const int node = tile * TM + warp * 32 + lane;
            mbar_wait(&bar[BAR_G1_DONE], ph);
            tc::fence_after_sync();
            // z = relu(acc + b1) -> bf16 hi/lo, in place: columns [16c, 16c+8) hi, [16c+8, 16c+16) lo of k-step c
#pragma unroll 1
            for (int c = 0; c < N1 / 16; c++)
            {
                uint32_t r[16];
                tc::ld16(lane_base + TC_Z + 16 * c, r);
                tc::wait_ld();
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int j = 0; j < 8; j++)
                {
                    const float2 b = *reinterpret_cast<const float2*>(b1s + 16 * c + 2 * j);
                    split2(relu_f(__uint_as_float(r[2 * j]) + b.x), relu_f(__uint_as_float(r[2 * j + 1]) + b.y), hi[j], lo[j]);
                }
                tc::st8(lane_base + TC_Z + 16 * c, hi);
                tc::st8(lane_base + TC_Z + 16 * c + 8, lo);
            }
            tc::wait_st();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar[BAR_A2_FULL]);

            mbar_wait(&bar[BAR_G2_DONE], ph);
            tc::fence_after_sync();
            float* out = p.h_out + (size_t)node * D;
            const bool live = node < p.num_nodes;
#pragma unroll 1
            for (int c = 0; c < N2 / 16; c++)
            {
                uint32_t r[16];
                tc::ld16(lane_base + TC_H + 16 * c, r);
                tc::wait_ld();
#pragma unroll
                for (int j = 0; j < 4; j++)
                {
                    if (16 * c + 4 * j < D)
                    {
                        const float4 b = ld_f4(b2s + 16 * c + 4 * j);
                        float4 o = make_float4(__uint_as_float(r[4 * j]) + b.x, __uint_as_float(r[4 * j + 1]) + b.y,
                                               __uint_as_float(r[4 * j + 2]) + b.z, __uint_as_float(r[4 * j + 3]) + b.w);
                        if (p.relu_out) o = make_float4(relu_f(o.x), relu_f(o.y), relu_f(o.z), relu_f(o.w));
                        if (live) stg_f4_stream(out + 16 * c + 4 * j, o);
                    }
                }
            }
        }
    }

    tc::fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP) tc::tmem_dealloc(tbase, TMEM_COLS);
}

}  // namespace

int gin_layer_tc_launch(const DeviceBatch& b, const GinWeights& w, int layer, const float* h_in, float* h_out, int sm_count, cudaStream_t s)
{
    static bool attr_set = false;
    if (!attr_set)
    {
        FG_CUDA(cudaFuncSetAttribute(gin_layer_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem::BYTES));
        attr_set = true;
    }
    GinTcParams p;
    p.h_in = h_in; p.h_out = h_out;
    p.in_ptr = b.in_ptr.as<int>(); p.src = b.src.as<int>(); p.code = b.code.as<uint8_t>();
    p.ee_raw = w.ee_raw.as<float>() + (size_t)layer * ED_FEATURE_PER_LAYER * D;
    p.wpack = w.wpack.as<unsigned char>() + (size_t)layer * 4 * WBLOCK;
    p.b1 = w.b1.as<float>() + (size_t)layer * N1;
    p.b2 = w.b2p.as<float>() + (size_t)layer * N2;
    p.num_nodes = (int)b.total_nodes;
    p.num_tiles = (int)ceil_div<long>(b.total_nodes, TM);
    p.relu_out = (layer != 4);
    const int grid = std::min(p.num_tiles, sm_count);
    gin_layer_tc_kernel<<<grid, NT, Smem::BYTES, s>>>(p);
    FG_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace fg

// GIN / GIN-VN layer on the B200 tensor cores: one persistent, warp-specialised kernel per layer.
//
// Reference work per layer (GIN/src/message_passing.cc:77-150, node_embedding.cc:23-201):
//   m_v = sum_{(u,v)} relu(h_u + EE_l[attr_uv]);  a_v = m_v + h_v  (eps is never loaded: SURVEY.md F4)
//   z = relu(W1 a + b1);  h'_v = W2 z + b2  (+ relu unless last layer)
//
// B200 mapping (one CTA per SM, tiles of 128 consecutive nodes, UMMA M = 128, cta_group::1):
//   * W1 / W2 are STATIONARY in shared memory for the whole launch, each as a bf16 hi + bf16 lo pair
//     (no-swizzle K-major canonical layout, 4 x 46,592 B, loaded once by bulk TMA);
//   * activations flow through TENSOR MEMORY.  8 gather warps build a_v in fp32 (16-byte gathers in CSR
//     order -> deterministic, no atomics; the tile's CSR slice is staged in shared memory by a loader warp
//     one tile ahead, feature rows are prefetched into L2 two tiles ahead), split it into bf16 hi/lo and
//     store it straight into TMEM with 16-lane x 256-bit tcgen05.st (the m16n8 fragment layout) as the A
//     operand of GEMM1;
//   * GEMM1 = 3 products (hi*hi + lo*hi + hi*lo) x 7 k-steps, fp32 accumulate in TMEM, issued as two N halves
//     (112 + 96) so that 8 epilogue warps convert the first half while the tensor core works on the second:
//     z = relu(acc + b1) is split to bf16 hi/lo and written back IN PLACE -- the 16 fp32 columns of k-step j
//     become its 8 hi + 8 lo operand columns for GEMM2 (3 x 13 MMAs, N = 112, started on the first half
//     while the second is being converted);
//   * the second epilogue adds b2 (+ relu) and streams h' to HBM, one full 32-byte sector per row and store;
//   * registers are redistributed with setmaxnreg: gather warps grow, epilogue / MMA / loader warps shrink.
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

// warp roles (warpgroup aligned, for setmaxnreg): 0-7 epilogue, 8-15 gather, 16 MMA issuer, 17 loader, 18-19 idle
#ifndef FG_EPI_WARPS
#define FG_EPI_WARPS 8
#endif
constexpr int EPI_WARPS = FG_EPI_WARPS, GATHER_WARPS = 8;
constexpr int MMA_WARP = EPI_WARPS + GATHER_WARPS;
constexpr int LOAD_WARP = MMA_WARP + 1;
constexpr int NT = (MMA_WARP + 4) * 32;       // 512 / 640 threads
constexpr bool REBALANCE_REGS = EPI_WARPS == 8;                   // 640 threads: launch allocation 96 per thread
constexpr int REGS_EPI = 80, REGS_MISC = 72, REGS_GATHER = 120;
constexpr int N1A = 112, N1B = N1 - N1A;      // GEMM1 N halves = GEMM2 K halves (7 + 6 k-steps)

// tensor-memory columns
constexpr uint32_t TC_A1_HI = 0, TC_A1_LO = 56, TC_Z = 128, TC_H = 384;
constexpr uint32_t TMEM_COLS = 512;

// per-tile slice of the destination CSR staged in shared memory by the loader warp (double buffered)
constexpr int CSR_CAP = 1024;                 // in-edges of one tile; larger tiles read src/code from global
struct CsrBuf {
    int ptr[TM + 4];                          // absolute edge positions of the tile's rows (+1)
    int src[CSR_CAP];
    uint8_t code[CSR_CAP];
    int e0; int staged; int pad[2];
};

struct Smem {
    static constexpr int W = 0;                                   // W1_hi | W1_lo | W2_hi | W2_lo
    static constexpr int EE = W + 4 * WBLOCK;                     // [60][100] fp32 combined edge-embedding rows
    static constexpr int B1 = EE + (ED_COMBOS + 1) * D * 4;       // [208]   (EE row 60 = -3e38 sentinel for absent edge slots)
    static constexpr int B2 = B1 + N1 * 4;                        // [112]
    static constexpr int CSR = B2 + N2 * 4;                       // 2 x CsrBuf
    static constexpr int BAR = CSR + 2 * (int)sizeof(CsrBuf);
    static constexpr int TMEM_PTR = BAR + 16 * 8;
    static constexpr int BYTES = TMEM_PTR + 16;
};
static_assert(Smem::CSR % 16 == 0 && Smem::BAR % 8 == 0 && sizeof(CsrBuf) % 16 == 0, "alignment");
static_assert(Smem::BYTES <= 232448, "shared memory budget");

enum { BAR_W = 0, BAR_A1_FULL, BAR_G1A_DONE, BAR_G1B_DONE, BAR_A2A_FULL, BAR_A2B_FULL, BAR_G2_DONE, BAR_CSR_FULL /* 2 */, BAR_CSR_EMPTY = BAR_CSR_FULL + 2 /* 2 */ };

struct GinTcParams {
    const float* h_in; float* h_out;
    const int* in_ptr; const int* src; const uint8_t* code;
    const float* ee_comb;            // [60][100] this layer: ((0 + T[a0]) + T[5 + a1]) + T[11 + a2]
    const unsigned char* wpack;      // 4 x WBLOCK this layer
    const float* b1; const float* b2;   // [208], [112] zero padded
    int num_nodes; int num_tiles; int relu_out;
};

// relu that lets NaN through, like the reference's compare-select (GIN/src/util.h:20-25), in ONE instruction
__device__ __forceinline__ float relu_nan(float x)
{
    float y;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(y) : "f"(x), "f"(0.0f));
    return y;
}

template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// (x0, x1) -> packed bf16 pairs: hi = rn(x), lo = rn(x - hi); element 0 in the low half
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo)
{
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    const float r0 = x0 - __uint_as_float(hi << 16);
    const float r1 = x1 - __uint_as_float(hi & 0xFFFF0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}

// 16 lanes x 256 bit TMEM store: thread t supplies columns 2(t%4), 2(t%4)+1 of lane t/4 (r0, r1) and of lane t/4 + 8 (r2, r3)
// -- the layout of an m16n8 accumulator fragment (verified on hardware with tools/tc_probe2.cu)
__device__ __forceinline__ void st_16x256(uint32_t taddr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3)
{
    asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}

// The in-edges of one tile row, as far as they fit in registers (molecular graphs: in-degree <= 4 almost always);
// longer lists continue from edge 4 in the staged CSR / global memory.  Absent slots are never loaded: they
// read as h_u = 0 and point at the sentinel table row (-3e38): relu(-3e38 + 0) adds exactly 0.
struct RowEdges {
    const float* hv;         // own feature row, offset by the thread's float4 sub-chunk
    const float* hu[4];      // source rows of the first four in-edges (own row where absent)
    uint32_t t[4];           // shared-memory byte address of their edge-embedding rows (sentinel where absent)
    int deg; int eb;
};

__device__ __forceinline__ RowEdges load_row_edges(const GinTcParams& p, const float* ee, const CsrBuf& cb, int n0, int rows, int r, int qsub)
{
    RowEdges re;
    const bool live = r < rows;
    const int node = n0 + (live ? r : rows - 1);
    re.eb = live ? cb.ptr[r] : 0;
    re.deg = live ? cb.ptr[r + 1] - re.eb : 0;
    re.hv = p.h_in + (size_t)node * D + 4 * qsub;
    const uint32_t ee_addr = smem_u32(ee) + 16 * qsub;
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
        int u = node, c = ED_COMBOS;
        if (j < re.deg)
        {
            if (cb.staged) { u = cb.src[re.eb + j - cb.e0]; c = cb.code[re.eb + j - cb.e0]; }
            else { u = __ldg(p.src + re.eb + j); c = __ldg(p.code + re.eb + j); }
        }
        re.hu[j] = p.h_in + (size_t)u * D + 4 * qsub;
        re.t[j] = ee_addr + c * (D * 4);
    }
    return re;
}

// predicated 16-byte load (no branch, no memory traffic when `on` is false): absent edge slots read as h_u = 0
__device__ __forceinline__ float4 ldg_f4_if(const float* ptr, bool on)
{
    float4 v;
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "mov.f32 %0, 0f00000000;\n\tmov.f32 %1, 0f00000000;\n\tmov.f32 %2, 0f00000000;\n\tmov.f32 %3, 0f00000000;\n\t"
        "@p ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];\n\t}"
        : "=&f"(v.x), "=&f"(v.y), "=&f"(v.z), "=&f"(v.w)
        : "l"(ptr), "r"((int)on));
    return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// Software-pipelined gather of one tile row: `row_loads<KS>` issues the 16-byte loads of chunk q = 4 KS + qsub
// (own row + up to four source rows, addresses = pointer + immediate), `row_finish<KS>` consumes them:
// a_v[4q .. 4q+3] = sum over in-edges (CSR order) relu(h_u + EE[attr]) + h_v.  The loads of the NEXT half-step
// are always issued before the current one is consumed, so every thread keeps five loads in flight.
struct RowLoads { float4 hv; float4 hu[4]; };

template <int KS>
__device__ __forceinline__ void row_loads(const RowEdges& r, int qsub, RowLoads& L)
{
    constexpr int OFF = 16 * KS;      // floats
    const bool on = (KS < 6) || (qsub == 0);      // the last k-step only holds chunk 24 (sub-chunk 0); the rest is zero padding
    L.hv = ldg_f4_if(r.hv + OFF, on);
#pragma unroll
    for (int j = 0; j < 4; j++) L.hu[j] = ldg_f4_if(r.hu[j] + OFF, on && j < r.deg);
}

template <int KS>
__device__ __forceinline__ float4 row_finish(const GinTcParams& p, const float* ee, const CsrBuf& cb, const RowEdges& r, int qsub, bool live,
                                             const RowLoads& L)
{
    constexpr int OFF = 16 * KS;
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
        const float4 t = lds_f4(r.t[j] + 4 * OFF);
        m.x += relu_nan(t.x + L.hu[j].x); m.y += relu_nan(t.y + L.hu[j].y); m.z += relu_nan(t.z + L.hu[j].z); m.w += relu_nan(t.w + L.hu[j].w);
    }
    if (r.deg > 4)
    {
        // long in-edge lists (virtual nodes, kNN graphs): continue in rounds of four edges, loads first
        const int q = 4 * KS + qsub;
        if (q < Q)
        {
            const float* hq = p.h_in + 4 * q;
            const int e_end = r.eb + r.deg;
            for (int e = r.eb + 4; e < e_end; e += 4)
            {
                float4 hu[4];
                int c[4];
#pragma unroll
                for (int j = 0; j < 4; j++)
                {
                    const bool on = e + j < e_end;
                    int u = 0;
                    c[j] = ED_COMBOS;
                    if (on)
                    {
                        if (cb.staged) { u = cb.src[e + j - cb.e0]; c[j] = cb.code[e + j - cb.e0]; }
                        else { u = __ldg(p.src + e + j); c[j] = __ldg(p.code + e + j); }
                    }
                    hu[j] = ldg_f4_if(hq + (size_t)u * D, on);
                }
#pragma unroll
                for (int j = 0; j < 4; j++)
                {
                    const float4 t = ld_f4(ee + c[j] * D + 4 * q);
                    m.x += relu_nan(t.x + hu[j].x); m.y += relu_nan(t.y + hu[j].y); m.z += relu_nan(t.z + hu[j].z); m.w += relu_nan(t.w + hu[j].w);
                }
            }
        }
    }
    float4 a = make_float4(m.x + L.hv.x, m.y + L.hv.y, m.z + L.hv.z, m.w + L.hv.w);
    if (!live || (KS == 6 && qsub != 0)) a = make_float4(0.f, 0.f, 0.f, 0.f);
    return a;
}

// split one k-step of both rows into bf16 hi/lo and store it to TMEM (16 lanes x 256 bit)
template <int KS>
__device__ __forceinline__ void store_step(const float4& a, const float4& b, uint32_t taddr, uint64_t* a1_free, int it)
{
    uint32_t ha0, la0, ha1, la1, hb0, lb0, hb1, lb1;
    split2(a.x, a.y, ha0, la0);
    split2(a.z, a.w, ha1, la1);
    split2(b.x, b.y, hb0, lb0);
    split2(b.z, b.w, hb1, lb1);
    if (KS == 0 && it > 0)
    {
        // A1 of the previous tile has been consumed once GEMM1 of that tile completed
        mbar_wait(a1_free, (it - 1) & 1);
        tc::fence_after_sync();
    }
    __syncwarp();
    st_16x256(taddr + TC_A1_HI + 8 * KS, ha0, ha1, hb0, hb1);
    st_16x256(taddr + TC_A1_LO + 8 * KS, la0, la1, lb0, lb1);
}

template <int KS>
__device__ __forceinline__ void gather_steps(const GinTcParams& p, const float* ee, const CsrBuf& cb, const RowEdges& ra, const RowEdges& rb, int qsub,
                                             bool live_a, bool live_b, uint32_t taddr, uint64_t* a1_free, int it, RowLoads& La, RowLoads& Lb)
{
    // on entry the loads of (KS, row a) are in flight in La
    row_loads<KS>(rb, qsub, Lb);
    const float4 a = row_finish<KS>(p, ee, cb, ra, qsub, live_a, La);
    if (KS < 6) row_loads<(KS < 6 ? KS + 1 : KS)>(ra, qsub, La);
    const float4 b = row_finish<KS>(p, ee, cb, rb, qsub, live_b, Lb);
    store_step<KS>(a, b, taddr, a1_free, it);
    if constexpr (KS < 6) gather_steps<KS + 1>(p, ee, cb, ra, rb, qsub, live_a, live_b, taddr, a1_free, it, La, Lb);
}

// z = relu(acc + b1) for 16 accumulator columns of this thread's row -> bf16 hi/lo, written back in place
__device__ __forceinline__ void convert_chunk(uint32_t zaddr, const float* b1c)
{
    uint32_t r[16];
    tc::ld16(zaddr, r);
    tc::wait_ld();
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
        const float4 b = ld_f4(b1c + 4 * j);
        split2(relu_nan(__uint_as_float(r[4 * j]) + b.x), relu_nan(__uint_as_float(r[4 * j + 1]) + b.y), hi[2 * j], lo[2 * j]);
        split2(relu_nan(__uint_as_float(r[4 * j + 2]) + b.z), relu_nan(__uint_as_float(r[4 * j + 3]) + b.w), hi[2 * j + 1], lo[2 * j + 1]);
    }
    tc::st8(zaddr, hi);
    tc::st8(zaddr + 8, lo);
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
        mbar_init(&bar[BAR_G1A_DONE], 1);
        mbar_init(&bar[BAR_G1B_DONE], 1);
        mbar_init(&bar[BAR_A2A_FULL], EPI_WARPS);
        mbar_init(&bar[BAR_A2B_FULL], EPI_WARPS);
        mbar_init(&bar[BAR_G2_DONE], 1);
        for (int i = 0; i < 2; i++)
        {
            mbar_init(&bar[BAR_CSR_FULL + i], 1);
            mbar_init(&bar[BAR_CSR_EMPTY + i], GATHER_WARPS);
        }
        fence_mbar_init();
    }
    if (warp == MMA_WARP)
    {
        tc::tmem_alloc(tmem_ptr, TMEM_COLS);
        tc::tmem_relinquish();
    }
    for (int i = tid; i < ED_COMBOS * Q; i += NT) st_f4(ee + 4 * i, ldg_f4(p.ee_comb + 4 * i));
    for (int i = tid; i < D; i += NT) ee[ED_COMBOS * D + i] = -3.0e38f;
    for (int i = tid; i < N1; i += NT) b1s[i] = __ldg(p.b1 + i);
    for (int i = tid; i < N2; i += NT) b2s[i] = __ldg(p.b2 + i);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = *tmem_ptr;

    if (warp >= MMA_WARP)
    {
        if (REBALANCE_REGS) reg_dec<REGS_MISC>();
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
                const uint32_t idesc1a = tc::idesc_bf16(TM, N1A), idesc1b = tc::idesc_bf16(TM, N1B), idesc2 = tc::idesc_bf16(TM, N2);
                int it = 0;
                for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, it++)
                {
                    const uint32_t ph = it & 1;
                    mbar_wait(&bar[BAR_A1_FULL], ph);
                    tc::fence_after_sync();
                    // GEMM1, N half a (z columns 0..111) then half b (112..207): B rows are 16 bytes apart inside a k-chunk
#pragma unroll
                    for (int nh = 0; nh < 2; nh++)
                    {
                        bool acc = false;
#pragma unroll
                        for (int prod = 0; prod < 3; prod++)
                        {
                            const uint32_t a_col = tbase + (prod == 1 ? TC_A1_LO : TC_A1_HI);
                            const uint32_t b_addr = w_addr + (prod == 2 ? WBLOCK : 0) + (nh ? N1A * 16 : 0);
#pragma unroll
                            for (int j = 0; j < K1 / 16; j++)
                            {
                                tc::mma_ts(tbase + TC_Z + (nh ? N1A : 0), a_col + 8 * j, tc::smem_desc(b_addr + 2 * j * N1 * 16, N1 * 16, 128),
                                           nh ? idesc1b : idesc1a, acc);
                                acc = true;
                            }
                        }
                        tc::commit(&bar[nh ? BAR_G1B_DONE : BAR_G1A_DONE]);
                    }
                    // GEMM2, K half a (k-steps 0..6, operand columns converted from z half a) then half b (7..12)
                    bool acc = false;
#pragma unroll
                    for (int kh = 0; kh < 2; kh++)
                    {
                        mbar_wait(&bar[kh ? BAR_A2B_FULL : BAR_A2A_FULL], ph);
                        tc::fence_after_sync();
#pragma unroll
                        for (int prod = 0; prod < 3; prod++)
                        {
                            const uint32_t a_col = tbase + TC_Z + (prod == 1 ? 8 : 0);
                            const uint32_t b_addr = w_addr + 2 * WBLOCK + (prod == 2 ? WBLOCK : 0);
#pragma unroll
                            for (int j = (kh ? N1A / 16 : 0); j < (kh ? K2 / 16 : N1A / 16); j++)
                            {
                                tc::mma_ts(tbase + TC_H, a_col + 16 * j, tc::smem_desc(b_addr + 2 * j * N2 * 16, N2 * 16, 128), idesc2, acc);
                                acc = true;
                            }
                        }
                    }
                    tc::commit(&bar[BAR_G2_DONE]);
                }
            }
        }
        else if (warp == LOAD_WARP)
        {
            // ===== loader warp: stages the next tile's CSR slice in shared memory and prefetches feature rows into L2 =====
            CsrBuf* csr = reinterpret_cast<CsrBuf*>(smem + Smem::CSR);
            int it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, it++)
            {
                CsrBuf& cb = csr[it & 1];
                if (it >= 2) mbar_wait(&bar[BAR_CSR_EMPTY + (it & 1)], ((it >> 1) - 1) & 1);
                const int n0 = tile * TM;
                const int rows = min(TM, p.num_nodes - n0);
                {
                    // feature rows of the tile after next -> L2, 4 rows (1,600 B) per lane
                    const int ahead = tile + 2 * gridDim.x;
                    const int rn = ahead < p.num_tiles ? min(TM, p.num_nodes - ahead * TM) : 0;
                    const int r0 = 4 * lane, nr = min(4, rn - r0);
                    if (nr > 0)
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.h_in + ((size_t)ahead * TM + r0) * D), "r"(nr * D * 4) : "memory");
                }
                for (int i = lane; i <= rows; i += 32) cb.ptr[i] = __ldg(p.in_ptr + n0 + i);
                __syncwarp();
                const int e0 = cb.ptr[0], ne = cb.ptr[rows] - e0;
                const bool staged = ne <= CSR_CAP;
                if (staged)
                {
                    for (int i = lane; i < ne; i += 32)
                    {
                        cb.src[i] = __ldg(p.src + e0 + i);
                        cb.code[i] = __ldg(p.code + e0 + i);
                    }
                }
                if (lane == 0) { cb.e0 = e0; cb.staged = staged ? 1 : 0; }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar[BAR_CSR_FULL + (it & 1)]);
            }
        }
    }
    else if (warp >= EPI_WARPS)
    {
        if (REBALANCE_REGS) reg_inc<REGS_GATHER>();
        // ===== gather warps: build the A operand of GEMM1 directly in tensor memory =====
        // warp -> TMEM lane quadrant (warp % 4) and 16-row half; thread t -> rows t/4 and t/4 + 8 of that half and,
        // per k-step ks, the float4 chunk q = 4 ks + t % 4 (k = 16 ks + 4 (t%4) .. +3 = TMEM columns 2(t%4), 2(t%4)+1)
        const int gw = warp - EPI_WARPS;          // 0..7
        const int quad = warp & 3, half = gw >> 2;
        const uint32_t taddr = tbase + ((uint32_t)(quad * 32 + half * 16) << 16);
        const int r_a = quad * 32 + half * 16 + (lane >> 2), r_b = r_a + 8;
        const int qsub = lane & 3;
        const CsrBuf* csr = reinterpret_cast<const CsrBuf*>(smem + Smem::CSR);
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, it++)
        {
            const int n0 = tile * TM;
            const int rows = min(TM, p.num_nodes - n0);
            const CsrBuf& cb = csr[it & 1];
            mbar_wait(&bar[BAR_CSR_FULL + (it & 1)], (it >> 1) & 1);
            const RowEdges ra = load_row_edges(p, ee, cb, n0, rows, r_a, qsub), rb = load_row_edges(p, ee, cb, n0, rows, r_b, qsub);
            const bool live_a = r_a < rows, live_b = r_b < rows;
            RowLoads La, Lb;
            row_loads<0>(ra, qsub, La);
            gather_steps<0>(p, ee, cb, ra, rb, qsub, live_a, live_b, taddr, &bar[BAR_G1B_DONE], it, La, Lb);
            tc::wait_st();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0)
            {
                mbar_arrive(&bar[BAR_A1_FULL]);
                mbar_arrive(&bar[BAR_CSR_EMPTY + (it & 1)]);
            }
        }
    }
    else
    {
        if (REBALANCE_REGS) reg_dec<REGS_EPI>();
        // ===== epilogue warps: two per TMEM lane quadrant =====
        constexpr int PER_QUAD = EPI_WARPS / 4;      // warps per TMEM lane quadrant
        const int quad = warp & 3, pp = warp >> 2;
        const uint32_t lane_base = tbase + ((uint32_t)(quad * 32) << 16);
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, it++)
        {
            const uint32_t ph = it & 1;
            // z = relu(acc + b1) -> bf16 hi/lo, in place (thread = row): columns [16c, 16c+8) hi, [16c+8, 16c+16) lo of
            // k-step c; the two warps of a quadrant take alternate chunks
            mbar_wait(&bar[BAR_G1A_DONE], ph);
            tc::fence_after_sync();
#pragma unroll 1
            for (int c = pp; c < N1A / 16; c += PER_QUAD) convert_chunk(lane_base + TC_Z + 16 * c, b1s + 16 * c);
            tc::wait_st();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar[BAR_A2A_FULL]);

            mbar_wait(&bar[BAR_G1B_DONE], ph);
            tc::fence_after_sync();
#pragma unroll 1
            for (int c = N1A / 16 + (PER_QUAD == 2 ? (pp ^ 1) : 0); c < N1 / 16; c += PER_QUAD) convert_chunk(lane_base + TC_Z + 16 * c, b1s + 16 * c);
            tc::wait_st();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar[BAR_A2B_FULL]);

            mbar_wait(&bar[BAR_G2_DONE], ph);
            tc::fence_after_sync();
            // h' = acc + b2 (+ relu): 16-lane x 256-bit TMEM loads give thread t columns 8g + 2(t%4), +1 of rows t/4 and
            // t/4 + 8, so the four lanes of a row write one full 32-byte sector per store instruction; warp pp of the
            // quadrant takes its 16-row half
#pragma unroll
            for (int hh = pp; hh < 2; hh += PER_QUAD)
            {
            const int row_a = tile * TM + quad * 32 + hh * 16 + (lane >> 2), row_b = row_a + 8;
            const uint32_t ta = lane_base + ((uint32_t)(hh * 16) << 16) + TC_H;
#pragma unroll
            for (int g4 = 0; g4 < 13; g4 += 4)
            {
                uint32_t r[16];
                asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                               "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                             : "r"(ta + 8 * g4)
                             : "memory");
                tc::wait_ld();
#pragma unroll
                for (int g = 0; g < 4; g++)
                {
                    const int col = 8 * (g4 + g) + 2 * (lane & 3);
                    if (8 * (g4 + g) < D && col < D)
                    {
                        const float2 bb = *reinterpret_cast<const float2*>(b2s + col);
                        float2 oa = make_float2(__uint_as_float(r[4 * g]) + bb.x, __uint_as_float(r[4 * g + 1]) + bb.y);
                        float2 ob = make_float2(__uint_as_float(r[4 * g + 2]) + bb.x, __uint_as_float(r[4 * g + 3]) + bb.y);
                        if (p.relu_out)
                        {
                            oa = make_float2(relu_nan(oa.x), relu_nan(oa.y));
                            ob = make_float2(relu_nan(ob.x), relu_nan(ob.y));
                        }
                        if (row_a < p.num_nodes) *reinterpret_cast<float2*>(p.h_out + (size_t)row_a * D + col) = oa;
                        if (row_b < p.num_nodes) *reinterpret_cast<float2*>(p.h_out + (size_t)row_b * D + col) = ob;
                    }
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
    FG_TRY(opt_in_smem(reinterpret_cast<const void*>(&gin_layer_tc_kernel), Smem::BYTES));
    GinTcParams p;
    p.h_in = h_in; p.h_out = h_out;
    p.in_ptr = b.in_ptr.as<int>(); p.src = b.src.as<int>(); p.code = b.code.as<uint8_t>();
    p.ee_comb = w.ee_comb.as<float>() + (size_t)layer * ED_COMBOS * D;
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

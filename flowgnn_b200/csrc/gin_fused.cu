// GIN / GIN-VN layer, ONE kernel per layer: TMA-staged graph-aligned tiles, shared-memory edge gather, tcgen05 node MLP.
//
// Reference work per layer (GIN/src/message_passing.cc:77-150, node_embedding.cc:23-201):
//   m_v = sum_{(u,v)} relu(h_u + EE_l[attr_uv]);  a_v = m_v + h_v  (eps is never loaded: SURVEY.md F4)
//   z = relu(W1 a + b1);  h'_v = W2 z + b2  (+ relu unless last layer)
//
// What changed against gin_tc2.cu (same CTA pair, same weight image, same MMA / epilogue protocol): the edge gather no
// longer goes through the L1 / LSU global path.  In-edge sources are nodes of the SAME graph and a graph's rows are
// contiguous in HBM, so prep.cu packs whole graphs into tiles of <= 128 rows and the kernel lands a tile's feature rows
// (and its row descriptors) in shared memory with bulk-TMA copies (cp.async.bulk + mbarrier, issued by a producer warp:
// no LSU wavefronts, every row read from HBM exactly once).  The gather warps read own row, source rows and
// edge-embedding rows with conflict-free 128-byte shared-memory pieces, reduce in CSR order and keep the tile's x = m + h
// as packed bf16 hi/lo IN REGISTERS (24 per thread); once all 16 gather warps are done reading, the A tile of the GEMM
// is written IN PLACE over the stage.  Two such buffers alternate: while GEMM1 consumes the A tile of tile t in one, the
// gather reads the stage of tile t+1 in the other, and the buffer GEMM1 releases receives the rows of tile t+2.
//
// Per CTA (896 threads; cluster of two CTAs = one UMMA M = 256 tile = two graph-aligned tiles):
//   warps 0-7   epilogue (as gin_tc2.cu): z = relu(acc) -> bf16 hi/lo in tensor memory, h' = acc (+relu) -> HBM, or the fused head
//   warps 8-23  gather: 8 rows each as two passes of 4 rows, 8 lanes per row, chunk 8 ks + j in step ks < 3; chunk 24 by lanes 0..7
//   warp 24     (leader CTA) MMA issuer: GEMM1 SS (3 products x 7 k-steps, N halves 112 + 96), GEMM2 TS (3 x 13, N = 128)
//   warp 25     producer: bulk-TMA copies of a tile's rows + descriptors into the free buffer, L2 prefetch two tiles ahead
// Option mp_only (node transform = identity): the same gather writes x = m + h straight to h_out, the MMA
// and epilogue warps idle.
// Tiles whose graph exceeds 128 nodes are flagged external: their source rows are read from global memory.
#include "internal.cuh"
#include "layers.cuh"
#include "tc.cuh"
#include "pair.cuh"
#include "gin_wpack.cuh"

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <set>
#include <type_traits>

namespace fg {

namespace {

using namespace ginw;
using namespace pair;

constexpr int LBO_A = TM * 16 + 32;                     // +32: the 8-byte stores of a warp hit every bank group exactly twice
constexpr int A_BYTES = K1_CHUNKS * LBO_A;              // one hi or lo buffer
constexpr int ZERO_BYTES = TM * 16;
constexpr int ROW_BYTES = D * 4;                        // 400
constexpr int STAGE_BYTES = TM * ROW_BYTES;             // 51,200
constexpr int DESC_BYTES = TM * 16;                     // 2,048

constexpr int EPI_WARPS = 8, GATHER_WARPS = 16;
constexpr int MMA_WARP = EPI_WARPS + GATHER_WARPS;
// 28 warps (7 per scheduler): launched with 72 registers per thread, setmaxnreg moves them to where they are needed -- the
// gather warps keep a tile's x in registers (80), the epilogue stays at 72, the last warpgroup (the MMA issuer and three
// idle warps) gets 40: every operand descriptor of the issuer is a compile-time constant (the tile loop is unrolled by
// two, one body per buffer), so it holds next to nothing.  (Tried: 24 registers with the descriptors in a shared-memory /
// constant-memory table -- ptxas moves them through vector registers and local memory, GEMM1 then takes 2.2 us to issue
// instead of ~1; 25 warps x 80 registers without setmaxnreg does not launch: 7 warps x 2,560 registers on one scheduler
// exceed its 16,384.)  The TMA producer is thread 0 of the epilogue: it learns that GEMM1(t) has consumed a buffer from
// the barrier it waits on anyway.
constexpr int NT = (MMA_WARP + 4) * 32;       // 896
constexpr int REGS_LAUNCH = 72;
constexpr int REGS_EPI = 72, REGS_MISC = 40, REGS_GATHER = 80;
static_assert(32 * (EPI_WARPS * REGS_EPI + GATHER_WARPS * REGS_GATHER + 4 * REGS_MISC) <= NT * REGS_LAUNCH, "setmaxnreg pool");
constexpr int ROWS_PER_WARP = TM / GATHER_WARPS;   // 8

// tensor-memory columns: a ring of three 112-column z halves (tile t: GEMM1 half a -> ring slot 2t % 3, half b -> (2t + 1) % 3;
// 96 of the 112 columns used by half b) and the h' accumulator.  With three halves GEMM1 of tile t+1 can run while the
// epilogue still converts and GEMM2 still reads the z of tile t.
constexpr uint32_t TC_ZR = 0, ZR_COLS = N1A, TC_H = 3 * ZR_COLS;
static_assert(TC_H + N2 <= 512, "tensor memory columns");
constexpr uint32_t TMEM_COLS = 512;

constexpr int G1_MMAS = 2 * 3 * K1_STEPS, G2_MMAS = 3 * K2_STEPS;       // 42, 39
constexpr int BUF_BYTES = 2 * A_BYTES;                  // one buffer: stage rows [128][100] fp32 + descriptors [128] int4, later A hi | A lo
struct Smem {
    static constexpr int W = 0;
    static constexpr int BUF = W + W_BYTES;                         // [2][BUF_BYTES]
    static constexpr int ZERO = BUF + 2 * BUF_BYTES;                // K chunk 13 of the A tiles (must lie above them: LBO >= 0); rows beyond a tile
    static constexpr int EE = ZERO + ZERO_BYTES;                    // [61][100] fp32 combined edge-embedding rows; row 60 = sentinel
    static constexpr int BAR = EE + (ED_COMBOS + 1) * D * 4;
    static constexpr int TMEM_PTR = BAR + 16 * 8;
    static constexpr int TILE = TMEM_PTR + 16;                      // [2] int2 tile record of each buffer (written by the producer)
    static constexpr int BYTES = TILE + 16;
};
static_assert(Smem::ZERO % 16 == 0 && Smem::BUF % 16 == 0 && BUF_BYTES % 16 == 0 && Smem::EE % 16 == 0 && Smem::BAR % 8 == 0, "alignment");
static_assert(BUF_BYTES >= STAGE_BYTES + DESC_BYTES, "a buffer holds the stage rows and the descriptors");
static_assert(Smem::BYTES <= 232448, "shared memory budget");

enum { BAR_W = 0, BAR_A_FULL /* 2 */, BAR_G1A_DONE = BAR_A_FULL + 2 /* 2 */, BAR_G1B_DONE = BAR_G1A_DONE + 2 /* 2 */, BAR_A2A_FULL = BAR_G1B_DONE + 2, BAR_A2B_FULL, BAR_G2_DONE,
       BAR_STAGE_FULL /* 2 */, BAR_BUF_FREE = BAR_STAGE_FULL + 2 /* 2 */, BAR_COUNT = BAR_BUF_FREE + 2 };
static_assert(BAR_COUNT <= 16, "barrier slots");

struct GinFusedParams {
    const float* h_in; float* h_out;
    const int* in_ptr; const int* src; const uint8_t* code;
    const int4* row_desc;            // [N] per tile: the rows' descriptors (first four in-edges, packed) ordered by in-degree, the row's
                                     // position inside the tile in .y bits 24..30 (prep.cu::sort_tile_rows_kernel); nullptr: no in-edges at all
    const int2* tiles;               // graph-aligned tiles (first node, rows | external << 30) (prep.cu::pack_tiles_kernel)
    const int* tile_count;
    const float* ee_comb;            // [60][100] this layer
    const unsigned char* wpack;      // [2 ranks][W_BYTES] this layer
    int num_nodes; int relu_out; int mp_only;
    const float* head_w; float* node_dot;      // last layer: fused prediction head (see gin_tc2.cu)
    unsigned long long* trace;                 // -DFG_TC2_TRACE only: timeline of pair 0 (tools/trace_fused.py): [role][tile][event] globaltimer ns
};

#ifdef FG_TC2_TRACE
__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define TRACE(role, it, ev) do { if (p.trace && pair == 0 && rank == 0 && (it) < 64) p.trace[((role) * 64 + (it)) * 8 + (ev)] = gtime(); } while (0)
// whole-kernel marks of CTA 0 (slot 60) and of the last CTA (slot 61): entry, prologue done, loops done
#define TRACE_K(ev) do { if (p.trace && tid == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) p.trace[(1 * 64 + (blockIdx.x == 0 ? 0 : 2) + ((ev) >> 1)) * 8 + 6 + ((ev) & 1)] = gtime(); } while (0)
#else
#define TRACE_K(ev) do { } while (0)
#define TRACE(role, it, ev) do { } while (0)
#endif

__device__ __forceinline__ int2 tile_of(const GinFusedParams& p, int t, int ntiles)
{
    int2 v = make_int2(0, 0);
    if (t >= ntiles) return v;
    // no tile list: the node MLP alone (second launch of a dense-graph layer) needs no graph alignment -- plain 128-row tiles, all full
    if (!p.tiles) return make_int2(t * TM, min(TM, p.num_nodes - t * TM));
    asm volatile("ld.global.nc.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p.tiles + t));
    return v;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ int4 lds_i4(uint32_t addr)
{
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// Shared-memory address of dynamic shared memory in a kernel without static shared memory (the kernel checks it): with it
// every operand descriptor of the MMA issuer is a compile-time constant.
constexpr uint32_t SMEM_BASE = 0x400;
__host__ __device__ constexpr uint64_t desc_imm(uint32_t off, uint32_t lbo)
{
    return (uint64_t)(((SMEM_BASE + off) >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
}

// compile-time loop: f(std::integral_constant<int, I>) for I = 0 .. N-1
template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F&& f)
{
    if constexpr (I < N)
    {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

__device__ __forceinline__ void stg_f4(float* ptr, const float4& v)
{
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// One destination row as seen by one of its 8 threads: shared-memory byte addresses of this thread's 16-byte chunk
// (step 0) of the row itself, of the source rows of its first four in-edges and of their edge-embedding rows.
// Absent slots point at the row itself and the sentinel table row (relu(h - 3e38) adds exactly 0); a row beyond the end of
// the tile reads the shared zero block, so its x is 0 without a select.
struct RowCtx {
    uint32_t own;
    uint32_t so[4];
    uint32_t tb[4];
    int deg;
    int node;            // global node id
};

// m += relu(t + h) with the two additions as FADD2 (same IEEE result per element as four FADD)
__device__ __forceinline__ void acc_edge2(float4& m, const float4& t, const float4& h)
{
    const float2 a = add2(make_float2(t.x, t.y), make_float2(h.x, h.y)), b = add2(make_float2(t.z, t.w), make_float2(h.z, h.w));
    const float2 ma = add2(make_float2(m.x, m.y), make_float2(relu_nan(a.x), relu_nan(a.y)));
    const float2 mb = add2(make_float2(m.z, m.w), make_float2(relu_nan(b.x), relu_nan(b.y)));
    m = make_float4(ma.x, ma.y, mb.x, mb.y);
}

// x = h_v + sum over in-edges (CSR order) of relu(h_u + EE[code]) for this thread's chunk at byte offset OFF of the row.
// NQ = number of descriptor slots the four rows of this warp instruction need (the longest of their in-edge lists, at
// most 4): the loads are unconditional -- an absent slot costs a shared-memory wavefront but no instruction to predicate
// it, and the kernel is bound by instruction issue, not by the shared-memory pipe.  LONG: a row has more than four
// in-edges (virtual nodes, kNN graphs): rounds of four further edges from the CSR arrays.
template <int NQ, bool LONG, int OFF>
__device__ __forceinline__ float4 gather_chunk(const GinFusedParams& p, const RowCtx& r, int maxdeg, uint32_t stage_thr, uint32_t ee_thr, int tile_start)
{
    const float4 hv = lds_f4(r.own + OFF);
    float4 hu[2], tt[2];
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
    // two slots per batch of loads: 20 registers of loads in flight instead of 36 (the 24 packed results of the tile live in
    // registers too), and the other 15 gather warps cover the second round trip to shared memory
#pragma unroll
    for (int q0 = 0; q0 < NQ; q0 += 2)
    {
#pragma unroll
        for (int q = q0; q < q0 + 2 && q < NQ; q++) { hu[q - q0] = lds_f4(r.so[q] + OFF); tt[q - q0] = lds_f4(r.tb[q] + OFF); }
#pragma unroll
        for (int q = q0; q < q0 + 2 && q < NQ; q++) acc_edge2(m, tt[q - q0], hu[q - q0]);
    }
    if constexpr (LONG)
    {
        const int eb = __ldg(p.in_ptr + r.node);
        int u[4], c[4];
#pragma unroll
        for (int q = 0; q < 4; q++)
        {
            const bool ok = 4 + q < r.deg;
            u[q] = ok ? __ldg(p.src + eb + 4 + q) : r.node;
            c[q] = ok ? (int)__ldg(p.code + eb + 4 + q) : ED_COMBOS;
        }
#pragma unroll 1
        for (int e4 = 4; e4 < maxdeg; e4 += 4)
        {
#pragma unroll
            for (int q0 = 0; q0 < 4; q0 += 2)
            {
#pragma unroll
                for (int q = q0; q < q0 + 2; q++)
                {
                    // an exhausted slot reads the row itself (u = node) with the sentinel table row: adds exactly 0
                    hu[q - q0] = lds_f4(stage_thr + (uint32_t)(u[q] - tile_start) * ROW_BYTES + OFF);
                    tt[q - q0] = lds_f4(ee_thr + c[q] * ROW_BYTES + OFF);
                    const bool ok = e4 + 4 + q < r.deg;
                    u[q] = ok ? __ldg(p.src + eb + e4 + 4 + q) : r.node;
                    c[q] = ok ? (int)__ldg(p.code + eb + e4 + 4 + q) : ED_COMBOS;
                }
#pragma unroll
                for (int q = q0; q < q0 + 2; q++) acc_edge2(m, tt[q - q0], hu[q - q0]);
            }
        }
    }
    const float2 xa = add2(make_float2(m.x, m.y), make_float2(hv.x, hv.y)), xb = add2(make_float2(m.z, m.w), make_float2(hv.z, hv.w));
    return make_float4(xa.x, xa.y, xb.x, xb.y);
}

// One chunk of x: mp_only -> stored to h_out right away; otherwise split into packed bf16 hi / lo pairs that stay in
// registers until the whole tile has been read (the A tile is then written over the stage, in place).
struct Packed { uint32_t h0, h1, l0, l1; };
__device__ __forceinline__ Packed keep_chunk(const float4& x, bool mp_only, bool live, float* out_chunk)
{
    Packed k = {0u, 0u, 0u, 0u};
    if (mp_only) { if (live) stg_f4(out_chunk, x); }
    else
    {
        split2(x.x, x.y, k.h0, k.l0);
        split2(x.z, x.w, k.h1, k.l1);
    }
    return k;
}
__device__ __forceinline__ void sts_packed(uint32_t a_dst, const Packed& k)
{
    sts_v2(a_dst, k.h0, k.h1);
    sts_v2(a_dst + A_BYTES, k.l0, k.l1);
}

// chunks 8 ks + j, ks < 3, of the pass's row (this thread: lane j of the row)
template <int NQ, bool LONG>
__device__ __forceinline__ void gather_pass(const GinFusedParams& p, const RowCtx& r, int maxdeg, uint32_t stage_thr, uint32_t ee_thr, int tile_start,
                                            bool mp_only, bool live, float* out_thr, Packed (&k)[3])
{
    k[0] = keep_chunk(gather_chunk<NQ, LONG, 0>(p, r, maxdeg, stage_thr, ee_thr, tile_start), mp_only, live, out_thr);
    k[1] = keep_chunk(gather_chunk<NQ, LONG, 128>(p, r, maxdeg, stage_thr, ee_thr, tile_start), mp_only, live, out_thr + 32);
    k[2] = keep_chunk(gather_chunk<NQ, LONG, 256>(p, r, maxdeg, stage_thr, ee_thr, tile_start), mp_only, live, out_thr + 64);
}

// The last float4 of a row (columns 96..99, chunk 24) does not fit 8 lanes x 3 steps: once per tile lanes 0..7 of a warp
// take it for the warp's 8 rows, a lane per row, walking the row's in-edges one by one (same CSR order).
__device__ __forceinline__ float4 tail_chunk(const GinFusedParams& p, const int4& d, int node, int deg, uint32_t own, uint32_t ee_tail, bool ext,
                                             uint32_t stage_tail, int tile_start)
{
    const float4 hv = lds_f4(own);
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
    const int dq[4] = {d.x, d.y, d.z, d.w};
    const float* h_tail = p.h_in + 4 * (Q - 1);
#pragma unroll
    for (int q = 0; q < 4; q++)
        if (q < deg)
        {
            const int rel = (dq[q] & 0xFFFF) - 32768;
            const float4 hu = ext ? ldg_f4(h_tail + (size_t)(node + rel) * D) : lds_f4(own + (uint32_t)(rel * ROW_BYTES));
            acc_edge2(m, lds_f4(ee_tail + ((dq[q] >> 16) & 0x3F) * ROW_BYTES), hu);
        }
    if (deg > 4)
    {
        const int eb = __ldg(p.in_ptr + node);
        for (int e = 4; e < deg; e++)
        {
            const int u = __ldg(p.src + eb + e), c = __ldg(p.code + eb + e);
            const float4 hu = ext ? ldg_f4(h_tail + (size_t)u * D) : lds_f4(stage_tail + (uint32_t)(u - tile_start) * ROW_BYTES);
            acc_edge2(m, lds_f4(ee_tail + c * ROW_BYTES), hu);
        }
    }
    return make_float4(m.x + hv.x, m.y + hv.y, m.z + hv.z, m.w + hv.w);
}

// A chunk of a row of a graph with more than 128 nodes: source rows may lie outside the stage and are read from global
// memory.  Rare (molhiv: 0.1 % of the graphs), kept out of line so that it does not cost the fast path registers.
__device__ __noinline__ float4 gather_chunk_ext(const float* h_chunk, const int* in_ptr, const int* src, const uint8_t* code, int node, int deg,
                                                uint32_t own, uint32_t ee_chunk)
{
    const float4 hv = lds_f4(own);
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
    const int eb = __ldg(in_ptr + node);
    for (int e = 0; e < deg; e++)
    {
        const int u = __ldg(src + eb + e), c = __ldg(code + eb + e);
        acc_edge2(m, lds_f4(ee_chunk + c * ROW_BYTES), ldg_f4(h_chunk + (size_t)u * D));
    }
    return make_float4(m.x + hv.x, m.y + hv.y, m.z + hv.z, m.w + hv.w);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NT, 1) gin_layer_fused_kernel(GinFusedParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    float* ee = reinterpret_cast<float*>(smem + Smem::EE);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + Smem::BAR);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + Smem::TMEM_PTR);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const bool mp_only = p.mp_only != 0;
    TRACE_K(0);

    if (tid == 0)
    {
        mbar_init(&bar[BAR_W], 1);
        // one barrier per buffer: the gather of tile t+1 does not wait for GEMM1(t), so both A tiles can be complete before the
        // MMA warp looks -- a single barrier would advance two phases and the parity wait would never return
        mbar_init(&bar[BAR_A_FULL], 2 * GATHER_WARPS);
        mbar_init(&bar[BAR_A_FULL + 1], 2 * GATHER_WARPS);
        // two each, by tile parity: GEMM1(t+1) is issued before the epilogue has looked at GEMM1(t), a single barrier could
        // advance two phases and the parity wait would never return
        for (int i = 0; i < 2; i++) { mbar_init(&bar[BAR_G1A_DONE + i], 1); mbar_init(&bar[BAR_G1B_DONE + i], 1); }
        mbar_init(&bar[BAR_A2A_FULL], 2 * EPI_WARPS);
        mbar_init(&bar[BAR_A2B_FULL], 2 * EPI_WARPS);
        mbar_init(&bar[BAR_G2_DONE], 1);
        for (int i = 0; i < 2; i++)
        {
            mbar_init(&bar[BAR_STAGE_FULL + i], 1);
            mbar_init(&bar[BAR_BUF_FREE + i], GATHER_WARPS);        // mp_only: the 16 gather warps are done reading the buffer
        }
        fence_mbar_init();
        if (!mp_only)
        {
            mbar_arrive_expect_tx(&bar[BAR_W], W_BYTES);
            tma_load_1d(smem + Smem::W, p.wpack + (size_t)rank * W_BYTES, W_BYTES, &bar[BAR_W]);
        }
    }
    __syncthreads();
    cluster_sync();          // both CTAs are running and their barriers are initialised
    if (warp == MMA_WARP)
    {
        tmem_alloc2(tmem_ptr, TMEM_COLS);
        tmem_relinquish2();
    }
    for (int i = tid; i < ED_COMBOS * Q; i += NT) st_f4(ee + 4 * i, ldg_f4(p.ee_comb + 4 * i));
    for (int i = tid; i < D; i += NT) ee[ED_COMBOS * D + i] = -3.0e38f;       // absent edge slots: relu(-3e38 + h) adds exactly 0
    for (int i = tid; i < ZERO_BYTES / 16; i += NT) st_f4(reinterpret_cast<float*>(smem + Smem::ZERO) + 4 * i, make_float4(0.f, 0.f, 0.f, 0.f));
    fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    cluster_sync();
    tc::fence_after_sync();
    const uint32_t tbase = *tmem_ptr;
    const uint32_t buf_base = smem_u32(smem + Smem::BUF);
    // Programmatic dependent launch: everything above overlapped the previous kernel's tail; from here on the kernel reads
    // what the previous kernels of the stream wrote (h_in, and in the first layer the tiles / descriptors of prep.cu)
    TRACE_K(1);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    TRACE_K(2);
#ifdef FG_TC2_TRACE
    if (p.trace && tid == 0) p.trace[2048 + blockIdx.x] = gtime();             // per-CTA loop start / end (tools/trace_fused.py)
#endif
    const int ntiles = p.tiles ? __ldg(p.tile_count) : (p.num_nodes + TM - 1) / TM;
    if (p.trace && blockIdx.x == 0 && tid == 0) p.trace[7] = (unsigned long long)ntiles;
    const int npt = (ntiles + 1) >> 1;                       // pair tiles
    // producer (one thread): the tile's feature rows + row descriptors -> buffer s (two bulk copies onto one barrier), its
    // record for the gather warps, and an L2 prefetch for the tile two further on
    auto produce = [&](int it, int pt) {
        const int s = it & 1;
        const int2 ti = tile_of(p, 2 * pt + (int)rank, ntiles);
        const int nrows = ti.y & 0xFFFF;
        const uint32_t bytes = (uint32_t)nrows * ROW_BYTES;
        unsigned char* dst = smem + Smem::BUF + s * BUF_BYTES;
        TRACE(2, it, 4);
        reinterpret_cast<int2*>(smem + Smem::TILE)[s] = ti;          // released to the gather warps by the arrival below
        mbar_arrive_expect_tx(&bar[BAR_STAGE_FULL + s], bytes + (p.row_desc ? nrows * 16 : 0));
        if (bytes)
        {
            tma_load_1d(dst, p.h_in + (size_t)ti.x * D, bytes, &bar[BAR_STAGE_FULL + s]);
            if (p.row_desc) tma_load_1d(dst + STAGE_BYTES, p.row_desc + ti.x, nrows * 16, &bar[BAR_STAGE_FULL + s]);
        }
        const int2 tn = tile_of(p, 2 * (pt + 2 * npairs) + (int)rank, ntiles);
        if (tn.y & 0xFFFF)
        {
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.h_in + (size_t)tn.x * D), "r"((tn.y & 0xFFFF) * ROW_BYTES) : "memory");
            if (p.row_desc) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.row_desc + tn.x), "r"((tn.y & 0xFFFF) * 16) : "memory");
        }
    };

    if (warp >= MMA_WARP)
    {
        reg_dec<REGS_MISC>();
        if (warp == MMA_WARP)
        {
            // ===== MMA issuer: the converged warp of the leader CTA, the tcgen05 instructions predicated on elect.sync.  (In
            // gin_tc2.cu this made the layer slower: the MMAs then hit shared memory in bursts and starved the gather.  Here
            // the gather warps are waiting for GEMM1 to release the A tile, so GEMM1 should run as fast as the tensor pipe can.)
            if (rank == 0 && !mp_only)
            {
                constexpr uint32_t IDESC1A = tc::idesc_bf16(2 * TM, N1A), IDESC1B = tc::idesc_bf16(2 * TM, N1B), IDESC2 = tc::idesc_bf16(2 * TM, N2);
                // the immediates of mma_*_tab assume these two bases
                if ((smem_u32(smem) & 0xFFFFFFu) != SMEM_BASE || tbase != 0) asm volatile("trap;");
                // Issue order: GEMM1 of tile t is interleaved with GEMM2 of tile t-1 --
                //   wait A(t), G1a(t) -> ring slot 2t % 3;  wait z_a(t-1) converted, G2a(t-1);  G1b(t) -> slot (2t+1) % 3 (the slot
                //   G2a(t-1) has just read);  wait z_b(t-1) converted, G2b(t-1)
                // so the tensor pipe works on GEMM1(t) while the epilogue converts tile t-1.  Every smem descriptor is a compile-time
                // constant (B selects the buffer, the tile loop is unrolled by two); ring columns are run-time values.
                auto g1_half = [&](auto BSEL, auto NH, uint32_t zcol) {
                    constexpr int B = decltype(BSEL)::value, nh = decltype(NH)::value;
                    static_for<0, 3 * K1_STEPS>([&](auto I) {
                        constexpr int e = decltype(I)::value, prod = e / K1_STEPS, j = e % K1_STEPS;
                        constexpr uint32_t lbo_b = nh ? LBO_W1B : LBO_W1A;
                        constexpr uint32_t a_start = Smem::BUF + B * BUF_BYTES + (prod == 1 ? A_BYTES : 0) + 2 * j * LBO_A;
                        // the last k-step pairs chunk 12 with the shared zero block (k = 104..111 does not exist)
                        constexpr uint64_t a_desc = desc_imm(a_start, j < K1_STEPS - 1 ? (uint32_t)LBO_A : (uint32_t)Smem::ZERO - a_start);
                        constexpr uint64_t b_desc = desc_imm(Smem::W + (nh ? (prod == 2 ? OFF_W1B_LO : OFF_W1B_HI) : (prod == 2 ? OFF_W1A_LO : OFF_W1A_HI)) + 2 * j * lbo_b, lbo_b);
                        mma_ss2_elect(zcol, a_desc, b_desc, nh ? IDESC1B : IDESC1A, e != 0);
                    });
                    commit2_elect(&bar[(nh ? BAR_G1B_DONE : BAR_G1A_DONE) + B]);
                };
                // GEMM2 K half kh of the tile whose parity is ph: A = z (tensor memory: hi at +0, lo at +8 of every 16-column k-step)
                auto g2_half = [&](auto KH, uint32_t zcol, uint32_t ph) {
                    constexpr int kh = decltype(KH)::value, NJA = N1A / 16, nj = kh ? K2_STEPS - NJA : NJA;
                    mbar_wait_park(&bar[kh ? BAR_A2B_FULL : BAR_A2A_FULL], ph);
                    tc::fence_after_sync();
                    static_for<0, 3 * nj>([&](auto I) {
                        constexpr int i = decltype(I)::value, prod = i / nj, jj = i % nj, j = (kh ? NJA : 0) + jj;
                        mma_ts2_elect(TC_H, zcol + (prod == 1 ? 8 : 0) + 16 * jj, desc_imm(Smem::W + (prod == 2 ? OFF_W2_LO : OFF_W2_HI) + 2 * j * LBO_W2, LBO_W2), IDESC2,
                                      (kh | i) != 0);
                    });
                    if constexpr (kh == 1) commit2_elect(&bar[BAR_G2_DONE]);
                };
                auto mma_tile = [&](auto BSEL, int it) {
                    constexpr int B = decltype(BSEL)::value;
                    const uint32_t za = TC_ZR + ((2 * it) % 3) * ZR_COLS, zb = TC_ZR + ((2 * it + 1) % 3) * ZR_COLS;
                    const uint32_t pza = TC_ZR + ((2 * it + 1) % 3) * ZR_COLS, pzb = TC_ZR + ((2 * it + 2) % 3) * ZR_COLS;   // slots of tile it - 1: (2it-2) % 3, (2it-1) % 3
                    if (lane == 0) TRACE(0, it, 0);
                    mbar_wait_park(&bar[BAR_A_FULL + B], (it >> 1) & 1);
                    tc::fence_after_sync();
                    if (lane == 0) TRACE(0, it, 1);
                    g1_half(BSEL, std::integral_constant<int, 0>{}, za);
                    if (it > 0) g2_half(std::integral_constant<int, 0>{}, pza, (uint32_t)((it - 1) & 1));
                    g1_half(BSEL, std::integral_constant<int, 1>{}, zb);
                    if (lane == 0) TRACE(0, it, 2);
                    if (it > 0) g2_half(std::integral_constant<int, 1>{}, pzb, (uint32_t)((it - 1) & 1));
                    if (lane == 0) TRACE(0, it, 5);
                };
                int it = 0;
                for (int pt = pair; pt < npt; pt += 2 * npairs, it += 2)
                {
                    mma_tile(std::integral_constant<int, 0>{}, it);
                    if (pt + npairs < npt) mma_tile(std::integral_constant<int, 1>{}, it + 1);
                }
                // GEMM2 of the last tile
                const int nt = pair < npt ? (npt - 1 - pair) / npairs + 1 : 0;
                if (nt > 0)
                {
                    g2_half(std::integral_constant<int, 0>{}, TC_ZR + ((2 * (nt - 1)) % 3) * ZR_COLS, (uint32_t)((nt - 1) & 1));
                    g2_half(std::integral_constant<int, 1>{}, TC_ZR + ((2 * (nt - 1) + 1) % 3) * ZR_COLS, (uint32_t)((nt - 1) & 1));
                }
            }
        }
    }
    else if (warp >= EPI_WARPS)
    {
        reg_inc<REGS_GATHER>();
        // ===== gather warps =====
        const int gw = warp - EPI_WARPS;
        const int g = lane >> 3, j = lane & 7;
        const uint32_t ee_thr = smem_u32(ee) + 16 * j;
        const uint32_t bar_full0 = mapa(smem_u32(&bar[BAR_A_FULL]), 0);
        const uint32_t zero_thr = smem_u32(smem + Smem::ZERO) + 16 * j;            // 2 KB of zeros: the "row" of a slot beyond the tile
        const bool has_desc = p.row_desc != nullptr;
        const int4 empty = make_int4(32768 | (ED_COMBOS << 16), 32768 | (ED_COMBOS << 16), 32768 | (ED_COMBOS << 16), 32768 | (ED_COMBOS << 16));

        // one pass = the four rows of this warp instruction (thread: lane j of row slot), chunks 8 ks + j -> k[0..2]
        auto do_pass = [&](int slot, int start, int rows, bool ext, uint32_t stage_thr, uint32_t desc_base, Packed (&k)[3], int& R_out) {
            // slot -> row: the tile's descriptors are ordered by in-degree, so the four rows of a warp instruction (almost
            // always) have in-edge lists of one length; a slot beyond the tile stands for itself
            const bool live = slot < rows;
            const int4 d = (live && has_desc) ? lds_i4(desc_base + slot * 16) : empty;
            const int R = (live && has_desc) ? ((d.y >> 24) & 0x7F) : slot;
            R_out = R;
            RowCtx r;
            r.node = start + (live ? R : 0);
            r.own = live ? stage_thr + (uint32_t)R * ROW_BYTES : zero_thr;
            r.so[0] = r.own + (uint32_t)(((d.x & 0xFFFF) - 32768) * ROW_BYTES);
            r.so[1] = r.own + (uint32_t)(((d.y & 0xFFFF) - 32768) * ROW_BYTES);
            r.so[2] = r.own + (uint32_t)(((d.z & 0xFFFF) - 32768) * ROW_BYTES);
            r.so[3] = r.own + (uint32_t)(((d.w & 0xFFFF) - 32768) * ROW_BYTES);
            r.tb[0] = ee_thr + ((d.x >> 16) & 0x3F) * ROW_BYTES;
            r.tb[1] = ee_thr + ((d.y >> 16) & 0x3F) * ROW_BYTES;
            r.tb[2] = ee_thr + ((d.z >> 16) & 0x3F) * ROW_BYTES;
            r.tb[3] = ee_thr + ((d.w >> 16) & 0x3F) * ROW_BYTES;
            r.deg = live ? (int)((unsigned)d.x >> 24) : 0;
            if (r.deg == 255) r.deg = __ldg(p.in_ptr + r.node + 1) - __ldg(p.in_ptr + r.node);
            int maxdeg = max(r.deg, __shfl_xor_sync(FULL, r.deg, 8));
            maxdeg = max(maxdeg, __shfl_xor_sync(FULL, maxdeg, 16));
            float* out_thr = p.h_out + (size_t)r.node * D + 4 * j;
            if (ext)
            {
#pragma unroll
                for (int ks = 0; ks < 3; ks++)
                    k[ks] = keep_chunk(gather_chunk_ext(p.h_in + 4 * (8 * ks + j), p.in_ptr, p.src, p.code, r.node, r.deg, r.own + 128 * ks, ee_thr + 128 * ks),
                                       mp_only, live, out_thr + 32 * ks);
            }
            else
            {
                // one specialisation per slot count of this warp instruction's four rows: a single uniform branch per pass
                switch (maxdeg)
                {
                case 0: gather_pass<0, false>(p, r, maxdeg, stage_thr, ee_thr, start, mp_only, live, out_thr, k); break;
                case 1: gather_pass<1, false>(p, r, maxdeg, stage_thr, ee_thr, start, mp_only, live, out_thr, k); break;
                case 2: gather_pass<2, false>(p, r, maxdeg, stage_thr, ee_thr, start, mp_only, live, out_thr, k); break;
                case 3: gather_pass<3, false>(p, r, maxdeg, stage_thr, ee_thr, start, mp_only, live, out_thr, k); break;
                case 4: gather_pass<4, false>(p, r, maxdeg, stage_thr, ee_thr, start, mp_only, live, out_thr, k); break;
                default: gather_pass<4, true>(p, r, maxdeg, stage_thr, ee_thr, start, mp_only, live, out_thr, k); break;
                }
            }
        };

        int it = 0;
        for (int pt = pair; pt < npt; pt += npairs, it++)
        {
            const int s = it & 1;
            if (gw == 0 && lane == 0) TRACE(2, it, 0);
            mbar_wait_park(&bar[BAR_STAGE_FULL + s], (it >> 1) & 1);
            if (gw == 0 && lane == 0) TRACE(2, it, 2);
            // the tile's record and row descriptors arrived with its rows: no global load, no register prefetch in these warps
            const int2 ti = reinterpret_cast<const int2*>(smem + Smem::TILE)[s];
            const int start = ti.x, rows = ti.y & 0xFFFF;
            const bool ext = (ti.y >> 30) & 1;
            const uint32_t buf = buf_base + s * BUF_BYTES;
            const uint32_t stage_thr = buf + 16 * j, desc_base = buf + STAGE_BYTES;
            Packed k0[3], k1[3], kt = {0u, 0u, 0u, 0u};
            int R0, R1, Rt = 0;
            // the slots are ordered by in-degree: pass 0 takes a group of four from the lower half, pass 1 from the upper half, so
            // that every warp gets a cheap and an expensive group (the A tile waits for the slowest warp)
            do_pass(gw * 4 + g, start, rows, ext, stage_thr, desc_base, k0, R0);
            do_pass(TM / 2 + gw * 4 + g, start, rows, ext, stage_thr, desc_base, k1, R1);
            // chunk 24 of the warp's 8 rows: lanes 0..7, a lane per row
            if (lane < ROWS_PER_WARP)
            {
                const int slot = (lane & 4) * (TM / 8) + gw * 4 + (lane & 3);       // the warp's 8 slots: gw*4 + {0..3}, 64 + gw*4 + {0..3}
                const bool live = slot < rows;
                const int4 d = (live && has_desc) ? lds_i4(desc_base + slot * 16) : empty;
                Rt = (live && has_desc) ? ((d.y >> 24) & 0x7F) : slot;
                const int node = start + (live ? Rt : 0);
                int deg = live ? (int)((unsigned)d.x >> 24) : 0;
                if (deg == 255) deg = __ldg(p.in_ptr + node + 1) - __ldg(p.in_ptr + node);
                const uint32_t stage_tail = buf + 16 * (Q - 1);
                const uint32_t own = live ? stage_tail + (uint32_t)Rt * ROW_BYTES : zero_thr - 16 * j;
                const float4 x = tail_chunk(p, d, node, deg, own, smem_u32(ee) + 16 * (Q - 1), ext, stage_tail, start);
                kt = keep_chunk(x, mp_only, live, p.h_out + (size_t)node * D + 4 * (Q - 1));
            }
            __syncwarp();
            if (gw == 0 && lane == 0) TRACE(2, it, 3);
            if (mp_only)
            {
                if (lane == 0) mbar_arrive(&bar[BAR_BUF_FREE + s]);          // the buffer may be refilled
                continue;
            }
            // every gather warp is done reading the stage: the A tile goes over it, in place (bf16 hi at +0, lo at +A_BYTES;
            // thread's 8-byte slot of row R in step ks: K chunk pair 4 ks + j / 2, half j % 2)
            asm volatile("bar.sync 1, %0;" ::"n"(GATHER_WARPS * 32) : "memory");
            {
                const uint32_t a_thr = buf + (j >> 1) * LBO_A + (j & 1) * 8;
#pragma unroll
                for (int ks = 0; ks < 3; ks++)
                {
                    sts_packed(a_thr + R0 * 16 + ks * (4 * LBO_A), k0[ks]);
                    sts_packed(a_thr + R1 * 16 + ks * (4 * LBO_A), k1[ks]);
                }
                if (lane < ROWS_PER_WARP)
                {
                    // K chunk 12 of the row: k = 96..99 from chunk 24, the bias column k = 100 (constant 1 in the hi tile: W1 carries
                    // b1 there), k = 101..103 zero
                    const uint32_t dst = buf + ((Q - 1) >> 1) * LBO_A + Rt * 16;
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(kt.h0), "r"(kt.h1), "r"(0x00003F80u), "r"(0u) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst + A_BYTES), "r"(kt.l0), "r"(kt.l1), "r"(0u), "r"(0u) : "memory");
                }
            }
            // the tile is visible to the tensor core (async proxy) and reported to the leader CTA
            fence_proxy_async();
            __syncwarp();
            if (lane == 0)
            {
                if (it == 0 && gw == 0) mbar_wait_park(&bar[BAR_W], 0);      // this CTA's weights have landed
                mbar_arrive_cluster(bar_full0 + 8 * s);
                if (gw == 0) TRACE(2, it, 1);
            }
        }
    }
    else if (mp_only)
    {
        // mp_only: no MMA, no epilogue -- thread 0 is just the producer, a buffer is free when the 16 gather warps have read it
        if (tid == 0)
        {
            int it = 0;
            for (int pt = pair; pt < npt; pt += npairs, it++)
            {
                if (it >= 2) mbar_wait_park(&bar[BAR_BUF_FREE + (it & 1)], ((it >> 1) - 1) & 1);
                produce(it, pt);
            }
        }
    }
    else
    {
        // ===== epilogue warps: two per TMEM lane quadrant (as gin_tc2.cu; rows beyond the tile are not stored) =====
        // thread 0 is also the TMA producer: tiles 0 and 1 up front, tile it + 2 as soon as GEMM1(it) has consumed its buffer
        if (tid == 0)
        {
            if (pair < npt) produce(0, pair);
            if (pair + npairs < npt) produce(1, pair + npairs);
        }
        const int quad = warp & 3, pp = warp >> 2;
        const uint32_t lane_base = tbase + ((uint32_t)(quad * 32) << 16);
        const uint32_t bar_a2a0 = mapa(smem_u32(&bar[BAR_A2A_FULL]), 0), bar_a2b0 = mapa(smem_u32(&bar[BAR_A2B_FULL]), 0);
        int it = 0;
        for (int pt = pair; pt < npt; pt += npairs, it++)
        {
            const uint32_t ph = it & 1;
            const int2 ti = tile_of(p, 2 * pt + (int)rank, ntiles);
            const int rows = ti.y & 0xFFFF;
            mbar_wait_park(&bar[BAR_G1A_DONE + (it & 1)], (it >> 1) & 1);
            tc::fence_after_sync();
            if (tid == 0) TRACE(1, it, 0);
            const uint32_t za = lane_base + TC_ZR + ((2 * it) % 3) * ZR_COLS, zb = lane_base + TC_ZR + ((2 * it + 1) % 3) * ZR_COLS;
            convert_range(za, pp, N1A / 16);
            tc::wait_st();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(bar_a2a0);
            if (tid == 0) TRACE(1, it, 1);

            mbar_wait_park(&bar[BAR_G1B_DONE + (it & 1)], (it >> 1) & 1);
            tc::fence_after_sync();
            // GEMM1 is complete (both N halves; the commit is multicast to both CTAs): the A tile is consumed, this CTA's buffer
            // receives the rows of tile it + 2
            if (tid == 0 && pt + 2 * npairs < npt) produce(it + 2, pt + 2 * npairs);
            if (tid == 0) TRACE(1, it, 2);
            convert_range(zb, pp ^ 1, N1B / 16);
            tc::wait_st();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(bar_a2b0);
            if (tid == 0) TRACE(1, it, 3);

            mbar_wait_park(&bar[BAR_G2_DONE], ph);
            tc::fence_after_sync();
            if (tid == 0) TRACE(1, it, 4);
            // 16-lane x 256-bit TMEM loads give thread t columns 8g + 2(t%4), +1 of rows t/4 and t/4 + 8
            const int ra = quad * 32 + pp * 16 + (lane >> 2), rb = ra + 8;
            const bool va = ra < rows, vb = rb < rows;
            const long row_a = (long)ti.x + ra, row_b = (long)ti.x + rb;
            const uint32_t ta = lane_base + ((uint32_t)(pp * 16) << 16) + TC_H;
            auto ld_h = [&](int g4, uint32_t (&r)[16]) {
                asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                               "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                             : "r"(ta + 8 * g4)
                             : "memory");
            };
            auto st_h = [&](int g4, const uint32_t (&r)[16]) {
#pragma unroll
                for (int gg = 0; gg < 4; gg++)
                {
                    const int col = 8 * (g4 + gg) + 2 * (lane & 3);
                    if (8 * (g4 + gg) < D && col < D)
                    {
                        float2 oa = make_float2(__uint_as_float(r[4 * gg]), __uint_as_float(r[4 * gg + 1]));
                        float2 ob = make_float2(__uint_as_float(r[4 * gg + 2]), __uint_as_float(r[4 * gg + 3]));
                        if (p.relu_out)
                        {
                            oa = make_float2(relu_nan(oa.x), relu_nan(oa.y));
                            ob = make_float2(relu_nan(ob.x), relu_nan(ob.y));
                        }
                        if (va) asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(p.h_out + (size_t)row_a * D + col), "f"(oa.x), "f"(oa.y) : "memory");
                        if (vb) asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(p.h_out + (size_t)row_b * D + col), "f"(ob.x), "f"(ob.y) : "memory");
                    }
                }
            };
            float dot_a = 0.f, dot_b = 0.f;
            auto dot_h = [&](int g4, const uint32_t (&r)[16]) {
#pragma unroll
                for (int gg = 0; gg < 4; gg++)
                {
                    const int col = 8 * (g4 + gg) + 2 * (lane & 3);
                    if (8 * (g4 + gg) < D && col < D)
                    {
                        const float2 wv = __ldg(reinterpret_cast<const float2*>(p.head_w + col));
                        dot_a = fmaf(__uint_as_float(r[4 * gg + 1]), wv.y, fmaf(__uint_as_float(r[4 * gg]), wv.x, dot_a));
                        dot_b = fmaf(__uint_as_float(r[4 * gg + 3]), wv.y, fmaf(__uint_as_float(r[4 * gg + 2]), wv.x, dot_b));
                    }
                }
            };
            {
                uint32_t r0[16], r1[16];
                ld_h(0, r0);
                if (p.head_w == nullptr)
                {
                    tc::wait_ld(); ld_h(4, r1); st_h(0, r0);
                    tc::wait_ld(); ld_h(8, r0); st_h(4, r1);
                    tc::wait_ld(); ld_h(12, r1); st_h(8, r0);
                    tc::wait_ld(); st_h(12, r1);
                }
                else
                {
                    tc::wait_ld(); ld_h(4, r1); dot_h(0, r0);
                    tc::wait_ld(); ld_h(8, r0); dot_h(4, r1);
                    tc::wait_ld(); ld_h(12, r1); dot_h(8, r0);
                    tc::wait_ld(); dot_h(12, r1);
                    dot_a += __shfl_xor_sync(FULL, dot_a, 1); dot_a += __shfl_xor_sync(FULL, dot_a, 2);
                    dot_b += __shfl_xor_sync(FULL, dot_b, 1); dot_b += __shfl_xor_sync(FULL, dot_b, 2);
                    if ((lane & 3) == 0)
                    {
                        if (va) p.node_dot[row_a] = dot_a;
                        if (vb) p.node_dot[row_b] = dot_b;
                    }
                }
            }
            if (tid == 0) TRACE(1, it, 5);
        }
    }

    // both CTAs must be done with tensor memory, shared memory and each other's barriers before either leaves
    tc::fence_before_sync();
    __syncthreads();
    TRACE_K(3);
#ifdef FG_TC2_TRACE
    if (p.trace && tid == 0) p.trace[2048 + 256 + blockIdx.x] = gtime();
#endif
    __syncwarp();
    cluster_sync();
    if (warp == MMA_WARP) tmem_dealloc2(tbase, TMEM_COLS);
}

}  // namespace

extern unsigned long long* gin_tc2_trace_buffer;      // set through flowgnn_b200_debug_trace (api.cu)

int gin_layer_fused_launch(const DeviceBatch& b, const GinWeights& w, int layer, const float* h_in, float* h_out, int sm_count, cudaStream_t s,
                           const float* head_w, float* node_dot, bool mlp_only, int mp_only)
{
    FG_TRY(opt_in_smem(reinterpret_cast<const void*>(&gin_layer_fused_kernel), Smem::BYTES));
    GinFusedParams p;
    p.h_in = h_in; p.h_out = h_out;
    p.in_ptr = b.in_ptr.as<int>(); p.src = b.src.as<int>(); p.code = b.code.as<uint8_t>();
    p.row_desc = mlp_only ? nullptr : b.row_desc_sorted.as<int4>();      // nullptr: every row without in-edges = the node MLP alone (dense graphs, gin.cu)
    p.tiles = mlp_only ? nullptr : b.tiles.as<int2>(); p.tile_count = b.tile_count.as<int>();
    p.ee_comb = w.ee_comb.as<float>() + (size_t)layer * ED_COMBOS * D;
    p.wpack = w.wpack2.as<unsigned char>() + (size_t)layer * 2 * W_BYTES;
    p.num_nodes = (int)b.total_nodes;
    p.relu_out = (layer != 4);
    p.mp_only = mp_only;
    p.head_w = head_w; p.node_dot = node_dot;
    p.trace = gin_tc2_trace_buffer;
    const long tiles_bound = mlp_only ? (b.total_nodes + TM - 1) / TM : b.max_tiles;
    const int pairs = (int)std::max<long>(1, std::min<long>((tiles_bound + 1) / 2, sm_count / 2));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = Smem::BYTES; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    FG_CUDA(cudaLaunchKernelEx(&cfg, gin_layer_fused_kernel, p));
    return 0;
}

}  // namespace fg

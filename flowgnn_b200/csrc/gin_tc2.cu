// GIN / GIN-VN layer on the B200 tensor cores, CTA-PAIR version: one persistent cluster of two CTAs per TPC
// (tcgen05 cta_group::2, UMMA M = 256 = 2 x 128 consecutive nodes), warp specialised.
//
// Reference work per layer (GIN/src/message_passing.cc:77-150, node_embedding.cc:23-201):
//   m_v = sum_{(u,v)} relu(h_u + EE_l[attr_uv]);  a_v = m_v + h_v  (eps is never loaded: SURVEY.md F4)
//   z = relu(W1 a + b1);  h'_v = W2 z + b2  (+ relu unless last layer)
//
// Why a pair: with cta_group::2 each CTA holds only HALF of every weight matrix (the N rows of the B operand are
// split over the pair), 94 KB instead of 186 KB.  The shared memory this frees holds a double-buffered bf16 hi/lo
// tile of a_v in the UMMA canonical K-major layout, so the edge gather is no longer tied to the tensor-memory
// fragment layout (4 lanes per feature row, 64-byte pieces, ~11 L1 wavefronts per load instruction): it uses 8 lanes
// per row (128-byte pieces, 4 rows per instruction) and twice as many gather warps.
//
// Per CTA (896 threads):
//   warps 0-7   epilogue: z = relu(acc) -> bf16 hi/lo written back to tensor memory IN PLACE (A operand of GEMM2), the
//               TMEM load of the next 16-column chunk in flight while one is converted; h' = acc (+ relu) streamed to
//               HBM, one full 32-byte sector per row and store.  The biases are inside the GEMMs (constant-1 column
//               k = 100 of the A tile, z column 200 == 1).  In the LAST layer the four lanes of a row multiply h' with the
//               prediction weights instead and store one float per node (the head is fused, nothing else leaves)
//   warps 8-23  gather: 8 rows each, as two passes of 4 rows; thread (g = lane / 8, j = lane % 8) owns row 4 pass + g
//               and, per step ks < 3, the float4 chunk 8 ks + j of that row: own row + up to four source rows are
//               loaded with addresses = pointer + immediate (from the node's 16-byte row descriptor, prep.cu), reduced
//               in CSR order (deterministic), split to bf16 hi/lo and stored to the A tile.  Software pipeline over
//               steps, passes and tiles on two load buffers; chunk 24 is gathered once per tile with a lane per row
//   warp 24     (leader CTA only) MMA issuer: GEMM1 = 3 products (hi*hi + lo*hi + hi*lo) x 7 k-steps from shared
//               memory (SS), as two N halves (112 + 96) so that the conversion of the first half overlaps the second;
//               GEMM2 = 3 x 13 k-steps with A = z from tensor memory (TS), N = 128 (TS needs N % 32 == 0 for pairs)
//   warp 25     L2 prefetch of the feature rows and the CSR slice two tiles ahead
// Barriers that the MMA issuer waits on live in the leader CTA and are arrived on remotely by the peer's warps;
// tcgen05.commit multicasts completion to both CTAs.  Launched with programmatic stream serialization: the prologue
// (barriers, tensor memory, weights, tables) overlaps the previous layer's tail, griddepcontrol.wait guards h_in / h_out.
// What bounds the kernel (the SM's shared-memory / L1 data path) and what was tried: DESIGN.md 5.1.
#include "internal.cuh"
#include "layers.cuh"
#include "tc.cuh"
#include "pair.cuh"
#include "gin_wpack.cuh"

#include <algorithm>
#include <cstdlib>
#include <type_traits>

namespace fg {

namespace {

using namespace ginw;

// A operand tile in shared memory: canonical no-swizzle K-major, byte(r, k) = (k / 8) * LBO_A + r * 16 + (k % 8) * 2;
// the 32-byte pad per chunk makes the 8-byte stores of a warp (4 rows x 8 lanes) hit every bank group exactly twice
constexpr int LBO_A = TM * 16 + 32;
constexpr int A_BYTES = K1_CHUNKS * LBO_A;              // one hi or lo buffer: 26,832
constexpr int ZERO_BYTES = TM * 16;                     // K chunk 13 of every A buffer (and the tail of W2_lo) reads from here

// warp roles (warpgroup aligned, for setmaxnreg): 0-7 epilogue, 8-23 gather, 24 MMA issuer, 25 L2 prefetch, 26-27 idle
constexpr int EPI_WARPS = 8, GATHER_WARPS = 16;
constexpr int MMA_WARP = EPI_WARPS + GATHER_WARPS, LOAD_WARP = MMA_WARP + 1;
constexpr int NT = (MMA_WARP + 4) * 32;       // 896
constexpr int REGS_LAUNCH = 72;               // 65,536 / 896 rounded down to a multiple of 8
constexpr int REGS_EPI = 80, REGS_MISC = 24, REGS_GATHER = 80;
// setmaxnreg moves registers inside the CTA's launch allocation: the new sizes must fit it or the increase never returns
static_assert(32 * (EPI_WARPS * REGS_EPI + GATHER_WARPS * REGS_GATHER + 4 * REGS_MISC) <= NT * REGS_LAUNCH, "setmaxnreg pool");
constexpr int ROWS_PER_WARP = TM / GATHER_WARPS;   // 8
constexpr int LPR = 8;                             // lanes per row
constexpr int GSTEPS = 3;                          // steps per row: chunk 8 ks + j; chunk 24 is gathered once per tile with one lane per row

// tensor-memory columns
constexpr uint32_t TC_Z = 0, TC_H = 256;
constexpr uint32_t TMEM_COLS = 512;

struct Smem {
    static constexpr int W = 0;
    static constexpr int A = W + W_BYTES;                           // [2 stages][hi, lo][A_BYTES]
    static constexpr int ZERO = A + 4 * A_BYTES;                    // K chunk 13 of every A buffer (must lie above them: LBO >= 0)
    static constexpr int EE = ZERO + ZERO_BYTES;                    // [61][100] fp32 combined edge-embedding rows; row 60 = sentinel
    static constexpr int BAR = EE + (ED_COMBOS + 1) * D * 4;
    static constexpr int TMEM_PTR = BAR + 16 * 8;
    static constexpr int BYTES = TMEM_PTR + 16;
};
static_assert(Smem::ZERO % 16 == 0 && Smem::A % 16 == 0 && Smem::EE % 16 == 0 && Smem::BAR % 8 == 0, "alignment");
static_assert(Smem::BYTES <= 232448, "shared memory budget");

enum { BAR_W = 0, BAR_A_FULL /* 2 */, BAR_A_FREE = BAR_A_FULL + 2 /* 2 */, BAR_G1A_DONE = BAR_A_FREE + 2, BAR_G1B_DONE, BAR_A2A_FULL, BAR_A2B_FULL, BAR_G2_DONE };

struct GinTc2Params {
    const float* h_in; float* h_out;
    const int* in_ptr; const int* src; const uint8_t* code;
    const int4* row_desc;            // [N] first four in-edges of every node, packed (prep.cu)
    const float* ee_comb;            // [60][100] this layer: ((0 + T[a0]) + T[5 + a1]) + T[11 + a2]
    const unsigned char* wpack;      // [2 ranks][W_BYTES] this layer
    int num_nodes; int num_pair_tiles; int relu_out;
    // last layer only: instead of storing h' (400 B per node, re-read by the pooling kernel) the epilogue reduces every row
    // with the prediction weights and stores node_dot[v] = <h'_v, w_pred> (4 B per node): mean_v(h'_v) . w == mean_v(h'_v . w)
    const float* head_w; float* node_dot;
    // experiments, compiled in with -DFG_TC2_TRACE only: dbg 1 = no in-edges, 2 = no h' stores, 4 = no z conversion (wrong
    // results, for bottleneck elimination); trace = timeline of pair 0 (tools/trace_gin.py): [role][tile][event] globaltimer ns
    int dbg;
    unsigned long long* trace;
};

using namespace pair;

// ---- one destination row as seen by one of its 8 threads ----------------------------------------------------------
// Pointers to this thread's 16-byte chunk (step 0) of the row itself and of the source rows of its first four
// in-edges, the shared-memory addresses of their edge-embedding rows, the in-degree.  Everything comes from the
// node's 16-byte row descriptor (prep.cu): no in_ptr -> src/code pointer chase.  Absent slots are never loaded (they
// read as 0) and point at the sentinel table row; in-edges beyond the fourth are read from the CSR arrays.
struct RowEdges {
    const float* hv;
    const float* hu[4];
    uint32_t t[4];
    int deg;
};

// predicated 16-byte load (no branch, no memory traffic when `on` is false): absent edge slots read as h_u = 0.
// The pointers stay live across the steps, which keeps ptxas from loading over the address registers.
__device__ __forceinline__ float4 ldg_f4_pred(const float* ptr, bool on)
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

struct RowLoads { float4 hv; float4 hu[4]; };

template <int KS>
__device__ __forceinline__ void row_loads(const float* hv, const float* const (&hu)[4], int deg, int j, RowLoads& L)
{
    constexpr int OFF = 4 * LPR * KS;                        // floats
    const bool on = true;
    L.hv = ldg_f4_pred(hv + OFF, on);
#pragma unroll
    for (int s = 0; s < 4; s++) L.hu[s] = ldg_f4_pred(hu[s] + OFF, on && s < deg);
}

// reduce one step of one row (CSR order), split to bf16 hi / lo and store this thread's 4 columns to the A tile
template <int KS>
__device__ __forceinline__ void row_finish(const GinTc2Params& p, const RowLoads& L, const uint32_t (&t)[4], int deg, int maxdeg, int node, bool live, int j,
                                           const float* h_thr, uint32_t ee_thr, uint32_t a_dst)
{
    constexpr int OFF = 4 * LPR * KS;
    const bool on = true;
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < 4; q++) acc_edge(m, lds_f4(t[q] + 4 * OFF), L.hu[q]);
    if (maxdeg > 4)
    {
        // long in-edge lists (virtual nodes, kNN graphs): rounds of four edges from the CSR arrays, loads first
        const int eb = __ldg(p.in_ptr + min(node, p.num_nodes - 1));
        for (int e4 = 4; e4 < maxdeg; e4 += 4)
        {
            float4 hx[4];
            uint32_t tx[4];
#pragma unroll
            for (int q = 0; q < 4; q++)
            {
                const bool ok = on && e4 + q < deg;
                int u = 0, c = ED_COMBOS;
                if (ok) { u = __ldg(p.src + eb + e4 + q); c = __ldg(p.code + eb + e4 + q); }
                hx[q] = ldg_f4_pred(h_thr + (size_t)u * D + OFF, ok);
                tx[q] = ee_thr + c * (D * 4) + 4 * OFF;
            }
#pragma unroll
            for (int q = 0; q < 4; q++) acc_edge(m, lds_f4(tx[q]), hx[q]);
        }
    }
    if (on)
    {
        uint32_t h0, l0, h1, l1;
        split2(live ? m.x + L.hv.x : 0.f, live ? m.y + L.hv.y : 0.f, h0, l0);
        split2(live ? m.z + L.hv.z : 0.f, live ? m.w + L.hv.w : 0.f, h1, l1);
        sts_v2(a_dst + KS * (4 * LBO_A), h0, h1);
        sts_v2(a_dst + KS * (4 * LBO_A) + A_BYTES, l0, l1);
    }
}

__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#ifdef FG_TC2_TRACE
#define TRACE(role, it, ev) do { if (p.trace && pair == 0 && rank == 0 && (it) < 64) p.trace[((role) * 64 + (it)) * 8 + (ev)] = gtime(); } while (0)
#define DBG(bit) (p.dbg & (bit))
#else
#define TRACE(role, it, ev) do { } while (0)
#define DBG(bit) false
#endif

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NT, 1) gin_layer_tc2_kernel(GinTc2Params p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    float* ee = reinterpret_cast<float*>(smem + Smem::EE);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + Smem::BAR);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + Smem::TMEM_PTR);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (tid == 0)
    {
        mbar_init(&bar[BAR_W], 1);
        for (int i = 0; i < 2; i++)
        {
            mbar_init(&bar[BAR_A_FULL + i], 2 * GATHER_WARPS);
            mbar_init(&bar[BAR_A_FREE + i], 1);
        }
        mbar_init(&bar[BAR_G1A_DONE], 1);
        mbar_init(&bar[BAR_G1B_DONE], 1);
        mbar_init(&bar[BAR_A2A_FULL], 2 * EPI_WARPS);
        mbar_init(&bar[BAR_A2B_FULL], 2 * EPI_WARPS);
        mbar_init(&bar[BAR_G2_DONE], 1);
        fence_mbar_init();
        // this CTA's half of the weights: one bulk copy, waited for by the first gather warp before it reports tile 0
        mbar_arrive_expect_tx(&bar[BAR_W], W_BYTES);
        tma_load_1d(smem + Smem::W, p.wpack + (size_t)rank * W_BYTES, W_BYTES, &bar[BAR_W]);
    }
    __syncthreads();
    cluster_sync();          // both CTAs are running and their barriers are initialised
    if (warp == MMA_WARP)
    {
        tmem_alloc2(tmem_ptr, TMEM_COLS);
        tmem_relinquish2();
    }
    for (int i = tid; i < ED_COMBOS * Q; i += NT) st_f4(ee + 4 * i, ldg_f4(p.ee_comb + 4 * i));
    for (int i = tid; i < D; i += NT) ee[ED_COMBOS * D + i] = -3.0e38f;       // absent edge slots: relu(-3e38 + 0) adds exactly 0
    // A buffers + zero block (k = 101..103 of every row stays zero for the whole launch)
    for (int i = tid; i < (Smem::EE - Smem::A) / 16; i += NT) st_f4(reinterpret_cast<float*>(smem + Smem::A) + 4 * i, make_float4(0.f, 0.f, 0.f, 0.f));
    __syncthreads();
    // bias column: a_hi[row][k = 100] = 1 in both stages, never overwritten (the gather writes k < 100 only)
    for (int i = tid; i < 2 * TM; i += NT)
        *reinterpret_cast<uint16_t*>(smem + Smem::A + (i / TM) * 2 * A_BYTES + (D / 8) * LBO_A + (i % TM) * 16 + (D % 8) * 2) = 0x3F80;
    fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    cluster_sync();
    tc::fence_after_sync();
    const uint32_t tbase = *tmem_ptr;
    const uint32_t a_base = smem_u32(smem + Smem::A);
    // Programmatic dependent launch: this grid may have been started while the previous kernel of the stream (the previous
    // layer) was still draining -- everything above (barriers, tensor memory, weights, tables: nothing the previous kernel
    // writes) overlapped its tail.  From here on the kernel reads h_in and writes h_out: wait for the previous grid, and let
    // the next one start its own prologue as soon as this grid's CTAs retire.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp >= MMA_WARP)
    {
        reg_dec<REGS_MISC>();
        if (warp == MMA_WARP)
        {
            // ===== MMA issuer: one thread of the leader CTA =====
            // (issuing from the converged warp under elect.sync removes ptxas' per-MMA elect-and-retry loop, but measured
            // 20 % slower: the MMAs then hit shared memory in bursts and starve the gather -- see DESIGN.md 5.1)
            if (rank == 0 && lane == 0)
            {
                const uint32_t w_addr = smem_u32(smem + Smem::W);
                const uint32_t zero_addr = smem_u32(smem + Smem::ZERO);
                const uint32_t idesc1a = tc::idesc_bf16(2 * TM, N1A), idesc1b = tc::idesc_bf16(2 * TM, N1B), idesc2 = tc::idesc_bf16(2 * TM, N2);
                int it = 0;
                for (int t = pair; t < p.num_pair_tiles; t += npairs, it++)
                {
                    const uint32_t ph = it & 1, s = it & 1;
                    TRACE(0, it, 0);
                    mbar_wait_park(&bar[BAR_A_FULL + s], (it >> 1) & 1);
                    tc::fence_after_sync();
                    TRACE(0, it, 1);
                    // GEMM1, N half a (z columns 0..111) then half b (112..207)
                    if constexpr (G1_SINGLE)
                    {
                        const uint32_t idesc1 = tc::idesc_bf16(2 * TM, N1);
                        bool acc = false;
#pragma unroll
                        for (int prod = 0; prod < 3; prod++)
                        {
                            const uint32_t a_addr = a_base + (2 * s + (prod == 1 ? 1 : 0)) * A_BYTES;
                            const uint32_t b_addr = w_addr + (prod == 2 ? OFF_W1A_LO + W1B_BYTES : OFF_W1A_HI);      // W1_hi | W1_lo, 21,632 B each
#pragma unroll
                            for (int j = 0; j < K1_STEPS; j++)
                            {
                                const uint32_t a_start = a_addr + 2 * j * LBO_A;
                                const uint32_t a_lbo = (j < K1_STEPS - 1) ? (uint32_t)LBO_A : zero_addr - a_start;
                                mma_ss2(tbase + TC_Z, tc::smem_desc(a_start, a_lbo, 128), tc::smem_desc(b_addr + 2 * j * LBO_W1, LBO_W1, 128), idesc1, acc);
                                acc = true;
                            }
                        }
                        commit2(&bar[BAR_G1A_DONE]);
                        commit2(&bar[BAR_G1B_DONE]);
                    }
                    else
                    {
#pragma unroll
                    for (int nh = 0; nh < 2; nh++)
                    {
                        bool acc = false;
                        const uint32_t lbo_b = nh ? LBO_W1B : LBO_W1A;
#pragma unroll
                        for (int prod = 0; prod < 3; prod++)
                        {
                            const uint32_t a_addr = a_base + (2 * s + (prod == 1 ? 1 : 0)) * A_BYTES;
                            const uint32_t b_addr = w_addr + (nh ? (prod == 2 ? OFF_W1B_LO : OFF_W1B_HI) : (prod == 2 ? OFF_W1A_LO : OFF_W1A_HI));
#pragma unroll
                            for (int j = 0; j < K1_STEPS; j++)
                            {
                                const uint32_t a_start = a_addr + 2 * j * LBO_A;
                                // the last k-step pairs chunk 12 with the shared zero block (k = 104..111 does not exist)
                                const uint32_t a_lbo = (j < K1_STEPS - 1) ? (uint32_t)LBO_A : zero_addr - a_start;
                                mma_ss2(tbase + TC_Z + (nh ? N1A : 0), tc::smem_desc(a_start, a_lbo, 128), tc::smem_desc(b_addr + 2 * j * lbo_b, lbo_b, 128),
                                        nh ? idesc1b : idesc1a, acc);
                                acc = true;
                            }
                        }
                        commit2(&bar[nh ? BAR_G1B_DONE : BAR_G1A_DONE]);
                    }
                    }
                    commit2(&bar[BAR_A_FREE + s]);
                    TRACE(0, it, 2);
                    // GEMM2, K half a (k-steps 0..6, operand columns converted from z half a) then half b (7..12)
                    bool acc = false;
#pragma unroll
                    for (int kh = 0; kh < 2; kh++)
                    {
                        mbar_wait_park(&bar[kh ? BAR_A2B_FULL : BAR_A2A_FULL], ph);
                        tc::fence_after_sync();
                        TRACE(0, it, 3 + kh);
#pragma unroll
                        for (int prod = 0; prod < 3; prod++)
                        {
                            const uint32_t a_col = tbase + TC_Z + (prod == 1 ? 8 : 0);
                            const uint32_t b_addr = w_addr + (prod == 2 ? OFF_W2_LO : OFF_W2_HI);
#pragma unroll
                            for (int j = (kh ? N1A / 16 : 0); j < (kh ? K2_STEPS : N1A / 16); j++)
                            {
                                mma_ts2(tbase + TC_H, a_col + 16 * j, tc::smem_desc(b_addr + 2 * j * LBO_W2, LBO_W2, 128), idesc2, acc);
                                acc = true;
                            }
                        }
                    }
                    commit2(&bar[BAR_G2_DONE]);
                    TRACE(0, it, 5);
                }
            }
        }
        else if (warp == LOAD_WARP)
        {
            // ===== L2 prefetch of the feature rows of the tile after next: 4 rows (1,600 B) per lane =====
            for (int t = pair + 2 * npairs; t < p.num_pair_tiles; t += npairs)
            {
                const long n0 = ((long)t * 2 + rank) * TM;
                const int rn = (int)min((long)TM, (long)p.num_nodes - n0);
                const int r0 = 4 * lane, nr = min(4, rn - r0);
                if (nr > 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.h_in + (n0 + r0) * D), "r"(nr * D * 4) : "memory");
                // ... and its CSR slice (row pointers, sources, codes), so that the gather warps' dependent loads hit L2
                if (rn > 0)
                {
                    const int e_lo = __ldg(p.in_ptr + n0), e_hi = __ldg(p.in_ptr + n0 + rn);
                    const int eb = e_lo + 32 * lane;                       // 32 edges per lane: 128 B of sources, 32 B of codes
                    if (eb < e_hi)
                    {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.src + eb) : "memory");
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.code + eb) : "memory");
                    }
                    if (lane < 5) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.in_ptr + n0 + 32 * lane) : "memory");
                }
                // pace the prefetch: do not run more than two tiles ahead of the tensor pipe
                mbar_wait_park(&bar[BAR_G1A_DONE], ((t - pair) / npairs - 2) & 1);
            }
        }
    }
    else if (warp >= EPI_WARPS)
    {
        if constexpr (REGS_GATHER > REGS_LAUNCH) reg_inc<REGS_GATHER>(); else reg_dec<REGS_GATHER>();
        // ===== gather warps: a_v for 8 rows each, written as bf16 hi/lo into the shared-memory A tile =====
        // Software pipeline over (tile, pass, step): the loads of the next step -- of the next pass, of the next tile --
        // are always in flight while the current one is reduced; the row descriptor is prefetched one pass ahead.
        const int gw = warp - EPI_WARPS;
        const int g = lane >> 3, j = lane & 7;
        const uint32_t ee_thr = smem_u32(ee) + 16 * j;
        const float* h_thr = p.h_in + 4 * j;
        const uint32_t bar_full0 = mapa(smem_u32(&bar[BAR_A_FULL]), 0);
        // this thread's 8-byte slot in row (8 gw + g) of the A tile, step 0: chunk j -> K chunk pair j / 2, half j % 2
        const uint32_t a_thr = a_base + (j >> 1) * LBO_A + (j & 1) * 8 + (gw * ROWS_PER_WARP + g) * 16;
        const int tile_rows = 2 * npairs * TM;                                    // row distance between this CTA's tiles
        const int last = p.num_nodes - 1;

        // decode a row descriptor into pointers / table addresses
        auto decode_ptrs = [&](const int4& d, int node, const float*& hv, const float* (&hu)[4]) {
            const int nc = min(node, last);
            hv = h_thr + (size_t)nc * D;
            hu[0] = h_thr + (size_t)(nc + (d.x & 0xFFFF) - 32768) * D;
            hu[1] = h_thr + (size_t)(nc + (d.y & 0xFFFF) - 32768) * D;
            hu[2] = h_thr + (size_t)(nc + (d.z & 0xFFFF) - 32768) * D;
            hu[3] = h_thr + (size_t)(nc + (d.w & 0xFFFF) - 32768) * D;
        };
        auto decode_tabs = [&](const int4& d, int node, uint32_t (&t)[4], int& deg, int& maxdeg) {
            t[0] = ee_thr + ((d.x >> 16) & 0x3F) * (D * 4);
            t[1] = ee_thr + ((d.y >> 16) & 0x3F) * (D * 4);
            t[2] = ee_thr + ((d.z >> 16) & 0x3F) * (D * 4);
            t[3] = ee_thr + ((d.w >> 16) & 0x3F) * (D * 4);
            deg = node <= last ? (int)((unsigned)d.x >> 24) : 0;
            if (DBG(1)) deg = 0;
            if (deg == 255) deg = __ldg(p.in_ptr + min(node, last) + 1) - __ldg(p.in_ptr + min(node, last));
            // longest in-edge list of the four rows of this pass (warp-uniform trip count of the tail rounds)
            maxdeg = max(deg, __shfl_xor_sync(FULL, deg, 8));
            maxdeg = max(maxdeg, __shfl_xor_sync(FULL, maxdeg, 16));
        };

        int node = (pair * 2 + (int)rank) * TM + gw * ROWS_PER_WARP + g;         // this thread's row in pass 0 of its first tile
        const float* hv;
        const float* hu[4];
        uint32_t tab[4];
        int deg, maxdeg;
        RowLoads La, Lb;
        int4 dn;                                                                  // descriptor of the NEXT pass's row
        {
            const int4 d0 = __ldg(p.row_desc + min(node, last));
            decode_ptrs(d0, node, hv, hu);
            decode_tabs(d0, node, tab, deg, maxdeg);
            row_loads<0>(hv, hu, deg, j, La);
            dn = __ldg(p.row_desc + min(node + 4, last));
        }
        int it = 0;
        for (int t = pair; t < p.num_pair_tiles; t += npairs, it++)
        {
            const int s = it & 1;
            if (gw == 0 && lane == 0) TRACE(2, it, 0);
            // one pass = 3 steps on 2 load buffers, so the buffers swap roles from pass to pass: X holds steps 0 and 2,
            // Y step 1 and then step 0 of the NEXT pass
            auto do_pass = [&](int pass, RowLoads& X, RowLoads& Y) {
                const bool live = node <= last;
                const uint32_t a_dst = a_thr + 2 * s * A_BYTES + pass * (4 * 16);
                const int node_next = pass == 0 ? node + 4 : node - 4 + tile_rows;
                row_loads<1>(hv, hu, deg, j, Y);
                if (pass == 0 && it >= 2) mbar_wait_park(&bar[BAR_A_FREE + s], ((it >> 1) - 1) & 1);
                row_finish<0>(p, X, tab, deg, maxdeg, node, live, j, h_thr, ee_thr, a_dst);
                // the next row's descriptor was requested together with X's loads: it is here by now.  Copying it out at this
                // point (and not where it is decoded) keeps its scoreboard from serialising behind the loads issued below.
                const int4 dc = make_int4(reg_copy(dn.x), reg_copy(dn.y), reg_copy(dn.z), reg_copy(dn.w));
                row_loads<2>(hv, hu, deg, j, X);
                row_finish<1>(p, Y, tab, deg, maxdeg, node, live, j, h_thr, ee_thr, a_dst);
                // the pointers of this row are no longer needed: switch them to the next pass's row and start its loads
                const int deg_next = (node_next <= last && !DBG(1)) ? (int)((unsigned)dc.x >> 24) : 0;
                decode_ptrs(dc, node_next, hv, hu);
                row_loads<0>(hv, hu, deg_next, j, Y);
                dn = __ldg(p.row_desc + min(node + tile_rows, last));       // row of the pass after next
                row_finish<2>(p, X, tab, deg, maxdeg, node, live, j, h_thr, ee_thr, a_dst);
                decode_tabs(dc, node_next, tab, deg, maxdeg);
                node = node_next;
            };
            do_pass(0, La, Lb);
            do_pass(1, Lb, La);
            // The last float4 of every row (columns 96..99, chunk 24) does not fit 8 lanes x 3 steps: lanes 0..7 gather it
            // for the warp's 8 rows, one lane per row (the gather warps have the slack for the exposed latency).
            {
                const int R = gw * ROWS_PER_WARP + (lane & 7);
                const int nodeT = (t * 2 + (int)rank) * TM + R, nc = min(nodeT, last);
                const bool mine = lane < 8, liveT = nodeT <= last;
                const float* h24 = p.h_in + 4 * (Q - 1);
                const int4 d = __ldg(p.row_desc + nc);
                int dgT = mine && liveT ? (int)((unsigned)d.x >> 24) : 0;
                if (dgT == 255) dgT = __ldg(p.in_ptr + nc + 1) - __ldg(p.in_ptr + nc);
                const float4 hvT = ldg_f4_pred(h24 + (size_t)nc * D, mine);
                const float4 x0 = ldg_f4_pred(h24 + (size_t)(nc + (d.x & 0xFFFF) - 32768) * D, 0 < dgT);
                const float4 x1 = ldg_f4_pred(h24 + (size_t)(nc + (d.y & 0xFFFF) - 32768) * D, 1 < dgT);
                const float4 x2 = ldg_f4_pred(h24 + (size_t)(nc + (d.z & 0xFFFF) - 32768) * D, 2 < dgT);
                const float4 x3 = ldg_f4_pred(h24 + (size_t)(nc + (d.w & 0xFFFF) - 32768) * D, 3 < dgT);
                float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
                acc_edge(m, ld_f4(ee + ((d.x >> 16) & 0x3F) * D + 4 * (Q - 1)), x0);
                acc_edge(m, ld_f4(ee + ((d.y >> 16) & 0x3F) * D + 4 * (Q - 1)), x1);
                acc_edge(m, ld_f4(ee + ((d.z >> 16) & 0x3F) * D + 4 * (Q - 1)), x2);
                acc_edge(m, ld_f4(ee + ((d.w >> 16) & 0x3F) * D + 4 * (Q - 1)), x3);
                if (dgT > 4)
                {
                    const int eb = __ldg(p.in_ptr + nc);
                    for (int e = eb + 4; e < eb + dgT; e++)
                        acc_edge(m, ld_f4(ee + (int)__ldg(p.code + e) * D + 4 * (Q - 1)), __ldg(reinterpret_cast<const float4*>(h24 + (size_t)__ldg(p.src + e) * D)));
                }
                if (mine)
                {
                    uint32_t h0, l0, h1, l1;
                    split2(liveT ? m.x + hvT.x : 0.f, liveT ? m.y + hvT.y : 0.f, h0, l0);
                    split2(liveT ? m.z + hvT.z : 0.f, liveT ? m.w + hvT.w : 0.f, h1, l1);
                    const uint32_t dst = a_base + 2 * s * A_BYTES + ((Q - 1) >> 1) * LBO_A + R * 16;
                    sts_v2(dst, h0, h1);
                    sts_v2(dst + A_BYTES, l0, l1);
                }
            }
            // make the tile visible to the tensor core (async proxy) and report it to the leader CTA
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
    else
    {
        if constexpr (REGS_EPI > REGS_LAUNCH) reg_inc<REGS_EPI>(); else reg_dec<REGS_EPI>();
        // ===== epilogue warps: two per TMEM lane quadrant =====
        constexpr int PER_QUAD = EPI_WARPS / 4;
        const int quad = warp & 3, pp = warp >> 2;
        const uint32_t lane_base = tbase + ((uint32_t)(quad * 32) << 16);
        const uint32_t bar_a2a0 = mapa(smem_u32(&bar[BAR_A2A_FULL]), 0), bar_a2b0 = mapa(smem_u32(&bar[BAR_A2B_FULL]), 0);
        int it = 0;
        for (int t = pair; t < p.num_pair_tiles; t += npairs, it++)
        {
            const uint32_t ph = it & 1;
            // z = relu(acc + b1) -> bf16 hi/lo, in place (thread = row): columns [16c, 16c+8) hi, [16c+8, 16c+16) lo of
            // k-step c; the two warps of a quadrant take alternate chunks
            mbar_wait_park(&bar[BAR_G1A_DONE], ph);
            tc::fence_after_sync();
            if (tid == 0) TRACE(1, it, 0);
            if (!DBG(4)) convert_range(lane_base + TC_Z, pp, N1A / 16);
            tc::wait_st();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(bar_a2a0);
            if (tid == 0) TRACE(1, it, 1);

            mbar_wait_park(&bar[BAR_G1B_DONE], ph);
            tc::fence_after_sync();
            if (tid == 0) TRACE(1, it, 2);
            if (!DBG(4)) convert_range(lane_base + TC_Z, N1A / 16 + (pp ^ 1), N1 / 16);
            tc::wait_st();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(bar_a2b0);
            if (tid == 0) TRACE(1, it, 3);

            mbar_wait_park(&bar[BAR_G2_DONE], ph);
            tc::fence_after_sync();
            if (tid == 0) TRACE(1, it, 4);
            // h' = acc (+ relu; b2 is already in acc through the bias column k = 200 of W2): 16-lane x 256-bit TMEM loads give thread t columns 8g + 2(t%4), +1 of rows t/4 and
            // t/4 + 8, so the four lanes of a row write one full 32-byte sector per store instruction; warp pp of the
            // quadrant takes its 16-row half
            const long row_a = ((long)t * 2 + rank) * TM + quad * 32 + pp * 16 + (lane >> 2), row_b = row_a + 8;
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
                for (int g = 0; g < 4; g++)
                {
                    const int col = 8 * (g4 + g) + 2 * (lane & 3);
                    if (8 * (g4 + g) < D && col < D)
                    {
                        float2 oa = make_float2(__uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]));
                        float2 ob = make_float2(__uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3]));
                        if (p.relu_out)
                        {
                            oa = make_float2(relu_nan(oa.x), relu_nan(oa.y));
                            ob = make_float2(relu_nan(ob.x), relu_nan(ob.y));
                        }
                        // written once, read by the next launch: do not allocate in L1
                        if (row_a < p.num_nodes && !DBG(2))
                            asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(p.h_out + (size_t)row_a * D + col), "f"(oa.x), "f"(oa.y) : "memory");
                        if (row_b < p.num_nodes && !DBG(2))
                            asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(p.h_out + (size_t)row_b * D + col), "f"(ob.x), "f"(ob.y) : "memory");
                    }
                }
            };
            float dot_a = 0.f, dot_b = 0.f;
            auto dot_h = [&](int g4, const uint32_t (&r)[16]) {
#pragma unroll
                for (int g = 0; g < 4; g++)
                {
                    const int col = 8 * (g4 + g) + 2 * (lane & 3);
                    if (8 * (g4 + g) < D && col < D)
                    {
                        const float2 wv = __ldg(reinterpret_cast<const float2*>(p.head_w + col));
                        dot_a = fmaf(__uint_as_float(r[4 * g + 1]), wv.y, fmaf(__uint_as_float(r[4 * g]), wv.x, dot_a));
                        dot_b = fmaf(__uint_as_float(r[4 * g + 3]), wv.y, fmaf(__uint_as_float(r[4 * g + 2]), wv.x, dot_b));
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
                    // the four lanes of a row hold disjoint columns: add them up in a fixed order
                    dot_a += __shfl_xor_sync(FULL, dot_a, 1); dot_a += __shfl_xor_sync(FULL, dot_a, 2);
                    dot_b += __shfl_xor_sync(FULL, dot_b, 1); dot_b += __shfl_xor_sync(FULL, dot_b, 2);
                    if ((lane & 3) == 0)
                    {
                        if (row_a < p.num_nodes) p.node_dot[row_a] = dot_a;
                        if (row_b < p.num_nodes) p.node_dot[row_b] = dot_b;
                    }
                }
            }
            if (tid == 0) TRACE(1, it, 5);
        }
    }

    // both CTAs must be done with tensor memory, shared memory and each other's barriers before either leaves
    tc::fence_before_sync();
    __syncthreads();
    __syncwarp();
    cluster_sync();
    if (warp == MMA_WARP) tmem_dealloc2(tbase, TMEM_COLS);
}

}  // namespace

// bytes of one layer's weight image (both ranks) -- api.cu packs with gin_tc2_pack_layer
size_t gin_tc2_pack_bytes() { return 2 * (size_t)W_BYTES; }

// W1 [200][100], W2 [100][200] (reference "[out][in]") -> per-rank bf16 hi/lo blocks in the stationary B layout.
// Rank r holds z columns 56r..56r+55 (block 1A) and 112+48r..112+48r+47 (block 1B) of W1 and output columns
// 64r..64r+63 of W2; rows beyond the real matrix and k beyond the real K are zero.
// The biases ride along as one more K column: the A tile has a constant 1 at k = 100, so W1[z][100] = b1[z]; the extra
// row z = 200 of W1 is (0, ..., 0, 1) so that z column 200 == relu(1) == 1, and W2[o][200] = b2[o].
void gin_tc2_pack_layer(const float* w1, const float* b1, const float* w2, const float* b2, unsigned char* dst, uint16_t (*bf16_rn)(float),
                        float (*bf16_to_float)(uint16_t))
{
    std::fill(dst, dst + 2 * (size_t)W_BYTES, (unsigned char)0);
    auto put = [&](unsigned char* hi_blk, unsigned char* lo_blk, int rows, int n_local, int k, float x) {
        const size_t off = (size_t)(k / 8) * rows * 16 + (size_t)n_local * 16 + (size_t)(k % 8) * 2;
        const uint16_t hi = bf16_rn(x);
        const uint16_t lo = bf16_rn(x - bf16_to_float(hi));
        hi_blk[off] = (unsigned char)(hi & 0xFF); hi_blk[off + 1] = (unsigned char)(hi >> 8);
        lo_blk[off] = (unsigned char)(lo & 0xFF); lo_blk[off + 1] = (unsigned char)(lo >> 8);
    };
    auto w1_row = [&](unsigned char* hi_blk, unsigned char* lo_blk, int rows, int n, int z) {
        if (z < 200)
        {
            for (int k = 0; k < D; k++) put(hi_blk, lo_blk, rows, n, k, w1[(size_t)z * D + k]);
            put(hi_blk, lo_blk, rows, n, D, b1[z]);
        }
        else if (z == 200) put(hi_blk, lo_blk, rows, n, D, 1.0f);
    };
    for (int r = 0; r < 2; r++)
    {
        unsigned char* img = dst + (size_t)r * W_BYTES;
        if (G1_SINGLE)
        {
            // one block per CTA: W1_hi at offset 0, W1_lo behind it (the two half blocks of the other variant add up to the same size)
            for (int n = 0; n < N1 / 2; n++) w1_row(img + OFF_W1A_HI, img + OFF_W1A_LO + W1B_BYTES, N1 / 2, n, (N1 / 2) * r + n);
        }
        else
        {
            for (int n = 0; n < N1A / 2; n++) w1_row(img + OFF_W1A_HI, img + OFF_W1A_LO, N1A / 2, n, (N1A / 2) * r + n);
            for (int n = 0; n < N1B / 2; n++) w1_row(img + OFF_W1B_HI, img + OFF_W1B_LO, N1B / 2, n, N1A + (N1B / 2) * r + n);
        }
        for (int n = 0; n < N2 / 2; n++)
        {
            const int o = (N2 / 2) * r + n;
            if (o >= D) continue;
            for (int k = 0; k < 200; k++) put(img + OFF_W2_HI, img + OFF_W2_LO, N2 / 2, n, k, w2[(size_t)o * 200 + k]);
            put(img + OFF_W2_HI, img + OFF_W2_LO, N2 / 2, n, 200, b2[o]);
        }
    }
}

// mean over a graph's nodes of the per-node head products + bias (finalize, GIN/src/finalize.cc:36-115, with the
// Linear(100 -> 1) already applied per node by the last layer's epilogue)
__global__ void __launch_bounds__(256) gin_pool_dot_kernel(const float* __restrict__ node_dot, const int* __restrict__ node_off,
                                                           const int* __restrict__ nn, const float* __restrict__ pred_b, float* __restrict__ out,
                                                           int num_graphs)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= num_graphs) return;
    const int n = nn[g];
    const float* y = node_dot + node_off[g];
    float s = 0.f;
    for (int r = 0; r < n; r++) s += __ldg(y + r);
    out[g] = s / (float)n + __ldg(pred_b);
}

int gin_pool_dot_launch(const float* node_dot, const DeviceBatch& b, const float* pred_b, cudaStream_t s)
{
    if (b.num_graphs <= 0) return 0;
    gin_pool_dot_kernel<<<ceil_div(b.num_graphs, 256), 256, 0, s>>>(node_dot, b.node_off.as<int>(), b.nums_of_nodes.as<int>(), pred_b, b.out.as<float>(),
                                                                     b.num_graphs);
    FG_CUDA(cudaGetLastError());
    return 0;
}

unsigned long long* gin_tc2_trace_buffer = nullptr;      // set through flowgnn_b200_debug_trace (api.cu)

int gin_layer_tc2_launch(const DeviceBatch& b, const GinWeights& w, int layer, const float* h_in, float* h_out, int sm_count, cudaStream_t s,
                         const float* head_w, float* node_dot, const int4* row_desc)
{
    FG_TRY(opt_in_smem(reinterpret_cast<const void*>(&gin_layer_tc2_kernel), Smem::BYTES));
    GinTc2Params p;
    p.h_in = h_in; p.h_out = h_out;
    p.in_ptr = b.in_ptr.as<int>(); p.src = b.src.as<int>(); p.code = b.code.as<uint8_t>();
    p.row_desc = row_desc ? row_desc : b.row_desc.as<int4>();     // override: "no in-edges" descriptors = node MLP only (gin.cu)
    p.ee_comb = w.ee_comb.as<float>() + (size_t)layer * ED_COMBOS * D;
    p.wpack = w.wpack2.as<unsigned char>() + (size_t)layer * 2 * W_BYTES;
    p.num_nodes = (int)b.total_nodes;
    p.num_pair_tiles = (int)ceil_div<long>(b.total_nodes, 2 * TM);
    p.relu_out = (layer != 4);
    p.head_w = head_w; p.node_dot = node_dot;
    static const int dbg_env = [] { const char* e = std::getenv("FLOWGNN_B200_DBG"); return e ? std::atoi(e) : 0; }();
    p.dbg = dbg_env;
    p.trace = gin_tc2_trace_buffer;
    const int pairs = std::max(1, std::min(p.num_pair_tiles, sm_count / 2));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = Smem::BYTES; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    FG_CUDA(cudaLaunchKernelEx(&cfg, gin_layer_tc2_kernel, p));
    return 0;
}

}  // namespace fg

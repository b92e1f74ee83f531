// ONE kernel per message-passing step for the models whose node transform is a single dense layer (GCN: 100 -> 100,
// DGN: 200 -> 100, GAT: 64 -> 128): the aggregation is the A producer INSIDE the tensor-core GEMM kernel, so the bf16 A
// operand never makes the round trip through HBM that tcgemm.cuh's two-launch form pays (800-1,000 B per node and step).
// Reference fusion point: */src/conv_layer.cc, where message passing and node transform share one DATAFLOW region.
//
// Persistent CTA per SM, 640 threads = 20 warps, five per scheduler, so that a thread may hold 96 registers (a sixth warp on one
// scheduler would cap everybody at 80, and the gather spills below ~90: with ~200 KB of shared memory the L1 left over is too
// small for 640 threads' spill slots):
//   warps 0-3    epilogue: TMEM lane group = warp, thread = row, 16 columns per tcgen05.ld; the model's functor turns a
//                piece into stores, and preloads what it adds to the NEXT piece before it waits for this one.
//                Lane 0 of warp 0 is also the tcgen05.mma issuer (hi*hi + lo*hi + hi*lo per K = 16 step -- 3 x bf16 keeps
//                the fp32 contract, DESIGN.md --, M = 128, N = NPAD, both operands from shared memory): it issues tile t's
//                MMAs right before the warps drain tile t, which costs no time while the gather is the longer stage
//   warps 4-19   gather: 8 rows of the 128-row tile each, 16 lanes x 4 K slots = one 64-wide K chunk of a row, four rows per
//                lane; the model's functor computes the lane's values, they are split into bf16 hi + lo and stored straight
//                into the stage.  Lane 0 of the first gather warp prefetches the tile after next into L2.
// The step's weights (NCHUNK blocks, tcgemm.cuh's no-swizzle layout and host-side packer) are loaded ONCE per CTA by bulk TMA
// and stay resident.  A stage = (tile, K chunk): A hi tile | A lo tile (16 KB each) in the 128-byte-swizzle K-major layout
// (row r = 128 consecutive bytes, 16-byte unit u of row r stored at unit u ^ (r % 8)), which makes a half-warp's row store
// conflict-free: its 16 lanes cover one 128-byte line.  The two operand descriptors are independent.
#pragma once

#include "internal.cuh"
#include "tc.cuh"
#include "tcgemm.cuh"

namespace fg {
namespace tcf {

constexpr int TM = 128;
constexpr int KC = 64;
constexpr int A_TILE = TM * 128;             // one bf16 tile of a chunk: 128 rows x 128 bytes
constexpr int A_BLOCK = 2 * A_TILE;          // hi | lo
constexpr int EPI_WARPS = 4;
constexpr int GATHER_WARPS = 16;
constexpr int FIRST_GATHER_WARP = EPI_WARPS;
constexpr int NT = (FIRST_GATHER_WARP + GATHER_WARPS) * 32;      // 640

template <class Model>
struct Cfg {
    using G = tcg::Cfg<Model::NPAD>;
    static constexpr int B_BLOCK = G::B_BLOCK;
    static constexpr int W_BYTES = Model::NCHUNK * B_BLOCK;
    static constexpr int STAGES = 3;
    static constexpr int STAGE0 = W_BYTES;                        // A stages behind the resident weights
    static constexpr int ACC_COLS = 128;
    static constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;
    static constexpr int BAR = STAGE0 + STAGES * A_BLOCK;         // full[STAGES], empty[STAGES], acc_full[2], acc_empty[2], w
    static constexpr int TMEM_PTR = BAR + (2 * STAGES + 5) * 8;
    static constexpr int BYTES = TMEM_PTR + 16 + 1024;            // + slack to align the base to 1,024 bytes (swizzle atom)
    static_assert(Model::NPAD <= 128 && W_BYTES % 1024 == 0 && BYTES <= 232448, "shared memory layout");
};

#ifdef __CUDACC__

// K-major, 128-byte swizzle: 8-row groups 1,024 bytes apart; the K = 16 steps of a chunk advance the start address by 32 bytes
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// four consecutive K slots (k % 4 == 0) of row r of a chunk: bf16 hi and lo
__device__ __forceinline__ void put4(unsigned char* hi_tile, int r, int k, const float4& x)
{
    uint32_t h0, l0, h1, l1;
    tcg::split2(x.x, x.y, h0, l0);
    tcg::split2(x.z, x.w, h1, l1);
    const int off = r * 128 + ((((k >> 3) ^ (r & 7)) << 4) | ((k & 7) << 1));
    *reinterpret_cast<uint2*>(hi_tile + off) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(hi_tile + A_TILE + off) = make_uint2(l0, l1);
}

// [p, p + bytes) -> L2, any alignment and size (the bulk prefetch wants 16-byte granules: the range is widened to them)
__device__ __forceinline__ void prefetch_l2(const void* p, int bytes)
{
    if (bytes <= 0) return;
    const uintptr_t a = reinterpret_cast<uintptr_t>(p) & ~uintptr_t(15);
    const uint32_t n = (uint32_t)((reinterpret_cast<uintptr_t>(p) + bytes + 15 - a) & ~uintptr_t(15));
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(n) : "memory");
}

// lanes that share a row in the gather: 16 (4 K slots per lane, four rows per lane; default) or 8 (Model::LPR = 8: 8 K slots, two rows)
template <class Model, class = void> struct LanesPerRow { static constexpr int value = 16; };
template <class Model> struct LanesPerRow<Model, decltype((void)Model::LPR)> { static constexpr int value = Model::LPR; };
template <class Model> constexpr int lanes_per_row() { return LanesPerRow<Model>::value; }

struct Args {
    const unsigned char* wpack;      // [NCHUNK][Cfg<NPAD>::B_BLOCK] this step
    int num_nodes; int num_tiles;
};

// Model:  static constexpr int NCHUNK, NPAD;
//         static constexpr unsigned ksteps(int c)                       bit j set: K = 16 step j of chunk c holds real columns
//         struct Rows; __device__ Rows rows_begin(const int (&v)[4], const bool (&live)[4]) const    per-row state shared by the chunks
//         __device__ bool gather4(const Rows&, const int (&v)[4], const bool (&live)[4], int c, int j, float4 (&x)[4]) const
//                                                                       the lane's K slots 4j .. 4j+3 of chunk c for its four rows;
//                                                                       false: nothing to store (slots outside the issued steps)
//         __device__ void prefetch_tile(int v0, int rows) const;                       one thread, two tiles ahead: whatever the gather will read -> L2
//         struct Pre;  __device__ Pre preload(int v, bool live, int d0) const;         what the epilogue adds to columns d0 .. d0+15
//         __device__ bool row_begin(int v, bool live) const;                           after the accumulator is complete; false: skip row
//         struct RowState; __device__ RowState row_state() const;
//         __device__ void store(int v, int d0, const uint32_t (&acc)[16], const Pre&, RowState&) const;
//         __device__ void row_end(int v, const RowState&) const;
template <class Model>
__global__ void __launch_bounds__(NT, 1) fused_kernel(Args g, Model m)
{
    constexpr int NCHUNK = Model::NCHUNK, NPAD = Model::NPAD;
    using C = Cfg<Model>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + C::BAR);
    uint64_t* bar_full = bar;
    uint64_t* bar_empty = bar + C::STAGES;
    uint64_t* bar_acc_full = bar + 2 * C::STAGES;
    uint64_t* bar_acc_empty = bar + 2 * C::STAGES + 2;
    uint64_t* bar_w = bar + 2 * C::STAGES + 4;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + C::TMEM_PTR);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0)
    {
        for (int i = 0; i < C::STAGES; i++) { mbar_init(&bar_full[i], GATHER_WARPS); mbar_init(&bar_empty[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&bar_acc_full[i], 1); mbar_init(&bar_acc_empty[i], EPI_WARPS * 32); }
        mbar_init(bar_w, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(bar_w, C::W_BYTES);
        for (int c = 0; c < NCHUNK; c++) tma_load_1d(smem + c * C::B_BLOCK, g.wpack + (size_t)c * C::B_BLOCK, C::B_BLOCK, bar_w);
    }
    if (warp == 0)
    {
        tc::tmem_alloc(tmem_ptr, C::TMEM_COLS);
        tc::tmem_relinquish();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = *tmem_ptr;

    if (warp < EPI_WARPS)
    {
        const int row = warp * 32 + lane;
        const uint32_t idesc = tc::idesc_bf16(TM, NPAD);
        const uint32_t w_addr = smem_u32(smem);
        uint32_t it = 0, n = 0;
        if (tid == 0) tcg::mbar_wait_park(bar_w, 0);
        for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, it++)
        {
            const uint32_t a = it & 1;
            if (warp == 0)
            {
                if (lane == 0)
                {
                    tcg::mbar_wait_park(&bar_acc_empty[a], ((it >> 1) & 1) ^ 1);
                    tc::fence_after_sync();
                    const uint32_t d_tmem = tbase + a * C::ACC_COLS;
                    bool first = true;
#pragma unroll
                    for (int c = 0; c < NCHUNK; c++, n++)
                    {
                        const uint32_t s = n % C::STAGES;
                        tcg::mbar_wait_park(&bar_full[s], (n / C::STAGES) & 1);
                        tc::fence_after_sync();
                        const uint32_t a_addr = smem_u32(smem + C::STAGE0 + s * A_BLOCK), b_addr = w_addr + c * C::B_BLOCK;
#pragma unroll
                        for (int j = 0; j < KC / 16; j++)
                        {
                            if (!((Model::ksteps(c) >> j) & 1)) continue;
                            const uint64_t a_hi = smem_desc_sw128(a_addr + 32 * j);
                            const uint64_t a_lo = smem_desc_sw128(a_addr + A_TILE + 32 * j);
                            const uint64_t b_hi = tc::smem_desc(b_addr + 2 * j * C::G::LBO_B, C::G::LBO_B, 128);
                            const uint64_t b_lo = tc::smem_desc(b_addr + C::G::B_HALF + 2 * j * C::G::LBO_B, C::G::LBO_B, 128);
                            tcg::mma_ss(d_tmem, a_hi, b_hi, idesc, !first);
                            tcg::mma_ss(d_tmem, a_lo, b_hi, idesc, true);
                            tcg::mma_ss(d_tmem, a_hi, b_lo, idesc, true);
                            first = false;
                        }
                        tc::commit(&bar_empty[s]);                    // the stage is free once these MMAs have read it
                    }
                    tc::commit(&bar_acc_full[a]);
                }
                __syncwarp();
            }
            const int v = tile * TM + row;
            const bool live = v < g.num_nodes;
            typename Model::Pre pre = m.preload(v, live, 0);
            tcg::mbar_wait_park(&bar_acc_full[a], (it >> 1) & 1);
            tc::fence_after_sync();
            const bool on = m.row_begin(v, live);
            const uint32_t taddr = tbase + a * C::ACC_COLS + ((uint32_t)(warp * 32) << 16);
            typename Model::RowState st = m.row_state();                // what a row carries from piece to piece (GAT: its scores)
#pragma unroll 1
            for (int d0 = 0; d0 < NPAD; d0 += 16)
            {
                uint32_t acc[16];
                tc::ld16(taddr + d0, acc);
                typename Model::Pre nxt = m.preload(v, live && d0 + 16 < NPAD, d0 + 16);
                tc::wait_ld();
                if (on) m.store(v, d0, acc, pre, st);
                pre = nxt;
            }
            if (on) m.row_end(v, st);
            tc::fence_before_sync();
            tcg::mbar_arrive(&bar_acc_empty[a]);
        }
    }
    else
    {
        const int gw = warp - FIRST_GATHER_WARP;
        uint32_t n = 0;
        if (gw == 0 && lane == 0 && (int)(blockIdx.x + gridDim.x) < g.num_tiles)
            m.prefetch_tile((blockIdx.x + gridDim.x) * TM, min(TM, g.num_nodes - (int)(blockIdx.x + gridDim.x) * TM));
        for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x)
        {
            const int ahead = tile + 2 * gridDim.x;
            if (gw == 0 && lane == 0 && ahead < g.num_tiles) m.prefetch_tile(ahead * TM, min(TM, g.num_nodes - ahead * TM));
            if constexpr (lanes_per_row<Model>() == 4)
            {
                // four lanes per row, 16 K slots per lane as four pieces of four (which four is the model's choice, Model::kslot), one row
                // per lane group: a warp walks its eight rows together
                const int j = lane & 3;
                const int r0 = gw * (TM / GATHER_WARPS) + (lane >> 2);
                const int v = tile * TM + r0;
                const bool live = v < g.num_nodes;
                const typename Model::Rows rows = m.rows_begin(v, live);
#pragma unroll 1
                for (int c = 0; c < NCHUNK; c++, n++)
                {
                    const uint32_t s = n % C::STAGES;
                    tcg::mbar_wait_park(&bar_empty[s], ((n / C::STAGES) & 1) ^ 1);
                    unsigned char* hi = smem + C::STAGE0 + s * A_BLOCK;
                    float4 x[4];
                    if (m.gather1(rows, v, live, c, j, x))
                    {
#pragma unroll
                        for (int i = 0; i < 4; i++) put4(hi, r0, m.kslot(j, i), x[i]);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) tcg::mbar_arrive(&bar_full[s]);
                }
                continue;
            }
            else if constexpr (lanes_per_row<Model>() == 8)
            {
                // eight lanes per row, 8 K slots per lane (one 16-byte unit of the swizzled row), two rows per lane: the per-row work
                // (indices, predicates, attention weights) is amortised over twice the columns
                const int sub = lane >> 3, j = lane & 7;
                const int r0 = gw * (TM / GATHER_WARPS) + sub;
                const int v[2] = {tile * TM + r0, tile * TM + r0 + 4};
                const bool live[2] = {v[0] < g.num_nodes, v[1] < g.num_nodes};
                const typename Model::Rows rows = m.rows_begin(v, live);
#pragma unroll 1
                for (int c = 0; c < NCHUNK; c++, n++)
                {
                    const uint32_t s = n % C::STAGES;
                    tcg::mbar_wait_park(&bar_empty[s], ((n / C::STAGES) & 1) ^ 1);
                    unsigned char* hi = smem + C::STAGE0 + s * A_BLOCK;
                    float4 x[2][2];
                    if (m.gather2(rows, v, live, c, j, x))
                    {
#pragma unroll
                        for (int p = 0; p < 2; p++) { put4(hi, r0 + 4 * p, 8 * j, x[p][0]); put4(hi, r0 + 4 * p, 8 * j + 4, x[p][1]); }
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) tcg::mbar_arrive(&bar_full[s]);
                }
                continue;
            }
            else
            {
            const int sub = lane >> 4, j = lane & 15;
            // the warp's eight rows: lane (sub, j) owns rows r0, r0 + 2, r0 + 4, r0 + 6, walked together
            const int r0 = gw * (TM / GATHER_WARPS) + sub;
            const int v[4] = {tile * TM + r0, tile * TM + r0 + 2, tile * TM + r0 + 4, tile * TM + r0 + 6};
            const bool live[4] = {v[0] < g.num_nodes, v[1] < g.num_nodes, v[2] < g.num_nodes, v[3] < g.num_nodes};
            const typename Model::Rows rows = m.rows_begin(v, live);     // what every chunk of these rows needs (CSR positions ...): loaded once
#pragma unroll 1
            for (int c = 0; c < NCHUNK; c++, n++)
            {
                const uint32_t s = n % C::STAGES;
                tcg::mbar_wait_park(&bar_empty[s], ((n / C::STAGES) & 1) ^ 1);
                unsigned char* hi = smem + C::STAGE0 + s * A_BLOCK;
                float4 x[4];
                if (m.gather4(rows, v, live, c, j, x))
                {
#pragma unroll
                    for (int p = 0; p < 4; p++) put4(hi, r0 + 2 * p, 4 * j, x[p]);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) tcg::mbar_arrive(&bar_full[s]);
            }
            }
        }
    }

    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, C::TMEM_COLS);
}

#endif  // __CUDACC__

}  // namespace tcf
}  // namespace fg

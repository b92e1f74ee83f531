// GIN / GIN-VN layer, CTA-pair version with the tile's feature rows STAGED IN SHARED MEMORY by bulk TMA.
//
// Same math and the same pair structure as gin_tc2.cu (tcgen05 cta_group::2, half of every weight matrix per CTA,
// biases folded into the GEMMs, row descriptors).  What changes is where the gather reads from and writes to:
//   * the 128 consecutive feature rows of a CTA tile are one contiguous 51,200-byte block of HBM: a loader thread copies
//     it into shared memory with ONE cp.async.bulk per tile, double buffered, one tile ahead.  In molecular batches a
//     node's neighbours are a few rows away, so ~90 % of the source rows of a tile are in that block: the gather reads
//     them with ~30-cycle shared-memory loads instead of ~700-cycle L2 hits (the gather warps of gin_tc2.cu spend half
//     their time on the long scoreboard), and a 64-byte piece costs one shared-memory wavefront where a 128-byte global
//     piece costs two L1 wavefronts plus tag lookups.  Sources outside the block (graphs that straddle a tile
//     boundary, virtual nodes) are read from global memory through the same GENERIC pointers;
//   * the A operand of GEMM1 goes to TENSOR memory again (16-lane x 256-bit tcgen05.st, as in gin_tc.cu), which is
//     what frees the shared memory for the two stages; GEMM1 is a TS MMA (N halves 128 + 96: TS needs N % 32 == 0
//     under cta_group::2).  A is single buffered: the gather of tile t+1 waits for GEMM1 of tile t -- affordable once
//     the gather itself is short.
//
// Per CTA (896 threads):
//   warps 0-7   epilogue (as gin_tc2.cu)
//   warps 8-23  gather: warp w owns the 16 rows of TMEM lane quadrant w % 4, half (w / 4) % 2, and the k-steps of
//               parity (w - 8) / 8; thread t owns rows t/4 and t/4 + 8 of them and, in step ks, the float4 chunk
//               4 ks + t % 4 (the m16n8 fragment layout of the 16x256b store)
//   warp 24     MMA issuer (leader CTA), warp 25 loader (TMA stages)
#include "internal.cuh"
#include "layers.cuh"
#include "tc.cuh"
#include "pair.cuh"

#include <algorithm>
#include <type_traits>

namespace fg {

namespace {

using namespace pair;

constexpr int D = 100;
constexpr int Q = D / 4;
constexpr int TM = 128;                       // nodes per CTA tile (pair tile = 256)
constexpr int N1A = 128, N1B = 96, N1 = N1A + N1B;   // GEMM1 N halves (z columns), whole pair; 201 used
constexpr int N2 = 128;                       // GEMM2 N (100 used)
constexpr int K1_STEPS = 7, K2_STEPS = 13;    // K = 16 per step
constexpr int K1_CHUNKS = 13;                 // stored 8-element K chunks of W1 (k < 104; the operand's chunk 13 is zero)
constexpr int K2_CHUNKS = 26;                 // k < 200 plus the bias column k = 200
constexpr int K2A_STEPS = N1A / 16;           // GEMM2 k-steps fed by the first z half

constexpr int LBO_W1A = (N1A / 2) * 16, LBO_W1B = (N1B / 2) * 16, LBO_W2 = (N2 / 2) * 16;
constexpr int W1A_BYTES = LBO_W1A * K1_CHUNKS, W1B_BYTES = LBO_W1B * K1_CHUNKS, W2_BYTES = LBO_W2 * K2_CHUNKS;
constexpr int OFF_W1A_HI = 0, OFF_W1A_LO = OFF_W1A_HI + W1A_BYTES, OFF_W1B_HI = OFF_W1A_LO + W1A_BYTES, OFF_W1B_LO = OFF_W1B_HI + W1B_BYTES,
              OFF_W2_HI = OFF_W1B_LO + W1B_BYTES, OFF_W2_LO = OFF_W2_HI + W2_BYTES;
constexpr int W_BYTES = OFF_W2_LO + W2_BYTES;           // 99,840
constexpr int STAGE_BYTES = TM * D * 4;                 // 51,200

constexpr int EPI_WARPS = 8, GATHER_WARPS = 16;
constexpr int MMA_WARP = EPI_WARPS + GATHER_WARPS, LOAD_WARP = MMA_WARP + 1;
constexpr int NT = (MMA_WARP + 4) * 32;       // 896
constexpr int REGS_LAUNCH = 72;
constexpr int REGS_EPI = 72, REGS_MISC = 40, REGS_GATHER = 80;
static_assert(32 * (EPI_WARPS * REGS_EPI + GATHER_WARPS * REGS_GATHER + 4 * REGS_MISC) <= NT * REGS_LAUNCH, "setmaxnreg pool");

// tensor-memory columns: A operand of GEMM1 (bf16 hi | lo, 7 k-steps x 8 columns each), z, h'
constexpr uint32_t TC_A1_HI = 0, TC_A1_LO = 56, TC_Z = 128, TC_H = 384;
constexpr uint32_t TMEM_COLS = 512;
static_assert(TC_Z + N1 <= TC_H && TC_H + N2 <= TMEM_COLS, "tensor memory budget");

struct Smem {
    static constexpr int W = 0;
    static constexpr int EE = W + W_BYTES;                          // [61][100] fp32 combined edge-embedding rows; row 60 = sentinel
    static constexpr int STAGE = EE + (ED_COMBOS + 1) * D * 4;      // [2][128][100] fp32 feature rows of the tile
    static constexpr int BAR = STAGE + 2 * STAGE_BYTES;
    static constexpr int TMEM_PTR = BAR + 16 * 8;
    static constexpr int BYTES = TMEM_PTR + 16;
};
static_assert(Smem::EE % 16 == 0 && Smem::STAGE % 16 == 0 && Smem::BAR % 8 == 0, "alignment");
static_assert(Smem::BYTES <= 232448, "shared memory budget");

enum { BAR_W = 0, BAR_STAGE_FULL /* 2 */, BAR_STAGE_FREE = BAR_STAGE_FULL + 2 /* 2 */, BAR_A1_FULL = BAR_STAGE_FREE + 2, BAR_G1A_DONE, BAR_G1B_DONE,
       BAR_A2A_FULL, BAR_A2B_FULL, BAR_G2_DONE };

struct GinTc3Params {
    const float* h_in; float* h_out;
    const int* in_ptr; const int* src; const uint8_t* code;
    const int4* row_desc;            // [N] first four in-edges of every node, packed (prep.cu)
    const float* ee_comb;            // [60][100] this layer
    const unsigned char* wpack;      // [2 ranks][W_BYTES] this layer
    int num_nodes; int num_pair_tiles; int relu_out;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 16 lanes x 256 bit TMEM store: thread t supplies columns 2(t%4), 2(t%4)+1 of lane t/4 (r0, r1) and of lane t/4 + 8 (r2, r3)
__device__ __forceinline__ void st_16x256(uint32_t taddr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3)
{
    asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}

// predicated 16-byte load through a GENERIC pointer (shared-memory stage or global memory): absent slots read as 0
__device__ __forceinline__ float4 ld_f4_pred(const float* ptr, bool on)
{
    float4 v;
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "mov.f32 %0, 0f00000000;\n\tmov.f32 %1, 0f00000000;\n\tmov.f32 %2, 0f00000000;\n\tmov.f32 %3, 0f00000000;\n\t"
        "@p ld.v4.f32 {%0,%1,%2,%3}, [%4];\n\t}"
        : "=&f"(v.x), "=&f"(v.y), "=&f"(v.z), "=&f"(v.w)
        : "l"(ptr), "r"((int)on));
    return v;
}

// One destination row as seen by one of its 4 threads: generic pointers to this thread's chunk (step 0) of the row
// itself and of the source rows of its first four in-edges, the shared-memory addresses of their edge-embedding rows.
struct RowEdges {
    const float* hv;
    const float* hu[4];
    uint32_t t[4];
    int deg;
    int node;
};

// a_v[4q .. 4q+3], q = 4 KS + qsub: sum over in-edges (CSR order) relu(h_u + EE[attr]) + h_v
template <int KS>
__device__ __forceinline__ float4 row_step(const GinTc3Params& p, const RowEdges& r, int qsub, bool live, uint32_t ee_thr, const float* h_thr,
                                           const float* stage_thr, int n0, int rows)
{
    constexpr int OFF = 16 * KS;                             // floats
    const bool on = (KS < K1_STEPS - 1) || (qsub == 0);      // the last k-step only holds chunk 24 (sub-chunk 0)
    const float4 hv = ld_f4_pred(r.hv + OFF, on);
    float4 hu[4];
#pragma unroll
    for (int s = 0; s < 4; s++) hu[s] = ld_f4_pred(r.hu[s] + OFF, on && s < r.deg);
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s = 0; s < 4; s++) acc_edge(m, lds_f4(r.t[s] + 4 * OFF), hu[s]);
    if (r.deg > 4 && on)
    {
        // long in-edge lists (virtual nodes, kNN graphs): continue from the CSR arrays (divergent per row: no shuffles here)
        const int eb = __ldg(p.in_ptr + r.node), e_end = eb + r.deg;
        for (int e = eb + 4; e < e_end; e++)
        {
            const int u = __ldg(p.src + e), c = __ldg(p.code + e);
            const int ru = u - n0;
            const float* pu = ((unsigned)ru < (unsigned)rows) ? stage_thr + ru * D : h_thr + (size_t)u * D;
            acc_edge(m, lds_f4(ee_thr + c * (D * 4) + 4 * OFF), ld_f4_pred(pu + OFF, true));
        }
    }
    float4 a = make_float4(m.x + hv.x, m.y + hv.y, m.z + hv.z, m.w + hv.w);
    if (!live || !on) a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (KS == K1_STEPS - 1 && qsub == 1) a.x = 1.0f;         // the bias column: k = 100 is a constant 1 for every row
    return a;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NT, 1) gin_layer_tc3_kernel(GinTc3Params p)
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
            mbar_init(&bar[BAR_STAGE_FULL + i], 1);
            mbar_init(&bar[BAR_STAGE_FREE + i], GATHER_WARPS);
        }
        mbar_init(&bar[BAR_A1_FULL], 2 * GATHER_WARPS);
        mbar_init(&bar[BAR_G1A_DONE], 1);
        mbar_init(&bar[BAR_G1B_DONE], 1);
        mbar_init(&bar[BAR_A2A_FULL], 2 * EPI_WARPS);
        mbar_init(&bar[BAR_A2B_FULL], 2 * EPI_WARPS);
        mbar_init(&bar[BAR_G2_DONE], 1);
        fence_mbar_init();
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
    tc::fence_before_sync();
    __syncthreads();
    cluster_sync();
    tc::fence_after_sync();
    const uint32_t tbase = *tmem_ptr;

    if (warp >= MMA_WARP)
    {
        reg_dec<REGS_MISC>();
        if (warp == MMA_WARP)
        {
            // ===== MMA issuer: one thread of the leader CTA =====
            if (rank == 0 && lane == 0)
            {
                const uint32_t w_addr = smem_u32(smem + Smem::W);
                const uint32_t idesc1a = tc::idesc_bf16(2 * TM, N1A), idesc1b = tc::idesc_bf16(2 * TM, N1B), idesc2 = tc::idesc_bf16(2 * TM, N2);
                int it = 0;
                for (int t = pair; t < p.num_pair_tiles; t += npairs, it++)
                {
                    const uint32_t ph = it & 1;
                    mbar_wait_park(&bar[BAR_A1_FULL], ph);
                    tc::fence_after_sync();
                    // GEMM1 (A from tensor memory), N half a (z columns 0..127) then half b (128..223)
#pragma unroll
                    for (int nh = 0; nh < 2; nh++)
                    {
                        bool acc = false;
                        const uint32_t lbo_b = nh ? LBO_W1B : LBO_W1A;
#pragma unroll
                        for (int prod = 0; prod < 3; prod++)
                        {
                            const uint32_t a_col = tbase + (prod == 1 ? TC_A1_LO : TC_A1_HI);
                            const uint32_t b_addr = w_addr + (nh ? (prod == 2 ? OFF_W1B_LO : OFF_W1B_HI) : (prod == 2 ? OFF_W1A_LO : OFF_W1A_HI));
#pragma unroll
                            for (int j = 0; j < K1_STEPS; j++)
                            {
                                mma_ts2(tbase + TC_Z + (nh ? N1A : 0), a_col + 8 * j, tc::smem_desc(b_addr + 2 * j * lbo_b, lbo_b, 128),
                                        nh ? idesc1b : idesc1a, acc);
                                acc = true;
                            }
                        }
                        commit2(&bar[nh ? BAR_G1B_DONE : BAR_G1A_DONE]);
                    }
                    // GEMM2, K half a (operand columns converted from z half a) then half b
                    bool acc = false;
#pragma unroll
                    for (int kh = 0; kh < 2; kh++)
                    {
                        mbar_wait_park(&bar[kh ? BAR_A2B_FULL : BAR_A2A_FULL], ph);
                        tc::fence_after_sync();
#pragma unroll
                        for (int prod = 0; prod < 3; prod++)
                        {
                            const uint32_t a_col = tbase + TC_Z + (prod == 1 ? 8 : 0);
                            const uint32_t b_addr = w_addr + (prod == 2 ? OFF_W2_LO : OFF_W2_HI);
#pragma unroll
                            for (int j = (kh ? K2A_STEPS : 0); j < (kh ? K2_STEPS : K2A_STEPS); j++)
                            {
                                mma_ts2(tbase + TC_H, a_col + 16 * j, tc::smem_desc(b_addr + 2 * j * LBO_W2, LBO_W2, 128), idesc2, acc);
                                acc = true;
                            }
                        }
                    }
                    commit2(&bar[BAR_G2_DONE]);
                }
            }
        }
        else if (warp == LOAD_WARP)
        {
            // ===== loader: one bulk copy per tile, double buffered, as far ahead as the stages allow =====
            if (lane == 0)
            {
                int it = 0;
                for (int t = pair; t < p.num_pair_tiles; t += npairs, it++)
                {
                    const int s = it & 1;
                    const long n0 = ((long)t * 2 + rank) * TM;
                    const int rows = (int)max(0L, min((long)TM, (long)p.num_nodes - n0));
                    if (it >= 2) mbar_wait_park(&bar[BAR_STAGE_FREE + s], ((it >> 1) - 1) & 1);
                    if (rows > 0)
                    {
                        mbar_arrive_expect_tx(&bar[BAR_STAGE_FULL + s], rows * D * 4);
                        tma_load_1d(smem + Smem::STAGE + s * STAGE_BYTES, p.h_in + n0 * D, rows * D * 4, &bar[BAR_STAGE_FULL + s]);
                    }
                    else mbar_arrive(&bar[BAR_STAGE_FULL + s]);
                }
            }
        }
    }
    else if (warp >= EPI_WARPS)
    {
        if constexpr (REGS_GATHER > REGS_LAUNCH) reg_inc<REGS_GATHER>(); else reg_dec<REGS_GATHER>();
        // ===== gather warps: build the A operand of GEMM1 in tensor memory from the staged tile =====
        const int gw = warp - EPI_WARPS;                    // 0..15
        const int quad = warp & 3, half = (gw >> 2) & 1, par = gw >> 3;
        const uint32_t taddr = tbase + ((uint32_t)(quad * 32 + half * 16) << 16);
        const int r_a = quad * 32 + half * 16 + (lane >> 2), r_b = r_a + 8;
        const int qsub = lane & 3;
        const uint32_t ee_thr = smem_u32(ee) + 16 * qsub;
        const float* h_thr = p.h_in + 4 * qsub;
        const uint32_t bar_full0 = mapa(smem_u32(&bar[BAR_A1_FULL]), 0);
        const int last = p.num_nodes - 1;

        int it = 0;
        for (int t = pair; t < p.num_pair_tiles; t += npairs, it++)
        {
            const int s = it & 1;
            const int n0 = (t * 2 + (int)rank) * TM;
            const int rows = max(0, min(TM, p.num_nodes - n0));
            const float* stage_thr = reinterpret_cast<const float*>(smem + Smem::STAGE + s * STAGE_BYTES) + 4 * qsub;   // generic pointer
            // decode: sources inside the staged block are read from shared memory, the others from global memory
            auto decode = [&](const int4& d, int r, RowEdges& re) {
                re.node = min(n0 + r, last);
                re.deg = r < rows ? (int)((unsigned)d.x >> 24) : 0;
                if (re.deg == 255) re.deg = __ldg(p.in_ptr + re.node + 1) - __ldg(p.in_ptr + re.node);
                re.hv = stage_thr + min(r, max(rows - 1, 0)) * D;
                const int dq[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                for (int q = 0; q < 4; q++)
                {
                    const int delta = (dq[q] & 0xFFFF) - 32768;
                    const int ru = r + delta;
                    re.hu[q] = ((unsigned)ru < (unsigned)rows) ? stage_thr + ru * D : h_thr + (size_t)(re.node + delta) * D;
                    re.t[q] = ee_thr + ((dq[q] >> 16) & 0x3F) * (D * 4);
                }
            };
            mbar_wait_park(&bar[BAR_STAGE_FULL + s], (it >> 1) & 1);
            // Row a first, all of this warp's k-steps, keeping its bf16 hi/lo words; then row b, storing both rows of a
            // k-step with one 16x256b instruction.  One row's pointers at a time: the register budget is 80.
            uint32_t sa_hi[4][2], sa_lo[4][2];
            {
                RowEdges re;
                decode(__ldg(p.row_desc + min(n0 + r_a, last)), r_a, re);
                const bool live = r_a < rows;
                auto step_a = [&](auto ks_tag, int i) {
                    constexpr int KS = decltype(ks_tag)::value;
                    const float4 a = row_step<KS>(p, re, qsub, live, ee_thr, h_thr, stage_thr, n0, rows);
                    split2(a.x, a.y, sa_hi[i][0], sa_lo[i][0]);
                    split2(a.z, a.w, sa_hi[i][1], sa_lo[i][1]);
                };
                if (par == 0)
                {
                    step_a(std::integral_constant<int, 0>{}, 0);
                    step_a(std::integral_constant<int, 2>{}, 1);
                    step_a(std::integral_constant<int, 4>{}, 2);
                    step_a(std::integral_constant<int, 6>{}, 3);
                }
                else
                {
                    step_a(std::integral_constant<int, 1>{}, 0);
                    step_a(std::integral_constant<int, 3>{}, 1);
                    step_a(std::integral_constant<int, 5>{}, 2);
                }
            }
            {
                RowEdges re;
                decode(__ldg(p.row_desc + min(n0 + r_b, last)), r_b, re);
                const bool live = r_b < rows;
                if (it > 0)
                {
                    // the A operand is single buffered: GEMM1 of the previous tile must have consumed it
                    mbar_wait_park(&bar[BAR_G1B_DONE], (it - 1) & 1);
                    tc::fence_after_sync();
                }
                auto step_b = [&](auto ks_tag, int i) {
                    constexpr int KS = decltype(ks_tag)::value;
                    const float4 b = row_step<KS>(p, re, qsub, live, ee_thr, h_thr, stage_thr, n0, rows);
                    uint32_t hb0, lb0, hb1, lb1;
                    split2(b.x, b.y, hb0, lb0);
                    split2(b.z, b.w, hb1, lb1);
                    __syncwarp();
                    st_16x256(taddr + TC_A1_HI + 8 * KS, sa_hi[i][0], sa_hi[i][1], hb0, hb1);
                    st_16x256(taddr + TC_A1_LO + 8 * KS, sa_lo[i][0], sa_lo[i][1], lb0, lb1);
                };
                if (par == 0)
                {
                    step_b(std::integral_constant<int, 0>{}, 0);
                    step_b(std::integral_constant<int, 2>{}, 1);
                    step_b(std::integral_constant<int, 4>{}, 2);
                    step_b(std::integral_constant<int, 6>{}, 3);
                }
                else
                {
                    step_b(std::integral_constant<int, 1>{}, 0);
                    step_b(std::integral_constant<int, 3>{}, 1);
                    step_b(std::integral_constant<int, 5>{}, 2);
                }
            }
            tc::wait_st();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0)
            {
                if (it == 0 && gw == 0) mbar_wait_park(&bar[BAR_W], 0);      // this CTA's weights have landed
                mbar_arrive(&bar[BAR_STAGE_FREE + s]);
                mbar_arrive_cluster(bar_full0);
            }
        }
    }
    else
    {
        if constexpr (REGS_EPI > REGS_LAUNCH) reg_inc<REGS_EPI>(); else reg_dec<REGS_EPI>();
        // ===== epilogue warps: two per TMEM lane quadrant =====
        const int quad = warp & 3, pp = warp >> 2;
        const uint32_t lane_base = tbase + ((uint32_t)(quad * 32) << 16);
        const uint32_t bar_a2a0 = mapa(smem_u32(&bar[BAR_A2A_FULL]), 0), bar_a2b0 = mapa(smem_u32(&bar[BAR_A2B_FULL]), 0);
        int it = 0;
        for (int t = pair; t < p.num_pair_tiles; t += npairs, it++)
        {
            const uint32_t ph = it & 1;
            // z = relu(acc) -> bf16 hi/lo, in place (thread = row); the two warps of a quadrant take alternate 16-column chunks
            mbar_wait_park(&bar[BAR_G1A_DONE], ph);
            tc::fence_after_sync();
            convert_range(lane_base + TC_Z, pp, K2A_STEPS);
            tc::wait_st();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(bar_a2a0);

            mbar_wait_park(&bar[BAR_G1B_DONE], ph);
            tc::fence_after_sync();
            convert_range(lane_base + TC_Z, K2A_STEPS + pp, K2_STEPS);
            tc::wait_st();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(bar_a2b0);

            mbar_wait_park(&bar[BAR_G2_DONE], ph);
            tc::fence_after_sync();
            // h' = acc (+ relu): 16-lane x 256-bit TMEM loads, four lanes of a row write one full 32-byte sector per store
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
                        if (row_a < p.num_nodes) *reinterpret_cast<float2*>(p.h_out + (size_t)row_a * D + col) = oa;
                        if (row_b < p.num_nodes) *reinterpret_cast<float2*>(p.h_out + (size_t)row_b * D + col) = ob;
                    }
                }
            };
            {
                uint32_t r0[16], r1[16];
                ld_h(0, r0);
                tc::wait_ld(); ld_h(4, r1); st_h(0, r0);
                tc::wait_ld(); ld_h(8, r0); st_h(4, r1);
                tc::wait_ld(); ld_h(12, r1); st_h(8, r0);
                tc::wait_ld(); st_h(12, r1);
            }
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

size_t gin_tc3_pack_bytes() { return 2 * (size_t)W_BYTES; }

// As gin_tc2_pack_layer, for N halves 128 + 96: rank r holds z columns 64r..64r+63 (block 1A) and 128+48r..128+48r+47
// (block 1B) of W1 and output columns 64r..64r+63 of W2; bias column k = 100 in W1, the row z = 200 = (0,...,0,1),
// bias column k = 200 in W2.
void gin_tc3_pack_layer(const float* w1, const float* b1, const float* w2, const float* b2, unsigned char* dst, uint16_t (*bf16_rn)(float),
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
        for (int n = 0; n < N1A / 2; n++) w1_row(img + OFF_W1A_HI, img + OFF_W1A_LO, N1A / 2, n, (N1A / 2) * r + n);
        for (int n = 0; n < N1B / 2; n++) w1_row(img + OFF_W1B_HI, img + OFF_W1B_LO, N1B / 2, n, N1A + (N1B / 2) * r + n);
        for (int n = 0; n < N2 / 2; n++)
        {
            const int o = (N2 / 2) * r + n;
            if (o >= D) continue;
            for (int k = 0; k < 200; k++) put(img + OFF_W2_HI, img + OFF_W2_LO, N2 / 2, n, k, w2[(size_t)o * 200 + k]);
            put(img + OFF_W2_HI, img + OFF_W2_LO, N2 / 2, n, 200, b2[o]);
        }
    }
}

int gin_layer_tc3_launch(const DeviceBatch& b, const GinWeights& w, int layer, const float* h_in, float* h_out, int sm_count, cudaStream_t s)
{
    FG_TRY(opt_in_smem(reinterpret_cast<const void*>(&gin_layer_tc3_kernel), Smem::BYTES));
    GinTc3Params p;
    p.h_in = h_in; p.h_out = h_out;
    p.in_ptr = b.in_ptr.as<int>(); p.src = b.src.as<int>(); p.code = b.code.as<uint8_t>();
    p.row_desc = b.row_desc.as<int4>();
    p.ee_comb = w.ee_comb.as<float>() + (size_t)layer * ED_COMBOS * D;
    p.wpack = w.wpack3.as<unsigned char>() + (size_t)layer * 2 * W_BYTES;
    p.num_nodes = (int)b.total_nodes;
    p.num_pair_tiles = (int)ceil_div<long>(b.total_nodes, 2 * TM);
    p.relu_out = (layer != 4);
    const int pairs = std::max(1, std::min(p.num_pair_tiles, sm_count / 2));
    gin_layer_tc3_kernel<<<2 * pairs, NT, Smem::BYTES, s>>>(p);
    FG_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace fg

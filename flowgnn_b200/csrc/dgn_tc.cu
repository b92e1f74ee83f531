// DGN layer with the node transform on the B200 tensor cores (option "dgn_tc"; dgn.cu is the FFMA version and documents
// the reference pipeline, DGN/src/message_passing.cc:121-153, node_embedding.cc:106-183).
//
// Per layer three launches (tcgemm.cuh):
//   dgn_aggregate_kernel        m0 = sum h_u, m1 = sum h_u eig_w_uv;  a1 = m0/outdeg(v), a2 = |(m1 - B_v h_v)/A_v|  -> bf16 hi/lo
//                               A blocks, K = [a1 | pad | a2 | pad] = 2 x 128; rows with a non-finite value are flagged
//   tcg::gemm_kernel<4, 112>    acc = [a1 | a2] W^T (hi*hi + lo*hi + hi*lo on tcgen05); epilogue h' = h + relu(acc + b)
//   dgn_exact_rows_kernel       flagged rows (out-degree 0: a1 = m0/0) in fp32 with the FFMA kernel's evaluation order, so
//                               that the result is non-finite exactly where the reference's is -- a bf16
//                               split of inf is inf + NaN.
// The expressions of the aggregate kernel are those of dgn.cu, in the same order.
#include "internal.cuh"
#include "layers.cuh"
#include "tcgemm.cuh"
#include "fused_tc.cuh"

#include <algorithm>
#include <cstdlib>

namespace fg {

namespace {

constexpr int D = 100;
constexpr int DP = 104;
constexpr int KA = 2 * D;
constexpr int KPART = 128;                   // padded width of one part of the A row
constexpr int NCHUNK = 4;                    // K = 256
constexpr int NPAD = 112;
constexpr int G8 = 13;                       // column groups of eight with real columns (the last one: 96..99)
constexpr int AG_WARPS = 8;

struct DgnAggParams {
    const float* h_in;
    const int* in_ptr; const int* src; const float* eig_w; const int* out_deg;
    const float* abssum; const float* wsum;
    unsigned char* apack; unsigned char* nonfinite;
    long num_nodes;
};

// a1, a2 of columns c..c+3 of node v, in-edges in CSR order (the loop of dgn.cu)
__device__ __forceinline__ void dgn_aggregate4(const DgnAggParams& p, long v, int c, int eb, int ee, float deg, float abssum, float wsum,
                                               float4& a1, float4& a2)
{
    float4 m0 = make_float4(0.f, 0.f, 0.f, 0.f), m1 = m0;
    for (int e = eb; e < ee; e++)
    {
        const int u = __ldg(p.src + e);
        const float w = __ldg(p.eig_w + e);
        const float4 hu = ldg_f4(p.h_in + (size_t)u * D + c);
        m0.x += hu.x; m0.y += hu.y; m0.z += hu.z; m0.w += hu.w;
        m1.x += hu.x * w; m1.y += hu.y * w; m1.z += hu.z * w; m1.w += hu.w * w;
    }
    const float4 hv = ldg_f4(p.h_in + (size_t)v * D + c);
    a1 = make_float4(m0.x / deg, m0.y / deg, m0.z / deg, m0.w / deg);
    a2 = make_float4(fabsf((m1.x - wsum * hv.x) / abssum), fabsf((m1.y - wsum * hv.y) / abssum), fabsf((m1.z - wsum * hv.z) / abssum),
                     fabsf((m1.w - wsum * hv.w) / abssum));
}

__device__ __forceinline__ bool finite4(const float4& x) { return isfinite(x.x) && isfinite(x.y) && isfinite(x.z) && isfinite(x.w); }

__global__ void __launch_bounds__(AG_WARPS * 32) dgn_aggregate_kernel(DgnAggParams p)
{
    const int lane = threadIdx.x & 31;
    const long warp = blockIdx.x * (long)AG_WARPS + (threadIdx.x >> 5), nwarps = (long)gridDim.x * AG_WARPS;
    const long items = ((p.num_nodes + 31) / 32) * G8;
    for (long item = warp; item < items; item += nwarps)
    {
        const long rb = item / G8;
        const int g8 = (int)(item - rb * G8);
        const long v = rb * 32 + lane;
        if (v >= p.num_nodes) continue;
        const int c0 = 8 * g8;
        const bool two = c0 + 4 < D;
        const int eb = __ldg(p.in_ptr + v), ee = __ldg(p.in_ptr + v + 1);
        const float deg = (float)__ldg(p.out_deg + v);
        float abssum = __ldg(p.abssum + v);
        if (abssum == 0.0f) abssum = 0.0001220703125f;           // ap_fixed_epsilon of <16,3> = 2^-13
        const float wsum = __ldg(p.wsum + v);
        float x1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, x2[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float4 a1, a2;
        dgn_aggregate4(p, v, c0, eb, ee, deg, abssum, wsum, a1, a2);
        bool fin = finite4(a1) && finite4(a2);
        x1[0] = a1.x; x1[1] = a1.y; x1[2] = a1.z; x1[3] = a1.w;
        x2[0] = a2.x; x2[1] = a2.y; x2[2] = a2.z; x2[3] = a2.w;
        if (two)
        {
            dgn_aggregate4(p, v, c0 + 4, eb, ee, deg, abssum, wsum, a1, a2);
            fin = fin && finite4(a1) && finite4(a2);
            x1[4] = a1.x; x1[5] = a1.y; x1[6] = a1.z; x1[7] = a1.w;
            x2[4] = a2.x; x2[5] = a2.y; x2[6] = a2.z; x2[7] = a2.w;
        }
        if (!fin) p.nonfinite[v] = 1;
        tcg::put8<NCHUNK>(p.apack, v, c0, x1);
        tcg::put8<NCHUNK>(p.apack, v, KPART + c0, x2);
        if (g8 == G8 - 1)
            for (int k0 = 8 * G8; k0 < KPART; k0 += 8)
            {
                tcg::put8_zero<NCHUNK>(p.apack, v, k0);                       // K padding of both parts
                tcg::put8_zero<NCHUNK>(p.apack, v, KPART + k0);
            }
    }
}

// h'[v][d] = h[v][d] + relu(acc[d] + b[d]); flagged rows are written by dgn_exact_rows_kernel
struct DgnEpi {
    const float* b; const float* h_in; float* h_out; const unsigned char* nonfinite;
    struct State { bool skip; };
    __device__ State begin(int v, bool live) const { return State{!live || __ldg(nonfinite + v) != 0}; }
    __device__ void store(const State& st, int v, int d0, const uint32_t (&acc)[16]) const
    {
        if (st.skip) return;
#pragma unroll
        for (int j = 0; j < 16; j += 4)
            if (d0 + j < D)
            {
                const float4 bb = ldg_f4(b + d0 + j);
                const float4 hv = ldg_f4(h_in + (size_t)v * D + d0 + j);
                stg_f4_stream(h_out + (size_t)v * D + d0 + j,
                              make_float4(hv.x + relu_f(__uint_as_float(acc[j]) + bb.x), hv.y + relu_f(__uint_as_float(acc[j + 1]) + bb.y),
                                          hv.z + relu_f(__uint_as_float(acc[j + 2]) + bb.z), hv.w + relu_f(__uint_as_float(acc[j + 3]) + bb.w)));
            }
    }
};

// flagged rows: [a1 | a2] in fp32, acc = sum_k a_k Wt[k][n] with k ascending (the FFMA kernel's order), a warp per row
constexpr int EX_WARPS = 8;
__global__ void __launch_bounds__(EX_WARPS * 32) dgn_exact_rows_kernel(DgnAggParams p, const float* __restrict__ wt, const float* __restrict__ b,
                                                                       float* __restrict__ h_out)
{
    __shared__ __align__(16) float agg[EX_WARPS][KA];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int warp = blockIdx.x * EX_WARPS + wid, nwarps = gridDim.x * EX_WARPS;
    float* a = agg[wid];
    for (long v0 = (long)warp * 32; v0 < p.num_nodes; v0 += (long)nwarps * 32)
    {
        const long vl = v0 + lane;
        unsigned todo = __ballot_sync(0xFFFFFFFFu, vl < p.num_nodes && __ldg(p.nonfinite + vl) != 0);
        while (todo)
        {
            const int i = __ffs(todo) - 1;
            todo &= todo - 1;
            const long v = v0 + i;
            __syncwarp();
            if (lane < D / 4)
            {
                const int eb = __ldg(p.in_ptr + v), ee = __ldg(p.in_ptr + v + 1);
                const float deg = (float)__ldg(p.out_deg + v);
                float abssum = __ldg(p.abssum + v);
                if (abssum == 0.0f) abssum = 0.0001220703125f;
                const float wsum = __ldg(p.wsum + v);
                float4 a1, a2;
                dgn_aggregate4(p, v, 4 * lane, eb, ee, deg, abssum, wsum, a1, a2);
                st_f4(a + 4 * lane, a1);
                st_f4(a + D + 4 * lane, a2);
            }
            __syncwarp();
            for (int n = lane; n < D; n += 32)
            {
                float acc = 0.f;
                for (int k = 0; k < KA; k++) acc = fmaf(a[k], __ldg(wt + (size_t)k * DP + n), acc);
                h_out[(size_t)v * D + n] = __ldg(p.h_in + (size_t)v * D + n) + relu_f(acc + __ldg(b + n));
            }
        }
    }
}

// ---- option dgn_fused (default): the layer as ONE launch, the aggregation as the A producer inside the GEMM kernel (fused_tc.cuh) ----
// K layout: chunk c = [a1 columns 32c .. 32c+31 | a2 columns 32c .. 32c+31], so that one chunk needs 128 bytes of every
// neighbour row and both halves of a lane group walk the same in-edges.  Chunk 3 holds columns 96..99 only: K steps 0 (a1) and 2 (a2).
struct DgnFused {
    static constexpr int NCHUNK = 4, NPAD = fg::NPAD;
    static constexpr unsigned ksteps(int c) { return c < 3 ? 0xFu : 0x5u; }
    DgnAggParams p;
    const float* b; float* h_out;

    // Four lanes per row (fused_tc.cuh, LPR = 4): lanes 0-1 of a group accumulate m0 = sum h_u (-> a1) for columns 32c + 16j .. + 15,
    // lanes 2-3 m1 = sum h_u eig_w (-> a2) for the same columns, over the SAME in-edge walk; a warp walks its eight rows together.
    static constexpr int LPR = 4;
    static __device__ __forceinline__ int kslot(int j, int i) { return 16 * j + 4 * i; }
    struct Rows { int e0, end; };
    __device__ __forceinline__ Rows rows_begin(int v, bool live) const
    {
        Rows r;
        r.e0 = live ? __ldg(p.in_ptr + v) : 0;
        r.end = live ? __ldg(p.in_ptr + v + 1) : 0;
        return r;
    }
    __device__ __forceinline__ bool gather1(const Rows& rows, int v, bool live, int c, int j, float4 (&x)[4]) const
    {
        const int part = j >> 1, col = 32 * c + 16 * (j & 1);
        if (c == 3 && (j & 1)) return false;                       // chunk 3: columns 96..111 only (K steps 0 and 2)
#pragma unroll
        for (int i = 0; i < 4; i++) x[i] = make_float4(0.f, 0.f, 0.f, 0.f);      // K padding and rows past the batch hold zeros
        if (!live) return true;
        const int np = (col + 16 <= D) ? 4 : (D - col) / 4;         // real float4 pieces of this lane: 4, or 1 for columns 96..99
        // what the end of the walk needs, requested before it starts
        float fin0, fin1 = 0.f;                                      // part 0: out-degree; part 1: A_v = sum |eig_w|, B_v = sum eig_w
        float4 hv[4], m[4];
#pragma unroll
        for (int i = 0; i < 4; i++) hv[i] = m[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (part)
        {
            fin0 = __ldg(p.abssum + v); fin1 = __ldg(p.wsum + v);
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (i < np) hv[i] = ldg_f4(p.h_in + (size_t)v * D + col + 4 * i);
        }
        else fin0 = (float)__ldg(p.out_deg + v);
        int e = rows.e0, u = 0;
        float w = 1.0f;
        if (e < rows.end) { u = __ldg(p.src + e); if (part) w = __ldg(p.eig_w + e); }
        // in-edges in CSR order (the loop of dgn_aggregate4 above); the source and weight of the NEXT in-edge are requested together
        // with the CURRENT neighbour row
        while (e < rows.end)
        {
            float4 hu[4];
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (i < np) hu[i] = ldg_f4(p.h_in + (size_t)u * D + col + 4 * i);
            int un = 0;
            float wn = 1.0f;
            if (e + 1 < rows.end) { un = __ldg(p.src + e + 1); if (part) wn = __ldg(p.eig_w + e + 1); }
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (i < np)
                {
                    if (part) { m[i].x += hu[i].x * w; m[i].y += hu[i].y * w; m[i].z += hu[i].z * w; m[i].w += hu[i].w * w; }
                    else { m[i].x += hu[i].x; m[i].y += hu[i].y; m[i].z += hu[i].z; m[i].w += hu[i].w; }
                }
            e++;
            u = un; w = wn;
        }
        // the reference divides (m0 / deg, (m1 - B h) / A); here one IEEE reciprocal per row and a multiplication per column:
        // <= 1 ulp from the quotient (the bar is 1e-4), and deg = 0 still gives inf * m0 = +-inf or NaN exactly where m0 / 0 does
        const float r = 1.0f / (part ? (fin0 == 0.0f ? 0.0001220703125f : fin0) : fin0);      // ap_fixed_epsilon of <16,3> = 2^-13
        float probe = 0.f;
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (i < np)
            {
                float4 a;
                if (part)
                    a = make_float4(fabsf((m[i].x - fin1 * hv[i].x) * r), fabsf((m[i].y - fin1 * hv[i].y) * r), fabsf((m[i].z - fin1 * hv[i].z) * r),
                                    fabsf((m[i].w - fin1 * hv[i].w) * r));
                else
                    a = make_float4(m[i].x * r, m[i].y * r, m[i].z * r, m[i].w * r);
                probe += (a.x * 0.0f + a.y * 0.0f) + (a.z * 0.0f + a.w * 0.0f);        // x * 0 is 0 for finite x and NaN otherwise
                x[i] = a;
            }
        if (probe != 0.0f) p.nonfinite[v] = 1;                      // before this warp's arrival on the stage: the epilogue sees it
        return true;
    }
    __device__ __forceinline__ void prefetch_tile(int v0, int rows) const
    {
        tcf::prefetch_l2(p.h_in + (size_t)v0 * D, rows * D * 4);
        tcf::prefetch_l2(p.in_ptr + v0, rows * 4 + 4);
        tcf::prefetch_l2(p.out_deg + v0, rows * 4);
        tcf::prefetch_l2(p.abssum + v0, rows * 4);
        tcf::prefetch_l2(p.wsum + v0, rows * 4);
        const int e0 = __ldg(p.in_ptr + v0), e1 = __ldg(p.in_ptr + v0 + rows);
        tcf::prefetch_l2(p.src + e0, (e1 - e0) * 4);
        tcf::prefetch_l2(p.eig_w + e0, (e1 - e0) * 4);
    }
    // the residual h[v][d0 .. d0+15], requested one accumulator piece ahead
    struct Pre { float4 h[4]; };
    __device__ __forceinline__ Pre preload(int v, bool live, int d0) const
    {
        Pre r;
#pragma unroll
        for (int j = 0; j < 4; j++) r.h[j] = (live && d0 + 4 * j < D) ? ldg_f4(p.h_in + (size_t)v * D + d0 + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
        return r;
    }
    __device__ __forceinline__ bool row_begin(int v, bool live) const { return live && __ldcg(p.nonfinite + v) == 0; }     // flagged rows: dgn_exact_rows_kernel
    struct RowState {};
    __device__ __forceinline__ RowState row_state() const { return RowState{}; }
    __device__ __forceinline__ void row_end(int, const RowState&) const {}
    __device__ __forceinline__ void store(int v, int d0, const uint32_t (&acc)[16], const Pre& pre, RowState&) const
    {
#pragma unroll
        for (int j = 0; j < 16; j += 4)
            if (d0 + j < D)
            {
                const float4 bb = ldg_f4(b + d0 + j);
                const float4 hv = pre.h[j / 4];
                stg_f4_stream(h_out + (size_t)v * D + d0 + j,
                              make_float4(hv.x + relu_f(__uint_as_float(acc[j]) + bb.x), hv.y + relu_f(__uint_as_float(acc[j + 1]) + bb.y),
                                          hv.z + relu_f(__uint_as_float(acc[j + 2]) + bb.z), hv.w + relu_f(__uint_as_float(acc[j + 3]) + bb.w)));
            }
    }
};

}  // namespace

size_t dgn_tc_pack_bytes() { return (size_t)NCHUNK * tcg::Cfg<NPAD>::B_BLOCK; }

// W_l [100][200] ("[out][part*100 + in]") -> four [112 x 64] hi | lo chunks, k' = part*128 + in
void dgn_tc_pack_layer(const float* w, unsigned char* dst, uint16_t (*bf16_rn)(float), float (*bf16_to_float)(uint16_t))
{
    tcg::pack_weights<NPAD>(w, D, KA, NCHUNK, [](int k) { return (k % KPART) < D ? (k / KPART) * D + (k % KPART) : -1; }, dst, bf16_rn,
                            bf16_to_float);
}

// W_l for dgn_fused: k' = chunk * 64 + part * 32 + (in % 32), chunk = in / 32
void dgn_fused_pack_layer(const float* w, unsigned char* dst, uint16_t (*bf16_rn)(float), float (*bf16_to_float)(uint16_t))
{
    tcg::pack_weights<NPAD>(w, D, KA, NCHUNK, [](int k) { const int col = 32 * (k / 64) + k % 32; return col < D ? ((k % 64) / 32) * D + col : -1; }, dst,
                            bf16_rn, bf16_to_float);
}

int dgn_layer_fused_launch(DeviceBatch& b, const DgnWeights& w, int l, const float* h_in, float* h_out, int sm_count, cudaStream_t s)
{
    using C = tcf::Cfg<DgnFused>;
    FG_TRY(opt_in_smem(reinterpret_cast<const void*>(&tcf::fused_kernel<DgnFused>), C::BYTES));
    const long N = b.total_nodes;
    if (N == 0) return 0;
    FG_TRY(b.nonfinite.reserve((size_t)N + 16));
    FG_TRY(zero_bytes_launch(b.nonfinite.ptr, (size_t)N, s));
    DgnAggParams p{};
    p.h_in = h_in;
    p.in_ptr = b.in_ptr.as<int>(); p.src = b.src.as<int>(); p.eig_w = b.edge_w.as<float>(); p.out_deg = b.out_deg.as<int>();
    p.abssum = b.node_w0.as<float>(); p.wsum = b.node_w1.as<float>();
    p.apack = nullptr; p.nonfinite = b.nonfinite.as<unsigned char>();
    p.num_nodes = N;
    const float* bias = w.b.as<float>() + (size_t)l * DP;
    tcf::Args g{};
    g.wpack = w.wpack_fused.as<unsigned char>() + (size_t)l * dgn_tc_pack_bytes();
    g.num_nodes = (int)N; g.num_tiles = (int)ceil_div<long>(N, tcf::TM);
    DgnFused m{p, bias, h_out};
    tcf::fused_kernel<DgnFused><<<std::min(g.num_tiles, sm_count), tcf::NT, C::BYTES, s>>>(g, m);
    FG_CUDA(cudaGetLastError());
    {
        const int blocks = (int)std::min<long>(ceil_div<long>(N, 32 * EX_WARPS), (long)sm_count * 8);
        dgn_exact_rows_kernel<<<blocks, EX_WARPS * 32, 0, s>>>(p, w.wt.as<float>() + (size_t)l * KA * DP, bias, h_out);
        FG_CUDA(cudaGetLastError());
    }
    return 0;
}

int dgn_layer_tc_launch(DeviceBatch& b, const DgnWeights& w, int l, const float* h_in, float* h_out, int sm_count, cudaStream_t s)
{
    using C = tcg::Cfg<NPAD>;
    FG_TRY(opt_in_smem(reinterpret_cast<const void*>(&tcg::gemm_kernel<NCHUNK, NPAD, DgnEpi>), C::BYTES));
    const long N = b.total_nodes;
    const int num_tiles = (int)ceil_div<long>(N, tcg::TM);
    FG_TRY(b.apack.reserve((size_t)num_tiles * NCHUNK * tcg::A_BLOCK));
    FG_TRY(b.nonfinite.reserve((size_t)N + 16));
    FG_TRY(zero_bytes_launch(b.nonfinite.ptr, (size_t)N, s));
    DgnAggParams p{};
    p.h_in = h_in;
    p.in_ptr = b.in_ptr.as<int>(); p.src = b.src.as<int>(); p.eig_w = b.edge_w.as<float>(); p.out_deg = b.out_deg.as<int>();
    p.abssum = b.node_w0.as<float>(); p.wsum = b.node_w1.as<float>();
    p.apack = b.apack.as<unsigned char>(); p.nonfinite = b.nonfinite.as<unsigned char>();
    p.num_nodes = N;
    {
        const int blocks = (int)std::min<long>(ceil_div<long>(ceil_div<long>(N, 32) * G8, AG_WARPS), (long)sm_count * 8);
        dgn_aggregate_kernel<<<blocks, AG_WARPS * 32, 0, s>>>(p);
        FG_CUDA(cudaGetLastError());
    }
    const float* bias = w.b.as<float>() + (size_t)l * DP;
    {
        tcg::GemmArgs g{};
        g.apack = b.apack.as<unsigned char>();
        g.wpack = w.wpack_tc.as<unsigned char>() + (size_t)l * dgn_tc_pack_bytes();
        g.num_nodes = (int)N; g.num_tiles = num_tiles;
        DgnEpi epi{bias, h_in, h_out, b.nonfinite.as<unsigned char>()};
        tcg::gemm_kernel<NCHUNK, NPAD, DgnEpi><<<std::min(num_tiles, sm_count), tcg::NT, C::BYTES, s>>>(g, epi);
        FG_CUDA(cudaGetLastError());
    }
    {
        const int blocks = (int)std::min<long>(ceil_div<long>(N, 32 * EX_WARPS), (long)sm_count * 8);
        dgn_exact_rows_kernel<<<blocks, EX_WARPS * 32, 0, s>>>(p, w.wt.as<float>() + (size_t)l * KA * DP, bias, h_out);
        FG_CUDA(cudaGetLastError());
    }
    return 0;
}

}  // namespace fg

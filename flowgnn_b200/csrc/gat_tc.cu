// GAT with its dense layers on the B200 tensor cores (option "gat_tc", default; gat.cu is the FP32 version and documents the
// reference pipeline, GAT/src/GAT_compute.cc:47-108, message_passing.cc:94-157, node_embedding.cc:98-271, finalize.cc:46-112).
//
// The layer is re-cut so that ONE dense product per launch is left (fused_tc.cuh): both linear maps applied to a layer's output
// o_l -- next layer's projection W_proj_{l+1} and next layer's skip projection W_skip_{l+1} -- form one [64] x [64 x 128] GEMM:
//   gat_embed_kernel (gat.cu)      hproj_0 = x W_proj_0, skip_0 = x W_skip_0 (raw integer features), S_0, T_0
//   fused_kernel<GatFused>, l<4    gather: msg = softmax-weighted sum of hproj_l over self + in-neighbours (weights
//                                  exp(leaky(S_l[v] + T_l[u])), no max subtraction, self loop first, CSR order);
//                                  o_l = ELU(msg + skip_l) -> bf16 hi/lo A tile (K = 64: one chunk);
//                                  GEMM [W_proj_{l+1} ; W_skip_{l+1}]; epilogue: hproj_{l+1}, skip_{l+1}, and the scores
//                                  S_{l+1}, T_{l+1} = <hproj_{l+1} per head, a_src / a_tgt> accumulated over the accumulator pieces
//   gat_final_kernel               l = 4: emb = (sum_h msg + sum_h skip_4) / 4   (finalize.cc:46-112), then the shared pool + head
// exp(leaky(S_v + T_u)) is evaluated ONCE per (edge, head): lane j < 4 of the 8 lanes that share a row evaluates head j and the four
// weights travel by shuffle (the FP32 kernel evaluates all four in each of its 16 column threads).
#include "internal.cuh"
#include "layers.cuh"
#include "fused_tc.cuh"

#include <algorithm>

namespace fg {

namespace {

constexpr int HF = 64;               // heads * dims, index d * 4 + h
constexpr int NH = 4;

// Eight lanes share a row: lane j holds dims 2j and 2j + 1 (eight consecutive floats of the [dim][head] row), lanes 0..3 of the
// group also evaluate the attention weight of head j.  Two rows per lane (a warp covers 4 + 4 rows).
struct GatAttend {
    const float* hproj; const float* S; const float* T;
    const int* in_ptr; const int* src;

    struct Rows { int e0[2], end[2]; };
    __device__ __forceinline__ Rows rows_begin(const int (&v)[2], const bool (&live)[2]) const
    {
        Rows r;
#pragma unroll
        for (int q = 0; q < 2; q++)
        {
            r.e0[q] = live[q] ? __ldg(in_ptr + v[q]) : 0;
            r.end[q] = live[q] ? __ldg(in_ptr + v[q] + 1) : 0;
        }
        return r;
    }

    // msg[q][i] = (sum_u w_u hproj_u) / (sum_u w_u) for dim 2j + i (four heads) of the lane's two rows; u = v first, then the
    // in-edges in CSR order.  The eight lanes of a row execute this together (the weights are exchanged with shuffles inside the group).
    __device__ __forceinline__ void attend2(const Rows& rows, const int (&v)[2], const bool (&live)[2], int j, float4 (&msg)[2][2]) const
    {
        const int base = threadIdx.x & 24, hsel = j & 3;
        const unsigned group_mask = 0xFFu << base;
        float sv[2];
        float4 num[2][2], den[2];
        int e[2];
#pragma unroll
        for (int q = 0; q < 2; q++)
        {
            sv[q] = live[q] ? __ldg(S + (size_t)v[q] * NH + hsel) : 0.f;
            num[q][0] = num[q][1] = den[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            e[q] = rows.e0[q] - 1;                                  // position e0 - 1 stands for the self loop
        }
        // the source of the NEXT step is requested together with the current rows
        int u[2];
#pragma unroll
        for (int q = 0; q < 2; q++) u[q] = v[q];
        while (true)
        {
            const bool act0 = live[0] && e[0] < rows.end[0], act1 = live[1] && e[1] < rows.end[1];
            if (!(act0 | act1)) break;                              // the eight lanes of a row agree on this
            const bool act[2] = {act0, act1};
            float tu[2];
            float4 hu[2][2];
            int un[2];
#pragma unroll
            for (int q = 0; q < 2; q++)
            {
                un[q] = v[q];
                if (act[q])
                {
                    tu[q] = __ldg(T + (size_t)u[q] * NH + hsel);
                    hu[q][0] = ldg_f4(hproj + (size_t)u[q] * HF + 8 * j);
                    hu[q][1] = ldg_f4(hproj + (size_t)u[q] * HF + 8 * j + 4);
                    if (e[q] + 1 < rows.end[q]) un[q] = __ldg(src + e[q] + 1);
                }
            }
#pragma unroll
            for (int q = 0; q < 2; q++)
            {
                float w = 0.f;
                if (act[q])
                {
                    float sc = sv[q] + tu[q];
                    sc = (sc < 0.f) ? sc * 0.2f : sc;               // leaky relu, slope 0.2 (message_passing.cc:126-127)
                    w = __expf(sc);
                }
                // within a group every lane has the same act[q]; the other three groups of the warp may be elsewhere in their walks
                const float w0 = __shfl_sync(group_mask, w, base + 0), w1 = __shfl_sync(group_mask, w, base + 1);
                const float w2 = __shfl_sync(group_mask, w, base + 2), w3 = __shfl_sync(group_mask, w, base + 3);
                if (act[q])
                {
                    den[q].x += w0; den[q].y += w1; den[q].z += w2; den[q].w += w3;
#pragma unroll
                    for (int i = 0; i < 2; i++)
                    {
                        num[q][i].x += w0 * hu[q][i].x; num[q][i].y += w1 * hu[q][i].y; num[q][i].z += w2 * hu[q][i].z; num[q][i].w += w3 * hu[q][i].w;
                    }
                    e[q]++;
                    u[q] = un[q];
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 2; q++)
        {
            // one IEEE reciprocal per head instead of a division per element: <= 1 ulp from the quotient
            const float4 rd = make_float4(1.0f / den[q].x, 1.0f / den[q].y, 1.0f / den[q].z, 1.0f / den[q].w);
#pragma unroll
            for (int i = 0; i < 2; i++)
                msg[q][i] = live[q] ? make_float4(num[q][i].x * rd.x, num[q][i].y * rd.y, num[q][i].z * rd.z, num[q][i].w * rd.w) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
};

__device__ __forceinline__ float elu_f(float o) { return (o <= 0.f) ? __expf(o) - 1.0f : o; }

struct GatFused {
    static constexpr int NCHUNK = 1, NPAD = 128;
    static constexpr unsigned ksteps(int) { return 0xFu; }
    GatAttend at;
    const float* skip;                                                    // skip_l [N][64]
    float* hproj_out; float* skip_out; float* S_out; float* T_out;        // layer l + 1
    const float* a_src; const float* a_tgt;                               // [64] of layer l + 1, index d * 4 + h

    static constexpr int LPR = 8;                                          // lanes per row (fused_tc.cuh): 8 K slots per lane, two rows per lane
    using Rows = GatAttend::Rows;
    __device__ __forceinline__ Rows rows_begin(const int (&v)[2], const bool (&live)[2]) const { return at.rows_begin(v, live); }
    __device__ __forceinline__ bool gather2(const Rows& rows, const int (&v)[2], const bool (&live)[2], int, int j, float4 (&x)[2][2]) const
    {
        float4 msg[2][2];
        at.attend2(rows, v, live, j, msg);
#pragma unroll
        for (int q = 0; q < 2; q++)
#pragma unroll
            for (int i = 0; i < 2; i++)
            {
                x[q][i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (!live[q]) continue;
                const float4 sk = ldg_f4(skip + (size_t)v[q] * HF + 8 * j + 4 * i);
                x[q][i] = make_float4(elu_f(msg[q][i].x + sk.x), elu_f(msg[q][i].y + sk.y), elu_f(msg[q][i].z + sk.z), elu_f(msg[q][i].w + sk.w));      // node_embedding.cc:176-195
            }
        return true;
    }
    __device__ __forceinline__ void prefetch_tile(int v0, int rows) const
    {
        tcf::prefetch_l2(at.hproj + (size_t)v0 * HF, rows * HF * 4);
        tcf::prefetch_l2(skip + (size_t)v0 * HF, rows * HF * 4);
        tcf::prefetch_l2(at.S + (size_t)v0 * NH, rows * NH * 4);
        tcf::prefetch_l2(at.T + (size_t)v0 * NH, rows * NH * 4);
        tcf::prefetch_l2(at.in_ptr + v0, rows * 4 + 4);
        const int e0 = __ldg(at.in_ptr + v0), e1 = __ldg(at.in_ptr + v0 + rows);
        tcf::prefetch_l2(at.src + e0, (e1 - e0) * 4);
    }
    struct Pre {};
    __device__ __forceinline__ Pre preload(int, bool, int) const { return Pre{}; }
    __device__ __forceinline__ bool row_begin(int, bool live) const { return live; }
    struct RowState { float s[4], t[4]; };
    __device__ __forceinline__ RowState row_state() const { return RowState{{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}; }
    // accumulator columns 0..63 = hproj_{l+1} (index d * 4 + h), 64..127 = skip_{l+1}
    __device__ __forceinline__ void store(int v, int d0, const uint32_t (&acc)[16], const Pre&, RowState& st) const
    {
        float* dst = (d0 < HF) ? hproj_out + (size_t)v * HF + d0 : skip_out + (size_t)v * HF + (d0 - HF);
#pragma unroll
        for (int k = 0; k < 16; k += 4)
            stg_f4_stream(dst + k, make_float4(__uint_as_float(acc[k]), __uint_as_float(acc[k + 1]), __uint_as_float(acc[k + 2]), __uint_as_float(acc[k + 3])));
        if (d0 < HF)
        {
            // S[v][h] = sum_d hproj'[v][d][h] a_src[d][h], d ascending (node_embedding.cc:236-262)
#pragma unroll
            for (int k = 0; k < 16; k++)
            {
                const float r = __uint_as_float(acc[k]);
                st.s[k & 3] = fmaf(r, __ldg(a_src + d0 + k), st.s[k & 3]);
                st.t[k & 3] = fmaf(r, __ldg(a_tgt + d0 + k), st.t[k & 3]);
            }
        }
    }
    __device__ __forceinline__ void row_end(int v, const RowState& st) const
    {
        *reinterpret_cast<float4*>(S_out + (size_t)v * NH) = make_float4(st.s[0], st.s[1], st.s[2], st.s[3]);
        *reinterpret_cast<float4*>(T_out + (size_t)v * NH) = make_float4(st.t[0], st.t[1], st.t[2], st.t[3]);
    }
};

// last layer: emb[v][d] = (sum_h msg[d][h] + sum_h skip_4[d][h]) / 4 (finalize.cc:46-112); eight lanes per row, eight rows per warp
__global__ void __launch_bounds__(256) gat_final_kernel(GatAttend at, const float* __restrict__ skip, float* __restrict__ emb, int num_nodes)
{
    const int lane = threadIdx.x & 31, sub = lane >> 3, j = lane & 7;
    const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5, nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long v0 = 8 * warp; v0 < num_nodes; v0 += 8 * nwarps)
    {
        const int v[2] = {(int)v0 + sub, (int)v0 + 4 + sub};
        const bool live[2] = {v[0] < num_nodes, v[1] < num_nodes};
        const GatAttend::Rows rows = at.rows_begin(v, live);
        float4 msg[2][2];
        at.attend2(rows, v, live, j, msg);
#pragma unroll
        for (int q = 0; q < 2; q++)
            if (live[q])
            {
                float of[2];
#pragma unroll
                for (int i = 0; i < 2; i++)
                {
                    const float4 sk = ldg_f4(skip + (size_t)v[q] * HF + 8 * j + 4 * i);
                    float o = 0.f;
                    o += msg[q][i].x; o += msg[q][i].y; o += msg[q][i].z; o += msg[q][i].w;
                    o += sk.x; o += sk.y; o += sk.z; o += sk.w;
                    of[i] = o / 4.0f;
                }
                *reinterpret_cast<float2*>(emb + (size_t)v[q] * 16 + 2 * j) = make_float2(of[0], of[1]);
            }
    }
}

}  // namespace

size_t gat_tc_pack_bytes() { return (size_t)tcg::Cfg<128>::B_BLOCK; }

// [W_proj ; W_skip] of one layer, given k-major ([k = di*4+hi][n = do*4+ho]) as api.cu holds them -> one [128 x 64] hi | lo block
void gat_tc_pack_layer(const float* projt, const float* skipt, unsigned char* dst, uint16_t (*bf16_rn)(float), float (*bf16_to_float)(uint16_t))
{
    float w[128 * 64];
    for (int n = 0; n < 64; n++)
        for (int k = 0; k < 64; k++) { w[n * 64 + k] = projt[k * 64 + n]; w[(64 + n) * 64 + k] = skipt[k * 64 + n]; }
    tcg::pack_weights<128>(w, 128, 64, 1, [](int k) { return k; }, dst, bf16_rn, bf16_to_float);
}

// layer l = 0..3 as one fused launch (reads hproj_l, skip_l, S_l, T_l; writes those of layer l + 1)
int gat_layer_tc_launch(const DeviceBatch& b, const GatWeights& w, int l, const float* hproj, const float* skip, const float* S, const float* T,
                        float* hproj_out, float* skip_out, float* S_out, float* T_out, int sm_count, cudaStream_t s)
{
    using C = tcf::Cfg<GatFused>;
    FG_TRY(opt_in_smem(reinterpret_cast<const void*>(&tcf::fused_kernel<GatFused>), C::BYTES));
    const long N = b.total_nodes;
    if (N == 0) return 0;
    tcf::Args g{};
    g.wpack = w.wpack_tc.as<unsigned char>() + (size_t)(l + 1) * gat_tc_pack_bytes();
    g.num_nodes = (int)N; g.num_tiles = (int)ceil_div<long>(N, tcf::TM);
    GatFused m{GatAttend{hproj, S, T, b.in_ptr.as<int>(), b.src.as<int>()}, skip, hproj_out, skip_out, S_out, T_out,
               w.a_src.as<float>() + (size_t)(l + 1) * HF, w.a_tgt.as<float>() + (size_t)(l + 1) * HF};
    tcf::fused_kernel<GatFused><<<std::min(g.num_tiles, sm_count), tcf::NT, C::BYTES, s>>>(g, m);
    FG_CUDA(cudaGetLastError());
    return 0;
}

int gat_final_launch(const DeviceBatch& b, const float* hproj, const float* skip, const float* S, const float* T, float* emb, int sm_count, cudaStream_t s)
{
    const long N = b.total_nodes;
    if (N == 0) return 0;
    gat_final_kernel<<<(int)std::min<long>(ceil_div<long>(N, 64), (long)sm_count * 8), 256, 0, s>>>(
        GatAttend{hproj, S, T, b.in_ptr.as<int>(), b.src.as<int>()}, skip, emb, (int)N);
    FG_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace fg

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
// exp(leaky(S_v + T_u)) is evaluated ONCE per (edge, head): lane j of the 4 lanes that share a row evaluates head j and the four
// weights travel by shuffle (the FP32 kernel evaluates all four in each of its 16 column threads).
#include "internal.cuh"
#include "layers.cuh"
#include "fused_tc.cuh"

#include <algorithm>

namespace fg {

namespace {

constexpr int HF = 64;               // heads * dims, index d * 4 + h
constexpr int NH = 4;

// Four lanes share a row: lane j holds dims j, 4 + j, 8 + j, 12 + j (its i-th load is the 16 bytes at 64 i + 16 j of the [dim][head]
// row, so the four lanes of a row read 64 contiguous bytes per instruction: full sectors) and evaluates the attention weight of head j.  One row per lane group: a warp walks its eight rows together.
struct GatAttend {
    const float* hproj; const float* S; const float* T;
    const int* in_ptr; const int* src;

    struct Rows { int e0, end; };
    __device__ __forceinline__ Rows rows_begin(int v, bool live) const
    {
        Rows r;
        r.e0 = live ? __ldg(in_ptr + v) : 0;
        r.end = live ? __ldg(in_ptr + v + 1) : 0;
        return r;
    }

    // msg[i] = (sum_u w_u hproj_u) / (sum_u w_u) for dim 4i + j (four heads); u = v first, then the in-edges in CSR order.  The four
    // lanes of a row execute this together (the weights are exchanged with shuffles inside the group).
    __device__ __forceinline__ void attend1(const Rows& rows, int v, bool live, int j, float4 (&msg)[4]) const
    {
        const int base = threadIdx.x & 28;
        const unsigned group_mask = 0xFu << base;
        const float sv = live ? __ldg(S + (size_t)v * NH + j) : 0.f;
        float4 num[4], den = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 4; i++) num[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        int e = rows.e0 - 1;                                        // position e0 - 1 stands for the self loop
        int u = v;
        while (live && e < rows.end)                                // the four lanes of a row agree on this
        {
            // the source of the NEXT step is requested together with the current row
            const float tu = __ldg(T + (size_t)u * NH + j);
            float4 hu[4];
#pragma unroll
            for (int i = 0; i < 4; i++) hu[i] = ldg_f4(hproj + (size_t)u * HF + 16 * i + 4 * j);
            int un = v;
            if (e + 1 < rows.end) un = __ldg(src + e + 1);
            float sc = sv + tu;
            sc = (sc < 0.f) ? sc * 0.2f : sc;                       // leaky relu, slope 0.2 (message_passing.cc:126-127)
            const float w = __expf(sc);
            const float w0 = __shfl_sync(group_mask, w, base + 0), w1 = __shfl_sync(group_mask, w, base + 1);
            const float w2 = __shfl_sync(group_mask, w, base + 2), w3 = __shfl_sync(group_mask, w, base + 3);
            den.x += w0; den.y += w1; den.z += w2; den.w += w3;
#pragma unroll
            for (int i = 0; i < 4; i++)
            {
                num[i].x += w0 * hu[i].x; num[i].y += w1 * hu[i].y; num[i].z += w2 * hu[i].z; num[i].w += w3 * hu[i].w;
            }
            e++;
            u = un;
        }
        // one IEEE reciprocal per head instead of a division per element: <= 1 ulp from the quotient
        const float4 rd = make_float4(1.0f / den.x, 1.0f / den.y, 1.0f / den.z, 1.0f / den.w);
#pragma unroll
        for (int i = 0; i < 4; i++)
            msg[i] = live ? make_float4(num[i].x * rd.x, num[i].y * rd.y, num[i].z * rd.z, num[i].w * rd.w) : make_float4(0.f, 0.f, 0.f, 0.f);
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

    static constexpr int LPR = 4;                                          // lanes per row (fused_tc.cuh): 16 K slots per lane, one row per lane group
    using Rows = GatAttend::Rows;
    __device__ __forceinline__ Rows rows_begin(int v, bool live) const { return at.rows_begin(v, live); }
    static __device__ __forceinline__ int kslot(int j, int i) { return 16 * i + 4 * j; }      // piece i of lane j: dim 4i + j, heads 0..3
    __device__ __forceinline__ bool gather1(const Rows& rows, int v, bool live, int, int j, float4 (&x)[4]) const
    {
        float4 msg[4];
        at.attend1(rows, v, live, j, msg);
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!live) continue;
            const float4 sk = ldg_f4(skip + (size_t)v * HF + 16 * i + 4 * j);
            x[i] = make_float4(elu_f(msg[i].x + sk.x), elu_f(msg[i].y + sk.y), elu_f(msg[i].z + sk.z), elu_f(msg[i].w + sk.w));      // node_embedding.cc:176-195
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

// last layer: emb[v][d] = (sum_h msg[d][h] + sum_h skip_4[d][h]) / 4 (finalize.cc:46-112); four lanes per row, eight rows per warp
__global__ void __launch_bounds__(256) gat_final_kernel(GatAttend at, const float* __restrict__ skip, float* __restrict__ emb, int num_nodes)
{
    const int lane = threadIdx.x & 31, sub = lane >> 2, j = lane & 3;
    const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5, nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long v0 = 8 * warp; v0 < num_nodes; v0 += 8 * nwarps)
    {
        const int v = (int)v0 + sub;
        const bool live = v < num_nodes;
        const GatAttend::Rows rows = at.rows_begin(v, live);
        float4 msg[4];
        at.attend1(rows, v, live, j, msg);
        if (live)
        {
            float of[4];
#pragma unroll
            for (int i = 0; i < 4; i++)
            {
                const float4 sk = ldg_f4(skip + (size_t)v * HF + 16 * i + 4 * j);
                float o = 0.f;
                o += msg[i].x; o += msg[i].y; o += msg[i].z; o += msg[i].w;
                o += sk.x; o += sk.y; o += sk.z; o += sk.w;
                of[i] = o / 4.0f;
            }
            #pragma unroll
            for (int i = 0; i < 4; i++) emb[(size_t)v * 16 + 4 * i + j] = of[i];
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

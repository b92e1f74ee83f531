// DGN forward in the reference's OWN arithmetic: ap_fixed<16,3> (DGN/src/dcl.h:54-55), option "fixed_point".
//
// 16-bit values with 13 fraction bits (range [-4, 4)); see gin_fixed.cu for the rules (exact wide intermediates, floor on
// assignment, wrap to 16 bits, integer division toward zero).  Statements that differ from GIN's:
//   message[v][1] += h_u * eig_w            one floored product per edge                     (message_passing.cc:150)
//   a1 = message_1 / degree                 raw sdiv out-degree                              (node_embedding.cc:145)
//   a2 = |FM_TYPE((message_2 - eigw_sum * h) / eig_abssum)|                                   (node_embedding.cc:146)
//        the numerator is exact with 26 fraction bits; ap_fixed_base::operator/ shifts it left by the divisor's 13 and
//        divides the raw integers toward zero (a 64-bit division), the cast floors to 13 bits and wraps, hls::abs wraps
//   addend = a1 * w0 + a2 * w1              ONE floor for the sum of the two products        (node_embedding.cc:152-155)
// A node with out-degree 0 divides by zero in the reference (undefined; the fp32 flavour gives NaN): the checker's
// emulation defines that quotient as 0 and so does this kernel.
// Checked bit for bit against the reference's unmodified sources compiled over an ap_fixed emulation (tests/test_fixed_point.py).
#include "internal.cuh"
#include "fixed.cuh"

#include <algorithm>

namespace fg {

namespace {

constexpr int D = 100;
constexpr int F = 13;
constexpr int TM = 64;               // nodes per tile
constexpr int NT = 416;              // 400 compute threads: 25 output groups x 16 node groups of 4
constexpr int W_OFF = 0;                                // int32 [100 k][2][100]  w << 3
constexpr int A1_OFF = W_OFF + 4 * D * 2 * D;           // int32 [TM][100]  a1 << 16
constexpr int A2_OFF = A1_OFF + 4 * TM * D;             // int32 [TM][100]  a2 << 16
constexpr int SMEM_BYTES = A2_OFF + 4 * TM * D;         // 131,200

// ---- load_graph, eigenvector part (DGN/src/load_inputs.cc:104-111): eig_w = phi_u - phi_v per in-edge of v, and its sums ----
__device__ __forceinline__ int to_fixed13(float x)
{
    const float scaled = floorf(x * 8192.0f);                    // exact scaling, AP_TRN
    if (!(fabsf(scaled) < 9.0e18f)) return 0;                    // inf / NaN: undefined in the reference's host cast
    return wrap16((int)(long long)scaled);                       // AP_WRAP
}

__global__ void dgn_fixed_eig_kernel(const float* __restrict__ eig, const int* __restrict__ in_ptr, const int* __restrict__ src, int16_t* __restrict__ ew,
                                     short2* __restrict__ sums, long num_nodes)
{
    for (long v = blockIdx.x * (long)blockDim.x + threadIdx.x; v < num_nodes; v += (long)gridDim.x * blockDim.x)
    {
        const int pv = to_fixed13(__ldg(eig + 4 * v + 1));
        int sa = 0, sw = 0;
        const int e1 = __ldg(in_ptr + v + 1);
        for (int e = __ldg(in_ptr + v); e < e1; e++)
        {
            const int d = wrap16(to_fixed13(__ldg(eig + 4 * (long)__ldg(src + e) + 1)) - pv);
            ew[e] = (int16_t)d;
            sa += abs16(d);
            sw += d;
        }
        sums[v] = make_short2((short)sa, (short)sw);             // eig_abssums, eigw_sums
    }
}

struct DgnFixedParams {
    const int16_t* h_in; int16_t* h_out;
    const int* in_ptr; const int* src; const int16_t* ew; const short2* sums; const int* out_deg;
    const int* w; const int* b;      // [100][2][100] raw << 3, [100] raw
    int num_nodes; int num_tiles;
};

__global__ void __launch_bounds__(NT, 1) dgn_fixed_layer_kernel(DgnFixedParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    int* ws = reinterpret_cast<int*>(smem + W_OFF);
    int* a1s = reinterpret_cast<int*>(smem + A1_OFF);
    int* a2s = reinterpret_cast<int*>(smem + A2_OFF);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    for (int i = tid; i < D * 2 * D / 4; i += NT) reinterpret_cast<int4*>(ws)[i] = __ldg(reinterpret_cast<const int4*>(p.w) + i);
    __syncthreads();

    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x)
    {
        const int n0 = tile * TM;
        const int rows = min(TM, p.num_nodes - n0);

        // ---- message passing + the two activations: a warp per node, lanes 0..24 hold four columns each ----
        for (int r = wid; r < TM; r += NT / 32)
        {
            int a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0};
            if (r < rows && lane < D / 4)
            {
                const int v = n0 + r;
                int m0[4] = {0, 0, 0, 0}, m1[4] = {0, 0, 0, 0};
                const int e1 = __ldg(p.in_ptr + v + 1);
                for (int e = __ldg(p.in_ptr + v); e < e1; e++)
                {
                    const int u = __ldg(p.src + e);
                    const int w = (int)__ldg(p.ew + e) << (16 - F);
                    const short4 hu = __ldg(reinterpret_cast<const short4*>(p.h_in + (size_t)u * D) + lane);
                    const int h[4] = {hu.x, hu.y, hu.z, hu.w};
#pragma unroll
                    for (int j = 0; j < 4; j++) { m0[j] += h[j]; m1[j] = mad_hi(h[j] << 16, w, m1[j]); }
                }
                const short4 hv4 = __ldg(reinterpret_cast<const short4*>(p.h_in + (size_t)v * D) + lane);
                const int hv[4] = {hv4.x, hv4.y, hv4.z, hv4.w};
                const short2 sm = __ldg(p.sums + v);
                const long long abssum = sm.x == 0 ? 1 : sm.x;                   // ap_fixed_epsilon<WT_TYPE>() (node_embedding.cc:126-129)
                const int od = __ldg(p.out_deg + v);
#pragma unroll
                for (int j = 0; j < 4; j++)
                {
                    a1[j] = od ? wrap16(m0[j]) / od : 0;
                    const long long num = ((long long)wrap16(m1[j]) << F) - (long long)sm.y * hv[j];      // 26 fraction bits, exact
                    const long long quo = (num * (1ll << F)) / abssum;                                    // toward zero, 26 fraction bits
                    a2[j] = abs16((int)(quo >> F));                                                       // floor to 13 bits, wrap, |.|
                }
            }
            if (lane < D / 4)
            {
                *reinterpret_cast<int4*>(a1s + r * D + 4 * lane) = make_int4(a1[0] << 16, a1[1] << 16, a1[2] << 16, a1[3] << 16);
                *reinterpret_cast<int4*>(a2s + r * D + 4 * lane) = make_int4(a2[0] << 16, a2[1] << 16, a2[2] << 16, a2[3] << 16);
            }
        }
        __syncthreads();

        // ---- acc = b + sum_k floor((a1_k w0_k + a2_k w1_k) / 2^13); h' = h + relu(acc) ----
        if (tid < 400)
        {
            const int og = tid % 25, ng = tid / 25;
            int acc[4][4];
            const int4 b = __ldg(reinterpret_cast<const int4*>(p.b) + og);
#pragma unroll
            for (int i = 0; i < 4; i++) { acc[i][0] = b.x; acc[i][1] = b.y; acc[i][2] = b.z; acc[i][3] = b.w; }
#pragma unroll 1
            for (int k = 0; k < D; k += 4)
            {
                int4 x1[4], x2[4];
#pragma unroll
                for (int i = 0; i < 4; i++)
                {
                    x1[i] = *reinterpret_cast<const int4*>(a1s + (4 * ng + i) * D + k);
                    x2[i] = *reinterpret_cast<const int4*>(a2s + (4 * ng + i) * D + k);
                }
#pragma unroll
                for (int kk = 0; kk < 4; kk++)
                {
                    const int4 w0 = *reinterpret_cast<const int4*>(ws + (k + kk) * 2 * D + 4 * og);
                    const int4 w1 = *reinterpret_cast<const int4*>(ws + (k + kk) * 2 * D + D + 4 * og);
#pragma unroll
                    for (int i = 0; i < 4; i++)
                    {
                        const int u1 = kk == 0 ? x1[i].x : kk == 1 ? x1[i].y : kk == 2 ? x1[i].z : x1[i].w;
                        const int u2 = kk == 0 ? x2[i].x : kk == 1 ? x2[i].y : kk == 2 ? x2[i].z : x2[i].w;
                        acc[i][0] += (int)(mad_wide(u2, w1.x, mul_wide(u1, w0.x)) >> 32);
                        acc[i][1] += (int)(mad_wide(u2, w1.y, mul_wide(u1, w0.y)) >> 32);
                        acc[i][2] += (int)(mad_wide(u2, w1.z, mul_wide(u1, w0.z)) >> 32);
                        acc[i][3] += (int)(mad_wide(u2, w1.w, mul_wide(u1, w0.w)) >> 32);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; i++)
            {
                const int r = 4 * ng + i;
                if (r < rows)
                {
                    const short4 hv = __ldg(reinterpret_cast<const short4*>(p.h_in + (size_t)(n0 + r) * D) + og);
                    reinterpret_cast<short4*>(p.h_out + (size_t)(n0 + r) * D)[og] =
                        make_short4((short)(hv.x + relu16(acc[i][0])), (short)(hv.y + relu16(acc[i][1])), (short)(hv.z + relu16(acc[i][2])),
                                    (short)(hv.w + relu16(acc[i][3])));
                }
            }
        }
        __syncthreads();
    }
}

// finalize (DGN/src/finalize.cc:14-105): mean pool toward zero, then Linear 100 -> 50 (relu), 50 -> 25 (relu), 25 -> 1, every product floored
constexpr int HEAD_WARPS = 8;
__global__ void __launch_bounds__(HEAD_WARPS * 32) dgn_fixed_pool_head_kernel(const int16_t* __restrict__ h, const int* __restrict__ node_off, const int* __restrict__ nn,
                                                                            const int* __restrict__ w0, const int* __restrict__ b0, const int* __restrict__ w1,
                                                                            const int* __restrict__ b1, const int* __restrict__ w2, const int* __restrict__ b2,
                                                                            float* __restrict__ out, int num_graphs)
{
    __shared__ int s_w0[100 * 50], s_w1[50 * 25], s_w2[25], s_b0[50], s_b1[25];
    __shared__ __align__(16) int s_x[HEAD_WARPS][100];
    __shared__ int s_y[HEAD_WARPS][50];
    for (int i = threadIdx.x; i < 5000; i += blockDim.x) s_w0[i] = __ldg(w0 + i);
    for (int i = threadIdx.x; i < 1250; i += blockDim.x) s_w1[i] = __ldg(w1 + i);
    if (threadIdx.x < 25) { s_w2[threadIdx.x] = __ldg(w2 + threadIdx.x); s_b1[threadIdx.x] = __ldg(b1 + threadIdx.x); }
    if (threadIdx.x < 50) s_b0[threadIdx.x] = __ldg(b0 + threadIdx.x);
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int g = blockIdx.x * HEAD_WARPS + wid; g < num_graphs; g += gridDim.x * HEAD_WARPS)
    {
        const int n = __ldg(nn + g);
        const size_t base = (size_t)__ldg(node_off + g);
        if (lane < D / 4)
        {
            int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            for (int r = 0; r < n; r++)
            {
                const short4 x = __ldg(reinterpret_cast<const short4*>(h + (base + r) * D) + lane);
                s0 += x.x; s1 += x.y; s2 += x.z; s3 += x.w;
            }
            *reinterpret_cast<int4*>(&s_x[wid][4 * lane]) =
                make_int4((n ? wrap16(s0) / n : 0) << 16, (n ? wrap16(s1) / n : 0) << 16, (n ? wrap16(s2) / n : 0) << 16, (n ? wrap16(s3) / n : 0) << 16);
        }
        __syncwarp();
        for (int o = lane; o < 50; o += 32)
        {
            int acc = s_b0[o];
            for (int k = 0; k < 100; k++) acc = mad_hi(s_x[wid][k], s_w0[k * 50 + o], acc);
            s_y[wid][o] = relu16(acc) << 16;
        }
        __syncwarp();
        int z = 0;
        if (lane < 25)
        {
            int acc = s_b1[lane];
            for (int k = 0; k < 50; k++) acc = mad_hi(s_y[wid][k], s_w1[k * 25 + lane], acc);
            z = mad_hi(relu16(acc) << 16, s_w2[lane], 0);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
        if (lane == 0) out[g] = (float)wrap16(z + __ldg(b2)) * (1.0f / 8192.0f);
        __syncwarp();
    }
}

}  // namespace

__global__ void fixed_embed_kernel(const int* __restrict__ feat, const int16_t* __restrict__ table, FixedEmbedOffsets off, int16_t* __restrict__ h, long num_nodes)
{
    const long total = num_nodes * (D / 2);
    for (long item = blockIdx.x * (long)blockDim.x + threadIdx.x; item < total; item += (long)gridDim.x * blockDim.x)
    {
        const long v = item / (D / 2);
        const int q = (int)(item - v * (D / 2));
        int s0 = 0, s1 = 0;
#pragma unroll
        for (int f = 0; f < ND_FEATURE; f++)
        {
            const int row = off.off[f] + __ldg(feat + v * ND_FEATURE + f);
            const short2 t = __ldg(reinterpret_cast<const short2*>(table + (size_t)row * D) + q);
            s0 += t.x; s1 += t.y;
        }
        reinterpret_cast<short2*>(h + v * D)[q] = make_short2((short)s0, (short)s1);
    }
}

int fixed_embed_launch(const int* feat, const int16_t* table, const FixedEmbedOffsets& off, int16_t* h, long num_nodes, int sm_count, cudaStream_t s)
{
    if (num_nodes <= 0) return 0;
    const int blocks = (int)std::min<long>(ceil_div<long>(num_nodes * (D / 2), 256), (long)sm_count * 16);
    fixed_embed_kernel<<<blocks, 256, 0, s>>>(feat, table, off, h, num_nodes);
    FG_CUDA(cudaGetLastError());
    return 0;
}

int dgn_fixed_forward(DeviceBatch& b, const DgnWeights& w, int sm_count, cudaStream_t s, int* launches)
{
    const long N = b.total_nodes;
    const int G = b.num_graphs;
    if (G == 0) return 0;
    FG_TRY(opt_in_smem(reinterpret_cast<const void*>(&dgn_fixed_layer_kernel), SMEM_BYTES));
    FG_TRY(b.act[0].reserve(sizeof(int16_t) * (size_t)std::max<long>(N, 1) * D));
    FG_TRY(b.act[1].reserve(sizeof(int16_t) * (size_t)std::max<long>(N, 1) * D));
    FG_TRY(b.edge_w.reserve(sizeof(int16_t) * (size_t)(b.total_edges + 1)));
    FG_TRY(b.node_w0.reserve(sizeof(short2) * (size_t)(N + 1)));
    int16_t* cur = b.act[0].as<int16_t>();
    int16_t* nxt = b.act[1].as<int16_t>();
    if (N > 0)
    {
        FixedEmbedOffsets off;
        for (int f = 0; f < ND_FEATURE; f++) off.off[f] = 119 * f;               // nine [119][100] tables (DGN/src/load_inputs.cc:133-137)
        FG_TRY(fixed_embed_launch(b.node_feature.as<int>(), w.fx_emb.as<int16_t>(), off, cur, N, sm_count, s));
        dgn_fixed_eig_kernel<<<(int)std::min<long>(ceil_div<long>(N, 256), (long)sm_count * 8), 256, 0, s>>>(
            b.node_eigen.as<float>(), b.in_ptr.as<int>(), b.src.as<int>(), b.edge_w.as<int16_t>(), b.node_w0.as<short2>(), N);
        FG_CUDA(cudaGetLastError());
        (*launches) += 2;
        const int tiles = (int)ceil_div<long>(N, TM);
        for (int l = 0; l < 4; l++)
        {
            DgnFixedParams p;
            p.h_in = cur; p.h_out = nxt;
            p.in_ptr = b.in_ptr.as<int>(); p.src = b.src.as<int>(); p.ew = b.edge_w.as<int16_t>(); p.sums = b.node_w0.as<short2>();
            p.out_deg = b.out_deg.as<int>();
            p.w = w.fx_w.as<int>() + (size_t)l * D * 2 * D; p.b = w.fx_b.as<int>() + l * D;
            p.num_nodes = (int)N; p.num_tiles = tiles;
            dgn_fixed_layer_kernel<<<std::min(tiles, sm_count), NT, SMEM_BYTES, s>>>(p);
            FG_CUDA(cudaGetLastError());
            (*launches)++;
            std::swap(cur, nxt);
        }
    }
    dgn_fixed_pool_head_kernel<<<std::min(ceil_div(G, HEAD_WARPS), sm_count * 4), HEAD_WARPS * 32, 0, s>>>(
        cur, b.node_off.as<int>(), b.nums_of_nodes.as<int>(), w.fx_m0w.as<int>(), w.fx_m0b.as<int>(), w.fx_m1w.as<int>(), w.fx_m1b.as<int>(),
        w.fx_m2w.as<int>(), w.fx_m2b.as<int>(), b.out.as<float>(), G);
    FG_CUDA(cudaGetLastError());
    (*launches)++;
    return 0;
}

}  // namespace fg

// Kernels shared by all models: global mean pool + prediction head ("finalize").
#include "internal.cuh"
#include "layers.cuh"

namespace fg {

namespace {

constexpr int HEAD_WARPS = 8;
constexpr int HEAD_MAXDIM = 128;

// One warp per graph.  Reference: global_mean_pooling (GIN/src/finalize.cc:36-115, same in GAT/PNA/DGN)
// then `linear` / `linear_output_stationary` / `linear_input_stationary` (*/src/linear.cc:11-149):
// bias first, then products in dim_in order; relu between layers, none on the last.
__global__ void __launch_bounds__(HEAD_WARPS * 32) pool_head_kernel(HeadParams p)
{
    __shared__ float s_buf[HEAD_WARPS][2][HEAD_MAXDIM];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int g = blockIdx.x * HEAD_WARPS + wid;
    if (g >= p.num_graphs) return;
    const int n = p.nn[g];
    const size_t base = (size_t)p.node_off[g];
    const int q4 = p.dim / 4;

    float* cur = s_buf[wid][0];
    float* nxt = s_buf[wid][1];
    if (lane < q4)
    {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < n; r++)
        {
            const float4 x = ldg_f4(p.x + (base + r) * p.dim + 4 * lane);
            s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w;
        }
        const float fn = (float)n;
        st_f4(cur + 4 * lane, make_float4(s.x / fn, s.y / fn, s.z / fn, s.w / fn));
    }
    __syncwarp();
    for (int l = 0; l < p.num_layers; l++)
    {
        const int din = p.dims[l], dout = p.dims[l + 1];
        const bool last = (l == p.num_layers - 1);
        for (int o = 0; o < dout; o++)
        {
            float part = 0.f;
            for (int i = lane; i < din; i += 32) part = fmaf(cur[i], __ldg(p.w[l] + (size_t)o * din + i), part);
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
            float r = part + __ldg(p.b[l] + o);
            if (!last) r = relu_f(r);
            if (lane == 0)
            {
                if (last) p.out[g] = r;        // NUM_TASK == 1
                else nxt[o] = r;
            }
        }
        __syncwarp();
        float* t = cur; cur = nxt; nxt = t;
    }
}

}  // namespace

int launch_pool_head(const HeadParams& p, cudaStream_t stream)
{
    if (p.num_graphs <= 0) return 0;
    if (p.dim > HEAD_MAXDIM || p.dim % 4 != 0) { set_last_error("pool_head: unsupported dim"); return FG_ERR_INVALID; }
    pool_head_kernel<<<ceil_div(p.num_graphs, HEAD_WARPS), HEAD_WARPS * 32, 0, stream>>>(p);
    FG_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace fg

// Kernels shared by all models: global mean pool + prediction head ("finalize").
#include <algorithm>

#include "internal.cuh"
#include "layers.cuh"

namespace fg {

namespace {

constexpr int HEAD_WARPS = 8;
constexpr int HEAD_MAXDIM = 128;

// One warp per graph, persistent blocks.  Reference: global_mean_pooling (GIN/src/finalize.cc:36-115, same in GAT/PNA/DGN)
// then `linear` / `linear_output_stationary` / `linear_input_stationary` (*/src/linear.cc:11-149): bias first, then the
// products in dim_in order; relu between layers, none on the last.
// The head weights (<= 6,275 floats) are staged once per block in shared memory, transposed to [in][out]: lane = output,
// every output is ONE sequential chain  r = b;  r += x_i * w_oi  (i ascending) -- the reference's own summation order, with
// separate multiply and add like its fp32 build (no FMA contraction) -- and there is no shuffle reduction per output (the
// first version spent 61 five-step shuffle trees per graph: 2.5 ms for the 437,929 graphs of the molpcba workload).
constexpr int HEAD_MAXW = 8192;
__global__ void __launch_bounds__(HEAD_WARPS * 32) pool_head_kernel(HeadParams p)
{
    __shared__ float s_w[HEAD_MAXW];
    __shared__ float s_b[3][HEAD_MAXDIM];
    __shared__ float s_buf[HEAD_WARPS][2][HEAD_MAXDIM];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int woff[4] = {0, 0, 0, 0};
    for (int l = 0; l < p.num_layers; l++)
    {
        const int din = p.dims[l], dout = p.dims[l + 1];
        woff[l + 1] = woff[l] + din * dout;
        for (int idx = threadIdx.x; idx < din * dout; idx += blockDim.x)
        {
            const int o = idx / din, i = idx - o * din;
            s_w[woff[l] + i * dout + o] = __ldg(p.w[l] + idx);
        }
        for (int o = threadIdx.x; o < dout; o += blockDim.x) s_b[l][o] = __ldg(p.b[l] + o);
    }
    __syncthreads();
    const int q4 = p.dim / 4;
    for (int g = blockIdx.x * HEAD_WARPS + wid; g < p.num_graphs; g += gridDim.x * HEAD_WARPS)
    {
        const int n = p.nn[g];
        const size_t base = (size_t)p.node_off[g];
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
            const float* wt = s_w + woff[l];
            for (int o = lane; o < dout; o += 32)
            {
                float r = s_b[l][o];
                for (int i = 0; i < din; i++) r = __fadd_rn(r, __fmul_rn(cur[i], wt[i * dout + o]));
                if (!last) r = relu_f(r);
                if (last) p.out[g] = r;        // NUM_TASK == 1
                else nxt[o] = r;
            }
            __syncwarp();
            float* t = cur; cur = nxt; nxt = t;
        }
        __syncwarp();
    }
}

}  // namespace

namespace {
__global__ void fill_kernel(float* __restrict__ out, float value, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = value;
}
}  // namespace

int fill_outputs(float* out, float value, int n, cudaStream_t stream)
{
    if (n <= 0) return 0;
    fill_kernel<<<std::min(ceil_div(n, 256), 1024), 256, 0, stream>>>(out, value, n);
    FG_CUDA(cudaGetLastError());
    return 0;
}

int launch_pool_head(const HeadParams& p, cudaStream_t stream)
{
    if (p.num_graphs <= 0) return 0;
    if (p.dim > HEAD_MAXDIM || p.dim % 4 != 0) { set_last_error("pool_head: unsupported dim"); return FG_ERR_INVALID; }
    int wtotal = 0;
    for (int l = 0; l < p.num_layers; l++) wtotal += p.dims[l] * p.dims[l + 1];
    if (wtotal > HEAD_MAXW) { set_last_error("pool_head: head too large"); return FG_ERR_INVALID; }
    pool_head_kernel<<<std::min(ceil_div(p.num_graphs, HEAD_WARPS), 148 * 8), HEAD_WARPS * 32, 0, stream>>>(p);
    FG_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace fg

// GCN forward on B200.
//
// Reference pipeline, GCN/src/GCN_compute.cc:50-102: load_input_node_embeddings, then for each layer
// l = 0..4  node_embedding_multi_pe(l) -> message_passing(l)  (GCN/src/conv_layer.cc:22-40), then finalize.
// The node transform of layer l first FINISHES layer l-1 (self term, BatchNorm, relu) and then applies
// Linear_l (GCN/src/node_embedding.cc:98-146), so the fused unit here is
//   gcn_layer_kernel<FIRST>   embedding                                   -> p_0 = W_0 h0 + b_0
//   gcn_layer_kernel<MIDDLE>  message passing over p_{l-1} + finish l-1   -> p_l = W_l a + b_l     (l = 1..4)
//   gcn_layer_kernel<FINAL>   message passing over p_4 + finish 4 (no relu) -> x      (GCN/src/finalize.cc:39-115)
// followed by the shared mean-pool + Linear(100 -> 1) head.
//
// Math: m_v = sum_{(u,v)} norm_uv relu(p_u + EE_l[attr]);
// q_v = m_v + relu(p_v + root_l)/(outdeg(v)+1);  BN(q) = ((q - mean)/sqrt(var + 2^-10)) * gamma + beta.
#include "internal.cuh"
#include "layers.cuh"

namespace fg {

namespace {

constexpr int D = 100;
constexpr int DP = 104;
constexpr int Q = D / 4;
constexpr int NT = 224;

using Gemm = TileGemm<D, DP, 4, NT>;

enum { FIRST = 0, MIDDLE = 1, FINAL = 2 };

struct GcnLayerParams {
    const float* p_in; float* p_out;
    const int* feat; const float* ne_table;                       // FIRST
    const int* in_ptr; const int* src; const uint8_t* code; const float* norm; const int* out_deg;
    const float* ee_comb;                                         // [60][100] of the layer being finished
    const float* root; const float* bn_mean; const float* bn_sqrt_var; const float* bn_weight; const float* bn_bias;
    const float* wt; const float* b;                              // Linear of the layer being started
    int num_nodes; int num_tiles;
};

struct GcnSmem {
    static constexpr int BAR = 0;
    static constexpr int PTR = 16;
    static constexpr int SRC = PTR + 4 * 80;
    static constexpr int CODE = SRC + 4 * EDGE_CAP;
    static constexpr int NORM = CODE + EDGE_CAP;
    static constexpr int TAB = NORM + 4 * EDGE_CAP;
    static constexpr int VEC = TAB + 4 * ED_COMBOS * D;            // root, mean, sqrt_var, gamma, beta: 5 x 100 floats (padded to 512)
    static constexpr int HS = VEC + 4 * 512;
    static constexpr int A = HS + 2 * 4 * TILE_M * D;
    static constexpr int WBUF = A + 4 * TILE_M * D;
    static constexpr int BYTES = WBUF + 4 * Gemm::WBUF_FLOATS;
};

template <int MODE>
__global__ void __launch_bounds__(NT, 1) gcn_layer_kernel(GcnLayerParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    using S = GcnSmem;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + S::BAR);
    TileCsr csr;
    csr.ptr = reinterpret_cast<int*>(smem + S::PTR);
    csr.src = reinterpret_cast<int*>(smem + S::SRC);
    csr.code = reinterpret_cast<uint8_t*>(smem + S::CODE);
    csr.w = reinterpret_cast<float*>(smem + S::NORM);
    float* tab = reinterpret_cast<float*>(smem + S::TAB);
    float* vec = reinterpret_cast<float*>(smem + S::VEC);
    float* hs = reinterpret_cast<float*>(smem + S::HS);
    float* As = reinterpret_cast<float*>(smem + S::A);
    float* wbuf = reinterpret_cast<float*>(smem + S::WBUF);

    const int tid = threadIdx.x;
    if (MODE != FIRST)
    {
        if (tid == 0)
        {
            mbar_init(&bar[0], 1);
            mbar_init(&bar[1], 1);
            fence_mbar_init();
        }
        for (int i = tid; i < ED_COMBOS * Q; i += NT) st_f4(tab + 4 * i, ldg_f4(p.ee_comb + 4 * i));
        for (int i = tid; i < D; i += NT)
        {
            vec[i] = __ldg(p.root + i);
            vec[100 + i] = __ldg(p.bn_mean + i);
            vec[200 + i] = __ldg(p.bn_sqrt_var + i);
            vec[300 + i] = __ldg(p.bn_weight + i);
            vec[400 + i] = __ldg(p.bn_bias + i);
        }
    }
    __syncthreads();

    int tile = blockIdx.x;
    if (MODE != FIRST && tile < p.num_tiles && tid == 0)
    {
        const int rows0 = min(TILE_M, p.num_nodes - tile * TILE_M);
        mbar_arrive_expect_tx(&bar[0], rows0 * D * 4);
        tma_load_1d(hs, p.p_in + (size_t)tile * TILE_M * D, rows0 * D * 4, &bar[0]);
    }

    const EmbedOffsets eo = concat_table_offsets();
    for (int it = 0; tile < p.num_tiles; tile += gridDim.x, it++)
    {
        const int buf = it & 1;
        const int n0 = tile * TILE_M;
        const int rows = min(TILE_M, p.num_nodes - n0);

        if (MODE == FIRST)
        {
            // layer 0 consumes the input embedding directly (GCN/src/node_embedding.cc:124-127)
            for (int item = tid; item < rows * Q; item += NT)
            {
                const int v = item / Q, q = item - v * Q;
                st_f4(As + v * D + 4 * q, embed_chunk<D>(p.feat + (size_t)(n0 + v) * ND_FEATURE, p.ne_table, eo, q));
            }
        }
        else
        {
            float* hcur = hs + buf * TILE_M * D;
            const int next = tile + gridDim.x;
            if (next < p.num_tiles && tid == 0)
            {
                const int rows_n = min(TILE_M, p.num_nodes - next * TILE_M);
                mbar_arrive_expect_tx(&bar[buf ^ 1], rows_n * D * 4);
                tma_load_1d(hs + (buf ^ 1) * TILE_M * D, p.p_in + (size_t)next * TILE_M * D, rows_n * D * 4, &bar[buf ^ 1]);
            }
            stage_tile_csr<NT, true, true>(csr, p.in_ptr, p.src, p.code, p.norm, n0, rows);
            mbar_wait(&bar[buf], (it >> 1) & 1);
            __syncthreads();

            for (int item = tid; item < rows * Q; item += NT)
            {
                const int v = item / Q, q = item - v * Q;
                const int eb = csr.ptr[v] - csr.e0, ee = csr.ptr[v + 1] - csr.e0;
                float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int e = eb; e < ee; e++)
                {
                    int u, c; float nrm;
                    if (csr.staged) { u = csr.src[e]; c = csr.code[e]; nrm = csr.w[e]; }
                    else { u = __ldg(p.src + csr.e0 + e); c = __ldg(p.code + csr.e0 + e); nrm = __ldg(p.norm + csr.e0 + e); }
                    const int ul = u - n0;
                    const float4 pu = ((unsigned)ul < (unsigned)rows) ? ld_f4(hcur + ul * D + 4 * q) : ldg_f4(p.p_in + (size_t)u * D + 4 * q);
                    const float4 t = ld_f4(tab + c * D + 4 * q);
                    m.x += nrm * relu_f(t.x + pu.x); m.y += nrm * relu_f(t.y + pu.y);
                    m.z += nrm * relu_f(t.z + pu.z); m.w += nrm * relu_f(t.w + pu.w);
                }
                // finish the layer: self term, BatchNorm (inference), relu unless this is the last layer
                const float4 pv = ld_f4(hcur + v * D + 4 * q);
                const float degp1 = (float)(__ldg(p.out_deg + n0 + v) + 1);
                const float4 rt = ld_f4(vec + 4 * q), mu = ld_f4(vec + 100 + 4 * q), sv = ld_f4(vec + 200 + 4 * q);
                const float4 ga = ld_f4(vec + 300 + 4 * q), be = ld_f4(vec + 400 + 4 * q);
                float4 a;
                a.x = (m.x + relu_f(pv.x + rt.x) / degp1 - mu.x) / sv.x * ga.x + be.x;
                a.y = (m.y + relu_f(pv.y + rt.y) / degp1 - mu.y) / sv.y * ga.y + be.y;
                a.z = (m.z + relu_f(pv.z + rt.z) / degp1 - mu.z) / sv.z * ga.z + be.z;
                a.w = (m.w + relu_f(pv.w + rt.w) / degp1 - mu.w) / sv.w * ga.w + be.w;
                if (MODE == FINAL) stg_f4_stream(p.p_out + (size_t)(n0 + v) * D + 4 * q, a);
                else st_f4(As + v * D + 4 * q, make_float4(relu_f(a.x), relu_f(a.y), relu_f(a.z), relu_f(a.w)));
            }
        }
        __syncthreads();

        if (MODE != FINAL)
        {
            const int tx = tid % Gemm::CT, ty = tid / Gemm::CT;
            float acc[8][4];
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int n = 0; n < 4; n++) acc[i][n] = 0.f;
            Gemm::run(As, D, p.wt, wbuf, acc);
            if (ty < Gemm::RT && tx * 4 < D)
            {
                const float4 bb = ldg_f4(p.b + tx * 4);
#pragma unroll
                for (int i = 0; i < 8; i++)
                {
                    const int r = ty + Gemm::RT * i;
                    if (r < rows)
                        stg_f4_stream(p.p_out + (size_t)(n0 + r) * D + tx * 4,
                                      make_float4(acc[i][0] + bb.x, acc[i][1] + bb.y, acc[i][2] + bb.z, acc[i][3] + bb.w));
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace

int gcn_forward(DeviceBatch& b, const GcnWeights& w, const RunOptions& opt, int sm_count, cudaStream_t s, int* launches)
{
    const long N = b.total_nodes;
    if (b.num_graphs == 0) return 0;
    FG_TRY(b.act[0].reserve(sizeof(float) * (size_t)N * D));
    FG_TRY(b.act[1].reserve(sizeof(float) * (size_t)N * D));
    float* h[2] = {b.act[0].as<float>(), b.act[1].as<float>()};
    FG_TRY(opt_in_smem(reinterpret_cast<const void*>(&gcn_layer_kernel<FIRST>), GcnSmem::BYTES));
    FG_TRY(opt_in_smem(reinterpret_cast<const void*>(&gcn_layer_kernel<MIDDLE>), GcnSmem::BYTES));
    FG_TRY(opt_in_smem(reinterpret_cast<const void*>(&gcn_layer_kernel<FINAL>), GcnSmem::BYTES));
    const int num_tiles = (int)ceil_div<long>(N, TILE_M);
    const int grid = min(num_tiles, sm_count);
    int nl = 0;
    for (int l = 0; l <= 5; l++)
    {
        if (opt.timer) FG_TRY(opt.timer->mark(s));
        if (opt.gcn_tc && opt.gcn_fused)
        {
            FG_TRY(gcn_step_fused_launch(b, w, l, h[(l + 1) & 1], h[l & 1], sm_count, s));
            nl++;
            continue;
        }
        if (opt.gcn_tc && l < 5)
        {
            FG_TRY(gcn_step_tc_launch(b, w, l, h[(l + 1) & 1], h[l & 1], sm_count, s));
            nl += 2;
            continue;
        }
        GcnLayerParams p{};
        p.p_in = h[(l + 1) & 1]; p.p_out = h[l & 1];
        p.feat = b.node_feature.as<int>(); p.ne_table = w.ne_table.as<float>();
        p.in_ptr = b.in_ptr.as<int>(); p.src = b.src.as<int>(); p.code = b.code.as<uint8_t>(); p.norm = b.edge_w.as<float>();
        p.out_deg = b.out_deg.as<int>();
        if (l > 0)
        {
            const size_t k = (size_t)(l - 1);
            p.ee_comb = w.ee_comb.as<float>() + k * ED_COMBOS * D;
            p.root = w.root.as<float>() + k * D; p.bn_mean = w.bn_mean.as<float>() + k * D; p.bn_sqrt_var = w.bn_sqrt_var.as<float>() + k * D;
            p.bn_weight = w.bn_weight.as<float>() + k * D; p.bn_bias = w.bn_bias.as<float>() + k * D;
        }
        if (l < 5) { p.wt = w.wt.as<float>() + (size_t)l * D * DP; p.b = w.b.as<float>() + (size_t)l * DP; }
        p.num_nodes = (int)N; p.num_tiles = num_tiles;
        if (l == 0) gcn_layer_kernel<FIRST><<<grid, NT, GcnSmem::BYTES, s>>>(p);
        else if (l < 5) gcn_layer_kernel<MIDDLE><<<grid, NT, GcnSmem::BYTES, s>>>(p);
        else gcn_layer_kernel<FINAL><<<grid, NT, GcnSmem::BYTES, s>>>(p);
        FG_CUDA(cudaGetLastError());
        nl++;
    }
    if (opt.timer) FG_TRY(opt.timer->mark(s));
    HeadParams hp{};
    hp.x = h[5 & 1]; hp.dim = D; hp.node_off = b.node_off.as<int>(); hp.nn = b.nums_of_nodes.as<int>(); hp.num_graphs = b.num_graphs;
    hp.w[0] = w.pred_w.as<float>(); hp.b[0] = w.pred_b.as<float>(); hp.dims[0] = D; hp.dims[1] = 1; hp.num_layers = 1;
    hp.out = b.out.as<float>();
    FG_TRY(launch_pool_head(hp, s));
    nl++;
    if (launches) *launches += nl;
    return 0;
}

}  // namespace fg

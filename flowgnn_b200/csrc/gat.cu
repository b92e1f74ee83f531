// GAT forward on B200 (5 layers, 4 heads x 16; activations laid out [v][dim][head] like the reference).
//
// Reference pipeline, GAT/src/GAT_compute.cc:47-108: load_graph (CSR by destination, self loop first),
// load_input_node_embeddings, then per layer MP -> adapter -> node transform
// (GAT/src/conv_layer.cc:29-133); the last layer's transform is `finalize`.
//   embed   : hproj0 = x W_proj0 (raw integer features), S0/T0 = <hproj0_head, a_src/a_tgt>      (load_inputs.cc:168-227)
//   layer l : per destination v over self + in-neighbours u:  e = S_l[v] + T_l[u];  e<0 -> 0.2e;  w = exp(e)
//             (no max subtraction); msg = sum w*hproj_u / sum w                           (message_passing.cc:94-157)
//             l < 4:  o = msg + Wskip_l o_prev; ELU; hproj' = Wproj_{l+1} o; S', T'       (node_embedding.cc:98-271)
//             l = 4:  emb = (sum_h msg + sum_ho Wskip_4 o_prev) / 4                       (finalize.cc:46-112)
// then the shared mean-pool + Linear(16 -> 1) head.
// the reference reads node features WITHOUT the per-graph offset; `feat_bug` reproduces that.
#include "internal.cuh"
#include "layers.cuh"

namespace fg {

namespace {

constexpr int HF = 64;               // heads * dims
constexpr int NH = 4;
constexpr int QF = HF / 4;           // 16 float4 per row: chunk q holds dim q, heads 0..3
constexpr int LDT = HF + 4;          // padded row stride of the shared tiles
constexpr int NT = 128;

using Gemm = TileGemm<HF, HF, 4, NT, 16>;

// ---- embedding: one warp per graph ----
__global__ void __launch_bounds__(128) gat_embed_kernel(const int* __restrict__ feat, const int* __restrict__ node_off,
                                                        const int* __restrict__ nn, int num_graphs, int feat_bug,
                                                        const float* __restrict__ proj0, const float* __restrict__ a_src,
                                                        const float* __restrict__ a_tgt, float* __restrict__ hproj,
                                                        float* __restrict__ o_prev, float* __restrict__ S, float* __restrict__ T,
                                                        const float* __restrict__ skipt0)
{
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (g >= num_graphs) return;
    const int n = nn[g];
    const size_t nb = (size_t)node_off[g];
    float w0[ND_FEATURE], w1[ND_FEATURE];
#pragma unroll
    for (int f = 0; f < ND_FEATURE; f++) { w0[f] = __ldg(proj0 + f * HF + lane); w1[f] = __ldg(proj0 + f * HF + lane + 32); }
    // gat_tc: the layer-0 skip projection of the raw features (k = 4 f: dim f, head 0) is written instead of the features themselves
    float k0[ND_FEATURE], k1[ND_FEATURE];
#pragma unroll
    for (int f = 0; f < ND_FEATURE; f++)
    {
        k0[f] = skipt0 ? __ldg(skipt0 + 4 * f * HF + lane) : 0.f;
        k1[f] = skipt0 ? __ldg(skipt0 + 4 * f * HF + lane + 32) : 0.f;
    }
    const float as0 = __ldg(a_src + lane), as1 = __ldg(a_src + lane + 32);
    const float at0 = __ldg(a_tgt + lane), at1 = __ldg(a_tgt + lane + 32);
    for (int v = 0; v < n; v++)
    {
        const int* row = feat + (feat_bug ? (size_t)v : nb + v) * ND_FEATURE;
        float x[ND_FEATURE];
#pragma unroll
        for (int f = 0; f < ND_FEATURE; f++) x[f] = (float)__ldg(row + f);
        float h0 = 0.f, h1 = 0.f;
#pragma unroll
        for (int f = 0; f < ND_FEATURE; f++) { h0 += x[f] * w0[f]; h1 += x[f] * w1[f]; }
        float* hp = hproj + (nb + v) * HF;
        float* op = o_prev + (nb + v) * HF;
        hp[lane] = h0;
        hp[lane + 32] = h1;
        // raw features sit at head 0 of dims 0..8 (GAT/src/load_inputs.cc:192-193)
        float o0 = 0.f, o1 = 0.f;
#pragma unroll
        for (int f = 0; f < ND_FEATURE; f++)
        {
            if (lane == 4 * f) o0 = x[f];
            if (lane + 32 == 4 * f) o1 = x[f];
        }
        if (skipt0)
        {
            o0 = 0.f; o1 = 0.f;
#pragma unroll
            for (int f = 0; f < ND_FEATURE; f++) { o0 = fmaf(x[f], k0[f], o0); o1 = fmaf(x[f], k1[f], o1); }      // k ascending, as TileGemm sums
        }
        op[lane] = o0;
        op[lane + 32] = o1;
        float s = h0 * as0 + h1 * as1, t = h0 * at0 + h1 * at1;     // lanes with equal lane%4 share a head
#pragma unroll
        for (int d = 4; d < 32; d <<= 1)
        {
            s += __shfl_xor_sync(0xffffffffu, s, d);
            t += __shfl_xor_sync(0xffffffffu, t, d);
        }
        if (lane < NH) { S[(nb + v) * NH + lane] = s; T[(nb + v) * NH + lane] = t; }
    }
}

struct GatLayerParams {
    const float* hproj_in; float* hproj_out;     // LAST: hproj_out receives emb [N][16]
    const float* o_in; float* o_out;
    const float* S_in; const float* T_in; float* S_out; float* T_out;
    const int* in_ptr; const int* src;
    const float* skipt; const float* projt; const float* a_src; const float* a_tgt;    // skip of layer l; proj/scores of layer l+1
    int num_nodes; int num_tiles;
};

struct GatSmem {
    static constexpr int BAR = 0;
    static constexpr int PTR = 16;
    static constexpr int SRC = PTR + 4 * 80;
    static constexpr int SS = SRC + 4 * EDGE_CAP;                 // S of the tile's nodes [TILE_M] float4
    static constexpr int TS = SS + 16 * TILE_M;
    static constexpr int HS = TS + 16 * TILE_M;                   // 2 x [TILE_M][64]
    static constexpr int M = HS + 2 * 4 * TILE_M * HF;            // messages, later the projected tile
    static constexpr int A1 = M + 4 * TILE_M * LDT;               // o_prev tile, later the activated tile
    static constexpr int WBUF = A1 + 4 * TILE_M * LDT;
    static constexpr int BYTES = WBUF + 4 * Gemm::WBUF_FLOATS;
};

template <bool LAST>
__global__ void __launch_bounds__(NT, 2) gat_layer_kernel(GatLayerParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    using SM = GatSmem;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SM::BAR);
    TileCsr csr;
    csr.ptr = reinterpret_cast<int*>(smem + SM::PTR);
    csr.src = reinterpret_cast<int*>(smem + SM::SRC);
    csr.code = nullptr; csr.w = nullptr;
    float* ss = reinterpret_cast<float*>(smem + SM::SS);
    float* ts = reinterpret_cast<float*>(smem + SM::TS);
    float* hs = reinterpret_cast<float*>(smem + SM::HS);
    float* Ms = reinterpret_cast<float*>(smem + SM::M);
    float* A1 = reinterpret_cast<float*>(smem + SM::A1);
    float* wbuf = reinterpret_cast<float*>(smem + SM::WBUF);

    const int tid = threadIdx.x;
    if (tid == 0)
    {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    int tile = blockIdx.x;
    if (tile < p.num_tiles && tid == 0)
    {
        const int rows0 = min(TILE_M, p.num_nodes - tile * TILE_M);
        mbar_arrive_expect_tx(&bar[0], rows0 * HF * 4);
        tma_load_1d(hs, p.hproj_in + (size_t)tile * TILE_M * HF, rows0 * HF * 4, &bar[0]);
    }

    for (int it = 0; tile < p.num_tiles; tile += gridDim.x, it++)
    {
        const int buf = it & 1;
        const int n0 = tile * TILE_M;
        const int rows = min(TILE_M, p.num_nodes - n0);
        float* hcur = hs + buf * TILE_M * HF;
        const int next = tile + gridDim.x;
        if (next < p.num_tiles && tid == 0)
        {
            const int rows_n = min(TILE_M, p.num_nodes - next * TILE_M);
            mbar_arrive_expect_tx(&bar[buf ^ 1], rows_n * HF * 4);
            tma_load_1d(hs + (buf ^ 1) * TILE_M * HF, p.hproj_in + (size_t)next * TILE_M * HF, rows_n * HF * 4, &bar[buf ^ 1]);
        }
        for (int i = tid; i < rows; i += NT)
        {
            st_f4(ss + 4 * i, ldg_f4(p.S_in + (size_t)(n0 + i) * NH));
            st_f4(ts + 4 * i, ldg_f4(p.T_in + (size_t)(n0 + i) * NH));
        }
        for (int i = tid; i < rows * QF; i += NT)
        {
            const int v = i / QF, q = i - v * QF;
            st_f4(A1 + v * LDT + 4 * q, ldg_f4(p.o_in + (size_t)(n0 + v) * HF + 4 * q));
        }
        stage_tile_csr<NT, false, false>(csr, p.in_ptr, p.src, nullptr, nullptr, n0, rows);
        mbar_wait(&bar[buf], (it >> 1) & 1);
        __syncthreads();

        // ---- attention-weighted gather (self loop first, then in-edges in CSR order) ----
        for (int item = tid; item < rows * QF; item += NT)
        {
            const int v = item / QF, q = item - v * QF;
            const float4 sv = ld_f4(ss + 4 * v);
            float4 num = make_float4(0.f, 0.f, 0.f, 0.f), den = num;
            const int eb = csr.ptr[v] - csr.e0, ee = csr.ptr[v + 1] - csr.e0;
            for (int e = eb - 1; e < ee; e++)
            {
                int ul;
                float4 tu, hu;
                if (e < eb) ul = v;
                else ul = (csr.staged ? csr.src[e] : __ldg(p.src + csr.e0 + e)) - n0;
                if ((unsigned)ul < (unsigned)rows) { tu = ld_f4(ts + 4 * ul); hu = ld_f4(hcur + ul * HF + 4 * q); }
                else
                {
                    const size_t u = (size_t)(ul + n0);
                    tu = ldg_f4(p.T_in + u * NH);
                    hu = ldg_f4(p.hproj_in + u * HF + 4 * q);
                }
                float4 sc = make_float4(sv.x + tu.x, sv.y + tu.y, sv.z + tu.z, sv.w + tu.w);
                sc.x = (sc.x < 0.f) ? sc.x * 0.2f : sc.x; sc.y = (sc.y < 0.f) ? sc.y * 0.2f : sc.y;
                sc.z = (sc.z < 0.f) ? sc.z * 0.2f : sc.z; sc.w = (sc.w < 0.f) ? sc.w * 0.2f : sc.w;
                const float4 w = make_float4(expf(sc.x), expf(sc.y), expf(sc.z), expf(sc.w));
                den.x += w.x; den.y += w.y; den.z += w.z; den.w += w.w;
                num.x += w.x * hu.x; num.y += w.y * hu.y; num.z += w.z * hu.z; num.w += w.w * hu.w;
            }
            st_f4(Ms + v * LDT + 4 * q, make_float4(num.x / den.x, num.y / den.y, num.z / den.z, num.w / den.w));
        }
        __syncthreads();

        // ---- skip projection of the previous activations ----
        const int tx = tid % Gemm::CT, ty = tid / Gemm::CT;
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int n = 0; n < 4; n++) acc[i][n] = 0.f;
        Gemm::run(A1, LDT, p.skipt, wbuf, acc);

        if (LAST)
        {
            // emb[d] = (sum_h msg[d][h] + sum_ho skip[d][ho]) / NUM_HEADS; thread columns = (d = tx, ho = 0..3)
#pragma unroll
            for (int i = 0; i < 8; i++)
            {
                const int r = ty + Gemm::RT * i;
                if (r < rows)
                {
                    const float4 m = ld_f4(Ms + r * LDT + 4 * tx);
                    float of = 0.f;
                    of += m.x; of += m.y; of += m.z; of += m.w;
                    of += acc[i][0]; of += acc[i][1]; of += acc[i][2]; of += acc[i][3];
                    p.hproj_out[(size_t)(n0 + r) * 16 + tx] = of / 4.0f;
                }
            }
        }
        else
        {
            // o = msg + skip; ELU (o <= 0 -> exp(o) - 1); keep for the next layer and as the next GEMM's input
#pragma unroll
            for (int i = 0; i < 8; i++)
            {
                const int r = ty + Gemm::RT * i;
                const float4 m = ld_f4(Ms + r * LDT + 4 * tx);
                float4 o = make_float4(m.x + acc[i][0], m.y + acc[i][1], m.z + acc[i][2], m.w + acc[i][3]);
                o.x = (o.x <= 0.f) ? expf(o.x) - 1.0f : o.x; o.y = (o.y <= 0.f) ? expf(o.y) - 1.0f : o.y;
                o.z = (o.z <= 0.f) ? expf(o.z) - 1.0f : o.z; o.w = (o.w <= 0.f) ? expf(o.w) - 1.0f : o.w;
                st_f4(A1 + r * LDT + 4 * tx, o);
                if (r < rows) stg_f4_stream(p.o_out + (size_t)(n0 + r) * HF + 4 * tx, o);
            }
            __syncthreads();
            // ---- next layer's projection ----
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int n = 0; n < 4; n++) acc[i][n] = 0.f;
            Gemm::run(A1, LDT, p.projt, wbuf, acc);
#pragma unroll
            for (int i = 0; i < 8; i++)
            {
                const int r = ty + Gemm::RT * i;
                const float4 hp = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
                st_f4(Ms + r * LDT + 4 * tx, hp);
                if (r < rows) stg_f4_stream(p.hproj_out + (size_t)(n0 + r) * HF + 4 * tx, hp);
            }
            __syncthreads();
            // ---- next layer's scores: S[v][h] = sum_d hproj'[v][d][h] a_src[d][h], d ascending ----
            for (int item = tid; item < rows * NH; item += NT)
            {
                const int v = item / NH, h = item - v * NH;
                float s = 0.f, t = 0.f;
#pragma unroll
                for (int d = 0; d < 16; d++)
                {
                    const float r = Ms[v * LDT + d * NH + h];
                    s = r * __ldg(p.a_src + d * NH + h) + s;
                    t = r * __ldg(p.a_tgt + d * NH + h) + t;
                }
                p.S_out[(size_t)(n0 + v) * NH + h] = s;
                p.T_out[(size_t)(n0 + v) * NH + h] = t;
            }
        }
        __syncthreads();
    }
}

}  // namespace

int gat_forward(DeviceBatch& b, const GatWeights& w, const RunOptions& opt, int sm_count, cudaStream_t s, int* launches)
{
    const long N = b.total_nodes;
    if (b.num_graphs == 0) return 0;
    for (int i = 0; i < 4; i++) FG_TRY(b.act[i].reserve(sizeof(float) * (size_t)N * HF));
    for (int i = 0; i < 4; i++) FG_TRY(b.score[i].reserve(sizeof(float) * (size_t)N * NH));
    float* hp[2] = {b.act[0].as<float>(), b.act[1].as<float>()};
    float* o[2] = {b.act[2].as<float>(), b.act[3].as<float>()};
    float* S[2] = {b.score[0].as<float>(), b.score[1].as<float>()};
    float* T[2] = {b.score[2].as<float>(), b.score[3].as<float>()};
    int nl = 0;
    gat_embed_kernel<<<ceil_div(b.num_graphs, 4), 128, 0, s>>>(b.node_feature.as<int>(), b.node_off.as<int>(), b.nums_of_nodes.as<int>(),
                                                             b.num_graphs, opt.gat_node_offset_bug ? 1 : 0, w.proj0.as<float>(),
                                                             w.a_src.as<float>(), w.a_tgt.as<float>(), hp[0], o[0], S[0], T[0],
                                                             opt.gat_tc ? w.skipt.as<float>() : nullptr);
    FG_CUDA(cudaGetLastError());
    nl++;
    if (opt.gat_tc)
    {
        // o[] holds skip_l instead of o_{l-1}: the fused kernel of layer l produces hproj, skip and the scores of layer l + 1 (gat_tc.cu)
        for (int l = 0; l < 4; l++)
        {
            if (opt.timer) FG_TRY(opt.timer->mark(s));
            FG_TRY(gat_layer_tc_launch(b, w, l, hp[l & 1], o[l & 1], S[l & 1], T[l & 1], hp[(l + 1) & 1], o[(l + 1) & 1], S[(l + 1) & 1], T[(l + 1) & 1],
                                       sm_count, s));
            nl++;
        }
        if (opt.timer) FG_TRY(opt.timer->mark(s));
        FG_TRY(gat_final_launch(b, hp[0], o[0], S[0], T[0], hp[1], sm_count, s));
        nl++;
        if (opt.timer) FG_TRY(opt.timer->mark(s));
        HeadParams hd{};
        hd.x = hp[1]; hd.dim = 16; hd.node_off = b.node_off.as<int>(); hd.nn = b.nums_of_nodes.as<int>(); hd.num_graphs = b.num_graphs;
        hd.w[0] = w.pred_w.as<float>(); hd.b[0] = w.pred_b.as<float>(); hd.dims[0] = 16; hd.dims[1] = 1; hd.num_layers = 1;
        hd.out = b.out.as<float>();
        FG_TRY(launch_pool_head(hd, s));
        nl++;
        if (launches) *launches += nl;
        return 0;
    }
    FG_TRY(opt_in_smem(reinterpret_cast<const void*>(&gat_layer_kernel<false>), GatSmem::BYTES));
    FG_TRY(opt_in_smem(reinterpret_cast<const void*>(&gat_layer_kernel<true>), GatSmem::BYTES));
    const int num_tiles = (int)ceil_div<long>(N, TILE_M);
    const int grid = min(num_tiles, sm_count * 2);
    for (int l = 0; l < 5; l++)
    {
        if (opt.timer) FG_TRY(opt.timer->mark(s));
        GatLayerParams p{};
        p.hproj_in = hp[l & 1]; p.hproj_out = hp[(l + 1) & 1];
        p.o_in = o[l & 1]; p.o_out = o[(l + 1) & 1];
        p.S_in = S[l & 1]; p.T_in = T[l & 1]; p.S_out = S[(l + 1) & 1]; p.T_out = T[(l + 1) & 1];
        p.in_ptr = b.in_ptr.as<int>(); p.src = b.src.as<int>();
        p.skipt = w.skipt.as<float>() + (size_t)l * HF * HF;
        if (l < 4)
        {
            p.projt = w.projt.as<float>() + (size_t)(l + 1) * HF * HF;
            p.a_src = w.a_src.as<float>() + (size_t)(l + 1) * HF;
            p.a_tgt = w.a_tgt.as<float>() + (size_t)(l + 1) * HF;
        }
        p.num_nodes = (int)N; p.num_tiles = num_tiles;
        if (l < 4) gat_layer_kernel<false><<<grid, NT, GatSmem::BYTES, s>>>(p);
        else gat_layer_kernel<true><<<grid, NT, GatSmem::BYTES, s>>>(p);
        FG_CUDA(cudaGetLastError());
        nl++;
    }
    if (opt.timer) FG_TRY(opt.timer->mark(s));
    HeadParams hd{};
    hd.x = hp[1]; hd.dim = 16; hd.node_off = b.node_off.as<int>(); hd.nn = b.nums_of_nodes.as<int>(); hd.num_graphs = b.num_graphs;
    hd.w[0] = w.pred_w.as<float>(); hd.b[0] = w.pred_b.as<float>(); hd.dims[0] = 16; hd.dims[1] = 1; hd.num_layers = 1;
    hd.out = b.out.as<float>();
    FG_TRY(launch_pool_head(hd, s));
    nl++;
    if (launches) *launches += nl;
    return 0;
}

}  // namespace fg

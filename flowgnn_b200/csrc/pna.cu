// PNA forward on B200.
//
// Reference pipeline, PNA/src/PNA_compute.cc:44-98: stage 0 embeds; stages 1..4 run
// node_embedding_multi_pe(l-1) then message passing / finalize (PNA/src/conv_layer.cc:37-104).
// Fused unit here, one launch per layer l = 0..3:
//   message passing: per (v, d) the four planes  S = sum h_u, Q = sum h_u^2, min, max over in-edges,
//                    initialised 0, 0, +(32 - 2^-10), -32   (PNA/src/message_passing.cc:88-147, util.h:34-46)
//   node transform : mean = S/n, std = sqrt(relu(Q/n - mean^2)), n = max(indeg, 1);
//                    t = log(outdeg+1)/avg_deg, s = avg_deg/log(outdeg+1) (s == 0 -> 1);
//                    acc = b + sum_in [T0 + T1 t + T2 s], T_k = sum_aggr w[k][aggr] * aggr;  h <- h + relu(acc)
//                    (PNA/src/node_embedding.cc:106-215)
// The 12-way weighted sum is one [rows x 320] x [320 x 240] GEMM (k = aggr*80 + in, n = scaler*80 + out)
// followed by out = b + G0 + t G1 + s G2.  A node with out-degree 0 has s = inf; for it the reference's
// per-input association decides between inf and NaN, so such rows are evaluated by a
// warp that follows the reference's expression term by term.
#include "internal.cuh"
#include "layers.cuh"

#include <algorithm>

namespace fg {

namespace {

constexpr int D = 80;
constexpr int Q = D / 4;
constexpr int KA = 4 * D;        // 320: [mean | min | max | std] x 80   (aggregator_t order, PNA/src/dcl.h:29-35)
constexpr int NC = 3 * D;        // 240: [none | t | scale] x 80          (scaler_t order, PNA/src/dcl.h:37-42)
constexpr int NT = 256;
constexpr int LDA = KA + 4;   // row stride of the aggregate tile (bank spread)

using Gemm = TileGemm<KA, NC, 8, NT>;

struct PnaLayerParams {
    const float* h_in; float* h_out;
    const int* in_ptr; const int* src; const int* out_deg;
    const float* wcat; const float* w_ref; const float* b;
    float avg_deg;
    int num_nodes; int num_tiles;
};

struct PnaSmem {
    static constexpr int BAR = 0;
    static constexpr int PTR = 16;
    static constexpr int SRC = PTR + 4 * 80;
    static constexpr int HS = SRC + 4 * EDGE_CAP;
    static constexpr int A = HS + 2 * 4 * TILE_M * D;          // [TILE_M][320]; reused as Z [TILE_M][240] after the GEMM
    static constexpr int WBUF = A + 4 * TILE_M * LDA;
    static constexpr int BYTES = WBUF + 4 * Gemm::WBUF_FLOATS;
};

__global__ void __launch_bounds__(NT, 1) pna_layer_kernel(PnaLayerParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    using S = PnaSmem;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + S::BAR);
    TileCsr csr;
    csr.ptr = reinterpret_cast<int*>(smem + S::PTR);
    csr.src = reinterpret_cast<int*>(smem + S::SRC);
    csr.code = nullptr; csr.w = nullptr;
    float* hs = reinterpret_cast<float*>(smem + S::HS);
    float* As = reinterpret_cast<float*>(smem + S::A);
    float* wbuf = reinterpret_cast<float*>(smem + S::WBUF);

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0)
    {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    // ap_fixed_max / ap_fixed_min of ap_fixed<16,6>
    const float fm_max = 32.0f - 0.0009765625f, fm_min = -32.0f;

    int tile = blockIdx.x;
    if (tile < p.num_tiles && tid == 0)
    {
        const int rows0 = min(TILE_M, p.num_nodes - tile * TILE_M);
        mbar_arrive_expect_tx(&bar[0], rows0 * D * 4);
        tma_load_1d(hs, p.h_in + (size_t)tile * TILE_M * D, rows0 * D * 4, &bar[0]);
    }

    for (int it = 0; tile < p.num_tiles; tile += gridDim.x, it++)
    {
        const int buf = it & 1;
        const int n0 = tile * TILE_M;
        const int rows = min(TILE_M, p.num_nodes - n0);
        float* hcur = hs + buf * TILE_M * D;
        const int next = tile + gridDim.x;
        if (next < p.num_tiles && tid == 0)
        {
            const int rows_n = min(TILE_M, p.num_nodes - next * TILE_M);
            mbar_arrive_expect_tx(&bar[buf ^ 1], rows_n * D * 4);
            tma_load_1d(hs + (buf ^ 1) * TILE_M * D, p.h_in + (size_t)next * TILE_M * D, rows_n * D * 4, &bar[buf ^ 1]);
        }
        stage_tile_csr<NT, false, false>(csr, p.in_ptr, p.src, nullptr, nullptr, n0, rows);
        mbar_wait(&bar[buf], (it >> 1) & 1);
        __syncthreads();

        // ---- message passing: four planes per (v, d), then mean / min / max / std ----
        for (int item = tid; item < rows * Q; item += NT)
        {
            const int v = item / Q, q = item - v * Q;
            const int eb = csr.ptr[v] - csr.e0, ee = csr.ptr[v + 1] - csr.e0;
            float s[4] = {0.f, 0.f, 0.f, 0.f}, sq[4] = {0.f, 0.f, 0.f, 0.f};
            float mn[4] = {fm_max, fm_max, fm_max, fm_max}, mx[4] = {fm_min, fm_min, fm_min, fm_min};
            for (int e = eb; e < ee; e++)
            {
                const int u = csr.staged ? csr.src[e] : __ldg(p.src + csr.e0 + e);
                const int ul = u - n0;
                const float4 hu = ((unsigned)ul < (unsigned)rows) ? ld_f4(hcur + ul * D + 4 * q) : ldg_f4(p.h_in + (size_t)u * D + 4 * q);
                const float x[4] = {hu.x, hu.y, hu.z, hu.w};
#pragma unroll
                for (int j = 0; j < 4; j++)
                {
                    s[j] += x[j];
                    sq[j] += x[j] * x[j];
                    if (x[j] < mn[j]) mn[j] = x[j];
                    if (x[j] > mx[j]) mx[j] = x[j];
                }
            }
            int in_deg = ee - eb;
            if (in_deg == 0) in_deg = 1;
            const float fn = (float)in_deg;
            float mean[4], sd[4];
#pragma unroll
            for (int j = 0; j < 4; j++)
            {
                mean[j] = s[j] / fn;
                sd[j] = sqrtf(relu_f(sq[j] / fn - mean[j] * mean[j]));
            }
            float* a = As + v * LDA + 4 * q;
            st_f4(a, make_float4(mean[0], mean[1], mean[2], mean[3]));
            st_f4(a + D, make_float4(mn[0], mn[1], mn[2], mn[3]));
            st_f4(a + 2 * D, make_float4(mx[0], mx[1], mx[2], mx[3]));
            st_f4(a + 3 * D, make_float4(sd[0], sd[1], sd[2], sd[3]));
        }
        __syncthreads();

        // ---- rows whose out-degree is 0: follow the reference's expression literally (one warp per row) ----
        for (int v = wid; v < rows; v += NT / 32)
        {
            if (__ldg(p.out_deg + n0 + v) != 0) continue;
            const float log_degree = logf(1.0f);
            const float t = log_degree / p.avg_deg;
            float scale = p.avg_deg / log_degree;
            if (scale == 0) scale = 1;
            const float* a = As + v * LDA;
            for (int o = lane; o < D; o += 32)
            {
                float acc = 0.f;
                for (int i = 0; i < D; i++)
                {
                    const float mean = a[i], mnv = a[D + i], mxv = a[2 * D + i], sdv = a[3 * D + i];
                    const float* w = p.w_ref + (size_t)o * 12 * D + i;       // [scaler][aggr][in], aggr: mean, min, max, std
#define WREF(sc, ag) __ldg(w + ((sc) * 4 + (ag)) * D)
                    const float t0 = __fadd_rn(__fadd_rn(__fmul_rn(mean, WREF(0, 0)), __fmul_rn(sdv, WREF(0, 3))),
                                               __fadd_rn(__fmul_rn(mnv, WREF(0, 1)), __fmul_rn(mxv, WREF(0, 2))));
                    const float t1 = __fadd_rn(__fadd_rn(__fmul_rn(mean, WREF(1, 0)), __fmul_rn(sdv, WREF(1, 3))),
                                               __fadd_rn(__fmul_rn(mnv, WREF(1, 1)), __fmul_rn(mxv, WREF(1, 2))));
                    const float t2 = __fadd_rn(__fadd_rn(__fmul_rn(mean, WREF(2, 0)), __fmul_rn(sdv, WREF(2, 3))),
                                               __fadd_rn(__fmul_rn(mnv, WREF(2, 1)), __fmul_rn(mxv, WREF(2, 2))));
#undef WREF
                    const float addend = __fadd_rn(t0, __fadd_rn(__fmul_rn(t1, t), __fmul_rn(t2, scale)));
                    acc = __fadd_rn(addend, (i == 0) ? __ldg(p.b + o) : acc);
                }
                p.h_out[(size_t)(n0 + v) * D + o] = hcur[v * D + o] + relu_f(acc);
            }
        }

        // ---- [rows x 320] x [320 x 240] ----
        const int tx = tid % Gemm::CT, ty = tid / Gemm::CT;
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int n = 0; n < 8; n++) acc[i][n] = 0.f;
        Gemm::run(As, LDA, p.wcat, wbuf, acc);
        float* Zs = As;                              // A is dead after the GEMM's final barrier
        if (ty < Gemm::RT)
        {
#pragma unroll
            for (int i = 0; i < 8; i++)
            {
                float* z = Zs + (ty + Gemm::RT * i) * NC;
                st_f4(z + Gemm::col(tx, 0), make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
                st_f4(z + Gemm::col(tx, 1), make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]));
            }
        }
        __syncthreads();

        // ---- combine scalers, relu, residual ----
        for (int item = tid; item < rows * Q; item += NT)
        {
            const int v = item / Q, q = item - v * Q;
            const int od = __ldg(p.out_deg + n0 + v);
            if (od == 0) continue;                    // written by the exact path above
            const float log_degree = logf((float)(od + 1));
            const float t = log_degree / p.avg_deg;
            float scale = p.avg_deg / log_degree;
            if (scale == 0) scale = 1;
            const float* z = Zs + v * NC + 4 * q;
            const float4 g0 = ld_f4(z), g1 = ld_f4(z + D), g2 = ld_f4(z + 2 * D);
            const float4 bb = ldg_f4(p.b + 4 * q);
            const float4 hv = ld_f4(hcur + v * D + 4 * q);
            float4 o;
            o.x = hv.x + relu_f(bb.x + g0.x + (g1.x * t + g2.x * scale));
            o.y = hv.y + relu_f(bb.y + g0.y + (g1.y * t + g2.y * scale));
            o.z = hv.z + relu_f(bb.z + g0.z + (g1.z * t + g2.z * scale));
            o.w = hv.w + relu_f(bb.w + g0.w + (g1.w * t + g2.w * scale));
            stg_f4_stream(p.h_out + (size_t)(n0 + v) * D + 4 * q, o);
        }
        __syncthreads();
    }
}

}  // namespace

int pna_forward(DeviceBatch& b, const PnaWeights& w, const RunOptions& opt, int sm_count, cudaStream_t s, int* launches)
{
    const long N = b.total_nodes;
    if (b.num_graphs == 0) return 0;
    FG_TRY(b.act[0].reserve(sizeof(float) * (size_t)N * D));
    FG_TRY(b.act[1].reserve(sizeof(float) * (size_t)N * D));
    float* h[2] = {b.act[0].as<float>(), b.act[1].as<float>()};
    int nl = 0;
    {
        const int blocks = (int)std::min<long>(ceil_div<long>(N, 8), (long)sm_count * 8);      // 8 warps per block, a warp per node
        const cudaStream_t es = embed_stream(opt, s);      // overlaps the CSR / tile build (api.cu::compute_on)
        if (b.perm_active) { FG_TRY(node_map_launch(b, es)); nl++; }      // re-ordered graphs: row -> caller-order node
        embed4_kernel<D><<<blocks, 256, 0, es>>>(b.node_feature.as<int>(), w.ne_table.as<float>(), w.ne_table4.as<float>(), h[0], N,
                                                 b.perm_active ? b.node_map.as<int>() : nullptr);
        FG_CUDA(cudaGetLastError());
        FG_TRY(embed_join(opt, es, s));
        nl++;
    }
    FG_TRY(opt_in_smem(reinterpret_cast<const void*>(&pna_layer_kernel), PnaSmem::BYTES));
    const int num_tiles = (int)ceil_div<long>(N, TILE_M);
    const int grid = min(num_tiles, sm_count);
    for (int l = 0; l < 4; l++)
    {
        if (opt.timer) FG_TRY(opt.timer->mark(s));
        if (opt.pna_fused && opt.pna_tc)
        {
            // one kernel for message passing + node transform, then the few rows that need fp32
            FG_TRY(b.nonfinite.reserve((size_t)N + 16));
            FG_TRY(zero_bytes_launch(b.nonfinite.ptr, (size_t)N, s));
            FG_TRY(pna_layer_fused_launch(b, w, l, h[l & 1], h[(l + 1) & 1], sm_count, s));
            FG_TRY(pna_exact_rows_launch(b, w, l, h[l & 1], h[(l + 1) & 1], sm_count, s));
            nl += 3;                                   // zero the flags, the fused layer, the exact rows
            continue;
        }
        if (opt.pna_tc)
        {
            FG_TRY(pna_layer_tc_launch(b, w, l, h[l & 1], h[(l + 1) & 1], sm_count, s));
            nl += 4;
            continue;
        }
        PnaLayerParams p{};
        p.h_in = h[l & 1]; p.h_out = h[(l + 1) & 1];
        p.in_ptr = b.in_ptr.as<int>(); p.src = b.src.as<int>(); p.out_deg = b.out_deg.as<int>();
        p.wcat = w.wcat.as<float>() + (size_t)l * KA * NC;
        p.w_ref = w.w_ref.as<float>() + (size_t)l * D * 12 * D;
        p.b = w.b.as<float>() + (size_t)l * D;
        p.avg_deg = w.avg_deg;
        p.num_nodes = (int)N; p.num_tiles = num_tiles;
        pna_layer_kernel<<<grid, NT, PnaSmem::BYTES, s>>>(p);
        FG_CUDA(cudaGetLastError());
        nl++;
    }
    if (opt.timer) FG_TRY(opt.timer->mark(s));
    HeadParams hp{};
    hp.x = h[0]; hp.dim = D; hp.node_off = b.node_off.as<int>(); hp.nn = b.nums_of_nodes.as<int>(); hp.num_graphs = b.num_graphs;
    hp.w[0] = w.m1w.as<float>(); hp.b[0] = w.m1b.as<float>();
    hp.w[1] = w.m2w.as<float>(); hp.b[1] = w.m2b.as<float>();
    hp.w[2] = w.m3w.as<float>(); hp.b[2] = w.m3b.as<float>();
    hp.dims[0] = D; hp.dims[1] = 40; hp.dims[2] = 20; hp.dims[3] = 1; hp.num_layers = 3;
    hp.out = b.out.as<float>();
    FG_TRY(launch_pool_head(hp, s));
    nl++;
    if (launches) *launches += nl;
    return 0;
}

}  // namespace fg

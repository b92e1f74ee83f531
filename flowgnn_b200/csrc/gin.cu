// GIN / GIN-VN forward on B200.
//
// Replaces the reference's per-graph pipeline GIN/src/GIN_compute.cc:44-98:
//   stage 0      load_input_node_embeddings        (load_inputs.cc:174-220)      -> embed_table_kernel
//   stages 1..5  message_passing_pe scatter        (message_passing.cc:77-150)  \_ gin_layer_kernel
//                node_embedding_multi_pe MLP       (node_embedding.cc:23-201)   /  (one launch per layer)
//   stage 5      finalize: mean pool + Linear      (finalize.cc:14-115)          -> pool_head_kernel
// GIN-VN is the same kernel on host-augmented graphs.
//
// Math per layer l:  m_v = sum_{(u,v)} relu(h_u + EE_l[attr_uv]);
// z = W1_l (m_v + h_v) + b1_l;  h_v <- W2_l relu(z) + b2_l  (+ relu unless l == 4).
// eps is never loaded by the reference kernel, so (1 + eps) == 1.
#include "internal.cuh"
#include "layers.cuh"

#include <algorithm>

namespace fg {

namespace {

constexpr int D = 100;               // EMB_DIM
constexpr int H = 200;               // MLP_1_OUT
constexpr int HP = 208;              // H padded to the 8-wide thread tile
constexpr int DP = 104;              // D padded to the 4-wide thread tile
constexpr int Q = D / 4;             // float4 chunks per row
constexpr int NT = 224;              // threads per CTA: 26 column-threads x 8 row-threads (+16 copy helpers)
constexpr int LDZ = 204;             // leading dimension of the hidden tile (bank spread, multiple of 4)

using Gemm1 = TileGemm<D, HP, 8, NT>;
using Gemm2 = TileGemm<H, DP, 4, NT>;

struct GinLayerParams {
    const float* h_in; float* h_out;
    const int* in_ptr; const int* src; const uint8_t* code;
    const float* ee_comb;            // [60][100] this layer
    const float* w1t; const float* b1; const float* w2t; const float* b2;
    int num_nodes; int num_tiles; int relu_out;
};

template <bool MP_ONLY>
struct GinSmem {
    // byte offsets into dynamic shared memory (all multiples of 16)
    static constexpr int BAR = 0;                                   // 2 mbarriers
    static constexpr int PTR = 16;                                  // (TILE_M + 1) ints, padded
    static constexpr int SRC = PTR + 4 * 80;
    static constexpr int CODE = SRC + 4 * EDGE_CAP;
    static constexpr int TAB = CODE + EDGE_CAP;
    static constexpr int HS = TAB + 4 * ED_COMBOS * D;              // 2 x [TILE_M][D]
    static constexpr int A = HS + 2 * 4 * TILE_M * D;
    static constexpr int Z = A + 4 * TILE_M * D;
    static constexpr int WBUF = Z + 4 * TILE_M * LDZ;
    static constexpr int END_FULL = WBUF + 4 * (Gemm1::WBUF_FLOATS > Gemm2::WBUF_FLOATS ? Gemm1::WBUF_FLOATS : Gemm2::WBUF_FLOATS);
    static constexpr int BYTES = MP_ONLY ? A : END_FULL;
};

template <bool MP_ONLY>
__global__ void __launch_bounds__(NT, 1) gin_layer_kernel(GinLayerParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    using S = GinSmem<MP_ONLY>;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + S::BAR);
    TileCsr csr;
    csr.ptr = reinterpret_cast<int*>(smem + S::PTR);
    csr.src = reinterpret_cast<int*>(smem + S::SRC);
    csr.code = reinterpret_cast<uint8_t*>(smem + S::CODE);
    csr.w = nullptr;
    float* tab = reinterpret_cast<float*>(smem + S::TAB);
    float* hs = reinterpret_cast<float*>(smem + S::HS);
    float* As = reinterpret_cast<float*>(smem + S::A);
    float* Zs = reinterpret_cast<float*>(smem + S::Z);
    float* wbuf = reinterpret_cast<float*>(smem + S::WBUF);

    const int tid = threadIdx.x;
    if (tid == 0)
    {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }
    for (int i = tid; i < ED_COMBOS * Q; i += NT) st_f4(tab + 4 * i, ldg_f4(p.ee_comb + 4 * i));
    __syncthreads();

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

        // stream the next tile's rows in while this one computes
        const int next = tile + gridDim.x;
        if (next < p.num_tiles && tid == 0)
        {
            const int rows_n = min(TILE_M, p.num_nodes - next * TILE_M);
            mbar_arrive_expect_tx(&bar[buf ^ 1], rows_n * D * 4);
            tma_load_1d(hs + (buf ^ 1) * TILE_M * D, p.h_in + (size_t)next * TILE_M * D, rows_n * D * 4, &bar[buf ^ 1]);
        }

        stage_tile_csr<NT, true, false>(csr, p.in_ptr, p.src, p.code, nullptr, n0, rows);
        mbar_wait(&bar[buf], (it >> 1) & 1);
        __syncthreads();

        // ---- message passing: gather + edge embedding + relu + segmented sum, CSR order ----
        for (int item = tid; item < rows * Q; item += NT)
        {
            const int v = item / Q, q = item - v * Q;
            const int eb = csr.ptr[v] - csr.e0, ee = csr.ptr[v + 1] - csr.e0;
            float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int e = eb; e < ee; e++)
            {
                int u, c;
                if (csr.staged) { u = csr.src[e]; c = csr.code[e]; }
                else { u = __ldg(p.src + csr.e0 + e); c = __ldg(p.code + csr.e0 + e); }
                const int ul = u - n0;
                const float4 hu = ((unsigned)ul < (unsigned)rows) ? ld_f4(hcur + ul * D + 4 * q) : ldg_f4(p.h_in + (size_t)u * D + 4 * q);
                const float4 t = ld_f4(tab + c * D + 4 * q);
                m.x += relu_f(t.x + hu.x); m.y += relu_f(t.y + hu.y); m.z += relu_f(t.z + hu.z); m.w += relu_f(t.w + hu.w);
            }
            const float4 hv = ld_f4(hcur + v * D + 4 * q);
            const float4 a = make_float4(m.x + hv.x, m.y + hv.y, m.z + hv.z, m.w + hv.w);
            if (MP_ONLY) stg_f4_stream(p.h_out + (size_t)(n0 + v) * D + 4 * q, a);
            else st_f4(As + v * D + 4 * q, a);
        }
        __syncthreads();

        if (!MP_ONLY)
        {
            const int tx = tid % 26, ty = tid / 26;
            // ---- z = relu(W1 a + b1) ----
            {
                float acc[8][8];
#pragma unroll
                for (int i = 0; i < 8; i++)
#pragma unroll
                    for (int n = 0; n < 8; n++) acc[i][n] = 0.f;
                Gemm1::run(As, D, p.w1t, wbuf, acc);
                if (ty < Gemm1::RT)
                {
                    // two groups of four hidden columns per thread (layers.cuh); the second group of the last two column
                    // threads is padding (columns 200..207)
                    const int c0 = Gemm1::col(tx, 0), c1 = Gemm1::col(tx, 1);
                    const float4 ba = ldg_f4(p.b1 + c0), bb = ldg_f4(p.b1 + c1);
#pragma unroll
                    for (int i = 0; i < 8; i++)
                    {
                        float* z = Zs + (ty + Gemm1::RT * i) * LDZ;
                        st_f4(z + c0, make_float4(relu_f(acc[i][0] + ba.x), relu_f(acc[i][1] + ba.y), relu_f(acc[i][2] + ba.z), relu_f(acc[i][3] + ba.w)));
                        if (c1 < H)
                            st_f4(z + c1, make_float4(relu_f(acc[i][4] + bb.x), relu_f(acc[i][5] + bb.y), relu_f(acc[i][6] + bb.z), relu_f(acc[i][7] + bb.w)));
                    }
                }
            }
            __syncthreads();
            // ---- h' = W2 z + b2 (relu unless last layer) ----
            {
                float acc[8][4];
#pragma unroll
                for (int i = 0; i < 8; i++)
#pragma unroll
                    for (int n = 0; n < 4; n++) acc[i][n] = 0.f;
                Gemm2::run(Zs, LDZ, p.w2t, wbuf, acc);
                if (ty < Gemm2::RT && tx * 4 < D)
                {
                    const float4 bb = ldg_f4(p.b2 + tx * 4);
#pragma unroll
                    for (int i = 0; i < 8; i++)
                    {
                        const int r = ty + Gemm2::RT * i;
                        if (r < rows)
                        {
                            float4 o = make_float4(acc[i][0] + bb.x, acc[i][1] + bb.y, acc[i][2] + bb.z, acc[i][3] + bb.w);
                            if (p.relu_out) o = make_float4(relu_f(o.x), relu_f(o.y), relu_f(o.z), relu_f(o.w));
                            stg_f4_stream(p.h_out + (size_t)(n0 + r) * D + tx * 4, o);
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ---- the edge gather-scatter on its own ("mp_only": node transform = identity, h <- m + h) --------------------
// One warp per destination row: lanes 0..24 own one float4 chunk each, so every feature row (own and source rows)
// is read as one coalesced 400-byte access, the 60 combined edge-embedding rows sit in shared memory, in-edges are
// added in CSR order.  This is the HBM-roofline variant of BASELINE.json's metric.
constexpr int GATHER_WARPS = 8;

__global__ void __launch_bounds__(GATHER_WARPS * 32) gin_gather_kernel(const float* __restrict__ h_in, float* __restrict__ h_out,
                                                                       const int* __restrict__ in_ptr, const int* __restrict__ src,
                                                                       const uint8_t* __restrict__ code, const float* __restrict__ ee_comb,
                                                                       int num_nodes)
{
    __shared__ __align__(16) float tab[ED_COMBOS * D];
    for (int i = threadIdx.x; i < ED_COMBOS * Q; i += GATHER_WARPS * 32) st_f4(tab + 4 * i, ldg_f4(ee_comb + 4 * i));
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * GATHER_WARPS + (threadIdx.x >> 5), nwarps = gridDim.x * GATHER_WARPS;
    if (lane >= Q) return;
    for (int v = warp; v < num_nodes; v += nwarps)
    {
        const int eb = __ldg(in_ptr + v), ee = __ldg(in_ptr + v + 1);
        const float4 hv = ldg_f4(h_in + (size_t)v * D + 4 * lane);
        float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int e = eb; e < ee; e++)
        {
            const int u = __ldg(src + e), c = __ldg(code + e);
            const float4 hu = ldg_f4(h_in + (size_t)u * D + 4 * lane);
            const float4 t = ld_f4(tab + c * D + 4 * lane);
            m.x += relu_f(t.x + hu.x); m.y += relu_f(t.y + hu.y); m.z += relu_f(t.z + hu.z); m.w += relu_f(t.w + hu.w);
        }
        stg_f4_stream(h_out + (size_t)v * D + 4 * lane, make_float4(m.x + hv.x, m.y + hv.y, m.z + hv.z, m.w + hv.w));
    }
}


// ---- staged gather: the message-passing half of a layer for DENSE graphs (hep10k kNN graphs: 16 in-edges per node) ------
// With many in-edges per node every feature row is re-read once per out-edge, and the row-per-warp kernel above becomes
// L2-bandwidth-bound (35 M edges x 400 B = 14 GB per launch on the hep10k workload: 2.04 ms = 6.9 TB/s out of L2).
// The sources of a node are nodes of the SAME graph and a graph's rows are contiguous in HBM, so this kernel reads every
// row from HBM exactly once: a persistent CTA per SM packs consecutive whole graphs into items of up to SG_ROWS rows,
// one bulk-TMA copy per item (cp.async.bulk + mbarrier, double buffered: item i+1 streams in while item i is reduced)
// lands them in shared memory and all in-edge reads hit shared memory.  Warp 0 packs and issues the copies, 31 warps
// take a destination row each: the edge records of a row are read 32 at a time (coalesced) and broadcast by shuffle,
// in-edges are added in CSR order (the reference's order, GIN/src/message_passing.cc:136-145).  Graphs with more than
// SG_ROWS nodes are read from global memory by the same code.
constexpr int SG_THREADS = 1024;
constexpr int SG_ROWS = 240;
struct SgSmem {
    float stage[2][SG_ROWS * D];      // 2 x 96,000 B
    float tab[ED_COMBOS * D];         // 24,000 B
    uint64_t bar[2];
    int item[2][4];                   // first node, end node, staged flag, the one bond code of all edges (or -1)
    int next_row[2];                  // next row of the item to hand out
    alignas(16) int scratch[SG_THREADS];   // per warp: the row offsets of 32 edge records
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// two fp32 additions in one instruction (FADD2, sm_100): the same IEEE round-to-nearest result per element
__device__ __forceinline__ void add2(float& d0, float& d1, float a0, float a1, float b0, float b1)
{
    unsigned long long ua, ub, ud;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(ub) : "f"(b0), "f"(b1));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(ud) : "l"(ua), "l"(ub));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(ud));
}
// signbit(x) ? 0 : x with NaN passing through, one instruction (FMNMX.NAN); the sign of a zero does not reach a sum
__device__ __forceinline__ float relu_max(float x)
{
    float y;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(y) : "f"(x), "f"(0.0f));
    return y;
}
// m += relu(t + hu), four columns
__device__ __forceinline__ void sg_edge(float4& m, const float4& t, const float4& hu)
{
    float4 x;
    add2(x.x, x.y, t.x, t.y, hu.x, hu.y);
    add2(x.z, x.w, t.z, t.w, hu.z, hu.w);
    add2(m.x, m.y, m.x, m.y, relu_max(x.x), relu_max(x.y));
    add2(m.z, m.w, m.z, m.w, relu_max(x.z), relu_max(x.w));
}

// Rows of an item are handed out one at a time through a counter in shared memory (a virtual node's row has three times
// the in-edges of the others: a static split leaves the other warps waiting at the item barrier).  The row loop is a
// two-deep software pipeline in registers: while row v is reduced, the first 32 edge records of the next row and the
// in_ptr pair of the row after next are in flight, so no global-load latency sits between two rows.
// PRE: every edge of the item carries the same bond code c and the stage already holds relu(h_u + EE[c]) (transformed in
// place, once per node instead of once per edge): an edge is one shuffle, one 16-byte shared-memory read and two FADD2.
template <bool STAGED, bool PRE>
__device__ __forceinline__ void sg_rows(const float* __restrict__ rows, const float* tab, const float* __restrict__ h_glob,
                                        float* __restrict__ h_out, const int* __restrict__ in_ptr, const int* __restrict__ src,
                                        const uint8_t* __restrict__ code, int nb, int ne, int* next_row, int* scratch, int lane)
{
    // `rows` is indexed by (node - base): the stage for STAGED (base = nb), h_in otherwise (base = 0)
    const int base = STAGED ? nb : 0;
    const int col = 4 * min(lane, Q - 1);                        // lanes 25..31 shadow lane 24 (same addresses, no store)
    rows += col;
    tab += col;
    auto grab = [&]() {
        int r = 0;
        if (lane == 0) r = atomicAdd(next_row, 1);
        return nb + __shfl_sync(0xFFFFFFFFu, r, 0);
    };
    auto chunk = [&](int e0, int ee, int& us, int& cs) {
        const int idx = e0 + lane;
        us = base; cs = -1;                                       // raw node id: `- base` happens where the record is consumed,
        if (idx < ee) { us = __ldg(src + idx); cs = (int)__ldg(code + idx); }      // not here, where it would wait for the load
    };
    // The edge-embedding row of the current bond code stays in registers: kNN graphs carry ONE code on every edge
    // (hep10k: edge_attr == 0), which halves the shared-memory reads per edge.  Chunks with mixed codes look every edge up.
    int c_cur = -1;
    float4 t_cur = make_float4(0.f, 0.f, 0.f, 0.f);

    int v = grab(), eb = 0, ee = 0, us = 0, cs = -1;
    if (v < ne) { eb = __ldg(in_ptr + v); ee = __ldg(in_ptr + v + 1); }
    int vn = grab(), ebn = 0, een = 0;
    if (vn < ne) { ebn = __ldg(in_ptr + vn); een = __ldg(in_ptr + vn + 1); }
    chunk(eb, ee, us, cs);
    while (v < ne)
    {
        const int v2 = grab();
        int eb2 = 0, ee2 = 0, usn, csn;
        if (v2 < ne) { eb2 = __ldg(in_ptr + v2); ee2 = __ldg(in_ptr + v2 + 1); }
        chunk(ebn, een, usn, csn);
        const float4 hv = (STAGED && !PRE) ? ld_f4(rows + (v - base) * D) : ldg_f4(h_glob + (size_t)v * D + col);
        float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int e0 = eb; e0 < ee; e0 += 32)
        {
            if (e0 != eb) chunk(e0, ee, us, cs);
            const int cnt = min(32, ee - e0);
            us -= base;
            if (PRE)
            {
                // the 32 row offsets go through a per-warp scratch line: one 16-byte broadcast read hands four of them to
                // every lane (a shuffle per edge would cost the shared-memory pipe four times as many wavefronts)
                scratch[lane] = us * D;
                __syncwarp();
                for (int k = 0; k < cnt; k += 4)
                {
                    const int4 o = *reinterpret_cast<const int4*>(scratch + k);
                    const float4 r0 = ld_f4(rows + o.x);
                    const float4 r1 = ld_f4(rows + (k + 1 < cnt ? o.y : o.x));
                    const float4 r2 = ld_f4(rows + (k + 2 < cnt ? o.z : o.x));
                    const float4 r3 = ld_f4(rows + (k + 3 < cnt ? o.w : o.x));
                    add2(m.x, m.y, m.x, m.y, r0.x, r0.y);
                    add2(m.z, m.w, m.z, m.w, r0.z, r0.w);
                    if (k + 1 < cnt) { add2(m.x, m.y, m.x, m.y, r1.x, r1.y); add2(m.z, m.w, m.z, m.w, r1.z, r1.w); }
                    if (k + 2 < cnt) { add2(m.x, m.y, m.x, m.y, r2.x, r2.y); add2(m.z, m.w, m.z, m.w, r2.z, r2.w); }
                    if (k + 3 < cnt) { add2(m.x, m.y, m.x, m.y, r3.x, r3.y); add2(m.z, m.w, m.z, m.w, r3.z, r3.w); }
                }
                __syncwarp();
                continue;
            }
            const int c0 = __shfl_sync(0xFFFFFFFFu, cs, 0);
            if (__all_sync(0xFFFFFFFFu, cs == c0 || cs < 0))
            {
                if (c0 != c_cur) { t_cur = ld_f4(tab + c0 * D); c_cur = c0; }
                if (STAGED) us *= D;                              // element offset of the source row inside the stage
#pragma unroll 4
                for (int k = 0; k < cnt; k++)
                {
                    const int q = __shfl_sync(0xFFFFFFFFu, us, k);
                    sg_edge(m, t_cur, STAGED ? ld_f4(rows + q) : ldg_f4(rows + (size_t)q * D));
                }
            }
            else
            {
#pragma unroll 2
                for (int k = 0; k < cnt; k++)
                {
                    const int q = __shfl_sync(0xFFFFFFFFu, us, k), c = __shfl_sync(0xFFFFFFFFu, cs, k);
                    sg_edge(m, ld_f4(tab + c * D), STAGED ? ld_f4(rows + q * D) : ldg_f4(rows + (size_t)q * D));
                }
            }
        }
        if (lane < Q) stg_f4_stream(h_out + (size_t)v * D + col, make_float4(m.x + hv.x, m.y + hv.y, m.z + hv.z, m.w + hv.w));
        v = vn; eb = ebn; ee = een; us = usn; cs = csn;
        vn = v2; ebn = eb2; een = ee2;
    }
}

__global__ void __launch_bounds__(SG_THREADS, 1) gin_gather_staged_kernel(const float* __restrict__ h_in, float* __restrict__ h_out,
                                                                           const int* __restrict__ in_ptr, const int* __restrict__ src,
                                                                           const uint8_t* __restrict__ code,
                                                                           const float* __restrict__ ee_comb,
                                                                           const int* __restrict__ node_off, int num_graphs)
{
    extern __shared__ __align__(128) unsigned char sg_raw[];
    SgSmem& sm = *reinterpret_cast<SgSmem*>(sg_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g_hi = (int)((long)num_graphs * (blockIdx.x + 1) / gridDim.x);
    int g_next = (int)((long)num_graphs * blockIdx.x / gridDim.x);           // used by thread 0 only

    for (int i = tid; i < ED_COMBOS * Q; i += SG_THREADS) st_f4(sm.tab + 4 * i, ldg_f4(ee_comb + 4 * i));
    if (tid == 0)
    {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        fence_mbar_init();
    }
    // warp 0: pack the next run of whole graphs (<= SG_ROWS rows) into `slot`, start its copy, and -- while the copy is in
    // flight -- check whether all in-edges of the item carry one bond code (record[3]: that code, else -1)
    auto issue = [&](int slot) {
        int nb = 0, ne = 0, staged = 0;
        if (lane == 0)
        {
            while (g_next < g_hi && ne == nb)                      // skip graphs without nodes
            {
                nb = __ldg(node_off + g_next);
                ne = __ldg(node_off + g_next + 1);
                g_next++;
            }
            if (ne > nb && ne - nb <= SG_ROWS)
            {
                staged = 1;
                while (g_next < g_hi)
                {
                    const int nx = __ldg(node_off + g_next + 1);
                    if (nx - nb > SG_ROWS) break;
                    ne = nx;
                    g_next++;
                }
            }
        }
        nb = __shfl_sync(0xFFFFFFFFu, nb, 0); ne = __shfl_sync(0xFFFFFFFFu, ne, 0); staged = __shfl_sync(0xFFFFFFFFu, staged, 0);
        int ucode = -1;
        if (staged)
        {
            if (lane == 0)
            {
                const uint32_t bytes = (uint32_t)(ne - nb) * (D * 4);
                fence_proxy_async();
                mbar_arrive_expect_tx(&sm.bar[slot], bytes);
                tma_load_1d(sm.stage[slot], h_in + (size_t)nb * D, bytes, &sm.bar[slot]);
            }
            const int eb = __ldg(in_ptr + nb), ee = __ldg(in_ptr + ne);
            if (ee > eb)
            {
                // 16 bytes per load; the bytes outside [eb, ee) of the first and last vector are masked out
                const int c0 = (int)__ldg(code + eb);
                const uint32_t pat = (uint32_t)c0 * 0x01010101u;
                const long lo = (long)eb, hi = (long)ee;
                uint32_t diff = 0;
#pragma unroll 2
                for (long p = (lo & ~15L) + 16 * lane; p < hi; p += 512)
                {
                    const uint4 w4 = __ldg(reinterpret_cast<const uint4*>(code + p));
                    const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                    for (int k = 0; k < 4; k++)
                    {
                        const long pw = p + 4 * k;
                        const int b0 = (int)max(lo - pw, 0L), b1 = (int)min(hi - pw, 4L);      // valid bytes [b0, b1) of this word
                        if (b1 > b0) diff |= (w[k] ^ pat) & (0xFFFFFFFFu >> (8 * (4 - b1))) & (0xFFFFFFFFu << (8 * b0));
                    }
                }
                if (__all_sync(0xFFFFFFFFu, diff == 0)) ucode = c0;
            }
        }
        else if (lane == 0) mbar_arrive(&sm.bar[slot]);            // keeps the phase count in step with the item count
        if (lane == 0)
        {
            sm.item[slot][0] = nb; sm.item[slot][1] = ne; sm.item[slot][2] = staged; sm.item[slot][3] = ucode;
            sm.next_row[slot] = 0;
        }
    };
    if (warp == 0) issue(0);
    for (int it = 0;; it++)
    {
        const int s = it & 1;
        __syncthreads();                                          // item it-1 is done: its slot may be refilled; item it's record is visible
        const int nb = sm.item[s][0], ne = sm.item[s][1], staged = sm.item[s][2], ucode = sm.item[s][3];
        if (ne == nb) break;                                      // CTA-uniform: no graphs left
        if (warp == 0) { issue(s ^ 1); continue; }
        if (!staged) { sg_rows<false, false>(h_in, sm.tab, h_in, h_out, in_ptr, src, code, nb, ne, &sm.next_row[s], sm.scratch + 32 * warp, lane); continue; }
        mbar_wait(&sm.bar[s], (it >> 1) & 1);
        if (ucode < 0) { sg_rows<true, false>(sm.stage[s], sm.tab, h_in, h_out, in_ptr, src, code, nb, ne, &sm.next_row[s], sm.scratch + 32 * warp, lane); continue; }
        // one code on every edge: stage <- relu(stage + EE[code]) in place, then the edges only add
        if (lane < Q)
        {
            const float4 t = ld_f4(sm.tab + ucode * D + 4 * lane);
            for (int r = warp - 1; r < ne - nb; r += SG_THREADS / 32 - 1)
            {
                float* x = sm.stage[s] + r * D + 4 * lane;
                const float4 h = ld_f4(x);
                st_f4(x, make_float4(relu_max(h.x + t.x), relu_max(h.y + t.y), relu_max(h.z + t.z), relu_max(h.w + t.w)));
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(SG_THREADS - 32) : "memory");          // the 31 row warps
        sg_rows<true, true>(sm.stage[s], sm.tab, h_in, h_out, in_ptr, src, code, nb, ne, &sm.next_row[s], sm.scratch + 32 * warp, lane);
        fence_proxy_async();                                      // these generic writes precede the next bulk copy into this stage
    }
}

// Row descriptors "no in-edges" (prep.cu::empty_row_desc): the CTA-pair kernel run on them is the node MLP alone.
__global__ void fill_empty_desc_kernel(int4* d, long n)
{
    const int e = 32768 | (ED_COMBOS << 16);
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) d[i] = make_int4(e, e, e, e);
}

}  // namespace

int gin_forward(DeviceBatch& b, const GinWeights& w, const RunOptions& opt, int sm_count, cudaStream_t s, int* launches)
{
    const long N = b.total_nodes;
    if (b.num_graphs == 0) return 0;
    if (!b.has_attr) { set_last_error("GIN needs edge_attr"); return FG_ERR_INVALID; }
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

    const int num_tiles = (int)ceil_div<long>(N, TILE_M);
    FG_TRY(opt_in_smem(reinterpret_cast<const void*>(&gin_layer_kernel<false>), GinSmem<false>::BYTES));
    FG_TRY(opt_in_smem(reinterpret_cast<const void*>(&gin_layer_kernel<true>), GinSmem<true>::BYTES));
    // Dense graphs (average in-degree >= 6: hep10k kNN graphs, GIN-VN on them) run every layer as TWO launches: the staged
    // gather (each row read from HBM once, in-edge reads from shared memory) writes x = m + h, and the CTA-pair kernel run
    // on "no in-edges" descriptors applies the node MLP.  Sparse graphs (molecules) keep the single fused launch.
    const bool staged = opt.gin_staged < 0 ? (b.total_edges >= 6 * N) : (opt.gin_staged != 0);
    const bool split_layer = staged && !opt.mp_only && !opt.gin_ffma;
    auto staged_gather = [&](const float* x_in, float* x_out, int l) -> int {
        FG_TRY(opt_in_smem(reinterpret_cast<const void*>(&gin_gather_staged_kernel), (int)sizeof(SgSmem)));
        const int grid = std::max(1, std::min(b.num_graphs, sm_count));
        gin_gather_staged_kernel<<<grid, SG_THREADS, sizeof(SgSmem), s>>>(x_in, x_out, b.in_ptr.as<int>(), b.src.as<int>(), b.code.as<uint8_t>(),
                                                                         w.ee_comb.as<float>() + (size_t)l * ED_COMBOS * D,
                                                                         b.node_off.as<int>(), b.num_graphs);
        FG_CUDA(cudaGetLastError());
        return 0;
    };
    if (split_layer && opt.gin_tc2)
    {
        FG_TRY(b.row_desc0.reserve(sizeof(int4) * (size_t)(N + 1)));
        fill_empty_desc_kernel<<<sm_count * 4, 256, 0, s>>>(b.row_desc0.as<int4>(), N + 1);
        FG_CUDA(cudaGetLastError());
        nl++;
    }
    // the layer kernel: gin_fused.cu (TMA-staged graph-aligned tiles, shared-memory gather), or with option gin_tc2 the
    // round-1 CTA-pair kernel that gathers through L1 from global memory
    // mlp_only: every row without in-edges = the node MLP alone (second launch of a dense-graph layer)
    auto pair_layer = [&](int l, const float* x_in, float* x_out, const float* head_w, float* node_dot, bool mlp_only) -> int {
        if (opt.gin_tc2) return gin_layer_tc2_launch(b, w, l, x_in, x_out, sm_count, s, head_w, node_dot, mlp_only ? b.row_desc0.as<int4>() : nullptr);
        return gin_layer_fused_launch(b, w, l, x_in, x_out, sm_count, s, head_w, node_dot, mlp_only, 0);
    };
    bool fused_head = false;
    for (int l = 0; l < 5; l++)
    {
        if (opt.timer && (l == 0 || !opt.timer_group)) FG_TRY(opt.timer->mark(s));
        if (split_layer)
        {
            // h[0] -> gather -> h[1] -> node MLP -> h[0]
            FG_TRY(staged_gather(h[0], h[1], l));
            nl++;
            if (l == 4 && !opt.gin_unfused_head)
            {
                FG_TRY(b.node_dot.reserve(sizeof(float) * (size_t)(N + 1)));
                FG_TRY(pair_layer(l, h[1], h[0], w.pred_w.as<float>(), b.node_dot.as<float>(), true));
                fused_head = true;
            }
            else FG_TRY(pair_layer(l, h[1], h[0], nullptr, nullptr, true));
            nl++;
            continue;
        }
        GinLayerParams p;
        p.h_in = h[l & 1]; p.h_out = h[(l + 1) & 1];
        p.in_ptr = b.in_ptr.as<int>(); p.src = b.src.as<int>(); p.code = b.code.as<uint8_t>();
        p.ee_comb = w.ee_comb.as<float>() + (size_t)l * ED_COMBOS * D;
        p.w1t = w.w1t.as<float>() + (size_t)l * D * HP; p.b1 = w.b1.as<float>() + (size_t)l * HP;
        p.w2t = w.w2t.as<float>() + (size_t)l * H * DP; p.b2 = w.b2.as<float>() + (size_t)l * DP;
        p.num_nodes = (int)N; p.num_tiles = num_tiles; p.relu_out = (l != 4);
        if (!opt.mp_only && !opt.gin_ffma)
        {
            if (l == 4 && !opt.gin_unfused_head)
            {
                // last layer: the epilogue applies the prediction weights per node, only 4 bytes per node leave the kernel
                FG_TRY(b.node_dot.reserve(sizeof(float) * (size_t)(N + 1)));
                FG_TRY(pair_layer(l, p.h_in, p.h_out, w.pred_w.as<float>(), b.node_dot.as<float>(), false));
                fused_head = true;
            }
            else FG_TRY(pair_layer(l, p.h_in, p.h_out, nullptr, nullptr, false));
            nl++;
            continue;
        }
        if (opt.mp_only == 1 && opt.gin_staged <= 0) FG_TRY(gin_layer_fused_launch(b, w, l, p.h_in, p.h_out, sm_count, s, nullptr, nullptr, false, 1));
        else if (opt.mp_only && staged) FG_TRY(staged_gather(p.h_in, p.h_out, l));
        else if (opt.mp_only)
        {
            const int grid = (int)std::min<long>(ceil_div<long>(N, GATHER_WARPS), (long)sm_count * 8);
            gin_gather_kernel<<<grid, GATHER_WARPS * 32, 0, s>>>(p.h_in, p.h_out, p.in_ptr, p.src, p.code, p.ee_comb, (int)N);
        }
        else
        {
            const int grid = min(num_tiles, sm_count);
            gin_layer_kernel<false><<<grid, NT, GinSmem<false>::BYTES, s>>>(p);
        }
        FG_CUDA(cudaGetLastError());
        nl++;
    }

    if (opt.timer) FG_TRY(opt.timer->mark(s));
    if (fused_head)
    {
        FG_TRY(gin_pool_dot_launch(b.node_dot.as<float>(), b, w.pred_b.as<float>(), s));
        nl++;
        if (launches) *launches += nl;
        return 0;
    }
    HeadParams hp{};
    hp.x = split_layer ? h[0] : h[1]; hp.dim = D; hp.node_off = b.node_off.as<int>(); hp.nn = b.nums_of_nodes.as<int>(); hp.num_graphs = b.num_graphs;
    hp.w[0] = w.pred_w.as<float>(); hp.b[0] = w.pred_b.as<float>(); hp.dims[0] = D; hp.dims[1] = 1; hp.num_layers = 1;
    hp.out = b.out.as<float>();
    FG_TRY(launch_pool_head(hp, s));
    nl++;
    if (launches) *launches += nl;
    return 0;
}

}  // namespace fg

// GCN forward with the Linear_l node transform on the B200 tensor cores (option "gcn_tc"; gcn.cu is the FFMA version and
// documents the reference pipeline, GCN/src/GCN_compute.cc:50-102, node_embedding.cc:98-146, message_passing.cc).
//
// Per step l = 0..4 two launches (tcgemm.cuh):
//   gcn_aggregate_kernel<FIRST>    a = x0 (input embedding)                                            } bf16 hi/lo A blocks,
//   gcn_aggregate_kernel<MIDDLE>   message passing over p_{l-1}, self term, BatchNorm_{l-1}, relu -> a } K = 100 padded to 128
//   tcg::gemm_kernel<2, 112>       p_l = W_l a + b_l   (hi*hi + lo*hi + hi*lo on tcgen05, epilogue adds the bias)
// and the last step (message passing over p_4, BatchNorm_4, no relu, no Linear) stays gcn_layer_kernel<FINAL> (gcn.cu).
// The expressions of the aggregate kernel are those of gcn.cu, in the same order.
#include "internal.cuh"
#include "layers.cuh"
#include "tcgemm.cuh"
#include "fused_tc.cuh"

#include <algorithm>

namespace fg {

namespace {

constexpr int D = 100;
constexpr int NCHUNK = 2;                    // K = 128
constexpr int NPAD = 112;
constexpr int G8 = 13;                       // column groups of eight that hold real columns (the last one: 96..99)
constexpr int AG_WARPS = 8;
constexpr int KC_COLS = 64;

struct GcnAggParams {
    const float* p_in;
    const int* feat; const float* ne_table;                       // FIRST
    const int* in_ptr; const int* src; const uint8_t* code; const float* norm; const int* out_deg;
    const float* ee_comb;                                         // [60][100] of the layer being finished
    const float* root; const float* bn_mean; const float* bn_sqrt_var; const float* bn_weight; const float* bn_bias;
    unsigned char* apack;
    long num_nodes;
};

template <bool FIRST>
__global__ void __launch_bounds__(AG_WARPS * 32) gcn_aggregate_kernel(GcnAggParams p)
{
    const int lane = threadIdx.x & 31;
    const long warp = blockIdx.x * (long)AG_WARPS + (threadIdx.x >> 5), nwarps = (long)gridDim.x * AG_WARPS;
    const long items = ((p.num_nodes + 31) / 32) * G8;
    const EmbedOffsets eo = concat_table_offsets();
    for (long item = warp; item < items; item += nwarps)
    {
        const long rb = item / G8;
        const int g8 = (int)(item - rb * G8);
        const long v = rb * 32 + lane;
        if (v >= p.num_nodes) continue;
        const int c0 = 8 * g8;
        const bool two = c0 + 4 < D;                              // the last group holds columns 96..99 only
        float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (FIRST)
        {
            // layer 0 consumes the input embedding directly (GCN/src/node_embedding.cc:124-127)
            const float4 x0 = embed_chunk<D>(p.feat + (size_t)v * ND_FEATURE, p.ne_table, eo, 2 * g8);
            a[0] = x0.x; a[1] = x0.y; a[2] = x0.z; a[3] = x0.w;
            if (two)
            {
                const float4 x1 = embed_chunk<D>(p.feat + (size_t)v * ND_FEATURE, p.ne_table, eo, 2 * g8 + 1);
                a[4] = x1.x; a[5] = x1.y; a[6] = x1.z; a[7] = x1.w;
            }
        }
        else
        {
            const int eb = __ldg(p.in_ptr + v), ee = __ldg(p.in_ptr + v + 1);
            const float degp1 = (float)(__ldg(p.out_deg + v) + 1);
#pragma unroll
            for (int half = 0; half < 2; half++)
            {
                if (half == 1 && !two) break;
                const int c = c0 + 4 * half;
                float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int e = eb; e < ee; e++)
                {
                    const int u = __ldg(p.src + e), cd = __ldg(p.code + e);
                    const float nrm = __ldg(p.norm + e);
                    const float4 pu = ldg_f4(p.p_in + (size_t)u * D + c);
                    const float4 t = ldg_f4(p.ee_comb + cd * D + c);
                    m.x += nrm * relu_f(t.x + pu.x); m.y += nrm * relu_f(t.y + pu.y);
                    m.z += nrm * relu_f(t.z + pu.z); m.w += nrm * relu_f(t.w + pu.w);
                }
                // finish the layer: self term, BatchNorm (inference), relu
                const float4 pv = ldg_f4(p.p_in + (size_t)v * D + c);
                const float4 rt = ldg_f4(p.root + c), mu = ldg_f4(p.bn_mean + c), sv = ldg_f4(p.bn_sqrt_var + c);
                const float4 ga = ldg_f4(p.bn_weight + c), be = ldg_f4(p.bn_bias + c);
                float4 r;
                r.x = (m.x + relu_f(pv.x + rt.x) / degp1 - mu.x) / sv.x * ga.x + be.x;
                r.y = (m.y + relu_f(pv.y + rt.y) / degp1 - mu.y) / sv.y * ga.y + be.y;
                r.z = (m.z + relu_f(pv.z + rt.z) / degp1 - mu.z) / sv.z * ga.z + be.z;
                r.w = (m.w + relu_f(pv.w + rt.w) / degp1 - mu.w) / sv.w * ga.w + be.w;
                a[4 * half] = relu_f(r.x); a[4 * half + 1] = relu_f(r.y); a[4 * half + 2] = relu_f(r.z); a[4 * half + 3] = relu_f(r.w);
            }
        }
        tcg::put8<NCHUNK>(p.apack, v, c0, a);
        if (g8 == G8 - 1)
            for (int k0 = 8 * G8; k0 < NCHUNK * tcg::KC; k0 += 8) tcg::put8_zero<NCHUNK>(p.apack, v, k0);      // K padding
    }
}

// p_l[v][d] = acc[d] + b_l[d]
struct GcnEpi {
    const float* b; float* p_out;
    struct State {};
    __device__ State begin(int, bool) const { return State{}; }
    __device__ void store(const State&, int v, int d0, const uint32_t (&acc)[16]) const
    {
#pragma unroll
        for (int j = 0; j < 16; j += 4)
            if (d0 + j < D)
            {
                const float4 bb = ldg_f4(b + d0 + j);
                stg_f4_stream(p_out + (size_t)v * D + d0 + j, make_float4(__uint_as_float(acc[j]) + bb.x, __uint_as_float(acc[j + 1]) + bb.y,
                                                                         __uint_as_float(acc[j + 2]) + bb.z, __uint_as_float(acc[j + 3]) + bb.w));
            }
    }
};

// ---- option gcn_fused (default): the same step as ONE launch, the aggregation as the A producer inside the GEMM kernel (fused_tc.cuh) ----
// message passing over p_{l-1} + self term + BatchNorm_{l-1} (+ relu) of columns c .. c+3: the expressions of gcn_aggregate_kernel
// above, in the same order (in-edges in CSR order)
struct GcnRowMath {
    const float* p_in;
    const int* in_ptr; const int* src; const uint8_t* code; const float* norm; const int* out_deg; const int4* row_desc;
    const float* ee_comb; const float* root; const float* bn_mean; const float* bn_sqrt_var; const float* bn_weight; const float* bn_bias;

    // Four rows per lane.  Round trip 1: the four rows' descriptors (first four in-edges packed by prep.cu: relative source | bond
    // code, in-degree), CSR positions and out-degrees.  Then, pair of rows by pair of rows, ONE more round trip: the own rows, every
    // neighbour row and edge norm at once.  In-edges beyond the fourth (rare in molecules) walk the CSR arrays.
    template <bool RELU>
    __device__ __forceinline__ void finish4(const int (&v)[4], const bool (&live)[4], int c, float4 (&out)[4]) const
    {
        int4 d[4];
        int eb[4], od[4];
#pragma unroll
        for (int p = 0; p < 4; p++)
        {
            d[p] = make_int4(0, 0, 0, 0); eb[p] = od[p] = 0;
            if (live[p]) { d[p] = __ldg(row_desc + v[p]); eb[p] = __ldg(in_ptr + v[p]); od[p] = __ldg(out_deg + v[p]); }
        }
#pragma unroll
        for (int p = 0; p < 4; p++)
        {
            const int dq[4] = {d[p].x, d[p].y, d[p].z, d[p].w};
            const int deg = live[p] ? (int)((unsigned)d[p].x >> 24) : 0;               // capped at 255 by prep.cu
            float4 pu[4], pv = make_float4(0.f, 0.f, 0.f, 0.f);
            float nr[4];
            if (live[p]) pv = ldg_f4(p_in + (size_t)v[p] * D + c);
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (q < deg)
                {
                    pu[q] = ldg_f4(p_in + (size_t)(v[p] + (dq[q] & 0xFFFF) - 32768) * D + c);
                    nr[q] = __ldg(norm + eb[p] + q);
                }
            float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (q < deg)
                {
                    const float4 t = ldg_f4(ee_comb + ((dq[q] >> 16) & 0xFF) * D + c);        // 24 KB table, L1 resident
                    m.x += nr[q] * relu_f(t.x + pu[q].x); m.y += nr[q] * relu_f(t.y + pu[q].y);
                    m.z += nr[q] * relu_f(t.z + pu[q].z); m.w += nr[q] * relu_f(t.w + pu[q].w);
                }
            if (deg > 4)
            {
                const int end = __ldg(in_ptr + v[p] + 1);
                for (int e = eb[p] + 4; e < end; e++)
                {
                    const int u = __ldg(src + e), cd = __ldg(code + e);
                    const float n1 = __ldg(norm + e);
                    const float4 p1 = ldg_f4(p_in + (size_t)u * D + c);
                    const float4 t = ldg_f4(ee_comb + cd * D + c);
                    m.x += n1 * relu_f(t.x + p1.x); m.y += n1 * relu_f(t.y + p1.y); m.z += n1 * relu_f(t.z + p1.z); m.w += n1 * relu_f(t.w + p1.w);
                }
            }
            // self term, BatchNorm (inference), relu.  The two divisions of the reference's expression (by deg + 1 and by
            // sqrt(var + eps)) are multiplications by reciprocals here: <= 2 ulp from the IEEE quotient, against a 1e-4 bar
            const float4 rt = ldg_f4(root + c), mu = ldg_f4(bn_mean + c), sv = ldg_f4(bn_sqrt_var + c);
            const float4 ga = ldg_f4(bn_weight + c), be = ldg_f4(bn_bias + c);
            const float4 g = make_float4(__fdividef(ga.x, sv.x), __fdividef(ga.y, sv.y), __fdividef(ga.z, sv.z), __fdividef(ga.w, sv.w));
            const float rdeg = __fdividef(1.0f, (float)(od[p] + 1));
            float4 r;
            r.x = (m.x + relu_f(pv.x + rt.x) * rdeg - mu.x) * g.x + be.x;
            r.y = (m.y + relu_f(pv.y + rt.y) * rdeg - mu.y) * g.y + be.y;
            r.z = (m.z + relu_f(pv.z + rt.z) * rdeg - mu.z) * g.z + be.z;
            r.w = (m.w + relu_f(pv.w + rt.w) * rdeg - mu.w) * g.w + be.w;
            if (RELU) { r.x = relu_f(r.x); r.y = relu_f(r.y); r.z = relu_f(r.z); r.w = relu_f(r.w); }
            out[p] = live[p] ? r : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }

    // Eight lanes per row (fused_tc.cuh, LPR = 8): two rows per lane, eight columns c .. c+7 per row (n8 = how many of them are real:
    // 8, or 4 for the row's last piece 96..99).  The rows' CSR walks advance together; the (source, code, norm) of a row's NEXT in-edge
    // are requested together with its CURRENT neighbour row.  The per-edge work that does not depend on the column -- indices,
    // predicates, addresses -- is amortised over eight columns instead of four.
    template <bool RELU>
    __device__ __forceinline__ void finish2x8(const int (&v)[2], const bool (&live)[2], int c, bool two, float4 (&out)[2][2]) const
    {
        int e[2], end[2], u[2], cd[2];
        float nr[2];
        float4 m[2][2], pv[2][2];
        float degp1[2];
#pragma unroll
        for (int q = 0; q < 2; q++)
        {
            e[q] = live[q] ? __ldg(in_ptr + v[q]) : 0;
            end[q] = live[q] ? __ldg(in_ptr + v[q] + 1) : 0;
            degp1[q] = live[q] ? (float)(__ldg(out_deg + v[q]) + 1) : 1.0f;
            m[q][0] = m[q][1] = pv[q][0] = pv[q][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live[q]) { pv[q][0] = ldg_f4(p_in + (size_t)v[q] * D + c); if (two) pv[q][1] = ldg_f4(p_in + (size_t)v[q] * D + c + 4); }
        }
#pragma unroll
        for (int q = 0; q < 2; q++)
        {
            u[q] = 0; cd[q] = 0; nr[q] = 0.f;
            if (e[q] < end[q]) { u[q] = __ldg(src + e[q]); cd[q] = __ldg(code + e[q]); nr[q] = __ldg(norm + e[q]); }
        }
        while ((e[0] < end[0]) | (e[1] < end[1]))
        {
            float4 pu[2][2];
            int un[2], cn[2];
            float nn[2];
#pragma unroll
            for (int q = 0; q < 2; q++)
            {
                un[q] = 0; cn[q] = 0; nn[q] = 0.f;
                if (e[q] < end[q])
                {
                    pu[q][0] = ldg_f4(p_in + (size_t)u[q] * D + c);
                    if (two) pu[q][1] = ldg_f4(p_in + (size_t)u[q] * D + c + 4);
                    if (e[q] + 1 < end[q]) { un[q] = __ldg(src + e[q] + 1); cn[q] = __ldg(code + e[q] + 1); nn[q] = __ldg(norm + e[q] + 1); }
                }
            }
#pragma unroll
            for (int q = 0; q < 2; q++)
                if (e[q] < end[q])
                {
                    const float4 t0 = ldg_f4(ee_comb + cd[q] * D + c);           // 24 KB table, L1 resident
                    m[q][0].x += nr[q] * relu_f(t0.x + pu[q][0].x); m[q][0].y += nr[q] * relu_f(t0.y + pu[q][0].y);
                    m[q][0].z += nr[q] * relu_f(t0.z + pu[q][0].z); m[q][0].w += nr[q] * relu_f(t0.w + pu[q][0].w);
                    if (two)
                    {
                        const float4 t1 = ldg_f4(ee_comb + cd[q] * D + c + 4);
                        m[q][1].x += nr[q] * relu_f(t1.x + pu[q][1].x); m[q][1].y += nr[q] * relu_f(t1.y + pu[q][1].y);
                        m[q][1].z += nr[q] * relu_f(t1.z + pu[q][1].z); m[q][1].w += nr[q] * relu_f(t1.w + pu[q][1].w);
                    }
                    e[q]++;
                    u[q] = un[q]; cd[q] = cn[q]; nr[q] = nn[q];
                }
        }
        // self term, BatchNorm (inference), relu.  The two divisions of the reference's expression (by deg + 1 and by sqrt(var + eps))
        // are multiplications by reciprocals here: <= 2 ulp from the IEEE quotient, against a 1e-4 bar
        const float rdeg[2] = {__fdividef(1.0f, degp1[0]), __fdividef(1.0f, degp1[1])};
#pragma unroll
        for (int i = 0; i < 2; i++)
        {
            if (i == 1 && !two) { out[0][1] = out[1][1] = make_float4(0.f, 0.f, 0.f, 0.f); break; }
            const int ci = c + 4 * i;
            const float4 rt = ldg_f4(root + ci), mu = ldg_f4(bn_mean + ci), sv = ldg_f4(bn_sqrt_var + ci);
            const float4 ga = ldg_f4(bn_weight + ci), be = ldg_f4(bn_bias + ci);
            const float4 g = make_float4(__fdividef(ga.x, sv.x), __fdividef(ga.y, sv.y), __fdividef(ga.z, sv.z), __fdividef(ga.w, sv.w));
#pragma unroll
            for (int q = 0; q < 2; q++)
            {
                float4 r;
                r.x = (m[q][i].x + relu_f(pv[q][i].x + rt.x) * rdeg[q] - mu.x) * g.x + be.x;
                r.y = (m[q][i].y + relu_f(pv[q][i].y + rt.y) * rdeg[q] - mu.y) * g.y + be.y;
                r.z = (m[q][i].z + relu_f(pv[q][i].z + rt.z) * rdeg[q] - mu.z) * g.z + be.z;
                r.w = (m[q][i].w + relu_f(pv[q][i].w + rt.w) * rdeg[q] - mu.w) * g.w + be.w;
                if (RELU) { r.x = relu_f(r.x); r.y = relu_f(r.y); r.z = relu_f(r.z); r.w = relu_f(r.w); }
                out[q][i] = live[q] ? r : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
};

template <bool FIRST>
struct GcnFused {
    static constexpr int NCHUNK = fg::NCHUNK, NPAD = fg::NPAD;
    static constexpr unsigned ksteps(int c) { return c == 0 ? 0xFu : 0x7u; }      // K = 112: columns 64..111 are three steps of chunk 1
    GcnRowMath r;
    const int* feat; const float* ne_table;
    const float* b; float* p_out;

    static constexpr int LPR = FIRST ? 16 : 8;          // the embedding step is nine table lookups per piece: more lanes, fewer pieces per lane
    struct Rows {};
    __device__ __forceinline__ Rows rows_begin(const int (&)[2], const bool (&)[2]) const { return Rows{}; }
    __device__ __forceinline__ bool gather2(const Rows&, const int (&v)[2], const bool (&live)[2], int c, int j, float4 (&x)[2][2]) const
    {
        const int col = KC_COLS * c + 8 * j;
        if (col >= NPAD) return false;
#pragma unroll
        for (int q = 0; q < 2; q++) x[q][0] = x[q][1] = make_float4(0.f, 0.f, 0.f, 0.f);      // K padding and rows past the batch hold zeros
        if (col >= D) return true;
        const bool two = col + 4 < D;                                                        // the row's last piece holds columns 96..99 only
        if (FIRST)
        {
#pragma unroll
            for (int q = 0; q < 2; q++)                                                      // GCN/src/node_embedding.cc:124-127
                if (live[q])
                {
                    x[q][0] = embed_chunk<D>(feat + (size_t)v[q] * ND_FEATURE, ne_table, concat_table_offsets(), col / 4);
                    if (two) x[q][1] = embed_chunk<D>(feat + (size_t)v[q] * ND_FEATURE, ne_table, concat_table_offsets(), col / 4 + 1);
                }
        }
        else r.template finish2x8<true>(v, live, col, two, x);
        return true;
    }
    __device__ __forceinline__ Rows rows_begin(const int (&)[4], const bool (&)[4]) const { return Rows{}; }
    __device__ __forceinline__ bool gather4(const Rows&, const int (&v)[4], const bool (&live)[4], int c, int j, float4 (&x)[4]) const
    {
        const int col = KC_COLS * c + 4 * j;
        if (col >= NPAD) return false;
#pragma unroll
        for (int p = 0; p < 4; p++) x[p] = make_float4(0.f, 0.f, 0.f, 0.f);      // K padding and rows past the batch hold zeros
        if (col >= D) return true;
        if (FIRST)
        {
#pragma unroll
            for (int p = 0; p < 4; p++)                                             // GCN/src/node_embedding.cc:124-127
                if (live[p]) x[p] = embed_chunk<D>(feat + (size_t)v[p] * ND_FEATURE, ne_table, concat_table_offsets(), col / 4);
        }
        else r.template finish4<true>(v, live, col, x);
        return true;
    }
    // one thread, two tiles ahead: the tile's rows (graphs are contiguous, so nearly every source row lies in the tile's own
    // row range) and its CSR slice into L2, so that the gather's dependent loads do not wait for HBM
    __device__ __forceinline__ void prefetch_tile(int v0, int rows) const
    {
        if (FIRST) { tcf::prefetch_l2(feat + (size_t)v0 * ND_FEATURE, rows * ND_FEATURE * 4); return; }
        tcf::prefetch_l2(r.p_in + (size_t)v0 * D, rows * D * 4);
        tcf::prefetch_l2(r.in_ptr + v0, rows * 4 + 4);
        tcf::prefetch_l2(r.out_deg + v0, rows * 4);
        tcf::prefetch_l2(r.row_desc + v0, rows * 16);
        const int e0 = __ldg(r.in_ptr + v0), e1 = __ldg(r.in_ptr + v0 + rows);
        tcf::prefetch_l2(r.src + e0, (e1 - e0) * 4);
        tcf::prefetch_l2(r.norm + e0, (e1 - e0) * 4);
        tcf::prefetch_l2(r.code + e0, e1 - e0);
    }
    struct Pre {};
    __device__ __forceinline__ Pre preload(int, bool, int) const { return Pre{}; }
    __device__ __forceinline__ bool row_begin(int, bool live) const { return live; }
    struct RowState {};
    __device__ __forceinline__ RowState row_state() const { return RowState{}; }
    __device__ __forceinline__ void row_end(int, const RowState&) const {}
    __device__ __forceinline__ void store(int v, int d0, const uint32_t (&acc)[16], const Pre&, RowState&) const
    {
#pragma unroll
        for (int j = 0; j < 16; j += 4)
            if (d0 + j < D)
            {
                const float4 bb = ldg_f4(b + d0 + j);
                stg_f4_stream(p_out + (size_t)v * D + d0 + j, make_float4(__uint_as_float(acc[j]) + bb.x, __uint_as_float(acc[j + 1]) + bb.y,
                                                                         __uint_as_float(acc[j + 2]) + bb.z, __uint_as_float(acc[j + 3]) + bb.w));
            }
    }
};

// the last step has no Linear: message passing over p_4, self term, BatchNorm_4, no relu (GCN/src/node_embedding.cc:128-146);
// a warp per four rows, 25 lanes x 4 columns
__global__ void __launch_bounds__(256, 2) gcn_final_kernel(GcnRowMath r, float* __restrict__ h_out, long num_nodes)
{
    const int lane = threadIdx.x & 31;
    const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5, nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long v0 = 4 * warp; v0 < num_nodes; v0 += 4 * nwarps)
        if (lane < D / 4)
        {
            const int v[4] = {(int)v0, (int)v0 + 1, (int)v0 + 2, (int)v0 + 3};
            const bool live[4] = {true, v0 + 1 < num_nodes, v0 + 2 < num_nodes, v0 + 3 < num_nodes};
            float4 x[4];
            r.finish4<false>(v, live, 4 * lane, x);
#pragma unroll
            for (int p = 0; p < 4; p++)
                if (live[p]) stg_f4_stream(h_out + (v0 + p) * D + 4 * lane, x[p]);
        }
}

}  // namespace

size_t gcn_tc_pack_bytes() { return (size_t)NCHUNK * tcg::Cfg<NPAD>::B_BLOCK; }

// W_l [100][100] ("[out][in]") -> two [112 x 64] hi | lo chunks
void gcn_tc_pack_layer(const float* w, unsigned char* dst, uint16_t (*bf16_rn)(float), float (*bf16_to_float)(uint16_t))
{
    tcg::pack_weights<NPAD>(w, D, D, NCHUNK, [](int k) { return k < D ? k : -1; }, dst, bf16_rn, bf16_to_float);
}

int gcn_step_tc_launch(DeviceBatch& b, const GcnWeights& w, int l, const float* p_in, float* p_out, int sm_count, cudaStream_t s)
{
    using C = tcg::Cfg<NPAD>;
    FG_TRY(opt_in_smem(reinterpret_cast<const void*>(&tcg::gemm_kernel<NCHUNK, NPAD, GcnEpi>), C::BYTES));
    const long N = b.total_nodes;
    const int num_tiles = (int)ceil_div<long>(N, tcg::TM);
    FG_TRY(b.apack.reserve((size_t)num_tiles * NCHUNK * tcg::A_BLOCK));
    GcnAggParams p{};
    p.p_in = p_in;
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
    p.apack = b.apack.as<unsigned char>();
    p.num_nodes = N;
    const int blocks = (int)std::min<long>(ceil_div<long>(ceil_div<long>(N, 32) * G8, AG_WARPS), (long)sm_count * 8);
    if (l == 0) gcn_aggregate_kernel<true><<<blocks, AG_WARPS * 32, 0, s>>>(p);
    else gcn_aggregate_kernel<false><<<blocks, AG_WARPS * 32, 0, s>>>(p);
    FG_CUDA(cudaGetLastError());

    tcg::GemmArgs g{};
    g.apack = b.apack.as<unsigned char>();
    g.wpack = w.wpack_tc.as<unsigned char>() + (size_t)l * gcn_tc_pack_bytes();
    g.num_nodes = (int)N; g.num_tiles = num_tiles;
    GcnEpi epi{w.b.as<float>() + (size_t)l * 104, p_out};
    tcg::gemm_kernel<NCHUNK, NPAD, GcnEpi><<<std::min(num_tiles, sm_count), tcg::NT, C::BYTES, s>>>(g, epi);
    FG_CUDA(cudaGetLastError());
    return 0;
}

static GcnRowMath gcn_row_math(const DeviceBatch& b, const GcnWeights& w, int l, const float* p_in)
{
    GcnRowMath r{};
    r.p_in = p_in;
    r.in_ptr = b.in_ptr.as<int>(); r.src = b.src.as<int>(); r.code = b.code.as<uint8_t>(); r.norm = b.edge_w.as<float>(); r.out_deg = b.out_deg.as<int>();
    r.row_desc = b.row_desc.as<int4>();
    if (l > 0)
    {
        const size_t k = (size_t)(l - 1);
        r.ee_comb = w.ee_comb.as<float>() + k * ED_COMBOS * D;
        r.root = w.root.as<float>() + k * D; r.bn_mean = w.bn_mean.as<float>() + k * D; r.bn_sqrt_var = w.bn_sqrt_var.as<float>() + k * D;
        r.bn_weight = w.bn_weight.as<float>() + k * D; r.bn_bias = w.bn_bias.as<float>() + k * D;
    }
    return r;
}

// step l = 0..4 as one launch; l = 5: the final message passing + BatchNorm
int gcn_step_fused_launch(DeviceBatch& b, const GcnWeights& w, int l, const float* p_in, float* p_out, int sm_count, cudaStream_t s)
{
    const long N = b.total_nodes;
    if (N == 0) return 0;                        // a batch of empty graphs: nothing to launch (grid 0 is an invalid configuration)
    if (l == 5)
    {
        gcn_final_kernel<<<(int)std::min<long>(ceil_div<long>(N, 32), (long)sm_count * 8), 256, 0, s>>>(gcn_row_math(b, w, 5, p_in), p_out, N);
        FG_CUDA(cudaGetLastError());
        return 0;
    }
    using C = tcf::Cfg<GcnFused<true>>;
    tcf::Args g{};
    g.wpack = w.wpack_tc.as<unsigned char>() + (size_t)l * gcn_tc_pack_bytes();
    g.num_nodes = (int)N; g.num_tiles = (int)ceil_div<long>(N, tcf::TM);
    const int grid = std::min(g.num_tiles, sm_count);
    if (l == 0)
    {
        FG_TRY(opt_in_smem(reinterpret_cast<const void*>(&tcf::fused_kernel<GcnFused<true>>), C::BYTES));
        GcnFused<true> m{gcn_row_math(b, w, 0, p_in), b.node_feature.as<int>(), w.ne_table.as<float>(), w.b.as<float>(), p_out};
        tcf::fused_kernel<GcnFused<true>><<<grid, tcf::NT, C::BYTES, s>>>(g, m);
    }
    else
    {
        FG_TRY(opt_in_smem(reinterpret_cast<const void*>(&tcf::fused_kernel<GcnFused<false>>), C::BYTES));
        GcnFused<false> m{gcn_row_math(b, w, l, p_in), nullptr, nullptr, w.b.as<float>() + (size_t)l * 104, p_out};
        tcf::fused_kernel<GcnFused<false>><<<grid, tcf::NT, C::BYTES, s>>>(g, m);
    }
    FG_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace fg

// C ABI (include/flowgnn_b200.h): contexts, load_weights, batch upload, and the reference-compatible
// <MODEL>_compute_graphs entry points.
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <utility>
#include <vector>

#include "../../include/flowgnn_b200.h"
#include "host_stage.h"
#include "internal.cuh"
#include "layers.cuh"
#include "tc.cuh"

namespace fg {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }

int opt_in_smem(const void* kernel, int bytes)
{
    static std::mutex mu;
    static std::set<std::pair<const void*, int>> done;
    int dev = 0;
    FG_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (done.count({kernel, dev})) return 0;
    FG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done.insert({kernel, dev});
    return 0;
}

int DevBuf::reserve(size_t bytes)
{
    if (bytes <= cap && ptr) return 0;
    if (bytes == 0) bytes = 16;
    if (ptr) { cudaFree(ptr); ptr = nullptr; cap = 0; }
    bytes = (bytes + 255) & ~size_t(255);
    FG_CUDA(cudaMalloc(&ptr, bytes));
    cap = bytes;
    return 0;
}
void DevBuf::release()
{
    if (ptr) cudaFree(ptr);
    ptr = nullptr; cap = 0;
}

void DeviceBatch::release()
{
    if (h_pack) { cudaFreeHost(h_pack); h_pack = nullptr; h_pack_cap = 0; }
    if (h_pack_done) { cudaEventDestroy(h_pack_done); h_pack_done = nullptr; }
    for (DevBuf* b : {&node_off_perm, &edge_off_perm, &tiles_perm, &node_off_in, &edge_off_in, &node_map, &packed_in}) b->release();
    DevBuf* all[] = {&nums_of_nodes, &nums_of_edges, &node_feature, &edge_list, &edge_attr, &node_eigen, &node_off, &edge_off,
                     &in_ptr, &src, &code, &edge_w, &out_deg, &node_w0, &node_w1, &row_desc, &row_desc_sorted, &row_desc0, &sort_tmp, &big_tab, &status, &tiles, &tile_count, &node_dot, &apack, &nonfinite,
                     &act[0], &act[1], &act[2], &act[3], &score[0], &score[1], &score[2], &score[3], &out};
    for (DevBuf* b : all) b->release();
}

int LayerTimer::mark(cudaStream_t s)
{
    if (marks >= MAX_MARKS) return 0;
    if (marks >= created)
    {
        FG_CUDA(cudaEventCreate(&ev[created]));
        created++;
    }
    FG_CUDA(cudaEventRecord(ev[marks], s));
    marks++;
    return 0;
}

static int upload(DevBuf& dst, const std::vector<float>& v, cudaStream_t s)
{
    FG_TRY(dst.reserve(v.size() * sizeof(float)));
    FG_CUDA(cudaMemcpyAsync(dst.ptr, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice, s));
    FG_CUDA(cudaStreamSynchronize(s));       // `v` is a temporary
    return 0;
}
static int upload(DevBuf& dst, const float* p, size_t n, cudaStream_t s)
{
    return upload(dst, std::vector<float>(p, p + n), s);
}

// edge-embedding rows for the 60 distinct bond-attribute triples, summed exactly as the reference
// does per edge: ((0 + T[a0]) + T[5 + a1]) + T[11 + a2]   (GIN/src/message_passing.cc:136-142)
// (WT_TYPE)float of the reference's host (GIN/src/host_load.cc:60-97): ap_fixed<16,I>, AP_TRN (floor) and AP_WRAP (low 16 bits)
static int16_t to_fixed16(float x, int frac_bits)
{
    const double scaled = std::floor(std::ldexp((double)x, frac_bits));
    if (!(std::fabs(scaled) < 9.0e18)) return 0;                     // inf / NaN: undefined in the reference
    return (int16_t)(uint16_t)(long long)scaled;
}

template <typename T>
static int upload_raw(DevBuf& dst, const std::vector<T>& v, cudaStream_t s)
{
    FG_TRY(dst.reserve(v.size() * sizeof(T)));
    FG_CUDA(cudaMemcpyAsync(dst.ptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    FG_CUDA(cudaStreamSynchronize(s));
    return 0;
}

static std::vector<float> combine_edge_embedding(const float* ee, int layers, int dim)
{
    std::vector<float> out((size_t)layers * ED_COMBOS * dim);
    for (int l = 0; l < layers; l++)
        for (int a0 = 0; a0 < 5; a0++)
            for (int a1 = 0; a1 < 6; a1++)
                for (int a2 = 0; a2 < 2; a2++)
                {
                    const int c = a0 * 12 + a1 * 2 + a2;
                    const float* t = ee + (size_t)l * ED_FEATURE_PER_LAYER * dim;
                    for (int d = 0; d < dim; d++)
                    {
                        float s = 0.0f;
                        s += t[(0 + a0) * dim + d];
                        s += t[(5 + a1) * dim + d];
                        s += t[(11 + a2) * dim + d];
                        out[((size_t)l * ED_COMBOS + c) * dim + d] = s;
                    }
                }
    return out;
}

// combined node-embedding tables of embed4_kernel (layers.cuh): A = T0; B = T1 + T2; C = T3 + T4; E = ((T5 + T6) + T7) + T8
static std::vector<float> combine_node_embedding(const float* t, int dim)
{
    static const int off[9] = {0, 119, 123, 135, 147, 157, 163, 169, 171};
    std::vector<float> out((size_t)431 * dim);
    auto row = [&](int f, int x) { return t + (size_t)(off[f] + x) * dim; };
    for (int d = 0; d < dim; d++)
    {
        for (int x0 = 0; x0 < 119; x0++) out[(size_t)x0 * dim + d] = row(0, x0)[d];
        for (int x1 = 0; x1 < 4; x1++)
            for (int x2 = 0; x2 < 12; x2++) out[(size_t)(119 + x1 * 12 + x2) * dim + d] = row(1, x1)[d] + row(2, x2)[d];
        for (int x3 = 0; x3 < 12; x3++)
            for (int x4 = 0; x4 < 10; x4++) out[(size_t)(167 + x3 * 10 + x4) * dim + d] = row(3, x3)[d] + row(4, x4)[d];
        for (int x5 = 0; x5 < 6; x5++)
            for (int x6 = 0; x6 < 6; x6++)
                for (int x7 = 0; x7 < 2; x7++)
                    for (int x8 = 0; x8 < 2; x8++)
                    {
                        float s = row(5, x5)[d] + row(6, x6)[d];
                        s += row(7, x7)[d];
                        s += row(8, x8)[d];
                        out[(size_t)(287 + ((x5 * 6 + x6) * 2 + x7) * 2 + x8) * dim + d] = s;
                    }
    }
    return out;
}

// [layers][n][k] (reference "[out][in]") -> k-major [layers][k][np] with zero-padded columns
static std::vector<float> transpose_pad(const float* w, int layers, int n, int k, int np)
{
    std::vector<float> out((size_t)layers * k * np, 0.0f);
    for (int l = 0; l < layers; l++)
        for (int o = 0; o < n; o++)
            for (int i = 0; i < k; i++) out[((size_t)l * k + i) * np + o] = w[((size_t)l * n + o) * k + i];
    return out;
}
// fp32 -> bf16, round to nearest even (weights are finite)
static uint16_t bf16_rn(float x)
{
    uint32_t u;
    std::memcpy(&u, &x, 4);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static float bf16_to_float(uint16_t h)
{
    const uint32_t u = (uint32_t)h << 16;
    float x;
    std::memcpy(&x, &u, 4);
    return x;
}

static std::vector<float> pad_rows(const float* b, int layers, int n, int np)
{
    std::vector<float> out((size_t)layers * np, 0.0f);
    for (int l = 0; l < layers; l++)
        for (int o = 0; o < n; o++) out[(size_t)l * np + o] = b[(size_t)l * n + o];
    return out;
}

}  // namespace fg

using namespace fg;

// tile packing of one chunk (pack_graphs below): re-ordered offsets, tile list, caller-order offsets
struct PackOut { std::vector<int32_t> po, pe, pt, pin, pie; };

struct flowgnn_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    DeviceBatch batch;
    bool batch_ready = false;
    // host-pointer entry points: the batch is cut into chunks that alternate between two device batches, so that
    // the H2D copy of chunk i+1 (copy_stream) overlaps the kernels of chunk i (stream)
    static constexpr int PIPE = 2;      // measured: a third buffer lets the uploads run ahead but slows the kernels more than it saves
    DeviceBatch pipe[PIPE];
    cudaStream_t copy_stream = nullptr;
    cudaStream_t aux_stream = nullptr;                 // the input embedding runs here, next to the CSR / tile build (RunOptions::aux)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t up_done[PIPE] = {}, buf_free[PIPE] = {};
    cudaStream_t d2h_stream = nullptr;                 // predictions of chunk i leave on their own stream: behind a long upload the small
    cudaEvent_t fwd_done[PIPE] = {};                   // download took 0.25-0.35 ms to be served, and chunk i+1's kernels were queued behind it
    float* h_out = nullptr; size_t h_out_cap = 0;      // pinned staging of the predictions
    // narrowed uploads (host_stage.h): thread pool, the pinned block of the call in flight, its narrowing run
    std::unique_ptr<HostPool> pool;
    uint8_t* stage_h = nullptr; size_t stage_cap = 0;
    cudaEvent_t stage_done = nullptr;                  // behind the last copy that read stage_h
    NarrowRun narrow;
    PackOut packed[NarrowRun::MAX_CHUNKS];             // computed by the pool next to the narrowing
    int* h_status = nullptr;                           // pinned, one word per chunk
    GinWeights gin; GcnWeights gcn; GatWeights gat; PnaWeights pna; DgnWeights dgn;
    bool loaded[NUM_MODELS] = {false, false, false, false, false};
    uint64_t weight_hash[NUM_MODELS] = {0, 0, 0, 0, 0};
    RunOptions opt;
    LayerTimer timer;
    int time_layers = 0;
    int last_launches = 0;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

const int kNumWeights[NUM_MODELS] = {8, 11, 6, 10, 9};

int load_gin(flowgnn_ctx* c, const float* const* w)
{
    cudaStream_t s = c->stream;
    GinWeights& g = c->gin;
    FG_TRY(upload(g.ne_table, w[0], (size_t)ND_FEATURE_TOTAL * 100, s));
    FG_TRY(upload(g.ne_table4, combine_node_embedding(w[0], 100), s));
    FG_TRY(upload(g.ee_comb, combine_edge_embedding(w[1], 5, 100), s));
    FG_TRY(upload(g.w1t, transpose_pad(w[2], 5, 200, 100, 208), s));
    FG_TRY(upload(g.b1, pad_rows(w[3], 5, 200, 208), s));
    FG_TRY(upload(g.w2t, transpose_pad(w[4], 5, 100, 200, 104), s));
    FG_TRY(upload(g.b2, pad_rows(w[5], 5, 100, 104), s));
    {
        const size_t per_layer = gin_tc2_pack_bytes();
        std::vector<unsigned char> pack(5 * per_layer);
        for (int l = 0; l < 5; l++)
            gin_tc2_pack_layer(w[2] + (size_t)l * 200 * 100, w[3] + (size_t)l * 200, w[4] + (size_t)l * 100 * 200, w[5] + (size_t)l * 100,
                               pack.data() + (size_t)l * per_layer, bf16_rn, bf16_to_float);
        FG_TRY(g.wpack2.reserve(pack.size()));
        FG_CUDA(cudaMemcpyAsync(g.wpack2.ptr, pack.data(), pack.size(), cudaMemcpyHostToDevice, s));
        FG_CUDA(cudaStreamSynchronize(s));
    }
    FG_TRY(upload(g.pred_w, w[6], 100, s));
    FG_TRY(upload(g.pred_b, w[7], 1, s));
    // option "fixed_point": the same weights as ap_fixed<16,6> bit patterns
    {
        constexpr int F = 10;
        std::vector<int16_t> ne((size_t)ND_FEATURE_TOTAL * 100), ee((size_t)5 * ED_COMBOS * 100);
        for (size_t i = 0; i < ne.size(); i++) ne[i] = to_fixed16(w[0][i], F);
        for (int l = 0; l < 5; l++)
            for (int a0 = 0; a0 < 5; a0++)
                for (int a1 = 0; a1 < 6; a1++)
                    for (int a2 = 0; a2 < 2; a2++)
                        for (int d = 0; d < 100; d++)
                        {
                            const float* t = w[1] + (size_t)l * ED_FEATURE_PER_LAYER * 100;
                            const int sum = to_fixed16(t[(0 + a0) * 100 + d], F) + to_fixed16(t[(5 + a1) * 100 + d], F) + to_fixed16(t[(11 + a2) * 100 + d], F);
                            ee[((size_t)l * ED_COMBOS + a0 * 12 + a1 * 2 + a2) * 100 + d] = (int16_t)sum;
                        }
        std::vector<int32_t> w1((size_t)5 * 100 * 200), b1(5 * 200), w2((size_t)5 * 200 * 100), b2(5 * 100), pw(100), pb(1);
        for (int l = 0; l < 5; l++)
        {
            for (int o = 0; o < 200; o++)
                for (int k = 0; k < 100; k++) w1[((size_t)l * 100 + k) * 200 + o] = (int32_t)to_fixed16(w[2][((size_t)l * 200 + o) * 100 + k], F) << (16 - F);
            for (int o = 0; o < 100; o++)
                for (int k = 0; k < 200; k++) w2[((size_t)l * 200 + k) * 100 + o] = (int32_t)to_fixed16(w[4][((size_t)l * 100 + o) * 200 + k], F) << (16 - F);
            for (int o = 0; o < 200; o++) b1[l * 200 + o] = to_fixed16(w[3][l * 200 + o], F);
            for (int o = 0; o < 100; o++) b2[l * 100 + o] = to_fixed16(w[5][l * 100 + o], F);
        }
        for (int k = 0; k < 100; k++) pw[k] = (int32_t)to_fixed16(w[6][k], F) << (16 - F);
        pb[0] = to_fixed16(w[7][0], F);
        FG_TRY(upload_raw(g.fx_ne, ne, s)); FG_TRY(upload_raw(g.fx_ee, ee, s));
        FG_TRY(upload_raw(g.fx_w1, w1, s)); FG_TRY(upload_raw(g.fx_b1, b1, s));
        FG_TRY(upload_raw(g.fx_w2, w2, s)); FG_TRY(upload_raw(g.fx_b2, b2, s));
        FG_TRY(upload_raw(g.fx_pw, pw, s)); FG_TRY(upload_raw(g.fx_pb, pb, s));
    }
    return 0;
}

int load_gcn(flowgnn_ctx* c, const float* const* w)
{
    cudaStream_t s = c->stream;
    GcnWeights& g = c->gcn;
    FG_TRY(upload(g.ne_table, w[0], (size_t)ND_FEATURE_TOTAL * 100, s));
    FG_TRY(upload(g.ee_comb, combine_edge_embedding(w[1], 5, 100), s));
    FG_TRY(upload(g.wt, transpose_pad(w[2], 5, 100, 100, 104), s));
    FG_TRY(upload(g.b, pad_rows(w[3], 5, 100, 104), s));
    {
        std::vector<unsigned char> pack(5 * gcn_tc_pack_bytes());
        for (int l = 0; l < 5; l++) gcn_tc_pack_layer(w[2] + (size_t)l * 100 * 100, pack.data() + (size_t)l * gcn_tc_pack_bytes(), bf16_rn, bf16_to_float);
        FG_TRY(g.wpack_tc.reserve(pack.size()));
        FG_CUDA(cudaMemcpyAsync(g.wpack_tc.ptr, pack.data(), pack.size(), cudaMemcpyHostToDevice, s));
        FG_CUDA(cudaStreamSynchronize(s));
    }
    FG_TRY(upload(g.root, w[4], 500, s));
    FG_TRY(upload(g.bn_weight, w[5], 500, s));
    FG_TRY(upload(g.bn_bias, w[6], 500, s));
    FG_TRY(upload(g.bn_mean, w[7], 500, s));
    std::vector<float> sv(500);
    // bn_sqrt_var = sqrt(var + ap_fixed_epsilon) with epsilon = 2^-10 for <16,6> (GCN/src/load_inputs.cc:32)
    for (int i = 0; i < 500; i++) sv[i] = std::sqrt(w[8][i] + (float)(1.0 / (1 << 10)));
    FG_TRY(upload(g.bn_sqrt_var, sv, s));
    FG_TRY(upload(g.pred_w, w[9], 100, s));
    FG_TRY(upload(g.pred_b, w[10], 1, s));
    return 0;
}

int load_gat(flowgnn_ctx* c, const float* const* w)
{
    cudaStream_t s = c->stream;
    GatWeights& g = c->gat;
    // reference layouts: scoring [l][h][d]; proj/skip [l][ho][do][hi][di]; activations are [v][d][h]
    std::vector<float> a_tgt(5 * 64), a_src(5 * 64), projt(5 * 64 * 64), skipt(5 * 64 * 64), proj0(9 * 64);
    for (int l = 0; l < 5; l++)
        for (int h = 0; h < 4; h++)
            for (int d = 0; d < 16; d++)
            {
                a_tgt[l * 64 + d * 4 + h] = w[0][(l * 4 + h) * 16 + d];
                a_src[l * 64 + d * 4 + h] = w[1][(l * 4 + h) * 16 + d];
            }
    for (int l = 0; l < 5; l++)
        for (int ho = 0; ho < 4; ho++)
            for (int dd = 0; dd < 16; dd++)
                for (int hi = 0; hi < 4; hi++)
                    for (int di = 0; di < 16; di++)
                    {
                        const size_t srci = ((((size_t)l * 4 + ho) * 16 + dd) * 4 + hi) * 16 + di;
                        const size_t dsti = ((size_t)l * 64 + (di * 4 + hi)) * 64 + (dd * 4 + ho);
                        projt[dsti] = w[2][srci];
                        skipt[dsti] = w[3][srci];
                    }
    // layer-0 projection sees the raw features at head_in 0, dim_in f < 9 (GAT/src/load_inputs.cc:203-215)
    for (int f = 0; f < 9; f++)
        for (int ho = 0; ho < 4; ho++)
            for (int dd = 0; dd < 16; dd++) proj0[f * 64 + dd * 4 + ho] = w[2][((((size_t)0 * 4 + ho) * 16 + dd) * 4 + 0) * 16 + f];
    FG_TRY(upload(g.a_tgt, a_tgt, s));
    FG_TRY(upload(g.a_src, a_src, s));
    FG_TRY(upload(g.projt, projt, s));
    FG_TRY(upload(g.skipt, skipt, s));
    FG_TRY(upload(g.proj0, proj0, s));
    {
        std::vector<unsigned char> pack(5 * gat_tc_pack_bytes());
        for (int l = 0; l < 5; l++)
            gat_tc_pack_layer(projt.data() + (size_t)l * 64 * 64, skipt.data() + (size_t)l * 64 * 64, pack.data() + (size_t)l * gat_tc_pack_bytes(), bf16_rn, bf16_to_float);
        FG_TRY(upload_raw(g.wpack_tc, pack, s));
    }
    FG_TRY(upload(g.pred_w, w[4], 16, s));
    FG_TRY(upload(g.pred_b, w[5], 1, s));
    return 0;
}

int load_pna(flowgnn_ctx* c, const float* const* w)
{
    cudaStream_t s = c->stream;
    PnaWeights& g = c->pna;
    FG_TRY(upload(g.ne_table, w[0], (size_t)ND_FEATURE_TOTAL * 80, s));
    FG_TRY(upload(g.ne_table4, combine_node_embedding(w[0], 80), s));
    // reference [l][out][scaler][aggr][in] -> wcat[l][aggr*80 + in][scaler*80 + out]
    std::vector<float> wcat((size_t)4 * 320 * 240);
    for (int l = 0; l < 4; l++)
        for (int o = 0; o < 80; o++)
            for (int sc = 0; sc < 3; sc++)
                for (int a = 0; a < 4; a++)
                    for (int i = 0; i < 80; i++)
                        wcat[((size_t)l * 320 + a * 80 + i) * 240 + sc * 80 + o] = w[1][((((size_t)l * 80 + o) * 3 + sc) * 4 + a) * 80 + i];
    FG_TRY(upload(g.wcat, wcat, s));
    {
        std::vector<unsigned char> pack(4 * pna_tc_pack_bytes());
        for (int l = 0; l < 4; l++) pna_tc_pack_layer(wcat.data() + (size_t)l * 320 * 240, pack.data() + (size_t)l * pna_tc_pack_bytes(), bf16_rn, bf16_to_float);
        FG_TRY(g.wpack_tc.reserve(pack.size()));
        FG_CUDA(cudaMemcpyAsync(g.wpack_tc.ptr, pack.data(), pack.size(), cudaMemcpyHostToDevice, s));
        FG_CUDA(cudaStreamSynchronize(s));
    }
    {
        std::vector<unsigned char> pack(4 * pna_fused_pack_bytes());
        for (int l = 0; l < 4; l++) pna_fused_pack_layer(wcat.data() + (size_t)l * 320 * 240, pack.data() + (size_t)l * pna_fused_pack_bytes(), bf16_rn, bf16_to_float);
        FG_TRY(g.wpack_fused.reserve(pack.size()));
        FG_CUDA(cudaMemcpyAsync(g.wpack_fused.ptr, pack.data(), pack.size(), cudaMemcpyHostToDevice, s));
        FG_CUDA(cudaStreamSynchronize(s));
    }
    FG_TRY(upload(g.w_ref, w[1], (size_t)4 * 80 * 12 * 80, s));
    FG_TRY(upload(g.b, w[2], 320, s));
    FG_TRY(upload(g.m1w, w[3], 40 * 80, s));
    FG_TRY(upload(g.m1b, w[4], 40, s));
    FG_TRY(upload(g.m2w, w[5], 20 * 40, s));
    FG_TRY(upload(g.m2b, w[6], 20, s));
    FG_TRY(upload(g.m3w, w[7], 20, s));
    FG_TRY(upload(g.m3b, w[8], 1, s));
    g.avg_deg = w[9][0];
    return 0;
}

int load_dgn(flowgnn_ctx* c, const float* const* w)
{
    cudaStream_t s = c->stream;
    DgnWeights& g = c->dgn;
    FG_TRY(upload(g.emb, w[0], (size_t)9 * 119 * 100, s));
    FG_TRY(upload(g.wt, transpose_pad(w[1], 4, 100, 200, 104), s));
    FG_TRY(upload(g.w_ref, w[1], (size_t)4 * 100 * 200, s));
    {
        std::vector<unsigned char> pack(4 * dgn_tc_pack_bytes());
        for (int l = 0; l < 4; l++) dgn_tc_pack_layer(w[1] + (size_t)l * 100 * 200, pack.data() + (size_t)l * dgn_tc_pack_bytes(), bf16_rn, bf16_to_float);
        FG_TRY(g.wpack_tc.reserve(pack.size()));
        FG_CUDA(cudaMemcpyAsync(g.wpack_tc.ptr, pack.data(), pack.size(), cudaMemcpyHostToDevice, s));
        FG_CUDA(cudaStreamSynchronize(s));
        for (int l = 0; l < 4; l++) dgn_fused_pack_layer(w[1] + (size_t)l * 100 * 200, pack.data() + (size_t)l * dgn_tc_pack_bytes(), bf16_rn, bf16_to_float);
        FG_TRY(g.wpack_fused.reserve(pack.size()));
        FG_CUDA(cudaMemcpyAsync(g.wpack_fused.ptr, pack.data(), pack.size(), cudaMemcpyHostToDevice, s));
        FG_CUDA(cudaStreamSynchronize(s));
    }
    FG_TRY(upload(g.b, pad_rows(w[2], 4, 100, 104), s));
    FG_TRY(upload(g.m0w, w[3], 50 * 100, s));
    FG_TRY(upload(g.m0b, w[4], 50, s));
    FG_TRY(upload(g.m1w, w[5], 25 * 50, s));
    FG_TRY(upload(g.m1b, w[6], 25, s));
    FG_TRY(upload(g.m2w, w[7], 25, s));
    FG_TRY(upload(g.m2b, w[8], 1, s));
    // option "fixed_point": the same weights as ap_fixed<16,3> bit patterns (DGN/src/host_load.cc casts with (WT_TYPE)float)
    {
        constexpr int F = 13;
        std::vector<int16_t> emb((size_t)9 * 119 * 100);
        for (size_t i = 0; i < emb.size(); i++) emb[i] = to_fixed16(w[0][i], F);
        std::vector<int32_t> fw((size_t)4 * 100 * 200), fb(400), m0w(5000), m0b(50), m1w(1250), m1b(25), m2w(25), m2b(1);
        for (int l = 0; l < 4; l++)
            for (int o = 0; o < 100; o++)
                for (int part = 0; part < 2; part++)
                    for (int k = 0; k < 100; k++)
                        fw[(((size_t)l * 100 + k) * 2 + part) * 100 + o] = (int32_t)to_fixed16(w[1][(((size_t)l * 100 + o) * 2 + part) * 100 + k], F) << (16 - F);
        for (int i = 0; i < 400; i++) fb[i] = to_fixed16(w[2][i], F);
        for (int o = 0; o < 50; o++) for (int k = 0; k < 100; k++) m0w[k * 50 + o] = (int32_t)to_fixed16(w[3][o * 100 + k], F) << (16 - F);
        for (int o = 0; o < 50; o++) m0b[o] = to_fixed16(w[4][o], F);
        for (int o = 0; o < 25; o++) for (int k = 0; k < 50; k++) m1w[k * 25 + o] = (int32_t)to_fixed16(w[5][o * 50 + k], F) << (16 - F);
        for (int o = 0; o < 25; o++) m1b[o] = to_fixed16(w[6][o], F);
        for (int k = 0; k < 25; k++) m2w[k] = (int32_t)to_fixed16(w[7][k], F) << (16 - F);
        m2b[0] = to_fixed16(w[8][0], F);
        FG_TRY(upload_raw(g.fx_emb, emb, s)); FG_TRY(upload_raw(g.fx_w, fw, s)); FG_TRY(upload_raw(g.fx_b, fb, s));
        FG_TRY(upload_raw(g.fx_m0w, m0w, s)); FG_TRY(upload_raw(g.fx_m0b, m0b, s)); FG_TRY(upload_raw(g.fx_m1w, m1w, s));
        FG_TRY(upload_raw(g.fx_m1b, m1b, s)); FG_TRY(upload_raw(g.fx_m2w, m2w, s)); FG_TRY(upload_raw(g.fx_m2b, m2b, s));
    }
    return 0;
}

int check_ctx(flowgnn_ctx* ctx)
{
    if (!ctx) { set_last_error("null context"); return FG_ERR_INVALID; }
    return 0;
}

// bytes the calling thread's last host-pointer entry-point call moved over PCIe (flowgnn_b200_last_transfer_bytes)
thread_local uint64_t g_h2d_bytes = 0, g_d2h_bytes = 0;

int copy_in(DevBuf& dst, const void* src, size_t bytes, cudaStream_t s)
{
    FG_TRY(dst.reserve(bytes));
    if (bytes) FG_CUDA(cudaMemcpyAsync(dst.ptr, src, bytes, cudaMemcpyHostToDevice, s));
    g_h2d_bytes += bytes;
    return 0;
}

// device-side rejections (prep.cu) -> error code + text
int status_error(int st)
{
    if (st & 2) { set_last_error("edge_list holds a node id outside [0, num_of_nodes)"); return FG_ERR_INVALID; }
    if (st & 4) { set_last_error("edge_attr holds a value outside the bond vocabulary {5, 6, 2} (GIN/src/host_load.cc:5-6)"); return FG_ERR_INVALID; }
    if (st & 1) { set_last_error("invalid node / edge count of a graph"); return FG_ERR_LIMIT; }
    if (st & 8) { set_last_error("an edge spans more than 32,767 node positions (GIN, GCN row descriptors; reference cap: MAX_NODE = 500 nodes per graph)"); return FG_ERR_LIMIT; }
    return 0;
}

int check_status(flowgnn_ctx* ctx)
{
    int st = 0;
    FG_CUDA(cudaMemcpyAsync(&st, ctx->batch.status.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(cudaStreamSynchronize(ctx->stream));
    return status_error(st);
}

}  // namespace

namespace fg { extern unsigned long long* gin_tc2_trace_buffer; }

extern "C" {

// debugging aid (not part of include/flowgnn_b200.h): device buffer of 3 x 64 x 8 timestamps filled by pair 0 of the GIN layer kernel
int flowgnn_b200_debug_trace(unsigned long long** dev_buffer, int enable)
{
    static unsigned long long* buf = nullptr;
    if (!buf) FG_CUDA(cudaMalloc(&buf, 4096 * sizeof(unsigned long long)));
    if (enable) FG_CUDA(cudaMemset(buf, 0, 4096 * sizeof(unsigned long long)));
    fg::gin_tc2_trace_buffer = enable ? buf : nullptr;
    if (dev_buffer) *dev_buffer = buf;
    return 0;
}

const char* flowgnn_b200_last_error(void) { return g_last_error.c_str(); }

int flowgnn_b200_create(flowgnn_ctx** out, int device)
{
    if (!out) { set_last_error("null out pointer"); return FG_ERR_INVALID; }
    *out = nullptr;
    int count = 0;
    FG_CUDA(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) { set_last_error("no such CUDA device"); return FG_ERR_INVALID; }
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    FG_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
    {
        set_last_error(std::string("flowgnn_b200 is built for sm_100a only; device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor));
        return FG_ERR_INVALID;
    }
    std::unique_ptr<flowgnn_ctx> c(new flowgnn_ctx);
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    // FLOWGNN_B200_TC_ALL=0/1 overrides the default (1) of the "gcn_tc" / "dgn_tc" options (used to run the whole GPU suite on either path)
    if (const char* e = std::getenv("FLOWGNN_B200_TC_ALL")) c->opt.gcn_tc = c->opt.dgn_tc = std::atoi(e) != 0;
    FG_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    FG_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    FG_CUDA(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
    FG_CUDA(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
    FG_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    FG_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    for (int i = 0; i < flowgnn_ctx::PIPE; i++)
    {
        FG_CUDA(cudaEventCreateWithFlags(&c->up_done[i], cudaEventDisableTiming));
        FG_CUDA(cudaEventCreateWithFlags(&c->buf_free[i], cudaEventDisableTiming));
        FG_CUDA(cudaEventCreateWithFlags(&c->fwd_done[i], cudaEventDisableTiming));
    }
    FG_CUDA(cudaMallocHost(&c->h_status, sizeof(int) * 64));
    FG_CUDA(cudaEventCreate(&c->ev0));
    FG_CUDA(cudaEventCreate(&c->ev1));
    *out = c.release();
    return 0;
}

int flowgnn_b200_destroy(flowgnn_ctx* ctx)
{
    if (!ctx) return 0;
    DeviceGuard guard(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copy_stream);
    ctx->batch.release();
    for (int i = 0; i < flowgnn_ctx::PIPE; i++)
    {
        ctx->pipe[i].release();
        cudaEventDestroy(ctx->up_done[i]);
        cudaEventDestroy(ctx->buf_free[i]);
        cudaEventDestroy(ctx->fwd_done[i]);
    }
    if (ctx->h_out) cudaFreeHost(ctx->h_out);
    if (ctx->h_status) cudaFreeHost(ctx->h_status);
    ctx->pool.reset();
    if (ctx->stage_h) cudaFreeHost(ctx->stage_h);
    if (ctx->stage_done) cudaEventDestroy(ctx->stage_done);
    cudaStreamDestroy(ctx->copy_stream);
    if (ctx->d2h_stream) { cudaStreamSynchronize(ctx->d2h_stream); cudaStreamDestroy(ctx->d2h_stream); }
    if (ctx->aux_stream) { cudaStreamSynchronize(ctx->aux_stream); cudaStreamDestroy(ctx->aux_stream); }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    DevBuf* w[] = {&ctx->gin.ne_table, &ctx->gin.ne_table4, &ctx->pna.ne_table4, &ctx->gin.ee_comb, &ctx->gin.w1t, &ctx->gin.b1, &ctx->gin.w2t, &ctx->gin.b2, &ctx->gin.wpack2, &ctx->gin.pred_w, &ctx->gin.pred_b,
                   &ctx->gcn.ne_table, &ctx->gcn.ee_comb, &ctx->gcn.wpack_tc, &ctx->gcn.wt, &ctx->gcn.b, &ctx->gcn.root, &ctx->gcn.bn_mean, &ctx->gcn.bn_sqrt_var,
                   &ctx->gcn.bn_weight, &ctx->gcn.bn_bias, &ctx->gcn.pred_w, &ctx->gcn.pred_b,
                   &ctx->pna.ne_table, &ctx->pna.wcat, &ctx->pna.wpack_tc, &ctx->pna.wpack_fused, &ctx->pna.w_ref, &ctx->pna.b, &ctx->pna.m1w, &ctx->pna.m1b, &ctx->pna.m2w, &ctx->pna.m2b,
                   &ctx->pna.m3w, &ctx->pna.m3b,
                   &ctx->dgn.emb, &ctx->dgn.wt, &ctx->dgn.wpack_tc, &ctx->dgn.wpack_fused, &ctx->dgn.fx_emb, &ctx->dgn.fx_w, &ctx->dgn.fx_b, &ctx->dgn.fx_m0w, &ctx->dgn.fx_m0b,
                   &ctx->dgn.fx_m1w, &ctx->dgn.fx_m1b, &ctx->dgn.fx_m2w, &ctx->dgn.fx_m2b, &ctx->gin.fx_ne, &ctx->gin.fx_ee, &ctx->gin.fx_w1, &ctx->gin.fx_b1, &ctx->gin.fx_w2, &ctx->gin.fx_b2,
                   &ctx->gin.fx_pw, &ctx->gin.fx_pb, &ctx->dgn.w_ref, &ctx->dgn.b, &ctx->dgn.m0w, &ctx->dgn.m0b, &ctx->dgn.m1w, &ctx->dgn.m1b,
                   &ctx->dgn.m2w, &ctx->dgn.m2b,
                   &ctx->gat.wpack_tc, &ctx->gat.proj0, &ctx->gat.projt, &ctx->gat.skipt, &ctx->gat.a_src, &ctx->gat.a_tgt, &ctx->gat.pred_w, &ctx->gat.pred_b};
    for (DevBuf* b : w) b->release();
    for (int i = 0; i < ctx->timer.created; i++) cudaEventDestroy(ctx->timer.ev[i]);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return 0;
}

int flowgnn_b200_set_option(flowgnn_ctx* ctx, const char* name, int value)
{
    FG_TRY(check_ctx(ctx));
    if (!name) { set_last_error("null option name"); return FG_ERR_INVALID; }
    if (!std::strcmp(name, "mp_only")) ctx->opt.mp_only = value;
    else if (!std::strcmp(name, "gin_ffma")) ctx->opt.gin_ffma = value;
    else if (!std::strcmp(name, "gin_tc2")) ctx->opt.gin_tc2 = value;
    else if (!std::strcmp(name, "gin_staged")) ctx->opt.gin_staged = value;
    else if (!std::strcmp(name, "pna_tc")) ctx->opt.pna_tc = value;
    else if (!std::strcmp(name, "pna_fused")) ctx->opt.pna_fused = value;
    else if (!std::strcmp(name, "gcn_tc")) ctx->opt.gcn_tc = value;
    else if (!std::strcmp(name, "gcn_fused")) ctx->opt.gcn_fused = value;
    else if (!std::strcmp(name, "dgn_fused")) ctx->opt.dgn_fused = value;
    else if (!std::strcmp(name, "dgn_tc")) ctx->opt.dgn_tc = value;
    else if (!std::strcmp(name, "gin_unfused_head")) ctx->opt.gin_unfused_head = value;
    else if (!std::strcmp(name, "gat_node_offset_bug")) ctx->opt.gat_node_offset_bug = value;
    else if (!std::strcmp(name, "gat_tc")) ctx->opt.gat_tc = value;
    else if (!std::strcmp(name, "embed_overlap")) ctx->opt.embed_overlap = value;
    else if (!std::strcmp(name, "pack_graphs")) ctx->opt.pack_graphs = value;
    else if (!std::strcmp(name, "host_stage")) ctx->opt.host_stage = value;
    else if (!std::strcmp(name, "fixed_point")) ctx->opt.fixed_point = value;
    else if (!std::strcmp(name, "time_layers")) ctx->time_layers = value;
    else { set_last_error(std::string("unknown option ") + name); return FG_ERR_INVALID; }
    return 0;
}

int flowgnn_b200_load_weights(flowgnn_ctx* ctx, int model, const float* const* weights, int num_weights)
{
    FG_TRY(check_ctx(ctx));
    if (model < 0 || model >= NUM_MODELS) { set_last_error("unknown model id"); return FG_ERR_INVALID; }
    if (!weights || num_weights != kNumWeights[model])
    {
        set_last_error("load_weights: expected " + std::to_string(kNumWeights[model]) + " weight arrays");
        return FG_ERR_INVALID;
    }
    for (int i = 0; i < num_weights; i++)
        if (!weights[i]) { set_last_error("load_weights: null weight array " + std::to_string(i)); return FG_ERR_INVALID; }
    DeviceGuard guard(ctx->device);
    ctx->loaded[model] = false;
    switch (model)
    {
    case MODEL_GIN: FG_TRY(load_gin(ctx, weights)); break;
    case MODEL_GCN: FG_TRY(load_gcn(ctx, weights)); break;
    case MODEL_GAT: FG_TRY(load_gat(ctx, weights)); break;
    case MODEL_PNA: FG_TRY(load_pna(ctx, weights)); break;
    case MODEL_DGN: FG_TRY(load_dgn(ctx, weights)); break;
    }
    ctx->loaded[model] = true;
    ctx->weight_hash[model] = 0;
    return 0;
}

}  // extern "C"

namespace {

void pack_graphs(const int32_t* nn, const int32_t* ne, int G, std::vector<int32_t>& node_off, std::vector<int32_t>& edge_off, std::vector<int32_t>& tiles,
                 std::vector<int32_t>& node_in, std::vector<int32_t>& edge_in);

// Narrowed upload (host_stage.h): queue the narrowing of ALL chunks of a call on the pool and return; run_reference_entry() then ships
// chunk after chunk as each one completes (NarrowRun::wait_chunk) while the pool is already working on the next ones.
struct StagedChunk {
    const uint8_t* h = nullptr;            // the chunk's narrowed block in pinned memory
    const void* part[3] = {nullptr, nullptr, nullptr};   // or, h == nullptr: the caller's own narrow arrays (flowgnn_b200_upload_batch_packed)
    NarrowPlan plan;
    bool ok[3] = {false, false, false};    // feat / edge_list / edge_attr were narrowed and every value fits
    cudaEvent_t done = nullptr;            // recorded behind the copy of the block
};

int stage_begin_run(flowgnn_ctx* ctx, NarrowRun::Chunk* chunks, int n, std::function<void(int)> first)
{
    const int want_threads = HostPool::default_threads();              // (the environment may change between calls: tests, probes)
    if (!ctx->pool || ctx->pool->threads() != want_threads) ctx->pool.reset(new HostPool(want_threads));
    const size_t bytes = NarrowRun::layout(chunks, n);
    if (!ctx->stage_done) FG_CUDA(cudaEventCreateWithFlags(&ctx->stage_done, cudaEventDisableTiming));
    else FG_CUDA(cudaEventSynchronize(ctx->stage_done));     // the previous call's copies have left the block
    if (bytes > ctx->stage_cap)
    {
        if (ctx->stage_h) cudaFreeHost(ctx->stage_h);
        ctx->stage_h = nullptr; ctx->stage_cap = 0;
        const size_t want = bytes + bytes / 4 + 4096;
        FG_CUDA(cudaMallocHost(&ctx->stage_h, want));
        ctx->stage_cap = want;
    }
    ctx->narrow.start(*ctx->pool, chunks, n, ctx->stage_h, std::move(first));
    return 0;
}

// `staged`: this chunk's narrowed node_feature / edge_list / edge_attr (nullptr: plain copies from the caller's arrays).  `write_after`: event the stream waits for before anything lands in buffers the batch's previous kernels read
// (the narrowed block itself goes into a buffer of its own, so its copy does not wait).
int upload_into(DeviceBatch& b, cudaStream_t s, int num_graphs, int64_t total_nodes, int64_t total_edges, const int32_t* nums_of_nodes,
                const int32_t* nums_of_edges, const int32_t* node_feature, const int32_t* edge_list, const int32_t* edge_attr,
                const float* node_eigen, const StagedChunk* staged = nullptr, cudaEvent_t write_after = nullptr,
                const PackOut* prepacked = nullptr)
{
    if (num_graphs < 0 || total_nodes < 0 || total_edges < 0) { set_last_error("negative size"); return FG_ERR_INVALID; }
    if (total_nodes >= (int64_t(1) << 31) / 100 * 4 || total_edges >= (int64_t(1) << 31) - 64)
    {
        set_last_error("batch too large for 32-bit node/edge positions; split it");
        return FG_ERR_LIMIT;
    }
    const bool own_narrow = staged && !staged->h;        // the caller's arrays are narrow already: there are no int32 ones
    if (num_graphs > 0 && (!nums_of_nodes || !nums_of_edges || (total_nodes && !node_feature && !own_narrow) || (total_edges && !edge_list && !own_narrow)))
    {
        set_last_error("null batch array");
        return FG_ERR_INVALID;
    }
    // the counts drive every offset on the device: check them here, where they are host memory (O(G), microseconds)
    int64_t sum_n = 0, sum_e = 0;
    int max_n = 0;
    for (int g = 0; g < num_graphs; g++)
    {
        const int n = nums_of_nodes[g], e = nums_of_edges[g];
        if (n < 0 || e < 0) { set_last_error("negative entry in nums_of_nodes / nums_of_edges (graph " + std::to_string(g) + ")"); return FG_ERR_INVALID; }
        sum_n += n; sum_e += e;
        max_n = std::max(max_n, n);
    }
    if (sum_n != total_nodes || sum_e != total_edges)
    {
        set_last_error("total_nodes / total_edges do not match the sums of nums_of_nodes / nums_of_edges");
        return FG_ERR_INVALID;
    }
    b.num_graphs = num_graphs; b.total_nodes = total_nodes; b.total_edges = total_edges;
    b.max_graph_nodes = max_n;
    b.has_attr = edge_attr != nullptr; b.has_eigen = node_eigen != nullptr;

    // re-ordered graph offsets + tile list for the models with graph-aligned tiles (used by prep.cu when option pack_graphs is on).
    // They travel from a pinned staging block owned by the batch, so that the copies stay asynchronous (the chunked entry points
    // overlap this upload with the previous chunk's kernels); the graph counts ride along (the caller's arrays may be pageable).
    b.has_perm = false;
    const bool pack = num_graphs > 0 && total_nodes > 0;
    static thread_local PackOut local;
    if (pack && !prepacked) pack_graphs(nums_of_nodes, nums_of_edges, num_graphs, local.po, local.pe, local.pt, local.pin, local.pie);
    const PackOut& pk = prepacked ? *prepacked : local;
    const std::vector<int32_t>&po = pk.po, &pe = pk.pe, &pt = pk.pt, &pin = pk.pin, &pie = pk.pie;
    if (pack) b.tiles_perm_count = (long)pt.size() / 2;

    bool narrow_ok[3] = {false, false, false};
    const size_t n_feat = (size_t)ND_FEATURE * (size_t)total_nodes, n_edge = 2 * (size_t)total_edges, n_attr = edge_attr ? 3 * (size_t)total_edges : 0;
    if (staged)
    {
        const NarrowPlan& plan = staged->plan;
        narrow_ok[0] = staged->ok[0] && plan.n_feat == n_feat;
        narrow_ok[1] = staged->ok[1] && plan.n_edge == n_edge;
        narrow_ok[2] = staged->ok[2] && edge_attr && plan.n_attr == n_attr;
        const size_t blob = plan.off_eig;                   // (node_eigen, if it rides in the block, goes straight to its own buffer below)
        FG_TRY(b.packed_in.reserve(blob));
        if (staged->h)
        {
            if (blob) FG_CUDA(cudaMemcpyAsync(b.packed_in.ptr, staged->h, blob, cudaMemcpyHostToDevice, s));
            g_h2d_bytes += blob;
        }
        else
        {
            const size_t off[3] = {plan.off_feat, plan.off_edge, plan.off_attr}, len[3] = {plan.n_feat, 2 * plan.n_edge, plan.n_attr};
            for (int a = 0; a < 3; a++)
                if (len[a] && staged->part[a])
                {
                    FG_CUDA(cudaMemcpyAsync(b.packed_in.as<uint8_t>() + off[a], staged->part[a], len[a], cudaMemcpyHostToDevice, s));
                    g_h2d_bytes += len[a];
                }
        }
        if (staged->done) FG_CUDA(cudaEventRecord(staged->done, s));
    }
    if (write_after) FG_CUDA(cudaStreamWaitEvent(s, write_after, 0));

    if (pack)
    {
        const size_t G = (size_t)num_graphs;
        const size_t words = po.size() + pe.size() + pt.size() + 1 + pin.size() + pie.size() + 2 * G;
        if (!b.h_pack_done) FG_CUDA(cudaEventCreateWithFlags(&b.h_pack_done, cudaEventDisableTiming));
        else FG_CUDA(cudaEventSynchronize(b.h_pack_done));   // the previous upload of this batch has left the staging block
        if (words > b.h_pack_cap)
        {
            if (b.h_pack) cudaFreeHost(b.h_pack);
            b.h_pack = nullptr; b.h_pack_cap = 0;
            FG_CUDA(cudaMallocHost(&b.h_pack, sizeof(int32_t) * (words + words / 4)));
            b.h_pack_cap = words + words / 4;
        }
        int32_t* h = b.h_pack;
        const std::vector<int32_t>* parts[5] = {&po, &pe, &pt, &pin, &pie};
        DevBuf* dst[5] = {&b.node_off_perm, &b.edge_off_perm, &b.tiles_perm, &b.node_off_in, &b.edge_off_in};
        for (int k = 0; k < 5; k++)
        {
            size_t n = parts[k]->size();
            std::memcpy(h, parts[k]->data(), sizeof(int32_t) * n);
            if (k == 2) h[n++] = (int32_t)b.tiles_perm_count;          // the count rides behind the tile list
            FG_TRY(copy_in(*dst[k], h, sizeof(int32_t) * n, s));
            h += n;
        }
        std::memcpy(h, nums_of_nodes, sizeof(int32_t) * G);
        std::memcpy(h + G, nums_of_edges, sizeof(int32_t) * G);
        FG_TRY(copy_in(b.nums_of_nodes, h, sizeof(int32_t) * G, s));
        FG_TRY(copy_in(b.nums_of_edges, h + G, sizeof(int32_t) * G, s));
        FG_CUDA(cudaEventRecord(b.h_pack_done, s));
        b.has_perm = true;
    }
    else
    {
        FG_TRY(copy_in(b.nums_of_nodes, nums_of_nodes, sizeof(int) * (size_t)num_graphs, s));
        FG_TRY(copy_in(b.nums_of_edges, nums_of_edges, sizeof(int) * (size_t)num_graphs, s));
    }

    // the three big arrays: widened on the device from the narrowed block, or copied as they are
    if (narrow_ok[0]) FG_TRY(b.node_feature.reserve(sizeof(int) * n_feat));
    else FG_TRY(copy_in(b.node_feature, node_feature, sizeof(int) * n_feat, s));
    if (narrow_ok[1]) FG_TRY(b.edge_list.reserve(sizeof(int) * n_edge));
    else FG_TRY(copy_in(b.edge_list, edge_list, sizeof(int) * n_edge, s));
    if (edge_attr)
    {
        if (narrow_ok[2]) FG_TRY(b.edge_attr.reserve(sizeof(int) * n_attr));
        else FG_TRY(copy_in(b.edge_attr, edge_attr, sizeof(int) * n_attr, s));
    }
    if (narrow_ok[0] || narrow_ok[1] || (edge_attr && narrow_ok[2]))
    {
        const NarrowPlan& plan = staged->plan;
        FG_TRY(unpack_inputs_launch(b.packed_in.as<uint8_t>(), plan.off_edge, plan.off_attr, b.node_feature.as<int32_t>(), narrow_ok[0] ? n_feat : 0,
                                    b.edge_list.as<int32_t>(), narrow_ok[1] ? n_edge : 0, b.edge_attr.as<int32_t>(),
                                    edge_attr && narrow_ok[2] ? n_attr : 0, s));
    }
    if (node_eigen)
    {
        const bool via_block = staged && staged->plan.n_eig == 4 * (size_t)total_nodes && total_nodes > 0;
        FG_TRY(copy_in(b.node_eigen, via_block ? static_cast<const void*>(staged->h + staged->plan.off_eig) : node_eigen,
                       sizeof(float) * 4 * (size_t)total_nodes, s));
        if (via_block && staged->done) FG_CUDA(cudaEventRecord(staged->done, s));
    }
    FG_TRY(b.out.reserve(sizeof(float) * (size_t)(num_graphs + 1)));
    return 0;
}

// Tile packing.  The GIN and PNA layer kernels work on tiles of whole graphs, at most 128 rows each (prep.cu).  Closing a tile whenever
// the next graph does not fit fills the tiles of a molecule batch to ~89 %; re-ordering the graphs inside windows of 256 by best-fit
// decreasing fills them to ~98 %, i.e. 9 % fewer tiles for every layer launch.  The order only decides WHERE a graph's rows live in
// HBM (node_off_perm / edge_off_perm, indexed by the caller's graph number): inputs are read and predictions written in caller order.
// O(G) on the host, hidden behind the upload it precedes.
void pack_graphs(const int32_t* nn, const int32_t* ne, int G, std::vector<int32_t>& node_off, std::vector<int32_t>& edge_off, std::vector<int32_t>& tiles,
                 std::vector<int32_t>& node_in, std::vector<int32_t>& edge_in)
{
    // caller-order offsets (what scan_offsets_kernel computes on the device when the graphs are not re-ordered)
    node_in.resize((size_t)G + 1); edge_in.resize((size_t)G + 1);
    {
        int32_t a = 0, c = 0;
        for (int g = 0; g < G; g++) { node_in[g] = a; edge_in[g] = c; a += nn[g]; c += ne[g]; }
        node_in[G] = a; edge_in[G] = c;
    }
    constexpr int W = 256, T = 128;
    node_off.assign((size_t)G + 1, 0);
    edge_off.assign((size_t)G + 1, 0);
    tiles.clear();
    tiles.reserve((size_t)G / 4 + 16);
    int32_t pos = 0, epos = 0;
    int order[W], bin_of[W], bin_count[W + 1];
    static thread_local std::vector<int> by_cap[T + 1];     // bins with exactly c free rows
    for (int w0 = 0; w0 < G; w0 += W)
    {
        const int m = std::min(W, G - w0);
        // graphs above a tile first, in caller order, as runs of "external" tiles (bit 30); then the others by size, largest first
        int cnt[T + 2] = {0};
        for (int i = 0; i < m; i++)
        {
            const int n = nn[w0 + i];
            if (n > T)
            {
                node_off[w0 + i] = pos; edge_off[w0 + i] = epos;
                for (int r = 0; r < n; r += T) { tiles.push_back(pos + r); tiles.push_back(std::min(T, n - r) | (1 << 30)); }
                pos += n; epos += ne[w0 + i];
            }
            else cnt[T - n]++;                              // bucket 0 = size 128 ... bucket 128 = size 0
        }
        int start[T + 2];
        start[0] = 0;
        for (int c = 0; c <= T; c++) start[c + 1] = start[c] + cnt[c];
        const int small = start[T + 1];
        for (int i = 0; i < m; i++) { const int n = nn[w0 + i]; if (n <= T) order[start[T - n]++] = i; }
        // best fit: the fullest bin that still takes the graph
        uint64_t mask[3] = {0, 0, 0};                       // capacities 0..128 with a bin
        int bins = 0;
        for (int k = 0; k < small; k++)
        {
            const int i = order[k], n = nn[w0 + i];
            int c = -1;
            if (n > 0)
                for (int word = n >> 6, bit = n & 63; word < 3; word++, bit = 0)
                {
                    const uint64_t mm = mask[word] & (~0ull << bit);
                    if (mm) { c = word * 64 + __builtin_ctzll(mm); break; }
                }
            int b;
            if (n == 0) b = bins ? 0 : -1;                   // an empty graph takes no rows: anywhere
            else if (c >= 0)
            {
                b = by_cap[c].back(); by_cap[c].pop_back();
                if (by_cap[c].empty()) mask[c >> 6] &= ~(1ull << (c & 63));
            }
            else { b = bins++; c = T; }
            if (n > 0)
            {
                const int rest = c - n;
                by_cap[rest].push_back(b);
                mask[rest >> 6] |= 1ull << (rest & 63);
            }
            bin_of[k] = b;
        }
        for (int c = 0; c <= T; c++) by_cap[c].clear();
        // lay the bins out one after the other (stable inside a bin)
        int bstart[W + 2], fill[W + 1], members[W];
        for (int q = 0; q <= bins; q++) bin_count[q] = 0;
        for (int k = 0; k < small; k++) if (bin_of[k] >= 0) bin_count[bin_of[k]]++;
        bstart[0] = 0;
        for (int q = 0; q < bins; q++) { bstart[q + 1] = bstart[q] + bin_count[q]; fill[q] = bstart[q]; }
        for (int k = 0; k < small; k++) if (bin_of[k] >= 0) members[fill[bin_of[k]]++] = k;
        for (int q = 0; q < bins; q++)
        {
            const int first = pos;
            for (int idx = bstart[q]; idx < bstart[q + 1]; idx++)
            {
                const int g = w0 + order[members[idx]];
                node_off[g] = pos; edge_off[g] = epos;
                pos += nn[g]; epos += ne[g];
            }
            if (pos > first) { tiles.push_back(first); tiles.push_back(pos - first); }
        }
        for (int k = 0; k < small; k++)
            if (bin_of[k] < 0) { const int g = w0 + order[k]; node_off[g] = pos; edge_off[g] = epos; epos += ne[g]; }      // empty graphs of a window without bins
    }
    node_off[G] = pos; edge_off[G] = epos;
}

// load_graph + the full forward of `model` on a resident batch, all on stream `s`
int compute_on(flowgnn_ctx* ctx, DeviceBatch& b, cudaStream_t s, int model)
{
    if (b.num_graphs == 0) return 0;
    if (model == MODEL_DGN && !b.has_eigen) { set_last_error("DGN needs node_eigen"); return FG_ERR_INVALID; }
    if ((model == MODEL_GIN || model == MODEL_GCN) && !b.has_attr) { set_last_error("GIN/GCN need edge_attr"); return FG_ERR_INVALID; }
    const bool fixed = ctx->opt.fixed_point != 0;
    if (b.total_nodes == 0 && !fixed)
    {
        // every graph is empty: the mean pool is 0 / 0, and the reference's fp32 flavour carries that NaN through every head
        FG_TRY(b.status.reserve(sizeof(int)));
        FG_TRY(zero_bytes_launch(b.status.ptr, sizeof(int), s));
        FG_TRY(fill_outputs(b.out.as<float>(), std::nanf(""), b.num_graphs, s));
        ctx->last_launches += 2;
        return 0;
    }
    if (fixed && model != MODEL_GIN && model != MODEL_DGN)
    {
        set_last_error("option fixed_point: only GIN / GIN-VN (ap_fixed<16,6>) and DGN (ap_fixed<16,3>) run in the reference's fixed-point arithmetic");
        return FG_ERR_INVALID;
    }
    const int flags = fixed ? 0 : (model == MODEL_GCN) ? (PREP_GCN_NORM | PREP_ROW_DESC) : (model == MODEL_DGN) ? PREP_DGN_EIG : (model == MODEL_GIN || model == MODEL_PNA) ? (PREP_ROW_DESC | PREP_TILES) : 0;
    const bool keep_attr = b.has_attr;
    if (model != MODEL_GIN && model != MODEL_GCN) b.has_attr = false;     // GAT/PNA/DGN kernels take no edge_attr
    ctx->opt.aux = ctx->opt.embed_overlap ? ctx->aux_stream : nullptr;
    ctx->opt.ev_fork = ctx->ev_fork; ctx->opt.ev_join = ctx->ev_join;
    if (ctx->opt.aux) FG_CUDA(cudaEventRecord(ctx->ev_fork, s));
    // graphs re-ordered for tile packing: not for dense batches, whose GIN layers go through the staged gather (it walks graphs in caller order)
    const bool dense = (ctx->opt.gin_staged < 0) ? (b.total_edges >= 6 * b.total_nodes) : (ctx->opt.gin_staged != 0);
    const bool use_perm = ctx->opt.pack_graphs && !(model == MODEL_GIN && dense);
    int rc = prep_batch(b, flags, s, &ctx->last_launches, use_perm);
    b.has_attr = keep_attr;
    FG_TRY(rc);
    ctx->timer.marks = 0;
    ctx->opt.timer = ctx->time_layers ? &ctx->timer : nullptr;
    ctx->opt.timer_group = ctx->time_layers == 2;
    if (fixed)
        return model == MODEL_GIN ? gin_fixed_forward(b, ctx->gin, ctx->sm_count, s, &ctx->last_launches)
                                  : dgn_fixed_forward(b, ctx->dgn, ctx->sm_count, s, &ctx->last_launches);
    switch (model)
    {
    case MODEL_GIN: FG_TRY(gin_forward(b, ctx->gin, ctx->opt, ctx->sm_count, s, &ctx->last_launches)); break;
    case MODEL_GCN: FG_TRY(gcn_forward(b, ctx->gcn, ctx->opt, ctx->sm_count, s, &ctx->last_launches)); break;
    case MODEL_GAT: FG_TRY(gat_forward(b, ctx->gat, ctx->opt, ctx->sm_count, s, &ctx->last_launches)); break;
    case MODEL_PNA: FG_TRY(pna_forward(b, ctx->pna, ctx->opt, ctx->sm_count, s, &ctx->last_launches)); break;
    case MODEL_DGN: FG_TRY(dgn_forward(b, ctx->dgn, ctx->opt, ctx->sm_count, s, &ctx->last_launches)); break;
    }
    return 0;
}

}  // namespace

extern "C" {

int flowgnn_b200_upload_batch(flowgnn_ctx* ctx, int num_graphs, int64_t total_nodes, int64_t total_edges,
                              const int32_t* nums_of_nodes, const int32_t* nums_of_edges, const int32_t* node_feature,
                              const int32_t* edge_list, const int32_t* edge_attr, const float* node_eigen)
{
    FG_TRY(check_ctx(ctx));
    DeviceGuard guard(ctx->device);
    ctx->batch_ready = false;
    FG_TRY(upload_into(ctx->batch, ctx->stream, num_graphs, total_nodes, total_edges, nums_of_nodes, nums_of_edges, node_feature, edge_list,
                       edge_attr, node_eigen));
    ctx->batch_ready = true;
    return 0;
}

int flowgnn_b200_upload_batch_packed(flowgnn_ctx* ctx, int num_graphs, int64_t total_nodes, int64_t total_edges,
                                     const int32_t* nums_of_nodes, const int32_t* nums_of_edges, const uint8_t* node_feature,
                                     const uint16_t* edge_list, const uint8_t* edge_attr, const float* node_eigen)
{
    FG_TRY(check_ctx(ctx));
    DeviceGuard guard(ctx->device);
    ctx->batch_ready = false;
    if ((total_nodes > 0 && !node_feature) || (total_edges > 0 && !edge_list)) { set_last_error("null batch array"); return FG_ERR_INVALID; }
    StagedChunk sc;
    const bool which[4] = {true, true, edge_attr != nullptr, false};
    sc.plan.layout((size_t)std::max<int64_t>(total_nodes, 0), (size_t)std::max<int64_t>(total_edges, 0), which);
    sc.part[0] = node_feature; sc.part[1] = edge_list; sc.part[2] = edge_attr;
    sc.ok[0] = sc.ok[1] = true; sc.ok[2] = edge_attr != nullptr;
    // upload_into only tests the third int32 pointer for "the model's batch has edge attributes": hand it a non-null token
    const int32_t* attr_token = edge_attr ? reinterpret_cast<const int32_t*>(edge_attr) : nullptr;
    FG_TRY(upload_into(ctx->batch, ctx->stream, num_graphs, total_nodes, total_edges, nums_of_nodes, nums_of_edges, nullptr, nullptr, attr_token,
                       node_eigen, &sc));
    ctx->batch_ready = true;
    return 0;
}

int flowgnn_b200_compute(flowgnn_ctx* ctx, int model, float* elapsed_ms)
{
    FG_TRY(check_ctx(ctx));
    if (model < 0 || model >= NUM_MODELS) { set_last_error("unknown model id"); return FG_ERR_INVALID; }
    if (!ctx->loaded[model]) { set_last_error("compute: load_weights has not been called for this model"); return FG_ERR_STATE; }
    if (!ctx->batch_ready) { set_last_error("compute: no batch uploaded"); return FG_ERR_STATE; }
    DeviceGuard guard(ctx->device);
    cudaStream_t s = ctx->stream;
    ctx->last_launches = 0;
    if (elapsed_ms) *elapsed_ms = 0.f;
    if (ctx->batch.num_graphs == 0) return 0;
    if (elapsed_ms) FG_CUDA(cudaEventRecord(ctx->ev0, s));
    FG_TRY(compute_on(ctx, ctx->batch, s, model));
    if (elapsed_ms)
    {
        FG_CUDA(cudaEventRecord(ctx->ev1, s));
        FG_CUDA(cudaEventSynchronize(ctx->ev1));
        FG_CUDA(cudaEventElapsedTime(elapsed_ms, ctx->ev0, ctx->ev1));
        FG_CUDA(cudaGetLastError());
    }
    return 0;
}

int flowgnn_b200_download(flowgnn_ctx* ctx, float* out, int num_graphs)
{
    FG_TRY(check_ctx(ctx));
    if (num_graphs < 0 || num_graphs > ctx->batch.num_graphs || (num_graphs && !out)) { set_last_error("download: bad size"); return FG_ERR_INVALID; }
    DeviceGuard guard(ctx->device);
    if (num_graphs) FG_CUDA(cudaMemcpyAsync(out, ctx->batch.out.ptr, sizeof(float) * (size_t)num_graphs, cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->batch.status.ptr) FG_TRY(check_status(ctx));
    else FG_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int flowgnn_b200_last_launch_count(flowgnn_ctx* ctx) { return ctx ? ctx->last_launches : 0; }
long flowgnn_b200_tile_count(flowgnn_ctx* ctx) { return ctx && ctx->batch_ready && ctx->batch.has_perm ? ctx->batch.tiles_perm_count : 0; }

int flowgnn_b200_last_layer_ms(flowgnn_ctx* ctx, float* out, int max_layers)
{
    if (!ctx || !out || max_layers <= 0) return 0;
    DeviceGuard guard(ctx->device);
    const int n = ctx->timer.marks - 1;
    if (n <= 0) return 0;
    if (cudaEventSynchronize(ctx->timer.ev[n]) != cudaSuccess) return 0;
    int k = 0;
    for (; k < n && k < max_layers; k++)
        if (cudaEventElapsedTime(&out[k], ctx->timer.ev[k], ctx->timer.ev[k + 1]) != cudaSuccess) break;
    return k;
}
void* flowgnn_b200_stream(flowgnn_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

// Page-lock caller-owned host memory so that the host-pointer entry points can overlap its upload with the kernels
// (pageable memory is staged by the driver, synchronously).  The caller unpins before freeing.
int flowgnn_b200_pin_host(void* ptr, size_t bytes)
{
    if (!ptr || !bytes) return 0;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) == cudaSuccess && at.type == cudaMemoryTypeHost) return 0;      // pinned already
    (void)cudaGetLastError();
    FG_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return 0;
}
int flowgnn_b200_unpin_host(void* ptr)
{
    if (!ptr) return 0;
    FG_CUDA(cudaHostUnregister(ptr));
    return 0;
}

void flowgnn_b200_last_transfer_bytes(uint64_t* h2d, uint64_t* d2h)
{
    if (h2d) *h2d = g_h2d_bytes;
    if (d2h) *d2h = g_d2h_bytes;
}

uint32_t flowgnn_b200_narrow_words(const int32_t* src, size_t n, int width, void* dst, int threads)
{
    if ((width != 1 && width != 2) || (n && (!src || !dst))) return 0xFFFFFFFFu;
    constexpr size_t SLICE = 64 * 1024;
    const int jobs = (int)((n + SLICE - 1) / SLICE);
    std::vector<uint32_t> seen((size_t)jobs, 0u);
    HostPool pool(std::max(1, std::min(threads, 64)));
    pool.start(jobs, [&](int j) {
        const size_t i0 = (size_t)j * SLICE, len = std::min(SLICE, n - i0);
        seen[(size_t)j] = width == 1 ? narrow_u8(src + i0, static_cast<uint8_t*>(dst) + i0, len)
                                     : narrow_u16(src + i0, static_cast<uint16_t*>(dst) + i0, len);
    });
    pool.finish();
    uint32_t all = 0;
    for (uint32_t v : seen) all |= v;
    return all;
}

int flowgnn_b200_synchronize(flowgnn_ctx* ctx)
{
    FG_TRY(check_ctx(ctx));
    DeviceGuard guard(ctx->device);
    FG_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

}  // extern "C"

// ---- Part 1: the reference's kernel entry points ---------------------------------------------------

namespace {

struct DefaultCtx {
    flowgnn_ctx* ctx = nullptr;
    ~DefaultCtx() { if (ctx) flowgnn_b200_destroy(ctx); }
};

int default_ctx(flowgnn_ctx** out)
{
    static thread_local DefaultCtx holder;
    int dev = 0;
    FG_CUDA(cudaGetDevice(&dev));
    if (holder.ctx && holder.ctx->device != dev) { flowgnn_b200_destroy(holder.ctx); holder.ctx = nullptr; }
    if (!holder.ctx) FG_TRY(flowgnn_b200_create(&holder.ctx, dev));
    *out = holder.ctx;
    return 0;
}

// FNV-1a over the weight bytes, 8 bytes at a time: re-upload only when the contents change
uint64_t hash_weights(const float* const* w, const size_t* counts, int n, size_t set)
{
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < n; i++)
    {
        const unsigned char* p = reinterpret_cast<const unsigned char*>(w[i] + set * counts[i]);
        const size_t bytes = counts[i] * sizeof(float);
        size_t k = 0;
        // four independent multiply chains (the single FNV chain is latency-bound at ~2.5 GB/s; ~1 MB of weights per call)
        uint64_t a = h, b = h ^ 0x9E3779B97F4A7C15ull, c = h ^ 0xC2B2AE3D27D4EB4Full, d = h ^ 0x165667B19E3779F9ull;
        for (; k + 32 <= bytes; k += 32)
        {
            uint64_t v[4];
            std::memcpy(v, p + k, 32);
            a = (a ^ v[0]) * 1099511628211ull; b = (b ^ v[1]) * 1099511628211ull;
            c = (c ^ v[2]) * 1099511628211ull; d = (d ^ v[3]) * 1099511628211ull;
        }
        h = ((a ^ (b >> 29)) * 1099511628211ull ^ c) * 1099511628211ull ^ (d >> 31);
        for (; k + 8 <= bytes; k += 8) { uint64_t v; std::memcpy(&v, p + k, 8); h = (h ^ v) * 1099511628211ull; }
        for (; k < bytes; k++) h = (h ^ p[k]) * 1099511628211ull;
    }
    return h ? h : 1;
}

// Walk the batch as runs of graphs that share a weight set (reload_weights[g] != 0 starts the next
// set; GIN/src/GIN_compute.cc:49-63) and run each through the extended interface.
// the big arrays in the narrow layout of the packed dataset files instead of the reference ABI's int32 words (flowgnn_b200_compute_graphs_packed)
struct PackedInputs { const uint8_t* feat; const uint16_t* edges; const uint8_t* attr; };

int run_reference_entry(int model, int num_graphs, const int* nn, const int* ne, const int* reload, float* out,
                        const int32_t* feat, const int32_t* edges, const int32_t* attr, const float* eig,
                        const float* const* weights, const size_t* counts, const PackedInputs* packed = nullptr)
{
    if (num_graphs < 0) { set_last_error("num_graphs < 0"); return FG_ERR_INVALID; }
    if (num_graphs == 0) return 0;
    std::vector<int> one_set;
    if (packed)
    {
        if (!packed->feat || !packed->edges) { set_last_error("null argument"); return FG_ERR_INVALID; }
        if (!reload) { one_set.assign((size_t)num_graphs, 0); one_set[0] = 1; reload = one_set.data(); }
        // upload_into only asks whether the int32 attribute pointer is null ("this model's batch has edge attributes")
        attr = packed->attr ? reinterpret_cast<const int32_t*>(packed->attr) : nullptr;
    }
    else if (!feat || !edges) { set_last_error("null argument"); return FG_ERR_INVALID; }
    if (!nn || !ne || !reload || !out) { set_last_error("null argument"); return FG_ERR_INVALID; }
    flowgnn_ctx* ctx = nullptr;
    FG_TRY(default_ctx(&ctx));
    g_h2d_bytes = 0; g_d2h_bytes = 0;
    const int nw = kNumWeights[model];
    // the counts drive every offset, on the host (chunk extents, the narrowing slices, the tile packing) and on the device: check them
    // before anything reads through them
    for (int k = 0; k < num_graphs; k++)
        if (nn[k] < 0 || ne[k] < 0)
        {
            set_last_error("negative entry in nums_of_nodes / nums_of_edges (graph " + std::to_string(k) + ")");
            return FG_ERR_INVALID;
        }
    long set = -1;
    int64_t node_base = 0, edge_base = 0;
    int g = 0;
    while (g < num_graphs)
    {
        if (reload[g]) set++;
        if (set < 0) { set_last_error("reload_weights[0] must be non-zero (the reference would index weight set -1)"); return FG_ERR_INVALID; }
        int g1 = g + 1;
        int64_t n_run = nn[g], e_run = ne[g];
        while (g1 < num_graphs && !reload[g1]) { n_run += nn[g1]; e_run += ne[g1]; g1++; }

        // Cut the run into chunks and alternate between two device batches: the H2D copy of chunk i+1 (copy stream)
        // overlaps the kernels of chunk i (compute stream).  Chunks are whole graphs, so results do not depend on the cut;
        // the first chunk is small so that the kernels start early, and the weight hash runs while it is in flight.
        // the reference's GAT reads node features from the START of the batch buffer for every graph.
        const bool gat_bug = (model == MODEL_GAT) && ctx->opt.gat_node_offset_bug;
        // Narrowed uploads (host_stage.h).  Option / FLOWGNN_B200_HOST_STAGE: 0 off, else a mask of the arrays to narrow (1 node_feature,
        // 2 edge_list, 4 edge_attr).  Automatic: everything when the caller's arrays are pageable (the plain copy would be staged by the
        // driver, one thread, synchronously: 4.0 M graphs/s on the 41k-graph GIN batch, narrowed 16 M); for page-locked arrays everything
        // when 8 host threads are free for this GPU, node_feature + edge_attr (4 bytes -> 1) with 6 or 7, else the plain copies.  Measured
        // on a 16-core B200 box, ms per call of the 41k-graph batch (profiles/r2k_e2e_probe.txt): plain copies 2.90; everything narrowed
        // 2.46 / 2.49 / 2.91 / 3.26 with 12 / 8 / 6 / 4 threads; node_feature + edge_attr only 2.55 / 2.57 / 2.55 / 3.06.  Two ranks on a
        // 24-core box (9 threads each, profiles/r2k_e2e_scale_2gpu.txt): plain 2.96, features + attributes 2.93, everything 3.10 -- the
        // pools of several ranks compete for the host's memory system, so edge_list is narrowed only by a rank that is alone.
        int stage_mask = ctx->opt.host_stage;
        if (const char* e = std::getenv("FLOWGNN_B200_HOST_STAGE")) stage_mask = std::atoi(e);
        bool pinned = false;
        if (stage_mask != 0)
        {
            cudaPointerAttributes at;
            pinned = cudaPointerGetAttributes(&at, packed ? static_cast<const void*>(packed->feat) : feat) == cudaSuccess && at.type != cudaMemoryTypeUnregistered;
            (void)cudaGetLastError();
        }
        if (stage_mask < 0)
        {
            const int threads = HostPool::default_threads();
            const bool alone = HostPool::local_ranks() == 1;     // several ranks' pools share the host's memory system
            stage_mask = !pinned || (alone && threads >= 8) ? 7 : threads >= 6 ? 5 : 0;
        }
        stage_mask &= 7;
        if (packed) stage_mask = 7;                          // nothing to narrow: the pool only computes the chunks' tile packing
        // DGN's node_eigen keeps its format; a pageable array can go through the block too, so that the pool reads it instead of the driver
        // (off unless FLOWGNN_B200_STAGE_EIG=1: measured on the 41k-graph DGN batch from pageable memory, 5.05 ms per call with node_eigen
        // through the block -- the middle chunk's kernels then take 2.0 ms instead of 1.2 ms, not understood -- against 4.48 ms with the
        // plain copy of the 16 B per node, which the driver stages)
        static const bool eig_through_block = [] { const char* e = std::getenv("FLOWGNN_B200_STAGE_EIG"); return e && std::atoi(e) != 0; }();
        const bool stage_eig = stage_mask != 0 && eig && !pinned && !packed && eig_through_block;
        const bool staged = stage_mask != 0;
        const int run_graphs = g1 - g;
        int nchunks = 1;
        int bounds[17];
        for (int i = 0; i <= 16; i++) bounds[i] = g1;
        bounds[0] = g;
        {
            // graded schedule (a small first chunk starts the kernels early): weights 1, 2, 3, 3, ... for the plain copies (the upload is the
            // longer stage: 3.87 / 3.18 / 3.11 / 3.19 / 3.54 ms for 1 / 2 / 3 / 4 / 6 chunks of the 41k-graph batch), 1, 3, 6, 6, ... for the
            // narrowed upload (the kernels are: 2.60 ms with 1 : 2 : 3, 2.44-2.50 ms with 1 : 3 : 6; 1 : 2 : 4 2.52, four chunks 2.68);
            // FLOWGNN_B200_GRADE="1,2,4" overrides the weights (and the chunk count) for measurements
            int want = run_graphs >= 16384 ? 3 : run_graphs >= 8192 ? 2 : 1;
            const int wmax = staged ? 6 : 3;
            // large inputs (PNA on 437,929 graphs: 945 MB; GIN-VN on 40,000 hep10k graphs: 785 MB): weight-3 chunks of about 100 MB of caller
            // bytes, so that the first chunk (what nothing overlaps) and the last chunk's kernels (what nothing follows) stay small
            {
                const int64_t run_bytes = (int64_t)sizeof(int) * (ND_FEATURE * n_run + (attr ? 5 : 2) * e_run) + (eig ? 16 * n_run : 0);
                const int64_t unit = (96 << 20) / wmax;                      // caller bytes per unit of weight
                const int64_t total_w = (run_bytes + unit - 1) / unit;
                const int64_t head = staged ? 4 : 3;                          // weight of the two graded chunks in front
                if (want == 3 && total_w > head + wmax) want = (int)std::min<int64_t>(16, 2 + (total_w - head + wmax - 1) / wmax);
            }
            if (const char* e = std::getenv("FLOWGNN_B200_CHUNKS")) want = std::max(1, std::min(16, std::atoi(e)));
            int weight[16];
            for (int i = 0; i < 16; i++) weight[i] = staged ? (i == 0 ? 1 : i == 1 ? 3 : 6) : std::min(i + 1, 3);
            if (const char* e = std::getenv("FLOWGNN_B200_GRADE"))
            {
                int k = 0;
                for (const char* p = e; *p && k < 16; k++)
                {
                    weight[k] = std::max(1, std::atoi(p));
                    while (*p && *p != ',') p++;
                    if (*p == ',') p++;
                }
                if (k > 0 && run_graphs >= 8192) want = k;
            }
            if (run_graphs < 2 * want) want = 1;
            nchunks = want;
            int total_w = 0;
            for (int i = 0; i < want; i++) total_w += weight[i];
            int acc_w = 0;
            for (int i = 0; i < want - 1; i++)
            {
                acc_w += weight[i];
                bounds[i + 1] = g + (int)((int64_t)run_graphs * acc_w / total_w);
            }
            bounds[want] = g1;
        }
        if ((size_t)run_graphs > ctx->h_out_cap)
        {
            if (ctx->h_out) cudaFreeHost(ctx->h_out);
            ctx->h_out = nullptr; ctx->h_out_cap = 0;
            FG_CUDA(cudaMallocHost(&ctx->h_out, sizeof(float) * (size_t)run_graphs));
            ctx->h_out_cap = (size_t)run_graphs;
        }
        ctx->last_launches = 0;
        ctx->timer.marks = 0;
        // chunk extents in the caller's arrays
        int64_t cnode[17], cedge[17];
        cnode[0] = node_base; cedge[0] = edge_base;
        for (int ci = 0; ci < nchunks; ci++)
        {
            int64_t n_c = 0, e_c = 0;
            for (int k = bounds[ci]; k < bounds[ci + 1]; k++) { n_c += nn[k]; e_c += ne[k]; }
            cnode[ci + 1] = cnode[ci] + n_c; cedge[ci + 1] = cedge[ci] + e_c;
        }
        constexpr int P = flowgnn_ctx::PIPE;
        auto chunk_feat = [&](int ci) { return gat_bug ? feat : feat + ND_FEATURE * cnode[ci]; };
        NarrowRun::Chunk nchunk[NarrowRun::MAX_CHUNKS];
        if (staged)
        {
            const bool which[4] = {(stage_mask & 1) != 0, (stage_mask & 2) != 0, (stage_mask & 4) != 0 && attr, stage_eig};
            for (int ci = 0; ci < nchunks; ci++)
            {
                nchunk[ci].plan.layout((size_t)(cnode[ci + 1] - cnode[ci]), (size_t)(cedge[ci + 1] - cedge[ci]), which);
                if (packed) continue;                        // no slices to narrow
                nchunk[ci].src[0] = which[0] ? chunk_feat(ci) : nullptr;
                nchunk[ci].src[1] = which[1] ? edges + 2 * cedge[ci] : nullptr;
                nchunk[ci].src[2] = which[2] ? attr + 3 * cedge[ci] : nullptr;
                nchunk[ci].eig = which[3] ? eig + 4 * cnode[ci] : nullptr;
            }
        }
        // FLOWGNN_B200_E2E_TRACE=1: host and device timeline of the call on stderr (tools/e2e_probe.py)
        struct Trace {
            bool on = std::getenv("FLOWGNN_B200_E2E_TRACE") != nullptr;
            std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
            std::vector<std::pair<std::string, double>> host;
            std::vector<std::pair<std::string, cudaEvent_t>> dev;
            void h(const std::string& what) { if (on) host.push_back({what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count()}); }
            void d(const std::string& what, cudaStream_t st) { if (!on) return; cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); dev.push_back({what, e}); }
            void dump()
            {
                if (!on) return;
                for (auto& x : host) std::fprintf(stderr, "  host %-28s %8.3f ms\n", x.first.c_str(), x.second);
                for (size_t i = 1; i < dev.size(); i++)
                {
                    float ms = 0.f; cudaEventElapsedTime(&ms, dev[0].second, dev[i].second);
                    std::fprintf(stderr, "  dev  %-28s %8.3f ms after the first device mark\n", dev[i].first.c_str(), ms);
                }
                for (auto& x : dev) cudaEventDestroy(x.second);
            }
        } trace;
        trace.d("start (copy stream)", ctx->copy_stream);
        auto issue_upload = [&](int ci) -> int {
            const int c0 = bounds[ci], c1 = bounds[ci + 1];
            DeviceBatch& db = ctx->pipe[ci % P];
            cudaEvent_t free_ev = ci >= P ? ctx->buf_free[ci % P] : nullptr;
            StagedChunk sc;
            if (staged)
            {
                ctx->narrow.wait_chunk(*ctx->pool, ci, sc.ok);                  // (this thread narrows slices too while it waits)
                trace.h("narrowed chunk " + std::to_string(ci));
                sc.h = ctx->stage_h + ctx->narrow.chunk(ci).base;
                sc.plan = ctx->narrow.chunk(ci).plan;
                sc.done = ctx->stage_done;
                if (packed)
                {
                    sc.h = nullptr; sc.done = nullptr;
                    sc.part[0] = packed->feat + (gat_bug ? 0 : ND_FEATURE * cnode[ci]);
                    sc.part[1] = packed->edges + 2 * cedge[ci];
                    sc.part[2] = packed->attr ? packed->attr + 3 * cedge[ci] : nullptr;
                    sc.ok[0] = sc.ok[1] = true; sc.ok[2] = packed->attr != nullptr;
                }
            }
            else if (free_ev) { FG_CUDA(cudaStreamWaitEvent(ctx->copy_stream, free_ev, 0)); free_ev = nullptr; }
            FG_TRY(upload_into(db, ctx->copy_stream, c1 - c0, cnode[ci + 1] - cnode[ci], cedge[ci + 1] - cedge[ci], nn + c0, ne + c0,
                               packed ? nullptr : chunk_feat(ci), packed ? nullptr : edges + 2 * cedge[ci],
                               packed ? attr : attr ? attr + 3 * cedge[ci] : nullptr, eig ? eig + 4 * cnode[ci] : nullptr,
                               staged ? &sc : nullptr, free_ev, staged ? &ctx->packed[ci] : nullptr));
            FG_CUDA(cudaEventRecord(ctx->up_done[ci % P], ctx->copy_stream));
            trace.h("upload issued " + std::to_string(ci));
            trace.d("upload done " + std::to_string(ci), ctx->copy_stream);
            return 0;
        };
        // any early return below leaves copies and kernels in flight that read the caller's buffers and write the pinned
        // staging words: drain both streams (and the narrowing run) before handing control back
        struct Drain {
            flowgnn_ctx* c; bool armed = true;
            ~Drain()
            {
                if (c->pool) c->narrow.finish(*c->pool);
                if (!armed) return;
                cudaStreamSynchronize(c->copy_stream); cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->d2h_stream);
            }
        } drain{ctx};
        auto ensure_weights = [&]() -> int {
            const uint64_t h = hash_weights(weights, counts, nw, (size_t)set);
            if (!ctx->loaded[model] || ctx->weight_hash[model] != h)
            {
                std::vector<const float*> ptrs(nw);
                for (int i = 0; i < nw; i++) ptrs[i] = weights[i] + (size_t)set * counts[i];
                FG_TRY(flowgnn_b200_load_weights(ctx, model, ptrs.data(), nw));
                ctx->weight_hash[model] = h;
            }
            return 0;
        };
        auto issue_compute = [&](int ci) -> int {
            const int c0 = bounds[ci], c1 = bounds[ci + 1];
            DeviceBatch& db = ctx->pipe[ci % P];
            FG_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->up_done[ci % P], 0));
            trace.d("compute begins " + std::to_string(ci), ctx->stream);
            FG_TRY(compute_on(ctx, db, ctx->stream, model));
            trace.d("compute ends " + std::to_string(ci), ctx->stream);
            trace.h("compute issued " + std::to_string(ci));
            FG_CUDA(cudaEventRecord(ctx->fwd_done[ci % P], ctx->stream));
            FG_CUDA(cudaStreamWaitEvent(ctx->d2h_stream, ctx->fwd_done[ci % P], 0));
            if (c1 > c0)
            {
                FG_CUDA(cudaMemcpyAsync(ctx->h_out + (c0 - g), db.out.ptr, sizeof(float) * (size_t)(c1 - c0), cudaMemcpyDeviceToHost, ctx->d2h_stream));
                FG_CUDA(cudaMemcpyAsync(ctx->h_status + ci, db.status.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->d2h_stream));
                g_d2h_bytes += sizeof(float) * (size_t)(c1 - c0) + sizeof(int);
            }
            else ctx->h_status[ci] = 0;
            FG_CUDA(cudaEventRecord(ctx->buf_free[ci % P], ctx->d2h_stream));     // the batch is free once its predictions have left
            return 0;
        };
        if (staged)
        {
            // the pool narrows all chunks in order; the weight hash runs next to chunk 0's narrowing
            FG_TRY(stage_begin_run(ctx, nchunk, nchunks, [&](int ci) {
                PackOut& pk = ctx->packed[ci];                                  // the chunk's tile packing, off this thread's critical path
                const int c0 = bounds[ci], c1 = bounds[ci + 1];
                if (c1 > c0 && cnode[ci + 1] > cnode[ci]) pack_graphs(nn + c0, ne + c0, c1 - c0, pk.po, pk.pe, pk.pt, pk.pin, pk.pie);
            }));
            trace.h("narrowing queued");
            FG_TRY(ensure_weights());
            trace.h("weights checked");
            for (int ci = 0; ci < nchunks; ci++)
            {
                FG_TRY(issue_upload(ci));
                FG_TRY(issue_compute(ci));
            }
        }
        else
        {
            FG_TRY(issue_upload(0));
            FG_TRY(ensure_weights());                                                   // runs while chunk 0 is in flight
            for (int ci = 0; ci < nchunks; ci++)
            {
                if (ci == 0) for (int k = 1; k < P && k < nchunks; k++) FG_TRY(issue_upload(k));   // the other buffers fill right away
                FG_TRY(issue_compute(ci));
                if (ci + P < nchunks) FG_TRY(issue_upload(ci + P));                     // reuses this chunk's buffer once it is free
            }
        }
        FG_CUDA(cudaStreamSynchronize(ctx->d2h_stream));                  // (behind the last chunk's kernels)
        FG_CUDA(cudaStreamSynchronize(ctx->stream));
        trace.h("synchronised");
        trace.dump();
        drain.armed = false;
        FG_CUDA(cudaGetLastError());
        int st = 0;
        for (int ci = 0; ci < nchunks; ci++) st |= ctx->h_status[ci];
        FG_TRY(status_error(st));
        std::memcpy(out + g, ctx->h_out, sizeof(float) * (size_t)run_graphs);
        node_base += n_run; edge_base += e_run;
        g = g1;
    }
    return 0;
}

}  // namespace

extern "C" {

// The host-pointer pipeline of the entry points below (chunks alternating between two device batches, uploads overlapping kernels,
// one weight set) for a caller that keeps its dataset in the packed layout: uint8 node features, uint16 graph-local edge ids, uint8 bond
// attributes (NULL for GAT / PNA / DGN), fp32 node_eigen (DGN, else NULL).  `weights`: the model's arrays in the order of its
// <MODEL>_compute_graphs argument list.
int flowgnn_b200_compute_graphs_packed(int model, int num_graphs, const int32_t* nums_of_nodes, const int32_t* nums_of_edges, float* out,
                                       const uint8_t* node_feature, const uint16_t* edge_list, const uint8_t* edge_attr, const float* node_eigen,
                                       const float* const* weights, int num_weights)
{
    static const size_t counts[NUM_MODELS][11] = {
        {173 * 100, 5 * 13 * 100, 5 * 200 * 100, 5 * 200, 5 * 100 * 200, 5 * 100, 100, 1},
        {173 * 100, 5 * 13 * 100, 5 * 100 * 100, 500, 500, 500, 500, 500, 500, 100, 1},
        {5 * 64, 5 * 64, 5 * 4096, 5 * 4096, 16, 1},
        {173 * 80, 4 * 80 * 12 * 80, 320, 40 * 80, 40, 20 * 40, 20, 20, 1, 1},
        {9 * 119 * 100, 4 * 100 * 200, 400, 5000, 50, 1250, 25, 25, 1}};
    if (model < 0 || model >= NUM_MODELS) { set_last_error("unknown model id"); return FG_ERR_INVALID; }
    if (!weights || num_weights != kNumWeights[model]) { set_last_error("expected " + std::to_string(kNumWeights[model]) + " weight arrays"); return FG_ERR_INVALID; }
    for (int i = 0; i < num_weights; i++) if (!weights[i]) { set_last_error("null weight argument"); return FG_ERR_INVALID; }
    if ((model == MODEL_GIN || model == MODEL_GCN) && !edge_attr && num_graphs > 0) { set_last_error("GIN / GCN need edge_attr"); return FG_ERR_INVALID; }
    if (model == MODEL_DGN && !node_eigen && num_graphs > 0) { set_last_error("DGN needs node_eigen"); return FG_ERR_INVALID; }
    const PackedInputs in{node_feature, edge_list, (model == MODEL_GIN || model == MODEL_GCN) ? edge_attr : nullptr};
    return run_reference_entry(model, num_graphs, nums_of_nodes, nums_of_edges, nullptr, out, nullptr, nullptr, nullptr,
                               model == MODEL_DGN ? node_eigen : nullptr, weights, counts[model], &in);
}

int GIN_compute_graphs(int num_graphs, int* nums_of_nodes, int* nums_of_edges, int* reload_weights, float* out,
                       const int32_t* node_feature_in, const int32_t* edge_list_in, const int32_t* edge_attr_in,
                       const float* node_embedding_weight_in, const float* edge_embedding_weight_in,
                       const float* node_mlp_1_weights, const float* node_mlp_1_bias, const float* node_mlp_2_weights,
                       const float* node_mlp_2_bias, const float* graph_pred_weights_in, const float* graph_pred_bias_in)
{
    const float* w[] = {node_embedding_weight_in, edge_embedding_weight_in, node_mlp_1_weights, node_mlp_1_bias,
                        node_mlp_2_weights, node_mlp_2_bias, graph_pred_weights_in, graph_pred_bias_in};
    const size_t counts[] = {173 * 100, 5 * 13 * 100, 5 * 200 * 100, 5 * 200, 5 * 100 * 200, 5 * 100, 100, 1};
    for (const float* p : w) if (!p) { set_last_error("null weight argument"); return FG_ERR_INVALID; }
    if (!edge_attr_in && num_graphs > 0) { set_last_error("GIN needs edge_attr_in"); return FG_ERR_INVALID; }
    return run_reference_entry(MODEL_GIN, num_graphs, nums_of_nodes, nums_of_edges, reload_weights, out, node_feature_in, edge_list_in,
                               edge_attr_in, nullptr, w, counts);
}

// The FPGA build's kernel ABI: FM_TYPE / WT_TYPE = ap_fixed<16,6>, i.e. every weight and result is an int16 bit pattern
// (value = raw / 1024; GIN/src/dcl.h:58-59, 77-95).  Runs gin_fixed.cu, which reproduces that arithmetic bit for bit.
int GIN_compute_graphs_fixed(int num_graphs, int* nums_of_nodes, int* nums_of_edges, int* reload_weights, int16_t* out,
                             const int32_t* node_feature_in, const int32_t* edge_list_in, const int32_t* edge_attr_in,
                             const int16_t* node_embedding_weight_in, const int16_t* edge_embedding_weight_in,
                             const int16_t* node_mlp_1_weights, const int16_t* node_mlp_1_bias, const int16_t* node_mlp_2_weights,
                             const int16_t* node_mlp_2_bias, const int16_t* graph_pred_weights_in, const int16_t* graph_pred_bias_in)
{
    const int16_t* w16[] = {node_embedding_weight_in, edge_embedding_weight_in, node_mlp_1_weights, node_mlp_1_bias,
                            node_mlp_2_weights, node_mlp_2_bias, graph_pred_weights_in, graph_pred_bias_in};
    const size_t counts[] = {173 * 100, 5 * 13 * 100, 5 * 200 * 100, 5 * 200, 5 * 100 * 200, 5 * 100, 100, 1};
    if (num_graphs < 0) { set_last_error("num_graphs < 0"); return FG_ERR_INVALID; }
    if (num_graphs == 0) return 0;
    for (const int16_t* p : w16) if (!p) { set_last_error("null weight argument"); return FG_ERR_INVALID; }
    if (!reload_weights || !out || !edge_attr_in) { set_last_error("null argument"); return FG_ERR_INVALID; }
    size_t sets = 0;
    for (int g = 0; g < num_graphs; g++) sets += reload_weights[g] != 0;
    // raw / 1024 is exact in fp32, and load_gin's floor(x * 1024) recovers the same bits
    std::vector<std::vector<float>> wf(8);
    const float* w[8];
    for (int i = 0; i < 8; i++)
    {
        wf[i].resize(std::max<size_t>(sets, 1) * counts[i]);
        for (size_t k = 0; k < sets * counts[i]; k++) wf[i][k] = (float)w16[i][k] * (1.0f / 1024.0f);
        w[i] = wf[i].data();
    }
    std::vector<float> outf((size_t)num_graphs);
    flowgnn_ctx* ctx = nullptr;
    FG_TRY(default_ctx(&ctx));
    const int keep = ctx->opt.fixed_point;
    ctx->opt.fixed_point = 1;
    const int rc = run_reference_entry(MODEL_GIN, num_graphs, nums_of_nodes, nums_of_edges, reload_weights, outf.data(), node_feature_in,
                                       edge_list_in, edge_attr_in, nullptr, w, counts);
    ctx->opt.fixed_point = keep;
    FG_TRY(rc);
    for (int g = 0; g < num_graphs; g++) out[g] = (int16_t)lrintf(outf[g] * 1024.0f);
    return 0;
}

int GCN_compute_graphs(int num_graphs, int* nums_of_nodes, int* nums_of_edges, int* reload_weights, float* out,
                       const int32_t* node_feature_in, const int32_t* edge_list_in, const int32_t* edge_attr_in,
                       const float* node_embedding_weight_in, const float* edge_embedding_weight_in, const float* convs_weight_in,
                       const float* convs_bias_in, const float* convs_root_emb_weight_in, const float* bn_weight_in,
                       const float* bn_bias_in, const float* bn_mean_in, const float* bn_var_in, const float* graph_pred_weights_in,
                       const float* graph_pred_bias_in)
{
    const float* w[] = {node_embedding_weight_in, edge_embedding_weight_in, convs_weight_in, convs_bias_in, convs_root_emb_weight_in,
                        bn_weight_in, bn_bias_in, bn_mean_in, bn_var_in, graph_pred_weights_in, graph_pred_bias_in};
    const size_t counts[] = {173 * 100, 5 * 13 * 100, 5 * 100 * 100, 500, 500, 500, 500, 500, 500, 100, 1};
    for (const float* p : w) if (!p) { set_last_error("null weight argument"); return FG_ERR_INVALID; }
    if (!edge_attr_in && num_graphs > 0) { set_last_error("GCN needs edge_attr_in"); return FG_ERR_INVALID; }
    return run_reference_entry(MODEL_GCN, num_graphs, nums_of_nodes, nums_of_edges, reload_weights, out, node_feature_in, edge_list_in,
                               edge_attr_in, nullptr, w, counts);
}

int GAT_compute_graphs(int num_graphs, int* nums_of_nodes, int* nums_of_edges, int* reload_weights, float* out,
                       const int32_t* node_feature_in, const int32_t* edge_list_in, const float* scoring_fn_target_in,
                       const float* scoring_fn_source_in, const float* linear_proj_weights_in, const float* skip_proj_weights_in,
                       const float* graph_pred_weights_in, const float* graph_pred_bias_in)
{
    const float* w[] = {scoring_fn_target_in, scoring_fn_source_in, linear_proj_weights_in, skip_proj_weights_in,
                        graph_pred_weights_in, graph_pred_bias_in};
    const size_t counts[] = {5 * 64, 5 * 64, 5 * 4096, 5 * 4096, 16, 1};
    for (const float* p : w) if (!p) { set_last_error("null weight argument"); return FG_ERR_INVALID; }
    return run_reference_entry(MODEL_GAT, num_graphs, nums_of_nodes, nums_of_edges, reload_weights, out, node_feature_in, edge_list_in,
                               nullptr, nullptr, w, counts);
}

int PNA_compute_graphs(int num_graphs, int* nums_of_nodes, int* nums_of_edges, int* reload_weights, float* out,
                       const int32_t* node_feature_in, const int32_t* edge_list_in, const float* node_embedding_weight_in,
                       const float* node_conv_weights_in, const float* node_conv_bias_in, const float* graph_mlp_1_weights_in,
                       const float* graph_mlp_1_bias_in, const float* graph_mlp_2_weights_in, const float* graph_mlp_2_bias_in,
                       const float* graph_mlp_3_weights_in, const float* graph_mlp_3_bias_in, const float* avg_deg_in)
{
    const float* w[] = {node_embedding_weight_in, node_conv_weights_in, node_conv_bias_in, graph_mlp_1_weights_in, graph_mlp_1_bias_in,
                        graph_mlp_2_weights_in, graph_mlp_2_bias_in, graph_mlp_3_weights_in, graph_mlp_3_bias_in, avg_deg_in};
    const size_t counts[] = {173 * 80, 4 * 80 * 12 * 80, 320, 40 * 80, 40, 20 * 40, 20, 20, 1, 1};
    for (const float* p : w) if (!p) { set_last_error("null weight argument"); return FG_ERR_INVALID; }
    return run_reference_entry(MODEL_PNA, num_graphs, nums_of_nodes, nums_of_edges, reload_weights, out, node_feature_in, edge_list_in,
                               nullptr, nullptr, w, counts);
}

int DGN_compute_graphs(int num_graphs, int* nums_of_nodes, int* nums_of_edges, int* reload_weights, float* out,
                       const int32_t* node_feature_in, const float* node_eigen_in, const int32_t* edge_list_in,
                       const float* embedding_h_atom_embedding_list_weights_in,
                       const float* layers_posttrans_fully_connected_0_linear_weight_in,
                       const float* layers_posttrans_fully_connected_0_linear_bias_in, const float* MLP_layer_FC_layers_0_weight_in,
                       const float* MLP_layer_FC_layers_0_bias_in, const float* MLP_layer_FC_layers_1_weight_in,
                       const float* MLP_layer_FC_layers_1_bias_in, const float* MLP_layer_FC_layers_2_weight_in,
                       const float* MLP_layer_FC_layers_2_bias_in)
{
    const float* w[] = {embedding_h_atom_embedding_list_weights_in, layers_posttrans_fully_connected_0_linear_weight_in,
                        layers_posttrans_fully_connected_0_linear_bias_in, MLP_layer_FC_layers_0_weight_in, MLP_layer_FC_layers_0_bias_in,
                        MLP_layer_FC_layers_1_weight_in, MLP_layer_FC_layers_1_bias_in, MLP_layer_FC_layers_2_weight_in,
                        MLP_layer_FC_layers_2_bias_in};
    const size_t counts[] = {9 * 119 * 100, 4 * 100 * 200, 400, 5000, 50, 1250, 25, 25, 1};
    for (const float* p : w) if (!p) { set_last_error("null weight argument"); return FG_ERR_INVALID; }
    if (!node_eigen_in && num_graphs > 0) { set_last_error("DGN needs node_eigen_in"); return FG_ERR_INVALID; }
    return run_reference_entry(MODEL_DGN, num_graphs, nums_of_nodes, nums_of_edges, reload_weights, out, node_feature_in, edge_list_in,
                               nullptr, node_eigen_in, w, counts);
}

}  // extern "C"

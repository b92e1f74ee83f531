// Internal structures shared by the C-ABI (api.cu), the graph preprocessing (prep.cu) and the
// per-model layer kernels.  Nothing here is part of the public interface (include/flowgnn_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace fg {

enum ModelId { MODEL_GIN = 0, MODEL_GCN = 1, MODEL_GAT = 2, MODEL_PNA = 3, MODEL_DGN = 4, NUM_MODELS = 5 };

// One growable device allocation.
struct DevBuf {
    void* ptr = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);      // grows (never shrinks); contents are NOT preserved
    void release();
    template <typename T> T* as() const { return static_cast<T*>(ptr); }
};

// The batch as the reference's host hands it to the kernel (GIN/src/host.cc:141-182), resident in HBM,
// plus everything `load_graph` derives from it (GIN/src/load_inputs.cc:87-172 and per-model variants).
struct DeviceBatch {
    int num_graphs = 0;
    long total_nodes = 0;
    long total_edges = 0;
    bool has_attr = false;
    bool has_eigen = false;
    int max_graph_nodes = 0;               // largest nums_of_nodes entry (host-side scan in upload_into, api.cu)

    // inputs (caller layout)
    DevBuf nums_of_nodes, nums_of_edges;   // int32 [G]
    DevBuf node_feature;                   // int32 [N][9]
    DevBuf edge_list;                      // int32 [E][2], graph-local ids
    DevBuf edge_attr;                      // int32 [E][3]
    DevBuf node_eigen;                     // float [N][4]
    DevBuf packed_in;                      // host-pointer entry points: one chunk's narrowed inputs (host_stage.h), widened by unpack_inputs_kernel

    // load_graph outputs: CSR by DESTINATION over global node ids, in-edges ordered (source, list order)
    DevBuf node_off, edge_off;             // int32 [G+1] first row / in-edge position of every graph (exclusive prefix sums in the order the rows are stored)
    // Tile packing (api.cu::pack_graphs, GIN / PNA on sparse graphs): graphs re-ordered inside windows of 256 so that whole graphs fill
    // the 128-row tiles to ~98 % instead of ~89 %.  Computed on the host at upload; prep.cu uses it when the model has tiles.
    DevBuf node_off_perm, edge_off_perm;   // int32 [G+1] re-ordered offsets, indexed by the caller's graph number
    DevBuf tiles_perm;                     // int2 [tiles_perm_count] + int32 count
    long tiles_perm_count = 0;
    bool has_perm = false, perm_active = false;
    DevBuf node_off_in, edge_off_in;       // int32 [G+1] caller-order offsets (where a re-ordered graph's inputs are read)
    DevBuf node_map;                       // int32 [N] caller-order node index of every re-ordered row (embedding lookup)
    int32_t* h_pack = nullptr; size_t h_pack_cap = 0;      // pinned staging of node_off_perm | edge_off_perm | tiles_perm (words)
    cudaEvent_t h_pack_done = nullptr;                     // recorded behind their copies
    DevBuf in_ptr;                         // int32 [N+1]
    DevBuf src;                            // int32 [E] global source node id
    DevBuf code;                           // uint8 [E] bond-attribute triple a0*12 + a1*2 + a2
    DevBuf edge_w;                         // float [E]  GCN: norm = dis[u]*dis[v]; DGN: eig_w = phi_u - phi_v
    DevBuf out_deg;                        // int32 [N]
    DevBuf node_w0, node_w1;               // float [N]  DGN: sum|eig_w|, sum eig_w over in-edges
    DevBuf row_desc;                       // int4 [N]  GIN: first four in-edges of every node, packed (prep.cu)
    DevBuf row_desc_sorted;                // int4 [N]  GIN: the descriptors of every tile ordered by in-degree, row position in .y bits 24..30 (prep.cu)
    DevBuf row_desc0;                      // int4 [N]  GIN, dense graphs: "no in-edges" descriptors (node MLP launch after the staged gather)
    DevBuf sort_tmp;                       // int32 [E] scratch for the two-pass stable sort
    DevBuf big_tab;                        // int32 [3][N] CSR-build tables of graphs above 1,024 nodes (allocated only if there is one)
    DevBuf status;                         // int32 [1] device-side limit violations
    DevBuf tiles;                          // int2 [max_tiles] GIN: graph-aligned tiles (first node, rows | external << 30) (prep.cu)
    DevBuf tile_count;                     // int32 [1] number of tiles
    long max_tiles = 0;                    // host-side upper bound of tile_count (grid sizing)

    // activations
    DevBuf act[4];                         // float [N][<=100] ping/pong (+2 extra for GAT)
    DevBuf score[4];                       // float [N][4] GAT source/target scores ping/pong
    DevBuf node_dot;                       // float [N]  GIN: <h'_v, w_pred> of the last layer (gin_tc2.cu)
    DevBuf apack;                          // PNA tensor-core path: bf16 hi/lo aggregate blocks [tiles][5][32768] (pna_tc.cu)
    DevBuf nonfinite;                      // uint8 [N]  PNA tensor-core path: rows whose aggregates are not finite
    DevBuf out;                            // float [G]

    void release();
};

enum PrepFlags { PREP_GCN_NORM = 1, PREP_DGN_EIG = 2, PREP_ROW_DESC = 4, PREP_TILES = 8 };

// graph preprocessing: offsets scan + per-graph CSR build (prep.cu)
int prep_batch(DeviceBatch& b, int flags, cudaStream_t stream, int* launches = nullptr, bool use_perm = false);
int node_map_launch(DeviceBatch& b, cudaStream_t stream);
int zero_bytes_launch(void* p, size_t bytes, cudaStream_t stream);   // a kernel, not a copy-engine memset (prep.cu)
int unpack_inputs_launch(const uint8_t* block, size_t off_edge, size_t off_attr, int32_t* feat, size_t n_feat, int32_t* edges, size_t n_edge,
                         int32_t* attr, size_t n_attr, cudaStream_t stream);

// ---- per-model device weights, repacked once by load_weights (api.cu) ----------------------------
struct GinWeights {
    DevBuf ne_table;     // [173][100]
    DevBuf ne_table4;    // [431][100] combined tables of embed4_kernel (layers.cuh)
    DevBuf ee_comb;      // [5][60][100]  ((0+T[a0])+T[5+a1])+T[11+a2]
    DevBuf w1t, b1;      // [5][100][208], [5][208]   k-major, N padded with zeros
    DevBuf w2t, b2;      // [5][200][104], [5][104]
    // CTA-pair tensor-core path (gin_tc2.cu): per layer and cluster rank, half of every weight block
    DevBuf wpack2;       // [5][2][gin_tc2_pack_bytes() / 2] bytes (weights and biases)
    DevBuf pred_w, pred_b;
    // option "fixed_point" (gin_fixed.cu): the weights as the reference's host casts them, (WT_TYPE)float = floor(x * 1024) mod 2^16
    // (GIN/src/host_load.cc:60-97).  Matrices are held k-major as raw << 6 (the mad.hi operand format), the rest raw.
    DevBuf fx_ne;        // int16 [173][100]
    DevBuf fx_ee;        // int16 [5][60][100]  sum of the three tables of a bond triple, mod 2^16
    DevBuf fx_w1, fx_b1; // int32 [5][100][200], [5][200]
    DevBuf fx_w2, fx_b2; // int32 [5][200][100], [5][100]
    DevBuf fx_pw, fx_pb; // int32 [100], [1]
};
struct GcnWeights {
    DevBuf ne_table, ee_comb;   // as GIN
    DevBuf wt, b;               // [5][100][104], [5][104]
    DevBuf wpack_tc;            // [5][2][28672] bytes: W_l as bf16 hi | lo K-chunks for tcg::gemm_kernel (gcn_tc.cu)
    DevBuf root;                // [5][100]
    DevBuf bn_mean, bn_sqrt_var, bn_weight, bn_bias;   // [5][100]; sqrt_var = sqrt(var + 2^-10)
    DevBuf pred_w, pred_b;
};
struct PnaWeights {
    DevBuf ne_table;            // [173][80]
    DevBuf ne_table4;           // [431][80] combined tables of embed4_kernel (layers.cuh)
    DevBuf wcat;                // [4][320][240]  k = aggr*80+in, n = scaler*80+out
    DevBuf w_ref;               // [4][80][3][4][80] reference layout (exact path for out-degree-0 nodes)
    DevBuf wpack_tc;            // [4][5][61440] bytes: wcat as bf16 hi | lo K-chunks for pna_gemm_kernel (pna_tc.cu)
    DevBuf wpack_fused;         // [4][5][61440] bytes: the same with K permuted (16 columns x 4 aggregates per chunk) for pna_fused.cu
    DevBuf b;                   // [4][80]
    DevBuf m1w, m1b, m2w, m2b, m3w, m3b;
    float avg_deg = 0.f;
};
struct DgnWeights {
    DevBuf emb;                 // [9][119][100]
    DevBuf wt;                  // [4][200][104]  k = part*100+in
    DevBuf w_ref;               // [4][100][200] reference layout (exact path for out-degree-0 nodes)
    DevBuf wpack_tc;            // [4][4][28672] bytes: W_l as bf16 hi | lo K-chunks for tcg::gemm_kernel (dgn_tc.cu)
    DevBuf wpack_fused;         // the same with K interleaved per 32 columns ([a1 | a2] per chunk) for the fused layer kernel (dgn_tc.cu)
    DevBuf b;                   // [4][104]
    DevBuf m0w, m0b, m1w, m1b, m2w, m2b;
    // option "fixed_point" (dgn_fixed.cu): ap_fixed<16,3> bit patterns, matrices k-major as raw << 3, biases raw
    DevBuf fx_emb;              // int16 [9][119][100]
    DevBuf fx_w, fx_b;          // int32 [4][100 k][2][100 out], [4][100]
    DevBuf fx_m0w, fx_m0b, fx_m1w, fx_m1b, fx_m2w, fx_m2b;   // int32 [100][50], [50], [50][25], [25], [25], [1]
};
struct GatWeights {
    DevBuf proj0;               // [9][64]  layer-0 projection of the raw features: [f][d*4+h]
    DevBuf projt;               // [5][64][64]  k = di*4+hi, n = do*4+ho  (layer 0 unused)
    DevBuf skipt;               // [5][64][64]
    DevBuf a_src, a_tgt;        // [5][64]  index d*4+h
    DevBuf pred_w, pred_b;      // [16], [1]
    DevBuf wpack_tc;            // [5][32768] bytes: [W_proj_l ; W_skip_l] as one [128 x 64] bf16 hi | lo block per layer (gat_tc.cu; layer 0 unused)
};

// Optional per-layer device timing: one event before each layer launch and one after the last.
struct LayerTimer {
    static constexpr int MAX_MARKS = 16;
    cudaEvent_t ev[MAX_MARKS] = {};
    int created = 0;
    int marks = 0;
    int mark(cudaStream_t s);
};

struct RunOptions {
    int mp_only = 0;                 // GIN: node transform = identity (roofline variant): 1 = the mp_only mode of the layer
                                     // kernel itself (gin_fused.cu), 2 = the stand-alone row-per-warp gather kernel (gin.cu)
    int gin_ffma = 0;                // GIN: node MLP on the FP32 FFMA pipe (on-device fp32 reference) instead of tcgen05
    int gin_tc2 = 0;                 // GIN: the round-1 CTA-pair kernel (gin_tc2.cu: gather through L1 from global memory) instead of gin_fused.cu
    int gin_unfused_head = 0;        // GIN pair kernel: store h' of the last layer and run pool_head_kernel instead of the fused head
    int gin_staged = -1;             // GIN: layer = staged shared-memory gather + node MLP launch (-1: when the average in-degree is >= 6)
    int gcn_tc = 1;                  // GCN: Linear_l on tcgen05 (gcn_tc.cu: aggregate -> bf16x3 GEMM); 0: the fused FFMA kernel (gcn.cu)
    int gcn_fused = 1;               // GCN: ONE launch per step (fused_tc.cuh: the aggregation is the A producer inside the GEMM kernel); 0: aggregate + GEMM launches
    int dgn_fused = 1;               // DGN: the same
    int dgn_tc = 1;                  // DGN: node transform on tcgen05 (dgn_tc.cu: aggregate -> bf16x3 GEMM -> fp32 rows); 0: FFMA kernel (dgn.cu)
    int pna_fused = 1;               // PNA: ONE kernel per layer (pna_fused.cu: the aggregation is the A producer inside the GEMM kernel); 0: pna_tc / FFMA
    int pna_tc = 1;                  // PNA: node transform on tcgen05 (pna_tc.cu: aggregate -> bf16x3 GEMM -> exact rows); 0: FFMA kernel (pna.cu)
    int gat_tc = 1;                  // GAT: the two dense maps of a layer as ONE tcgen05 GEMM inside a fused gather kernel (gat_tc.cu); 0: the FP32 kernel (gat.cu)
    int gat_node_offset_bug = 1;
    int fixed_point = 0;             // GIN, DGN: the reference's ap_fixed<16,6> / <16,3> arithmetic, bit for bit (gin_fixed.cu, dgn_fixed.cu)
    // The input embedding needs only node_feature, the CSR / tile build only the edge lists: the embedding kernel runs on `aux`
    // between `ev_fork` (recorded on the compute stream before the build is launched) and `ev_join` (awaited before layer 0)
    int pack_graphs = 1;             // GIN, PNA: graphs re-ordered inside windows of 256 so that whole graphs fill the 128-row tiles (api.cu::pack_graphs)
    int host_stage = -1;             // host-pointer entry points: inputs narrowed to u8 / u16 by a host thread pool into pinned memory (host_stage.h);
                                     // mask of arrays (1 node_feature, 2 edge_list, 4 edge_attr), 0 off, -1: all when the caller's arrays are
                                     // pageable, or page-locked with >= 8 host threads to spare for this GPU
    int embed_overlap = 1;           // GIN, PNA, DGN: the embedding launch on a second stream, concurrent with the CSR / tile build
    cudaStream_t aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    LayerTimer* timer = nullptr;     // set while option "time_layers" is on
    int timer_group = 0;             // time_layers == 2: one interval around ALL layer launches (events between the launches
                                     // would keep them from overlapping through programmatic dependent launch)
};

// stream for the embedding launch (aux after ev_fork, or the compute stream itself) / make the compute stream wait for it
inline cudaStream_t embed_stream(const RunOptions& opt, cudaStream_t s)
{
    if (!opt.aux) return s;
    if (cudaStreamWaitEvent(opt.aux, opt.ev_fork, 0) != cudaSuccess) return s;
    return opt.aux;
}
inline int embed_join(const RunOptions& opt, cudaStream_t es, cudaStream_t s)
{
    if (es == s) return 0;
    FG_CUDA(cudaEventRecord(opt.ev_join, es));
    FG_CUDA(cudaStreamWaitEvent(s, opt.ev_join, 0));
    return 0;
}
int gin_forward(DeviceBatch& b, const GinWeights& w, const RunOptions& opt, int sm_count, cudaStream_t s, int* launches);
int dgn_fixed_forward(DeviceBatch& b, const DgnWeights& w, int sm_count, cudaStream_t s, int* launches);
int gin_fixed_forward(DeviceBatch& b, const GinWeights& w, int sm_count, cudaStream_t s, int* launches);
int gin_layer_tc2_launch(const DeviceBatch& b, const GinWeights& w, int layer, const float* h_in, float* h_out, int sm_count, cudaStream_t s,
                         const float* head_w = nullptr, float* node_dot = nullptr, const int4* row_desc = nullptr);
int gin_layer_fused_launch(const DeviceBatch& b, const GinWeights& w, int layer, const float* h_in, float* h_out, int sm_count, cudaStream_t s,
                           const float* head_w = nullptr, float* node_dot = nullptr, bool mlp_only = false, int mp_only = 0);
int gin_pool_dot_launch(const float* node_dot, const DeviceBatch& b, const float* pred_b, cudaStream_t s);
size_t gin_tc2_pack_bytes();
void gin_tc2_pack_layer(const float* w1, const float* b1, const float* w2, const float* b2, unsigned char* dst, uint16_t (*bf16_rn)(float),
                        float (*bf16_to_float)(uint16_t));
int pna_layer_fused_launch(DeviceBatch& b, const PnaWeights& w, int layer, const float* h_in, float* h_out, int sm_count, cudaStream_t s);
int pna_exact_rows_launch(DeviceBatch& b, const PnaWeights& w, int layer, const float* h_in, float* h_out, int sm_count, cudaStream_t s);
size_t pna_fused_pack_bytes();
void pna_fused_pack_layer(const float* wcat, unsigned char* dst, uint16_t (*bf16_rn)(float), float (*bf16_to_float)(uint16_t));
int pna_layer_tc_launch(DeviceBatch& b, const PnaWeights& w, int layer, const float* h_in, float* h_out, int sm_count, cudaStream_t s);
size_t pna_tc_pack_bytes();
void pna_tc_pack_layer(const float* wcat, unsigned char* dst, uint16_t (*bf16_rn)(float), float (*bf16_to_float)(uint16_t));
int gcn_step_fused_launch(DeviceBatch& b, const GcnWeights& w, int l, const float* p_in, float* p_out, int sm_count, cudaStream_t s);
int gcn_step_tc_launch(DeviceBatch& b, const GcnWeights& w, int l, const float* p_in, float* p_out, int sm_count, cudaStream_t s);
size_t gcn_tc_pack_bytes();
void gcn_tc_pack_layer(const float* w, unsigned char* dst, uint16_t (*bf16_rn)(float), float (*bf16_to_float)(uint16_t));
int dgn_layer_fused_launch(DeviceBatch& b, const DgnWeights& w, int l, const float* h_in, float* h_out, int sm_count, cudaStream_t s);
void dgn_fused_pack_layer(const float* w, unsigned char* dst, uint16_t (*bf16_rn)(float), float (*bf16_to_float)(uint16_t));
int dgn_layer_tc_launch(DeviceBatch& b, const DgnWeights& w, int l, const float* h_in, float* h_out, int sm_count, cudaStream_t s);
size_t dgn_tc_pack_bytes();
void dgn_tc_pack_layer(const float* w, unsigned char* dst, uint16_t (*bf16_rn)(float), float (*bf16_to_float)(uint16_t));
size_t gat_tc_pack_bytes();
void gat_tc_pack_layer(const float* projt, const float* skipt, unsigned char* dst, uint16_t (*bf16_rn)(float), float (*bf16_to_float)(uint16_t));
int gat_layer_tc_launch(const DeviceBatch& b, const GatWeights& w, int l, const float* hproj, const float* skip, const float* S, const float* T,
                        float* hproj_out, float* skip_out, float* S_out, float* T_out, int sm_count, cudaStream_t s);
int gat_final_launch(const DeviceBatch& b, const float* hproj, const float* skip, const float* S, const float* T, float* emb, int sm_count, cudaStream_t s);
int gcn_forward(DeviceBatch& b, const GcnWeights& w, const RunOptions& opt, int sm_count, cudaStream_t s, int* launches);
int pna_forward(DeviceBatch& b, const PnaWeights& w, const RunOptions& opt, int sm_count, cudaStream_t s, int* launches);
int dgn_forward(DeviceBatch& b, const DgnWeights& w, const RunOptions& opt, int sm_count, cudaStream_t s, int* launches);
int gat_forward(DeviceBatch& b, const GatWeights& w, const RunOptions& opt, int sm_count, cudaStream_t s, int* launches);

}  // namespace fg

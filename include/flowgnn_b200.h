/*
 * flowgnn_b200 -- C ABI of the B200 (sm_100a) replacement for FlowGNN's kernel path.
 *
 * PART 1 is the drop-in boundary: one `extern "C"` symbol per model with exactly the argument list,
 * order and buffer layouts of the reference's HLS kernel tops, which its host binds by name
 * (`cl::Kernel(program, "GIN_compute_graphs")`, GIN/src/host.cc:48) and feeds with `setArg` in
 * declaration order (GIN/src/host.cc:184-200).  Differences, all forced by the platform:
 *   - FM_TYPE / WT_TYPE are fp32 (reference: ap_fixed<16,6>, DGN ap_fixed<16,3>);
 *   - the functions return int (0 = ok, otherwise a CUDA error code or FLOWGNN_ERR_*); the
 *     reference returns void and cannot fail;
 *   - all pointers are HOST pointers owned by the caller; the library copies to the GPU, runs, and
 *     copies `out` back before returning (the reference does the same through XRT buffers).
 * Batch layout (GIN/src/host.cc:110-182): graphs are concatenated; `edge_list_in` holds (u, v) pairs
 * with GRAPH-LOCAL node ids, u = source, v = destination; weight arrays carry a leading
 * "weight set" dimension that advances each time `reload_weights[g]` is non-zero
 * (GIN/src/GIN_compute.cc:49-63).
 *
 * PART 2 is the extended interface for callers that keep batches resident in HBM, want device
 * timing, or shard across GPUs (one context per GPU).
 *
 * Threading: one context per GPU; a context is used by one host thread at a time (the reference
 * kernel is single-instance and non-reentrant: its state lives in <MODEL>/src/globals.cc).
 * Part-1 functions use a lazily created per-thread default context on the current CUDA device.
 */
#ifndef FLOWGNN_B200_H
#define FLOWGNN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLOWGNN_ERR_INVALID 10001   /* bad argument */
#define FLOWGNN_ERR_LIMIT   10002   /* the batch exceeds a documented limit (32-bit node / edge positions per batch; GIN: an edge may
                                       span at most 32,767 node positions).  Graphs of any node count are accepted: up to 1,024
                                       nodes the CSR build keeps its tables in shared memory, above that in global memory */
#define FLOWGNN_ERR_STATE   10003   /* wrong call order (no weights / no batch) */

/* ------------------------------------------------------------------------------------------------
 * PART 1 -- reference-compatible entry points
 * ---------------------------------------------------------------------------------------------- */

/* Replaces GIN_compute_graphs, GIN/src/dcl.h:76-93 (also used for GIN-VN: the virtual node is a host-side
 * graph augmentation, GIN-VN/src/host_load.cc:125-153). */
int GIN_compute_graphs(
    int num_graphs, int* nums_of_nodes, int* nums_of_edges, int* reload_weights,
    float* out,                              /* [num_graphs][NUM_TASK = 1] */
    const int32_t* node_feature_in,          /* [sum N][9] */
    const int32_t* edge_list_in,             /* [sum E][2] */
    const int32_t* edge_attr_in,             /* [sum E][3] */
    const float* node_embedding_weight_in,   /* [][173][100] */
    const float* edge_embedding_weight_in,   /* [][5][13][100] */
    const float* node_mlp_1_weights,         /* [][5][200][100] */
    const float* node_mlp_1_bias,            /* [][5][200] */
    const float* node_mlp_2_weights,         /* [][5][100][200] */
    const float* node_mlp_2_bias,            /* [][5][100] */
    const float* graph_pred_weights_in,      /* [][1][100] */
    const float* graph_pred_bias_in);        /* [][1] */

/* The same entry point with the FPGA build's own scalar type: FM_TYPE = WT_TYPE = ap_fixed<16,6> (GIN/src/dcl.h:58-59), i.e.
 * every weight and every result is an int16 bit pattern, value = raw / 1024 -- what the reference's host produces with
 * `(WT_TYPE)float` (GIN/src/host_load.cc:60-97) and reads back from `result` (GIN/src/host.cc:160-166).  The arithmetic
 * is Vitis' bit for bit: exact wide intermediates, floor to 10 fraction bits on every assignment (AP_TRN), wrap to 16 bits
 * (AP_WRAP), integer division toward zero in the mean pool.  Also reachable through the float
 * entry points and Part 2 with flowgnn_b200_set_option("fixed_point", 1): fp32 weights are then cast as the reference's
 * host casts them and `out` holds raw / 1024. */
int GIN_compute_graphs_fixed(
    int num_graphs, int* nums_of_nodes, int* nums_of_edges, int* reload_weights,
    int16_t* out,                            /* [num_graphs][1] raw */
    const int32_t* node_feature_in, const int32_t* edge_list_in, const int32_t* edge_attr_in,
    const int16_t* node_embedding_weight_in, const int16_t* edge_embedding_weight_in,
    const int16_t* node_mlp_1_weights, const int16_t* node_mlp_1_bias,
    const int16_t* node_mlp_2_weights, const int16_t* node_mlp_2_bias,
    const int16_t* graph_pred_weights_in, const int16_t* graph_pred_bias_in);

/* Replaces GCN_compute_graphs, GCN/src/dcl.h:76-96. */
int GCN_compute_graphs(
    int num_graphs, int* nums_of_nodes, int* nums_of_edges, int* reload_weights,
    float* out,
    const int32_t* node_feature_in, const int32_t* edge_list_in, const int32_t* edge_attr_in,
    const float* node_embedding_weight_in,   /* [][173][100] */
    const float* edge_embedding_weight_in,   /* [][5][13][100] */
    const float* convs_weight_in,            /* [][5][100][100] */
    const float* convs_bias_in,              /* [][5][100] */
    const float* convs_root_emb_weight_in,   /* [][5][100] */
    const float* bn_weight_in, const float* bn_bias_in, const float* bn_mean_in, const float* bn_var_in, /* [][5][100] */
    const float* graph_pred_weights_in, const float* graph_pred_bias_in);

/* Replaces GAT_compute_graphs, GAT/src/dcl.h:79-93.  By default reproduces the reference's missing
 * per-graph node-feature offset (GAT/src/GAT_compute.cc:72); see
 * flowgnn_b200_set_option("gat_node_offset_bug"). */
int GAT_compute_graphs(
    int num_graphs, int* nums_of_nodes, int* nums_of_edges, int* reload_weights,
    float* out,
    const int32_t* node_feature_in, const int32_t* edge_list_in,
    const float* scoring_fn_target_in,       /* [][5][4][16] */
    const float* scoring_fn_source_in,       /* [][5][4][16] */
    const float* linear_proj_weights_in,     /* [][5][4][16][4][16] */
    const float* skip_proj_weights_in,       /* [][5][4][16][4][16] */
    const float* graph_pred_weights_in,      /* [][1][16] */
    const float* graph_pred_bias_in);

/* Replaces PNA_compute_graphs, PNA/src/dcl.h:92-110. */
int PNA_compute_graphs(
    int num_graphs, int* nums_of_nodes, int* nums_of_edges, int* reload_weights,
    float* out,
    const int32_t* node_feature_in, const int32_t* edge_list_in,
    const float* node_embedding_weight_in,   /* [][173][80] */
    const float* node_conv_weights_in,       /* [][4][80][3 scalers][4 aggregators][80] */
    const float* node_conv_bias_in,          /* [][4][80] */
    const float* graph_mlp_1_weights_in, const float* graph_mlp_1_bias_in,   /* [][40][80], [][40] */
    const float* graph_mlp_2_weights_in, const float* graph_mlp_2_bias_in,   /* [][20][40], [][20] */
    const float* graph_mlp_3_weights_in, const float* graph_mlp_3_bias_in,   /* [][1][20],  [][1]  */
    const float* avg_deg_in);                /* [] */

/* Replaces DGN_compute_graphs, DGN/src/dcl.h:72-90. */
int DGN_compute_graphs(
    int num_graphs, int* nums_of_nodes, int* nums_of_edges, int* reload_weights,
    float* out,
    const int32_t* node_feature_in,
    const float* node_eigen_in,              /* [sum N][4]; only column 1 is used (DGN/src/load_inputs.cc:105-106) */
    const int32_t* edge_list_in,
    const float* embedding_h_atom_embedding_list_weights_in,            /* [][9][119][100] */
    const float* layers_posttrans_fully_connected_0_linear_weight_in,   /* [][4][100][200] */
    const float* layers_posttrans_fully_connected_0_linear_bias_in,     /* [][4][100] */
    const float* MLP_layer_FC_layers_0_weight_in, const float* MLP_layer_FC_layers_0_bias_in,   /* [][50][100], [][50] */
    const float* MLP_layer_FC_layers_1_weight_in, const float* MLP_layer_FC_layers_1_bias_in,   /* [][25][50],  [][25] */
    const float* MLP_layer_FC_layers_2_weight_in, const float* MLP_layer_FC_layers_2_bias_in);  /* [][1][25],   [][1]  */

/* ------------------------------------------------------------------------------------------------
 * PART 2 -- extended interface (not in the reference)
 * ---------------------------------------------------------------------------------------------- */

typedef struct flowgnn_ctx flowgnn_ctx;

enum flowgnn_model { FLOWGNN_GIN = 0, FLOWGNN_GCN = 1, FLOWGNN_GAT = 2, FLOWGNN_PNA = 3, FLOWGNN_DGN = 4 };

/* Text of the last error raised on the calling thread ("" if none). */
const char* flowgnn_b200_last_error(void);

int flowgnn_b200_create(flowgnn_ctx** ctx, int device);
int flowgnn_b200_destroy(flowgnn_ctx* ctx);

/* Options:
 *   "mp_only"      GIN: node transform = identity (the edge gather-scatter roofline variant).  1 = the mp_only
 *                  mode of the layer kernel itself (gin_fused.cu), 2 = the stand-alone row-per-warp / staged gather kernels
 *   "gin_ffma"     GIN: node MLP on the FP32 FFMA pipe (the on-device fp32 reference) instead of tcgen05 bf16x3
 *   "gin_tc2"      GIN: the round-1 CTA-pair kernel (gather from global memory through L1) instead of gin_fused.cu (graph-
 *                  aligned tiles staged by TMA, shared-memory gather); both add in-edges in CSR order: bit-identical results
 *   "gin_staged"   GIN: a layer as two launches -- the shared-memory-staged gather and the node MLP; -1 (default): when the
 *                  batch averages >= 6 in-edges per node (hep10k kNN graphs), 0 / 1: never / always
 *   "gin_unfused_head"  GIN: store h' of the last layer and run the pooling kernel instead of the head fused in the epilogue
 *   "pna_fused"    PNA, default 1: ONE kernel per layer (pna_fused.cu: the aggregation is the A producer inside the tcgen05
 *                  GEMM kernel); 0: "pna_tc" decides
 *   "pna_tc"       PNA, default 1: aggregate kernel -> bf16x3 GEMM -> fp32 rows kernel; 0: the fused FP32 FFMA kernel, kept
 *                  as the on-device fp32 reference
 *   "gcn_tc" / "dgn_tc"  default 1: the dense layer of GCN / DGN on tcgen05 (aggregate -> GEMM, tcgemm.cuh); 0: the fused
 *                  FFMA kernels; environment FLOWGNN_B200_TC_ALL=0/1 sets both defaults
 *   "gcn_fused", "dgn_fused"  default 1: GCN step / DGN layer as ONE launch (fused_tc.cuh: the aggregation is the A producer inside the
 *                          tcgen05 GEMM kernel); 0: round 1's aggregate + GEMM launches (needs "gcn_tc" / "dgn_tc" = 1)
 *   "host_stage"   default -1 (automatic): the Part-1 entry points narrow node_feature / edge_attr to one byte and edge_list to two bytes
 *                  per value with a pool of host threads (FLOWGNN_B200_HOST_THREADS, default min(12, 3/4 of the usable cores /
 *                  LOCAL_WORLD_SIZE)) into pinned memory, copy 9 B per node + 7 B per edge instead of 36 + 20, and widen them on the device;
 *                  the caller's arrays may be pageable.  Value = mask of the arrays to narrow (1 node_feature, 2 edge_list, 4 edge_attr),
 *                  0 = off; automatic = all three for pageable arrays, and for page-locked ones when 8 threads are available.  An array
 *                  holding a value that does not fit is uploaded unchanged.  Results are bit-identical.  (FLOWGNN_B200_HOST_STAGE overrides.)
 *   "pack_graphs"  default 1: GIN / PNA store the graphs' rows in an order (windows of 256, best fit) that fills the 128-row tiles of the
 *                  layer kernels to ~98 %; inputs are read and `out` is written in the caller's order either way
 *   "gat_tc"       default 1: GAT's two dense maps per layer as ONE tcgen05 GEMM inside a fused gather kernel (gat_tc.cu); 0: the FP32 kernel
 *   "embed_overlap" default 1: GIN / PNA / DGN run the input embedding on a second stream, concurrent with the CSR / tile build
 *   "gat_node_offset_bug"  default 1
 *   "fixed_point"          default 0; 1: GIN / GIN-VN in the reference's ap_fixed<16,6> arithmetic, bit for bit (gin_fixed.cu);
 *                          other models return FLOWGNN_ERR_INVALID while it is set
 *   "time_layers"  1: see flowgnn_b200_last_layer_ms; 2 (GIN): ONE interval around all layer launches, which leaves them
 *                  adjacent in the stream so that programmatic dependent launch can overlap them
 * Environment: FLOWGNN_B200_CHUNKS=n overrides the number of chunks the host-pointer entry points cut a batch into
 * (default 3 for >= 16,384 graphs: upload of chunk i+1 overlaps the kernels of chunk i). */
int flowgnn_b200_set_option(flowgnn_ctx* ctx, const char* name, int value);

/* load_weights (GIN/src/load_inputs.cc:7-85 and per-model variants): upload ONE weight set and
 * repack it for the kernels.  `weights` lists host pointers in the model's Part-1 argument order
 * (GIN 8, GCN 11, GAT 6, PNA 10, DGN 9). */
int flowgnn_b200_load_weights(flowgnn_ctx* ctx, int model, const float* const* weights, int num_weights);

/* Copy a batch to the GPU (asynchronously on the context's stream; pinned host memory overlaps). */
int flowgnn_b200_upload_batch(
    flowgnn_ctx* ctx, int num_graphs, int64_t total_nodes, int64_t total_edges,
    const int32_t* nums_of_nodes, const int32_t* nums_of_edges,
    const int32_t* node_feature, const int32_t* edge_list,
    const int32_t* edge_attr /* may be NULL */, const float* node_eigen /* may be NULL */);

/* The same batch with the big arrays in the narrow layout of the packed dataset files (flowgnn_b200/dataset.py::save_packed):
 * node_feature as uint8 [N][9], edge_list as uint16 [E][2] (graph-local ids), edge_attr as uint8 [E][3] (or NULL), 9 B per node +
 * 7 B per edge over PCIe instead of the reference ABI's 36 + 20 (GIN/src/dcl.h:61-67).  Widened on the device into the int32
 * arrays the kernels read; everything else as flowgnn_b200_upload_batch.  A dataset whose values do not fit uses that call. */
int flowgnn_b200_upload_batch_packed(flowgnn_ctx* ctx, int num_graphs, int64_t total_nodes, int64_t total_edges,
                                     const int32_t* nums_of_nodes, const int32_t* nums_of_edges, const uint8_t* node_feature,
                                     const uint16_t* edge_list, const uint8_t* edge_attr, const float* node_eigen);

/* Run graph preprocessing (load_graph) + the full forward of `model` on the resident batch.
 * `elapsed_ms` (may be NULL) receives the device time between CUDA events around exactly that work;
 * passing it makes the call synchronous. */
int flowgnn_b200_compute(flowgnn_ctx* ctx, int model, float* elapsed_ms);

/* Copy the per-graph predictions of the last compute to host memory and synchronise. */
int flowgnn_b200_download(flowgnn_ctx* ctx, float* out, int num_graphs);

/* Kernels launched by the last flowgnn_b200_compute. */
int flowgnn_b200_last_launch_count(flowgnn_ctx* ctx);

/* Number of 128-row tiles of whole graphs the uploaded batch was packed into (GIN / PNA layer launches work on these; a tile
 * issues 128 rows of tensor-core work whatever its fill).  0 if the batch was not packed on the host. */
long flowgnn_b200_tile_count(flowgnn_ctx* ctx);

/* With option "time_layers" = 1, flowgnn_b200_compute brackets every per-layer kernel launch with CUDA
 * events; this returns the device time of each layer launch of the last compute (count written). */
int flowgnn_b200_last_layer_ms(flowgnn_ctx* ctx, float* out, int max_layers);

/* The context's cudaStream_t (as void*), for callers that want to order their own work. */
void* flowgnn_b200_stream(flowgnn_ctx* ctx);

int flowgnn_b200_synchronize(flowgnn_ctx* ctx);

/* Page-lock / release caller-owned host memory (cudaHostRegister without CUDA headers).  The reference host keeps its
 * batch in 4 KiB-aligned pageable vectors (common/includes/xcl2/xcl2.hpp:61-76) that XRT maps with CL_MEM_USE_HOST_PTR
 * (GIN/src/host.cc:141-182); pinning them once after the batch is built lets the Part-1 entry points overlap the upload
 * of chunk i+1 with the kernels of chunk i.  Pageable buffers work too, the copies are then staged by the driver. */
int flowgnn_b200_pin_host(void* ptr, size_t bytes);
int flowgnn_b200_unpin_host(void* ptr);

/* The host-side step of the narrowed upload of the Part-1 entry points (option "host_stage"): dst[i] = the low `width` bytes
 * (1 or 2) of src[i], computed by `threads` host threads (>= 1; the caller is one of them).  Returns the OR of all source
 * words: bits above the narrow width mean a value did not fit, and the entry points then upload that array unchanged.
 * Exported for tests and for callers that want to measure their host; no GPU is touched.  width outside {1, 2}: returns
 * 0xFFFFFFFF and writes nothing. */
uint32_t flowgnn_b200_narrow_words(const int32_t* src, size_t n, int width, void* dst, int threads);

/* <MODEL>_compute_graphs for a caller that keeps its dataset in the packed layout (flowgnn_b200/dataset.py::save_packed): uint8
 * node features [N][9], uint16 graph-local edge ids [E][2], uint8 bond attributes [E][3] (GIN / GCN; NULL otherwise), fp32
 * node_eigen [N][4] (DGN; NULL otherwise).  Same pipeline as the entry points of Part 1 -- chunks of whole graphs alternate between
 * two device batches, the upload of chunk i+1 overlaps the kernels of chunk i, predictions are written to `out` before the call
 * returns -- with 9 B per node + 7 B per edge crossing PCIe and no host-side narrowing.  model: enum flowgnn_model; weights: the
 * model's arrays in the order of its <MODEL>_compute_graphs argument list (one weight set). */
int flowgnn_b200_compute_graphs_packed(int model, int num_graphs, const int32_t* nums_of_nodes, const int32_t* nums_of_edges, float* out,
                                       const uint8_t* node_feature, const uint16_t* edge_list, const uint8_t* edge_attr,
                                       const float* node_eigen, const float* const* weights, int num_weights);

/* Bytes the calling thread's last <MODEL>_compute_graphs call copied host -> device and device -> host (batch inputs, offsets and
 * tile lists, predictions and status words; weights are uploaded only when their contents change and are not counted). */
void flowgnn_b200_last_transfer_bytes(uint64_t* h2d, uint64_t* d2h);

#ifdef __cplusplus
}
#endif
#endif

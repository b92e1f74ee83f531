/* TEST INFRASTRUCTURE (oracle).  Plain-C restatement of the six FlowGNN kernels in fp32.
 *
 * This is the CHECKER, never the product: only tests/, __graft_entry__.smoke() and bench.py's
 * CPU-baseline legs may link or load it.  The entry points take exactly the argument lists of
 * the reference's kernel tops (float for ap_fixed, int32 ids):
 *   GIN/src/dcl.h:76-93, GCN/src/dcl.h:76-96, GAT/src/dcl.h:79-93, PNA/src/dcl.h:92-110,
 *   DGN/src/dcl.h:72-90.
 * Unlike the reference (state in file-scope globals, <MODEL>/src/globals.cc) every function here is
 * re-entrant, so shards can run on several host threads.
 *
 * Parity pin: built with -O2 -ffp-contract=off this restatement is BIT-IDENTICAL to the
 * reference's own sources compiled against oracle/shim (oracle/_ref) on every shipped molhiv
 * graph for all six models -- tests/test_oracle.py checks that against tests/golden/.
 */
#ifndef FLOWGNN_ORACLE_H
#define FLOWGNN_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

void oracle_GIN_compute_graphs(
    int num_graphs, const int32_t* nums_of_nodes, const int32_t* nums_of_edges, const int32_t* reload_weights,
    float* out, const int32_t* node_feature_in, const int32_t* edge_list_in, const int32_t* edge_attr_in,
    const float* node_embedding_weight_in, const float* edge_embedding_weight_in,
    const float* node_mlp_1_weights, const float* node_mlp_1_bias,
    const float* node_mlp_2_weights, const float* node_mlp_2_bias,
    const float* graph_pred_weights_in, const float* graph_pred_bias_in);

void oracle_GCN_compute_graphs(
    int num_graphs, const int32_t* nums_of_nodes, const int32_t* nums_of_edges, const int32_t* reload_weights,
    float* out, const int32_t* node_feature_in, const int32_t* edge_list_in, const int32_t* edge_attr_in,
    const float* node_embedding_weight_in, const float* edge_embedding_weight_in,
    const float* convs_weight_in, const float* convs_bias_in, const float* convs_root_emb_weight_in,
    const float* bn_weight_in, const float* bn_bias_in, const float* bn_mean_in, const float* bn_var_in,
    const float* graph_pred_weights_in, const float* graph_pred_bias_in);

void oracle_GAT_compute_graphs(
    int num_graphs, const int32_t* nums_of_nodes, const int32_t* nums_of_edges, const int32_t* reload_weights,
    float* out, const int32_t* node_feature_in, const int32_t* edge_list_in,
    const float* scoring_fn_target_in, const float* scoring_fn_source_in,
    const float* linear_proj_weights_in, const float* skip_proj_weights_in,
    const float* graph_pred_weights_in, const float* graph_pred_bias_in);

void oracle_PNA_compute_graphs(
    int num_graphs, const int32_t* nums_of_nodes, const int32_t* nums_of_edges, const int32_t* reload_weights,
    float* out, const int32_t* node_feature_in, const int32_t* edge_list_in,
    const float* node_embedding_weight_in, const float* node_conv_weights_in, const float* node_conv_bias_in,
    const float* graph_mlp_1_weights_in, const float* graph_mlp_1_bias_in,
    const float* graph_mlp_2_weights_in, const float* graph_mlp_2_bias_in,
    const float* graph_mlp_3_weights_in, const float* graph_mlp_3_bias_in,
    const float* avg_deg_in);

void oracle_DGN_compute_graphs(
    int num_graphs, const int32_t* nums_of_nodes, const int32_t* nums_of_edges, const int32_t* reload_weights,
    float* out, const int32_t* node_feature_in, const float* node_eigen_in, const int32_t* edge_list_in,
    const float* embedding_h_atom_embedding_list_weights_in,
    const float* layers_posttrans_fully_connected_0_linear_weight_in,
    const float* layers_posttrans_fully_connected_0_linear_bias_in,
    const float* MLP_layer_FC_layers_0_weight_in, const float* MLP_layer_FC_layers_0_bias_in,
    const float* MLP_layer_FC_layers_1_weight_in, const float* MLP_layer_FC_layers_1_bias_in,
    const float* MLP_layer_FC_layers_2_weight_in, const float* MLP_layer_FC_layers_2_bias_in);

/* GAT only: 1 (default) reproduces the reference's missing per-graph node-feature offset
 * (GAT/src/GAT_compute.cc:72); 0 reads each graph's own features. */
void oracle_set_gat_node_offset_bug(int enabled);

#ifdef __cplusplus
}
#endif
#endif

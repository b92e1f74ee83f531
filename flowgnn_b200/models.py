"""Static description of the six FlowGNN kernels' weight arguments.

Every entry mirrors the argument list of the reference's kernel top function,
in declaration order, with the leading "weight-set" dimension dropped (the
reference indexes it with ``weights_ndx``; GIN/src/GIN_compute.cc:49-63):

* GIN / GIN-VN  -- GIN/src/dcl.h:76-93   (16 args, 8 of them weights)
* GCN           -- GCN/src/dcl.h:76-96   (19 args, 11 weights)
* GAT           -- GAT/src/dcl.h:79-93   (13 args, 6 weights)
* PNA           -- PNA/src/dcl.h:92-110  (17 args, 10 weights incl. avg_deg)
* DGN           -- DGN/src/dcl.h:72-90   (17 args, 9 weights)

``FM_TYPE``/``WT_TYPE`` are fp32 here (the fp32 flavour is the
parity target; the fixed-point flavour is a "next" row).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple

ND_FEATURE = 9
ND_FEATURE_TOTAL = 173
EDGE_ATTR = 3
ED_FEATURE_PER_LAYER = 13
NUM_TASK = 1

#: vocabulary sizes of the 9 atom features / 3 bond features (GIN/src/host_load.cc:5-6)
ND_FEATURE_TABLE = (119, 4, 12, 12, 10, 6, 6, 2, 2)
ED_FEATURE_TABLE = (5, 6, 2)
#: row offsets into the concatenated embedding tables (GIN/src/load_inputs.cc:5, message_passing.cc:3)
ND_FEATURE_OFFSETS = (0, 119, 123, 135, 147, 157, 163, 169, 171)
ED_FEATURE_OFFSETS = (0, 5, 11)

#: PNA's hard-coded average log-degree (PNA/src/host_load.cc:127)
PNA_AVG_DEG = 6.885701656341553


@dataclass(frozen=True)
class ModelSpec:
    name: str            # canonical tag: gin, ginvn, gcn, gat, pna, dgn
    symbol: str          # extern "C" entry point, as in the reference
    emb_dim: int
    num_layers: int
    uses_edge_attr: bool
    uses_eigen: bool
    virtual_node: bool
    weights: Tuple[Tuple[str, Tuple[int, ...]], ...]  # (arg name, shape without weight-set dim), kernel order

    @property
    def weight_names(self) -> List[str]:
        return [n for n, _ in self.weights]

    def weight_shapes(self) -> Dict[str, Tuple[int, ...]]:
        return dict(self.weights)


_GIN_WEIGHTS = (
    ("node_embedding_weight", (ND_FEATURE_TOTAL, 100)),
    ("edge_embedding_weight", (5, ED_FEATURE_PER_LAYER, 100)),
    ("node_mlp_1_weights", (5, 200, 100)),
    ("node_mlp_1_bias", (5, 200)),
    ("node_mlp_2_weights", (5, 100, 200)),
    ("node_mlp_2_bias", (5, 100)),
    ("graph_pred_weights", (NUM_TASK, 100)),
    ("graph_pred_bias", (NUM_TASK,)),
)

MODELS: Dict[str, ModelSpec] = {
    "gin": ModelSpec("gin", "GIN_compute_graphs", 100, 5, True, False, False, _GIN_WEIGHTS),
    # GIN-VN is the GIN kernel on host-augmented graphs (GIN-VN/src/host_load.cc:125-153).
    "ginvn": ModelSpec("ginvn", "GIN_compute_graphs", 100, 5, True, False, True, _GIN_WEIGHTS),
    "gcn": ModelSpec(
        "gcn", "GCN_compute_graphs", 100, 5, True, False, False,
        (
            ("node_embedding_weight", (ND_FEATURE_TOTAL, 100)),
            ("edge_embedding_weight", (5, ED_FEATURE_PER_LAYER, 100)),
            ("convs_weight", (5, 100, 100)),
            ("convs_bias", (5, 100)),
            ("convs_root_emb_weight", (5, 100)),
            ("bn_weight", (5, 100)),
            ("bn_bias", (5, 100)),
            ("bn_mean", (5, 100)),
            ("bn_var", (5, 100)),
            ("graph_pred_weights", (NUM_TASK, 100)),
            ("graph_pred_bias", (NUM_TASK,)),
        ),
    ),
    "gat": ModelSpec(
        "gat", "GAT_compute_graphs", 16, 5, False, False, False,
        (
            ("scoring_fn_target", (5, 4, 16)),
            ("scoring_fn_source", (5, 4, 16)),
            ("linear_proj_weights", (5, 4, 16, 4, 16)),
            ("skip_proj_weights", (5, 4, 16, 4, 16)),
            ("graph_pred_weights", (NUM_TASK, 16)),
            ("graph_pred_bias", (NUM_TASK,)),
        ),
    ),
    "pna": ModelSpec(
        "pna", "PNA_compute_graphs", 80, 4, False, False, False,
        (
            ("node_embedding_weight", (ND_FEATURE_TOTAL, 80)),
            ("node_conv_weights", (4, 80, 3, 4, 80)),   # [l][out][scaler][aggr][in]
            ("node_conv_bias", (4, 80)),
            ("graph_mlp_1_weights", (40, 80)),
            ("graph_mlp_1_bias", (40,)),
            ("graph_mlp_2_weights", (20, 40)),
            ("graph_mlp_2_bias", (20,)),
            ("graph_mlp_3_weights", (NUM_TASK, 20)),
            ("graph_mlp_3_bias", (NUM_TASK,)),
            ("avg_deg", (1,)),
        ),
    ),
    "dgn": ModelSpec(
        "dgn", "DGN_compute_graphs", 100, 4, False, True, False,
        (
            ("embedding_h_atom_embedding_list_weights", (9, 119, 100)),
            ("layers_posttrans_fully_connected_0_linear_weight", (4, 100, 200)),
            ("layers_posttrans_fully_connected_0_linear_bias", (4, 100)),
            ("MLP_layer_FC_layers_0_weight", (50, 100)),
            ("MLP_layer_FC_layers_0_bias", (50,)),
            ("MLP_layer_FC_layers_1_weight", (25, 50)),
            ("MLP_layer_FC_layers_1_bias", (25,)),
            ("MLP_layer_FC_layers_2_weight", (1, 25)),
            ("MLP_layer_FC_layers_2_bias", (1,)),
        ),
    ),
}

MODEL_NAMES = tuple(MODELS)


def get_model(name: str) -> ModelSpec:
    key = name.lower().replace("-", "").replace("_", "")
    if key not in MODELS:
        raise KeyError(f"unknown model {name!r}; expected one of {MODEL_NAMES}")
    return MODELS[key]

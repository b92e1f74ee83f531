"""Weight loading: the host-side mirror of the reference's ``load_weights()``.

The on-disk formats are the reference's own:

* GIN / GIN-VN -- ``gin_ep1_noBN_dim100.weights.all.bin`` (225,406 fp32), or the nine split
  files ``gin_ep1_*_dim100.bin`` that GIN/src/host_load.cc:24-58 reads.  ``eps`` is parsed
  but never handed to the kernel (GIN/src/host.cc:184-200 passes no eps).
* GCN -- ``gcn_ep1_dim100.weights.all.bin`` (GCN/src/host_load.cc:31-170).
* PNA -- ``pna_ep1_noBN_dim80.weights.all.bin`` (PNA/src/host_load.cc:23-68); ``avg_deg`` is
  the constant of PNA/src/host_load.cc:127.
* DGN -- ``dgn_ep1_noBN_dim100.weights.all.bin`` (DGN/src/host_load.cc:11-149); the nine
  per-feature embedding tables are placed at ``f*11900`` in a zero-filled [9][119][100] buffer.
* GAT -- eight split files ``gat_ep1_*_layer5.bin`` (GAT/src/host_load.cc:20-49), layer-0
  projections zero-padded into [5][4][16][4][16] (:69-91).

Values stay fp32 (the reference casts to ap_fixed here).
"""
from __future__ import annotations

import os
from typing import Dict

import numpy as np

from .models import (ND_FEATURE, ND_FEATURE_TABLE, PNA_AVG_DEG, ModelSpec, get_model)

Weights = Dict[str, np.ndarray]


def _read_f32(path: str, count: int = -1) -> np.ndarray:
    if not os.path.isfile(path):
        raise FileNotFoundError(f"weight file not found: {path}")
    return np.fromfile(path, dtype="<f4", count=count)


def _take(blob: np.ndarray, offset: int, shape) -> np.ndarray:
    n = int(np.prod(shape))
    if offset + n > blob.size:
        raise ValueError(f"weight blob too short: need {offset + n} floats, have {blob.size}")
    return np.ascontiguousarray(blob[offset:offset + n].reshape(shape))


def _load_gin(d: str) -> Weights:
    all_bin = os.path.join(d, "gin_ep1_noBN_dim100.weights.all.bin")
    w: Weights = {}
    if os.path.isfile(all_bin):
        blob = _read_f32(all_bin)
        if blob.size != 225406:
            raise ValueError(f"{all_bin}: expected 225406 floats, got {blob.size}")
        w["node_embedding_weight"] = _take(blob, 0, (173, 100))
        w1, b1, w2, b2, ee, eps = [], [], [], [], [], []
        for l in range(5):
            base = 17300 + 41601 * l
            eps.append(blob[base])
            w1.append(_take(blob, base + 1, (200, 100)))
            b1.append(_take(blob, base + 20001, (200,)))
            w2.append(_take(blob, base + 20201, (100, 200)))
            b2.append(_take(blob, base + 40201, (100,)))
            ee.append(_take(blob, base + 40301, (13, 100)))
        w["edge_embedding_weight"] = np.stack(ee)
        w["node_mlp_1_weights"] = np.stack(w1)
        w["node_mlp_1_bias"] = np.stack(b1)
        w["node_mlp_2_weights"] = np.stack(w2)
        w["node_mlp_2_bias"] = np.stack(b2)
        w["graph_pred_weights"] = _take(blob, 225305, (1, 100))
        w["graph_pred_bias"] = _take(blob, 225405, (1,))
        w["_eps_unused"] = np.asarray(eps, dtype=np.float32)
        return w
    p = lambda n: os.path.join(d, f"gin_ep1_{n}_dim100.bin")
    w["node_embedding_weight"] = _read_f32(p("nd_embed"), 17300).reshape(173, 100)
    w["edge_embedding_weight"] = _read_f32(p("ed_embed"), 6500).reshape(5, 13, 100)
    w["node_mlp_1_weights"] = _read_f32(p("mlp_1_weights"), 100000).reshape(5, 200, 100)
    w["node_mlp_1_bias"] = _read_f32(p("mlp_1_bias"), 1000).reshape(5, 200)
    w["node_mlp_2_weights"] = _read_f32(p("mlp_2_weights"), 100000).reshape(5, 100, 200)
    w["node_mlp_2_bias"] = _read_f32(p("mlp_2_bias"), 500).reshape(5, 100)
    w["graph_pred_weights"] = _read_f32(p("pred_weights"), 100).reshape(1, 100)
    w["graph_pred_bias"] = _read_f32(p("pred_bias"), 1).reshape(1)
    w["_eps_unused"] = _read_f32(p("eps"), 5)
    return w


def _load_gcn(d: str) -> Weights:
    blob = _read_f32(os.path.join(d, "gcn_ep1_dim100.weights.all.bin"))
    w: Weights = {"node_embedding_weight": _take(blob, 0, (173, 100))}
    cw, cb, cr, ee = [], [], [], []
    for l in range(5):
        base = 17300 + 11500 * l
        cw.append(_take(blob, base, (100, 100)))
        cb.append(_take(blob, base + 10000, (100,)))
        cr.append(_take(blob, base + 10100, (100,)))
        ee.append(_take(blob, base + 10200, (13, 100)))
    w["edge_embedding_weight"] = np.stack(ee)
    w["convs_weight"] = np.stack(cw)
    w["convs_bias"] = np.stack(cb)
    w["convs_root_emb_weight"] = np.stack(cr)
    bn = {k: [] for k in ("bn_weight", "bn_bias", "bn_mean", "bn_var")}
    for l in range(5):
        base = 74800 + 401 * l  # one scalar (num_batches_tracked) is skipped after each layer's var
        for i, k in enumerate(("bn_weight", "bn_bias", "bn_mean", "bn_var")):
            bn[k].append(_take(blob, base + 100 * i, (100,)))
    for k, v in bn.items():
        w[k] = np.stack(v)
    w["graph_pred_weights"] = _take(blob, 76805, (1, 100))
    w["graph_pred_bias"] = _take(blob, 76905, (1,))
    return w


def _load_pna(d: str) -> Weights:
    blob = _read_f32(os.path.join(d, "pna_ep1_noBN_dim80.weights.all.bin"))
    w: Weights = {"node_embedding_weight": _take(blob, 0, (173, 80))}
    cw, cb = [], []
    for l in range(4):
        base = 13840 + 76880 * l
        cw.append(_take(blob, base, (80, 3, 4, 80)))
        cb.append(_take(blob, base + 76800, (80,)))
    w["node_conv_weights"] = np.stack(cw)
    w["node_conv_bias"] = np.stack(cb)
    w["graph_mlp_1_weights"] = _take(blob, 321360, (40, 80))
    w["graph_mlp_1_bias"] = _take(blob, 324560, (40,))
    w["graph_mlp_2_weights"] = _take(blob, 324600, (20, 40))
    w["graph_mlp_2_bias"] = _take(blob, 325400, (20,))
    w["graph_mlp_3_weights"] = _take(blob, 325420, (1, 20))
    w["graph_mlp_3_bias"] = _take(blob, 325440, (1,))
    w["avg_deg"] = np.asarray([PNA_AVG_DEG], dtype=np.float32)
    return w


def _load_dgn(d: str) -> Weights:
    blob = _read_f32(os.path.join(d, "dgn_ep1_noBN_dim100.weights.all.bin"))
    emb = np.zeros((ND_FEATURE, 119, 100), dtype=np.float32)
    off = 0
    for f, rows in enumerate(ND_FEATURE_TABLE):
        emb[f, :rows] = _take(blob, off, (rows, 100))
        off += rows * 100
    w: Weights = {"embedding_h_atom_embedding_list_weights": emb}
    lw, lb = [], []
    for l in range(4):
        base = 17300 + 20100 * l
        lw.append(_take(blob, base, (100, 200)))
        lb.append(_take(blob, base + 20000, (100,)))
    w["layers_posttrans_fully_connected_0_linear_weight"] = np.stack(lw)
    w["layers_posttrans_fully_connected_0_linear_bias"] = np.stack(lb)
    w["MLP_layer_FC_layers_0_weight"] = _take(blob, 97700, (50, 100))
    w["MLP_layer_FC_layers_0_bias"] = _take(blob, 102700, (50,))
    w["MLP_layer_FC_layers_1_weight"] = _take(blob, 102750, (25, 50))
    w["MLP_layer_FC_layers_1_bias"] = _take(blob, 104000, (25,))
    w["MLP_layer_FC_layers_2_weight"] = _take(blob, 104025, (1, 25))
    w["MLP_layer_FC_layers_2_bias"] = _take(blob, 104050, (1,))
    return w


def _load_gat(d: str) -> Weights:
    p = lambda n: os.path.join(d, f"gat_ep1_{n}_layer5.bin")
    w: Weights = {}
    w["scoring_fn_target"] = _read_f32(p("scoring_fn_target"), 320).reshape(5, 4, 16)
    w["scoring_fn_source"] = _read_f32(p("scoring_fn_source"), 320).reshape(5, 4, 16)
    for kind in ("linear", "skip"):
        full = np.zeros((5, 4, 16, 4, 16), dtype=np.float32)
        l0 = _read_f32(p(f"{kind}_proj_weight_0"), 4 * 16 * 9).reshape(4, 16, 1, 9)
        full[0, :, :, 0, :9] = l0[:, :, 0, :]
        full[1:] = _read_f32(p(f"{kind}_proj_weight_1"), 4 * 4096).reshape(4, 4, 16, 4, 16)
        w[f"{kind}_proj_weights"] = full
    w["graph_pred_weights"] = _read_f32(p("pred_weights"), 16).reshape(1, 16)
    w["graph_pred_bias"] = _read_f32(p("pred_bias"), 1).reshape(1)
    return w


_LOADERS = {"gin": _load_gin, "ginvn": _load_gin, "gcn": _load_gcn, "pna": _load_pna, "dgn": _load_dgn, "gat": _load_gat}


def load_weights(model: str, directory: str) -> Weights:
    """Read one model's trained weights from ``directory`` in the reference's file formats."""
    spec = get_model(model)
    w = _LOADERS[spec.name](directory)
    return check_weights(spec, w)


def check_weights(spec: ModelSpec, w: Weights) -> Weights:
    out: Weights = {}
    for name, shape in spec.weights:
        if name not in w:
            raise KeyError(f"{spec.name}: missing weight {name}")
        a = np.ascontiguousarray(w[name], dtype=np.float32)
        if a.shape != tuple(shape):
            raise ValueError(f"{spec.name}: weight {name} has shape {a.shape}, expected {tuple(shape)}")
        out[name] = a
    return out


def random_weights(model: str, seed: int = 0) -> Weights:
    """Random-init weights of the right architecture (ranges follow the trained blobs)."""
    spec = get_model(model)
    rng = np.random.default_rng(seed)
    w: Weights = {}
    for name, shape in spec.weights:
        fan_in = shape[-1] if len(shape) > 1 else 1
        if name == "avg_deg":
            a = np.asarray([PNA_AVG_DEG], dtype=np.float32)
        elif name == "bn_var":
            a = rng.uniform(0.5, 1.5, size=shape).astype(np.float32)
        elif name == "bn_weight":
            a = rng.uniform(0.8, 1.2, size=shape).astype(np.float32)
        elif "embedding" in name or name.endswith("bias") or name in ("bn_bias", "bn_mean", "convs_root_emb_weight"):
            a = rng.uniform(-0.3, 0.3, size=shape).astype(np.float32)
        else:
            bound = 1.0 / np.sqrt(max(fan_in, 1))
            a = rng.uniform(-bound, bound, size=shape).astype(np.float32)
        w[name] = a
    if spec.name == "gat":
        # layer-0 projections only see head_in 0, dim_in < 9 (GAT/src/host_load.cc:69-78)
        for k in ("linear_proj_weights", "skip_proj_weights"):
            w[k][0, :, :, 1:, :] = 0
            w[k][0, :, :, 0, 9:] = 0
    return w

"""TEST INFRASTRUCTURE (oracle).  numpy restatements of the reference's GIN kernel in ap_fixed<16,6> arithmetic
(GIN/src/dcl.h:58-59; value = raw / 1024) and of its DGN kernel in ap_fixed<16,3> (DGN/src/dcl.h:54-55; value = raw / 8192),
on raw int16 bit patterns.

Each assignment to an FM_TYPE variable floors to 10 fraction bits (AP_TRN) and keeps the low 16 bits (AP_WRAP):
``acc += a * w`` is ``acc = wrap16(acc + ((a * w) >> 10))`` (linear.cc:41), additions wrap, relu tests the sign of the wrapped
value (util.h:20-25), and ``x / num_of_nodes`` is an integer division of the raw value toward zero (finalize.cc:112 through
ap_fixed_base::operator/).  Wrap-around sums are order independent, so the restatement vectorises freely.

Only tests/ may import this.  It is pinned against oracle/_ref/libflowgnn_ref_gin_fixed.so (tests/test_oracle.py).
"""
from __future__ import annotations

import numpy as np

_ND_OFF = np.array([0, 119, 123, 135, 147, 157, 163, 169, 171])      # GIN/src/load_inputs.cc:5
_ED_OFF = np.array([0, 5, 11])                                         # GIN/src/message_passing.cc:3
F = 10


def _wrap(x):
    return ((np.asarray(x, dtype=np.int64) + 32768) & 0xFFFF) - 32768


def _q(x, frac=F):
    return _wrap(np.floor(np.asarray(x, dtype=np.float64) * (1 << frac)).astype(np.int64))


def _tdiv(a, b):
    """Integer division toward zero (C++ `/`, which ap_fixed_base::operator/ applies to the raw operands); b != 0."""
    a = np.asarray(a, dtype=np.int64)
    b = np.asarray(b, dtype=np.int64)
    return (np.abs(a) // np.abs(b)) * np.sign(a) * np.sign(b)


def _abs16(x):
    """hls::abs on ap_fixed<16,I> returns the same type: -(-2^15) wraps back to -2^15."""
    return np.where(x < 0, _wrap(-x), x)


def _relu(x):
    return np.where(x < 0, 0, x)


def _linear(a, w, b, chunk=256):
    """wrap16(b + sum_k floor(a_k * w_ok / 1024)); a [N][K], w [O][K], b [O] (node_embedding.cc:119-127, 165-175)."""
    out = np.empty((a.shape[0], w.shape[0]), dtype=np.int64)
    for i in range(0, a.shape[0], chunk):
        prod = a[i:i + chunk, None, :] * w[None, :, :]
        out[i:i + chunk] = (prod >> F).sum(axis=2) + b[None, :]
    return _wrap(out)


def gin_fixed(batch, weights) -> np.ndarray:
    """Raw int16 prediction of every graph in ``batch`` (GIN/src/GIN_compute.cc:44-98 in ap_fixed<16,6>)."""
    ne = _q(weights["node_embedding_weight"])
    ee = _q(weights["edge_embedding_weight"])
    w1, b1 = _q(weights["node_mlp_1_weights"]), _q(weights["node_mlp_1_bias"])
    w2, b2 = _q(weights["node_mlp_2_weights"]), _q(weights["node_mlp_2_bias"])
    pw, pb = _q(weights["graph_pred_weights"]).reshape(-1), _q(weights["graph_pred_bias"]).reshape(-1)
    nn = np.asarray(batch.nums_of_nodes, dtype=np.int64)
    ned = np.asarray(batch.nums_of_edges, dtype=np.int64)
    node_off = np.concatenate([[0], np.cumsum(nn)])
    gid_e = np.repeat(np.arange(len(nn)), ned)
    src = batch.edge_list[:, 0].astype(np.int64) + node_off[gid_e]
    dst = batch.edge_list[:, 1].astype(np.int64) + node_off[gid_e]
    h = _wrap(ne[batch.node_feature.astype(np.int64) + _ND_OFF[None, :]].sum(axis=1))           # load_inputs.cc:174-220
    attr = batch.edge_attr.astype(np.int64) + _ED_OFF[None, :]
    for l in range(5):
        edge_embed = _wrap(ee[l][attr].sum(axis=1))                                                # message_passing.cc:136-143
        msg = np.zeros_like(h)
        np.add.at(msg, dst, _relu(_wrap(edge_embed + h[src])))                                     # :145-146
        a = _wrap(msg + h)                                                                         # node_embedding.cc:108, eps == 0
        z = _relu(_linear(a, w1[l], b1[l]))
        h = _linear(z, w2[l], b2[l])
        if l != 4:
            h = _relu(h)
    out = np.zeros(len(nn), dtype=np.int64)
    for g in range(len(nn)):
        s = _wrap(h[node_off[g]:node_off[g + 1]].sum(axis=0))
        n = int(nn[g])
        hg = (np.abs(s) // n) * np.sign(s) if n else np.zeros_like(s)                              # toward zero
        out[g] = _wrap(((hg * pw) >> F).sum() + pb[0])
    return out.astype(np.int16)


def dgn_fixed(batch, weights) -> np.ndarray:
    """Raw int16 prediction of every graph in ``batch`` (DGN/src/DGN_compute.cc:36-103 in ap_fixed<16,3>)."""
    FD = 13
    q = lambda x: _q(x, FD)
    emb = q(weights["embedding_h_atom_embedding_list_weights"])                                   # [9][119][100]
    w = q(weights["layers_posttrans_fully_connected_0_linear_weight"]).reshape(4, 100, 2, 100)      # [l][out][part][in]
    bias = q(weights["layers_posttrans_fully_connected_0_linear_bias"])
    m0w, m0b = q(weights["MLP_layer_FC_layers_0_weight"]), q(weights["MLP_layer_FC_layers_0_bias"])
    m1w, m1b = q(weights["MLP_layer_FC_layers_1_weight"]), q(weights["MLP_layer_FC_layers_1_bias"])
    m2w, m2b = q(weights["MLP_layer_FC_layers_2_weight"]), q(weights["MLP_layer_FC_layers_2_bias"])
    nn = np.asarray(batch.nums_of_nodes, dtype=np.int64)
    ned = np.asarray(batch.nums_of_edges, dtype=np.int64)
    N = int(nn.sum())
    node_off = np.concatenate([[0], np.cumsum(nn)])
    gid_e = np.repeat(np.arange(len(nn)), ned)
    src = batch.edge_list[:, 0].astype(np.int64) + node_off[gid_e]
    dst = batch.edge_list[:, 1].astype(np.int64) + node_off[gid_e]
    feat = batch.node_feature.astype(np.int64)
    h = _wrap(sum(emb[f][feat[:, f]] for f in range(9)))                                           # load_inputs.cc:114-168
    phi = q(batch.node_eigen[:, 1])                                                                # host cast, host_load.cc:207-211
    eig_w = _wrap(phi[src] - phi[dst])                                                             # load_inputs.cc:104-107
    abssum = np.zeros(N, dtype=np.int64); np.add.at(abssum, dst, _abs16(eig_w)); abssum = _wrap(abssum)
    wsum = np.zeros(N, dtype=np.int64); np.add.at(wsum, dst, eig_w); wsum = _wrap(wsum)
    outdeg = np.bincount(src, minlength=N).astype(np.int64)                                        # degree_table, load_inputs.cc:68
    den = np.where(abssum == 0, 1, abssum)                                                         # node_embedding.cc:126-129
    for l in range(4):
        m0 = np.zeros_like(h); np.add.at(m0, dst, h[src]); m0 = _wrap(m0)                          # message_passing.cc:149
        m1 = np.zeros_like(h); np.add.at(m1, dst, (h[src] * eig_w[:, None]) >> FD); m1 = _wrap(m1)  # :150
        a1 = np.where(outdeg[:, None] == 0, 0, _tdiv(m0, np.maximum(outdeg, 1)[:, None]))          # node_embedding.cc:145 (x / 0: emulation gives 0)
        num = (m1 << FD) - wsum[:, None] * h                                                       # 26 fraction bits
        a2 = _abs16(_wrap(_tdiv(num << FD, den[:, None]) >> FD))                                   # :146
        acc = np.empty_like(h)
        for i in range(0, N, 256):
            t = a1[i:i + 256, None, :] * w[l][None, :, 0, :] + a2[i:i + 256, None, :] * w[l][None, :, 1, :]
            acc[i:i + 256] = (t >> FD).sum(axis=2) + bias[l][None, :]                              # :152-157, one floor per pair of products
        h = _wrap(h + _relu(_wrap(acc)))                                                           # :176-178
    out = np.zeros(len(nn), dtype=np.int64)
    for g in range(len(nn)):
        s = _wrap(h[node_off[g]:node_off[g + 1]].sum(axis=0))
        n = int(nn[g])
        x = _tdiv(s, n) if n else np.zeros_like(s)                                                 # finalize.cc:101
        y = _relu(_wrap(((x[None, :] * m0w) >> FD).sum(axis=1) + m0b))                             # linear.cc, RELU defaults to true
        z = _relu(_wrap(((y[None, :] * m1w) >> FD).sum(axis=1) + m1b))
        out[g] = _wrap(((z * m2w.reshape(-1)) >> FD).sum() + m2b.reshape(-1)[0])
    return out.astype(np.int16)

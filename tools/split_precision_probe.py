#!/usr/bin/env python
"""Numerics probe (CPU): how much error do tensor-core operand splits add to the GIN forward?

Emulates the node-MLP GEMMs with operands split into tf32 / fp16 / bf16 pieces (products summed
in float64, so only the SPLIT error shows) and compares per-graph predictions with the golden
reference outputs in tests/golden/.  Used to choose the tcgen05 operand format; see DESIGN.md.
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flowgnn_b200.dataset import load_npz
from flowgnn_b200.weights import load_weights

def trunc_bits(x, drop):
    xi = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return (xi & np.uint32((0xFFFFFFFF << drop) & 0xFFFFFFFF)).view(np.float32)

def rn_bf16(x):
    xi = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((xi + 0x7FFF + ((xi >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)

def split(x, fmt, terms):
    parts = []
    r = x.astype(np.float32)
    for _ in range(terms):
        if fmt == 'tf32': p = trunc_bits(r, 13)
        elif fmt == 'fp16': p = r.astype(np.float16).astype(np.float32)
        elif fmt == 'bf16': p = rn_bf16(r)
        parts.append(p); r = (r - p).astype(np.float32)
    return parts

def gemm(a, w, mode):
    if mode == 'fp32': return (a.astype(np.float64) @ w.astype(np.float64).T).astype(np.float32)
    fmt, nterms, pairs = mode
    ap, wp = split(a, fmt, nterms), split(w, fmt, nterms)
    acc = 0
    for i, j in pairs: acc = acc + ap[i].astype(np.float64) @ wp[j].astype(np.float64).T
    return acc.astype(np.float32)

def gin_forward(b, w, mode, stats=None):
    N = b.total_nodes
    goff = b.node_offsets; eoff = b.edge_offsets
    gid_e = np.repeat(np.arange(b.num_graphs), b.nums_of_edges)
    u = b.edge_list[:,0] + goff[gid_e]; v = b.edge_list[:,1] + goff[gid_e]
    offs = np.array([0,119,123,135,147,157,163,169,171])
    h = w['node_embedding_weight'][b.node_feature + offs].sum(1).astype(np.float32)
    eo = np.array([0,5,11])
    for l in range(5):
        ee = w['edge_embedding_weight'][l][b.edge_attr + eo].sum(1)
        msg = np.maximum(h[u] + ee, 0)
        m = np.zeros_like(h); np.add.at(m, v, msg)
        a = m + h
        z = gemm(a, w['node_mlp_1_weights'][l], mode) + w['node_mlp_1_bias'][l]
        z = np.maximum(z, 0)
        if stats is not None: stats.append((np.abs(a).max(), np.abs(z).max()))
        h = gemm(z, w['node_mlp_2_weights'][l], mode) + w['node_mlp_2_bias'][l]
        if l != 4: h = np.maximum(h, 0)
    gid_n = np.repeat(np.arange(b.num_graphs), b.nums_of_nodes)
    pooled = np.zeros((b.num_graphs, 100)); np.add.at(pooled, gid_n, h)
    pooled /= b.nums_of_nodes[:,None]
    return (pooled @ w['graph_pred_weights'][0] + w['graph_pred_bias'][0]).astype(np.float32)

if __name__ == '__main__':
    G = os.path.join(ROOT, 'tests', 'golden')
    w = load_weights('gin', os.path.join(G, 'weights', 'GIN'))
    for ds, vn in (('molhiv', False), ('molpcba', False), ('hep10k', False), ('molhiv', True)):
        b = load_npz(os.path.join(G, ds + '.npz'))
        gold = np.load(os.path.join(G, f'golden_{ds}.npz'))['ginvn' if vn else 'gin']
        if vn: b = b.with_virtual_node()
        modes = {'fp32': 'fp32',
                 '3xTF32': ('tf32', 2, [(0,0),(0,1),(1,0)]),
                 '1xTF32': ('tf32', 1, [(0,0)]),
                 '3xFP16': ('fp16', 2, [(0,0),(0,1),(1,0)]),
                 '3xBF16': ('bf16', 3, [(0,0),(0,1),(1,0)]),
                 '6xBF16': ('bf16', 3, [(0,0),(0,1),(1,0),(1,1),(0,2),(2,0)])}
        for name, mode in modes.items():
            stats = []
            y = gin_forward(b, w, mode, stats)
            err = np.abs(y - gold) / np.maximum(1, np.abs(gold))
            print(f"{ds}{'+vn' if vn else ''} {name:7s} max scaled err {err.max():.3e} mean {err.mean():.3e}  max|a|,|z| per layer: " + ' '.join(f'{a:.0f}/{z:.0f}' for a, z in stats))

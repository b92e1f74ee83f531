"""CPU: the operand split used on the tensor cores keeps the fp32 contract.

The GIN, PNA, GCN and DGN node transforms run on tcgen05 as THREE bf16 products (hi*hi + lo*hi + hi*lo, x = hi + lo with
both parts bf16, round to nearest).  This test emulates the GCN forward in numpy (float64 accumulation, so that only the
operand split shows) on shipped molhiv graphs and compares with the golden reference outputs (tests/golden/, made from
the reference's own sources): the 3-product split must stay far inside the 1e-4 bar of BASELINE.json's north_star, and a
single bf16 product must NOT (which is why the cheaper format was rejected).
Math: SURVEY.md App. A "GCN" (GCN/src/node_embedding.cc:98-146, message_passing.cc, finalize.cc:94-109).
"""
import numpy as np

from conftest import assert_parity


def rn_bf16(x):
    xi = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((xi + 0x7FFF + ((xi >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)


def linear(a, w, products):
    """a [N][K] x w [O][K]^T with the operands split into bf16 pieces; products = list of (piece of a, piece of w)."""
    if products is None:
        return (a.astype(np.float64) @ w.astype(np.float64).T).astype(np.float32)
    a = a.astype(np.float32)
    a_hi = rn_bf16(a); a_lo = rn_bf16(a - a_hi)
    w_hi = rn_bf16(w); w_lo = rn_bf16(w - w_hi)
    ap, wp = (a_hi, a_lo), (w_hi, w_lo)
    acc = 0
    for i, j in products:
        acc = acc + ap[i].astype(np.float64) @ wp[j].astype(np.float64).T
    return acc.astype(np.float32)


def gcn_forward(b, w, products):
    goff = b.node_offsets
    gid_e = np.repeat(np.arange(b.num_graphs), b.nums_of_edges)
    u = b.edge_list[:, 0] + goff[gid_e]
    v = b.edge_list[:, 1] + goff[gid_e]
    N = b.total_nodes
    outdeg = np.bincount(u, minlength=N).astype(np.float32)
    dis = (1.0 / np.sqrt(outdeg + 1.0)).astype(np.float32)
    norm = (dis[u] * dis[v]).astype(np.float32)
    offs = np.array([0, 119, 123, 135, 147, 157, 163, 169, 171])
    a = w["node_embedding_weight"][b.node_feature + offs].sum(1).astype(np.float32)
    eo = np.array([0, 5, 11])
    for l in range(5):
        p = linear(a, w["convs_weight"][l], products) + w["convs_bias"][l]
        ee = w["edge_embedding_weight"][l][b.edge_attr + eo].sum(1)
        m = np.zeros_like(p)
        np.add.at(m, v, norm[:, None] * np.maximum(p[u] + ee, 0))
        q = m + np.maximum(p + w["convs_root_emb_weight"][l], 0) / (outdeg + 1.0)[:, None]
        x = (q - w["bn_mean"][l]) / np.sqrt(w["bn_var"][l] + np.float32(2.0 ** -10)) * w["bn_weight"][l] + w["bn_bias"][l]
        a = np.maximum(x, 0).astype(np.float32)
    gid_n = np.repeat(np.arange(b.num_graphs), b.nums_of_nodes)
    pooled = np.zeros((b.num_graphs, 100))
    np.add.at(pooled, gid_n, x)
    pooled /= b.nums_of_nodes[:, None]
    return (pooled @ w["graph_pred_weights"][0] + w["graph_pred_bias"][0]).astype(np.float32)


def scaled_err(y, ref):
    return float(np.max(np.abs(y - ref) / np.maximum(1.0, np.abs(ref))))


def test_three_product_bf16_split_keeps_the_fp32_contract(weights, datasets, golden):
    b = datasets["molhiv"].slice(0, 600)
    ref = golden["molhiv"]["gcn"][:600]
    w = weights["gcn"]
    exact = gcn_forward(b, w, None)
    assert_parity(exact, ref, what="numpy GCN restatement vs reference outputs")            # the emulation itself is right
    three = gcn_forward(b, w, [(0, 0), (1, 0), (0, 1)])
    assert scaled_err(three, ref) < 2e-5, scaled_err(three, ref)
    assert scaled_err(three, exact) < 1e-5, scaled_err(three, exact)
    one = gcn_forward(b, w, [(0, 0)])
    assert scaled_err(one, ref) > 1e-4, "a single bf16 product would have been enough -- revisit the operand format"

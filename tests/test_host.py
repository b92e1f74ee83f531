"""CPU: the host-side mirror of the reference's loaders (host.cc / host_load.cc restated)."""
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from flowgnn_b200 import dataset as D
from flowgnn_b200.models import MODELS, get_model
from flowgnn_b200.weights import load_weights, random_weights


def test_weight_shapes_follow_kernel_argument_lists(weights):
    for m, w in weights.items():
        spec = get_model(m)
        assert list(w) == spec.weight_names
        for name, shape in spec.weights:
            assert w[name].shape == tuple(shape) and w[name].dtype == np.float32
    assert sum(v.size for v in weights["gin"].values()) == 225406 - 5          # blob minus the five unused eps
    assert float(weights["pna"]["avg_deg"][0]) == np.float32(6.885701656341553)
    # GAT layer-0 projections only see head_in 0, dim_in < 9 (GAT/src/host_load.cc:69-78)
    for k in ("linear_proj_weights", "skip_proj_weights"):
        assert not weights["gat"][k][0, :, :, 1:, :].any() and not weights["gat"][k][0, :, :, 0, 9:].any()


def test_random_weights_have_the_right_architecture():
    for m in MODELS:
        w = random_weights(m, seed=3)
        for name, shape in get_model(m).weights:
            assert w[name].shape == tuple(shape)


def test_missing_weight_file_is_an_error(tmp_path):
    with pytest.raises(FileNotFoundError):
        load_weights("gcn", str(tmp_path))


def test_dataset_statistics_match_survey(datasets):
    b = datasets["molhiv"]
    assert (b.num_graphs, b.total_nodes, b.total_edges) == (4113, 103927, 228654)
    assert (b.nums_of_nodes.min(), b.nums_of_nodes.max()) == (6, 183)
    assert b.node_eigen is not None and b.edge_attr is not None
    # molhiv g1 is the analysis graph of GIN/src/dcl.h:47-54: 19 nodes, 40 edges
    assert (b.nums_of_nodes[0], b.nums_of_edges[0]) == (19, 40)
    h = datasets["hep10k"]
    assert not h.node_feature.any() and np.array_equal(h.nums_of_edges, h.nums_of_nodes * np.minimum(16, h.nums_of_nodes - 1))


def test_slice_select_concat_roundtrip(datasets):
    b = datasets["molhiv"].slice(0, 50)
    parts = [b.slice(0, 17), b.slice(17, 18), b.slice(18, 50)]
    c = D.concat(parts)
    for f in ("nums_of_nodes", "nums_of_edges", "node_feature", "edge_list", "edge_attr", "node_eigen"):
        assert np.array_equal(getattr(b, f), getattr(c, f))
    s = b.select([3, 3, 7])
    assert s.num_graphs == 3 and np.array_equal(s.slice(0, 1).edge_list, s.slice(1, 2).edge_list)
    assert b.slice(10, 10).num_graphs == 0 and b.tile(120).num_graphs == 120


def test_packed_and_npz_formats_roundtrip(datasets, tmp_path):
    b = datasets["molpcba"].slice(5, 40)
    b.save_packed(str(tmp_path / "x.fgb"))
    b.save_npz(str(tmp_path / "x.npz"))
    for c in (D.load_packed(str(tmp_path / "x.fgb")), D.load_npz(str(tmp_path / "x.npz"))):
        for f in ("nums_of_nodes", "nums_of_edges", "node_feature", "edge_list", "edge_attr", "node_eigen"):
            assert np.array_equal(getattr(b, f), getattr(c, f)), f
    with pytest.raises(ValueError):
        (tmp_path / "bad.fgb").write_bytes(b"nope" * 20)
        D.load_packed(str(tmp_path / "bad.fgb"))


def test_virtual_node_augmentation_matches_reference_host(datasets):
    """Literal restatement of GIN-VN/src/host_load.cc:125-153 for a few graphs."""
    b = datasets["molhiv"].slice(0, 5)
    v = b.with_virtual_node()
    assert np.array_equal(v.nums_of_nodes, b.nums_of_nodes + 1)
    assert np.array_equal(v.nums_of_edges, b.nums_of_edges + 2 * b.nums_of_nodes)
    for g in range(5):
        o, a = b.slice(g, g + 1), v.slice(g, g + 1)
        n, e = int(o.nums_of_nodes[0]), int(o.nums_of_edges[0])
        assert np.array_equal(a.node_feature[:n], o.node_feature) and not a.node_feature[n].any()
        assert np.array_equal(a.edge_list[:e], o.edge_list) and np.array_equal(a.edge_attr[:e], o.edge_attr)
        for nd in range(n):
            assert tuple(a.edge_list[e + 2 * nd]) == (nd, n) and tuple(a.edge_list[e + 2 * nd + 1]) == (n, nd)
        assert not a.edge_attr[e:].any()


def test_eigen_text_parser():
    txt = b"tensor([[-2.2942e-01, -3.0708e-01, -1.1516e-01, -5.3229e-01],\n        [ 1.0e+00,  2.5e-17,  3.0, -4.0]])"
    e = D.parse_eigen_text(txt, 2)
    assert e.shape == (2, 4) and e[0, 1] == np.float32(-0.30708) and e[1, 3] == -4.0
    with pytest.raises(ValueError):
        D.parse_eigen_text(txt, 3)


def test_reference_file_layout_reader(datasets, tmp_path):
    """Write three graphs in the reference's per-graph layout (CRLF info files) and read them back."""
    b = datasets["molhiv"].slice(0, 3)
    (tmp_path / "graphs" / "graph_info").mkdir(parents=True)
    (tmp_path / "graphs" / "graph_bin").mkdir(parents=True)
    (tmp_path / "DGN" / "eig").mkdir(parents=True)
    (tmp_path / "common" / "includes" / "dataset").mkdir(parents=True)
    (tmp_path / "common" / "includes" / "dataset" / "dataset_size.txt").write_text("3")
    for g in range(3):
        s = b.slice(g, g + 1)
        (tmp_path / "graphs" / "graph_info" / f"g{g + 1}_info.txt").write_bytes(f"{s.total_nodes}\r\n{s.total_edges}\r\n".encode())
        s.node_feature.tofile(tmp_path / "graphs" / "graph_bin" / f"g{g + 1}_node_feature.bin")
        s.edge_list.tofile(tmp_path / "graphs" / "graph_bin" / f"g{g + 1}_edge_list.bin")
        s.edge_attr.tofile(tmp_path / "graphs" / "graph_bin" / f"g{g + 1}_edge_attr.bin")
        rows = ",\n        ".join("[" + ", ".join(f"{x:.9e}" for x in r) + "]" for r in s.node_eigen)
        (tmp_path / "DGN" / "eig" / f"g{g + 1}.txt").write_text(f"tensor([{rows}])")
    c = D.load_dataset_dir(str(tmp_path), with_eigen=True)
    for f in ("nums_of_nodes", "nums_of_edges", "node_feature", "edge_list", "edge_attr"):
        assert np.array_equal(getattr(b, f), getattr(c, f)), f
    assert np.allclose(b.node_eigen, c.node_eigen, rtol=1e-6, atol=0)


def test_synthetic_generators_are_deterministic_and_shaped():
    a = D.synthetic_molecules(300, "molhiv", seed=7)
    b = D.synthetic_molecules(300, "molhiv", seed=7)
    assert np.array_equal(a.edge_list, b.edge_list) and np.array_equal(a.node_feature, b.node_feature)
    assert 21.0 < a.nums_of_nodes.mean() < 29.0 and 1.9 < a.total_edges / a.total_nodes < 2.5
    assert a.nums_of_nodes.min() >= 6 and a.nums_of_nodes.max() <= 183
    # symmetric, both directions adjacent, identical attrs, no isolated atoms, valid vocabularies
    assert np.array_equal(a.edge_list[0::2], a.edge_list[1::2][:, ::-1]) and np.array_equal(a.edge_attr[0::2], a.edge_attr[1::2])
    off = a.node_offsets
    for g in range(20):
        e = a.slice(g, g + 1).edge_list
        assert set(range(int(a.nums_of_nodes[g]))) == set(e[:, 0].tolist())
        assert e.max() < a.nums_of_nodes[g] and np.bincount(e[:, 0]).max() <= 4
    assert (a.node_feature < np.array([119, 4, 12, 12, 10, 6, 6, 2, 2])).all() and (a.edge_attr < np.array([5, 6, 2])).all()
    h = D.synthetic_hep(40, seed=7)
    assert np.array_equal(h.nums_of_edges, h.nums_of_nodes * np.minimum(16, h.nums_of_nodes - 1))
    assert (np.diff(h.slice(0, 1).edge_list[:, 0]) >= 0).all()


def test_shard_ranges_cover_and_balance(datasets):
    b = datasets["molhiv"]
    for world in (1, 2, 3, 8):
        r = D.shard_ranges(b, world)
        assert r[0] == 0 and r[-1] == b.num_graphs and (np.diff(r) >= 0).all() and len(r) == world + 1
        cost = [b.slice(int(r[k]), int(r[k + 1])).total_nodes for k in range(world)]
        assert max(cost) - min(cost) < 400
    tiny = b.slice(0, 2)
    r = D.shard_ranges(tiny, 8)
    assert r[0] == 0 and r[-1] == 2 and (np.diff(r) >= 0).all()


def test_batch_validation_rejects_inconsistent_arrays():
    with pytest.raises(ValueError):
        D.Batch(np.array([2]), np.array([1]), np.zeros((3, 9), np.int32), np.zeros((1, 2), np.int32))
    with pytest.raises(ValueError):
        D.Batch(np.array([2]), np.array([2]), np.zeros((2, 9), np.int32), np.zeros((1, 2), np.int32))


def test_header_declares_exactly_the_bound_symbols():
    from flowgnn_b200.capi import EXPORTED_SYMBOLS
    text = open(os.path.join(ROOT, "include", "flowgnn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text)) - {"defined"}
    assert declared == set(EXPORTED_SYMBOLS)

"""the reference's own arithmetic -- ap_fixed<16,6> for GIN / GIN-VN (GIN/src/dcl.h:58-59),
ap_fixed<16,3> for DGN (DGN/src/dcl.h:54-55) -- bit for bit.

The checker is the reference's UNMODIFIED kernel sources compiled over an emulation of Vitis' ap_fixed (oracle/shim_fixed/:
int16 storage, exact wide intermediates, floor on assignment, wrap on overflow, integer division toward zero) --
oracle/_ref/libflowgnn_ref_<model>_fixed.so -- and its committed outputs tests/golden/golden_fixed_<dataset>.npz
(tools/make_fixed_fixtures.py).  A vectorised numpy restatement (oracle/fixed_port.py) is pinned against both.
The bar is equality of every output bit.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import fixed_port, refbind


@pytest.fixture(scope="module")
def golden_fixed():
    return {ds: dict(np.load(os.path.join(GOLDEN, f"golden_fixed_{ds}.npz"))) for ds in ("molhiv", "molpcba", "hep10k")}


# ---- CPU: the cast, the emulation, the restatement ---------------------------------------------------------------------

def test_host_cast_floors_and_wraps():
    """(WT_TYPE)float, GIN/src/host_load.cc:60-97: AP_TRN rounds toward minus infinity, AP_WRAP keeps the low 16 bits."""
    from flowgnn_b200.capi import to_fixed
    x = np.array([0.0, 1.0, -1.0, 0.00097, -0.00001, 31.9995, 32.0, -32.0, -32.001, 33.5, 1.0 / 1024, -1.0 / 1024, 0.5004], dtype=np.float32)
    want = np.array([0, 1024, -1024, 0, -1, 32767, -32768, -32768, 32766, -31232, 1, -1, 512], dtype=np.int16)
    assert np.array_equal(to_fixed(x, 10), want)
    assert np.array_equal(refbind.to_fixed(x, 10), want)
    # ap_fixed<16,3> (DGN): 13 fraction bits, range [-4, 4)
    assert np.array_equal(to_fixed(np.array([3.99995, 4.0, -4.0, -0.00001], dtype=np.float32), 13), np.array([32767, -32768, -32768, -1], dtype=np.int16))


@pytest.mark.parametrize("ds,count", [("molhiv", 150), ("molpcba", 150), ("hep10k", 8)])
@pytest.mark.parametrize("vn", [False, True])
def test_restatement_matches_committed_reference_outputs(ds, count, vn, weights, datasets, golden_fixed):
    """oracle/fixed_port.py against what the reference's own sources computed (committed golden), no reference tree needed."""
    b = datasets[ds].slice(0, count)
    got = fixed_port.gin_fixed(b.with_virtual_node() if vn else b, weights["gin"])
    assert np.array_equal(got, golden_fixed[ds]["ginvn" if vn else "gin"][:count])


@pytest.mark.parametrize("ds,count", [("molhiv", 150), ("molpcba", 150), ("hep10k", 8)])
def test_dgn_restatement_matches_committed_reference_outputs(ds, count, weights, datasets, golden_fixed):
    got = fixed_port.dgn_fixed(datasets[ds].slice(0, count), weights["dgn"])
    assert np.array_equal(got, golden_fixed[ds]["dgn"][:count])


@pytest.mark.skipif(not refbind.have_ref_fixed("gin"), reason="oracle/_ref fixed build missing (needs /root/reference)")
@pytest.mark.parametrize("model", ["gin", "ginvn", "dgn"])
def test_reference_sources_over_the_emulation_reproduce_the_golden(model, weights, datasets, golden_fixed):
    b = datasets["molhiv"].slice(40, 100)
    got = refbind.run_reference_fixed(model, b.with_virtual_node() if model == "ginvn" else b, weights[model])
    assert np.array_equal(got, golden_fixed["molhiv"][model][40:100])


def test_fixed_point_is_not_the_fp32_flavour(golden, golden_fixed):
    """the two flavours differ by far more than the fp32 parity bar, so neither can stand in for the other."""
    d = np.abs(golden_fixed["molhiv"]["gin"].astype(np.float64) / 1024.0 - golden["molhiv"]["gin"])
    assert d.max() > 0.1 and np.median(d) > 1e-3


# ---- GPU: gin_fixed.cu through the C ABI --------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def ctx():
    from flowgnn_b200.capi import Context
    c = Context(0)
    yield c
    c.close()


def _run_fixed(ctx, model, batch, w):
    try:
        ctx.set_option("fixed_point", 1)
        y = ctx.run(model, batch, w)
    finally:
        ctx.set_option("fixed_point", 0)
    raw = y.astype(np.float64) * (8192.0 if model == "dgn" else 1024.0)
    assert np.array_equal(raw, np.round(raw)), "fixed-point results must be multiples of 2^-F"
    return raw.astype(np.int64).astype(np.int16)


@pytest.mark.gpu
@pytest.mark.parametrize("ds", ["molhiv", "molpcba", "hep10k"])
@pytest.mark.parametrize("vn", [False, True])
def test_gin_fixed_point_is_bit_exact(ds, vn, ctx, weights, datasets, golden_fixed):
    """Every shipped molhiv graph, 4,113 molpcba graphs, 200 dense hep10k graphs; with and without the virtual node."""
    want = golden_fixed[ds]["ginvn" if vn else "gin"]
    b = datasets[ds].slice(0, len(want))
    got = _run_fixed(ctx, "gin", b.with_virtual_node() if vn else b, weights["gin"])
    bad = np.flatnonzero(got != want)
    assert bad.size == 0, f"{bad.size} graphs differ, first {bad[:5]}: got {got[bad[:5]]} want {want[bad[:5]]}"


@pytest.mark.gpu
@pytest.mark.parametrize("ds", ["molhiv", "molpcba", "hep10k"])
def test_dgn_fixed_point_is_bit_exact(ds, ctx, weights, datasets, golden_fixed):
    want = golden_fixed[ds]["dgn"]
    got = _run_fixed(ctx, "dgn", datasets[ds].slice(0, len(want)), weights["dgn"])
    bad = np.flatnonzero(got != want)
    assert bad.size == 0, f"{bad.size} graphs differ, first {bad[:5]}: got {got[bad[:5]]} want {want[bad[:5]]}"


@pytest.mark.gpu
def test_dgn_fixed_point_agrees_with_restatement_on_random_graphs(ctx, weights):
    """Directed random graphs: nodes without out-edges (x / 0 is defined as 0 by the emulation), without in-edges (eig_abssum
    = 0 -> epsilon), eigenvector entries across the whole [-4, 4) range so that eig_w and its sums wrap."""
    from flowgnn_b200.dataset import Batch, concat
    rng = np.random.default_rng(11)
    parts = []
    for n in (1, 2, 63, 64, 65, 200, 1300):
        m = 2 * n if n > 1 else 0
        e = rng.integers(0, n, (m, 2)).astype(np.int32)
        feat = np.stack([rng.integers(0, 119, n) for _ in range(9)], 1).astype(np.int32)
        eig = rng.uniform(-4.2, 4.2, (n, 4)).astype(np.float32)
        parts.append(Batch(np.array([n]), np.array([m]), feat, e, None, eig))
    b = concat(parts)
    want = fixed_port.dgn_fixed(b, weights["dgn"])
    got = _run_fixed(ctx, "dgn", b, weights["dgn"])
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_gin_fixed_point_int16_entry_point(weights, datasets, golden_fixed):
    """GIN_compute_graphs_fixed: the FPGA build's kernel ABI (int16 bit patterns in and out)."""
    from flowgnn_b200.capi import compute_graphs_fixed
    b = datasets["molhiv"]
    assert np.array_equal(compute_graphs_fixed("gin", b, weights["gin"]), golden_fixed["molhiv"]["gin"])
    e = b.slice(0, 0)
    assert compute_graphs_fixed("gin", e, weights["gin"]).shape == (0,)


@pytest.mark.gpu
def test_gin_fixed_point_at_baseline_size(ctx, weights, datasets, golden_fixed):
    """BASELINE config C2's batch (41,127 graphs = the molhiv set tiled): the k-th copy of a graph must reproduce the golden
    bits wherever it lands in the batch, and the fp32 path must be untouched by the option afterwards."""
    b = datasets["molhiv"].tile(41127)
    got = _run_fixed(ctx, "gin", b, weights["gin"])
    want = np.resize(golden_fixed["molhiv"]["gin"], 41127)
    assert np.array_equal(got, want)
    y = ctx.run("gin", datasets["molhiv"].slice(0, 64))
    assert np.abs(y.astype(np.float64) * 1024 - np.round(y.astype(np.float64) * 1024)).max() > 0      # fp32 again


@pytest.mark.gpu
def test_gin_fixed_point_agrees_with_restatement_on_random_graphs(ctx, weights):
    """Graphs the fixtures do not hold: isolated nodes, a graph above 1,024 nodes, random weights large enough to wrap."""
    from flowgnn_b200.dataset import Batch, concat
    rng = np.random.default_rng(7)
    parts = []
    for n in (1, 2, 47, 48, 49, 96, 300, 1500):
        m = 3 * n if n > 1 else 0
        e = rng.integers(0, n, (m, 2)).astype(np.int32)
        attr = np.stack([rng.integers(0, k, m) for k in (5, 6, 2)], 1).astype(np.int32).reshape(m, 3)
        feat = np.stack([rng.integers(0, k, n) for k in (119, 4, 12, 12, 10, 6, 6, 2, 2)], 1).astype(np.int32)
        parts.append(Batch(np.array([n]), np.array([m]), feat, e, attr, None))
    b = concat(parts)
    w = {k: (v * 4.0).astype(np.float32) for k, v in weights["gin"].items()}     # pushes activations through the 16-bit wrap
    want = fixed_port.gin_fixed(b, w)
    got = _run_fixed(ctx, "gin", b, w)
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_fixed_point_option_rejects_models_without_it(ctx, weights, datasets):
    from flowgnn_b200.capi import FlowGNNError
    try:
        ctx.set_option("fixed_point", 1)
        with pytest.raises(FlowGNNError, match="fixed_point"):
            ctx.run("gcn", datasets["molhiv"].slice(0, 4), weights["gcn"])
    finally:
        ctx.set_option("fixed_point", 0)

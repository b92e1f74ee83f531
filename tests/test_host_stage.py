"""Narrowed upload of the host-pointer entry points (flowgnn_b200/csrc/host_stage.h, option ``host_stage``).

The reference's kernel ABI passes int32 words for nine atom features per node, two graph-local node ids and three bond attributes
per edge (GIN/src/dcl.h:61-67).  The entry points narrow them to u8 / u16 with a pool of host threads, copy the narrow block and
widen it on the device.  CPU: the host-side step.  GPU: predictions are bit-identical with the step on and off, for pageable and
page-locked callers, any thread count, and inputs that do not fit (those arrays travel unchanged)."""
import os

import numpy as np
import pytest

from flowgnn_b200 import capi


def _lib():
    if not os.path.isfile(capi.LIB_PATH):
        pytest.skip("libflowgnn_b200.so not built (run __graft_entry__.build())")
    return capi.load_library()


@pytest.mark.parametrize("threads", [1, 3, 8])
@pytest.mark.parametrize("width", [1, 2])
@pytest.mark.parametrize("n", [0, 1, 5, 65536, 65537, 1_000_003])
def test_narrow_words_matches_numpy(threads, width, n):
    _lib()
    rng = np.random.default_rng(n + width)
    src = rng.integers(0, 256 if width == 1 else 65536, size=n, dtype=np.int32)
    got, seen = capi.narrow_words(src, width, threads)
    assert np.array_equal(got, src.astype(np.uint8 if width == 1 else np.uint16))
    assert seen == (int(np.bitwise_or.reduce(src)) if n else 0)
    assert seen & ~((1 << (8 * width)) - 1) == 0


@pytest.mark.parametrize("bad", [-1, 256, 70000, np.iinfo(np.int32).min])
def test_narrow_words_reports_values_that_do_not_fit(bad):
    _lib()
    src = np.zeros(200_000, dtype=np.int32)
    src[123_456] = bad
    for width in (1, 2):
        _, seen = capi.narrow_words(src, width, 4)
        fits = 0 <= bad < (1 << (8 * width))
        assert (seen & ~((1 << (8 * width)) - 1) == 0) == fits


def test_narrow_words_rejects_other_widths():
    _lib()
    assert capi.narrow_words(np.zeros(4, np.int32), 4, 1)[1] == 0xFFFFFFFF


def test_narrow_run_probe_binary(tmp_path):
    """tools/hs_probe/narrow_probe.cc drives NarrowRun (all chunks' slices queued at once, the caller helping until its chunk is complete)
    over the bench batch's sizes and checks every narrowed value; built here with the host compiler, with and without the AVX2 path."""
    import shutil
    import subprocess
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "narrow_probe")
    subprocess.run([cxx, "-O2", "-std=c++17", "-pthread", os.path.join(root, "tools", "hs_probe", "narrow_probe.cc"),
                    os.path.join(root, "flowgnn_b200", "csrc", "host_stage.cc"), "-o", exe], check=True)
    for env in ({}, {"FLOWGNN_B200_NO_AVX2": "1"}):
        r = subprocess.run([exe, "3"], capture_output=True, text=True, env={**os.environ, **env}, timeout=120)
        assert r.returncode == 0 and "GB/s" in r.stdout, r.stdout + r.stderr


# ---- GPU --------------------------------------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def ctx():
    from flowgnn_b200.capi import Context
    c = Context(0)
    yield c
    c.close()


def _run_entry(model, batch, w, stage, threads=None, chunks=None):
    env = {"FLOWGNN_B200_HOST_STAGE": stage, "FLOWGNN_B200_HOST_THREADS": threads, "FLOWGNN_B200_CHUNKS": chunks}
    old = {k: os.environ.get(k) for k in env}
    try:
        for k, v in env.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)
        return capi.ReferenceCall(model, batch, w).run().copy()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["gin", "ginvn", "gcn", "gat", "pna", "dgn"])
def test_narrowed_upload_is_bit_identical(model, weights, datasets):
    from flowgnn_b200.dataset import synthetic_molecules
    b = datasets["molhiv"].slice(0, 1500) if model != "dgn" else synthetic_molecules(1500, "molhiv", seed=5, with_eigen=True)
    if model == "ginvn":
        b = datasets["hep10k"].slice(0, 200)
    want = _run_entry(model, b, weights[model], stage=0)
    assert np.isfinite(want).sum() > 0.9 * want.size
    # FLOWGNN_B200_HOST_STAGE is a mask of the arrays to narrow: 1 node_feature, 2 edge_list, 4 edge_attr
    for threads, mask in ((1, 7), (5, 7), (3, 5), (2, 2), (4, 1)):
        got = _run_entry(model, b, weights[model], stage=mask, threads=threads)
        assert np.array_equal(got.view(np.int32), want.view(np.int32)), (model, threads, mask)


@pytest.mark.gpu
@pytest.mark.parametrize("chunks", [1, 2, 3, 5])
def test_narrowed_upload_chunked_large_batch(chunks, ctx, weights, datasets):
    """A batch large enough for the chunked pipeline (two pinned blocks and two device batches alternate), every chunk count;
    default mode = pageable numpy arrays -> narrowed; against the device-resident one-shot path."""
    b = datasets["molpcba"].tile(20000)
    want = ctx.run("gin", b, weights["gin"])
    got = _run_entry("gin", b, weights["gin"], stage=None, chunks=chunks)
    again = _run_entry("gin", b, weights["gin"], stage=5, chunks=chunks)
    assert np.array_equal(got.view(np.int32), want.view(np.int32))
    assert np.array_equal(again.view(np.int32), want.view(np.int32))


@pytest.mark.gpu
def test_large_inputs_are_cut_into_more_chunks(ctx, weights, datasets):
    """Above ~200 MB of caller bytes the entry point cuts weight-3 chunks of about 100 MB (up to 16 chunks): PNA on a 300 MB batch,
    pageable arrays, against the device-resident one-shot run."""
    b = datasets["molpcba"].tile(200000)
    assert b.node_feature.nbytes + b.edge_list.nbytes > 250e6
    want = ctx.run("pna", b, weights["pna"])
    got = _run_entry("pna", b, weights["pna"], stage=None)
    assert np.array_equal(got.view(np.int32), want.view(np.int32))
    h2d, d2h = capi.last_transfer_bytes()
    assert h2d < 0.4 * (b.node_feature.nbytes + b.edge_list.nbytes) and d2h >= 4 * b.num_graphs


@pytest.mark.gpu
def test_narrowed_upload_with_pinned_caller_memory(weights, datasets):
    b = datasets["molhiv"].slice(0, 2000)
    want = _run_entry("gin", b, weights["gin"], stage=0)
    arrays = [b.node_feature, b.edge_list, b.edge_attr]
    for a in arrays:
        capi.pin_host(a)
    try:
        for stage in (0, 7, 5, None):
            got = _run_entry("gin", b, weights["gin"], stage=stage)
            assert np.array_equal(got.view(np.int32), want.view(np.int32)), stage
    finally:
        for a in arrays:
            capi.unpin_host(a)


@pytest.mark.gpu
def test_inputs_that_do_not_fit_travel_unchanged(weights, datasets):
    """Out-of-vocabulary node features (the embedding's fallback path) and ids that are not node ids: the array with such a value is
    uploaded as int32, so the device sees exactly what the caller passed -- same predictions, same error."""
    import copy
    b = copy.deepcopy(datasets["molhiv"].slice(0, 800))
    b.node_feature[17, 3] = 300            # does not fit a byte; out of vocabulary either way
    b.node_feature[400, 0] = -2
    want = _run_entry("gin", b, weights["gin"], stage=0)
    got = _run_entry("gin", b, weights["gin"], stage=7, threads=3)
    assert np.array_equal(got.view(np.int32), want.view(np.int32))
    # an edge id beyond 16 bits is an invalid edge: rejected with the same error on both paths
    b2 = copy.deepcopy(datasets["molhiv"].slice(0, 800))
    b2.edge_list[1000, 1] = 70000
    for stage in (0, 7):
        with pytest.raises(capi.FlowGNNError, match="node id"):
            _run_entry("gin", b2, weights["gin"], stage=stage)
    # ... and so is one that fits 16 bits but is not a node of its graph (caught on the device after the widening)
    b3 = copy.deepcopy(datasets["molhiv"].slice(0, 800))
    b3.edge_list[1000, 1] = 600
    for stage in (0, 7):
        with pytest.raises(capi.FlowGNNError, match="node id"):
            _run_entry("gin", b3, weights["gin"], stage=stage)
    # the entry point still works afterwards
    b4 = datasets["molhiv"].slice(0, 800)
    assert np.array_equal(_run_entry("gin", b4, weights["gin"], stage=7).view(np.int32), _run_entry("gin", b4, weights["gin"], stage=0).view(np.int32))


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["gin", "gat", "pna", "dgn"])
def test_packed_upload_is_bit_identical(model, ctx, weights, datasets):
    """Part 2: flowgnn_b200_upload_batch_packed takes the narrow arrays of the packed dataset files; same predictions as the int32 upload."""
    from flowgnn_b200.dataset import synthetic_molecules
    from flowgnn_b200.models import get_model
    b = datasets["molpcba"].slice(0, 3001) if model != "dgn" else synthetic_molecules(3001, "molhiv", seed=7, with_eigen=True)
    spec = get_model(model)
    want = ctx.run(model, b, weights[model])
    ctx.upload_packed_arrays(b.num_graphs, b.total_nodes, b.total_edges, b.nums_of_nodes, b.nums_of_edges, b.node_feature.astype(np.uint8),
                             b.edge_list.astype(np.uint16), b.edge_attr.astype(np.uint8) if spec.uses_edge_attr else None,
                             b.node_eigen if spec.uses_eigen else None)
    ctx.compute(model)
    got = ctx.download()
    assert np.array_equal(got.view(np.int32), want.view(np.int32)), model
    with pytest.raises(capi.FlowGNNError):
        ctx.upload_packed_arrays(b.num_graphs, b.total_nodes, b.total_edges, b.nums_of_nodes, b.nums_of_edges, None, None)


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["gin", "gcn", "gat", "pna", "dgn"])
def test_packed_entry_point_is_bit_identical(model, ctx, weights, datasets):
    """flowgnn_b200_compute_graphs_packed: the chunked host-pointer pipeline fed with the narrow dataset layout (large enough for
    several chunks; GAT with the reference's node-offset quirk reads features from the start of the caller's buffer)."""
    from flowgnn_b200.dataset import synthetic_molecules
    b = datasets["molpcba"].tile(20000) if model != "dgn" else synthetic_molecules(2500, "molhiv", seed=3, with_eigen=True).tile(20000)
    ctx.set_option("gat_node_offset_bug", 1)
    want = ctx.run(model, b, weights[model])
    call = capi.PackedCall(model, b, weights[model])
    got = call.run().copy()
    assert np.array_equal(got.view(np.int32), want.view(np.int32)), model
    assert np.array_equal(call.run().view(np.int32), want.view(np.int32)), model
    h2d, _ = capi.last_transfer_bytes()
    assert h2d < 0.45 * (b.node_feature.nbytes + b.edge_list.nbytes + (b.edge_attr.nbytes if model in ("gin", "gcn") else 0)
                         + (b.node_eigen.nbytes if model == "dgn" else 0)) + 4e6

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
MODEL_WEIGHT_DIR = {"gin": "GIN", "ginvn": "GIN", "gcn": "GCN", "gat": "GAT", "pna": "PNA", "dgn": "DGN"}
ALL_MODELS = ("gin", "ginvn", "gcn", "gat", "pna", "dgn")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA GPU (B200); run with -m gpu")


@pytest.fixture(scope="session")
def weights():
    from flowgnn_b200.weights import load_weights
    return {m: load_weights(m, os.path.join(GOLDEN, "weights", d)) for m, d in MODEL_WEIGHT_DIR.items()}


@pytest.fixture(scope="session")
def datasets():
    from flowgnn_b200.dataset import load_npz
    return {ds: load_npz(os.path.join(GOLDEN, f"{ds}.npz")) for ds in ("molhiv", "molpcba", "hep10k")}


@pytest.fixture(scope="session")
def golden():
    return {ds: dict(np.load(os.path.join(GOLDEN, f"golden_{ds}.npz"))) for ds in ("molhiv", "molpcba", "hep10k")}


def assert_parity(got, want, tol=1e-4, what=""):
    """The parity bar of BASELINE.json's north_star: |y - y_ref| <= 1e-4 * max(1, |y_ref|), with
    non-finite values matching positionally."""
    got = np.asarray(got, dtype=np.float32)
    want = np.asarray(want, dtype=np.float32)
    assert got.shape == want.shape, what
    nf_w, nf_g = ~np.isfinite(want), ~np.isfinite(got)
    assert np.array_equal(nf_w, nf_g), f"{what}: non-finite positions differ: ref {np.flatnonzero(nf_w)[:10]} vs got {np.flatnonzero(nf_g)[:10]}"
    assert np.array_equal(np.isnan(want), np.isnan(got)), f"{what}: NaN positions differ"
    if nf_w.any():
        assert np.array_equal(want[nf_w & ~np.isnan(want)], got[nf_w & ~np.isnan(want)]), f"{what}: infinities differ in sign"
    ok = ~nf_w
    err = np.abs(got[ok] - want[ok]) / np.maximum(1.0, np.abs(want[ok]))
    worst = int(np.argmax(err)) if err.size else 0
    assert err.size == 0 or err.max() <= tol, (
        f"{what}: max scaled error {err.max():.3e} > {tol:g} at graph {np.flatnonzero(ok)[worst]} "
        f"(got {got[ok][worst]!r}, ref {want[ok][worst]!r})")
    return float(err.max()) if err.size else 0.0

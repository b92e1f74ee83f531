"""CPU: the C-ABI library loads, exports every symbol include/flowgnn_b200.h declares, and fails
loudly (never silently falls back) when there is no GPU.  No compute calls here."""
import ctypes
import os

import numpy as np
import pytest

from flowgnn_b200 import capi


def _lib():
    if not os.path.isfile(capi.LIB_PATH):
        pytest.skip("libflowgnn_b200.so not built (run __graft_entry__.build())")
    return capi.load_library()


def test_library_exports_every_declared_symbol():
    lib = _lib()
    for name in capi.EXPORTED_SYMBOLS:
        assert hasattr(lib, name), name


def test_last_error_is_a_string():
    assert isinstance(_lib().flowgnn_b200_last_error(), bytes)


def test_no_gpu_is_an_error_not_a_fallback():
    lib = _lib()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.FlowGNNError):
        capi.Context(0)
    # the drop-in entry point reports failure through its return code as well
    from flowgnn_b200.dataset import Batch
    from flowgnn_b200.weights import random_weights
    b = Batch(np.array([2]), np.array([2]), np.zeros((2, 9), np.int32), np.array([[0, 1], [1, 0]], np.int32), np.zeros((2, 3), np.int32))
    with pytest.raises(capi.FlowGNNError):
        capi.compute_graphs("gin", b, random_weights("gin"))
    assert lib.flowgnn_b200_last_error() != b""


def test_product_package_never_imports_the_oracle():
    import pathlib
    pkg = pathlib.Path(capi.__file__).parent
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.cc")) + list(pkg.rglob("*.h")):
        txt = p.read_text(errors="replace")
        assert "refbind" not in txt and "flowgnn_oracle" not in txt and "oracle/" not in txt, p

"""TEST INFRASTRUCTURE (oracle).  ctypes bindings to the two CPU checkers:

* ``oracle/_ref/libflowgnn_ref_<model>.so`` -- the reference's own kernel sources compiled
  unmodified against the float shims (``kind="reference"``).  Its state lives in file-scope
  globals (``*/src/globals.cc``), so it is NOT re-entrant: one batch at a time per process.
* ``oracle/libflowgnn_oracle.so`` -- the plain-C restatement (``kind="port"``), re-entrant.

Both export the reference's kernel entry points (``GIN_compute_graphs`` ... with the argument
lists of ``*/src/dcl.h``); the port prefixes them with ``oracle_``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.
The product path (``flowgnn_b200``) never does.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")

_i32p = ctypes.POINTER(ctypes.c_int32)
_f32p = ctypes.POINTER(ctypes.c_float)


def _ptr(a: Optional[np.ndarray], ty):
    if a is None:
        return ctypes.cast(None, ty)
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ty)


def _has_avx2() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            txt = f.read()
        return " avx2 " in txt and " fma " in txt
    except OSError:
        return False


def ref_library_path(model: str, fast: bool = False) -> str:
    tag = model + ("_fast" if fast and _has_avx2() else "")
    return os.path.join(_REF_DIR, f"libflowgnn_ref_{tag}.so")


def port_library_path(fast: bool = False) -> str:
    return os.path.join(_HERE, "libflowgnn_oracle_fast.so" if fast and _has_avx2() else "libflowgnn_oracle.so")


def have_ref(model: str = "gin") -> bool:
    return os.path.isfile(ref_library_path(model))


def have_port() -> bool:
    return os.path.isfile(port_library_path())


_libs: Dict[str, ctypes.CDLL] = {}


def _load(path: str) -> ctypes.CDLL:
    if path not in _libs:
        if not os.path.isfile(path):
            raise FileNotFoundError(f"{path} is missing -- run `make -C oracle` (needs /root/reference for _ref)")
        _libs[path] = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
    return _libs[path]


def _call(fn, spec, batch, weights, gat_node_offset_bug=True, reload_weights=None):
    """Marshal a Batch + weight dict into the reference argument order for ``spec``."""
    G = batch.num_graphs
    out = np.zeros(G, dtype=np.float32)
    if reload_weights is None:
        reload_weights = np.zeros(G, dtype=np.int32)
        if G:
            reload_weights[0] = 1                       # GIN/src/host.cc:135
    nn = np.ascontiguousarray(batch.nums_of_nodes, dtype=np.int32)
    ne = np.ascontiguousarray(batch.nums_of_edges, dtype=np.int32)
    args = [ctypes.c_int(G), _ptr(nn, _i32p), _ptr(ne, _i32p), _ptr(reload_weights, _i32p), _ptr(out, _f32p),
            _ptr(batch.node_feature, _i32p)]
    if spec.uses_eigen:
        if batch.node_eigen is None:
            raise ValueError("DGN needs node_eigen")
        args.append(_ptr(batch.node_eigen, _f32p))
    args.append(_ptr(batch.edge_list, _i32p))
    if spec.uses_edge_attr:
        if batch.edge_attr is None:
            raise ValueError(f"{spec.name} needs edge_attr")
        args.append(_ptr(batch.edge_attr, _i32p))
    keep = [nn, ne, reload_weights]
    for name, _ in spec.weights:
        a = np.ascontiguousarray(weights[name], dtype=np.float32)
        keep.append(a)
        args.append(_ptr(a, _f32p))
    fn.restype = None
    fn(*args)
    return out


def run_reference(model: str, batch, weights, fast: bool = False) -> np.ndarray:
    """Predictions of the UNMODIFIED reference kernel (fp32 csim flavour) for every graph in ``batch``.

    GIN-VN: pass ``batch.with_virtual_node()`` -- the augmentation is host-side in the reference too.
    GAT reproduces the reference's missing node offset because it IS the reference."""
    from flowgnn_b200.models import get_model
    spec = get_model(model)
    lib = _load(ref_library_path(spec.name, fast))
    return _call(getattr(lib, spec.symbol), spec, batch, weights)


def run_port(model: str, batch, weights, fast: bool = False, gat_node_offset_bug: bool = True) -> np.ndarray:
    """Predictions of the plain-C restatement (oracle/flowgnn_oracle.c)."""
    from flowgnn_b200.models import get_model
    spec = get_model(model)
    lib = _load(port_library_path(fast))
    if spec.name == "gat":
        lib.oracle_set_gat_node_offset_bug(ctypes.c_int(1 if gat_node_offset_bug else 0))
    return _call(getattr(lib, "oracle_" + spec.symbol), spec, batch, weights)


# ---- bit-accurate ap_fixed flavour -------------------------------------------------------------------
# oracle/_ref/libflowgnn_ref_<model>_fixed.so = the same unmodified reference sources over oracle/shim_fixed/ (an emulation
# of Vitis' ap_fixed<16,I>: int16 storage, floor on assignment, wrap on overflow).  Arguments and results are int16 bit
# patterns, exactly the kernel ABI of the FPGA build (FM_TYPE / WT_TYPE arrays, */src/dcl.h).

FRAC_BITS = {"gin": 10, "ginvn": 10, "dgn": 13}
_i16p = ctypes.POINTER(ctypes.c_int16)


def to_fixed(x: np.ndarray, frac_bits: int) -> np.ndarray:
    """``(WT_TYPE)float`` of the reference's host (GIN/src/host_load.cc:60-97): floor(x * 2^F), low 16 bits."""
    q = np.floor(np.asarray(x, dtype=np.float64) * float(1 << frac_bits)).astype(np.int64)
    return (q & 0xFFFF).astype(np.uint16).view(np.int16)


def ref_fixed_library_path(model: str) -> str:
    return os.path.join(_REF_DIR, f"libflowgnn_ref_{model}_fixed.so")


def have_ref_fixed(model: str = "gin") -> bool:
    return os.path.isfile(ref_fixed_library_path(model))


def run_reference_fixed(model: str, batch, weights) -> np.ndarray:
    """Raw int16 predictions (value = raw / 2^F) of the unmodified reference kernel in emulated ap_fixed arithmetic.
    ``weights`` are the fp32 arrays of the weight files; they are cast as the reference's host casts them."""
    from flowgnn_b200.models import get_model
    spec = get_model(model)
    F = FRAC_BITS[spec.name]
    lib = _load(ref_fixed_library_path(spec.name))
    G = batch.num_graphs
    out = np.zeros(G, dtype=np.int16)
    reload_weights = np.zeros(G, dtype=np.int32)
    if G:
        reload_weights[0] = 1
    nn = np.ascontiguousarray(batch.nums_of_nodes, dtype=np.int32)
    ne = np.ascontiguousarray(batch.nums_of_edges, dtype=np.int32)
    args = [ctypes.c_int(G), _ptr(nn, _i32p), _ptr(ne, _i32p), _ptr(reload_weights, _i32p), _ptr(out, _i16p),
            _ptr(batch.node_feature, _i32p)]
    keep = [nn, ne, reload_weights]
    if spec.uses_eigen:
        eig = np.ascontiguousarray(to_fixed(batch.node_eigen, F))
        keep.append(eig)
        args.append(_ptr(eig, _i16p))
    args.append(_ptr(batch.edge_list, _i32p))
    if spec.uses_edge_attr:
        args.append(_ptr(batch.edge_attr, _i32p))
    for name, _ in spec.weights:
        a = np.ascontiguousarray(to_fixed(weights[name], F))
        keep.append(a)
        args.append(_ptr(a, _i16p))
    fn = getattr(lib, spec.symbol)
    fn.restype = None
    fn(*args)
    return out

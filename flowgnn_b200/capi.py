"""ctypes binding of ``libflowgnn_b200.so`` (the C ABI declared in ``include/flowgnn_b200.h``).

This is the only way Python reaches the CUDA kernels; there is no CPU fallback.  If the shared
library has not been built (``make -C flowgnn_b200/csrc`` or ``__graft_entry__.build()``), importing
works but every call raises :class:`FlowGNNError`.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional, Sequence

import numpy as np

from .dataset import Batch
from .models import ModelSpec, get_model
from .weights import Weights, check_weights

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libflowgnn_b200.so")

MODEL_IDS = {"gin": 0, "ginvn": 0, "gcn": 1, "gat": 2, "pna": 3, "dgn": 4}

#: every symbol include/flowgnn_b200.h declares
EXPORTED_SYMBOLS = (
    "GIN_compute_graphs", "GIN_compute_graphs_fixed", "GCN_compute_graphs", "GAT_compute_graphs", "PNA_compute_graphs", "DGN_compute_graphs",
    "flowgnn_b200_last_error", "flowgnn_b200_create", "flowgnn_b200_destroy", "flowgnn_b200_set_option",
    "flowgnn_b200_load_weights", "flowgnn_b200_upload_batch", "flowgnn_b200_upload_batch_packed", "flowgnn_b200_compute", "flowgnn_b200_download",
    "flowgnn_b200_last_launch_count", "flowgnn_b200_tile_count", "flowgnn_b200_last_layer_ms", "flowgnn_b200_stream", "flowgnn_b200_synchronize",
    "flowgnn_b200_pin_host", "flowgnn_b200_unpin_host", "flowgnn_b200_narrow_words", "flowgnn_b200_last_transfer_bytes", "flowgnn_b200_compute_graphs_packed",
)


class FlowGNNError(RuntimeError):
    pass


_lib: Optional[ctypes.CDLL] = None
_i32p = ctypes.POINTER(ctypes.c_int32)
_f32p = ctypes.POINTER(ctypes.c_float)


def load_library() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise FlowGNNError(f"{LIB_PATH} not built: run `make -C flowgnn_b200/csrc` (needs nvcc, targets sm_100a)")
        lib = ctypes.CDLL(LIB_PATH)
        lib.flowgnn_b200_last_error.restype = ctypes.c_char_p
        lib.flowgnn_b200_stream.restype = ctypes.c_void_p
        lib.flowgnn_b200_stream.argtypes = [ctypes.c_void_p]
        lib.flowgnn_b200_create.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]
        lib.flowgnn_b200_destroy.argtypes = [ctypes.c_void_p]
        lib.flowgnn_b200_set_option.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]
        lib.flowgnn_b200_load_weights.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(_f32p), ctypes.c_int]
        lib.flowgnn_b200_upload_batch.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
                                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                  ctypes.c_void_p, ctypes.c_void_p]
        lib.flowgnn_b200_upload_batch_packed.argtypes = lib.flowgnn_b200_upload_batch.argtypes
        lib.flowgnn_b200_compute.argtypes = [ctypes.c_void_p, ctypes.c_int, _f32p]
        lib.flowgnn_b200_download.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        lib.flowgnn_b200_last_launch_count.argtypes = [ctypes.c_void_p]
        lib.flowgnn_b200_tile_count.argtypes = [ctypes.c_void_p]
        lib.flowgnn_b200_tile_count.restype = ctypes.c_long
        lib.flowgnn_b200_synchronize.argtypes = [ctypes.c_void_p]
        lib.flowgnn_b200_last_layer_ms.argtypes = [ctypes.c_void_p, _f32p, ctypes.c_int]
        lib.flowgnn_b200_pin_host.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        lib.flowgnn_b200_unpin_host.argtypes = [ctypes.c_void_p]
        lib.flowgnn_b200_last_transfer_bytes.restype = None
        lib.flowgnn_b200_last_transfer_bytes.argtypes = [ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]
        lib.flowgnn_b200_narrow_words.restype = ctypes.c_uint32
        lib.flowgnn_b200_narrow_words.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        _lib = lib
    return _lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load_library().flowgnn_b200_last_error().decode(errors="replace")
        raise FlowGNNError(f"{what} failed with code {rc}: {msg}")


def _addr(a) -> Optional[int]:
    """Address of a numpy array, a torch tensor (e.g. pinned host memory) or None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise FlowGNNError("array must be C-contiguous")
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return int(a.data_ptr())
    raise TypeError(type(a))


def pin_host(a: np.ndarray) -> None:
    """Page-lock a caller-owned array (what a C++ host does once after building its batch vectors)."""
    _check(load_library().flowgnn_b200_pin_host(a.ctypes.data, a.nbytes), "pin_host")


def unpin_host(a: np.ndarray) -> None:
    _check(load_library().flowgnn_b200_unpin_host(a.ctypes.data), "unpin_host")


def last_transfer_bytes():
    """(host -> device, device -> host) bytes of this thread's last ``<MODEL>_compute_graphs`` call."""
    h2d, d2h = ctypes.c_uint64(0), ctypes.c_uint64(0)
    load_library().flowgnn_b200_last_transfer_bytes(ctypes.byref(h2d), ctypes.byref(d2h))
    return int(h2d.value), int(d2h.value)


def narrow_words(src: np.ndarray, width: int, threads: int = 1):
    """The host-side step of the narrowed upload (option ``host_stage``): int32 words -> u8 / u16 with ``threads`` host threads.
    Returns (narrowed array, OR of all source words).  No GPU is touched."""
    src = np.ascontiguousarray(src, dtype=np.int32)
    dst = np.zeros(src.size, dtype=np.uint8 if width == 1 else np.uint16)
    seen = load_library().flowgnn_b200_narrow_words(src.ctypes.data, src.size, int(width), dst.ctypes.data, int(threads))
    return dst, int(seen)


class Context:
    """One GPU's context (extended interface, Part 2 of the header)."""

    def __init__(self, device: int = 0):
        self._lib = load_library()
        self._h = ctypes.c_void_p()
        _check(self._lib.flowgnn_b200_create(ctypes.byref(self._h), device), "flowgnn_b200_create")
        self.device = device
        self._num_graphs = 0

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.flowgnn_b200_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_option(self, name: str, value: int) -> None:
        _check(self._lib.flowgnn_b200_set_option(self._h, name.encode(), int(value)), f"set_option({name})")

    def load_weights(self, model: str, weights: Weights) -> None:
        spec = get_model(model)
        w = check_weights(spec, weights)
        arrs = [w[n] for n in spec.weight_names]
        ptrs = (_f32p * len(arrs))(*[a.ctypes.data_as(_f32p) for a in arrs])
        _check(self._lib.flowgnn_b200_load_weights(self._h, MODEL_IDS[spec.name], ptrs, len(arrs)), "load_weights")

    def upload(self, batch: Batch) -> None:
        self.upload_arrays(batch.num_graphs, batch.total_nodes, batch.total_edges, batch.nums_of_nodes, batch.nums_of_edges,
                           batch.node_feature, batch.edge_list, batch.edge_attr, batch.node_eigen)

    def upload_arrays(self, num_graphs, total_nodes, total_edges, nums_of_nodes, nums_of_edges, node_feature, edge_list,
                      edge_attr=None, node_eigen=None) -> None:
        _check(self._lib.flowgnn_b200_upload_batch(self._h, int(num_graphs), int(total_nodes), int(total_edges),
                                                   _addr(nums_of_nodes), _addr(nums_of_edges), _addr(node_feature),
                                                   _addr(edge_list), _addr(edge_attr), _addr(node_eigen)), "upload_batch")
        self._num_graphs = int(num_graphs)

    def upload_packed_arrays(self, num_graphs, total_nodes, total_edges, nums_of_nodes, nums_of_edges, node_feature_u8, edge_list_u16,
                             edge_attr_u8=None, node_eigen=None) -> None:
        """``flowgnn_b200_upload_batch_packed``: the big arrays in the narrow layout of the packed dataset files (uint8 / uint16 / uint8)."""
        for a, dt in ((node_feature_u8, np.uint8), (edge_list_u16, np.uint16), (edge_attr_u8, np.uint8)):
            if isinstance(a, np.ndarray) and a.dtype != dt:
                raise FlowGNNError(f"packed upload wants {np.dtype(dt).name}, got {a.dtype}")
        _check(self._lib.flowgnn_b200_upload_batch_packed(self._h, int(num_graphs), int(total_nodes), int(total_edges),
                                                          _addr(nums_of_nodes), _addr(nums_of_edges), _addr(node_feature_u8),
                                                          _addr(edge_list_u16), _addr(edge_attr_u8), _addr(node_eigen)), "upload_batch_packed")
        self._num_graphs = int(num_graphs)

    def compute(self, model: str, timed: bool = True) -> float:
        ms = ctypes.c_float(0.0)
        _check(self._lib.flowgnn_b200_compute(self._h, MODEL_IDS[get_model(model).name], ctypes.byref(ms) if timed else None),
               "compute")
        return float(ms.value)

    def download(self, out=None) -> np.ndarray:
        if out is None:
            out = np.empty(self._num_graphs, dtype=np.float32)
        _check(self._lib.flowgnn_b200_download(self._h, _addr(out), self._num_graphs), "download")
        return out

    @property
    def tile_count(self) -> int:
        """128-row tiles of whole graphs the uploaded batch was packed into (0: not packed on the host)."""
        return int(self._lib.flowgnn_b200_tile_count(self._h))

    def synchronize(self) -> None:
        _check(self._lib.flowgnn_b200_synchronize(self._h), "synchronize")

    @property
    def last_launch_count(self) -> int:
        return int(self._lib.flowgnn_b200_last_launch_count(self._h))

    def last_layer_ms(self):
        """Device time of each per-layer launch of the last compute (needs set_option("time_layers", 1))."""
        buf = (ctypes.c_float * 16)()
        n = self._lib.flowgnn_b200_last_layer_ms(self._h, buf, 16)
        return [float(buf[i]) for i in range(n)]

    @property
    def stream(self) -> int:
        return int(self._lib.flowgnn_b200_stream(self._h) or 0)

    def run(self, model: str, batch: Batch, weights: Optional[Weights] = None) -> np.ndarray:
        """load (optional) + upload + compute + download."""
        if weights is not None:
            self.load_weights(model, weights)
        if get_model(model).virtual_node:
            batch = batch.with_virtual_node()
        self.upload(batch)
        self.compute(model, timed=False)
        return self.download()


class ReferenceCall:
    """A prepared call of the reference-compatible entry point ``<MODEL>_compute_graphs`` (Part 1 of the header)
    with HOST arrays, marshalled once: ``run()`` is then exactly the C call a host written in C++ would make
    (GIN/src/host.cc:184-209), with no per-call Python work."""

    def __init__(self, model: str, batch: Batch, weights: Weights, reload_weights: Optional[np.ndarray] = None,
                 weight_sets: Optional[Sequence[Weights]] = None):
        lib = load_library()
        spec: ModelSpec = get_model(model)
        sets = list(weight_sets) if weight_sets is not None else [weights]
        sets = [check_weights(spec, w) for w in sets]
        stacked: Dict[str, np.ndarray] = {n: np.ascontiguousarray(np.stack([w[n] for w in sets])) for n in spec.weight_names}
        G = batch.num_graphs
        if reload_weights is None:
            reload_weights = np.zeros(G, dtype=np.int32)
            if G:
                reload_weights[0] = 1
        reload_weights = np.ascontiguousarray(reload_weights, dtype=np.int32)
        self.out = np.zeros(G, dtype=np.float32)
        nn = np.ascontiguousarray(batch.nums_of_nodes, dtype=np.int32)
        ne = np.ascontiguousarray(batch.nums_of_edges, dtype=np.int32)
        args = [ctypes.c_int(G), nn.ctypes.data_as(_i32p), ne.ctypes.data_as(_i32p), reload_weights.ctypes.data_as(_i32p),
                self.out.ctypes.data_as(_f32p), batch.node_feature.ctypes.data_as(_i32p)]
        if spec.uses_eigen:
            if batch.node_eigen is None:
                raise FlowGNNError("DGN needs node_eigen")
            args.append(batch.node_eigen.ctypes.data_as(_f32p))
        args.append(batch.edge_list.ctypes.data_as(_i32p))
        if spec.uses_edge_attr:
            if batch.edge_attr is None:
                raise FlowGNNError(f"{spec.name} needs edge_attr")
            args.append(batch.edge_attr.ctypes.data_as(_i32p))
        for n in spec.weight_names:
            args.append(stacked[n].ctypes.data_as(_f32p))
        self._keep = (nn, ne, reload_weights, stacked, batch)
        self._args = args
        self._symbol = spec.symbol
        self._fn = getattr(lib, spec.symbol)
        self._fn.restype = ctypes.c_int

    def run(self) -> np.ndarray:
        _check(self._fn(*self._args), self._symbol)
        return self.out


class PackedCall:
    """A prepared call of ``flowgnn_b200_compute_graphs_packed``: the host-pointer pipeline of the entry points for a caller whose
    dataset is in the packed layout (uint8 features, uint16 edge ids, uint8 bond attributes).  Arrays may be numpy arrays or views of
    pinned torch tensors; ``run()`` is the bare C call."""

    def __init__(self, model: str, batch: Batch, weights: Weights, node_feature_u8=None, edge_list_u16=None, edge_attr_u8=None):
        lib = load_library()
        spec: ModelSpec = get_model(model)
        w = check_weights(spec, weights)
        arrs = [w[n] for n in spec.weight_names]
        self._wptrs = (_f32p * len(arrs))(*[a.ctypes.data_as(_f32p) for a in arrs])
        nf = batch.node_feature.astype(np.uint8) if node_feature_u8 is None else node_feature_u8
        el = batch.edge_list.astype(np.uint16) if edge_list_u16 is None else edge_list_u16
        ea = None
        if spec.uses_edge_attr:
            ea = batch.edge_attr.astype(np.uint8) if edge_attr_u8 is None else edge_attr_u8
        for a, dt in ((nf, np.uint8), (el, np.uint16), (ea, np.uint8)):
            if a is not None and a.dtype != dt:
                raise FlowGNNError(f"packed call wants {np.dtype(dt).name}, got {a.dtype}")
        eg = batch.node_eigen if spec.uses_eigen else None
        if spec.uses_eigen and eg is None:
            raise FlowGNNError("DGN needs node_eigen")
        self.out = np.zeros(batch.num_graphs, dtype=np.float32)
        nn = np.ascontiguousarray(batch.nums_of_nodes, dtype=np.int32)
        ne = np.ascontiguousarray(batch.nums_of_edges, dtype=np.int32)
        self._keep = (nn, ne, nf, el, ea, eg, arrs, batch)
        self._args = [ctypes.c_int(MODEL_IDS[spec.name]), ctypes.c_int(batch.num_graphs), ctypes.c_void_p(_addr(nn)), ctypes.c_void_p(_addr(ne)),
                      ctypes.c_void_p(_addr(self.out)), ctypes.c_void_p(_addr(nf)), ctypes.c_void_p(_addr(el)), ctypes.c_void_p(_addr(ea)),
                      ctypes.c_void_p(_addr(eg)), self._wptrs, ctypes.c_int(len(arrs))]
        self._fn = lib.flowgnn_b200_compute_graphs_packed
        self._fn.restype = ctypes.c_int

    def run(self) -> np.ndarray:
        _check(self._fn(*self._args), "flowgnn_b200_compute_graphs_packed")
        return self.out


def to_fixed(x, frac_bits: int = 10) -> np.ndarray:
    """``(WT_TYPE)float`` of the reference's host (GIN/src/host_load.cc:60-97): ap_fixed<16, 16 - frac_bits> bit patterns,
    floor(x * 2^F) (AP_TRN), low 16 bits (AP_WRAP)."""
    q = np.floor(np.asarray(x, dtype=np.float64) * float(1 << frac_bits)).astype(np.int64)
    return (q & 0xFFFF).astype(np.uint16).view(np.int16)


def compute_graphs_fixed(model: str, batch: Batch, weights: Weights) -> np.ndarray:
    """Call ``GIN_compute_graphs_fixed``: the kernel ABI of the FPGA build, int16 ap_fixed<16,6> bit patterns in and out
    (GIN / GIN-VN only).  ``weights`` are the fp32 arrays of the weight files, cast here as the reference's host casts them.
    Returns the raw int16 predictions (value = raw / 1024)."""
    lib = load_library()
    spec: ModelSpec = get_model(model)
    if spec.name not in ("gin", "ginvn"):
        raise FlowGNNError("the ap_fixed entry point exists for GIN / GIN-VN")
    w = check_weights(spec, weights)
    G = batch.num_graphs
    reload_weights = np.zeros(G, dtype=np.int32)
    if G:
        reload_weights[0] = 1
    out = np.zeros(G, dtype=np.int16)
    nn = np.ascontiguousarray(batch.nums_of_nodes, dtype=np.int32)
    ne = np.ascontiguousarray(batch.nums_of_edges, dtype=np.int32)
    i16p = ctypes.POINTER(ctypes.c_int16)
    fixed = [np.ascontiguousarray(to_fixed(w[n])) for n in spec.weight_names]
    args = [ctypes.c_int(G), nn.ctypes.data_as(_i32p), ne.ctypes.data_as(_i32p), reload_weights.ctypes.data_as(_i32p),
            out.ctypes.data_as(i16p), batch.node_feature.ctypes.data_as(_i32p), batch.edge_list.ctypes.data_as(_i32p),
            batch.edge_attr.ctypes.data_as(_i32p)] + [a.ctypes.data_as(i16p) for a in fixed]
    fn = lib.GIN_compute_graphs_fixed
    fn.restype = ctypes.c_int
    _check(fn(*args), "GIN_compute_graphs_fixed")
    return out


def compute_graphs(model: str, batch: Batch, weights: Weights, reload_weights: Optional[np.ndarray] = None,
                   weight_sets: Optional[Sequence[Weights]] = None) -> np.ndarray:
    """Call the reference-compatible entry point ``<MODEL>_compute_graphs`` (Part 1 of the header)
    with host arrays, exactly as the reference's host would (GIN/src/host.cc:184-209).

    ``weight_sets`` (optional) stacks several weight sets along the leading dimension; ``reload_weights``
    then marks the graphs at which the kernel advances to the next set (GIN/src/GIN_compute.cc:49-63).
    GIN-VN: the caller passes the virtual-node-augmented batch (as the reference's host does)."""
    return ReferenceCall(model, batch, weights, reload_weights, weight_sets).run()

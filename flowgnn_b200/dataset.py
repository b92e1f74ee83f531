"""Graph batches: the host-side mirror of the reference's graph loading.

A :class:`Batch` is exactly the set of buffers the reference's ``host.cc`` builds before it
enqueues the kernel (GIN/src/host.cc:110-182): per-graph node/edge counts plus the
concatenation over graphs of ``node_feature`` [sum N][9], ``edge_list`` [sum E][2] (graph-local
ids, u = source, v = destination), ``edge_attr`` [sum E][3] and, for DGN, ``node_eigen``
[sum N][4] (DGN/src/host.cc, DGN/src/host_load.cc:201-215).

Readers: the reference's per-graph files (``graphs/graph_info/g%d_info.txt``,
``graphs/graph_bin/g%d_{node_feature,edge_list,edge_attr}.bin``, ``DGN/eig/g%d.txt``), straight
from a dataset zip or an extracted directory; and a packed single-file format (``.fgb``/``.npz``)
that replaces the 3-4 tiny files per graph.

Also here: the GIN-VN virtual-node augmentation (GIN-VN/src/host_load.cc:125-153 and
GIN-VN/src/host.cc:133-134) and the synthetic molhiv-/molpcba-/hep10k-shaped generators used
for throughput runs.
"""
from __future__ import annotations

import io
import os
import re
import struct
import zipfile
from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence

import numpy as np

from .models import ED_FEATURE_TABLE, EDGE_ATTR, ND_FEATURE, ND_FEATURE_TABLE

PACK_MAGIC = b"FGNNPACK"
PACK_VERSION = 1
_FLAG_EDGE_ATTR = 1
_FLAG_EIGEN = 2


@dataclass
class Batch:
    nums_of_nodes: np.ndarray                  # int32 [G]
    nums_of_edges: np.ndarray                  # int32 [G]
    node_feature: np.ndarray                   # int32 [sum N, 9]
    edge_list: np.ndarray                      # int32 [sum E, 2]  (u, v) graph-local
    edge_attr: Optional[np.ndarray] = None     # int32 [sum E, 3]
    node_eigen: Optional[np.ndarray] = None    # float32 [sum N, 4]
    name: str = ""
    _node_off: Optional[np.ndarray] = field(default=None, repr=False)
    _edge_off: Optional[np.ndarray] = field(default=None, repr=False)

    def __post_init__(self):
        self.nums_of_nodes = np.ascontiguousarray(self.nums_of_nodes, dtype=np.int32)
        self.nums_of_edges = np.ascontiguousarray(self.nums_of_edges, dtype=np.int32)
        self.node_feature = np.ascontiguousarray(self.node_feature, dtype=np.int32).reshape(-1, ND_FEATURE)
        self.edge_list = np.ascontiguousarray(self.edge_list, dtype=np.int32).reshape(-1, 2)
        if self.edge_attr is not None:
            self.edge_attr = np.ascontiguousarray(self.edge_attr, dtype=np.int32).reshape(-1, EDGE_ATTR)
        if self.node_eigen is not None:
            self.node_eigen = np.ascontiguousarray(self.node_eigen, dtype=np.float32).reshape(-1, 4)
        self.validate()

    # ---- shape bookkeeping -------------------------------------------------------------
    @property
    def num_graphs(self) -> int:
        return int(self.nums_of_nodes.shape[0])

    @property
    def total_nodes(self) -> int:
        return int(self.node_feature.shape[0])

    @property
    def total_edges(self) -> int:
        return int(self.edge_list.shape[0])

    @property
    def node_offsets(self) -> np.ndarray:
        if self._node_off is None:
            self._node_off = np.concatenate([[0], np.cumsum(self.nums_of_nodes, dtype=np.int64)])
        return self._node_off

    @property
    def edge_offsets(self) -> np.ndarray:
        if self._edge_off is None:
            self._edge_off = np.concatenate([[0], np.cumsum(self.nums_of_edges, dtype=np.int64)])
        return self._edge_off

    def validate(self) -> None:
        if self.nums_of_nodes.shape != self.nums_of_edges.shape:
            raise ValueError("nums_of_nodes / nums_of_edges length mismatch")
        if int(self.nums_of_nodes.sum(dtype=np.int64)) != self.total_nodes:
            raise ValueError("sum(nums_of_nodes) != rows of node_feature")
        if int(self.nums_of_edges.sum(dtype=np.int64)) != self.total_edges:
            raise ValueError("sum(nums_of_edges) != rows of edge_list")
        if self.edge_attr is not None and self.edge_attr.shape[0] != self.total_edges:
            raise ValueError("edge_attr rows != edge_list rows")
        if self.node_eigen is not None and self.node_eigen.shape[0] != self.total_nodes:
            raise ValueError("node_eigen rows != node_feature rows")

    # ---- slicing / sharding ------------------------------------------------------------
    def slice(self, g0: int, g1: int) -> "Batch":
        """Graphs [g0, g1) as a new batch (views where possible)."""
        g0 = max(0, min(g0, self.num_graphs))
        g1 = max(g0, min(g1, self.num_graphs))
        n0, n1 = int(self.node_offsets[g0]), int(self.node_offsets[g1])
        e0, e1 = int(self.edge_offsets[g0]), int(self.edge_offsets[g1])
        return Batch(
            self.nums_of_nodes[g0:g1], self.nums_of_edges[g0:g1], self.node_feature[n0:n1], self.edge_list[e0:e1],
            None if self.edge_attr is None else self.edge_attr[e0:e1],
            None if self.node_eigen is None else self.node_eigen[n0:n1],
            name=self.name,
        )

    def select(self, graph_ids: Sequence[int]) -> "Batch":
        return concat([self.slice(int(g), int(g) + 1) for g in graph_ids], name=self.name)

    def tile(self, num_graphs: int) -> "Batch":
        """Repeat the batch cyclically until it holds ``num_graphs`` graphs ("real-tiled" batches)."""
        reps = -(-num_graphs // max(self.num_graphs, 1))
        out = concat([self] * reps, name=self.name)
        return out.slice(0, num_graphs)

    def with_virtual_node(self) -> "Batch":
        """GIN-VN augmentation: one extra all-zero-feature node N per graph and, after the real
        edges, the pairs (i, N), (N, i) for i = 0..N-1 with attr {0,0,0}
        (GIN-VN/src/host_load.cc:129,137-141,149-153; counts per GIN-VN/src/host.cc:133-134)."""
        G = self.num_graphs
        nn = self.nums_of_nodes.astype(np.int64)
        ne = self.nums_of_edges.astype(np.int64)
        new_nn = nn + 1
        new_ne = ne + 2 * nn
        node_off = self.node_offsets
        edge_off = self.edge_offsets
        new_node_off = np.concatenate([[0], np.cumsum(new_nn)])
        new_edge_off = np.concatenate([[0], np.cumsum(new_ne)])
        nf = np.zeros((int(new_node_off[-1]), ND_FEATURE), dtype=np.int32)
        el = np.zeros((int(new_edge_off[-1]), 2), dtype=np.int32)
        ea = np.zeros((int(new_edge_off[-1]), EDGE_ATTR), dtype=np.int32)
        # real nodes / edges keep their relative position inside each graph
        node_graph = np.repeat(np.arange(G), nn)
        nf[np.arange(self.total_nodes) + (new_node_off[:-1] - node_off[:-1])[node_graph]] = self.node_feature
        edge_graph = np.repeat(np.arange(G), ne)
        dst_rows = np.arange(self.total_edges) + (new_edge_off[:-1] - edge_off[:-1])[edge_graph]
        el[dst_rows] = self.edge_list
        if self.edge_attr is not None:
            ea[dst_rows] = self.edge_attr
        # virtual edges
        local = np.arange(self.total_nodes) - node_off[:-1][node_graph]          # i within graph
        base = (new_edge_off[:-1] + ne)[node_graph] + 2 * local
        vn = nn[node_graph].astype(np.int32)
        el[base, 0] = local
        el[base, 1] = vn
        el[base + 1, 0] = vn
        el[base + 1, 1] = local
        return Batch(new_nn, new_ne, nf, el, ea, None, name=self.name + "+vn")

    # ---- packed single-file format -----------------------------------------------------
    def save_packed(self, path: str) -> None:
        flags = (_FLAG_EDGE_ATTR if self.edge_attr is not None else 0) | (_FLAG_EIGEN if self.node_eigen is not None else 0)
        with open(path, "wb") as f:
            f.write(PACK_MAGIC)
            f.write(struct.pack("<IIQQQ", PACK_VERSION, flags, self.num_graphs, self.total_nodes, self.total_edges))
            for a in (self.nums_of_nodes, self.nums_of_edges, self.node_feature, self.edge_list):
                f.write(np.ascontiguousarray(a).tobytes())
            if self.edge_attr is not None:
                f.write(self.edge_attr.tobytes())
            if self.node_eigen is not None:
                f.write(self.node_eigen.tobytes())

    def save_reference_layout(self, root: str, first: int = 1) -> None:
        """Write the batch in the reference's per-graph file layout (what its host reads, GIN/src/host.cc:119-138 and
        DGN/src/host_load.cc:201-215): graphs/graph_info/g%d_info.txt, graphs/graph_bin/g%d_{node_feature,edge_list,
        edge_attr}.bin, DGN/eig/g%d.txt (a printed tensor) and common/includes/dataset/dataset_size.txt."""
        for d in ("graphs/graph_info", "graphs/graph_bin", "DGN/eig", "common/includes/dataset"):
            os.makedirs(os.path.join(root, d), exist_ok=True)
        no, eo = self.node_offsets, self.edge_offsets
        for i in range(self.num_graphs):
            g = first + i
            n0, n1, e0, e1 = int(no[i]), int(no[i + 1]), int(eo[i]), int(eo[i + 1])
            with open(os.path.join(root, f"graphs/graph_info/g{g}_info.txt"), "w") as f:
                f.write(f"{n1 - n0}\n{e1 - e0}")
            self.node_feature[n0:n1].astype("<i4").tofile(os.path.join(root, f"graphs/graph_bin/g{g}_node_feature.bin"))
            self.edge_list[e0:e1].astype("<i4").tofile(os.path.join(root, f"graphs/graph_bin/g{g}_edge_list.bin"))
            attr = self.edge_attr[e0:e1] if self.edge_attr is not None else np.zeros((e1 - e0, EDGE_ATTR), np.int32)
            attr.astype("<i4").tofile(os.path.join(root, f"graphs/graph_bin/g{g}_edge_attr.bin"))
            if self.node_eigen is not None:
                rows = ",\n        ".join("[" + ", ".join(repr(float(x)) for x in r) + "]" for r in self.node_eigen[n0:n1])
                with open(os.path.join(root, f"DGN/eig/g{g}.txt"), "w") as f:
                    f.write("tensor([" + rows + "])")
        with open(os.path.join(root, "common/includes/dataset/dataset_size.txt"), "w") as f:
            f.write(str(first + self.num_graphs - 1))

    def save_npz(self, path: str) -> None:
        arrays = dict(nums_of_nodes=self.nums_of_nodes, nums_of_edges=self.nums_of_edges,
                      node_feature=self.node_feature.astype(np.uint8) if self.node_feature.max(initial=0) < 256 else self.node_feature,
                      edge_list=self.edge_list.astype(np.uint16) if self.edge_list.max(initial=0) < 65536 else self.edge_list)
        if self.edge_attr is not None:
            arrays["edge_attr"] = self.edge_attr.astype(np.uint8)
        if self.node_eigen is not None:
            arrays["node_eigen"] = self.node_eigen
        np.savez_compressed(path, **arrays)


def concat(batches: Sequence[Batch], name: str = "") -> Batch:
    if not batches:
        return Batch(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros((0, ND_FEATURE), np.int32), np.zeros((0, 2), np.int32),
                     np.zeros((0, EDGE_ATTR), np.int32), None, name=name)
    has_attr = all(b.edge_attr is not None for b in batches)
    has_eig = all(b.node_eigen is not None for b in batches)
    return Batch(
        np.concatenate([b.nums_of_nodes for b in batches]), np.concatenate([b.nums_of_edges for b in batches]),
        np.concatenate([b.node_feature for b in batches]), np.concatenate([b.edge_list for b in batches]),
        np.concatenate([b.edge_attr for b in batches]) if has_attr else None,
        np.concatenate([b.node_eigen for b in batches]) if has_eig else None,
        name=name or batches[0].name,
    )


def load_packed(path: str) -> Batch:
    with open(path, "rb") as f:
        if f.read(8) != PACK_MAGIC:
            raise ValueError(f"{path}: not a FGNNPACK file")
        version, flags, G, N, E = struct.unpack("<IIQQQ", f.read(32))
        if version != PACK_VERSION:
            raise ValueError(f"{path}: unsupported pack version {version}")
        rd = lambda dt, n: np.frombuffer(f.read(n * np.dtype(dt).itemsize), dtype=dt, count=n)
        nn = rd("<i4", G)
        ne = rd("<i4", G)
        nf = rd("<i4", N * ND_FEATURE)
        el = rd("<i4", E * 2)
        ea = rd("<i4", E * EDGE_ATTR) if flags & _FLAG_EDGE_ATTR else None
        eg = rd("<f4", N * 4) if flags & _FLAG_EIGEN else None
    return Batch(nn, ne, nf, el, ea, eg, name=os.path.basename(path))


def load_npz(path: str) -> Batch:
    z = np.load(path)
    return Batch(z["nums_of_nodes"], z["nums_of_edges"], z["node_feature"], z["edge_list"],
                 z["edge_attr"] if "edge_attr" in z.files else None,
                 z["node_eigen"] if "node_eigen" in z.files else None, name=os.path.basename(path))


# ---- the reference's per-graph file layout -------------------------------------------------

_EIG_FLOAT = re.compile(rb"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?|nan|inf")


def parse_eigen_text(text: bytes, num_nodes: int) -> np.ndarray:
    """``DGN/eig/g%d.txt`` is a printed PyTorch tensor, N rows x 4 (DGN/src/host_load.cc:201-215).
    The reference's fscanf walk is restated as: take the numbers in order, 4 per node."""
    vals = np.array([float(x) for x in _EIG_FLOAT.findall(text.replace(b"tensor", b""))], dtype=np.float32)
    if vals.size < 4 * num_nodes:
        raise ValueError(f"eigen file has {vals.size} numbers, need {4 * num_nodes}")
    return vals[:4 * num_nodes].reshape(num_nodes, 4)


def _load_graph_files(read: Callable[[str], bytes], num_graphs: int, first: int, with_eigen: bool, name: str) -> Batch:
    nn, ne, nf, el, ea, eg = [], [], [], [], [], []
    for g in range(first, first + num_graphs):
        info = read(f"graphs/graph_info/g{g}_info.txt").split()       # "N\nE", CRLF in molhiv
        n, e = int(info[0]), int(info[1])
        nn.append(n)
        ne.append(e)
        nf.append(np.frombuffer(read(f"graphs/graph_bin/g{g}_node_feature.bin"), dtype="<i4", count=n * ND_FEATURE))
        el.append(np.frombuffer(read(f"graphs/graph_bin/g{g}_edge_list.bin"), dtype="<i4", count=e * 2))
        ea.append(np.frombuffer(read(f"graphs/graph_bin/g{g}_edge_attr.bin"), dtype="<i4", count=e * EDGE_ATTR))
        if with_eigen:
            eg.append(parse_eigen_text(read(f"DGN/eig/g{g}.txt"), n))
    return Batch(np.asarray(nn), np.asarray(ne), np.concatenate(nf), np.concatenate(el), np.concatenate(ea),
                 np.concatenate(eg) if with_eigen else None, name=name)


def dataset_size(read: Callable[[str], bytes]) -> int:
    return int(read("common/includes/dataset/dataset_size.txt").split()[0])


def load_dataset_zip(path: str, num_graphs: Optional[int] = None, first: int = 1, with_eigen: bool = True) -> Batch:
    """Read graphs ``first .. first+num_graphs-1`` (1-based ids, as the reference numbers them)
    from one of the reference's dataset zips (molhiv.zip / molpcba.zip / hep10k.zip)."""
    with zipfile.ZipFile(path) as zf:
        read = zf.read
        total = dataset_size(read)
        if num_graphs is None:
            num_graphs = total - first + 1
        return _load_graph_files(read, num_graphs, first, with_eigen, os.path.basename(path))


def load_dataset_dir(root: str, num_graphs: Optional[int] = None, first: int = 1, with_eigen: bool = False) -> Batch:
    """Same, from an extracted tree (``<root>/graphs/graph_info``, ``<root>/graphs/graph_bin``, ``<root>/DGN/eig``)."""
    def read(rel: str) -> bytes:
        with open(os.path.join(root, rel), "rb") as f:
            return f.read()
    if num_graphs is None:
        num_graphs = dataset_size(read) - first + 1
    return _load_graph_files(read, num_graphs, first, with_eigen, os.path.basename(root.rstrip("/")))


# ---- synthetic generators ------------------------------------------------

BASE_SEED = 20220427


def _mol_graph(rng: np.random.Generator, n: int, lam: float):
    deg = np.zeros(n, dtype=np.int64)
    bonds = []
    adj = set()
    for i in range(1, n):
        cand = np.flatnonzero(deg[:i] < 4)
        j = int(cand[rng.integers(cand.size)]) if cand.size else int(rng.integers(i))
        bonds.append((i, j))
        adj.add((min(i, j), max(i, j)))
        deg[i] += 1
        deg[j] += 1
    for _ in range(int(rng.poisson(lam))):
        cand = np.flatnonzero(deg < 4)
        if cand.size < 2:
            break
        a, b = (int(x) for x in rng.choice(cand, size=2, replace=False))
        key = (min(a, b), max(a, b))
        if key in adj:
            continue
        adj.add(key)
        bonds.append((a, b))
        deg[a] += 1
        deg[b] += 1
    return bonds


def synthetic_molecules(num_graphs: int, shape: str = "molhiv", seed: int = BASE_SEED, with_eigen: bool = False) -> Batch:
    """molhiv-/molpcba-shaped random molecules: lognormal node count, spanning tree with max
    valence 4 plus Poisson ring closures, both edge directions adjacent, no isolated atoms."""
    mu, sigma, nmin, nmax, lam = {
        "molhiv": (np.log(23.0), 0.42, 6, 183, 3.5),
        "molpcba": (np.log(26.0), 0.27, 4, 188, 3.7),
    }[shape]
    rng = np.random.default_rng([seed, 1])
    ns = np.clip(np.rint(np.exp(rng.normal(mu, sigma, size=num_graphs))), nmin, nmax).astype(np.int64)
    nn, ne, nf, el, ea, eg = [], [], [], [], [], []
    for g in range(num_graphs):
        n = int(ns[g])
        grng = np.random.default_rng([seed, 2, g])
        bonds = _mol_graph(grng, n, lam)
        e = np.empty((2 * len(bonds), 2), dtype=np.int32)
        e[0::2] = bonds
        e[1::2] = [(b, a) for a, b in bonds]
        attr = np.stack([grng.integers(0, v, size=len(bonds)) for v in ED_FEATURE_TABLE], axis=1).astype(np.int32)
        feat = np.stack([np.minimum(grng.geometric(0.35, size=n) + 4, 118)] +
                        [grng.integers(0, v, size=n) for v in ND_FEATURE_TABLE[1:]], axis=1).astype(np.int32)
        nn.append(n)
        ne.append(e.shape[0])
        nf.append(feat)
        el.append(e)
        ea.append(np.repeat(attr, 2, axis=0))
        if with_eigen:
            v = grng.normal(size=(n, 4)).astype(np.float32)
            eg.append(v / np.linalg.norm(v, axis=0, keepdims=True))
    return Batch(np.asarray(nn), np.asarray(ne), np.concatenate(nf), np.concatenate(el), np.concatenate(ea),
                 np.concatenate(eg) if with_eigen else None, name=f"synthetic-{shape}")


def synthetic_hep(num_graphs: int, seed: int = BASE_SEED) -> Batch:
    """hep10k-shaped graphs: N ~ N(49.1, 17) clamped to [5,123], directed kNN (k = min(16, N-1))
    over random 3-D points, sorted by source then distance, all features/attrs zero."""
    rng = np.random.default_rng([seed, 3])
    ns = np.clip(np.rint(rng.normal(49.1, 17.0, size=num_graphs)), 5, 123).astype(np.int64)
    nn, ne, el = [], [], []
    for g in range(num_graphs):
        n = int(ns[g])
        k = min(16, n - 1)
        pts = np.random.default_rng([seed, 4, g]).random((n, 3))
        d = ((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1)
        np.fill_diagonal(d, np.inf)
        nbr = np.argsort(d, axis=1, kind="stable")[:, :k]
        e = np.stack([np.repeat(np.arange(n), k), nbr.reshape(-1)], axis=1).astype(np.int32)
        nn.append(n)
        ne.append(e.shape[0])
        el.append(e)
    tn, te = int(np.sum(nn)), int(np.sum(ne))
    return Batch(np.asarray(nn), np.asarray(ne), np.zeros((tn, ND_FEATURE), np.int32), np.concatenate(el),
                 np.zeros((te, EDGE_ATTR), np.int32), None, name="synthetic-hep10k")


def shard_ranges(batch: Batch, world_size: int, node_cost: float = 1.0, edge_cost: float = 0.25):
    """Contiguous graph-index ranges balanced by sum(c_n*N + c_e*E).  Returns
    ``world_size + 1`` boundaries; rank k owns graphs [b[k], b[k+1])."""
    cost = node_cost * batch.nums_of_nodes.astype(np.float64) + edge_cost * batch.nums_of_edges.astype(np.float64)
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    targets = cum[-1] * np.arange(1, world_size) / world_size
    inner = np.searchsorted(cum, targets, side="left")
    bounds = np.concatenate([[0], inner, [batch.num_graphs]]).astype(np.int64)
    return np.maximum.accumulate(bounds)

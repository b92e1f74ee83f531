"""compute-sanitizer workload: every DEFAULT kernel of the six models on small batches (run under
`compute-sanitizer --tool memcheck|racecheck|synccheck`, see tools/sanitize.sh).

Covers: scan + both CSR-build kernels (small graphs, a 300-node graph, a 1,500-node graph on the global-memory tables),
tile packing and row sorting, the embedding kernels, the GIN fused layer kernel (gin_fused.cu: fused head and unfused), the staged
gather + node-MLP launches of dense graphs, the mp_only mode, the fused GCN / DGN step kernels (fused_tc.cuh) + gcn_final +
dgn_exact_rows, the fused PNA layer kernel (pna_fused.cu) + pna_exact_rows, GAT, the pooling / head kernels, the ap_fixed
kernels of GIN and DGN (option fixed_point), and the chunked host-pointer entry point (two streams, two device batches)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from flowgnn_b200.capi import Context, ReferenceCall  # noqa: E402
from flowgnn_b200.dataset import Batch, concat, load_npz, synthetic_molecules  # noqa: E402
from flowgnn_b200.weights import load_weights  # noqa: E402

gold = os.path.join(ROOT, "tests", "golden")
DIRS = {"gin": "GIN", "ginvn": "GIN", "gcn": "GCN", "gat": "GAT", "pna": "PNA", "dgn": "DGN"}
n_mol = int(os.environ.get("SAN_MOL", "150"))
n_hep = int(os.environ.get("SAN_HEP", "12"))
mol = synthetic_molecules(n_mol, "molhiv", seed=3, with_eigen=True)
hep = load_npz(os.path.join(gold, "hep10k.npz")).slice(0, n_hep)
rng = np.random.default_rng(1)


def chain(n):
    a = np.concatenate([np.arange(n - 1), rng.integers(0, n, n)])
    b = np.concatenate([np.arange(1, n), rng.integers(0, n, n)])
    e = np.stack([np.stack([a, b], 1), np.stack([b, a], 1)], 1).reshape(-1, 2).astype(np.int32)
    return Batch(np.array([n]), np.array([len(e)]), np.zeros((n, 9), np.int32), e, np.zeros((len(e), 3), np.int32),
                 rng.standard_normal((n, 4)).astype(np.float32))


mixed = concat([mol.slice(0, 20), chain(300), chain(1500), mol.slice(20, 40)])
with Context(0) as c:
    for model in ("gin", "ginvn", "gcn", "gat", "pna", "dgn"):
        w = load_weights(model, os.path.join(gold, "weights", DIRS[model]))
        y = c.run(model, mol, w)
        y2 = c.run(model, hep)
        y3 = c.run(model, mixed)
        print(model, "molecules", float(np.nanmax(np.abs(y))), "hep10k", float(np.nanmax(np.abs(y2))), "mixed", float(np.nanmax(np.abs(y3))), flush=True)
    w = load_weights("gin", os.path.join(gold, "weights", "GIN"))
    c.load_weights("gin", w)
    for opt, val in (("gin_unfused_head", 1), ("mp_only", 1), ("gin_staged", 1)):
        c.set_option(opt, val)
        y = c.run("gin", mol)
        c.set_option(opt, 0 if opt != "gin_staged" else -1)
        print("gin", opt, float(np.abs(y).max()), flush=True)
    for model in ("gin", "dgn"):
        c.set_option("fixed_point", 1)
        y = c.run(model, mixed, load_weights(model, os.path.join(gold, "weights", DIRS[model])))
        c.set_option("fixed_point", 0)
        print(model, "fixed_point", float(np.abs(y).max()), flush=True)
# the chunked host-pointer entry point (>= 8,192 graphs -> 2 chunks on two streams)
big = mol.tile(8192 + 64)
call = ReferenceCall("gin", big, w)
y = call.run()
print("entry point", y.shape, float(np.abs(y).max()), flush=True)
# the same with the plain int32 copies, and Part 2's packed upload (unpack_inputs_kernel on odd sizes)
os.environ["FLOWGNN_B200_HOST_STAGE"] = "0"
y0 = call.run().copy()
del os.environ["FLOWGNN_B200_HOST_STAGE"]
assert np.array_equal(y0.view(np.int32), call.run().view(np.int32))
with Context(0) as c:
    c.load_weights("gin", w)
    odd = mol.slice(0, 77)
    c.upload_packed_arrays(odd.num_graphs, odd.total_nodes, odd.total_edges, odd.nums_of_nodes, odd.nums_of_edges, odd.node_feature.astype(np.uint8),
                           odd.edge_list.astype(np.uint16), odd.edge_attr.astype(np.uint8))
    c.compute("gin")
    print("packed upload", float(np.abs(c.download()).max()), flush=True)
print("sanitize workload done", flush=True)

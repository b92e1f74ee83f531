"""End-to-end (host buffers) timing of GIN_compute_graphs for several chunk schedules.  usage: python tools/e2e_probe.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from flowgnn_b200.capi import ReferenceCall
from flowgnn_b200.dataset import synthetic_molecules
from flowgnn_b200.weights import load_weights
w = load_weights("gin", os.path.join(ROOT, "tests", "golden", "weights", "GIN"))
big = synthetic_molecules(2048, "molhiv", seed=11).tile(41127)
import torch
from flowgnn_b200.dataset import Batch
keep = []
def pin(a):
    t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, pin_memory=True)
    v = t.numpy(); v[...] = a; keep.append(t)
    return v
big = Batch(pin(big.nums_of_nodes), pin(big.nums_of_edges), pin(big.node_feature), pin(big.edge_list), pin(big.edge_attr), None, name=big.name)
call = ReferenceCall("gin", big, w)
for n in (os.environ.get("SCHEDULES", "4 1 2 3 4 6 8 12").split()):
    os.environ["FLOWGNN_B200_CHUNKS"] = n
    for _ in range(3): call.run()
    t0 = time.perf_counter()
    for _ in range(30): call.run()
    print(f"chunks={n}: {(time.perf_counter() - t0) / 30 * 1e3:.3f} ms per call", flush=True)
